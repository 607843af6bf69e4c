"""`B200ShardInferenceEngine` — the drop-in replacement for `TorchDynamicShardInferenceEngine`
(/root/reference/xotorch/inference/torch/sharded_inference_engine.py:37-424) on B200.

Same public surface (`encode / sample / decode / infer_tensor / ensure_shard / load_checkpoint`, attributes `.shard`,
`.tokenizer`, `.shard_downloader`), same ndim protocol (2 => int token ids [1,S], 3 => float32 hidden [1,S,D]; returns
float32 hidden [1,S,D] or last-position logits [1,V]), same 1-thread executor discipline — but every FLOP runs in
libxotb200.so through the C ABI (xotorch_b200/native.py).  There is no CPU fallback.

Deliberate deviations from reference DEFECTS: the inference_state that rides the ring is the tiny
dict {"curr_pos": int} instead of the JSON-encoded L x L mask (B5/B11, §0.9); a shard hands off the output of its
last layer (B1); KV capacity comes from the request (B7); strict weight loading (B6)."""
from __future__ import annotations

import asyncio
import functools
import os
import uuid
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Callable, Dict, Optional, Tuple

import numpy as np

from .inference_engine import InferenceEngine
from .shard import Shard
from .tokenizers import resolve_local_tokenizer

TEMP = 0.6      # sharded_inference_engine.py:34
TOP_K = 35      # sharded_inference_engine.py:35
MAX_GENERATED_TOKENS = 1024   # general_mha.py:152: KV length = prompt + 1024

# Device-resident hop (xotb200_forward_hop): when the next shard's engine lives in this process (NVLinkPeerHandle), the
# "tensor" a non-last shard returns — and PeerHandle.send_tensor carries — is this 4-word descriptor instead of the fp32
# hidden state; the payload was stored into the next engine's HBM by the producing kernel (SURVEY.md §8b "Hop API").
HOP_MAGIC = 0x584F5442   # "XOTB"


def make_hop_descriptor(seq: int, n_tokens: int, dim: int) -> np.ndarray:
  return np.array([[[HOP_MAGIC, int(seq), int(n_tokens), int(dim)]]], dtype=np.int64)


def parse_hop_descriptor(x: np.ndarray):
  """(seq, n_tokens, dim) if `x` is a hop descriptor, else None.  ndim 3 like a hidden state, but int64 [1, 1, 4] + magic."""
  if isinstance(x, np.ndarray) and x.dtype == np.int64 and x.shape == (1, 1, 4) and int(x[0, 0, 0]) == HOP_MAGIC:
    return int(x[0, 0, 1]), int(x[0, 0, 2]), int(x[0, 0, 3])
  return None


def reference_sample(logits: np.ndarray, temp: float, top_k: int, q: np.ndarray) -> np.ndarray:
  """torchtune.generation.sample restated on host arrays: /max(temp,1e-5), top-k mask,
  softmax, argmax(probs / q).  `q` ~ Exp(1) noise.  Returns int32 [B,1]."""
  x = logits.astype(np.float32) / np.float32(max(temp, 1e-5))
  if top_k is not None:
    k = min(int(top_k), x.shape[-1])
    pivot = np.partition(x, -k, axis=-1)[..., -k][..., None]
    x = np.where(x < pivot, -np.inf, x)
  x = x - x.max(axis=-1, keepdims=True)
  p = np.exp(x)
  p = p / p.sum(axis=-1, keepdims=True)
  return np.argmax(p / q, axis=-1).astype(np.int32).reshape(-1, 1)


class LocalDirDownloader:
  """Stands in for the reference's ShardDownloader (xotorch/download/shard_download.py:9-37): `ensure_shard` must
  return a local directory.  There is no network on the B200 box, so model ids map to local directories."""

  def __init__(self, model_dirs: Optional[Dict[str, os.PathLike]] = None):
    self.model_dirs = dict(model_dirs or {})

  async def ensure_shard(self, shard, inference_engine_name: str) -> Path:
    mid = shard.model_id
    if mid in self.model_dirs:
      return Path(self.model_dirs[mid])
    if Path(mid).is_dir():
      return Path(mid)
    env = os.environ.get("XOT_B200_MODEL_DIR")
    if env and (Path(env) / mid).is_dir():
      return Path(env) / mid
    raise FileNotFoundError(f"no local model directory for model id {mid!r} (set model_dirs or XOT_B200_MODEL_DIR)")


class B200ShardInferenceEngine(InferenceEngine):
  def __init__(self, shard_downloader=None, device: Optional[int] = None,
               native_factory: Optional[Callable] = None, max_concurrent_requests: int = 1,
               max_total_len: Optional[int] = None):
    """max_concurrent_requests > 1: the engine keeps that many KV caches and serves interleaved `request_id`s (the
    reference's Node tracks requests individually, but its engine is single-request: `self.request_id` latched at
    sharded_inference_engine.py:250, one KV cache, so two overlapping requests corrupt each other).  `max_total_len`
    then fixes the KV capacity per request (default: first prompt + 1024, like the reference's prompt + max_generated)."""
    self.shard: Optional[Shard] = None
    self.shard_downloader = shard_downloader if shard_downloader is not None else LocalDirDownloader()
    self.tokenizer = None
    self.native = None
    self.model_path: Optional[Path] = None
    self.model_config: Optional[dict] = None
    self.request_id: Optional[str] = None
    self.executor = ThreadPoolExecutor(max_workers=1)       # sharded_inference_engine.py:46
    self.uuid = str(uuid.uuid4())
    if device is None:
      dev = os.environ.get("TORCH_DEVICE", "cuda:0")        # same env the reference honours (:58-59)
      device = int(dev.split(":")[1]) if ":" in dev else 0
    self.device = device
    self._native_factory = native_factory                   # tests inject a stand-in; default = the CUDA library
    self._kv_capacity = 0
    self._pos = 0
    self.max_concurrent_requests = int(max_concurrent_requests)
    self.max_total_len = max_total_len
    self._slots: Dict[str, int] = {}                        # request_id -> KV slot, in least-recently-used order
    self._hop_next: Optional["B200ShardInferenceEngine"] = None   # engine of the next shard when it lives in this process
    self._hop_linked = None                                 # (own native, its native) the device link was made for
    self.hop_stats = {"device_out": 0, "device_in": 0, "host_out": 0, "host_in": 0}
    import torch
    self._rng = torch.Generator(device="cpu")
    self._rng.manual_seed(1234)                             # sharded_inference_engine.py:67-69

  # ---- helpers -------------------------------------------------------------------------------------------------
  async def _run(self, fn, *a):
    return await asyncio.get_running_loop().run_in_executor(self.executor, functools.partial(fn, *a))

  def _make_native(self, hf_config: dict):
    if self._native_factory is not None:
      return self._native_factory(hf_config, self.device)
    from xotorch_b200.native import NativeShard             # raises ImportError if libxotb200.so is missing
    return NativeShard(hf_config, self.device)

  # ---- device-resident hop ---------------------------------------------------------------------------------------
  def attach_next_engine(self, nxt: Optional["B200ShardInferenceEngine"]) -> None:
    """Called by NVLinkPeerHandle.connect(): the engine that runs the NEXT shard of the ring lives in this process (another GPU
    of the box, or this one).  From then on a non-last shard hands its hidden state over on the device."""
    self._hop_next = nxt
    self._hop_linked = None

  def _hop_ready(self) -> bool:
    """Link this engine's native object to the next engine's (lazily: both load their shard on first use, and a changed
    shard means a new native object).  Runs on this engine's executor thread."""
    nxt = self._hop_next
    if nxt is None or self.native is None or nxt.native is None or not hasattr(self.native, "forward_hop"):
      return False
    if self._hop_linked != (id(self.native), id(nxt.native)):
      self.native.link_next_local(nxt.native)               # enables peer access, clears the link's flags
      self._hop_linked = (id(self.native), id(nxt.native))
    return True

  # ---- reference API ---------------------------------------------------------------------------------------------
  async def ensure_shard(self, shard: Shard):
    """sharded_inference_engine.py:372-421: no-op for the same shard, else (re)load config, tokenizer, weights."""
    if self.shard == shard:
      return
    from xotorch_b200.native import load_hf_config
    self.model_path = Path(await self.shard_downloader.ensure_shard(shard, self.__class__.__name__))
    self.model_config = load_hf_config(self.model_path)
    if self.model_config["num_hidden_layers"] != shard.n_layers:
      raise ValueError(f"shard.n_layers={shard.n_layers} but the model has {self.model_config['num_hidden_layers']} layers")
    self.tokenizer = resolve_local_tokenizer(self.model_path, self.model_config["vocab_size"])

    def start_model():
      if self.native is not None:
        self.native.close()
      self.native = self._make_native(self.model_config)
      self.native.load_dir(self.model_path, shard.start_layer, shard.end_layer)   # strict loader
      self._kv_capacity = 0
    await self._run(start_model)
    self.shard = shard
    self._pos = 0

  async def encode(self, shard: Shard, prompt: str) -> np.ndarray:
    await self.ensure_shard(shard)
    def enc():
      ids = self.tokenizer.encode(prompt)
      return np.asarray(ids, dtype=np.int64).reshape(-1)
    return await self._run(enc)

  async def decode(self, shard: Shard, tokens: np.ndarray) -> str:
    await self.ensure_shard(shard)
    return await self._run(self.tokenizer.decode, np.asarray(tokens).reshape(-1).tolist())

  async def sample(self, x: np.ndarray, temp=TEMP, top_k=TOP_K) -> np.ndarray:
    """sharded_inference_engine.py:208-228.  Node calls sample(result, temp=0.0) (node.py:121): that is argmax except
    on exact ties, which the seeded Exp(1) noise decides — reproduced bit for bit with the same CPU generator."""
    def samp():
      import torch
      logits = np.asarray(x, dtype=np.float32).reshape(-1, x.shape[-1])
      q = torch.empty((logits.shape[0], logits.shape[1])).exponential_(1, generator=self._rng).numpy()
      return reference_sample(logits, temp, top_k, q)
    return await self._run(samp)

  async def infer_tensor(self, request_id: str, shard: Shard, input_data: np.ndarray,
                         inference_state: Optional[dict] = None) -> Tuple[np.ndarray, Optional[dict]]:
    """sharded_inference_engine.py:230-370.  Position protocol: only `curr_pos` crosses the hop; as in the reference
    (:356-359) only the LAST shard advances it.  `inference_state is None` (or curr_pos == 0) starts a new request:
    KV and position are reset (what `encode` does on the first shard, :133-147)."""
    await self.ensure_shard(shard)
    input_data = np.asarray(input_data)
    if input_data.ndim not in (2, 3):
      raise ValueError(f"input_data.ndim must be 2 (token ids) or 3 (hidden state), got {input_data.ndim}")
    start_pos = int(inference_state["curr_pos"]) if inference_state and "curr_pos" in inference_state else 0
    hop_in = parse_hop_descriptor(input_data)               # the previous shard already put the hidden state into our HBM
    if hop_in is not None and hop_in[2] != self.model_config["hidden_size"]:
      raise ValueError(f"hop descriptor for hidden size {hop_in[2]} reached a model of hidden size {self.model_config['hidden_size']}")
    n = hop_in[1] if hop_in is not None else input_data.shape[1]
    self.request_id = request_id

    def run_native(pos):
      """one step on the native engine; device-resident hop where the link allows it, host arrays otherwise"""
      last = shard.is_last_layer()
      host_hidden_in = hop_in is None and input_data.ndim == 3
      send_dev = (not last) and (not host_hidden_in) and self._hop_ready()
      if hop_in is None and not send_dev:                   # classic step on host arrays (xotb200_forward)
        self.hop_stats["host_in"] += 1 if host_hidden_in else 0
        self.hop_stats["host_out"] += 0 if last else 1
        return self.native.forward(input_data, pos)
      if not last and not send_dev:
        raise RuntimeError("received a device-hop descriptor but the next shard's engine is not attached to this one")
      seq_out, logits = self.native.forward_hop(input_data if hop_in is None else None, -1 if hop_in is None else hop_in[0], n, pos)
      if hop_in is not None:
        self.hop_stats["device_in"] += 1
      if last:
        return logits
      self.hop_stats["device_out"] += 1
      return make_hop_descriptor(seq_out, n, self.model_config["hidden_size"])

    def step():
      # KV capacity = prompt + 1024 like the reference (general_mha.py:152), clamped to the model's context: a prompt that fits
      # the context must not be refused because prompt + 1024 does not
      max_pos = int(self.model_config["max_position_embeddings"])
      if start_pos + n > max_pos:
        raise RuntimeError(f"{start_pos + n} positions exceed the model's context of {max_pos} (max_position_embeddings)")
      need = min(max(n + MAX_GENERATED_TOKENS, int((inference_state or {}).get("max_total_len", 0))), max_pos)
      if self.max_concurrent_requests <= 1:
        if start_pos == 0:
          self.native.begin_request(need)
          self._kv_capacity = need
        return run_native(start_pos)
      # several requests in flight: one KV slot per request id (xotb200_begin_batch / xotb200_select_request)
      if start_pos == 0:
        if self._kv_capacity == 0:
          self._kv_capacity = max(need, int(self.max_total_len or 0))
          self.native.begin_batch(self._kv_capacity, self.max_concurrent_requests)
          self._slots.clear()
        if need - MAX_GENERATED_TOKENS > self._kv_capacity:
          raise RuntimeError(f"prompt of {n} tokens exceeds the per-request KV capacity {self._kv_capacity} (max_total_len)")
        if request_id not in self._slots:
          if len(self._slots) >= self.max_concurrent_requests:   # evict the least recently used request's cache
            self._slots.pop(next(iter(self._slots)))
          used = set(self._slots.values())
          self._slots[request_id] = next(i for i in range(self.max_concurrent_requests) if i not in used)
      elif request_id not in self._slots:
        raise RuntimeError(f"request {request_id!r} has no KV cache on this engine (evicted or never started here)")
      slot = self._slots.pop(request_id)
      self._slots[request_id] = slot                             # most recently used last
      if start_pos + n > self._kv_capacity:
        raise RuntimeError(f"request {request_id!r} reached the KV capacity {self._kv_capacity}")
      self.native.select_request(slot)
      return run_native(start_pos)
    out = await self._run(step)
    if shard.is_last_layer():
      new_pos = start_pos + n          # the reference advances by 1 per decode step (:356-359); n > 1 continuation chunks advance by n
    else:
      new_pos = start_pos
    self._pos = new_pos
    return out, {"curr_pos": new_pos}

  def release_request(self, request_id: str) -> None:
    """Free the KV slot of a finished request (Node knows when a request ends: node.py:119-147)."""
    self._slots.pop(request_id, None)

  async def load_checkpoint(self, shard: Shard, path: str):
    await self.ensure_shard(shard)                          # sharded_inference_engine.py:423-424


def register_with_xotorch() -> bool:
  """Install this engine into an importable reference package (no plugin registry exists upstream: the factory is an
  if/elif and model cards are keyed by engine CLASS NAME — SURVEY.md §8b).  Returns False if xotorch is absent."""
  try:
    import xotorch.inference.inference_engine as ref_ie
    import xotorch.models as ref_models
  except Exception:
    return False
  ref_ie.InferenceEngine.register(B200ShardInferenceEngine)
  ref_ie.inference_engine_classes["b200"] = "B200ShardInferenceEngine"
  orig = ref_ie.get_inference_engine

  def patched(name, shard_downloader):
    if name == "b200":
      return B200ShardInferenceEngine(shard_downloader)
    return orig(name, shard_downloader)
  ref_ie.get_inference_engine = patched
  for card in ref_models.model_cards.values():
    repo = card.get("repo", {})
    if "TorchDynamicShardInferenceEngine" in repo:          # same HF repos; a local dir is resolved by LocalDirDownloader
      repo.setdefault("B200ShardInferenceEngine", repo["TorchDynamicShardInferenceEngine"])
  return True
