"""`NativeShard` — thin, synchronous Python owner of one `xotb200_engine*` (include/xotb200.h).  All compute is in
libxotb200.so; this class only marshals numpy buffers and raises on error codes."""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import ModelCfg, check


def cfg_from_hf(hf: dict) -> ModelCfg:
  """HF config.json dict -> xotb200_model_cfg, following load_model_config
  (/root/reference/xotorch/inference/llm_utils.py:79-126) with App. B2/B3 fixed (HF rope_scaling keys; unscaled RoPE
  when rope_scaling is null)."""
  dt = str(hf.get("torch_dtype", "bfloat16"))
  if dt in ("bfloat16", "bf16"):
    dtype = _lib.DTYPE_BF16
  elif dt in ("float16", "fp16", "half"):
    dtype = _lib.DTYPE_F16
  else:
    raise ValueError(f"torch_dtype {dt!r}: libxotb200 implements 16-bit weights (bfloat16, float16) with fp32 accumulation")
  rs = hf.get("rope_scaling")
  # Only what the kernels implement is accepted; anything else would load without error and give wrong logits
  # (the reference builds biased attention for Qwen2, general_mha.py:45-53; other RoPE scalings exist upstream in HF).
  if rs is not None:
    kind = rs.get("rope_type", rs.get("type"))
    if kind != "llama3":
      raise ValueError(f"rope_scaling type {kind!r}: libxotb200 implements plain RoPE and the 'llama3' scaling only")
  for key in ("attention_bias", "mlp_bias"):
    if hf.get(key):
      raise ValueError(f"{key}=true: biased projections are not implemented by libxotb200")
  if hf.get("hidden_act", "silu") != "silu":
    raise ValueError(f"hidden_act {hf.get('hidden_act')!r}: only SwiGLU (silu) MLPs are implemented")
  if hf.get("sliding_window") not in (None, 0):
    raise ValueError("sliding_window attention is not implemented by libxotb200")
  mt = hf.get("model_type")
  if mt not in (None, "llama", "mixtral", "mistral"):
    raise ValueError(f"model_type {mt!r}: libxotb200 implements Llama-style dense and Mixtral-style sparse-MoE decoders")
  c = ModelCfg()
  c.n_layers = hf["num_hidden_layers"]
  c.dim = hf["hidden_size"]
  c.ffn = hf["intermediate_size"]
  c.n_heads = hf["num_attention_heads"]
  c.n_kv_heads = hf["num_key_value_heads"]
  c.head_dim = hf.get("head_dim") or hf["hidden_size"] // hf["num_attention_heads"]
  c.vocab = hf["vocab_size"]
  c.max_pos = int(hf["max_position_embeddings"])
  c.rope_theta = float(hf.get("rope_theta", 10000.0))
  c.norm_eps = float(hf["rms_norm_eps"])
  c.tie_embeddings = 1 if hf.get("tie_word_embeddings", False) else 0
  c.dtype = dtype
  c.n_experts = int(hf.get("num_local_experts") or 0)          # Mixtral-style sparse MoE MLP
  c.experts_per_tok = int(hf.get("num_experts_per_tok") or 0)
  if rs is not None:
    c.rope_kind = _lib.ROPE_LLAMA3
    c.rope_factor = float(rs.get("factor", 32))
    c.rope_low_freq_factor = float(rs.get("low_freq_factor", 1))
    c.rope_high_freq_factor = float(rs.get("high_freq_factor", 4))
    c.rope_old_context_len = int(rs.get("original_max_position_embeddings", 8192))
  else:
    c.rope_kind = _lib.ROPE_PLAIN
    c.rope_factor, c.rope_low_freq_factor, c.rope_high_freq_factor, c.rope_old_context_len = 1.0, 1.0, 4.0, 8192
  return c


class NativeShard:
  def __init__(self, hf_config: dict, device: int = 0):
    self.lib = _lib.load()
    self.hf_config = dict(hf_config)
    self.cfg = cfg_from_hf(hf_config)
    self.device = device
    h = C.c_void_p()
    check(self.lib.xotb200_create(C.byref(self.cfg), device, C.byref(h)))
    self.h = h
    self.start = self.end = -1

  # -- lifetime ----------------------------------------------------------------------------------------------
  def close(self):
    if getattr(self, "h", None):
      self.lib.xotb200_destroy(self.h)
      self.h = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  @property
  def is_first(self) -> bool: return self.start == 0
  @property
  def is_last(self) -> bool: return self.end == self.cfg.n_layers - 1

  # -- weights -----------------------------------------------------------------------------------------------
  def load_dir(self, model_dir, start_layer: int, end_layer: int):
    check(self.lib.xotb200_load_shard(self.h, str(model_dir).encode(), start_layer, end_layer))
    self.start, self.end = start_layer, end_layer

  def set_shard(self, start_layer: int, end_layer: int):
    check(self.lib.xotb200_set_shard(self.h, start_layer, end_layer))
    self.start, self.end = start_layer, end_layer

  def init_synthetic(self, seed: int = 1234):
    check(self.lib.xotb200_init_synthetic(self.h, C.c_uint64(seed)))

  def load_tensor(self, hf_name: str, bf16_bits: np.ndarray) -> bool:
    a = np.ascontiguousarray(bf16_bits)
    assert a.dtype in (np.uint16, np.int16), "pass bf16 bit patterns (uint16)"
    shp = (C.c_int64 * a.ndim)(*a.shape)
    return check(self.lib.xotb200_load_tensor(self.h, hf_name.encode(), a.ctypes.data_as(C.c_void_p), shp, a.ndim)) == 0

  def check_loaded(self):
    check(self.lib.xotb200_check_loaded(self.h))

  # -- request / step ----------------------------------------------------------------------------------------
  def begin_request(self, max_total_len: int):
    check(self.lib.xotb200_begin_request(self.h, int(max_total_len)))

  def begin_batch(self, max_total_len: int, n_requests: int):
    """n_requests independent requests in flight through the ring, one KV cache each"""
    check(self.lib.xotb200_begin_batch(self.h, int(max_total_len), int(n_requests)))

  def select_request(self, request: int):
    check(self.lib.xotb200_select_request(self.h, int(request)))

  def forward(self, input_data: np.ndarray, pos0: int) -> np.ndarray:
    """ndim==2 int ids [1,S] (first shard) or ndim==3 float32 hidden [1,S,D] -> hidden [1,S,D] f32 or logits [1,V] f32."""
    if input_data.ndim == 2:
      ids = np.ascontiguousarray(input_data, dtype=np.int32)
      n = ids.shape[1]
      tok_p, hid_p = ids.ctypes.data_as(C.c_void_p), None
    elif input_data.ndim == 3:
      hid = np.ascontiguousarray(input_data, dtype=np.float32)
      n = hid.shape[1]
      tok_p, hid_p = None, hid.ctypes.data_as(C.c_void_p)
    else:
      raise ValueError(f"input_data.ndim must be 2 (token ids) or 3 (hidden state), got {input_data.ndim}")
    if self.is_last:
      out = np.empty((1, self.cfg.vocab), dtype=np.float32)
    else:
      out = np.empty((1, n, self.cfg.dim), dtype=np.float32)
    check(self.lib.xotb200_forward(self.h, tok_p, hid_p, n, int(pos0), out.ctypes.data_as(C.c_void_p)))
    return out

  def forward_device(self, ids: Optional[np.ndarray], hidden_dev_ptr: Optional[int], n_tokens: int, pos0: int,
                     out_dev_ptr: Optional[int]) -> Optional[np.ndarray]:
    """The same step with the hidden state in DEVICE memory (16-bit, engine dtype, [n_tokens, dim]) on both sides:
    `hidden_dev_ptr` / `out_dev_ptr` are raw device pointers (e.g. torch.Tensor.data_ptr() of a tensor on this GPU whose
    producer stream has been synchronised).  Returns the logits [1, V] on the last shard, else None."""
    tok_p = None
    if ids is not None:
      ids = np.ascontiguousarray(ids, dtype=np.int32)
      tok_p = ids.ctypes.data_as(C.c_void_p)
    logits = np.empty((1, self.cfg.vocab), dtype=np.float32) if self.is_last else None
    check(self.lib.xotb200_forward_device(self.h, tok_p, C.c_void_p(hidden_dev_ptr) if hidden_dev_ptr else None, int(n_tokens),
                                          int(pos0), C.c_void_p(out_dev_ptr) if out_dev_ptr else None,
                                          logits.ctypes.data_as(C.c_void_p) if logits is not None else None))
    return logits

  def forward_hop(self, ids: Optional[np.ndarray], seq_in: int, n_tokens: int, pos0: int) -> Tuple[Optional[int], Optional[np.ndarray]]:
    """One step with the hidden state handed over on the device (xotb200_forward_hop): first shard passes `ids` [1, S], other
    shards the hop number `seq_in` their predecessor returned.  Returns (hop number produced | None, logits [1, V] | None).
    For S == 1 a non-last shard returns without waiting for its kernel: the next shard's kernel acquires the flag itself."""
    tok_p = None
    if ids is not None:
      ids = np.ascontiguousarray(ids, dtype=np.int32)
      tok_p = ids.ctypes.data_as(C.c_void_p)
    logits = np.empty((1, self.cfg.vocab), dtype=np.float32) if self.is_last else None
    seq_out = C.c_int64(-1)
    check(self.lib.xotb200_forward_hop(self.h, tok_p, C.c_int64(int(seq_in)), int(n_tokens), int(pos0), C.byref(seq_out),
                                       logits.ctypes.data_as(C.c_void_p) if logits is not None else None))
    return (None if self.is_last else int(seq_out.value)), logits

  def last_token(self) -> int:
    t = C.c_int32()
    check(self.lib.xotb200_last_token(self.h, C.byref(t)))
    return int(t.value)

  # -- ring ----------------------------------------------------------------------------------------------------
  def inbox_export(self) -> bytes:
    buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
    check(self.lib.xotb200_inbox_export(self.h, buf))
    return buf.raw

  def link_next_ipc(self, handle: bytes):
    assert len(handle) == _lib.IPC_HANDLE_BYTES
    check(self.lib.xotb200_link_next_ipc(self.h, C.create_string_buffer(handle, _lib.IPC_HANDLE_BYTES)))

  def link_next_local(self, nxt: "NativeShard"):
    check(self.lib.xotb200_link_next_local(self.h, nxt.h))

  def post_to_next(self, step: int):
    check(self.lib.xotb200_post_to_next(self.h, int(step)))

  def ring_launch(self, n_steps: int, step0: int, pos0: int):
    check(self.lib.xotb200_ring_launch(self.h, int(n_steps), int(step0), int(pos0)))

  def ring_wait(self, n_steps: int = 0) -> Tuple[Optional[np.ndarray], float]:
    ms = C.c_float()
    toks = None
    if self.is_last and n_steps > 0:
      toks = np.empty(n_steps, dtype=np.int32)
      check(self.lib.xotb200_ring_wait(self.h, toks.ctypes.data_as(C.c_void_p), n_steps, C.byref(ms)))
    else:
      check(self.lib.xotb200_ring_wait(self.h, None, 0, C.byref(ms)))
    return toks, float(ms.value)

  def ring_launch_slots(self, n_steps: int, step0: int, kv_slots, positions):
    """ring launch over an explicit set of active requests: request i lives in KV cache kv_slots[i] at position positions[i]"""
    sl = np.ascontiguousarray(kv_slots, dtype=np.int32); ps = np.ascontiguousarray(positions, dtype=np.int32)
    assert sl.shape == ps.shape and sl.ndim == 1
    check(self.lib.xotb200_ring_launch_slots(self.h, int(n_steps), int(step0), sl.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p), int(sl.size)))

  def post_token(self, step: int, token: int, n_active: int):
    check(self.lib.xotb200_post_token(self.h, int(step), int(token), int(n_active)))

  def ring_poll(self, start: int = 0, max_n: int = 4096) -> Tuple[np.ndarray, int]:
    """(token ids [start : done) produced so far by the running ring launch, done) — does not wait, safe while the kernel runs"""
    buf = np.empty(max(max_n, 1), dtype=np.int32)
    done = C.c_int()
    check(self.lib.xotb200_ring_poll(self.h, buf.ctypes.data_as(C.c_void_p), int(start), int(max_n), C.byref(done)))
    n = max(0, min(int(done.value) - start, max_n))
    return buf[:n].copy(), int(done.value)

  def clear_request(self, request: int):
    check(self.lib.xotb200_clear_request(self.h, int(request)))

  def ring_next_step(self) -> int:
    """first ring step number not yet used on this engine (inbox flags are monotonic and never cleared)"""
    v = C.c_int()
    check(self.lib.xotb200_ring_next_step(self.h, C.byref(v)))
    return int(v.value)

  def ring_timestamps(self, n_steps: int) -> Tuple[np.ndarray, np.ndarray]:
    t_in = np.empty(n_steps, dtype=np.uint64)
    t_out = np.empty(n_steps, dtype=np.uint64)
    check(self.lib.xotb200_ring_timestamps(self.h, t_in.ctypes.data_as(C.c_void_p), t_out.ctypes.data_as(C.c_void_p), n_steps))
    return t_in, t_out

  def set_cta_limit(self, max_ctas: int):
    check(self.lib.xotb200_set_cta_limit(self.h, int(max_ctas)))

  def set_prefill_mode(self, mode: int):
    """0 auto, 1 decode kernel only, 2 tensor-core GEMM prefill only"""
    check(self.lib.xotb200_set_prefill_mode(self.h, int(mode)))

  def debug_gemm(self, a_bits: np.ndarray, w_bits: np.ndarray, tk: int) -> np.ndarray:
    """bf16 bit patterns in (uint16): A [S,K], W [N,K] -> out [S,N] bf16 bits, through the tcgen05 GEMM"""
    a = np.ascontiguousarray(a_bits, dtype=np.uint16); w = np.ascontiguousarray(w_bits, dtype=np.uint16)
    S, K = a.shape; N = w.shape[0]
    out = np.empty((S, N), dtype=np.uint16)
    check(self.lib.xotb200_debug_gemm(self.h, a.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), S, N, K, int(tk),
                                      out.ctypes.data_as(C.c_void_p)))
    return out

  def debug_routing(self, n_positions: int) -> np.ndarray:
    """[n_positions, local layers, 2] experts chosen by the router (ascending index) for request 0; -1 = not run"""
    nl = self.end - self.start + 1
    out = np.empty(n_positions * nl, dtype=np.int32)
    check(self.lib.xotb200_debug_routing(self.h, out.ctypes.data_as(C.POINTER(C.c_int32)), int(n_positions)))
    out = out.reshape(n_positions, nl)
    return np.stack([np.where(out < 0, -1, out & 0xff), np.where(out < 0, -1, (out >> 8) & 0xff)], axis=-1)

  def debug_timeline(self, step: int, cta: int = 0, read: bool = False):
    n = (4 * self.cfg.n_layers + 2) * 4 * 2 + 4 * 64 + 8
    if read == "all":      # + every CTA's (consume start, tiles done, landed stages) per op: [4 * n_layers + 1][256][3]
      n += (4 * self.cfg.n_layers + 1) * 256 * 3
    if read:
      out = np.zeros(n, dtype=np.uint64)
      check(self.lib.xotb200_debug_timeline(self.h, int(step), int(cta), out.ctypes.data_as(C.c_void_p), n))
      return out
    check(self.lib.xotb200_debug_timeline(self.h, int(step), int(cta), None, 0))
    return None

  # -- introspection ---------------------------------------------------------------------------------------------
  def kernel_launches(self) -> int:
    v = C.c_int64()
    check(self.lib.xotb200_kernel_launches(self.h, C.byref(v)))
    return int(v.value)

  def load_stats(self) -> dict:
    """the most recent load_dir: bytes into HBM, wall seconds, GB/s, and the seconds spent copying out of the file mapping"""
    b, t, tm = C.c_int64(), C.c_double(), C.c_double()
    check(self.lib.xotb200_load_stats(self.h, C.byref(b), C.byref(t), C.byref(tm)))
    return {"bytes": int(b.value), "seconds": float(t.value), "GBps": (b.value / t.value / 1e9) if t.value > 0 else 0.0,
            "memcpy_seconds": float(tm.value)}

  def weight_bytes(self) -> int:
    v = C.c_int64()
    check(self.lib.xotb200_weight_bytes(self.h, C.byref(v)))
    return int(v.value)

  def step_bytes(self, ctx: int) -> int:
    v = C.c_int64()
    check(self.lib.xotb200_step_bytes(self.h, int(ctx), C.byref(v)))
    return int(v.value)


def load_hf_config(model_dir) -> dict:
  with open(Path(model_dir) / "config.json", "r") as f:
    return json.load(f)
