"""Host-side logic of the drop-in boundary, on CPU, with the CUDA library replaced by a stand-in object
(`native_factory`) — no compute is claimed here; the arithmetic is tested in tests/test_gpu_parity.py.

Covers what the reference's Node relies on: API surface, ndim protocol, the tiny inference_state,
the temp=0 sampler semantics (bit-exact with the oracle's restatement of torchtune's sampler, same seeded generator),
ensure_shard caching, and — when /root/reference is present — the reference's OWN `Node` driving two of our engines
through the in-process peer handle (the drop-in proof)."""
import asyncio
import sys
import types
from pathlib import Path

import numpy as np
import pytest

from xotorch_b200.inference.b200_engine import B200ShardInferenceEngine, LocalDirDownloader, reference_sample
from xotorch_b200.inference.inference_engine import InferenceEngine, get_inference_engine
from xotorch_b200.inference.shard import Shard
from xotorch_b200.synth import PRESETS, hf_config, write_hf_dir

NAME = "tiny-llama-hd64"
SHAPE = PRESETS[NAME]


class FakeNative:
  """Records calls; produces deterministic arrays of the right shapes/dtypes."""
  instances = []

  def __init__(self, cfg, device):
    self.cfg, self.device, self.calls = cfg, device, []
    self.start = self.end = -1
    FakeNative.instances.append(self)

  def load_dir(self, d, a, b): self.start, self.end = a, b; self.calls.append(("load", a, b))
  def begin_request(self, n): self.calls.append(("begin", n))
  def begin_batch(self, n, b): self.calls.append(("begin_batch", n, b))
  def select_request(self, r): self.calls.append(("select", r))
  def close(self): self.calls.append(("close",))

  def link_next_local(self, nxt): self.calls.append(("link", id(nxt))); self.hop_seq = 0

  def forward_hop(self, ids, seq_in, n, pos0):
    """stand-in for xotb200_forward_hop: only the protocol (who gets ids / a hop number, what comes back)"""
    self.calls.append(("forward_hop", None if ids is None else ids.shape, seq_in, n, pos0))
    if self.end == self.cfg["num_hidden_layers"] - 1:
      out = np.zeros((1, self.cfg["vocab_size"]), np.float32)
      out[0, (7 * pos0 + 3 * n) % self.cfg["vocab_size"]] = 5.0
      return None, out
    self.hop_seq += 1
    return self.hop_seq - 1, None

  def forward(self, x, pos0):
    self.calls.append(("forward", x.ndim, x.shape, pos0))
    last = self.end == self.cfg["num_hidden_layers"] - 1
    n = x.shape[1]
    if last:
      out = np.zeros((1, self.cfg["vocab_size"]), np.float32)
      out[0, (7 * pos0 + 3 * n) % self.cfg["vocab_size"]] = 5.0
      return out
    base = x.astype(np.float32).sum() if x.ndim == 2 else float(x.sum())
    return np.full((1, n, self.cfg["hidden_size"]), np.float32(base % 97), np.float32)


@pytest.fixture(scope="module")
def model_dir(tmp_path_factory):
  return write_hf_dir(tmp_path_factory.mktemp("m") / NAME, SHAPE, 1234)


def _engine(model_dir):
  return B200ShardInferenceEngine(LocalDirDownloader({NAME: model_dir}), device=0, native_factory=FakeNative)


def test_is_an_inference_engine_and_factory():
  assert issubclass(B200ShardInferenceEngine, InferenceEngine)
  assert isinstance(get_inference_engine("b200", None), B200ShardInferenceEngine)
  with pytest.raises(ValueError):
    get_inference_engine("torch", None)


def test_ndim_and_position_protocol(model_dir):
  async def run():
    e0, e1 = _engine(model_dir), _engine(model_dir)
    s0, s1 = Shard(NAME, 0, 1, 4), Shard(NAME, 2, 3, 4)
    ids = np.arange(5, dtype=np.int64).reshape(1, 5)
    h, st = await e0.infer_tensor("r", s0, ids, None)
    assert h.shape == (1, 5, 512) and h.dtype == np.float32 and st == {"curr_pos": 0}      # non-last shard: position unchanged
    logits, st = await e1.infer_tensor("r", s1, h, st)
    assert logits.shape == (1, 2048) and logits.dtype == np.float32 and st == {"curr_pos": 5}
    tok = await e1.sample(logits, temp=0.0)
    assert tok.shape == (1, 1) and tok.dtype == np.int32
    h, st2 = await e0.infer_tensor("r", s0, tok.reshape(1, -1), st)
    assert h.shape == (1, 1, 512) and st2 == {"curr_pos": 5}
    logits, st3 = await e1.infer_tensor("r", s1, h, st2)
    assert st3 == {"curr_pos": 6}
    n0, n1 = e0.native, e1.native
    assert [c[0] for c in n0.calls] == ["load", "begin", "forward", "forward"]          # KV reset only on the prefill
    assert n0.calls[1] == ("begin", 5 + 1024)                                           # prompt + 1024 (general_mha.py:152)
    assert n0.calls[3] == ("forward", 2, (1, 1), 5) and n1.calls[3] == ("forward", 3, (1, 1, 512), 5)
    with pytest.raises(ValueError):
      await e0.infer_tensor("r", s0, np.zeros(3), None)
  asyncio.run(run())


def test_concurrent_request_ids_get_their_own_kv_slot(model_dir):
  """max_concurrent_requests > 1: interleaved request ids are routed to distinct KV slots (select_request before every
  forward), the least recently used request is evicted when the slots run out, and an evicted request fails loudly."""
  async def run():
    e = B200ShardInferenceEngine(LocalDirDownloader({NAME: model_dir}), device=0, native_factory=FakeNative,
                                 max_concurrent_requests=2, max_total_len=64)
    sh = Shard(NAME, 0, 3, 4)
    ids = np.arange(4, dtype=np.int64).reshape(1, 4)
    _, sa = await e.infer_tensor("a", sh, ids, None)
    _, sb = await e.infer_tensor("b", sh, ids, None)
    _, sa = await e.infer_tensor("a", sh, ids[:, :1], sa)
    _, sb = await e.infer_tensor("b", sh, ids[:, :1], sb)
    calls = [c for c in e.native.calls if c[0] in ("begin_batch", "select", "forward")]
    assert calls[0] == ("begin_batch", 4 + 1024, 2)                       # one allocation for both slots
    assert [c[1] for c in calls if c[0] == "select"] == [0, 1, 0, 1]
    assert [c[3] for c in calls if c[0] == "forward"] == [0, 0, 4, 4] and sa == sb == {"curr_pos": 5}
    _, sc = await e.infer_tensor("c", sh, ids, None)                      # third request: evicts "a" (least recently used)
    assert [c[1] for c in e.native.calls if c[0] == "select"][-1] == 0
    with pytest.raises(RuntimeError, match="no KV cache"):
      await e.infer_tensor("a", sh, ids[:, :1], sa)
    e.release_request("b")
    _, sd = await e.infer_tensor("d", sh, ids, None)                      # takes the slot "b" released
    assert [c[1] for c in e.native.calls if c[0] == "select"][-1] == 1
  asyncio.run(run())


def test_ensure_shard_is_cached_and_reloads_on_change(model_dir):
  async def run():
    e = _engine(model_dir)
    await e.ensure_shard(Shard(NAME, 0, 3, 4))
    n = e.native
    await e.ensure_shard(Shard(NAME, 0, 3, 4))
    assert e.native is n and [c[0] for c in n.calls] == ["load"]
    await e.ensure_shard(Shard(NAME, 0, 1, 4))
    assert e.native is not n and ("close",) in n.calls
    assert e.tokenizer.eos_token_id == -1 and e.shard == Shard(NAME, 0, 1, 4)
    with pytest.raises(ValueError):
      await e.ensure_shard(Shard(NAME, 0, 1, 7))
    with pytest.raises(FileNotFoundError):
      await B200ShardInferenceEngine(LocalDirDownloader({}), native_factory=FakeNative).ensure_shard(Shard("nope", 0, 0, 1))
  asyncio.run(run())


def test_encode_decode_infer_prompt(model_dir):
  async def run():
    e = _engine(model_dir)
    s = Shard(NAME, 0, 3, 4)
    ids = await e.encode(s, "5 6 7")
    assert ids.tolist() == [5, 6, 7]
    assert await e.decode(s, np.array([1, 2])) == "1 2"
    out, st = await e.infer_prompt("r", s, "5 6 7")
    assert out.shape == (1, 2048) and st == {"curr_pos": 3}
  asyncio.run(run())


def test_sampler_is_bit_identical_to_the_oracle_restatement(model_dir):
  import torch
  from oracle.ref_engine import ref_sample
  rng = np.random.default_rng(0)
  async def run():
    e = _engine(model_dir)
    await e.ensure_shard(Shard(NAME, 0, 3, 4))
    g = torch.Generator(device="cpu"); g.manual_seed(1234)
    for temp, k in [(0.0, 35), (0.0, 35), (0.6, 35), (1.0, 5), (0.0, 35)]:
      x = rng.standard_normal((1, 2048)).astype(np.float32)
      x = np.round(x * 8) / 8                     # coarse grid => exact ties among the top logits, like bf16 logits
      q = torch.empty((1, 2048)).exponential_(1, generator=g)
      want = ref_sample(torch.tensor(x).clone(), temperature=temp, top_k=k, q=q).numpy()
      got = await e.sample(x, temp=temp, top_k=k)
      assert np.array_equal(got, want)
  asyncio.run(run())
  # numpy restatement alone: ties go to the smallest q
  x = np.array([[1.0, 5.0, 5.0, 0.0]], np.float32)
  assert int(reference_sample(x, 0.0, 35, np.array([[1.0, 2.0, 0.5, 1.0]], np.float32))[0, 0]) == 2


# ------------------------------------------------------------------------------------------------------------------
# drop-in proof: the REFERENCE's own Node drives our engine (only where /root/reference exists: the build container)
# ------------------------------------------------------------------------------------------------------------------
from tests._refpkg import reference_root, import_reference


@pytest.mark.skipif(reference_root() is None, reason="the reference package is neither installed under baseline/_ref nor at /root/reference")
@pytest.mark.parametrize("device_hop", [False, True])
def test_reference_node_drives_b200_engines_in_a_two_node_ring(model_dir, tmp_path, monkeypatch, device_hop):
  import_reference(monkeypatch, tmp_path)
  from xotorch.orchestration.node import Node
  from xotorch.topology.ring_memory_weighted_partitioning_strategy import RingMemoryWeightedPartitioningStrategy
  from xotorch.topology.device_capabilities import DeviceCapabilities, DeviceFlops
  from xotorch.topology.topology import Topology
  from xotorch.inference.shard import Shard as RefShard
  from xotorch_b200.networking.nvlink_peer_handle import InProcessPeerHandle, NVLinkPeerHandle
  from unittest import mock

  caps = DeviceCapabilities(model="B200", chip="B200", memory=183359, flops=DeviceFlops(fp32=0, fp16=0, int8=0))

  async def run():
    engines = [_engine(model_dir), _engine(model_dir)]
    nodes = []
    for i, eng in enumerate(engines):
      server, discovery = mock.AsyncMock(), mock.AsyncMock()
      discovery.discover_peers = mock.AsyncMock(return_value=[])
      n = Node(f"node{i}", server, eng, discovery, eng.shard_downloader, RingMemoryWeightedPartitioningStrategy(),
               max_generate_tokens=4)
      n.device_capabilities = caps
      nodes.append(n)
    topo = Topology()
    for n in nodes:
      topo.update_node(n.id, caps)
    for i, n in enumerate(nodes):
      n.topology = topo
      if device_hop:      # the handle attaches the peer's engine to the local one: hidden states stay on the device
        n.peers = [NVLinkPeerHandle(nodes[1 - i].id, nodes[1 - i], local_engine=engines[i], device_capabilities=caps)]
        await n.peers[0].connect()
      else:
        n.peers = [InProcessPeerHandle(nodes[1 - i].id, nodes[1 - i], device_capabilities=caps)]
      n.select_best_inference_engine = mock.AsyncMock()       # App. B9: upstream would swap in the torch engine
    got = []
    # the last-shard node fires on_token locally and broadcasts the delta to its peers (node.py:136-137): listen on one
    nodes[0].on_token.register("t").on_next(lambda rid, toks, fin: got.append((rid, list(toks), fin)))
    base = RefShard(NAME, 0, 0, 4)
    # ring order: (memory, id) descending -> node1 first
    await nodes[1].process_prompt(base, "1 2 3", "req")
    for _ in range(200):
      await asyncio.sleep(0.01)
      if any(fin for _, _, fin in got):
        break
    assert any(fin for _, _, fin in got), got
    toks = [t for _, ts, _ in got for t in ts]
    assert len(toks) == 4
    # the ring really crossed both engines with the ndim protocol and the tiny state
    if not device_hop:
      c1 = [c for c in engines[1].native.calls if c[0] == "forward"]
      c0 = [c for c in engines[0].native.calls if c[0] == "forward"]
      assert c1[0][1] == 2 and c0[0][1] == 3 and len(c0) == 4
    else:
      # first shard (node1): the prompt's hidden state still goes by host array once (the peer engine has not loaded its shard
      # when the first hop is made), every later token leaves as a hop descriptor; last shard (node0) consumes hop numbers
      c1 = [c for c in engines[1].native.calls if c[0] in ("forward", "forward_hop")]
      c0 = [c for c in engines[0].native.calls if c[0] in ("forward", "forward_hop")]
      assert [c[0] for c in c1] == ["forward", "forward_hop", "forward_hop", "forward_hop"]
      assert [c[0] for c in c0] == ["forward", "forward_hop", "forward_hop", "forward_hop"]
      assert [c[2] for c in c0[1:]] == [0, 1, 2] and all(c[1] is None for c in c0[1:])        # hop numbers in, no ids
      assert all(c[1] == (1, 1) and c[2] == -1 for c in c1[1:])                               # ids in on the first shard
      assert nodes[1].peers[0].sent["descriptors"] == 3 and engines[1].hop_stats["device_out"] == 3
      assert engines[0].hop_stats["device_in"] == 3
  asyncio.run(run())


def test_kv_capacity_is_clamped_to_the_model_context_and_chunks_advance_by_n(model_dir):
  """prompt + 1024 must be clamped to max_position_embeddings (tiny model: 2048); a prompt that does not fit the
  context fails loudly; a multi-token continuation chunk advances curr_pos by its length."""
  async def run():
    e = _engine(model_dir)
    sh = Shard(NAME, 0, 3, 4)
    ids = np.zeros((1, 1500), dtype=np.int64)
    _, st = await e.infer_tensor("r", sh, ids, None)
    assert ("begin", 2048) in e.native.calls and st == {"curr_pos": 1500}             # 1500 + 1024 > 2048 -> clamped
    _, st = await e.infer_tensor("r", sh, ids[:, :3], st)                               # continuation chunk of 3 tokens
    assert st == {"curr_pos": 1503}
    with pytest.raises(RuntimeError, match="exceed the model's context"):
      await e.infer_tensor("q", sh, np.zeros((1, 2049), dtype=np.int64), None)
  asyncio.run(run())


def test_cfg_from_hf_rejects_what_the_kernels_do_not_implement():
  from xotorch_b200.native import cfg_from_hf
  base = hf_config(SHAPE)
  cfg_from_hf(base)
  cfg_from_hf({**base, "rope_scaling": {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                                         "original_max_position_embeddings": 8192}})
  for bad in ({"rope_scaling": {"type": "linear", "factor": 2.0}}, {"rope_scaling": {"rope_type": "yarn", "factor": 4.0}},
              {"attention_bias": True}, {"mlp_bias": True}, {"hidden_act": "gelu"}, {"sliding_window": 4096}, {"model_type": "qwen2"}):
    with pytest.raises(ValueError):
      cfg_from_hf({**base, **bad})
