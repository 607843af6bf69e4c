"""The drop-in proof on the GPU (-m gpu): the reference's OWN, unmodified `Node` (orchestration/node.py) drives two real
`B200ShardInferenceEngine`s (libxotb200.so through the C ABI) around a two-shard ring, hop by hop through
`NVLinkPeerHandle`, and the generated ids equal the CPU oracle's.

  Node.process_prompt -> engine.infer_prompt -> Node.process_inference_result -> Node.forward_tensor
       -> NVLinkPeerHandle.send_tensor(descriptor) -> peer Node.process_tensor -> engine.infer_tensor(descriptor) -> sample -> ...
  (/root/reference/xotorch/orchestration/node.py:109-147,193-208,361-380,424-443; grpc_peer_handle.py:111-130)

Also covered here, without the reference package: the API hop itself (`xotb200_forward_hop`) against the host-array API,
bit for bit, decode steps and prompts, bf16 and fp16, and `DistributedRing`-style pre-launched consumers.

The reference package comes from baseline/_ref (pip-installed from /root/reference, travels to the GPU box) — see tests/_refpkg.py."""
import asyncio

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests._refpkg import import_reference, reference_root
from xotorch_b200.synth import PRESETS, hf_config, synth_prompt, write_hf_dir

NAME = "tiny-llama-hd64"


def _oracle_ids(shape_name, prompt_ids, n_gen, sampler):
  from oracle.ref_engine import RefEngine, RefShard, model_config_from_hf, generate_ids
  from xotorch_b200.synth import synth_state_dict
  shape = PRESETS[shape_name]
  ref = RefEngine(model_config_from_hf(hf_config(shape)), synth_state_dict(shape, 1234), RefShard(shape_name, 0, shape.n_layers - 1, shape.n_layers))
  return generate_ids([ref], np.asarray(prompt_ids, dtype=np.int32).reshape(1, -1), n_gen, sampler=sampler)


def _ngpu():
  import torch
  return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(reference_root() is None, reason="the reference package is neither under baseline/_ref nor at /root/reference")
@pytest.mark.parametrize("device_hop,devices", [(True, (0, 0)), (False, (0, 0)), (True, (0, 1))])
def test_reference_node_ring_on_gpu_matches_oracle(tmp_path, monkeypatch, device_hop, devices):
  if max(devices) >= _ngpu():
    pytest.skip("needs 2 GPUs: the hop crosses NVLink (peer stores into the other GPU's HBM)")
  import_reference(monkeypatch, tmp_path)
  from unittest import mock
  from xotorch.orchestration.node import Node
  from xotorch.topology.ring_memory_weighted_partitioning_strategy import RingMemoryWeightedPartitioningStrategy
  from xotorch.topology.device_capabilities import DeviceCapabilities, DeviceFlops
  from xotorch.topology.topology import Topology
  from xotorch.inference.shard import Shard as RefShard
  from xotorch_b200.inference.b200_engine import B200ShardInferenceEngine, LocalDirDownloader
  from xotorch_b200.networking.nvlink_peer_handle import InProcessPeerHandle, NVLinkPeerHandle

  shape = PRESETS[NAME]
  d = write_hf_dir(tmp_path / NAME, shape, 1234)
  n_gen = 12
  caps = DeviceCapabilities(model="B200", chip="B200", memory=183359, flops=DeviceFlops(fp32=0, fp16=0, int8=0))

  async def run():
    # node0 (last shard) on devices[0], node1 (first shard) on devices[1]
    engines = [B200ShardInferenceEngine(LocalDirDownloader({NAME: d}), device=devices[i]) for i in range(2)]
    nodes = []
    for i, eng in enumerate(engines):
      server, discovery = mock.AsyncMock(), mock.AsyncMock()
      discovery.discover_peers = mock.AsyncMock(return_value=[])
      n = Node(f"node{i}", server, eng, discovery, eng.shard_downloader, RingMemoryWeightedPartitioningStrategy(), max_generate_tokens=n_gen)
      n.device_capabilities = caps
      nodes.append(n)
    topo = Topology()
    for n in nodes:
      topo.update_node(n.id, caps)
    for i, n in enumerate(nodes):
      n.topology = topo
      if device_hop:
        n.peers = [NVLinkPeerHandle(nodes[1 - i].id, nodes[1 - i], local_engine=engines[i], device_capabilities=caps)]
        await n.peers[0].connect()
      else:
        n.peers = [InProcessPeerHandle(nodes[1 - i].id, nodes[1 - i], device_capabilities=caps)]
      n.select_best_inference_engine = mock.AsyncMock()       # upstream would swap in its torch engine here
    got = []
    nodes[0].on_token.register("t").on_next(lambda rid, toks, fin: got.append((rid, list(toks), fin)))
    prompt = "17 4 1999 256 3 77 1024 5"
    # equal memory -> ring order (memory, id) descending: node1 holds layers 0-1, node0 layers 2-3
    await nodes[1].process_prompt(RefShard(NAME, 0, 0, shape.n_layers), prompt, "req")
    for _ in range(3000):
      await asyncio.sleep(0.01)
      if any(fin for _, _, fin in got):
        break
    assert any(fin for _, _, fin in got), got
    return [t for _, ts, _ in got for t in ts], engines, nodes

  toks, engines, nodes = asyncio.run(run())
  assert len(toks) == n_gen
  # Node samples with the reference sampler at temp 0 (node.py:121): argmax, seeded noise on exact ties
  prompt_ids = [17, 4, 1999, 256, 3, 77, 1024, 5]
  o_ids, _, gaps = _oracle_ids(NAME, prompt_ids, n_gen, "reference")
  k = int(np.argmax(np.array(toks) != o_ids)) if not np.array_equal(toks, o_ids) else n_gen
  assert k == n_gen or gaps[k] <= 0.08, f"ids {toks} vs oracle {o_ids.tolist()}: diverge at step {k} with a decisive gap {gaps[k]:.3f}"
  if device_hop:
    # every decode step's hidden state stayed on the device: descriptors on the wire, forward_hop on both engines
    first, last = engines[1], engines[0]
    assert first.hop_stats["device_out"] == n_gen - 1 and last.hop_stats["device_in"] == n_gen - 1
    assert nodes[1].peers[0].sent["descriptors"] == n_gen - 1
    assert first.hop_stats["host_out"] <= 1                      # at most the prompt (the peer engine loads its shard on first use)
  for e in engines:
    e.native.close()


def _pair(shape_name, split, devices=None):
  from xotorch_b200.native import NativeShard
  shape = PRESETS[shape_name]
  out = []
  for i, (a, b) in enumerate(split):
    e = NativeShard(hf_config(shape), device=0 if devices is None else devices[i])
    e.set_shard(a, b)
    e.init_synthetic(1234)
    out.append(e)
  return out


@pytest.mark.parametrize("shape_name,split", [("tiny-llama-hd64", [(0, 1), (2, 3)]), ("tiny-llama-hd64", [(0, 0), (1, 2), (3, 3)]),
                                              ("tiny-llama-hd128-f16", [(0, 0), (1, 2)]), ("tiny-mixtral", [(0, 1), (2, 3)])])
def test_forward_hop_equals_host_array_api(shape_name, split):
  _hop_vs_host(shape_name, split, None)


def test_forward_hop_across_two_gpus():
  if _ngpu() < 2:
    pytest.skip("needs 2 GPUs")
  _hop_vs_host("tiny-llama-hd64", [(0, 1), (2, 3)], [0, 1])
  _hop_vs_host("tiny-llama-hd128-f16", [(0, 0), (1, 2)], [1, 0])


def _hop_vs_host(shape_name, split, devices):
  """xotb200_forward_hop (hidden state handed over by the producing kernel / one peer copy) == xotb200_forward (fp32 host
  arrays), bit for bit: a 150-token prompt (tensor-core prefill + peer copy), a 5-token chunk (decode kernel, multi-token hop)
  and single decode steps (fused inbox hop, consumer launched before the producer has finished)."""
  shape = PRESETS[shape_name]
  prompt = synth_prompt(shape.vocab, 150)
  ref = _pair(shape_name, split, devices)
  for e in ref:
    e.begin_request(220)
  want, x, pos = [], prompt.astype(np.int32), 0
  chunks = [prompt.astype(np.int32), prompt[:, :5].astype(np.int32)] + [np.array([[7 + 3 * i]], dtype=np.int32) for i in range(20)]
  for c in chunks:
    data = c
    for e in ref:
      data = e.forward(data, pos)
    want.append(data.copy())
    pos += c.shape[1]
  hop = _pair(shape_name, split, devices)
  for i, e in enumerate(hop):
    e.begin_request(220)
    if i + 1 < len(hop):
      e.link_next_local(hop[i + 1])
  pos = 0
  for c, w in zip(chunks, want):
    seq, logits = hop[0].forward_hop(c, -1, c.shape[1], pos)
    for e in hop[1:]:
      seq, logits = e.forward_hop(None, seq, c.shape[1], pos)
    assert np.array_equal(logits, w), f"chunk of {c.shape[1]} tokens at pos {pos}"
    pos += c.shape[1]


def test_forward_hop_errors():
  from xotorch_b200._lib import XotB200Error
  a, b = _pair("tiny-llama-hd64", [(0, 1), (2, 3)])
  a.begin_request(32); b.begin_request(32)
  with pytest.raises(XotB200Error, match="not linked"):
    a.forward_hop(np.array([[1]], dtype=np.int32), -1, 1, 0)
  a.link_next_local(b)
  with pytest.raises(XotB200Error, match="hop number|hop to consume"):
    b.forward_hop(None, -1, 1, 0)
  with pytest.raises(XotB200Error, match="no prompt block"):
    b.forward_hop(None, 3, 4, 0)                               # nothing was pushed for hop 3
