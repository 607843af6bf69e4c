"""TEST INFRASTRUCTURE: regenerates tests/golden/*.  Run in the BUILD container (needs /root/reference for part B):
    python oracle/gen_golden.py
A. oracle goldens: for each tiny preset, prompt ids, greedy ids, logits of the prefill + first decode steps, top-2 gaps
   (pins the oracle against silent drift) and the max |oracle - HF LlamaForCausalLM| on the same weights.
B. reference-generated fixtures: the partitioning outputs of the REAL reference code
   (xotorch/topology/ring_memory_weighted_partitioning_strategy.py + partitioning_strategy.py), imported from
   /root/reference with a scapy stub, for the memories/layer counts our tests use.  Also the DummyInferenceEngine
   behaviour through the reference's own InferenceEngine ABC (the boundary fixture).
"""
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"

from oracle.ref_engine import RefEngine, RefShard, model_config_from_hf, generate_ids  # noqa: E402
from xotorch_b200.synth import PRESETS, hf_config, synth_state_dict, synth_prompt  # noqa: E402

TINY = ["tiny-llama-hd64", "tiny-llama-hd128", "tiny-llama-3.2", "tiny-mixtral"]


def hf_logits(shape, sd, prompt, dtype):
  from transformers import LlamaConfig, LlamaForCausalLM
  hc = {k: v for k, v in hf_config(shape).items() if k not in ("architectures", "model_type", "torch_dtype")}
  if shape.n_experts:   # the reference has no MoE path: the oracle's MoE block restates HF Mixtral (fused 3-D expert params)
    from transformers import MixtralConfig, MixtralForCausalLM
    m = MixtralForCausalLM(MixtralConfig(**hc)).to(dtype).eval()
    conv = {k: v.to(dtype) for k, v in sd.items() if "block_sparse_moe" not in k}
    for l in range(shape.n_layers):
      p = f"model.layers.{l}."
      ex = lambda e, w: sd[p + f"block_sparse_moe.experts.{e}.{w}.weight"]
      conv[p + "mlp.gate.weight"] = sd[p + "block_sparse_moe.gate.weight"].to(dtype)
      conv[p + "mlp.experts.gate_up_proj"] = torch.stack([torch.cat([ex(e, "w1"), ex(e, "w3")], 0) for e in range(shape.n_experts)]).to(dtype)
      conv[p + "mlp.experts.down_proj"] = torch.stack([ex(e, "w2") for e in range(shape.n_experts)]).to(dtype)
    m.load_state_dict(conv, strict=True)
    with torch.no_grad():
      return m(torch.tensor(prompt).long()).logits[0].float().numpy()
  m = LlamaForCausalLM(LlamaConfig(**hc)).to(dtype).eval()
  m.load_state_dict({k: v.to(dtype) for k, v in sd.items()}, strict=not shape.tie)
  with torch.no_grad():
    return m(torch.tensor(prompt).long()).logits[0].float().numpy()


def part_a():
  for name in TINY:
    shape = PRESETS[name]
    sd = synth_state_dict(shape, 1234)
    cfg = model_config_from_hf(hf_config(shape))
    prompt = synth_prompt(shape.vocab, 24)
    eng = RefEngine(cfg, sd, RefShard(name, 0, shape.n_layers - 1, shape.n_layers))
    ids, logits, gaps = generate_ids([eng], prompt, 24)
    eng_ref = RefEngine(cfg, sd, RefShard(name, 0, shape.n_layers - 1, shape.n_layers))
    ids_ref, _, _ = generate_ids([eng_ref], prompt, 24, sampler="reference")
    # fp32 oracle vs fp32 HF (tight), bf16 oracle vs bf16 HF (loose) on the prefill's last position
    eng32 = RefEngine(cfg, sd, RefShard(name, 0, shape.n_layers - 1, shape.n_layers), dtype=torch.float32)
    _, l32, _ = generate_ids([eng32], prompt, 1)
    hf32 = hf_logits(shape, sd, prompt, torch.float32)[-1]
    hf16 = hf_logits(shape, sd, prompt, torch.bfloat16)[-1]
    np.savez_compressed(GOLD / f"oracle_{name}.npz", prompt=prompt, ids=ids, ids_reference_sampler=ids_ref,
                        logits=logits[:4].astype(np.float32), gaps=gaps.astype(np.float32),
                        hf_fp32_maxdiff=np.float32(np.abs(l32[0] - hf32).max()),
                        hf_bf16_maxdiff=np.float32(np.abs(logits[0] - hf16).max()),
                        logit_std=np.float32(logits[0].std()))
    print(name, "ids", ids[:8], "hf32 diff", np.abs(l32[0] - hf32).max(), "hfbf16 diff", np.abs(logits[0] - hf16).max(),
          "ref-sampler ids equal argmax ids:", np.array_equal(ids, ids_ref), "min gap", gaps.min())


def part_b():
  ref = Path("/root/reference")
  if not ref.exists():
    print("no /root/reference: skipping part B")
    return
  stub = types.ModuleType("scapy"); stub_all = types.ModuleType("scapy.all")
  stub_all.get_if_addr = lambda *a, **k: "127.0.0.1"; stub_all.get_if_list = lambda *a, **k: []
  sys.modules["scapy"] = stub; sys.modules["scapy.all"] = stub_all
  sys.path.insert(0, str(ref))
  cwd = os.getcwd()
  with tempfile.TemporaryDirectory() as td:
    os.chdir(td)   # device_capabilities.py writes a log file into the CWD
    sys.dont_write_bytecode = True
    from xotorch.topology.ring_memory_weighted_partitioning_strategy import RingMemoryWeightedPartitioningStrategy
    from xotorch.topology.partitioning_strategy import map_partitions_to_shards
    from xotorch.topology.topology import Topology
    from xotorch.topology.device_capabilities import DeviceCapabilities, DeviceFlops
    cases = {
      "ref_test_3nodes": {"node1": 3000, "node2": 1000, "node3": 6000},
      "ref_test_rounding": {"node1": 128 * 1024, "node2": 192 * 1024, "node3": 128 * 1024},
      "b200_x2": {f"gpu{i}": 183359 for i in range(2)},
      "b200_x4": {f"gpu{i}": 183359 for i in range(4)},
      "b200_x8": {f"gpu{i}": 183359 for i in range(8)},
      "b200_x8_uneven": {f"gpu{i}": 183359 - 9000 * (i % 3) for i in range(8)},
      "three_uneven": {"a": 100, "b": 101, "c": 55},
    }
    out = {}
    for cname, mems in cases.items():
      topo = Topology()
      for nid, mem in mems.items():
        topo.update_node(nid, DeviceCapabilities(model="x", chip="x", memory=mem, flops=DeviceFlops(fp32=0, fp16=0, int8=0)))
      parts = RingMemoryWeightedPartitioningStrategy().partition(topo)
      entry = {"memories": mems, "partitions": [[p.node_id, p.start, p.end] for p in parts], "shards": {}}
      for nl in (4, 16, 32, 80):
        shards = map_partitions_to_shards(parts, nl, "m")
        entry["shards"][str(nl)] = [[s.start_layer, s.end_layer] for s in shards]
      out[cname] = entry
    # the reference's own literal golden vectors (test_map_partitions.py:10-37)
    from xotorch.topology.partitioning_strategy import Partition
    lit = []
    for fr in ([0.0, 0.42857, 0.71428, 0.99999], [0.0, 0.1, 0.2, 1.0]):
      parts = [Partition(f"n{i}", fr[i], fr[i + 1]) for i in range(3)]
      lit.append({"fractions": fr, "shards": [[s.start_layer, s.end_layer] for s in map_partitions_to_shards(parts, 32, "m")]})
    out["_literal_map_partitions_32"] = lit
    os.chdir(cwd)
  (GOLD / "partitioning_from_reference.json").write_text(json.dumps(out, indent=1))
  print("partitioning fixtures:", list(out.keys()))


if __name__ == "__main__":
  GOLD.mkdir(parents=True, exist_ok=True)
  part_a()
  part_b()
