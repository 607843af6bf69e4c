"""Tensor-core (tcgen05/TMEM/TMA) prefill path — GPU tests through the C ABI.
1. the GEMM kernel alone against numpy (bf16 inputs, fp32 accumulation, one bf16 rounding);
2. prefill through the GEMM path vs the CPU oracle and vs the decode-kernel path (both within the bf16 tolerance of
   tests/test_gpu_parity.py), including prompts longer than one 128-token tile and longer than one 1024-token chunk and sharded == unsharded bit-exactness."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from xotorch_b200.synth import PRESETS, hf_config, synth_prompt
from tests.test_gpu_parity import _oracle, _native, _api_generate, _oracle_with_routing, ATOL, RTOL


def _bf16_bits(x: torch.Tensor) -> np.ndarray:
  return x.to(torch.bfloat16).contiguous().view(torch.int16).numpy().view(np.uint16)


def _from_bits(b: np.ndarray) -> np.ndarray:
  return torch.from_numpy(b.view(np.int16).copy()).view(torch.bfloat16).float().numpy()


@pytest.mark.parametrize("S,N,K,tk", [(5, 64, 128, 64), (128, 768, 512, 512), (77, 256, 4096, 2048), (33, 16384, 1024, 512),
                                      (128, 4096, 2048, 2048)])
def test_tcgen05_gemm_matches_numpy(S, N, K, tk):
  from xotorch_b200.native import NativeShard
  e = NativeShard(hf_config(PRESETS["tiny-llama-hd64"]), 0)
  g = torch.Generator().manual_seed(S * 1000 + N)
  a = (torch.randn(S, K, generator=g) * 0.5).to(torch.bfloat16)
  w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16)
  out = _from_bits(e.debug_gemm(_bf16_bits(a), _bf16_bits(w), tk))
  ref = a.float().numpy().astype(np.float64) @ w.float().numpy().astype(np.float64).T
  ref_bf = torch.tensor(ref, dtype=torch.float32).to(torch.bfloat16).float().numpy()
  # fp32 accumulation in a different order + one bf16 rounding: at most 1 bf16 ulp away from the rounded exact result
  ulp = np.maximum(np.abs(ref_bf), 1e-3) * 2.0 ** -6   # one ulp of the upper binade at a binade boundary
  assert np.all(np.abs(out - ref_bf) <= ulp), f"max err {np.abs(out - ref_bf).max()} (|ref| max {np.abs(ref).max()})"
  assert np.mean(out == ref_bf) > 0.97


@pytest.mark.parametrize("shape_name,n_prompt", [("tiny-llama-hd64", 40), ("tiny-llama-hd128", 131), ("tiny-llama-3.2", 200),
                                                 ("tiny-llama-hd128-f16", 140), ("tiny-llama-hd64", 1100), ("tiny-llama-hd128", 900)])
def test_tc_prefill_matches_oracle_and_decode_kernel_path(shape_name, n_prompt):
  shape = PRESETS[shape_name]
  full = [(0, shape.n_layers - 1)]
  n_gen = 6
  prompt, (o_ids, o_logits, gaps) = _oracle(shape_name, full, n_prompt, n_gen)
  res = {}
  for mode in (1, 2):
    eng = _native(shape_name, full)
    eng[0].set_prefill_mode(mode)
    res[mode] = _api_generate(eng, prompt, n_gen, force_ids=o_ids)
  for mode, (ids, logits) in res.items():
    err = np.abs(logits - o_logits)
    assert (err <= ATOL + RTOL * np.abs(o_logits)).all(), f"mode {mode}: max err {err.max():.4f}"
  # the two prefill paths are two valid bf16 evaluations of the same model
  d = np.abs(res[1][1] - res[2][1])
  assert (d <= ATOL + RTOL * np.abs(o_logits)).all(), f"TC vs decode-kernel prefill: max diff {d.max():.4f}"


def test_tc_prefill_sharded_equals_unsharded_bit_exact():
  shape_name = "tiny-llama-hd64"
  shape = PRESETS[shape_name]
  prompt = synth_prompt(shape.vocab, 150)           # two tiles on every shard
  a = _native(shape_name, [(0, shape.n_layers - 1)])
  b = _native(shape_name, [(0, 0), (1, 2), (3, 3)])
  for e in a + b:
    e.set_prefill_mode(2)
  ids_f, logits_f = _api_generate(a, prompt, 5)
  ids_s, logits_s = _api_generate(b, prompt, 5)
  assert np.array_equal(ids_f, ids_s) and np.array_equal(logits_f, logits_s)


def test_tc_prefill_real_dims_vs_oracle():
  import dataclasses
  from oracle.ref_engine import RefEngine, RefShard, model_config_from_hf, generate_ids
  from oracle.synth_fast import synth_state_dict_fast
  from xotorch_b200.native import NativeShard
  shape = dataclasses.replace(PRESETS["llama-3-8b"], n_layers=2, name="llama-3-8b-cut2")
  sd = synth_state_dict_fast(shape, 1234)
  ref = RefEngine(model_config_from_hf(hf_config(shape)), sd, RefShard(shape.name, 0, 1, 2))
  prompt = synth_prompt(shape.vocab, 48)
  o_ids, o_logits, gaps = generate_ids([ref], prompt, 3)
  e = NativeShard(hf_config(shape), 0)
  e.set_shard(0, 1); e.init_synthetic(1234); e.set_prefill_mode(2)
  ids, logits = _api_generate([e], prompt, 3, force_ids=o_ids)
  err = np.abs(logits - o_logits); std = float(o_logits.std())
  assert float(np.sqrt((err ** 2).mean())) <= 0.02 * std and err.max() <= 0.10 * std, (err.max(), std)


@pytest.mark.parametrize("shape_name,n_prompt", [("tiny-mixtral", 40), ("tiny-mixtral", 300), ("tiny-mixtral-hd128", 140)])
def test_tc_prefill_sparse_moe_matches_oracle(shape_name, n_prompt):
  """Sparse-MoE prompt on the tensor cores (router -> per-expert row blocks -> tcgen05 GEMMs per expert -> combine) and
  through the decode kernel: each against the oracle re-run with that path's expert choices (near-ties asserted)."""
  from xotorch_b200.synth import synth_state_dict
  shape = PRESETS[shape_name]
  full = [(0, shape.n_layers - 1)]
  n_gen = 4
  prompt, (o_ids, _, _) = _oracle(shape_name, full, n_prompt, n_gen)
  sd = synth_state_dict(shape, 1234)
  for mode in (2, 1):
    eng = _native(shape_name, full)
    eng[0].set_prefill_mode(mode)
    ids, logits = _api_generate(eng, prompt, n_gen, force_ids=o_ids)
    o_logits, gaps = _oracle_with_routing(shape, sd, prompt, n_gen, o_ids, eng[0].debug_routing(n_prompt + n_gen - 1))
    err = np.abs(logits - o_logits)
    assert (err <= ATOL + RTOL * np.abs(o_logits)).all(), f"mode {mode}: max err {err.max():.4f}"

