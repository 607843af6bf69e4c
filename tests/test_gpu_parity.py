"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the
same seeded inputs.  Tolerances are stated where used.

Reference contract restated here: "An inference engine should work the same for any number of Shards"
(/root/reference/xotorch/inference/test_inference_engine.py:11,43-44 -> np.array_equal(full, sharded))."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from xotorch_b200.synth import PRESETS, hf_config, synth_state_dict, synth_prompt, write_hf_dir

# Logit tolerance (bf16 model, fp32 accumulation, different summation order than the CPU oracle):
# |delta| <= ATOL + RTOL*|logit|; bf16 has 8 significant bits (ulp = 2^-8 relative), a few ulps accumulate over depth.
ATOL, RTOL = 0.04, 0.02


def _oracle(shape_name, shards, n_prompt, n_gen, force_ids=None, seed=1234):
  from oracle.ref_engine import RefEngine, RefShard, model_config_from_hf, generate_ids
  shape = PRESETS[shape_name]
  sd = synth_state_dict(shape, seed)
  cfg = model_config_from_hf(hf_config(shape))
  engines = [RefEngine(cfg, sd, RefShard(shape_name, a, b, shape.n_layers)) for a, b in shards]
  prompt = synth_prompt(shape.vocab, n_prompt)
  return prompt, generate_ids(engines, prompt, n_gen, force_ids=force_ids)


# Sparse-MoE routing is a discrete choice on bf16 router logits (std ~0.45, ulp 2^-9..2^-7): two correct bf16 evaluations
# pick different experts when the 2nd and 3rd logit are within an ulp or two.  The oracle is therefore re-run with the
# CUDA path's expert choices imposed; it reports every decision where its own choice differs and how far (router-logit
# units) the imposed pair is from its own: that must be a near-tie, and rare.
ROUTE_TIE = 0.02


def _oracle_with_routing(shape, sd, prompt, n_gen, force_ids, routes):
  from oracle.ref_engine import RefEngine, RefShard, model_config_from_hf, generate_ids
  nl = shape.n_layers
  ref = RefEngine(model_config_from_hf(hf_config(shape)), sd, RefShard(shape.name, 0, nl - 1, nl))
  ref.model.forced_routes = {(l, p): (int(routes[p, l, 0]), int(routes[p, l, 1]))
                             for p in range(routes.shape[0]) for l in range(nl) if routes[p, l, 0] >= 0}
  ref.model.route_log = []
  ids, logits, gaps = generate_ids([ref], prompt, n_gen, force_ids=force_ids)
  log = ref.model.route_log
  assert len(log) == nl * (prompt.shape[1] + n_gen - 1) == len(ref.model.forced_routes)
  dis = [r for r in log if r[2] != r[3]]
  assert all(r[4] <= ROUTE_TIE for r in dis), f"routing differs from the oracle away from a near-tie: {[r for r in dis if r[4] > ROUTE_TIE]}"
  assert len(dis) <= max(2, 0.06 * len(log)), f"{len(dis)} of {len(log)} routing decisions differ"
  return logits, gaps


def _native(shape_name, shards, seed=1234, model_dir=None):
  from xotorch_b200.native import NativeShard
  shape = PRESETS[shape_name]
  out = []
  for a, b in shards:
    e = NativeShard(hf_config(shape), device=0)
    if model_dir is None:
      e.set_shard(a, b)
      e.init_synthetic(seed)
    else:
      e.load_dir(model_dir, a, b)
    out.append(e)
  return out


def _api_generate(engines, prompt, n_gen, force_ids=None, max_len=None):
  """Drive the host-buffer API exactly like Node does: ids -> shard0 -> ... -> logits -> argmax -> ids."""
  max_len = max_len or (prompt.shape[1] + n_gen + 1)
  for e in engines:
    e.begin_request(max_len)
  x, pos = prompt.astype(np.int32), 0
  ids, logits_all = [], []
  for step in range(n_gen):
    data = x
    for e in engines:
      data = e.forward(data, pos)
    logits_all.append(data[0].copy())
    tok = int(np.argmax(data[0]))
    assert tok == engines[-1].last_token(), "device argmax != host lowest-index argmax"
    if force_ids is not None:
      tok = int(force_ids[step])
    ids.append(tok)
    pos += x.shape[1]
    x = np.array([[tok]], dtype=np.int32)
  return np.array(ids, dtype=np.int32), np.stack(logits_all)


@pytest.mark.parametrize("shape_name", ["tiny-llama-hd64", "tiny-llama-hd128", "tiny-llama-3.2", "tiny-llama-hd64-f16",
                                        "tiny-llama-hd128-f16", "tiny-mixtral", "tiny-mixtral-hd128", "tiny-mixtral-f16"])
def test_logits_match_oracle_teacher_forced(shape_name):
  shape = PRESETS[shape_name]
  full = [(0, shape.n_layers - 1)]
  prompt, (o_ids, o_logits, gaps) = _oracle(shape_name, full, 24, 24)
  engines = _native(shape_name, full)
  ids, logits = _api_generate(engines, prompt, 24, force_ids=o_ids)
  if shape.n_experts:
    o_logits, gaps = _oracle_with_routing(shape, synth_state_dict(shape, 1234), prompt, 24, o_ids, engines[0].debug_routing(24 + 23))
  err = np.abs(logits - o_logits)
  tol = ATOL + RTOL * np.abs(o_logits)
  assert (err <= tol).all(), f"max err {err.max():.4f} (logit std {o_logits.std():.3f})"
  # wherever the oracle's top-2 gap is above the tolerance, the greedy choice must agree
  decided = gaps > 2 * (ATOL + RTOL * np.abs(o_logits).max())
  assert (np.argmax(logits, -1)[decided] == np.argmax(o_logits, -1)[decided]).all()


def test_greedy_ids_match_oracle():
  """Free-running greedy ids: equal to the oracle's up to the first near-tie (reported, not hidden)."""
  shape_name = "tiny-llama-hd64"
  shape = PRESETS[shape_name]
  full = [(0, shape.n_layers - 1)]
  prompt, (o_ids, o_logits, gaps) = _oracle(shape_name, full, 16, 32)
  engines = _native(shape_name, full)
  ids, logits = _api_generate(engines, prompt, 32)
  if not np.array_equal(ids, o_ids):
    # a divergence is only legitimate at a (near-)tie of the oracle's two best logits, and only towards the runner-up: the
    # 16-bit logits of the oracle can tie EXACTLY (gap 0.0), and which of the two wins then hangs on the fp32 summation order
    # (the tensor-core GEMV sums in a different order than the oracle's matmul).  Reported, not hidden:
    first = int(np.argmax(ids != o_ids))
    assert gaps[first] <= 2 * ATOL, f"ids diverge at step {first} although the oracle top-2 gap is {gaps[first]:.4f}"
    top2 = np.argsort(o_logits[first])[-2:]
    assert ids[first] in top2, f"step {first}: picked {ids[first]}, the oracle's two best are {top2.tolist()}"
    print(f"greedy ids agree for {first} steps, then diverge on a near-tie (oracle top-2 gap {gaps[first]:.4f}) towards the runner-up")


@pytest.mark.parametrize("shape_name,split", [("tiny-llama-hd64", [(0, 1), (2, 3)]), ("tiny-llama-hd64", [(0, 0), (1, 2), (3, 3)]),
                                              ("tiny-llama-3.2", [(0, 2), (3, 3)]), ("tiny-llama-hd64-f16", [(0, 1), (2, 3)]),
                                              ("tiny-mixtral", [(0, 0), (1, 2), (3, 3)])])
def test_sharded_equals_unsharded_bit_exact(shape_name, split):
  shape = PRESETS[shape_name]
  prompt = synth_prompt(shape.vocab, 12)
  ids_f, logits_f = _api_generate(_native(shape_name, [(0, shape.n_layers - 1)]), prompt, 8)
  ids_s, logits_s = _api_generate(_native(shape_name, split), prompt, 8)
  assert np.array_equal(ids_f, ids_s)
  assert np.array_equal(logits_f, logits_s)


@pytest.mark.parametrize("shape_name", ["tiny-llama-hd64", "tiny-llama-hd64-f16", "tiny-mixtral"])
def test_safetensors_loader_equals_device_synth(tmp_path, shape_name):
  shape = PRESETS[shape_name]
  d = write_hf_dir(tmp_path / "m", shape, 1234)
  prompt = synth_prompt(shape.vocab, 9)
  a = _api_generate(_native(shape_name, [(0, shape.n_layers - 1)]), prompt, 4)
  b = _api_generate(_native(shape_name, [(0, shape.n_layers - 1)], model_dir=d), prompt, 4)
  assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_strict_loader_reports_missing(tmp_path):
  from xotorch_b200.native import NativeShard
  from xotorch_b200._lib import XotB200Error
  shape = PRESETS["tiny-llama-hd64"]
  e = NativeShard(hf_config(shape), 0)
  e.set_shard(0, 1)
  with pytest.raises(XotB200Error, match="never loaded"):
    e.check_loaded()


@pytest.mark.parametrize("split,shape_name", [([(0, 3)], "tiny-llama-hd64"), ([(0, 1), (2, 3)], "tiny-llama-hd64"),
                                              ([(0, 0), (1, 2), (3, 3)], "tiny-llama-hd64"), ([(0, 1), (2, 3)], "tiny-llama-hd64-f16"),
                                              ([(0, 3)], "tiny-mixtral"), ([(0, 1), (2, 3)], "tiny-mixtral")])
def test_ring_mode_equals_api_mode(split, shape_name):
  """Device-resident ring (persistent kernel, inbox flags, on-device argmax feedback) == step-by-step host API."""
  shape = PRESETS[shape_name]
  prompt = synth_prompt(shape.vocab, 10)
  n_gen = 20
  ids_api, _ = _api_generate(_native(shape_name, [(0, shape.n_layers - 1)]), prompt, n_gen)
  engines = _native(shape_name, split)
  import torch
  n_sm = torch.cuda.get_device_properties(0).multi_processor_count
  for i, e in enumerate(engines):
    if len(engines) > 1:          # persistent kernels of several engines on ONE GPU must be co-resident
      e.set_cta_limit(n_sm // len(engines))
    e.link_next_local(engines[(i + 1) % len(engines)])
    e.begin_request(prompt.shape[1] + n_gen + 1)
  data = prompt.astype(np.int32)
  for e in engines:
    data = e.forward(data, 0)
  first_tok = engines[-1].last_token()
  assert first_tok == ids_api[0]
  step0 = 7
  engines[-1].post_to_next(step0)
  for e in engines:
    e.ring_launch(n_gen - 1, step0, prompt.shape[1])
  toks = None
  for e in engines:
    t, ms = e.ring_wait(n_gen - 1)
    toks = t if t is not None else toks
  assert np.array_equal(np.concatenate([[first_tok], toks]), ids_api)
  # a second run on the same engines (monotonic step numbers) gives the same ids
  for e in engines:
    e.begin_request(prompt.shape[1] + n_gen + 1)
  data = prompt.astype(np.int32)
  for e in engines:
    data = e.forward(data, 0)
  step0 += n_gen - 1
  engines[-1].post_to_next(step0)
  for e in engines:
    e.ring_launch(n_gen - 1, step0, prompt.shape[1])
  for e in engines:
    t, ms = e.ring_wait(n_gen - 1)
    toks = t if t is not None else toks
  assert np.array_equal(np.concatenate([[first_tok], toks]), ids_api)


# ------------------------------------------------------------------------------------------------------------------
# Real Llama dimensions (the TK=2048 kernel instantiation that bench.py measures), layer count cut so the CPU oracle
# finishes in seconds: same per-layer arithmetic, same head.
# ------------------------------------------------------------------------------------------------------------------
def _cut(shape_name, n_layers):
  import dataclasses
  # the 70B card is fp16 upstream; the engine is bf16-only so far: same dimensions, bf16 weights
  return dataclasses.replace(PRESETS[shape_name], n_layers=n_layers, name=f"{shape_name}-cut{n_layers}", dtype="bfloat16")


@pytest.mark.parametrize("shape_name,n_layers,n_prompt,n_gen", [("llama-3-8b", 2, 12, 6), ("llama-3.2-1b", 3, 12, 6),
                                                                 ("llama-3-70b", 2, 12, 4), ("mixtral-8x7b", 2, 10, 4)])
def test_real_dims_match_oracle_teacher_forced(shape_name, n_layers, n_prompt, n_gen):
  from oracle.ref_engine import RefEngine, RefShard, model_config_from_hf, generate_ids
  from oracle.synth_fast import synth_state_dict_fast
  from xotorch_b200.native import NativeShard
  shape = _cut(shape_name, n_layers)
  sd = synth_state_dict_fast(shape, 1234)
  cfg = model_config_from_hf(hf_config(shape))
  ref = RefEngine(cfg, sd, RefShard(shape.name, 0, n_layers - 1, n_layers))
  prompt = synth_prompt(shape.vocab, n_prompt)
  o_ids, o_logits, gaps = generate_ids([ref], prompt, n_gen)
  e = NativeShard(hf_config(shape), device=0)
  e.set_shard(0, n_layers - 1)
  e.init_synthetic(1234)
  ids, logits = _api_generate([e], prompt, n_gen, force_ids=o_ids)
  if shape.n_experts:
    o_logits, gaps = _oracle_with_routing(shape, sd, prompt, n_gen, o_ids, e.debug_routing(n_prompt + n_gen - 1))
  # Tolerance calibrated on CPU with two independent, equally valid bf16 implementations of the same model (this oracle
  # vs HF LlamaForCausalLM, same weights, 128k-entry logit vectors): they differ by rms 0.0136*std / max 0.070*std
  # (8B dims) and rms 0.0139*std / max 0.067*std (1B dims).  The CUDA path must be at least that close to the oracle:
  err = np.abs(logits - o_logits)
  std = float(o_logits.std())
  rms = float(np.sqrt((err ** 2).mean()))
  assert rms <= 0.02 * std, f"rms err {rms:.4f} vs logit std {std:.3f}"
  assert err.max() <= 0.10 * std, f"max err {err.max():.4f} vs logit std {std:.3f}"
  decided = gaps > 0.2 * std
  assert (np.argmax(logits, -1)[decided] == np.argmax(o_logits, -1)[decided]).all()
  # device ring == host API at these dimensions too (bit-exact)
  e.link_next_local(e)
  e.begin_request(n_prompt + n_gen + 1)
  lg = e.forward(prompt, 0)
  first = e.last_token()
  e.post_to_next(0)
  e.ring_launch(n_gen - 1, 0, n_prompt)
  toks, _ = e.ring_wait(n_gen - 1)
  ids_free, _ = _api_generate([e], prompt, n_gen)
  assert np.array_equal(np.concatenate([[first], toks]), ids_free)


def test_two_shards_real_dims_bit_exact_on_one_gpu():
  import torch
  from xotorch_b200.native import NativeShard
  shape = _cut("llama-3-8b", 2)
  prompt = synth_prompt(shape.vocab, 6)
  full = NativeShard(hf_config(shape), 0); full.set_shard(0, 1); full.init_synthetic(1234)
  ids_f, logits_f = _api_generate([full], prompt, 4)
  full.close()
  a = NativeShard(hf_config(shape), 0); a.set_shard(0, 0); a.init_synthetic(1234)
  b = NativeShard(hf_config(shape), 0); b.set_shard(1, 1); b.init_synthetic(1234)
  ids_s, logits_s = _api_generate([a, b], prompt, 4)
  assert np.array_equal(ids_f, ids_s) and np.array_equal(logits_f, logits_s)


# ------------------------------------------------------------------------------------------------------------------
# error behaviour of the boundary (the reference swallows OOM and returns None, App. B; we raise with a message)
# ------------------------------------------------------------------------------------------------------------------
def test_error_paths(tmp_path):
  from xotorch_b200.native import NativeShard
  from xotorch_b200._lib import XotB200Error
  shape = PRESETS["tiny-llama-hd64"]
  e = NativeShard(hf_config(shape), 0)
  with pytest.raises(XotB200Error, match="bad layer range"):
    e.set_shard(2, 9)
  e.set_shard(0, 1)
  e.init_synthetic(1)
  with pytest.raises(XotB200Error, match="begin_request"):
    e.forward(np.zeros((1, 3), np.int32), 0)
  e.begin_request(8)
  with pytest.raises(XotB200Error, match="KV cache overflow"):
    e.forward(np.zeros((1, 9), np.int32), 0)
  with pytest.raises(XotB200Error, match="expects"):
    e.forward(np.zeros((1, 2, shape.dim), np.float32), 0)            # first shard wants token ids
  with pytest.raises(XotB200Error, match="out of range"):
    e.forward(np.full((1, 2), shape.vocab + 5, np.int32), 0)
  with pytest.raises(XotB200Error, match="max_total_len"):
    e.begin_request(shape.max_pos + 1)
  with pytest.raises(XotB200Error, match="ring not linked"):
    e.ring_launch(1, 0, 0)
  with pytest.raises(XotB200Error, match="cannot open directory|no \\*.safetensors"):
    e.load_dir(tmp_path / "nope", 0, 1)
  # wrong-shaped tensor is rejected by name
  with pytest.raises(XotB200Error, match="shape mismatch"):
    e.load_tensor("model.layers.0.mlp.down_proj.weight", np.zeros((4, 4), np.uint16))
  assert e.load_tensor("model.layers.3.mlp.down_proj.weight", np.zeros((4, 4), np.uint16)) is False   # other shard: ignored


@pytest.mark.parametrize("split,B", [([(0, 3)], 3), ([(0, 1), (2, 3)], 2), ([(0, 0), (1, 2), (3, 3)], 4)])
def test_requests_in_flight_equal_independent_runs(split, B):
  """B requests interleaved step by step through the ring (own KV cache each) == B independent single-request runs."""
  import torch
  from xotorch_b200.ring import LocalRing
  shape_name = "tiny-llama-hd64"
  shape = PRESETS[shape_name]
  n_gen = 12
  g = torch.Generator().manual_seed(7)
  prompts = torch.randint(0, shape.vocab, (B, 11), generator=g).to(torch.int32).numpy()
  single = LocalRing(_native(shape_name, [(0, shape.n_layers - 1)]))
  want = np.stack([single.generate(prompts[r:r + 1], n_gen)[0] for r in range(B)])
  ring = LocalRing(_native(shape_name, split))
  got, ms = ring.generate_batch(prompts, n_gen)
  assert np.array_equal(got, want), (got, want)
  got2, _ = ring.generate_batch(prompts[::-1].copy(), n_gen)     # a second batch on the same engines
  assert np.array_equal(got2, want[::-1])
  one, _ = ring.generate(prompts[0:1], n_gen)                    # and back to a single request
  assert np.array_equal(one, want[0])


def test_engine_api_interleaved_request_ids(tmp_path):
  """B200ShardInferenceEngine(max_concurrent_requests=2): two request ids interleaved token by token through infer_tensor
  give exactly the logits of the same requests run one after the other on a single-request engine."""
  import asyncio
  from xotorch_b200.inference.b200_engine import B200ShardInferenceEngine, LocalDirDownloader
  from xotorch_b200.inference.shard import Shard
  name = "tiny-llama-hd64"
  shape = PRESETS[name]
  d = write_hf_dir(tmp_path / name, shape, 1234)
  sh = Shard(name, 0, shape.n_layers - 1, shape.n_layers)
  pa, pb = synth_prompt(shape.vocab, 9, seed=1), synth_prompt(shape.vocab, 14, seed=2)

  async def one(engine, rid, prompt, n):
    out, st = await engine.infer_tensor(rid, sh, prompt, None)
    logs = [out[0].copy()]
    for _ in range(n - 1):
      tok = np.argmax(out[0]).reshape(1, 1).astype(np.int64)
      out, st = await engine.infer_tensor(rid, sh, tok, st)
      logs.append(out[0].copy())
    return np.stack(logs)

  async def run():
    single = B200ShardInferenceEngine(LocalDirDownloader({name: d}), device=0)
    want_a, want_b = await one(single, "a", pa, 6), await one(single, "b", pb, 6)
    multi = B200ShardInferenceEngine(LocalDirDownloader({name: d}), device=0, max_concurrent_requests=2, max_total_len=64)
    oa, sa = await multi.infer_tensor("a", sh, pa, None)
    ob, sb = await multi.infer_tensor("b", sh, pb, None)
    got_a, got_b = [oa[0].copy()], [ob[0].copy()]
    for _ in range(5):                                     # strictly alternating decode steps of the two requests
      oa, sa = await multi.infer_tensor("a", sh, np.argmax(oa[0]).reshape(1, 1).astype(np.int64), sa)
      ob, sb = await multi.infer_tensor("b", sh, np.argmax(ob[0]).reshape(1, 1).astype(np.int64), sb)
      got_a.append(oa[0].copy()); got_b.append(ob[0].copy())
    assert np.array_equal(np.stack(got_a), want_a) and np.array_equal(np.stack(got_b), want_b)
  asyncio.run(run())


@pytest.mark.parametrize("shape_name", ["tiny-llama-hd64", "tiny-llama-hd64-f16"])
def test_forward_device_equals_host_buffer_forward(shape_name):
  """xotb200_forward_device (hidden state handed over in device memory, 16-bit) == xotb200_forward (fp32 host arrays)."""
  import torch
  shape = PRESETS[shape_name]
  nl = shape.n_layers
  prompt = synth_prompt(shape.vocab, 150)
  a, b = _native(shape_name, [(0, 1), (2, nl - 1)])
  for e in (a, b):
    e.begin_request(200)
  want = b.forward(a.forward(prompt, 0), 0)
  want2 = b.forward(a.forward(prompt[:, :1], 150), 150)
  a2, b2 = _native(shape_name, [(0, 1), (2, nl - 1)])
  for e in (a2, b2):
    e.begin_request(200)
  dt = torch.float16 if shape.dtype == "float16" else torch.bfloat16
  h = torch.empty((150, shape.dim), dtype=dt, device="cuda:0")
  assert a2.forward_device(prompt, None, 150, 0, h.data_ptr()) is None
  got = b2.forward_device(None, h.data_ptr(), 150, 0, None)
  h1 = torch.empty((1, shape.dim), dtype=dt, device="cuda:0")
  a2.forward_device(prompt[:, :1], None, 1, 150, h1.data_ptr())
  got2 = b2.forward_device(None, h1.data_ptr(), 1, 150, None)
  assert np.array_equal(got, want) and np.array_equal(got2, want2)



def test_streaming_loader_chunks_equal_device_synth_at_real_dims(tmp_path):
  """The streaming shard loader (mmap -> pinned double buffer -> DMA -> re-layout kernel, 64 MiB chunks) on tensors that need
  several chunks: Llama-3-8B dimensions, 1 layer + embedding + head (2.5 GB: embed/head 1 GB = 16 plain / K-panel chunks,
  gate/up 117 MB = 2 chunks into interleaved rows, q/k 32 MB with the head-wise _permute).  Bit-exact against the device
  generator, and the load rate is reported next to the round-1 per-tensor path (XOTB200_LOAD_STREAM=0)."""
  import dataclasses, json, os, time
  from xotorch_b200.native import NativeShard
  shape = dataclasses.replace(PRESETS["llama-3-8b"], n_layers=1, name="llama-3-8b-cut1")
  d = write_hf_dir(tmp_path / "m", shape, 1234, n_files=2)
  prompt = synth_prompt(shape.vocab, 9)
  a = NativeShard(hf_config(shape), 0); a.set_shard(0, 0); a.init_synthetic(1234)
  want = _api_generate([a], prompt, 3)
  a.close()
  rates = {}
  for mode in ("1", "0"):
    os.environ["XOTB200_LOAD_STREAM"] = mode
    try:
      b = NativeShard(hf_config(shape), 0)
      t0 = time.time()
      b.load_dir(d, 0, 0)
      st = b.load_stats()
      rates["streaming" if mode == "1" else "per-tensor (round 1)"] = {"GBps": round(st["GBps"], 2), "seconds": round(time.time() - t0, 2),
                                                                          "bytes": st["bytes"], "memcpy_seconds": round(st["memcpy_seconds"], 2)}
      got = _api_generate([b], prompt, 3)
      b.close()
    finally:
      os.environ.pop("XOTB200_LOAD_STREAM", None)
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1]), f"loader mode {mode}"
  print(f"[loader] {json.dumps(rates)}")
  out = Path(__file__).resolve().parent.parent / "gpurun_out"
  if out.is_dir():
    (out / "loader_rates.json").write_text(json.dumps(rates, indent=1))


def _rewrite_safetensors(path, mutate):
  """re-serialise one safetensors file after `mutate(header_dict, data_bytes) -> (header_dict, data_bytes)`"""
  import json, struct
  raw = Path(path).read_bytes()
  hlen = struct.unpack("<Q", raw[:8])[0]
  hdr, data = json.loads(raw[8:8 + hlen]), raw[8 + hlen:]
  hdr, data = mutate(hdr, data)
  hb = json.dumps(hdr, separators=(",", ":")).encode()
  Path(path).write_bytes(struct.pack("<Q", len(hb)) + hb + data)


def test_strict_loader_rejects_malformed_and_foreign_tensors(tmp_path):
  """a biased-attention checkpoint (tensor of an owned layer the engine has no use for) and a file whose
  data_offsets do not match the declared shape / the file size must be refused, not silently loaded."""
  import shutil
  from xotorch_b200.native import NativeShard
  from xotorch_b200._lib import XotB200Error
  shape = PRESETS["tiny-llama-hd64"]
  good = write_hf_dir(tmp_path / "good", shape, 1234, n_files=1)
  f = next(Path(good).glob("*.safetensors")).name

  def variant(name, mutate):
    d = tmp_path / name
    shutil.copytree(good, d)
    _rewrite_safetensors(d / f, mutate)
    return d

  def add_bias(h, data):
    h["model.layers.1.self_attn.q_proj.bias"] = {"dtype": "BF16", "shape": [shape.n_heads * shape.head_dim], "data_offsets": [len(data), len(data) + 2 * shape.n_heads * shape.head_dim]}
    return h, data + b"\0" * (2 * shape.n_heads * shape.head_dim)

  def short_offsets(h, data):
    e = h["model.layers.0.mlp.down_proj.weight"]
    e["data_offsets"] = [e["data_offsets"][0], e["data_offsets"][1] - 64]
    return h, data

  def past_eof(h, data):
    e = h["model.layers.0.mlp.down_proj.weight"]
    n = e["data_offsets"][1] - e["data_offsets"][0]
    e["data_offsets"] = [len(data) - 16, len(data) - 16 + n]
    return h, data

  e = NativeShard(hf_config(shape), 0)
  e.load_dir(good, 0, shape.n_layers - 1)                     # the untouched directory loads
  with pytest.raises(XotB200Error, match="no use for it"):
    e.load_dir(variant("bias", add_bias), 0, shape.n_layers - 1)
  e.load_dir(variant("bias_elsewhere", add_bias), 2, 3)       # ... but a shard that does not own layer 1 is not concerned
  for name, mut in (("short", short_offsets), ("eof", past_eof)):
    with pytest.raises(XotB200Error, match="data_offsets"):
      e.load_dir(variant(name, mut), 0, shape.n_layers - 1)
