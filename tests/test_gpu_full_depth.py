"""GPU parity at the BASELINE.json configurations themselves (-m gpu): FULL depth, the benchmark's own prompt/generation
lengths, through the C ABI, against the CPU oracle on the same seeded inputs.

The tiny-shape and 2-3-layer-cut tests (test_gpu_parity.py) pin the arithmetic; these pin what bf16 error does over
32 / 16 real layers and over a 512-token greedy run, i.e. exactly what bench.py measures:

  * config 2/3  Llama-3-8B  bf16, 32 layers, 128-token prompt, teacher-forced decode steps          (logits + argmax rule)
  * config 1    Llama-3.2-1B bf16, 16 layers, 128-token prompt / 512 generated, free-running + teacher-forced (ids rule)
  * config 4    Llama-3-70B  fp16 dimensions, 2-layer cut                                            (fp16 at real dims)

Tolerances (stated where used) are calibrated against the distance between two equally valid bf16 evaluations of the same
full-depth model on the CPU — this oracle vs HF LlamaForCausalLM on identical weights, `python -m oracle.calibrate_depth`:
    llama-3-8b   32 layers, 128 + 16 steps: rms 0.069 std, max 0.40 std, argmax equal at 15 of 17 steps
    llama-3.2-1b 16 layers, 128 + 63 steps: rms 0.031 std, max 0.17 std, argmax equal at 61 of 64 steps
(bf16 error grows with depth: the 2-3-layer cuts of test_gpu_parity.py sit at rms 0.014 / max 0.07).  The CUDA path has to be
at least that close to the oracle; what it actually reaches is written to the report.

Greedy ids: the north star asks for bit-exact ids on the same prompt.  Two correct bf16 evaluations cannot promise that on
random-weight models (top-2 logit gaps below one bf16 ulp occur every few dozen steps), so the rule asserted here is the
strongest one that is implementation-independent: at EVERY teacher-forced step whose oracle top-2 gap exceeds the logit
tolerance the CUDA argmax equals the oracle's; the number of near-tie steps is reported; free-running ids are identical up
to the first near-tie step.  Reports go to gpurun_out/parity_full_depth.json when that directory exists."""
import json
import os
import time
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from xotorch_b200.synth import PRESETS, hf_config, synth_prompt

ROOT = Path(__file__).resolve().parent.parent


def _report(name: str, rec: dict):
  print(f"[parity] {name}: {json.dumps(rec)}")
  d = ROOT / "gpurun_out"
  if d.is_dir():
    p = d / "parity_full_depth.json"
    try:
      cur = json.loads(p.read_text()) if p.exists() else {}
    except Exception:
      cur = {}
    cur[name] = rec
    p.write_text(json.dumps(cur, indent=1))


def _oracle_run(shape, n_prompt, n_gen):
  from oracle.ref_engine import RefEngine, RefShard, model_config_from_hf, generate_ids
  from oracle.synth_fast import synth_state_dict_fast
  import torch
  from tests._cpu import usable_cores
  torch.set_num_threads(usable_cores())         # NOT the affinity mask: the box runs under a CPU quota
  t = time.time()
  sd = synth_state_dict_fast(shape, 1234)
  ref = RefEngine(model_config_from_hf(hf_config(shape)), sd, RefShard(shape.name, 0, shape.n_layers - 1, shape.n_layers))
  prompt = synth_prompt(shape.vocab, n_prompt)
  o_ids, o_logits, gaps = generate_ids([ref], prompt, n_gen)
  return prompt, o_ids, o_logits, gaps, time.time() - t


def _native(shape):
  from xotorch_b200.native import NativeShard
  e = NativeShard(hf_config(shape), device=0)
  e.set_shard(0, shape.n_layers - 1)
  e.init_synthetic(1234)
  return e


def _forced(e, prompt, ids_forced, n_gen):
  """teacher-forced run through the host-buffer API: logits [n_gen, V]"""
  e.begin_request(prompt.shape[1] + n_gen + 1)
  x, pos, out = prompt.astype(np.int32), 0, []
  for s in range(n_gen):
    lg = e.forward(x, pos)
    assert int(np.argmax(lg[0])) == e.last_token(), "device argmax != host lowest-index argmax"
    out.append(lg[0].copy())
    pos += x.shape[1]
    x = np.array([[int(ids_forced[s])]], dtype=np.int32)
  return np.stack(out)


def _ring_ids(e, prompt, n_gen):
  """free-running greedy ids of the device-resident ring (what bench.py's `value` runs)"""
  e.link_next_local(e)
  e.begin_request(prompt.shape[1] + n_gen + 1)
  lg = e.forward(prompt.astype(np.int32), 0)
  first = e.last_token()
  e.post_to_next(0)
  e.ring_launch(n_gen - 1, 0, prompt.shape[1])
  toks, _ = e.ring_wait(n_gen - 1)
  return np.concatenate([[first], toks]).astype(np.int32)


def _check_logits(name, logits, o_logits, gaps, rms_tol, max_tol, extra=None):
  err = np.abs(logits - o_logits)
  std = float(o_logits.std())
  rms = float(np.sqrt((err ** 2).mean()))
  tie = 2.0 * max_tol * std                       # a top-2 gap below twice the logit tolerance cannot be decided
  decided = gaps > tie
  agree = np.argmax(logits, -1) == np.argmax(o_logits, -1)
  rec = {"steps": int(len(gaps)), "logit_std": round(std, 4), "rms_over_std": round(rms / std, 5), "max_over_std": round(float(err.max()) / std, 5),
         "rms_tol": rms_tol, "max_tol": max_tol, "near_tie_steps": int((~decided).sum()), "argmax_equal_steps": int(agree.sum()),
         "argmax_equal_on_decided": bool(agree[decided].all())}
  rec.update(extra or {})
  _report(name, rec)
  assert rms <= rms_tol * std, f"rms err {rms:.4f} vs logit std {std:.3f}"
  assert err.max() <= max_tol * std, f"max err {err.max():.4f} vs logit std {std:.3f}"
  assert agree[decided].all(), f"argmax differs at decided steps {np.nonzero(decided & ~agree)[0].tolist()}"
  # two CPU bf16 evaluations (oracle vs HF) agree on 88-95 % of the steps of these random-weight models (15/17, 61/64): most
  # top-2 gaps are inside the bf16 noise.  The agreement RATE is therefore a sanity bound, the per-step rule above is the test.
  assert agree.mean() >= 0.75, f"argmax agrees at only {int(agree.sum())} of {len(agree)} teacher-forced steps"
  return decided, agree


def test_llama3_8b_full_depth_teacher_forced():
  """BASELINE configs 2/3: the kernel instantiation bench.py times (<2048,128,4,12,dense>), all 32 layers, the benchmark's
  128-token prompt (tensor-core prefill), then 32 teacher-forced decode steps (persistent decode kernel)."""
  shape = PRESETS["llama-3-8b"]
  n_prompt, n_gen = 128, 33
  prompt, o_ids, o_logits, gaps, t_cpu = _oracle_run(shape, n_prompt, n_gen)
  e = _native(shape)
  logits = _forced(e, prompt, o_ids, n_gen)
  ring = _ring_ids(e, prompt, n_gen)
  decided, agree = _check_logits("llama-3-8b/32 layers/128+32 teacher-forced", logits, o_logits, gaps, 0.075, 0.45,
                                 {"oracle_cpu_s": round(t_cpu, 1), "ring_ids_equal_oracle_prefix": int(_common_prefix(ring, o_ids))})
  # free-running device ring: identical to the oracle's ids up to the first near-tie step
  k = _common_prefix(ring, o_ids)
  assert k == n_gen or not decided[k], f"ring ids diverge from the oracle at step {k} where the top-2 gap is decisive"


def _common_prefix(a, b) -> int:
  n = min(len(a), len(b))
  neq = np.nonzero(np.asarray(a[:n]) != np.asarray(b[:n]))[0]
  return int(neq[0]) if len(neq) else n


def test_llama32_1b_full_model_128_512():
  """BASELINE config 1 (the reference's CPU-runnable case): full 16-layer Llama-3.2-1B shape (tied head, llama3-scaled
  RoPE, hd 64), 128-token prompt / 512 generated: free-running ring ids + all 512 steps teacher-forced."""
  shape = PRESETS["llama-3.2-1b"]
  n_prompt, n_gen = 128, 512
  prompt, o_ids, o_logits, gaps, t_cpu = _oracle_run(shape, n_prompt, n_gen)
  e = _native(shape)
  ring = _ring_ids(e, prompt, n_gen)
  logits = _forced(e, prompt, o_ids, n_gen)
  k = _common_prefix(ring, o_ids)
  decided, agree = _check_logits("llama-3.2-1b/16 layers/128+512", logits, o_logits, gaps, 0.035, 0.20,
                                 {"oracle_cpu_s": round(t_cpu, 1), "free_running_ids_equal_prefix": k})
  assert k == n_gen or not decided[k], f"free-running ids diverge from the oracle at step {k} where the top-2 gap is decisive"


def test_llama3_70b_dims_fp16_cut():
  """BASELINE config 4 is fp16: Llama-3-70B dimensions (D 8192, F 28672, 64/8 heads -> the R=8, 8-stage instantiation) in
  IEEE half, 2-layer cut so the CPU oracle finishes in seconds."""
  import dataclasses
  shape = dataclasses.replace(PRESETS["llama-3-70b"], n_layers=2, name="llama-3-70b-cut2-f16")
  assert shape.dtype == "float16"
  n_prompt, n_gen = 12, 5
  prompt, o_ids, o_logits, gaps, t_cpu = _oracle_run(shape, n_prompt, n_gen)
  e = _native(shape)
  logits = _forced(e, prompt, o_ids, n_gen)
  _check_logits("llama-3-70b dims/fp16/2-layer cut", logits, o_logits, gaps, 0.02, 0.10, {"oracle_cpu_s": round(t_cpu, 1)})
  ring = _ring_ids(e, prompt, n_gen)
  e.begin_request(n_prompt + n_gen + 1)
  x, pos, api = prompt.astype(np.int32), 0, []
  for s in range(n_gen):
    lg = e.forward(x, pos)
    api.append(int(np.argmax(lg[0])))
    pos += x.shape[1]
    x = np.array([[api[-1]]], dtype=np.int32)
  assert np.array_equal(ring, np.array(api)), "device ring != host API (fp16, R=8)"
