#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the shard-decode hot path.

Metric (BASELINE.json): Llama-3-8B greedy decode tokens/sec (128-token prompt, 512 generated, batch 1) at N
ring shards (one process per GPU, launched by torchrun for N>1), plus the mean inter-shard hop latency.

A "step" is ONE request of the workload: a 128-token prefill (not timed, reported in `config`) followed by the timed
512-token decode.  `value` = K*512 / (decode time of K requests): the device-resident ring (weights, KV and
activations in HBM; on-device argmax feeds the next step; no host work per token), timed on the device with CUDA
events on the launching stream, max over ranks.  `e2e` = the same metric through the host-buffer plugin API the
reference's Node drives (`forward(ids/hidden ndarray) -> ndarray`, i.e. `InferenceEngine.infer_tensor`): every
token copies its input host->device and its result (fp32 logits / hidden state) device->host inside the timed region.

`python bench.py --impl reference` times the CPU restatement of the reference engine (oracle/, test infrastructure;
the reference itself cannot be imported: torchtune is absent) on the host cores — a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_PROMPT, N_GEN = 128, 512
TRAFFIC_FILE = ROOT / "profiles" / "r2_decode_ncu_traffic.json"   # written by tools/ncu_summarize.py from an `ncu --set full` capture of THIS build
MODEL = "llama-3-8b"
SEED = 1234


def _peaks():
  p = ROOT / "MEASURED_PEAKS.json"
  if p.exists():
    try:
      d = json.loads(p.read_text())
      return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
      pass
  return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
  """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index: int):
    self.gpu, self.proc, self.lines = gpu_index, None, []

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                    "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append((time.time(), line.strip()))

  def stop(self, t0: float, t1: float) -> dict:
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.15)
    self.proc.terminate()
    sm, mx, reasons = [], None, set()
    for ts, ln in self.lines:
      if ts < t0 or ts > t1 + 0.2:
        continue
      f = [x.strip() for x in ln.split(",")]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); mx = float(f[2])
      except ValueError:
        continue
      for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
        if v.lower().startswith("active"):
          reasons.add(name)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def usable_cores() -> int:
  """Host cores this process may actually use: min(sched affinity, cgroup cpu quota) — os.cpu_count() reports the
  whole host and oversubscribes a quota-limited container."""
  n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
  for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    try:
      txt = Path(p).read_text().split()
      if p.endswith("cpu.max"):
        if txt[0] != "max":
          n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
      else:
        q = int(txt[0])
        if q > 0:
          per = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
          n = min(n, max(1, int(q / per + 0.5)))
      break
    except Exception:
      continue
  return max(1, n)


def ring_partition(hf_cfg: dict, n: int, rank: int, how: str):
  """Layer range of ring position `rank` on n equal-HBM GPUs.  "layers": the reference rule
  (RingMemoryWeightedPartitioningStrategy + map_partitions_to_shards, topology/*).  "bytes": the same ring order with the
  bytes each shard streams per token balanced (the last shard also streams the lm_head) — xotorch_b200/topology/partitioning.py."""
  from xotorch_b200.topology.partitioning import (ByteBalancedRingPartitioningStrategy, RingMemoryWeightedPartitioningStrategy,
                                                   map_partitions_to_shards)
  mems = {f"gpu{n - 1 - i:02d}": 183359 for i in range(n)}          # ids chosen so that ring position i == rank i
  st = ByteBalancedRingPartitioningStrategy.for_model(hf_cfg) if how == "bytes" else RingMemoryWeightedPartitioningStrategy()
  shards = map_partitions_to_shards(st.partition(mems), hf_cfg["num_hidden_layers"], MODEL)
  assert len(shards) == n
  return shards[rank].start_layer, shards[rank].end_layer, [s.get_layer_count() for s in shards]


# ----------------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (oracle = test infrastructure; only place bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------------------------
_CPU_ENGINE = None


def cpu_reference_engine():
  """The CPU restatement of the reference engine (oracle/, test infrastructure) on the SAME workload: the full-depth model
  (all layers + lm_head, synthetic weights of the same seed), prefilled with the same 128-token prompt.  Built once per
  process (weight synthesis is the expensive part).  Returns (engine, state after the prefill, first token, cores)."""
  global _CPU_ENGINE
  if _CPU_ENGINE is None:
    import torch
    from oracle.ref_engine import RefEngine, RefShard, model_config_from_hf
    from oracle.synth_fast import synth_state_dict_fast
    from xotorch_b200.synth import PRESETS, hf_config, synth_prompt
    shape = PRESETS[MODEL]
    cores = usable_cores()
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    sd = synth_state_dict_fast(shape, SEED)
    t_w = time.perf_counter() - t0
    eng = RefEngine(model_config_from_hf(hf_config(shape)), sd, RefShard(MODEL, 0, shape.n_layers - 1, shape.n_layers))
    prompt = synth_prompt(shape.vocab, N_PROMPT)
    t0 = time.perf_counter()
    logits, state = eng.infer_tensor("cpu", eng.shard, prompt.astype(np.int32), None)      # prefill (never timed)
    tok = eng.sample(logits, temp=0.0)
    _CPU_ENGINE = {"eng": eng, "state": state, "tok": tok, "cores": cores, "weights_s": t_w, "prefill_s": time.perf_counter() - t0,
                   "layers": shape.n_layers}
  return _CPU_ENGINE


def cpu_decode_tokens(n_tokens: int) -> float:
  """n_tokens REAL greedy decode steps at full depth on the host cores (continuing the prefilled request: KV of
  prompt + 1024, full-length masked SDPA over the whole cache, GQA expand, fp32 logits, reference sampler at temperature 0 —
  the per-token cost of the reference engine does not depend on the position).  Returns the wall seconds."""
  c = cpu_reference_engine()
  eng = c["eng"]
  t0 = time.perf_counter()
  for _ in range(n_tokens):
    logits, c["state"] = eng.infer_tensor("cpu", eng.shard, c["tok"].astype(np.int32).reshape(1, 1), c["state"])
    c["tok"] = eng.sample(logits, temp=0.0)                                # node.py:121
  return time.perf_counter() - t0


def cpu_sample_desc(c, n_tokens, what):
  return (f"oracle RefEngine = CPU restatement of the reference torch engine (bf16 weights, KV prompt+1024, full-length masked SDPA, GQA expand, "
          f"reference sampler temp 0), FULL {MODEL} shape: all {c['layers']} layers + lm_head, {N_PROMPT}-token prefill ({c['prefill_s']:.1f} s, untimed; "
          f"weights synthesised in {c['weights_s']:.0f} s) then {n_tokens} real decode steps {what}; nothing extrapolated")


def workload_string(dtype_name: str) -> str:
  """one string for both arms (`config.workload`), so the driver can tell they measure the same thing"""
  return (f"{MODEL} shape, synthetic {dtype_name} weights (seed {SEED}), {N_PROMPT}-token prompt / {N_GEN} generated, batch 1, greedy; "
          f"metric = decode tokens/s")


def run_reference(args):
  """`--impl reference`: the reference's own CPU implementation of the path.  The reference engine cannot be imported
  (torchtune is absent here and on the box), so this is the oracle port, timed on all usable host cores.
  A step = a bounded sample of the workload: `tok_per_step` consecutive decode tokens of the full-depth model (at least 32
  timed tokens in total); W warm-up steps, then K timed steps; value = timed tokens / timed seconds."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  from xotorch_b200.synth import PRESETS
  t_start = time.time()
  sh = PRESETS[MODEL]
  need = 2 * (sh.n_layers * (sh.dim * (sh.n_heads + 2 * sh.n_kv_heads) * sh.head_dim + sh.n_heads * sh.head_dim * sh.dim
                             + 3 * sh.dim * sh.ffn * max(1, sh.n_experts)) + 2 * sh.vocab * sh.dim)
  try:
    avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
  except Exception:
    avail = 1 << 62
  if need > 0.7 * avail:       # 70B / Mixtral: the full model does not fit the host's RAM (report n/a, do not extrapolate)
    print(json.dumps({"impl": "reference", "unavailable": f"{MODEL}: {need / 1e9:.0f} GB of weights do not fit the host's {avail / 1e9:.0f} GB of available RAM"}), flush=True)
    return
  steps, warm = max(1, args.steps), max(0, args.warmup)
  tok_per_step = max(2, -(-32 // steps))
  c = cpu_reference_engine()
  for _ in range(warm):
    cpu_decode_tokens(tok_per_step)
  secs = [cpu_decode_tokens(tok_per_step) for _ in range(steps)]
  v = steps * tok_per_step / float(np.sum(secs))
  hop = None
  if args.gpus > 1 and not args.no_ref_hop:
    try:
      from tools.ref_hop_bench import measure_reference_hop
      hop = measure_reference_hop(dim=PRESETS[MODEL].dim)
    except Exception as ex:                      # the comparator must never take the arm down
      hop = {"unavailable": repr(ex)[:300]}
  desc = cpu_sample_desc(c, steps * tok_per_step, f"in {steps} steps of {tok_per_step} tokens after {warm} warm-up steps")
  line = {
    "impl": "reference", "metric": ("Llama-3-8B" if MODEL == "llama-3-8b" else MODEL) + f" greedy decode tokens/sec ({N_PROMPT}-prompt/{N_GEN}-gen, batch 1)", "value": round(v, 4),
    "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
    "ms_per_step": round(float(np.mean(secs)) * 1e3, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
    "dtype": "f16" if PRESETS[MODEL].dtype == "float16" else "bf16", "data": "synthetic",
    "config": {"workload": workload_string(PRESETS[MODEL].dtype), "shards": args.gpus,
               "step": f"{tok_per_step} consecutive decode tokens of the full-depth model on the host CPU (bounded sample of the {N_GEN}-token generation)"},
    "cpu_baseline": {"value": round(v, 4), "unit": "tokens/s", "cores": c["cores"], "kind": "port", "sample": desc},
    "e2e": {"value": round(v, 4), "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "hop": hop, "wall_s": round(time.time() - t_start, 1),
  }
  print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
  import torch
  import torch.distributed as dist
  from xotorch_b200.native import NativeShard
  from xotorch_b200.synth import PRESETS, hf_config, synth_prompt

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
  n = world
  if not torch.cuda.is_available():
    raise SystemExit("bench.py: no CUDA device — xotorch_b200 has no CPU fallback (use --impl reference for the CPU arm)")
  torch.cuda.set_device(local)
  if n > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  shape = PRESETS[MODEL]
  cfg = hf_config(shape)
  a, b, layer_counts = ring_partition(cfg, n, rank, args.partition) if n > 1 else (0, shape.n_layers - 1, [shape.n_layers])
  eng = NativeShard(cfg, device=local)
  eng.set_shard(a, b)
  eng.init_synthetic(SEED)
  first, last = eng.is_first, eng.is_last
  prev_r, next_r = (rank - 1) % n, (rank + 1) % n
  D, V = shape.dim, shape.vocab

  # ---- link the ring: every engine learns the inbox of the next one (CUDA IPC across processes) ----
  from xotorch_b200.ring import DistributedRing
  prompt = synth_prompt(V, N_PROMPT)
  max_len = N_PROMPT + N_GEN + 1
  dev = torch.device("cuda", local)
  ring = DistributedRing(eng, rank, n, D, V, comm_device=dev)
  ring.link()

  def prefill():
    """128-token prefill through the host-buffer API; hidden state handed over per shard (bytes = 2*S*D algorithmic)."""
    return ring.prefill(prompt, max_len)

  def prefill_dev():
    """the same prefill with the hidden state handed over GPU to GPU (NCCL send/recv of the 16-bit [S, D] tensor)"""
    return ring.prefill_device(prompt, max_len)

  def ring_decode(n_steps: int):
    """device-resident ring: last shard seeds shard 0 with the first generated token, then ONE persistent kernel/rank."""
    torch.cuda.synchronize()
    return ring.ring_decode(n_steps, N_PROMPT)

  def barrier():
    if n > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def max_over_ranks(v: float) -> float:
    if n == 1:
      return v
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  # ---- warm-up requests ----
  prefill_ms = None
  prefill_dev_ms = None
  if n > 1:                                     # GPU-to-GPU prefill hop (not part of any timed region): measured once, warm
    for _ in range(2):
      barrier(); t0 = time.perf_counter(); prefill_dev(); barrier(); prefill_dev_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
  for _ in range(max(args.warmup, 0)):
    t0 = time.perf_counter(); prefill(); barrier(); prefill_ms = (time.perf_counter() - t0) * 1e3
    ring_decode(N_GEN)
  # ---- timed: K requests, decode phases timed with CUDA events on the launching stream ----
  sampler = ClockSampler(local)
  launches0 = eng.kernel_launches()
  dec_ms, ring_tokens, hop_stats = [], None, []
  t_region0 = time.time()
  sampler.start()
  barrier()
  for k in range(args.steps):
    t0 = time.perf_counter(); prefill(); barrier(); prefill_ms = (time.perf_counter() - t0) * 1e3
    l0 = eng.kernel_launches()
    toks, ms = ring_decode(N_GEN)
    dec_ms.append(max_over_ranks(ms))
    launches_decode = eng.kernel_launches() - l0
    if toks is not None:
      ring_tokens = toks
    # hop accounting: per-step compute span on each rank (own clock) vs ring period on rank 0
    t_in, t_out = eng.ring_timestamps(N_GEN)
    comp = (t_out.astype(np.int64) - t_in.astype(np.int64)).astype(np.float64)     # ns, per step, this shard
    if n > 1:
      allc = [None] * n
      dist.all_gather_object(allc, comp)
      period = np.diff(t_in.astype(np.int64)).astype(np.float64)                   # rank-local ring period
      if rank == 0:
        tot = np.sum(np.stack(allc), axis=0)[:-1]
        hop_stats.append((period - tot) / n)                                       # mean per-hop ns in each ring turn
  barrier()
  t_region1 = time.time()
  clocks = sampler.stop(t_region0, t_region1)
  total_ms = float(np.sum(dec_ms))
  value = args.steps * N_GEN / (total_ms * 1e-3)
  gpu_launches = (eng.kernel_launches() - launches0)

  # ---- e2e: host-buffer API per token (H2D ids/hidden in, D2H logits/hidden out, every step) ----
  def api_request():
    prefill_out = prefill()
    tok = int(np.argmax(prefill_out[0])) if last else 0
    if n > 1:
      tok = ring._bcast_token(tok)
    barrier()
    t0 = time.perf_counter()
    out_ids = ring.api_decode_device(tok, N_PROMPT, N_GEN)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0), out_ids

  api_request()                                 # warm
  e2e_times, e2e_ids = [], None
  for _ in range(max(1, min(args.steps, 2))):
    dt, e2e_ids = api_request()
    e2e_times.append(max_over_ranks(dt))
  e2e_value = N_GEN / float(np.mean(e2e_times))
  # host<->device bytes of one request, summed over the ranks: ids in on the first shard, logits out on the last (every token),
  # plus the prompt's fp32 hidden state through the host-array API once per shard boundary
  h2d = (N_PROMPT + N_GEN) * 4 + (n - 1) * N_PROMPT * D * 4
  d2h = (N_GEN + 1) * V * 4 + (n - 1) * N_PROMPT * D * 4

  # ---- requests in flight (N>1): B = N independent batch-1 requests interleaved through the ring, so every shard is busy.
  #      Reported beside the headline, never instead of it: `value` stays the single-request metric BASELINE.json names. ----
  pipelined = None
  if n > 1:
    Bq, gen_q = min(n, 8), min(256, N_GEN)
    prompts_q = np.concatenate([synth_prompt(V, N_PROMPT, seed=SEED + 1 + r) for r in range(Bq)])
    ring.generate_batch(prompts_q, gen_q)                  # warm
    barrier()
    ids_q, ms_q = ring.generate_batch(prompts_q, gen_q)
    ms_q = max_over_ranks(ms_q)
    barrier()
    pipelined = {"requests_in_flight": Bq, "tokens_per_request": gen_q, "value": round(Bq * (gen_q - 1) / (ms_q * 1e-3), 2),
                 "unit": "tokens/s (all requests)", "per_request": round((gen_q - 1) / (ms_q * 1e-3), 2),
                 "how": "ring step s serves request s % B (own KV cache each, each still a batch-1 greedy stream); decode phase, CUDA events, max over ranks"}

  # ---- roofline of the decode step (algorithmic bytes of SURVEY.md §8d / DESIGN.md, live CUDA-event time) ----
  peak, peak_src = _peaks()
  ctx_avg = N_PROMPT + (N_GEN - 1) / 2.0
  step_bytes_local = eng.step_bytes(int(round(ctx_avg)))
  if n > 1:
    allb = [None] * n
    dist.all_gather_object(allb, step_bytes_local)
    step_bytes_total = int(sum(allb))
  else:
    step_bytes_total = step_bytes_local
  achieved = step_bytes_total * args.steps * N_GEN / (total_ms * 1e-3) / 1e9

  # ---- id checks (north star: greedy ids of the same prompt must not depend on how the model is sharded or driven) ----
  #  (1) device-resident ring == per-token API path (all N_GEN-1 ids they share);  (2) N>1: N shards == ONE shard holding all
  #  layers, run on the last rank's GPU after the timed regions (16 GB more on a 180 GB device);  sha1 of the ring ids so the
  #  lines of different N can be compared with each other.
  import hashlib
  checks = None
  if last:
    ring_l = [int(t) for t in ring_tokens[:N_GEN - 1]]
    checks = {"ring_ids_equal_api_ids": bool(e2e_ids is not None and ring_l == [int(t) for t in e2e_ids[:N_GEN - 1]]),
              "first_ids": ring_l[:8], "ids_sha1": hashlib.sha1(np.asarray(ring_l, dtype=np.int32).tobytes()).hexdigest()[:16],
              "n_shard_ids_equal_1_shard_ids": None}
    if n > 1 and not args.no_single_shard_check:
      one = NativeShard(cfg, device=local)
      one.set_shard(0, shape.n_layers - 1)
      one.init_synthetic(SEED)
      one.link_next_local(one)
      one.begin_request(max_len)
      one.forward(prompt, 0)
      s1 = one.ring_next_step()
      one.post_to_next(s1)
      one.ring_launch(N_GEN, s1, N_PROMPT)
      toks1, _ = one.ring_wait(N_GEN)
      checks["n_shard_ids_equal_1_shard_ids"] = bool([int(t) for t in toks1[:N_GEN - 1]] == ring_l)
      one.close()
  if n > 1:
    info = [None] * n
    dist.all_gather_object(info, checks)
    checks = [x for x in info if x is not None][0]
  first_ids = checks["first_ids"]

  traffic = {}
  if TRAFFIC_FILE.exists():
    try:
      traffic = json.loads(TRAFFIC_FILE.read_text())
    except Exception:
      traffic = {}
  if rank == 0:
    cpu = None
    if n == 1 and not args.no_cpu_baseline:
      try:
        c = cpu_reference_engine()
        cpu_decode_tokens(2)                                # warm
        n_cpu = 16
        v = n_cpu / cpu_decode_tokens(n_cpu)
        cpu = {"value": round(v, 4), "unit": "tokens/s", "cores": c["cores"], "kind": "port", "sample": cpu_sample_desc(c, n_cpu, "after 2 warm-up steps")}
      except Exception as ex:  # the CPU leg must never take the GPU number down with it
        cpu = {"value": None, "unit": "tokens/s", "cores": usable_cores(), "kind": "port", "sample": f"failed: {ex!r}"}
    hop = None
    if hop_stats:
      hs = np.concatenate(hop_stats) / 1e3
      hop = {"p50_us": round(float(np.median(hs)), 2), "p90_us": round(float(np.percentile(hs, 90)), 2),
             "how": "per ring turn: (rank-0 ring period - sum over shards of [flag observed -> result posted], each on its own globaltimer) / n_hops"}
    line = {
      "metric": ("Llama-3-8B" if MODEL == "llama-3-8b" else MODEL) + f" greedy decode tokens/sec ({N_PROMPT}-prompt/{N_GEN}-gen, batch 1)",
      "value": round(value, 2), "unit": "tokens/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 3),
      "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16" if shape.dtype == "float16" else "bf16", "data": "synthetic",
      "config": {"workload": workload_string(shape.dtype),
                 "step": "one request: prefill (untimed) + the timed decode of all generated tokens; weights generated on the device",
                 "shards": n, "layers_per_shard": layer_counts,
                 "partitioning": ("ring order by (memory, id) like the reference, equal HBM; " +
                                  ("bytes streamed per token balanced: the last shard also streams the lm_head (ByteBalancedRingPartitioningStrategy)"
                                   if args.partition == "bytes" and n > 1 else "equal layer counts (the reference rule)")),
                 "l2_policy": f"inputs larger than L2 ({step_bytes_total / 1e9:.1f} GB of weights + KV streamed per token vs 126 MB L2)",
                 "prefill_ms_host_api": round(prefill_ms, 1),
                 "prefill_ms_device_hop": None if prefill_dev_ms is None else round(prefill_dev_ms, 1), "first_ids": first_ids,
                 "ids_sha1": checks["ids_sha1"], "ring_ids_equal_api_ids": checks["ring_ids_equal_api_ids"],
                 "n_shard_ids_equal_1_shard_ids": checks["n_shard_ids_equal_1_shard_ids"]},
      "clocks": clocks, "gpu_launches": int(gpu_launches),
      "e2e": {"value": round(e2e_value, 2), "unit": "tokens/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
              "how": "one engine call per shard per token through the C ABI: token ids from host memory in, logits to host memory out, host argmax; "
                     + ("one engine call per rank per token (xotb200_forward_hop): the producing kernel stores the 16-bit hidden state [1, D] into the "
                        "next rank's inbox over NVLink and releases a flag the consuming kernel acquires (no NCCL, no stream sync per hop); "
                        "only the 4-byte token id travels back through torch.distributed" if n > 1 else "single shard: NativeShard.forward == InferenceEngine.infer_tensor with host ndarrays")},
      "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                   "traffic": traffic.get("bytes_per_token") if (MODEL == "llama-3-8b" and n == 1) else None,
                   "traffic_note": traffic.get("note"),
                   "peak_source": peak_src,
                   "kernel": "whole decode step (all shard kernels of one token), algorithmic bytes = weights once + valid KV",
                   "bytes_per_token": step_bytes_total},
      "hop": hop, "pipelined": pipelined, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
  ring.close()
  if n > 1:
    dist.barrier()
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=3)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--partition", default="bytes", choices=["bytes", "layers"], help="N>1: balance streamed bytes (default) or layer counts (reference rule)")
  ap.add_argument("--no-single-shard-check", action="store_true", help="N>1: skip re-running the request on ONE shard holding all layers")
  ap.add_argument("--no-ref-hop", action="store_true", help="--impl reference --gpus N>1: skip the reference gRPC hop comparator")
  # the other BASELINE.json configs through the same contract (defaults = the headline config):
  #   --model llama-3-70b --gpus 8 --prompt 4096      (fp16, must shard to fit)     --model mixtral-8x7b --gpus 4
  from xotorch_b200.synth import PRESETS
  ap.add_argument("--model", default=MODEL, choices=sorted(k for k in PRESETS if not k.startswith("tiny")))
  ap.add_argument("--prompt", type=int, default=N_PROMPT)
  ap.add_argument("--gen", type=int, default=N_GEN)
  args = ap.parse_args()
  globals().update(MODEL=args.model, N_PROMPT=args.prompt, N_GEN=args.gen)
  if args.impl == "reference":
    run_reference(args)
  else:
    run_b200(args)


if __name__ == "__main__":
  main()
