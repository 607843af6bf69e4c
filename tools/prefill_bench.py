"""Prefill time through the host-buffer API (NativeShard.forward with token ids -> last-position logits) for several prompt
lengths.  usage: python tools/prefill_bench.py [preset] [len ...]   (run on a B200)"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from xotorch_b200.native import NativeShard
from xotorch_b200.synth import PRESETS, hf_config, synth_prompt

name = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
lens = [int(x) for x in sys.argv[2:]] or [128, 1024, 4096]
shape = PRESETS[name]
e = NativeShard(hf_config(shape), 0)
e.set_shard(0, shape.n_layers - 1)
e.init_synthetic(1234)
params = e.step_bytes(0)                       # flop per token ~ 2 * parameters touched = bytes touched per token (16-bit weights)
for n in lens:
  prompt = synth_prompt(shape.vocab, n)
  best = 1e9
  for it in range(3):
    e.begin_request(n + 8)
    t0 = time.perf_counter()
    e.forward(prompt, 0)
    best = min(best, time.perf_counter() - t0)
  print(f"{name} prefill {n:5d} tokens: {best * 1e3:8.2f} ms  {n / best:9.0f} tok/s  ~{params * n / best / 1e12:6.1f} TFLOP/s (GEMMs only)", flush=True)
