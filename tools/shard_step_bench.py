"""Development aid: decode step time of ONE shard (a layer range of a preset) on one GPU, ring mode with the shard feeding
itself (hidden state or token loops through its own inbox).  usage: shard_step_bench.py <preset> <start> <end> [n_steps] [ctx]"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from xotorch_b200.synth import PRESETS, hf_config, synth_prompt
from xotorch_b200.native import NativeShard

name, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 256
ctx = int(sys.argv[5]) if len(sys.argv) > 5 else 128
shape = PRESETS[name]
e = NativeShard(hf_config(shape), 0)
e.set_shard(a, b)
e.init_synthetic(1234)
e.link_next_local(e)
step0 = 0
for it in range(3):
  e.begin_request(ctx + n + 1)
  if e.is_first:
    e.forward(synth_prompt(shape.vocab, ctx), 0)
  else:
    x = (np.random.default_rng(0).standard_normal((1, ctx, shape.dim)) * 0.5).astype(np.float32)
    e.forward(x, 0)
  e.post_to_next(step0)
  e.ring_launch(n, step0, ctx)
  toks, ms = e.ring_wait(n)
  step0 += n
  gb = e.weight_bytes() / 1e9
  print(json.dumps({"preset": name, "layers": [a, b], "iter": it, "us_per_step": round(ms / n * 1e3, 1), "weight_GB": round(gb, 2),
                    "weight_stream_GBs": round(gb / (ms / n * 1e-3), 1)}), flush=True)
