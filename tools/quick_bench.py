"""Quick single-GPU ring benchmark (development aid; bench.py is the contract)."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from xotorch_b200.synth import PRESETS, hf_config, synth_prompt
from xotorch_b200.native import NativeShard

name = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
n_prompt = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n_gen = int(sys.argv[3]) if len(sys.argv) > 3 else 512
shape = PRESETS[name]
t = time.time()
e = NativeShard(hf_config(shape), 0)
e.set_shard(0, shape.n_layers - 1)
e.init_synthetic(1234)
print("init", time.time() - t, "s; weight GB", e.weight_bytes() / 1e9, flush=True)
e.link_next_local(e)
prompt = synth_prompt(shape.vocab, n_prompt)
step0 = 0
for it in range(3):
  e.begin_request(n_prompt + n_gen + 1)
  t = time.time()
  logits = e.forward(prompt, 0)
  tp = time.time() - t
  e.post_to_next(step0)
  e.ring_launch(n_gen, step0, n_prompt)
  toks, ms = e.ring_wait(n_gen)
  step0 += n_gen
  ctx = n_prompt + n_gen // 2
  gbs = e.step_bytes(ctx) * n_gen / (ms * 1e-3) / 1e9
  print(json.dumps({"iter": it, "prefill_s": round(tp, 3), "decode_ms": round(ms, 2), "tok_s": round(n_gen / ms * 1e3, 1),
                    "ms_per_tok": round(ms / n_gen, 4), "achieved_GBs": round(gbs, 1), "first_ids": toks[:8].tolist(),
                    "launches": e.kernel_launches()}), flush=True)
