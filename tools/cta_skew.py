"""Development aid: per-op skew between the CTAs of the persistent decode kernel (one ring step).  For every op, every CTA
records when it starts consuming, when its tiles are done and how many ring stages had landed at the start; printed relative to
the earliest CTA.  usage: python tools/cta_skew.py [model] [n_prompt] [n_gen] [layer]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from xotorch_b200.synth import PRESETS, hf_config, synth_prompt
from xotorch_b200.native import NativeShard

name = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
n_prompt = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n_gen = int(sys.argv[3]) if len(sys.argv) > 3 else 512
layer = int(sys.argv[4]) if len(sys.argv) > 4 else 5
shape = PRESETS[name]
e = NativeShard(hf_config(shape), 0)
e.set_shard(0, shape.n_layers - 1)
e.init_synthetic(1234)
e.link_next_local(e)
prompt = synth_prompt(shape.vocab, n_prompt)
step0 = 0
for it in range(2):
  if it == 1:
    e.debug_timeline(200, 0)
  e.begin_request(n_prompt + n_gen + 1)
  e.forward(prompt, 0); e.post_to_next(step0); e.ring_launch(n_gen, step0, n_prompt)
  toks, ms = e.ring_wait(n_gen)
  step0 += n_gen
  print("tok/s", round(n_gen / ms * 1e3, 1))
raw = e.debug_timeline(-1, 0, read="all").astype(np.int64)
nops = 4 * shape.n_layers + 1
base = (2 * nops + 2) * 4 + 4 * 64 + 8
allc = raw[base:base + nops * 256 * 3].reshape(nops, 256, 3)
ncta = int((allc[0, :, 0] > 0).sum())
names = ["qkv", "o", "gu", "down"]
def show(op, label):
  a = allc[op, :ncta]
  t0 = a[:, 0].min()
  st, dn, ld = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3, a[:, 2]
  dur = dn - st
  q = lambda v: "min %.2f p10 %.2f med %.2f p90 %.2f max %.2f" % (v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max())
  print(f"{label:8s} start [{q(st)}]  tiles [{q(dur)}]  done [{q(dn)}]  landed min {ld.min()} med {int(np.median(ld))}")
  worst = np.argsort(dn)[-6:]
  print("         slowest CTAs (cta: start, tiles, landed):", [(int(c), round(float(st[c]), 2), round(float(dur[c]), 2), int(ld[c])) for c in worst])
for k in range(4):
  show(4 * layer + k, f"L{layer} {names[k]}")
show(nops - 1, "head")
# which CTAs are habitually late?  mean rank of the done time over all ops
late = np.zeros(ncta)
for op in range(nops):
  late += np.argsort(np.argsort(allc[op, :ncta, 1]))
late /= nops
print("habitually late CTAs (mean rank of 'tiles done', 0 = first):", [(int(c), round(float(late[c]), 1)) for c in np.argsort(late)[-10:]])
print("habitually early CTAs:", [(int(c), round(float(late[c]), 1)) for c in np.argsort(late)[:10]])
