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
if ":" in name:      # "llama-3-70b:8" = that preset cut to its first 8 layers (fits one GPU; per-layer behaviour is what is measured)
  import dataclasses
  base, nl = name.split(":")
  shape = dataclasses.replace(PRESETS[base], n_layers=int(nl), name=f"{base}-{nl}l")
else:
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

if len(sys.argv) > 4:   # timeline of one ring step
  e.debug_timeline(200, int(sys.argv[4]))
  e.begin_request(n_prompt + n_gen + 1)
  e.forward(prompt, 0); e.post_to_next(step0); e.ring_launch(n_gen, step0, n_prompt); e.ring_wait(n_gen); step0 += n_gen
  raw = e.debug_timeline(-1, 0, read=True).astype(np.int64)
  nops = 4 * shape.n_layers + 1
  stg = raw[(2 * nops + 2) * 4:(2 * nops + 2) * 4 + 256].reshape(4, 64)
  tl = raw[:(2 * nops + 2) * 4].reshape(-1, 4)
  landed = tl[nops + 1: 2 * nops + 1, 0]
  tl = tl[:nops]
  names = ["qkv", "o", "gu", "down"]
  pro = tl[:, 1] - tl[:, 0]; tiles = tl[:, 2] - tl[:, 1]; bar = tl[:, 3] - tl[:, 2]
  gap = np.concatenate([tl[1:, 0] - tl[:-1, 3], [0]])
  for k in range(4):
    sel = np.arange(k, nops - 1, 4)
    print(f"{names[k]:5s} prologue {pro[sel].mean()/1e3:6.2f} us  tiles {tiles[sel].mean()/1e3:6.2f} us  barrier(+attn) {bar[sel].mean()/1e3:6.2f} us  publish-gap {gap[sel].mean()/1e3:5.2f} us")
  for k in range(4):
    print(names[k], "landed stages at consume start (first 6 layers):", landed[k:24:4].tolist(), "mean", landed[k:nops-1:4].mean())
  print(f"head  prologue {pro[-1]/1e3:6.2f} us  tiles {tiles[-1]/1e3:6.2f} us  barrier {bar[-1]/1e3:6.2f} us")
  for k in range(4):
    row = stg[k]; row = row[row > 0]
    t0 = tl[20 + k, 1]
    print(names[k], "stage arrival (us after consume start):", [round((v - t0) / 1e3, 2) for v in row[:16]])
  for k in range(4):     # grid barrier closing op k of layer 5: local arrival (all 9 warps) -> atomic issued -> all CTAs arrived
    b = stg[k][40:45]
    if b[0] > 0:
      t2k = tl[20 + k, 2]
      print(f"{names[k]:5s} barrier of layer 5: consumers done -> epilogue warp done {(b[4]-t2k)/1e3:5.2f} | -> CTA arrived {(b[0]-t2k)/1e3:5.2f} | atomic issued +{(b[1]-b[0])/1e3:4.2f} | all CTAs seen +{(b[2]-b[1])/1e3:4.2f} us")
  st = stg[0][48:54]; t2 = tl[20, 2]
  print("attention chain (layer 5, CTA %s): tiles done -> B1 exit %.2f | K/V+softmax loop %.2f | warp+smem merge, partial store %.2f | ticket %.2f | (last CTA) split merge %.2f | B2 %.2f us" % ((sys.argv[4],) + tuple(round(x / 1e3, 2) for x in [st[0] - t2, st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[5] - st[4]])))
  rr = stg[0][54:64]; rr = rr[rr > 0]
  if len(rr):
    print("direct attention rounds (us after B1 exit):", [round((v - st[0]) / 1e3, 2) for v in rr], " loop end:", round((st[1] - st[0]) / 1e3, 2))
  print("step total us", (tl[-1, 3] - tl[0, 0]) / 1e3, " sum prologue", pro.sum() / 1e3, " tiles", tiles.sum() / 1e3, " barrier", bar.sum() / 1e3)
