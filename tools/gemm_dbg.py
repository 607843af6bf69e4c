import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from xotorch_b200.native import NativeShard
from xotorch_b200.synth import PRESETS, hf_config
def bits(x): return x.to(torch.bfloat16).contiguous().view(torch.int16).numpy().view(np.uint16)
def unb(b): return torch.from_numpy(b.view(np.int16).copy()).view(torch.bfloat16).float().numpy()
e = NativeShard(hf_config(PRESETS["tiny-llama-hd64"]), 0)
for (S,N,K,tk) in [(128,768,512,512),(128,768,512,512),(64,768,512,512),(128,1536,512,512),(128,768,1024,512),(128,768,512,64)]:
    g = torch.Generator().manual_seed(1)
    a = (torch.randn(S,K,generator=g)*0.5).to(torch.bfloat16); w=(torch.randn(N,K,generator=g)*0.05).to(torch.bfloat16)
    out = unb(e.debug_gemm(bits(a), bits(w), tk))
    ref = torch.tensor(a.float().numpy().astype(np.float64) @ w.float().numpy().astype(np.float64).T, dtype=torch.float32).to(torch.bfloat16).float().numpy()
    bad = np.abs(out-ref) > np.maximum(np.abs(ref),1e-3)*2.0**-6
    cols = np.where(bad.any(0))[0]; rows=np.where(bad.any(1))[0]
    print((S,N,K,tk), 'bad frac', bad.mean(), 'bad col blocks(64):', sorted(set((cols//64).tolist())), 'rows:', (rows.min(), rows.max()) if len(rows) else None)
