"""TEST INFRASTRUCTURE: how far apart are two equally valid bf16 evaluations of the SAME full-depth model?

Runs this oracle (oracle/ref_engine.py, interleaved RoPE + permuted q/k, torchtune-order ops) and HF transformers'
LlamaForCausalLM (half-split RoPE, un-permuted q/k) on identical synthetic weights at FULL depth, both in bf16 on the CPU,
and prints rms / max |delta logit| relative to the logit std, per position.  The number calibrates the tolerance of
tests/test_gpu_full_depth.py: the CUDA path has to be at least as close to the oracle as HF is.

  python -m oracle.calibrate_depth llama-3-8b 128 8      (model, prompt tokens, teacher-forced decode steps)
"""
import json
import sys
import time

import numpy as np
import torch

from oracle.ref_engine import RefEngine, RefShard, generate_ids, model_config_from_hf
from oracle.synth_fast import synth_state_dict_fast
from xotorch_b200.synth import PRESETS, hf_config, synth_prompt


def main():
  name = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
  n_prompt = int(sys.argv[2]) if len(sys.argv) > 2 else 128
  n_dec = int(sys.argv[3]) if len(sys.argv) > 3 else 8
  shape = PRESETS[name]
  dt = torch.float16 if shape.dtype == "float16" else torch.bfloat16
  t = time.time()
  sd = synth_state_dict_fast(shape, 1234)
  print(f"weights {time.time() - t:.1f} s", flush=True)
  prompt = synth_prompt(shape.vocab, n_prompt)
  ref = RefEngine(model_config_from_hf(hf_config(shape)), sd, RefShard(name, 0, shape.n_layers - 1, shape.n_layers))
  t = time.time()
  o_ids, o_logits, gaps = generate_ids([ref], prompt, n_dec + 1)
  print(f"oracle {time.time() - t:.1f} s  ids {o_ids.tolist()}", flush=True)
  from transformers import LlamaConfig, LlamaForCausalLM
  hc = {k: v for k, v in hf_config(shape).items() if k not in ("architectures", "model_type", "torch_dtype")}
  with torch.device("meta"):
    m = LlamaForCausalLM(LlamaConfig(**hc))
  full = dict(sd)
  if shape.tie:
    full["lm_head.weight"] = sd["model.embed_tokens.weight"]
  m.load_state_dict(full, assign=True, strict=True)
  m.model.rotary_emb = type(m.model.rotary_emb)(config=m.config)   # its inv_freq buffer is not in the state dict
  m = m.eval()
  t = time.time()
  toks = torch.tensor(np.concatenate([prompt[0], o_ids[:-1]])[None]).long()
  with torch.no_grad():
    # one causal pass over prompt + forced ids gives the logits of every teacher-forced step at once
    hf = m(toks).logits[0, n_prompt - 1:].float().numpy()
  print(f"hf {time.time() - t:.1f} s", flush=True)
  err = np.abs(hf - o_logits)
  std = float(o_logits.std())
  out = {"model": name, "n_prompt": n_prompt, "steps": n_dec + 1, "logit_std": std,
         "rms_over_std": float(np.sqrt((err ** 2).mean()) / std), "max_over_std": float(err.max() / std),
         "per_step_max_over_std": [round(float(e.max() / std), 4) for e in err],
         "argmax_equal": int((hf.argmax(-1) == o_logits.argmax(-1)).sum()), "oracle_top2_gaps_over_std": [round(float(g / std), 4) for g in gaps]}
  print(json.dumps(out))


if __name__ == "__main__":
  main()
