"""Synthetic Llama-style model directories (HF layout) with counter-based, bit-reproducible weights.

No real checkpoints exist on the build or GPU boxes, so every parity and bench run
uses weights from this generator.  The generator is *counter based* (splitmix64 of
``seed, tensor_id, flat_index``) and uses only integer arithmetic plus one fp32 multiply and one fp32 add,
so the very same values can be produced on the host (here, torch int64 ops) and on the device
(``xotb200_init_synthetic`` in csrc/xotb200.cu) bit for bit, without writing 16 GB of safetensors for the
8B bench.  Values are Irwin-Hall(4) ~ N(mean, std^2), rounded to bf16/fp16 (round-nearest-even).

The HF tensor names / shapes follow what the reference's loader consumes
(/root/reference/xotorch/inference/llm_utils.py:230-318, key map in SURVEY.md App. A.6).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, asdict
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

# tensor "kinds" inside one decoder layer; tensor_id = layer*16 + kind
K_ATTN_NORM, K_Q, K_K, K_V, K_O, K_MLP_NORM, K_GATE, K_UP, K_DOWN = range(9)
G_EMBED, G_NORM, G_HEAD = (1 << 20), (1 << 20) + 1, (1 << 20) + 2
# mixture-of-experts tensors of a layer: tensor_id = G_MOE + layer*64 + {0: router, 1 + 3*e + {0: w1 gate, 1: w3 up, 2: w2 down}}
G_MOE = 1 << 21

_M64 = (1 << 64) - 1
_C_SEED = 0x9E3779B97F4A7C15
_C_TID = 0xD1B54A32D192ED03
_C_M1 = 0xBF58476D1CE4E5B9
_C_M2 = 0x94D049BB133111EB
# std of the sum of four independent uniform{0..65535}
IH4_STD = float(np.sqrt(4.0 * (65536.0**2 - 1.0) / 12.0))


def _i64(v: int) -> int:
  """two's-complement int64 view of an unsigned 64-bit constant"""
  v &= _M64
  return v - (1 << 64) if v >= (1 << 63) else v


def tensor_base(seed: int, tensor_id: int) -> int:
  """64-bit counter base for a tensor (host mirror of the device computation)."""
  return (seed * _C_SEED + tensor_id * _C_TID) & _M64


def synth_values_f32(seed: int, tensor_id: int, start: int, count: int, std: float, mean: float) -> torch.Tensor:
  """fp32 values for flat HF indices [start, start+count) of a tensor (before the dtype rounding)."""
  base = _i64(tensor_base(seed, tensor_id))
  z = torch.arange(start, start + count, dtype=torch.int64) + base  # wraps mod 2^64
  z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _i64(_C_M1)
  z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _i64(_C_M2)
  z = z ^ ((z >> 31) & ((1 << 33) - 1))
  s = (z & 0xFFFF) + ((z >> 16) & 0xFFFF) + ((z >> 32) & 0xFFFF) + ((z >> 48) & 0xFFFF)
  v = (s - 131070).to(torch.float32)
  scale = torch.tensor(np.float32(std / IH4_STD))
  return v * scale + torch.tensor(np.float32(mean))


def synth_tensor(seed: int, tensor_id: int, shape: Tuple[int, ...], std: float, mean: float,
                 dtype: torch.dtype = torch.bfloat16, chunk: int = 1 << 24) -> torch.Tensor:
  n = int(np.prod(shape))
  out = torch.empty(n, dtype=dtype)
  for s in range(0, n, chunk):
    c = min(chunk, n - s)
    out[s:s + c] = synth_values_f32(seed, tensor_id, s, c, std, mean).to(dtype)
  return out.view(*shape)


@dataclass
class ModelShape:
  """Llama-style decoder shape (HF config.json field names in the comments)."""
  name: str
  n_layers: int          # num_hidden_layers
  dim: int               # hidden_size
  ffn: int               # intermediate_size
  n_heads: int           # num_attention_heads
  n_kv_heads: int        # num_key_value_heads
  head_dim: int
  vocab: int             # vocab_size
  rope_theta: float = 500000.0
  norm_eps: float = 1e-5
  max_pos: int = 8192    # max_position_embeddings
  tie: bool = False      # tie_word_embeddings
  rope_scaling: Optional[dict] = None
  dtype: str = "bfloat16"
  n_experts: int = 0     # num_local_experts (0 = dense SwiGLU MLP)
  top_k: int = 0         # num_experts_per_tok


_LLAMA32_SCALING = {"factor": 32.0, "high_freq_factor": 4.0, "low_freq_factor": 1.0,
                    "original_max_position_embeddings": 8192, "rope_type": "llama3"}

# Constants from SURVEY.md Appendix C.
PRESETS: Dict[str, ModelShape] = {
  # tiny parity shapes (oracle runs in milliseconds)
  "tiny-llama-hd64": ModelShape("tiny-llama-hd64", 4, 512, 1536, 8, 2, 64, 2048, max_pos=2048),
  "tiny-llama-hd128": ModelShape("tiny-llama-hd128", 3, 1024, 3072, 8, 2, 128, 4096, max_pos=2048),
  "tiny-llama-hd64-f16": ModelShape("tiny-llama-hd64-f16", 4, 512, 1536, 8, 2, 64, 2048, max_pos=2048, dtype="float16"),
  "tiny-llama-hd128-f16": ModelShape("tiny-llama-hd128-f16", 3, 1024, 3072, 8, 2, 128, 4096, max_pos=2048, dtype="float16"),
  "tiny-llama-3.2": ModelShape("tiny-llama-3.2", 4, 512, 1536, 8, 2, 64, 2048, max_pos=16384, tie=True,
                               rope_scaling=dict(_LLAMA32_SCALING)),
  "llama-3.2-1b": ModelShape("llama-3.2-1b", 16, 2048, 8192, 32, 8, 64, 128256, max_pos=131072, tie=True,
                             rope_scaling=dict(_LLAMA32_SCALING)),
  "llama-3-8b": ModelShape("llama-3-8b", 32, 4096, 14336, 32, 8, 128, 128256),
  "llama-3-70b": ModelShape("llama-3-70b", 80, 8192, 28672, 64, 8, 128, 128256, dtype="float16"),
  # Mixtral-style sparse MoE MLP: router + top-2 of 8 SwiGLU experts
  "tiny-mixtral": ModelShape("tiny-mixtral", 4, 512, 1024, 8, 2, 64, 2048, rope_theta=1e6, max_pos=2048, n_experts=8, top_k=2),
  "tiny-mixtral-hd128": ModelShape("tiny-mixtral-hd128", 3, 1024, 2048, 8, 2, 128, 4096, rope_theta=1e6, max_pos=2048,
                                   n_experts=4, top_k=2),
  "tiny-mixtral-f16": ModelShape("tiny-mixtral-f16", 4, 512, 1024, 8, 2, 64, 2048, rope_theta=1e6, max_pos=2048, n_experts=8,
                                 top_k=2, dtype="float16"),
  "mixtral-8x7b": ModelShape("mixtral-8x7b", 32, 4096, 14336, 32, 8, 128, 32000, rope_theta=1e6, max_pos=32768,
                             n_experts=8, top_k=2),
  # same dimensions with bf16 weights: what the engine runs today (fp16 weights are not implemented yet)
  "llama-3-70b-bf16": ModelShape("llama-3-70b-bf16", 80, 8192, 28672, 64, 8, 128, 128256),
}


def hf_config(shape: ModelShape) -> dict:
  if shape.n_experts:
    return {
      "architectures": ["MixtralForCausalLM"], "model_type": "mixtral", "hidden_act": "silu",
      "hidden_size": shape.dim, "intermediate_size": shape.ffn, "num_hidden_layers": shape.n_layers,
      "num_attention_heads": shape.n_heads, "num_key_value_heads": shape.n_kv_heads, "head_dim": shape.head_dim,
      "vocab_size": shape.vocab, "rms_norm_eps": shape.norm_eps, "rope_theta": shape.rope_theta, "rope_scaling": None,
      "max_position_embeddings": shape.max_pos, "tie_word_embeddings": False, "torch_dtype": shape.dtype,
      "attention_dropout": 0.0, "bos_token_id": 1, "eos_token_id": 2, "sliding_window": None,
      "num_local_experts": shape.n_experts, "num_experts_per_tok": shape.top_k, "router_jitter_noise": 0.0,
    }
  return {
    "architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_act": "silu",
    "hidden_size": shape.dim, "intermediate_size": shape.ffn, "num_hidden_layers": shape.n_layers,
    "num_attention_heads": shape.n_heads, "num_key_value_heads": shape.n_kv_heads, "head_dim": shape.head_dim,
    "vocab_size": shape.vocab, "rms_norm_eps": shape.norm_eps, "rope_theta": shape.rope_theta,
    "rope_scaling": shape.rope_scaling, "max_position_embeddings": shape.max_pos,
    "tie_word_embeddings": shape.tie, "torch_dtype": shape.dtype, "attention_bias": False,
    "attention_dropout": 0.0, "bos_token_id": 1, "eos_token_id": 2,
  }


def shape_from_hf_config(cfg: dict, name: str = "model") -> ModelShape:
  return ModelShape(
    name=name, n_layers=cfg["num_hidden_layers"], dim=cfg["hidden_size"], ffn=cfg["intermediate_size"],
    n_heads=cfg["num_attention_heads"], n_kv_heads=cfg["num_key_value_heads"],
    head_dim=cfg.get("head_dim") or cfg["hidden_size"] // cfg["num_attention_heads"], vocab=cfg["vocab_size"],
    rope_theta=float(cfg.get("rope_theta", 10000.0)), norm_eps=float(cfg["rms_norm_eps"]),
    max_pos=int(cfg["max_position_embeddings"]), tie=bool(cfg.get("tie_word_embeddings", False)),
    rope_scaling=cfg.get("rope_scaling"), dtype=str(cfg.get("torch_dtype", "bfloat16")),
    n_experts=int(cfg.get("num_local_experts") or 0), top_k=int(cfg.get("num_experts_per_tok") or 0))


W_STD = 0.02  # weights ~ N(0, 0.02^2), norm scales = 1 + N(0, 0.02^2)


def hf_tensor_specs(shape: ModelShape, layers: Optional[Iterable[int]] = None,
                    with_embed: bool = True, with_head: bool = True) -> List[Tuple[str, int, Tuple[int, ...], float, float]]:
  """(hf_name, tensor_id, shape, std, mean) for every tensor of the model (or a layer subset)."""
  D, F, Hq, Hkv, hd = shape.dim, shape.ffn, shape.n_heads, shape.n_kv_heads, shape.head_dim
  specs = []
  if with_embed:
    specs.append(("model.embed_tokens.weight", G_EMBED, (shape.vocab, D), W_STD, 0.0))
  for l in (range(shape.n_layers) if layers is None else layers):
    p, t = f"model.layers.{l}.", l * 16
    specs += [
      (p + "input_layernorm.weight", t + K_ATTN_NORM, (D,), W_STD, 1.0),
      (p + "self_attn.q_proj.weight", t + K_Q, (Hq * hd, D), W_STD, 0.0),
      (p + "self_attn.k_proj.weight", t + K_K, (Hkv * hd, D), W_STD, 0.0),
      (p + "self_attn.v_proj.weight", t + K_V, (Hkv * hd, D), W_STD, 0.0),
      (p + "self_attn.o_proj.weight", t + K_O, (D, Hq * hd), W_STD, 0.0),
      (p + "post_attention_layernorm.weight", t + K_MLP_NORM, (D,), W_STD, 1.0),
    ]
    if shape.n_experts:   # on-disk names of the Mixtral checkpoints (w1 = gate, w3 = up, w2 = down)
      m, tm = p + "block_sparse_moe.", G_MOE + l * 64
      specs.append((m + "gate.weight", tm, (shape.n_experts, D), W_STD, 0.0))
      for e in range(shape.n_experts):
        specs += [
          (m + f"experts.{e}.w1.weight", tm + 1 + 3 * e, (F, D), W_STD, 0.0),
          (m + f"experts.{e}.w3.weight", tm + 2 + 3 * e, (F, D), W_STD, 0.0),
          (m + f"experts.{e}.w2.weight", tm + 3 + 3 * e, (D, F), W_STD, 0.0),
        ]
    else:
      specs += [
        (p + "mlp.gate_proj.weight", t + K_GATE, (F, D), W_STD, 0.0),
        (p + "mlp.up_proj.weight", t + K_UP, (F, D), W_STD, 0.0),
        (p + "mlp.down_proj.weight", t + K_DOWN, (D, F), W_STD, 0.0),
      ]
  if with_head:
    specs.append(("model.norm.weight", G_NORM, (D,), W_STD, 1.0))
    if not shape.tie:
      specs.append(("lm_head.weight", G_HEAD, (shape.vocab, D), W_STD, 0.0))
  return specs


def torch_dtype(name: str) -> torch.dtype:
  return {"bfloat16": torch.bfloat16, "float16": torch.float16, "float32": torch.float32}[name]


def synth_state_dict(shape: ModelShape, seed: int = 1234, layers: Optional[Iterable[int]] = None,
                     with_embed: bool = True, with_head: bool = True) -> Dict[str, torch.Tensor]:
  dt = torch_dtype(shape.dtype)
  return {name: synth_tensor(seed, tid, shp, std, mean, dt)
          for name, tid, shp, std, mean in hf_tensor_specs(shape, layers, with_embed, with_head)}


def write_hf_dir(path: os.PathLike, shape: ModelShape, seed: int = 1234, n_files: int = 2) -> Path:
  """Write config.json + model-0000i-of-0000n.safetensors (+ index) in HF layout.  Returns the directory."""
  from safetensors.torch import save_file
  path = Path(path)
  path.mkdir(parents=True, exist_ok=True)
  with open(path / "config.json", "w") as f:
    json.dump(hf_config(shape), f, indent=1)
  sd = synth_state_dict(shape, seed)
  names = list(sd.keys())
  per = (len(names) + n_files - 1) // n_files
  weight_map = {}
  for i in range(n_files):
    part = names[i * per:(i + 1) * per]
    if not part:
      continue
    fn = f"model-{i + 1:05d}-of-{n_files:05d}.safetensors"
    save_file({k: sd[k].contiguous() for k in part}, str(path / fn), metadata={"format": "pt"})
    weight_map.update({k: fn for k in part})
  with open(path / "model.safetensors.index.json", "w") as f:
    json.dump({"metadata": {"seed": seed, "shape": asdict(shape)}, "weight_map": weight_map}, f, indent=1)
  return path


def synth_prompt(vocab: int, n: int = 128, seed: int = 42) -> np.ndarray:
  """prompt = n random ids from Generator(seed 42); returned as int32 [1, n]."""
  g = torch.Generator().manual_seed(seed)
  return torch.randint(0, vocab, (1, n), generator=g).to(torch.int32).numpy()
