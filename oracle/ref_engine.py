"""TEST INFRASTRUCTURE — CPU oracle for the shard-decode hot path.  NOT part of the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this file
(prompt rule ③).  The product path (xotorch_b200/) never imports oracle/ and fails loudly without its CUDA
library.

What this is
------------
A torchtune-free restatement, in plain eager PyTorch on the CPU, of the reference's
`TorchDynamicShardInferenceEngine` path (all citations relative to /root/reference):

* config mapping            xotorch/inference/llm_utils.py:79-126          -> `load_model_config`
* weight key map + _permute xotorch/inference/llm_utils.py:175-183,230-318 -> `permute_qk`, `RefShardModel.__init__`
* module graph / RoPE pick  xotorch/inference/torch/models/general_mha.py:23-142
* decoder forward           xotorch/inference/llm_utils.py:416-489          -> `RefShardModel.forward`
* generate() slicing        xotorch/inference/torch/models/general_mha.py:178-254 -> `RefEngine._generate`
* engine state / ndim proto xotorch/inference/torch/sharded_inference_engine.py:108-370 -> `RefEngine`
* sampler                   xotorch/inference/torch/sharded_inference_engine.py:208-228 -> `RefEngine.sample`, `ref_sample`
* canonical generate loop   xotorch/inference/torch/tests/test_llama3_full.py:34-180 -> `generate_ids`

The arithmetic itself lives in a third-party dependency that is absent here (torchtune >= 0.6.1, un-pinned,
setup.py:41-68; `import torchtune` fails in this container and there is no network), so its published
algorithm is restated from SURVEY.md Appendix A: RMSNorm (A.1), Llama3ScaledRoPE / RotaryPositionalEmbeddings
(A.2), MultiHeadAttention with KVCache + GQA expand + masked SDPA over the whole cache (A.3-A.4),
TransformerSelfAttentionLayer / FeedForward (A.5), `torchtune.generation.sample` (A.8).

Pinning status ("how do we know the oracle is right")
-----------------------------------------------------
The reference holds NO golden logits / known-answer vectors for this path and its own
implementation cannot be imported, so absolute numerical parity with *the reference* is **unpinned**.  What is
pinned instead (tests/test_oracle.py, fixtures under tests/golden/, generator oracle/gen_golden.py):
  1. independent cross-check against HF transformers `LlamaForCausalLM` (half-split RoPE, un-permuted q/k) on
     the same synthetic weights, fp32 and bf16;
  2. the reference's own contract (xotorch/inference/test_inference_engine.py:11,43-44): sharded == unsharded,
     bit-exact;
  3. golden logits / ids of this oracle, committed, so later edits cannot drift silently.

Deliberate deviations from reference *defects* (SURVEY.md Appendix B): B1 (a shard hands off the output of its
last layer), B2/B3 (HF rope_scaling keys; un-scaled RoPE when rope_scaling is null so Llama-3-8B/70B build), B5
(position is explicit).  Cost structure kept faithful for the CPU baseline: bf16 weights, KV cache of
prompt+1024, full-length masked SDPA, per-step GQA K/V materialisation, fp32 logits, reference sampler at
temperature 0.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

TEMP = 0.6    # sharded_inference_engine.py:34
TOP_K = 35    # sharded_inference_engine.py:35
MAX_GENERATED_TOKENS = 1024  # general_mha.py:152

_DTYPES = {"bfloat16": torch.bfloat16, "float16": torch.float16, "float32": torch.float32,
           "bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}


def load_model_config(model_config_path) -> dict:
  """llm_utils.py:79-126, with B2 fixed: the HF rope_scaling keys are read as HF names them."""
  with open(model_config_path, "r") as f:
    base = json.load(f)
  return model_config_from_hf(base)


def model_config_from_hf(base: dict) -> dict:
  cfg = {
    "rope_scaling": base.get("rope_scaling"),
    "embed_dim": base["hidden_size"],
    "num_heads": base["num_attention_heads"],
    "head_dim": base.get("head_dim") or base["hidden_size"] // base["num_attention_heads"],
    "num_kv_heads": base["num_key_value_heads"],
    "max_seq_len": base["max_position_embeddings"],
    "intermediate_dim": base["intermediate_size"],
    "attn_dropout": base.get("attention_dropout", 0.0),
    "norm_eps": base["rms_norm_eps"],
    "rope_base": base["rope_theta"],
    "vocab_size": base["vocab_size"],
    "num_layers": base["num_hidden_layers"],
    "attn_bias": base.get("attention_bias", False),
    "hidden_act": base.get("hidden_act", "silu"),
    "torch_dtype": _DTYPES.get(base.get("torch_dtype", "bfloat16"), torch.bfloat16),
    "tie_word_embeddings": bool(base.get("tie_word_embeddings", False)),
    "num_experts": int(base.get("num_local_experts") or 0),       # Mixtral-style sparse MoE MLP (0 = dense)
    "experts_per_tok": int(base.get("num_experts_per_tok") or 0),
  }
  rs = cfg["rope_scaling"]
  if rs is not None:
    cfg["rope_scaling_factor"] = float(rs.get("factor", 32))
    cfg["rope_low_freq_factor"] = float(rs.get("low_freq_factor", 1))
    cfg["rope_high_freq_factor"] = float(rs.get("high_freq_factor", 4))
    cfg["rope_old_context_len"] = int(rs.get("original_max_position_embeddings", 8192))
  return cfg


def permute_qk(t: torch.Tensor, n_heads: int, head_dim: int, dim: int) -> torch.Tensor:
  """llm_utils.py:175-183 `_permute`: HF half-split rows -> torchtune interleaved-pair rows."""
  return t.view(n_heads, 2, head_dim // 2, dim).transpose(1, 2).reshape(head_dim * n_heads, dim)


# --------------------------------------------------------------------------------------------------------------
# torchtune semantics
# --------------------------------------------------------------------------------------------------------------
def rope_theta_llama3(dim: int, base: float, scale_factor: float, low_freq_factor: float = 1.0,
                      high_freq_factor: float = 4.0, old_context_len: int = 8192) -> torch.Tensor:
  """A.2 Llama3ScaledRoPE.rope_init/apply_scaling (fp32 frequencies, python-float smoothing)."""
  freqs = 1.0 / (base ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
  low_wl, high_wl = old_context_len / low_freq_factor, old_context_len / high_freq_factor
  out = []
  for freq in freqs:
    wavelen = 2 * math.pi / freq
    if wavelen < high_wl:
      out.append(freq)
    elif wavelen > low_wl:
      out.append(freq / scale_factor)
    else:
      smooth = (old_context_len / wavelen - low_freq_factor) / (high_freq_factor - low_freq_factor)
      out.append((1 - smooth) * freq / scale_factor + smooth * freq)
  return torch.tensor(out, dtype=freqs.dtype)


def rope_theta_plain(dim: int, base: float) -> torch.Tensor:
  return 1.0 / (base ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))


def rope_cache(theta: torch.Tensor, max_seq_len: int) -> torch.Tensor:
  """A.2 build_rope_cache: [max_seq_len, dim/2, 2] = (cos, sin)(pos * theta), fp32."""
  seq_idx = torch.arange(max_seq_len, dtype=theta.dtype)
  idx_theta = torch.einsum("i, j -> ij", seq_idx, theta).float()
  return torch.stack([torch.cos(idx_theta), torch.sin(idx_theta)], dim=-1)


def apply_rope_interleaved(x: torch.Tensor, cache: torch.Tensor, input_pos: torch.Tensor) -> torch.Tensor:
  """A.2 forward on x [b, s, n_h, hd]; pairs are adjacent elements (2j, 2j+1); indexed by input_pos."""
  rc = cache[input_pos]                                  # [s, hd/2, 2] (or [b, s, hd/2, 2])
  xs = x.float().reshape(*x.shape[:-1], -1, 2)
  rc = rc.view(-1, xs.size(1), 1, xs.size(3), 2)
  out = torch.stack([xs[..., 0] * rc[..., 0] - xs[..., 1] * rc[..., 1],
                     xs[..., 1] * rc[..., 0] + xs[..., 0] * rc[..., 1]], -1)
  return out.flatten(3).type_as(x)


def rms_norm(x: torch.Tensor, scale: torch.Tensor, eps: float, variant: str = "old") -> torch.Tensor:
  """A.1.  'old' (torchtune<=0.6 / HF LlamaRMSNorm order): fp32 normalise, round to x.dtype, then multiply by
  scale in x.dtype (second rounding).  'new': F.rms_norm in fp32 with the weight, single rounding."""
  xf = x.float()
  if variant == "old":
    xn = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x)
    return xn * scale
  return F.rms_norm(xf, (x.shape[-1],), weight=scale.float(), eps=eps).to(x.dtype)


class RefKVCache:
  """A.4 torchtune KVCache: [b, Hkv, L, hd] zeros, internal cache_pos counter independent of input_pos."""

  def __init__(self, batch: int, max_seq_len: int, n_kv: int, hd: int, dtype: torch.dtype):
    self.k = torch.zeros(batch, n_kv, max_seq_len, hd, dtype=dtype)
    self.v = torch.zeros(batch, n_kv, max_seq_len, hd, dtype=dtype)
    self.pos = 0

  def reset(self):
    self.k.zero_(); self.v.zero_(); self.pos = 0

  def update(self, k: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    s = k.shape[2]
    assert self.pos + s <= self.k.shape[2], "KV cache overflow (reference B7)"
    self.k[:, :, self.pos:self.pos + s] = k
    self.v[:, :, self.pos:self.pos + s] = v
    self.pos += s
    return self.k, self.v


@dataclass
class RefShard:
  """Value-compatible with xotorch/inference/shard.py:4-39 (kept local so oracle/ imports nothing else)."""
  model_id: str
  start_layer: int
  end_layer: int
  n_layers: int

  def is_first_layer(self): return self.start_layer == 0
  def is_last_layer(self): return self.end_layer == self.n_layers - 1


class RefShardModel:
  """general_mha.py:23-142 + llm_utils.py:335-489 restated.  `weights` uses HF names/layout; q/k are permuted
  to interleaved order exactly as the reference loader does for llama ids (llm_utils.py:262-282)."""

  def __init__(self, config: dict, weights: Dict[str, torch.Tensor], shard, dtype: Optional[torch.dtype] = None,
               norm_variant: str = "old"):
    self.cfg, self.shard = config, shard
    self.dtype = dtype or config["torch_dtype"]
    self.norm_variant = norm_variant
    c = config
    D, Hq, Hkv, hd = c["embed_dim"], c["num_heads"], c["num_kv_heads"], c["head_dim"]
    self.D, self.Hq, self.Hkv, self.hd = D, Hq, Hkv, hd
    # RoPE selection: general_mha.py:33-63 (llama ids -> Llama3ScaledRoPE); B3: unscaled when rope_scaling is null
    if c.get("rope_scaling") is not None:
      theta = rope_theta_llama3(hd, c["rope_base"], c["rope_scaling_factor"], c.get("rope_low_freq_factor", 1.0),
                                c.get("rope_high_freq_factor", 4.0), c.get("rope_old_context_len", 8192))
    else:
      theta = rope_theta_plain(hd, c["rope_base"])
    self.theta = theta
    self._rope_cache = None
    self._rope_len = 0
    dt = self.dtype
    g = lambda name: weights[name].to(dt)
    self.layers: Dict[int, dict] = {}
    for i in range(shard.start_layer, shard.end_layer + 1):
      p = f"model.layers.{i}."
      self.layers[i] = {
        "sa_norm": g(p + "input_layernorm.weight"),
        "q": permute_qk(g(p + "self_attn.q_proj.weight"), Hq, hd, D).contiguous(),
        "k": permute_qk(g(p + "self_attn.k_proj.weight"), Hkv, hd, D).contiguous(),
        "v": g(p + "self_attn.v_proj.weight"),
        "o": g(p + "self_attn.o_proj.weight"),
        "mlp_norm": g(p + "post_attention_layernorm.weight"),
      }
      if c.get("num_experts", 0):
        m = p + "block_sparse_moe."
        self.layers[i]["router"] = g(m + "gate.weight")
        self.layers[i]["experts"] = [(g(m + f"experts.{e}.w1.weight"), g(m + f"experts.{e}.w2.weight"), g(m + f"experts.{e}.w3.weight"))
                                     for e in range(c["num_experts"])]
      else:
        self.layers[i].update({"w1": g(p + "mlp.gate_proj.weight"), "w2": g(p + "mlp.down_proj.weight"),
                               "w3": g(p + "mlp.up_proj.weight")})
    # B6: the reference allocates embedding + head on every shard; only first/last need them.
    self.tok_embeddings = g("model.embed_tokens.weight") if shard.is_first_layer() else None
    if shard.is_last_layer():
      self.norm = g("model.norm.weight")
      tied = c.get("tie_word_embeddings", False) or "lm_head.weight" not in weights
      self.output = g("model.embed_tokens.weight") if tied else g("lm_head.weight")
    self.caches: Dict[int, RefKVCache] = {}
    self.decoder_max_cache_seq_len = None

  # llm_utils.py:369-414
  def setup_caches(self, batch_size: int, max_seq_len: int):
    self.decoder_max_cache_seq_len = max_seq_len
    self.caches = {i: RefKVCache(batch_size, max_seq_len, self.Hkv, self.hd, self.dtype) for i in self.layers}

  def caches_are_enabled(self) -> bool:
    return bool(self.caches)

  def reset_caches(self):
    for c in self.caches.values():
      c.reset()

  def _rope(self, n: int) -> torch.Tensor:
    if self._rope_cache is None or self._rope_len < n:
      self._rope_len = max(n, 2048)
      self._rope_cache = rope_cache(self.theta, self._rope_len)
    return self._rope_cache

  def _attention(self, i: int, x: torch.Tensor, mask: torch.Tensor, input_pos: torch.Tensor) -> torch.Tensor:
    """A.3 MultiHeadAttention.forward with kv cache."""
    L = self.layers[i]
    b, s, _ = x.shape
    Hq, Hkv, hd = self.Hq, self.Hkv, self.hd
    rc = self._rope(int(input_pos.max()) + 1)
    q = F.linear(x, L["q"]).view(b, s, Hq, hd)
    q = apply_rope_interleaved(q, rc, input_pos).transpose(1, 2)
    k = F.linear(x, L["k"]).view(b, s, Hkv, hd)
    v = F.linear(x, L["v"]).view(b, s, Hkv, hd)
    k = apply_rope_interleaved(k, rc, input_pos).transpose(1, 2)
    v = v.transpose(1, 2)
    k, v = self.caches[i].update(k, v)                     # returns the ENTIRE cache [b, Hkv, L, hd]
    if Hq != Hkv:                                          # K5: materialised GQA expand
      r = Hq // Hkv
      k = k.unsqueeze(2).expand(b, Hkv, r, k.shape[2], hd).flatten(1, 2)
      v = v.unsqueeze(2).expand(b, Hkv, r, v.shape[2], hd).flatten(1, 2)
    out = F.scaled_dot_product_attention(q, k, v, attn_mask=mask[:, None, :, :], dropout_p=0.0, is_causal=False)
    out = out.transpose(1, 2).contiguous().view(b, s, Hq * hd)
    return F.linear(out, L["o"])

  def _layer(self, i: int, h: torch.Tensor, mask, input_pos) -> torch.Tensor:
    """A.5 TransformerSelfAttentionLayer + FeedForward (llm_utils.py:491-500)."""
    L = self.layers[i]
    eps = self.cfg["norm_eps"]
    xn = rms_norm(h, L["sa_norm"], eps, self.norm_variant)
    h = self._attention(i, xn, mask, input_pos) + h
    xm = rms_norm(h, L["mlp_norm"], eps, self.norm_variant)
    if "router" in L:
      return h + self._moe(L, xm, i, input_pos)
    mlp = F.linear(F.silu(F.linear(xm, L["w1"])) * F.linear(xm, L["w3"]), L["w2"])
    return h + mlp

  # Routing is a DISCRETE choice on bf16 router logits: two correct implementations whose activations differ by one bf16 ulp
  # can pick different experts when the 2nd and 3rd logit are (nearly) tied.  For parity tests the choice of another
  # implementation can be imposed (`forced_routes[(layer, position)] = (lo, hi)`); `route_log` then records this oracle's own
  # choice and how far (in router-logit units) the imposed pair is from being its own — a test asserts that is ~1 ulp.
  forced_routes: Optional[dict] = None
  route_log: Optional[list] = None

  def _moe(self, L: dict, xm: torch.Tensor, layer: int = -1, input_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Mixtral sparse-MoE MLP.  The reference has no MoE path (its `MOELayer`, llm_utils.py:502-590, is an unused stub:
    SURVEY.md §0.10), so this restates the third-party algorithm the config names: HF transformers 5.5.0
    models/mixtral/modeling_mixtral.py:101-116 (MixtralTopKRouter: linear -> fp32 softmax -> top-k -> renormalise) and
    :74-98 (MixtralExperts: per expert in ascending index, down(silu(gate x) * up x) * fp32 routing weight -> model dtype ->
    index_add_ into a model-dtype accumulator).  One deliberate pin: torch.topk leaves the order of EQUAL values
    unspecified; here (and in the CUDA kernel) ties go to the lower expert index."""
    b, s, D = xm.shape
    T = xm.reshape(-1, D)
    k = self.cfg["experts_per_tok"]
    probs = F.softmax(F.linear(T, L["router"]).float(), dim=-1)
    order = torch.sort(probs, dim=-1, descending=True, stable=True)
    top_v, top_i = order.values[:, :k], order.indices[:, :k]
    if self.forced_routes is not None or self.route_log is not None:
      logits = F.linear(T, L["router"]).float()
      pos_list = [int(p) for p in input_pos.reshape(-1)]
      top_i = top_i.clone()
      for t, pos in enumerate(pos_list):
        own = sorted(int(x) for x in top_i[t])
        used = own
        if self.forced_routes is not None and (layer, pos) in self.forced_routes:
          used = sorted(int(x) for x in self.forced_routes[(layer, pos)])
          top_i[t] = torch.tensor(used)
        if self.route_log is not None:
          thr = float(order.values[t, k - 1].log() - order.values[t, 0].log() + logits[t].max())   # own k-th largest logit
          self.route_log.append((layer, pos, tuple(own), tuple(used), max(0.0, thr - float(logits[t, used].min()))))
      top_v = torch.gather(probs, 1, top_i)
    top_v = top_v / top_v.sum(dim=-1, keepdim=True)
    final = torch.zeros_like(T)
    for e, (w1, w2, w3) in enumerate(L["experts"]):
      tok, slot = torch.where(top_i == e)
      if tok.numel() == 0:
        continue
      cur = T[tok]
      cur = F.linear(F.silu(F.linear(cur, w1)) * F.linear(cur, w3), w2)
      cur = cur * top_v[tok, slot, None]                  # fp32 weight: the product is fp32 ...
      final.index_add_(0, tok, cur.to(final.dtype))       # ... rounded to the model dtype, accumulated in the model dtype
    return final.view(b, s, D)

  @torch.inference_mode()
  def forward(self, tokens: Optional[torch.Tensor], *, mask: torch.Tensor, input_pos: torch.Tensor,
              hidden_state: Optional[torch.Tensor] = None) -> torch.Tensor:
    """llm_utils.py:416-489.  B1: returns the OUTPUT of the shard's last layer (the reference captures the
    input of layer `end_layer` by local index and drops a layer)."""
    if hidden_state is not None:
      h = hidden_state
    else:
      h = F.embedding(tokens.long(), self.tok_embeddings).to(self.dtype)
    if input_pos.ndim == 0:
      input_pos = input_pos[None]
    for i in range(self.shard.start_layer, self.shard.end_layer + 1):
      h = self._layer(i, h, mask, input_pos)
    if self.shard.is_last_layer():
      h = rms_norm(h, self.norm, self.cfg["norm_eps"], self.norm_variant)
      return F.linear(h, self.output)
    return h


def ref_sample(logits: torch.Tensor, temperature: float = 1.0, top_k: Optional[int] = None,
               q: Optional[torch.Tensor] = None) -> torch.Tensor:
  """A.8 torchtune.generation.sample."""
  logits = logits / max(temperature, 1e-5)
  if top_k is not None:
    v, _ = torch.topk(logits, min(top_k, logits.size(-1)))
    pivot = v.select(-1, -1).unsqueeze(-1)
    logits = torch.where(logits < pivot, -float("Inf"), logits)
  probs = F.softmax(logits, dim=-1)
  if q is None:
    q = torch.empty_like(probs).exponential_(1)
  return torch.argmax(probs / q, dim=-1, keepdim=True).to(dtype=torch.int)


class RefEngine:
  """Synchronous mirror of TorchDynamicShardInferenceEngine (sharded_inference_engine.py:37-370) minus the
  tokenizer/downloader: prompts are token ids (SURVEY.md §8d feeds `infer_tensor` ids directly)."""

  def __init__(self, config: dict, weights: Dict[str, torch.Tensor], shard, dtype=None, norm_variant="old",
               max_generated_tokens: int = MAX_GENERATED_TOKENS):
    self.cfg = config
    self.shard = shard
    self.model = RefShardModel(config, weights, shard, dtype, norm_variant)
    self.max_generated_tokens = max_generated_tokens
    self.rng = torch.Generator(device="cpu")
    self.rng.manual_seed(1234)                              # sharded_inference_engine.py:67-69
    self.tokens = None
    self.input_pos = None
    self.mask = None
    self.curr_pos = 0
    self.cache_setup = False

  # sharded_inference_engine.py:108-193 `encode` minus the tokenizer
  def begin(self, n_prompt_tokens: int):
    self.tokens = None
    self.curr_pos = 0
    total = n_prompt_tokens + self.max_generated_tokens
    if self.model.caches_are_enabled():
      self.model.reset_caches()
    if not self.cache_setup or self.model.decoder_max_cache_seq_len != total:
      self.model.setup_caches(1, total)
      self.cache_setup = True
    max_seq_len = self.model.decoder_max_cache_seq_len
    self.mask = torch.tril(torch.ones(total, max_seq_len, dtype=torch.bool)).unsqueeze(0)   # :179-184
    self.input_pos = torch.arange(0, total).unsqueeze(0)                                     # :186

  # general_mha.py:178-254
  def _generate(self, tokens, hidden_state, seq_len: int):
    if self.curr_pos > 0:
      input_pos = self.input_pos[:, self.curr_pos].contiguous()
      mask = self.mask[:, self.curr_pos, None, :].contiguous()
    else:
      mask = self.mask[:, :seq_len]
      input_pos = self.input_pos[:, :seq_len].squeeze(0)
    return self.model.forward(tokens, mask=mask, input_pos=input_pos, hidden_state=hidden_state)

  # sharded_inference_engine.py:230-370
  def infer_tensor(self, request_id: str, shard, input_data: np.ndarray,
                   inference_state: Optional[dict] = None) -> Tuple[np.ndarray, Optional[dict]]:
    # Position protocol (B5/B11): only `curr_pos` crosses the hop.  As in the reference (:356-359) only the LAST
    # shard advances it; the other shards pass through the position the step started at.  None => new request.
    self.curr_pos = int(inference_state["curr_pos"]) if inference_state and "curr_pos" in inference_state else 0
    hidden_state = None
    input_tensor = None
    if input_data.ndim == 3:
      hidden_state = torch.tensor(input_data).to(dtype=self.model.dtype)   # :254-258
      seq_len = hidden_state.shape[1]
    else:
      input_tensor = torch.tensor(input_data).to(dtype=torch.int)          # :259-263
      seq_len = input_tensor.shape[1]
    if self.curr_pos == 0:
      self.begin(seq_len)                                                    # prefill: reset state + caches
    if input_tensor is not None:
      self.tokens = input_tensor.clone() if self.tokens is None else torch.cat([self.tokens, input_tensor], dim=-1)
    out = self._generate(input_tensor, hidden_state, seq_len)
    if not self.shard.is_last_layer():
      return out.float().numpy(force=True), {"curr_pos": self.curr_pos}     # :351-354
    self.curr_pos = seq_len if self.curr_pos == 0 else self.curr_pos + 1    # :356-359
    return out[:, -1].float().numpy(force=True), {"curr_pos": self.curr_pos}  # :365-368

  # sharded_inference_engine.py:208-228
  def sample(self, x: np.ndarray, temp=TEMP, top_k=TOP_K) -> np.ndarray:
    logits = torch.tensor(x)
    q = torch.empty((logits.size(0), self.cfg["vocab_size"])).exponential_(1, generator=self.rng)
    return ref_sample(logits.clone(), temperature=temp, top_k=top_k, q=q).numpy(force=True)


def greedy_argmax(logits: np.ndarray) -> np.ndarray:
  """Lowest-index argmax over the (bf16-valued) fp32 logits: the engine's on-device 'greedy'."""
  return np.argmax(logits, axis=-1).astype(np.int32).reshape(-1, 1)


def generate_ids(engines: List[RefEngine], prompt_ids: np.ndarray, n_gen: int, sampler: str = "argmax",
                 force_ids: Optional[np.ndarray] = None, collect_logits: bool = True):
  """test_llama3_full.py:34-180-style loop over a ring of shard engines (1 engine = unsharded).
  sampler: 'argmax' (lowest index) or 'reference' (temp=0.0 path of node.py:121).  force_ids: teacher forcing.
  Returns (ids [n_gen], logits list [n_gen x V] or None, top2_gaps [n_gen])."""
  x = prompt_ids.astype(np.int32).reshape(1, -1)
  state = None
  ids, all_logits, gaps = [], [], []
  for step in range(n_gen):
    data = x
    for e in engines:
      data, state = e.infer_tensor("oracle", e.shard, data, state)
    logits = data                                           # [1, V] f32
    top2 = np.partition(logits[0], -2)[-2:]
    gaps.append(float(top2[1] - top2[0]))
    if collect_logits:
      all_logits.append(logits[0].copy())
    if sampler == "reference":
      tok = engines[-1].sample(logits, temp=0.0)
    else:
      tok = greedy_argmax(logits)
    if force_ids is not None:
      tok = np.array([[int(force_ids[step])]], dtype=np.int32)
    ids.append(int(tok[0, 0]))
    x = tok.astype(np.int32).reshape(1, 1)
  return np.array(ids, dtype=np.int32), (np.stack(all_logits) if collect_logits else None), np.array(gaps)


def load_hf_dir_weights(model_dir) -> Dict[str, torch.Tensor]:
  """Read every *.safetensors in the directory fully (llm_utils.py:205-218)."""
  from safetensors.torch import load_file
  out: Dict[str, torch.Tensor] = {}
  for f in sorted(Path(model_dir).glob("*.safetensors")):
    out.update(load_file(str(f)))
  return out
