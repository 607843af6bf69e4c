"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/xotb200.h declares
(no compute without a GPU), error paths, the ctypes struct layout, the synthetic generator twins."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from xotorch_b200 import _lib
from xotorch_b200.native import cfg_from_hf
from xotorch_b200.synth import PRESETS, hf_config, synth_tensor, tensor_base, IH4_STD

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
  hdr = (ROOT / "include" / "xotb200.h").read_text()
  declared = set(re.findall(r"\b(xotb200_[a-z_0-9]+)\s*\(", hdr))
  assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
  lib = _lib.load()
  for s in declared:
    assert getattr(lib, s) is not None
  assert b"sm_100a" in lib.xotb200_version()


def test_model_cfg_struct_matches_header_order():
  hdr = (ROOT / "include" / "xotb200.h").read_text()
  body = hdr[hdr.index("typedef struct xotb200_model_cfg {"):hdr.index("} xotb200_model_cfg;")]
  fields = re.findall(r"^\s*(?:int32_t|float)\s+([a-z_]+);", body, flags=re.M)
  assert fields == [f[0] for f in _lib.ModelCfg._fields_]
  assert C.sizeof(_lib.ModelCfg) == 4 * len(fields)


def test_create_fails_loudly_without_gpu_or_with_bad_cfg():
  lib = _lib.load()
  cfg = cfg_from_hf(hf_config(PRESETS["tiny-llama-hd64"]))
  cfg.head_dim = 96
  h = C.c_void_p()
  rc = lib.xotb200_create(C.byref(cfg), 0, C.byref(h))
  assert rc < 0 and b"head_dim" in lib.xotb200_last_error()
  if not torch.cuda.is_available():
    cfg = cfg_from_hf(hf_config(PRESETS["tiny-llama-hd64"]))
    rc = lib.xotb200_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -2, "without a CUDA device the library must fail (no CPU fallback)"
    with pytest.raises(_lib.XotB200Error):
      _lib.check(rc)


def test_cfg_from_hf_follows_reference_mapping():
  c = cfg_from_hf(hf_config(PRESETS["llama-3-8b"]))
  assert (c.n_layers, c.dim, c.ffn, c.n_heads, c.n_kv_heads, c.head_dim, c.vocab) == (32, 4096, 14336, 32, 8, 128, 128256)
  assert c.rope_kind == _lib.ROPE_PLAIN and c.tie_embeddings == 0          # App. B3: no rope_scaling -> unscaled
  c = cfg_from_hf(hf_config(PRESETS["llama-3.2-1b"]))
  assert c.rope_kind == _lib.ROPE_LLAMA3 and c.rope_factor == 32.0 and c.tie_embeddings == 1
  assert cfg_from_hf(hf_config(PRESETS["llama-3-70b"])).dtype == _lib.DTYPE_F16          # the upstream 70B card is fp16
  with pytest.raises(ValueError):
    cfg_from_hf({**hf_config(PRESETS["llama-3-8b"]), "torch_dtype": "float32"})


def test_synth_generators_agree():
  from oracle.synth_fast import synth_tensor_fast
  a = synth_tensor(1234, 37, (256, 512), 0.02, 0.0)
  b = synth_tensor_fast(1234, 37, (256, 512), 0.02, 0.0)
  assert torch.equal(a, b)
  for std in (0.02, 2e-5):   # IEEE-half twin of the generator (fp16 models), including subnormal halves
    ah = synth_tensor(1234, 38, (128, 512), std, 0.0, torch.float16)
    bh = synth_tensor_fast(1234, 38, (128, 512), std, 0.0, torch.float16)
    assert torch.equal(ah.view(torch.int16), bh.view(torch.int16))
  n = synth_tensor(7, (1 << 20) + 1, (512,), 0.02, 1.0).float()
  assert abs(float(n.mean()) - 1.0) < 0.01 and 0.015 < float(n.std()) < 0.025
  assert tensor_base(1, 1) == (0x9E3779B97F4A7C15 + 0xD1B54A32D192ED03) % (1 << 64)
  assert abs(IH4_STD - 37837.2) < 0.1
