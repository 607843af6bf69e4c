"""ctypes binding of libxotb200.so (include/xotb200.h).  This is the stub a maintainer of the reference would add
(see INTEGRATION.md).  There is NO fallback: if the CUDA library is missing or fails to load, importing the
engine raises — the product path never runs on the CPU or through oracle/."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libxotb200.so"

IPC_HANDLE_BYTES = 64
DTYPE_BF16, DTYPE_F16 = 0, 1
ROPE_PLAIN, ROPE_LLAMA3 = 0, 1
E_TIMEOUT = -5


class ModelCfg(C.Structure):
  """xotb200_model_cfg — mirror of the dict built by load_model_config
  (/root/reference/xotorch/inference/llm_utils.py:79-126)."""
  _fields_ = [
    ("n_layers", C.c_int32), ("dim", C.c_int32), ("ffn", C.c_int32), ("n_heads", C.c_int32),
    ("n_kv_heads", C.c_int32), ("head_dim", C.c_int32), ("vocab", C.c_int32), ("max_pos", C.c_int32),
    ("rope_theta", C.c_float), ("norm_eps", C.c_float), ("tie_embeddings", C.c_int32), ("dtype", C.c_int32),
    ("rope_kind", C.c_int32), ("rope_factor", C.c_float), ("rope_low_freq_factor", C.c_float),
    ("rope_high_freq_factor", C.c_float), ("rope_old_context_len", C.c_int32),
    ("n_experts", C.c_int32), ("experts_per_tok", C.c_int32),
  ]


class XotB200Error(RuntimeError):
  def __init__(self, code: int, msg: str):
    super().__init__(f"xotb200 error {code}: {msg}")
    self.code = code


_lib = None

_SIGS = {
  "xotb200_last_error": (C.c_char_p, []),
  "xotb200_version": (C.c_char_p, []),
  "xotb200_create": (C.c_int, [C.POINTER(ModelCfg), C.c_int, C.POINTER(C.c_void_p)]),
  "xotb200_destroy": (C.c_int, [C.c_void_p]),
  "xotb200_set_shard": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
  "xotb200_load_shard": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
  "xotb200_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
  "xotb200_init_synthetic": (C.c_int, [C.c_void_p, C.c_uint64]),
  "xotb200_check_loaded": (C.c_int, [C.c_void_p]),
  "xotb200_begin_request": (C.c_int, [C.c_void_p, C.c_int]),
  "xotb200_begin_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
  "xotb200_select_request": (C.c_int, [C.c_void_p, C.c_int]),
  "xotb200_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
  "xotb200_forward_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
  "xotb200_forward_hop": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_void_p]),
  "xotb200_last_token": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
  "xotb200_inbox_export": (C.c_int, [C.c_void_p, C.c_void_p]),
  "xotb200_link_next_ipc": (C.c_int, [C.c_void_p, C.c_void_p]),
  "xotb200_link_next_local": (C.c_int, [C.c_void_p, C.c_void_p]),
  "xotb200_post_to_next": (C.c_int, [C.c_void_p, C.c_int]),
  "xotb200_ring_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
  "xotb200_ring_launch_slots": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
  "xotb200_post_token": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, C.c_int]),
  "xotb200_ring_poll": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]),
  "xotb200_clear_request": (C.c_int, [C.c_void_p, C.c_int]),
  "xotb200_ring_wait": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
  "xotb200_ring_next_step": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
  "xotb200_ring_timestamps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
  "xotb200_set_cta_limit": (C.c_int, [C.c_void_p, C.c_int]),
  "xotb200_set_prefill_mode": (C.c_int, [C.c_void_p, C.c_int]),
  "xotb200_debug_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
  "xotb200_debug_routing": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int]),
  "xotb200_debug_timeline": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]),
  "xotb200_kernel_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
  "xotb200_load_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
  "xotb200_weight_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
  "xotb200_step_bytes": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
}

EXPORTED_SYMBOLS = tuple(_SIGS.keys())


def load(path: os.PathLike | None = None) -> C.CDLL:
  """dlopen the library (once) and declare every prototype of include/xotb200.h.  Raises if it is missing."""
  global _lib
  if _lib is not None and path is None:
    return _lib
  p = Path(path) if path else Path(os.environ.get("XOTB200_LIB", LIB_PATH))   # XOTB200_LIB: A/B-test another build
  if not p.exists():
    raise ImportError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      f"or `make -C xotorch_b200/csrc` — xotorch_b200 has no CPU fallback")
  lib = C.CDLL(str(p))
  for name, (res, args) in _SIGS.items():
    fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
    fn.restype, fn.argtypes = res, args
  if path is None:
    _lib = lib
  return lib


def check(rc: int) -> int:
  if rc < 0:
    raise XotB200Error(rc, load().xotb200_last_error().decode("utf-8", "replace"))
  return rc
