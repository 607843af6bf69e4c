"""TEST INFRASTRUCTURE: ctypes wrapper around oracle/synth_fast.c (falls back to the torch generator if gcc output
is unavailable).  Produces the same tensors as xotorch_b200.synth.synth_state_dict, much faster."""
import ctypes as C
from typing import Dict, Iterable, Optional

import numpy as np
import torch

from xotorch_b200 import synth
from .build import build, OUT

_lib = None


def usable_cores() -> int:
  """Host cores this process may actually use: min(sched affinity, cgroup cpu quota).  The GPU box shows all 128 cores of the
  host but runs the container under a quota of 16: a thread team sized by the affinity mask makes the CPU oracle crawl."""
  import os
  from pathlib import Path
  n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
  for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    try:
      txt = Path(p).read_text().split()
      if p.endswith("cpu.max"):
        if txt[0] != "max":
          n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
      else:
        q = int(txt[0])
        if q > 0:
          per = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
          n = min(n, max(1, int(q / per + 0.5)))
      break
    except Exception:
      continue
  return max(1, n)


def _load():
  global _lib
  if _lib is None:
    lib = C.CDLL(str(build()))
    lib.synth_set_threads.argtypes = [C.c_int]
    lib.synth_set_threads.restype = None
    lib.synth_set_threads(usable_cores())
    lib.synth_fill_bf16.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, C.c_float]
    lib.synth_fill_bf16.restype = None
    lib.synth_fill_f16.argtypes = lib.synth_fill_bf16.argtypes
    lib.synth_fill_f16.restype = None
    _lib = lib
  return _lib


def synth_tensor_fast(seed: int, tensor_id: int, shape, std: float, mean: float, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
  n = int(np.prod(shape))
  out = torch.empty(n, dtype=dtype)
  scale = np.float32(std / synth.IH4_STD)
  fill = _load().synth_fill_f16 if dtype == torch.float16 else _load().synth_fill_bf16
  fill(C.c_void_p(out.data_ptr()), C.c_uint64(synth.tensor_base(seed, tensor_id)), 0, n,
                          C.c_float(float(scale)), C.c_float(mean))
  return out.view(*shape)


def synth_state_dict_fast(shape: synth.ModelShape, seed: int = 1234, layers: Optional[Iterable[int]] = None,
                          with_embed: bool = True, with_head: bool = True) -> Dict[str, torch.Tensor]:
  assert shape.dtype in ("bfloat16", "float16")
  dt = synth.torch_dtype(shape.dtype)
  return {name: synth_tensor_fast(seed, tid, shp, std, mean, dt)
          for name, tid, shp, std, mean in synth.hf_tensor_specs(shape, layers, with_embed, with_head)}
