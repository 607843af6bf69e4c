"""Multi-GPU (one process, one engine per GPU) — skipped unless >= 2 B200s are visible.  Covers the in-process NVLink path:
`xotb200_link_next_local` across devices (cudaDeviceEnablePeerAccess, peer stores + system-scope flags inside the persistent
kernel) through `LocalRing`.  The multi-PROCESS path (CUDA IPC) is what bench.py runs under torchrun."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from xotorch_b200.synth import PRESETS, hf_config, synth_prompt

NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


def _engines(shape_name, split, devices):
  from xotorch_b200.native import NativeShard
  shape = PRESETS[shape_name]
  out = []
  for (a, b), dev in zip(split, devices):
    e = NativeShard(hf_config(shape), device=dev)
    e.set_shard(a, b)
    e.init_synthetic(1234)
    out.append(e)
  return out


@pytest.mark.skipif(NGPU < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("shape_name", ["tiny-llama-hd64", "tiny-llama-hd128-f16"])
def test_local_ring_across_gpus_equals_single_gpu(shape_name):
  from xotorch_b200.ring import LocalRing
  shape = PRESETS[shape_name]
  nl = shape.n_layers
  prompt = synth_prompt(shape.vocab, 20)
  n_gen = 24
  one = LocalRing(_engines(shape_name, [(0, nl - 1)], [0]))
  ids1, _ = one.generate(prompt, n_gen)
  n = min(NGPU, nl)
  cuts = [round(i * nl / n) for i in range(n + 1)]
  split = [(cuts[i], cuts[i + 1] - 1) for i in range(n)]
  ring = LocalRing(_engines(shape_name, split, list(range(n))))
  ids_n, ms = ring.generate(prompt, n_gen)
  assert np.array_equal(ids1, ids_n), (ids1, ids_n)
  ids_again, _ = ring.generate(prompt, n_gen)            # monotonic ring step numbers across requests
  assert np.array_equal(ids1, ids_again)


@pytest.mark.skipif(NGPU < 2, reason="needs >= 2 GPUs")
def test_real_dims_two_gpus():
  import dataclasses
  from xotorch_b200.native import NativeShard
  from xotorch_b200.ring import LocalRing
  shape = dataclasses.replace(PRESETS["llama-3-8b"], n_layers=4, name="llama-3-8b-cut4")
  prompt = synth_prompt(shape.vocab, 40)
  def mk(split, devs):
    out = []
    for (a, b), d in zip(split, devs):
      e = NativeShard(hf_config(shape), device=d); e.set_shard(a, b); e.init_synthetic(1234); out.append(e)
    return out
  ids1, _ = LocalRing(mk([(0, 3)], [0])).generate(prompt, 16)
  ids2, ms = LocalRing(mk([(0, 1), (2, 3)], [0, 1])).generate(prompt, 16)
  assert np.array_equal(ids1, ids2)
