"""Device-resident ring over the shards of one model on the GPUs of one box (single process, one engine per GPU).

Replaces the per-token Node.forward_tensor -> PeerHandle.send_tensor hop
(/root/reference/xotorch/orchestration/node.py:424-443) when every shard sits on the same NVSwitch box: after a
prefill through the host API, `generate` launches ONE persistent kernel per shard; activations and the argmax token
travel between the engines' inboxes over NVLink (peer stores + release/acquire flags), the host only collects the ids.
bench.py does the same across processes (torchrun) with CUDA-IPC mapped inboxes."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from xotorch_b200.native import NativeShard


class LocalRing:
  def __init__(self, engines: Sequence[NativeShard]):
    assert engines and engines[0].is_first and engines[-1].is_last, "engines must cover the model in ring order"
    self.engines = list(engines)
    n = len(self.engines)
    same_dev = len({e.device for e in self.engines}) < n
    if same_dev and n > 1:
      import torch
      sms = torch.cuda.get_device_properties(self.engines[0].device).multi_processor_count
      for e in self.engines:
        e.set_cta_limit(sms // n)           # persistent kernels sharing one GPU must be co-resident
    for i, e in enumerate(self.engines):
      e.link_next_local(self.engines[(i + 1) % n])
    self._step = 0

  def prefill(self, prompt_ids: np.ndarray, max_total_len: int) -> np.ndarray:
    """host-buffer API through all shards; returns the last-position logits [1, V]."""
    data = np.asarray(prompt_ids, dtype=np.int32).reshape(1, -1)
    for e in self.engines:
      e.begin_request(max_total_len)
    for e in self.engines:
      data = e.forward(data, 0)
    return data

  def generate(self, prompt_ids: np.ndarray, n_gen: int) -> Tuple[np.ndarray, float]:
    """greedy ids [n_gen] (on-device lowest-index argmax) and the decode time in ms (max over shards, CUDA events)."""
    s = int(np.asarray(prompt_ids).size)
    self.prefill(prompt_ids, s + n_gen + 1)
    first = self.engines[-1].last_token()
    if n_gen == 1:
      return np.array([first], dtype=np.int32), 0.0
    self.engines[-1].post_to_next(self._step)
    for e in self.engines:
      e.ring_launch(n_gen - 1, self._step, s)
    toks, ms = None, 0.0
    for e in self.engines:
      t, m = e.ring_wait(n_gen - 1)
      toks = t if t is not None else toks
      ms = max(ms, m)
    self._step += n_gen - 1
    return np.concatenate([[first], toks]).astype(np.int32), ms
