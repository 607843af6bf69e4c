"""Continuous batching over the device-resident ring + the `Node.on_token` bridge.

The reference schedules every token from Python: `Node.process_inference_result` samples on the last shard, checks EOS /
max tokens, fires `on_token.trigger_all(request_id, [token], is_finished)` and forwards the id to the first shard
(/root/reference/xotorch/orchestration/node.py:109-147); its engine holds ONE request.  Here the per-token loop lives in the
persistent kernels (xotorch_b200/ring.py) and this module is what remains on the host:

* `ContinuousRing` — a request table over the B KV caches of `xotb200_begin_batch`.  Requests are admitted while others are
  running: the ring runs in short launches (`chunk` tokens per active request); between two launches a waiting request is
  prefilled into a free KV slot and joins the active set, finished ones (max tokens reached or EOS seen) leave, the rest keep
  their caches and positions (`xotb200_ring_launch_slots`: per-slot position table).  While the active set is unchanged the next
  launch continues on the flags the previous one left in the first shard's inbox — no reseeding, no host work per token.
* tokens are STREAMED: the last shard's kernel writes every id to pinned host memory as it is produced and the scheduler
  thread polls it (`xotb200_ring_poll`) while the kernels keep running, calling `on_token(request_id, [id], finished)` once
  per token like the reference does.
* `NodeTokenBridge` — marshals those calls onto the asyncio loop thread (`AsyncCallback.set` calls `asyncio.create_task`,
  helpers.py:119-123, so `node.on_token.trigger_all` must run there: `loop.call_soon_threadsafe`)."""
from __future__ import annotations

import threading
import time
from collections import deque
from dataclasses import dataclass, field
from typing import Callable, Deque, Dict, List, Optional, Sequence

import numpy as np

TokenCallback = Callable[[str, List[int], bool], None]


@dataclass
class _Req:
  request_id: str
  prompt: np.ndarray                 # int32 [1, S]
  max_tokens: int
  eos_token_id: Optional[int]
  slot: int = -1
  pos: int = 0                       # position of the next token to be fed (== tokens in the KV cache)
  last_token: int = -1               # token that feeds the request's next ring step
  n_out: int = 0
  done: bool = False
  ids: List[int] = field(default_factory=list)


class ContinuousRing:
  def __init__(self, engines: Sequence, max_requests: int, max_total_len: int, chunk: int = 16,
               on_token: Optional[TokenCallback] = None):
    """engines: NativeShard objects covering the model in ring order (one per GPU, or several on one GPU with
    set_cta_limit); max_requests <= 8 KV caches of max_total_len positions each."""
    assert engines and engines[0].is_first and engines[-1].is_last, "engines must cover the model in ring order"
    assert 1 <= max_requests <= 8 and chunk >= 1
    self.engines, self.B, self.L, self.chunk = list(engines), int(max_requests), int(max_total_len), int(chunk)
    self.on_token = on_token
    n = len(self.engines)
    for i, e in enumerate(self.engines):
      e.link_next_local(self.engines[(i + 1) % n])
      e.begin_batch(self.L, self.B)
    self.waiting: Deque[_Req] = deque()
    self.active: List[_Req] = []
    self.free = list(range(self.B))
    self.finished: Dict[str, _Req] = {}
    self._lock = threading.Lock()
    self._seeded_for: Optional[tuple] = None      # (active request ids, next step) the first shard's inbox is seeded for
    self.stats = {"launches": 0, "prefills": 0, "reseeds": 0, "tokens": 0}

  # ---- client side (any thread) --------------------------------------------------------------------------------------
  def submit(self, request_id: str, prompt_ids, max_tokens: int, eos_token_id: Optional[int] = None) -> None:
    p = np.asarray(prompt_ids, dtype=np.int32).reshape(1, -1)
    if p.shape[1] + max_tokens > self.L:
      raise ValueError(f"request {request_id!r}: {p.shape[1]} prompt + {max_tokens} generated tokens exceed the KV capacity {self.L}")
    with self._lock:
      self.waiting.append(_Req(request_id, p, int(max_tokens), eos_token_id))

  def pending(self) -> bool:
    with self._lock:
      return bool(self.waiting) or bool(self.active)

  # ---- scheduler (one thread) ------------------------------------------------------------------------------------------
  def _emit(self, r: _Req, tok: int) -> None:
    r.ids.append(tok)
    r.n_out += 1
    self.stats["tokens"] += 1
    fin = r.n_out >= r.max_tokens or (r.eos_token_id is not None and tok == r.eos_token_id)   # node.py:119-124
    r.done = fin
    if self.on_token is not None:
      self.on_token(r.request_id, [tok], fin)

  def _admit(self) -> bool:
    """prefill waiting requests into free KV slots (between two ring launches); True if the active set changed"""
    changed = False
    while self.free:
      with self._lock:
        if not self.waiting:
          break
        r = self.waiting.popleft()
      r.slot = self.free.pop(0)
      data = r.prompt
      for e in self.engines:
        e.select_request(r.slot)
        e.clear_request(r.slot)                       # reset_caches for this request only
        data = e.forward(data, 0)
      r.pos = r.prompt.shape[1]
      self.stats["prefills"] += 1
      self._emit(r, self.engines[-1].last_token())    # the prefill's token is the request's first generated token
      r.last_token = r.ids[-1]
      if r.done:
        self._retire(r)
      else:
        self.active.append(r)
        changed = True
    return changed

  def _retire(self, r: _Req) -> None:
    self.free.append(r.slot)
    self.finished[r.request_id] = r

  def step(self) -> bool:
    """one scheduling round: admit, launch one chunk for the active set, stream its tokens, retire.  False when idle."""
    changed = self._admit()
    if not self.active:
      return bool(self.waiting)
    act = self.active
    n = len(act)
    tokens_each = min(self.chunk, min(r.max_tokens - r.n_out for r in act))
    last = self.engines[-1]
    key = tuple(r.request_id for r in act)
    if changed or self._seeded_for is None or self._seeded_for[0] != key:
      # new active set: restart the step numbering past everything any engine has seen and seed every request's next token
      step0 = max(e.ring_next_step() for e in self.engines) + 9
      step0 = (step0 + n - 1) // n * n
      for i, r in enumerate(act):
        last.post_token(step0 + i, r.last_token, n)
      self.stats["reseeds"] += 1
    else:
      step0 = self._seeded_for[1]                     # the previous launch already published these steps' tokens
    n_steps = tokens_each * n
    slots = [r.slot for r in act]
    poss = [r.pos for r in act]
    for e in self.engines:
      e.ring_launch_slots(n_steps, step0, slots, poss)
    self.stats["launches"] += 1
    # stream the ids while the kernels run: ring step s of this launch belongs to active request s % n
    seen, stopped = 0, set()
    t_end = time.monotonic() + 60.0
    while seen < n_steps:
      toks, done = last.ring_poll(seen, n_steps - seen)
      if len(toks) == 0:
        if time.monotonic() > t_end:
          break                                       # ring_wait below reports what went wrong
        time.sleep(0.0002)
        continue
      for k, tok in enumerate(toks):
        r = act[(seen + k) % n]
        if r.request_id in stopped:
          continue                                    # generated past EOS inside this chunk: discarded
        self._emit(r, int(tok))
        r.last_token = int(tok)
        if r.done:
          stopped.add(r.request_id)
      seen += len(toks)
    for e in self.engines:
      e.ring_wait(0)
    for r in act:
      r.pos += tokens_each
    self._seeded_for = (key, step0 + n_steps)
    still = []
    for r in act:
      if r.done:
        self._retire(r)
      else:
        still.append(r)
    if len(still) != len(act):
      self._seeded_for = None
    self.active = still
    return True

  def run_until_idle(self) -> Dict[str, List[int]]:
    while self.step() or self.pending():
      pass
    return {k: v.ids for k, v in self.finished.items()}


class NodeTokenBridge:
  """on_token callback for ContinuousRing that delivers to a reference `Node` the way its own last-shard code does
  (node.py:127-137): `node.on_token.trigger_all(request_id, [token], is_finished)` ON THE EVENT-LOOP THREAD, and the
  accumulated ids in `node.buffered_token_output[request_id]`."""

  def __init__(self, node, loop):
    self.node, self.loop = node, loop

  def __call__(self, request_id: str, tokens: List[int], finished: bool) -> None:
    def deliver():
      buf = getattr(self.node, "buffered_token_output", None)
      if buf is not None:
        ids, _ = buf.get(request_id, ([], False))
        buf[request_id] = (ids + list(tokens), finished)
      self.node.on_token.trigger_all(request_id, list(tokens), finished)
    self.loop.call_soon_threadsafe(deliver)
