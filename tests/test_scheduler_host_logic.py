"""Host logic of the continuous-batching scheduler (xotorch_b200/scheduler.py) on CPU, with a stand-in for the CUDA library
that emulates the ring protocol exactly as the kernels implement it: ring step s of a launch serves active request s % n at
position pos[i] + s // n, reads its input token from inbox slot (step0 + s) % (n + 1) if the flag there equals step + 1, and
publishes its output token for step + n.  No compute is claimed: the "model" is a hash of (token, position).  What is tested is
the bookkeeping — admission between launches, reseeding only when the active set changes, step numbers never reused, slot reuse,
EOS retirement, one on_token call per token in order."""
import numpy as np

from xotorch_b200.scheduler import ContinuousRing

V = 1000


def _next(tok: int, pos: int) -> int:
  return (tok * 7919 + pos * 104729 + 13) % V


class FakeRingShard:
  """one engine that is first AND last shard"""
  is_first = is_last = True

  def __init__(self):
    self.flags, self.tokens = {}, {}          # inbox slot -> flag value / token
    self.next_step, self.log, self.calls = 0, [], []
    self.kv_pos = {}                          # kv slot -> tokens in the cache (what a real engine's KV would hold)
    self.cur, self._last = 0, 0

  def link_next_local(self, other): assert other is self
  def begin_batch(self, L, B): self.L, self.B = L, B
  def select_request(self, r): self.cur = r
  def clear_request(self, r): self.kv_pos[r] = 0
  def ring_next_step(self): return self.next_step

  def forward(self, ids, pos0):
    assert pos0 == self.kv_pos[self.cur] == 0
    self.kv_pos[self.cur] = ids.shape[1]
    t = int(ids.sum()) % V
    for p in range(ids.shape[1]):
      t = _next(t, p)
    self._last = t
    return np.zeros((1, V), np.float32)

  def last_token(self): return self._last

  def post_token(self, step, token, n_active):
    assert step >= self.next_step, "a ring step number was reused"
    self.flags[step % (n_active + 1)], self.tokens[step % (n_active + 1)] = step + 1, token
    self.next_step = max(self.next_step, step + 1)
    self.calls.append(("post", step, n_active))

  def ring_launch_slots(self, n_steps, step0, slots, poss):
    n = len(slots)
    assert n_steps % n == 0 and step0 >= self.next_step - n, (step0, self.next_step)
    self.calls.append(("launch", n_steps, step0, tuple(slots), tuple(poss)))
    self.log = []
    for s in range(n_steps):
      step, i = step0 + s, s % n
      slot = step % (n + 1)
      assert self.flags.get(slot) == step + 1, f"ring step {step}: no token was published for it (flag {self.flags.get(slot)})"
      pos = poss[i] + s // n
      assert pos == self.kv_pos[slots[i]], "position table out of step with the KV cache"
      self.kv_pos[slots[i]] += 1
      out = _next(self.tokens[slot], pos)
      self.log.append(out)
      ns = step + n
      self.flags[ns % (n + 1)], self.tokens[ns % (n + 1)] = ns + 1, out
    self.next_step = max(self.next_step, step0 + n_steps)

  def ring_poll(self, start, max_n):
    return np.array(self.log[start:start + max_n], dtype=np.int32), len(self.log)

  def ring_wait(self, n): return None, 0.0


def _standalone(prompt, n):
  t = int(np.asarray(prompt).sum()) % V
  for p in range(len(prompt)):
    t = _next(t, p)
  out, pos = [t], len(prompt)
  while len(out) < n:
    out.append(_next(out[-1], pos))
    pos += 1
  return out


def test_admission_reseeding_and_step_numbers():
  rng = np.random.default_rng(3)
  reqs = [(f"r{i}", rng.integers(0, V, size=int(rng.integers(2, 30))).tolist(), int(rng.integers(1, 40))) for i in range(9)]
  eng = FakeRingShard()
  events = []
  ring = ContinuousRing([eng], max_requests=3, max_total_len=128, chunk=5, on_token=lambda rid, t, f: events.append((rid, t[0], f)))
  for rid, p, n in reqs[:4]:
    ring.submit(rid, p, n)
  ring.step(); ring.step(); ring.step()
  for rid, p, n in reqs[4:]:
    ring.submit(rid, p, n)
  got = ring.run_until_idle()
  assert got == {rid: _standalone(p, n) for rid, p, n in reqs}
  for rid, _, n in reqs:
    mine = [(t, f) for r, t, f in events if r == rid]
    assert [t for t, _ in mine] == got[rid] and [f for _, f in mine] == [False] * (n - 1) + [True]
  launches = [c for c in eng.calls if c[0] == "launch"]
  assert all(len(c[3]) <= 3 and len(set(c[3])) == len(c[3]) for c in launches)            # never more than the 3 KV slots, no slot twice
  # consecutive launches over the same active set continue on the published flags: no posts in between
  idx = [i for i, c in enumerate(eng.calls) if c[0] == "launch"]
  cont = [j for a, j in zip(idx, idx[1:]) if j == a + 1]
  assert cont and all(eng.calls[j][3] == eng.calls[j - 1][3] and eng.calls[j][2] == eng.calls[j - 1][2] + eng.calls[j - 1][1] for j in cont)
  assert ring.stats["reseeds"] < ring.stats["launches"] and ring.stats["prefills"] == 9
  assert not ring.active and sorted(ring.free) == [0, 1, 2]


def test_eos_and_capacity():
  eng = FakeRingShard()
  ring = ContinuousRing([eng], max_requests=2, max_total_len=64, chunk=8)
  want = _standalone([5, 6, 7], 30)
  ring.submit("a", [5, 6, 7], 30, eos_token_id=want[9])
  ring.submit("b", [1, 2], 12)
  got = ring.run_until_idle()
  assert got["a"] == want[:want.index(want[9]) + 1] and got["b"] == _standalone([1, 2], 12)
  import pytest
  with pytest.raises(ValueError, match="exceed the KV capacity"):
    ring.submit("c", list(range(60)), 10)
