"""Continuous batching on the device-resident ring (-m gpu): requests join and leave a running ring, each one still an
independent batch-1 greedy stream whose ids equal a stand-alone run; tokens are streamed to `on_token` one by one while the
kernels run; `NodeTokenBridge` delivers them on the asyncio loop thread like the reference's Node does (node.py:127-137)."""
import asyncio
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from xotorch_b200.synth import PRESETS, hf_config

NAME = "tiny-llama-hd64"


def _engines(split, limit=True):
  import torch
  from xotorch_b200.native import NativeShard
  shape = PRESETS[NAME]
  out = []
  for a, b in split:
    e = NativeShard(hf_config(shape), device=0)
    e.set_shard(a, b)
    e.init_synthetic(1234)
    out.append(e)
  if len(out) > 1 and limit:       # persistent kernels of several engines on ONE GPU must be co-resident
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for e in out:
      e.set_cta_limit(sms // len(out))
  return out


def _requests(n, seed=5):
  import torch
  shape = PRESETS[NAME]
  g = torch.Generator().manual_seed(seed)
  reqs = []
  for i in range(n):
    s = int(torch.randint(3, 40, (1,), generator=g))
    prompt = torch.randint(0, shape.vocab, (1, s), generator=g).to(torch.int32).numpy()
    reqs.append((f"req{i}", prompt, int(torch.randint(1, 45, (1,), generator=g))))
  return reqs


def _standalone(split, reqs):
  from xotorch_b200.ring import LocalRing
  ring = LocalRing(_engines(split))
  return {rid: ring.generate(p, n)[0].tolist() for rid, p, n in reqs}


@pytest.mark.parametrize("split,B,chunk", [([(0, 3)], 3, 4), ([(0, 1), (2, 3)], 2, 16), ([(0, 0), (1, 2), (3, 3)], 4, 7)])
def test_requests_join_and_leave_a_running_ring(split, B, chunk):
  from xotorch_b200.scheduler import ContinuousRing
  reqs = _requests(7)
  want = _standalone(split, reqs)
  events = []
  ring = ContinuousRing(_engines(split), max_requests=B, max_total_len=96, chunk=chunk,
                        on_token=lambda rid, toks, fin: events.append((rid, list(toks), fin)))
  for rid, p, n in reqs[:4]:
    ring.submit(rid, p, n)
  ring.step(); ring.step()                       # some requests are mid-flight ...
  for rid, p, n in reqs[4:]:
    ring.submit(rid, p, n)                       # ... when more arrive
  got = ring.run_until_idle()
  assert got == want
  # one on_token call per generated token, in order per request, exactly one `finished`, and it is the last
  for rid, _, n in reqs:
    mine = [(t, f) for r, t, f in events if r == rid]
    assert [t[0] for t, _ in mine] == want[rid] and all(len(t) == 1 for t, _ in mine)
    assert [f for _, f in mine] == [False] * (n - 1) + [True]
  assert ring.stats["prefills"] == 7 and ring.stats["tokens"] == sum(n for _, _, n in reqs)
  assert ring.stats["launches"] > ring.stats["reseeds"] >= 2          # unchanged active sets continue without reseeding


def test_eos_retires_a_request_early():
  from xotorch_b200.scheduler import ContinuousRing
  reqs = _requests(3, seed=11)
  want = _standalone([(0, 3)], [(rid, p, 30) for rid, p, _ in reqs])
  eos = want["req1"][4]                          # make the 5th token of req1 its EOS
  ring = ContinuousRing(_engines([(0, 3)]), max_requests=3, max_total_len=96, chunk=8)
  for rid, p, _ in reqs:
    ring.submit(rid, p, 30, eos_token_id=eos if rid == "req1" else None)
  got = ring.run_until_idle()
  cut = want["req1"].index(eos) + 1
  assert got["req1"] == want["req1"][:cut] and got["req0"] == want["req0"] and got["req2"] == want["req2"]


def test_node_token_bridge_delivers_on_the_loop_thread():
  from xotorch_b200.scheduler import ContinuousRing, NodeTokenBridge

  class Callbacks:                                 # the surface of the reference's AsyncCallbackSystem the bridge uses
    def __init__(self): self.calls = []
    def trigger_all(self, *a): self.calls.append((threading.get_ident(), a))

  class FakeNode:
    def __init__(self): self.on_token, self.buffered_token_output = Callbacks(), {}

  reqs = _requests(2, seed=3)
  want = _standalone([(0, 3)], reqs)

  async def main():
    loop = asyncio.get_running_loop()
    node = FakeNode()
    ring = ContinuousRing(_engines([(0, 3)]), max_requests=2, max_total_len=96, chunk=8, on_token=NodeTokenBridge(node, loop))
    for rid, p, n in reqs:
      ring.submit(rid, p, n)
    worker = threading.Thread(target=ring.run_until_idle)          # the ring runs on its own thread, like an engine executor
    worker.start()
    while worker.is_alive():
      await asyncio.sleep(0.005)
    await asyncio.sleep(0.05)
    return node, threading.get_ident()
  node, loop_thread = asyncio.run(main())
  assert {tid for tid, _ in node.on_token.calls} == {loop_thread}
  for rid, _, n in reqs:
    assert node.buffered_token_output[rid] == (want[rid], True)
    assert [a[1][0] for _, a in node.on_token.calls if a[0] == rid] == want[rid]
