"""N>1 host logic on CPU: world_size-2 and -3 `gloo` process groups drive `DistributedRing` with a stand-in for the CUDA
library (no compute is claimed; the device ring itself is covered on GPUs by tests/test_gpu_parity.py and bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xotorch_b200.ring import DistributedRing
from xotorch_b200.topology.partitioning import Partition, map_partitions_to_shards

D, V, NL = 16, 64, 6


class FakeShard:
  """Deterministic stand-in with the NativeShard surface DistributedRing uses."""

  def __init__(self, rank, start, end, mailbox=None, world=1):
    self.rank, self.start, self.end = rank, start, end
    self.linked, self.calls = None, []
    self.mailbox, self.world, self.hop_seq = mailbox, world, 0   # a directory stands in for the next engine's inbox

  @property
  def is_first(self): return self.start == 0
  @property
  def is_last(self): return self.end == NL - 1

  def inbox_export(self): return bytes([self.rank]) * 64
  def ring_next_step(self): return 0
  def link_next_ipc(self, h): self.linked = h[0]; self.hop_seq = 0

  def forward_hop(self, ids, seq_in, n, pos0):
    """xotb200_forward_hop stand-in: the hidden state goes through the mailbox directory, only hop NUMBERS cross the API"""
    import time
    self.calls.append(("forward_hop", None if ids is None else tuple(ids.shape), seq_in, n, pos0))
    if ids is not None:
      x = ids
    else:
      path = os.path.join(self.mailbox, f"to{self.rank}_{seq_in}.npy")
      for _ in range(2000):
        if os.path.exists(path):
          break
        time.sleep(0.005)
      x = np.load(path)
    out = self.forward(x, pos0)
    self.calls.pop()                                       # the nested forward() is not an API call
    if self.is_last:
      return None, out
    seq = self.hop_seq
    self.hop_seq += 1
    tmp = os.path.join(self.mailbox, f"tmp{self.rank}_{seq}.npy")
    np.save(tmp, out)
    os.replace(tmp, os.path.join(self.mailbox, f"to{(self.rank + 1) % self.world}_{seq}.npy"))
    return seq, None
  def link_next_local(self, other): self.linked = "self"
  def begin_request(self, n): self.calls.append(("begin", n))
  def begin_batch(self, n, b): self.calls.append(("begin_batch", n, b))
  def select_request(self, r): self.calls.append(("select", r))
  def last_token(self): return 1000 + [c[1] for c in self.calls if c[0] == "select"][-1]
  def post_to_next(self, step): self.calls.append(("post", step))
  def ring_launch(self, n_steps, step0, pos0): self.calls.append(("launch", n_steps, step0, pos0))
  def ring_wait(self, n_steps): return (np.arange(n_steps, dtype=np.int32) if self.is_last else None), 1.0

  def forward(self, x, pos0):
    self.calls.append(("forward", x.ndim, tuple(x.shape), pos0))
    if x.ndim == 2:
      h = np.zeros((1, x.shape[1], D), np.float32) + x[0][:, None].astype(np.float32)
    else:
      h = x.astype(np.float32)
    h = h + np.float32(self.end - self.start + 1)          # every layer adds 1
    if self.is_last:
      logits = np.zeros((1, V), np.float32)
      logits[0, int(h[0, -1, 0] + pos0) % V] = 1.0
      return logits
    return h


def _shards(world):
  parts, start = [], 0.0
  for i in range(world):
    end = round(start + 1.0 / world, 5)
    parts.append(Partition(str(i), start, end))
    start = end
  return map_partitions_to_shards(parts, NL, "m")


def _worker(rank, world, port, q, mailbox):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  sh = _shards(world)[rank]
  native = FakeShard(rank, sh.start_layer, sh.end_layer, mailbox, world)
  ring = DistributedRing(native, rank, world, D, V, comm_device="cpu")
  ring.link()
  prompt = np.arange(5, dtype=np.int32).reshape(1, -1)
  logits = ring.prefill(prompt, 5 + 8 + 1)
  tok = int(np.argmax(logits[0])) if native.is_last else 0
  tok = ring._bcast_token(tok)
  ids = ring.api_decode(tok, 5, 8)
  n_api = len(native.calls)
  ids_hop = ring.api_decode_device(tok, 5, 8)             # the same decode with the hidden state handed over by hop number
  hop_calls = native.calls[n_api:]
  del native.calls[n_api:]
  ring.step = 7                                            # a batch must start at a ring step that is a multiple of B
  n0 = len(native.calls)
  bids, _ = ring.generate_batch(np.arange(15, dtype=np.int32).reshape(3, 5), 4)
  q.put((rank, native.linked, tok, ids, native.calls[:n0], native.calls[n0:], None if bids is None else bids.tolist(), ring.step, ids_hop, hop_calls))
  dist.barrier()
  ring.close()
  dist.destroy_process_group()


def _free_port():
  s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_ring_host_logic(world, tmp_path):
  # single-process truth: one shard covering all layers
  one = FakeShard(0, 0, NL - 1)
  ring1 = DistributedRing(one, 0, 1, D, V)
  ring1.link()
  prompt = np.arange(5, dtype=np.int32).reshape(1, -1)
  tok1 = int(np.argmax(ring1.prefill(prompt, 14)[0]))
  ids1 = ring1.api_decode(tok1, 5, 8)
  assert one.linked == "self"

  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q, str(tmp_path))) for r in range(world)]
  for p in procs: p.start()
  res = sorted(q.get(timeout=120) for _ in range(world))
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for rank, linked, tok, ids, calls, bcalls, bids, step, ids_hop, hop_calls in res:
    assert linked == (rank + 1) % world                 # each rank mapped the inbox of the NEXT rank
    assert tok == tok1 and ids == ids1                  # sharded ring == unsharded, on every rank
    # device hop: one forward_hop per token per rank, ids only on the first shard, hop k consumed for token k elsewhere
    assert ids_hop == ids1 and [c[0] for c in hop_calls] == ["forward_hop"] * 8
    assert all((c[1] == (1, 1) and c[2] == -1) if rank == 0 else (c[1] is None and c[2] == k) for k, c in enumerate(hop_calls))
    assert [c[4] for c in hop_calls] == list(range(5, 13))
    fw = [c for c in calls if c[0] == "forward"]
    assert len(fw) == 9 and fw[0][3] == 0 and [c[3] for c in fw[1:]] == list(range(5, 13))
    assert fw[0][1] == (2 if rank == 0 else 3)          # ndim protocol: ids on the first shard, hidden elsewhere
    assert calls[0] == ("begin", 14)
    # requests in flight: B=3 requests, 4 tokens each -> ring steps 9..17 (7 rounded up to a multiple of 3), 9 steps
    assert bcalls[0] == ("begin_batch", 5 + 4 + 1, 3) and step == 9 + 9
    assert [c[1] for c in bcalls if c[0] == "select"] == [0, 1, 2]
    assert [c for c in bcalls if c[0] == "launch"] == [("launch", 9, 9, 5)]
    if rank == world - 1:
      assert [c[1] for c in bcalls if c[0] == "post"] == [9, 10, 11]      # request r seeds ring step step0 + r
      # ring step s carries request s % 3: ids[r] = [first token of r, steps r, r+3, r+6]
      assert bids == [[1000 + r, r, r + 3, r + 6] for r in range(3)]
    else:
      assert bids is None and not [c for c in bcalls if c[0] == "post"]
