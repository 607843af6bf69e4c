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
    self._sync_step()

  def _sync_step(self):
    """Ring step numbers are never reused on an engine (its inbox flags hold step+1 and are not cleared): a new ring over
    existing engines, or a ring recovering from a timeout, continues after the highest number any of them has seen."""
    self._step = max(e.ring_next_step() for e in self.engines)

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
    self._sync_step()
    step0 = self._step
    self.engines[-1].post_to_next(step0)
    toks, ms = None, 0.0
    try:
      for e in self.engines:
        e.ring_launch(n_gen - 1, step0, s)
      for e in self.engines:
        t, m = e.ring_wait(n_gen - 1)
        toks = t if t is not None else toks
        ms = max(ms, m)
    finally:
      self._sync_step()                     # also after XOTB200_ETIMEOUT: the numbers of the failed run stay burnt
    return np.concatenate([[first], toks]).astype(np.int32), ms

  def generate_batch(self, prompts: np.ndarray, n_gen: int) -> Tuple[np.ndarray, float]:
    """B independent requests (prompts [B, S], same length) kept in flight through the ring together: ring step s serves
    request s % B, so with B >= number of shards no GPU idles.  Returns (ids [B, n_gen], decode ms)."""
    prompts = np.asarray(prompts, dtype=np.int32)
    B, S = prompts.shape
    self._sync_step()
    self._step = (self._step + B - 1) // B * B                 # ring step numbers of a batch start at a multiple of B
    for e in self.engines:
      e.begin_batch(S + n_gen + 1, B)
    first = []
    for r in range(B):
      data = prompts[r:r + 1]
      for e in self.engines:
        e.select_request(r)
        data = e.forward(data, 0)
      first.append(self.engines[-1].last_token())
      if n_gen > 1:
        self.engines[-1].post_to_next(self._step + r)          # seeds request r's first ring step
    if n_gen == 1:
      return np.array(first, dtype=np.int32).reshape(B, 1), 0.0
    n_steps = B * (n_gen - 1)
    toks, ms = None, 0.0
    try:
      for e in self.engines:
        e.ring_launch(n_steps, self._step, S)
      for e in self.engines:
        t, m = e.ring_wait(n_steps)
        toks = t if t is not None else toks
        ms = max(ms, m)
    finally:
      self._sync_step()
    ids = np.concatenate([np.array(first, dtype=np.int32).reshape(B, 1), toks.reshape(n_gen - 1, B).T], axis=1)
    return ids, ms


class DistributedRing:
  """One process per shard (torchrun / torch.distributed): rank r owns shard r of the ring and GPU LOCAL_RANK.

  `torch.distributed` is plumbing only: it carries the CUDA-IPC inbox handles at link time, the fp32 hidden state of
  the HOST-API path (prefill and the per-token e2e path; bytes = 4*S*D per hop) and the 4-byte token id that closes the
  ring on that path.  The device-resident decode (`ring_decode`) uses no collective at all: activations and tokens move
  between the engines' inboxes by peer stores + flags inside the persistent kernel.

  Mirrors what the reference's Node does per request across peers (node.py:193-208 prefill on the first shard, :361-380
  process_tensor on the others, :424-443 forward to partition (i+1) % n, :109-147 sample on the last shard)."""

  def __init__(self, native, rank: int, world: int, dim: int, vocab: int, comm_device="cpu"):
    import torch.distributed as dist
    self.dist, self.native, self.rank, self.world = dist, native, rank, world
    self.dim, self.vocab = dim, vocab
    self.prev, self.next = (rank - 1) % world, (rank + 1) % world
    self.comm_device = comm_device
    self.first, self.last = native.is_first, native.is_last
    self.step = 0                      # monotonic ring step counter, identical on every rank
    self.hops = 0                      # API hops made on every link since link() (identical on every rank)
    self._shm = None                   # same-box mailbox for the token id that closes the ring on the per-token API path
    self._tok_seq = 0

  # -- link: every engine learns the inbox of the next one -----------------------------------------------------------
  def link(self):
    if self.world == 1:
      self.native.link_next_local(self.native)
      self.step = self.native.ring_next_step()
      return
    handles = [None] * self.world
    self.dist.all_gather_object(handles, (self.native.inbox_export(), self.native.ring_next_step()))
    self.native.link_next_ipc(handles[self.next][0])      # also restarts the link's API-hop numbering at 0
    self.hops = 0
    self._open_token_mailbox()
    self.step = max(h[1] for h in handles)      # never reuse a step number any engine of the ring has seen (flags are monotonic)
    self.dist.barrier()

  # -- ring closure on the per-token API path: 4 bytes from the last shard's host to everybody's host -----------------
  def _open_token_mailbox(self):
    """All ranks sit on one box: the token id goes through 16 bytes of POSIX shared memory (seq, token) instead of a
    collective — a NCCL broadcast costs a kernel launch and a stream synchronise per token on every rank (~40 us of a 2.9 ms
    token), the mailbox ~1 us.  Falls back to the broadcast if shared memory is unavailable."""
    from multiprocessing import shared_memory
    try:
      name = [None]
      if self.rank == 0:
        self._shm = shared_memory.SharedMemory(create=True, size=64)
        np.ndarray((2,), dtype=np.int64, buffer=self._shm.buf)[:] = 0
        name[0] = self._shm.name
      self.dist.broadcast_object_list(name, src=0)
      if self.rank != 0:
        self._shm = shared_memory.SharedMemory(name=name[0])
      self._tok_seq = 0
      self.dist.barrier()
    except Exception:
      self._shm = None

  def close(self):
    if self._shm is not None:
      try:
        self._shm.close()
        if self.rank == 0:
          self._shm.unlink()
      except Exception:
        pass
      self._shm = None

  def _pass_token(self, tok: int) -> int:
    """the last rank publishes `tok`, every rank returns it"""
    if self._shm is None:
      return self._bcast_token(tok)
    box = np.ndarray((2,), dtype=np.int64, buffer=self._shm.buf)
    self._tok_seq += 1
    if self.last:
      box[1] = tok
      box[0] = self._tok_seq            # x86 keeps the two stores in order
      return tok
    import time
    t_end = time.monotonic() + 30.0
    while int(box[0]) < self._tok_seq:
      if time.monotonic() > t_end:
        raise TimeoutError("no token from the last shard within 30 s")
    return int(box[1])

  # -- host-array hop (API-compatible path) -----------------------------------------------------------------------------
  def _send(self, arr: np.ndarray):
    import torch
    self.dist.send(torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(self.comm_device), self.next)

  def _recv(self, shape) -> np.ndarray:
    import torch
    t = torch.empty(shape, dtype=torch.float32, device=self.comm_device)
    self.dist.recv(t, self.prev)
    return t.cpu().numpy()

  def _bcast_token(self, tok: int) -> int:
    import torch
    t = torch.tensor([tok], dtype=torch.int64, device=self.comm_device)
    self.dist.broadcast(t, self.world - 1)
    return int(t.item())

  def forward_through_ring(self, ids_or_none: Optional[np.ndarray], n_tokens: int, pos0: int) -> Optional[np.ndarray]:
    """One pass of `n_tokens` positions through all shards with host buffers.  Rank 0 passes the ids; the last rank
    returns the logits [1, V]; every other rank returns None."""
    if self.world == 1:
      return self.native.forward(ids_or_none, pos0)
    if self.first:
      self._send(self.native.forward(ids_or_none, pos0))
      return None
    h = self._recv((1, n_tokens, self.dim))
    out = self.native.forward(h, pos0)
    if not self.last:
      self._send(out)
      return None
    return out

  def prefill(self, prompt_ids: np.ndarray, max_total_len: int) -> Optional[np.ndarray]:
    self.native.begin_request(max_total_len)
    ids = np.asarray(prompt_ids, dtype=np.int32).reshape(1, -1)
    return self.forward_through_ring(ids if self.first else None, ids.shape[1], 0)

  def prefill_device(self, prompt_ids: np.ndarray, max_total_len: int, begin: bool = True) -> Optional[np.ndarray]:
    """Prefill with a GPU-to-GPU hop: the 16-bit hidden state [S, D] goes from `xotb200_forward_device` of one rank to the
    next rank's by NCCL send/recv over NVLink (2*S*D bytes, the algorithmic size of the hop) — no host staging.
    Rank 0 passes the ids; the last rank returns the logits [1, V]; every other rank returns None."""
    import torch
    if begin:
      self.native.begin_request(max_total_len)
    ids = np.asarray(prompt_ids, dtype=np.int32).reshape(1, -1)
    S = ids.shape[1]
    if self.world == 1:
      return self.native.forward(ids, 0)
    dt = torch.float16 if self.native.cfg.dtype == 1 else torch.bfloat16
    h = out = None
    if not self.first:
      h = torch.empty((S, self.dim), dtype=dt, device=self.comm_device)
      self.dist.recv(h, self.prev)
      torch.cuda.current_stream().synchronize()            # the engine runs on its own stream
    if not self.last:
      out = torch.empty((S, self.dim), dtype=dt, device=self.comm_device)
    logits = self.native.forward_device(ids if self.first else None, h.data_ptr() if h is not None else None, S, 0,
                                        out.data_ptr() if out is not None else None)
    if out is not None:
      self.dist.send(out, self.next)
    return logits

  def api_decode(self, first_token: int, pos0: int, n_gen: int) -> List[int]:
    """Greedy decode through the host API, token by token (what Node drives): returns the n_gen ids on every rank."""
    tok, pos, out = first_token, pos0, []
    for _ in range(n_gen):
      ids = np.array([[tok]], dtype=np.int32)
      logits = self.forward_through_ring(ids if self.first else None, 1, pos)
      if self.last:
        tok = int(np.argmax(logits[0]))
      if self.world > 1:
        tok = self._bcast_token(tok)            # ring closure: the id travels back to the first shard
      out.append(tok)
      pos += 1
    return out

  def api_decode_device(self, first_token: int, pos0: int, n_gen: int) -> List[int]:
    """`api_decode` with the hop kept on the GPUs — what NVLinkPeerHandle does behind PeerHandle.send_tensor: every rank still
    makes ONE engine call per token (`xotb200_forward_hop`: token ids from host memory on the first shard, fp32 logits to host
    memory on the last — the user-visible API), but the hidden state is stored into the next rank's inbox by the producing
    kernel (peer store over NVLink + release flag through the CUDA-IPC mapping) and acquired there by the consuming kernel:
    no NCCL send/recv, no stream synchronisation per hop.  Non-last ranks return from the call at once, so a rank's kernel for
    token k is usually already waiting on its flag when the hidden state lands.  The 4-byte token id that closes the ring goes
    host to host through a shared-memory mailbox (the reference sends it to the first shard as a tensor, node.py:119-147)."""
    if self.world == 1:
      return self.api_decode(first_token, pos0, n_gen)
    tok, pos, ids_out = first_token, pos0, []
    for _ in range(n_gen):
      # every link of the ring carries exactly one hop per token, so the number of the hop to consume is a local count
      seq_out, logits = self.native.forward_hop(np.array([[tok]], dtype=np.int32) if self.first else None,
                                                -1 if self.first else self.hops, 1, pos)
      assert self.last or seq_out == self.hops, f"hop numbering out of step: produced {seq_out}, expected {self.hops}"
      self.hops += 1
      if self.last:
        tok = int(np.argmax(logits[0]))
      tok = self._pass_token(tok)               # ring closure: the id travels back to the first shard (same-box mailbox)
      ids_out.append(tok)
      pos += 1
    return ids_out

  # -- device-resident decode -------------------------------------------------------------------------------------------
  def ring_decode(self, n_steps: int, pos0: int) -> Tuple[Optional[np.ndarray], float]:
    """After a prefill: the last shard seeds shard 0's inbox with the first generated token, then every rank launches
    ONE persistent kernel for n_steps tokens.  Returns (ids on the last rank | None, device ms of this rank)."""
    s0 = self.step
    if self.last:
      self.native.post_to_next(s0)
    if self.world > 1:
      self.dist.barrier()
    try:
      self.native.ring_launch(n_steps, s0, pos0)
      toks, ms = self.native.ring_wait(n_steps)
    finally:
      self.step = s0 + n_steps                # also on XOTB200_ETIMEOUT: the numbers of a failed run stay burnt
    return toks, ms

  # -- B requests in flight: ring step s serves request s % B; with B >= world no GPU idles ------------
  def generate_batch(self, prompts: np.ndarray, n_gen: int) -> Tuple[Optional[np.ndarray], float]:
    """prompts [B, S] (same on every rank).  Prefill per request through the host-array path, then ONE persistent kernel
    per rank for B*(n_gen-1) ring steps.  Returns (ids [B, n_gen] on the last rank | None, device ms of this rank)."""
    prompts = np.asarray(prompts, dtype=np.int32)
    B, S = prompts.shape
    self.step = (self.step + B - 1) // B * B
    self.native.begin_batch(S + n_gen + 1, B)
    if self.world > 1:
      self.dist.barrier()                                  # every inbox is reset before anybody posts into one
    first = []
    for r in range(B):
      self.native.select_request(r)
      self.forward_through_ring(prompts[r:r + 1] if self.first else None, S, 0)
      if self.last:
        first.append(self.native.last_token())
        if n_gen > 1:
          self.native.post_to_next(self.step + r)
    n_steps = B * (n_gen - 1)
    toks, ms = None, 0.0
    if n_steps:
      if self.world > 1:
        self.dist.barrier()
      s0 = self.step
      try:
        self.native.ring_launch(n_steps, s0, S)
        toks, ms = self.native.ring_wait(n_steps)
      finally:
        self.step = s0 + n_steps
    if not self.last:
      return None, ms
    ids = np.array(first, dtype=np.int32).reshape(B, 1)
    if toks is not None:
      ids = np.concatenate([ids, toks.reshape(n_gen - 1, B).T], axis=1)
    return ids, ms

