"""`NVLinkPeerHandle` — the PeerHandle for a peer whose engine lives in THIS process on a GPU of the same NVSwitch box
(/root/reference/xotorch/networking/peer_handle.py:9-56; concrete signatures grpc_peer_handle.py:96-201; receiver side
grpc_server.py:73-88).  Injection point upstream: the `create_peer_handle` factory given to the discovery module
(main.py:146,155).

The reference moves every hop through gRPC + gzip + a JSON-encoded inference state (≈1.25 ms bare, ≈100 ms with its L x L
mask).  Here `connect()` attaches the peer's engine to the local one (`attach_next_engine`): from then on a
non-last shard's `infer_tensor` runs `xotb200_forward_hop` — the last epilogue of its decode kernel stores the 16-bit hidden
state straight into the peer engine's HBM over NVLink and releases a flag (a prompt is one peer copy) — and returns a
4-word DESCRIPTOR instead of the hidden state.  `send_tensor` carries that descriptor to the peer `Node.process_tensor`
(same call, same argument order as the gRPC receiver); the peer's engine recognises it and consumes the payload where it
already is.  No serialisation, no socket, no host copy of the activation, no stream synchronisation per hop.  Anything that
is not a descriptor (token ids closing the ring, host hidden states from an engine that is not attached) passes through
unchanged, so the handle also works between engines that cannot hop on the device.

The fully device-resident ring (xotorch_b200/ring.py: one persistent kernel per shard for all tokens) remains the fastest
path and is what bench.py's `value` measures; this handle is the same hop behind the reference's own per-token API."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from xotorch_b200.inference.b200_engine import parse_hop_descriptor


class NVLinkPeerHandle:
  def __init__(self, _id: str, node, local_engine=None, desc: str = "same-box NVLink peer", device_capabilities=None,
               address: str = "nvlink"):
    """`node`: the peer's Node (same process).  `local_engine`: the engine of the node that OWNS this handle (the sender);
    None keeps every hop on host arrays."""
    self._id, self._node, self._desc, self._caps, self._addr = _id, node, desc, device_capabilities, address
    self._local_engine = local_engine
    self._connected = False
    self.sent = {"descriptors": 0, "host_tensors": 0, "payload_bytes_on_wire": 0}

  def id(self) -> str: return self._id
  def addr(self) -> str: return self._addr
  def description(self) -> str: return self._desc
  def device_capabilities(self): return self._caps

  def _peer_engine(self):
    return getattr(self._node, "inference_engine", None)

  async def connect(self) -> None:
    peer = self._peer_engine()
    if self._local_engine is not None and hasattr(self._local_engine, "attach_next_engine") and hasattr(peer, "attach_next_engine"):
      self._local_engine.attach_next_engine(peer)           # hidden states for this peer now stay in HBM
    self._connected = True

  async def is_connected(self) -> bool: return self._connected

  async def disconnect(self) -> None:
    if self._local_engine is not None and hasattr(self._local_engine, "attach_next_engine"):
      self._local_engine.attach_next_engine(None)
    self._connected = False

  async def health_check(self) -> bool: return self._node is not None

  async def send_prompt(self, shard, prompt: str, inference_state: Optional[dict] = None, request_id: Optional[str] = None):
    return await self._node.process_prompt(shard, prompt, request_id, inference_state)

  async def send_tensor(self, shard, tensor: np.ndarray, inference_state: Optional[dict] = None, request_id: Optional[str] = None):
    # receiver side of grpc_server.py:73-88 without the wire: same call, same argument order
    hop = parse_hop_descriptor(tensor)
    if hop is not None:
      peer = self._peer_engine()
      if self._local_engine is None or getattr(self._local_engine, "_hop_next", None) is not peer:
        raise RuntimeError(f"hop descriptor {hop} was produced for another peer than {self._id}")
      self.sent["descriptors"] += 1
    else:
      self.sent["host_tensors"] += 1
    self.sent["payload_bytes_on_wire"] += int(np.asarray(tensor).nbytes)
    return await self._node.process_tensor(shard, tensor, request_id, inference_state)

  async def send_result(self, request_id: str, result: List[int], is_finished: bool) -> None:
    cb = getattr(self._node, "on_token", None)
    if cb is not None:
      cb.trigger_all(request_id, result, is_finished)

  async def send_opaque_status(self, request_id: str, status: str) -> None:
    cb = getattr(self._node, "on_opaque_status", None)
    if cb is not None:
      cb.trigger_all(request_id, status)

  async def collect_topology(self, visited: set, max_depth: int):
    return await self._node.collect_topology(visited, max_depth)


class InProcessPeerHandle(NVLinkPeerHandle):
  """The same handle with no local engine attached: every hop stays on host arrays (what round 1 shipped; kept for callers
  whose engines are not B200ShardInferenceEngine)."""

  def __init__(self, _id: str, node, desc: str = "in-process peer", device_capabilities=None, address: str = "inproc"):
    super().__init__(_id, node, None, desc, device_capabilities, address)
