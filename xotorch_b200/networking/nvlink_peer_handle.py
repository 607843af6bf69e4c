"""`InProcessPeerHandle` — PeerHandle for peers that live in the SAME process on different GPUs of one box
(/root/reference/xotorch/networking/peer_handle.py:9-56; concrete signatures grpc_peer_handle.py:96-201).

The reference moves every hop through gRPC + gzip + a JSON-encoded inference state (≈1.25 ms bare, ≈100 ms with its
L x L mask).  When all shards sit on one 8xB200 box there are two better hops:
  * API-compatible hop (this class): `send_tensor` hands the ndarray straight to the peer Node's `process_tensor`
    — no serialisation, no sockets; used when the reference's Node drives the ring step by step.
  * device-resident hop (xotorch_b200/ring.py + libxotb200's inbox flags): the activation never leaves HBM; the last
    kernel of shard i stores it into shard i+1's inbox over NVLink.  That is the path bench.py's `value` measures.
Injection point upstream: the `create_peer_handle` factory given to the discovery module (main.py:146,155)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np


class InProcessPeerHandle:
  def __init__(self, _id: str, node, desc: str = "in-process NVLink peer", device_capabilities=None, address: str = "inproc"):
    self._id, self._node, self._desc, self._caps, self._addr = _id, node, desc, device_capabilities, address
    self._connected = False

  def id(self) -> str: return self._id
  def addr(self) -> str: return self._addr
  def description(self) -> str: return self._desc
  def device_capabilities(self): return self._caps

  async def connect(self) -> None: self._connected = True
  async def is_connected(self) -> bool: return self._connected
  async def disconnect(self) -> None: self._connected = False
  async def health_check(self) -> bool: return self._node is not None

  async def send_prompt(self, shard, prompt: str, inference_state: Optional[dict] = None, request_id: Optional[str] = None):
    return await self._node.process_prompt(shard, prompt, request_id, inference_state)

  async def send_tensor(self, shard, tensor: np.ndarray, inference_state: Optional[dict] = None, request_id: Optional[str] = None):
    # receiver side of grpc_server.py:73-88 without the wire: same call, same argument order
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
