"""Same-box comparator for the hop number (SURVEY.md §8d last sentence): the REFERENCE's own hop —
`GRPCPeerHandle.send_tensor` (/root/reference/xotorch/networking/grpc/grpc_peer_handle.py:111-130) into the reference's own
`GRPCServer.SendTensor` (grpc_server.py:73-88) in a SECOND process on the same host (loopback), unmodified, from the package
installed under baseline/_ref (or /root/reference in the build container) — timed next to the device hop bench.py reports.

Measured per hop, payload = one decode hidden state [1, 1, dim] float32 (what the reference puts on the wire):
  * one-way : sender's clock just before send_tensor  ->  receiver's clock when its Node.process_tensor is entered
              (both processes read CLOCK_MONOTONIC of the same host).  This is the reference's "src shard done -> dst shard starts".
  * rtt     : the await of send_tensor itself (Node.forward_tensor awaits it, node.py:424-443).
Two inference_state variants: None ("bare": the floor of the transport) and the state the reference's torch engine really
ships every hop — tokens, input_pos and the full [1, L, L] boolean mask as nested lists in JSON (llm_utils.py:617-623,
grpc_peer_handle.py:203-224; L = 128 + 1024) — a handful of iterations, it is ~100 ms per hop.

Not part of the product; bench.py --impl reference --gpus N>1 calls measure_reference_hop().  No GPU involved."""
from __future__ import annotations

import asyncio
import json
import multiprocessing as mp
import os
import socket
import sys
import tempfile
import time
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _reference_root():
  for p in (ROOT / "baseline" / "_ref", Path("/root/reference")):
    if (p / "xotorch" / "networking" / "grpc" / "grpc_server.py").exists():
      return p
  return None


def _import_reference():
  root = _reference_root()
  if root is None:
    raise RuntimeError("reference package not found (baseline/_ref or /root/reference)")
  stub, stub_all = types.ModuleType("scapy"), types.ModuleType("scapy.all")
  stub_all.get_if_addr = lambda *a, **k: "127.0.0.1"
  stub_all.get_if_list = lambda *a, **k: []
  sys.modules.setdefault("scapy", stub)
  sys.modules.setdefault("scapy.all", stub_all)
  sys.dont_write_bytecode = True
  os.chdir(tempfile.mkdtemp(prefix="refhop_"))            # device_capabilities.py writes a log file into the CWD
  sys.path.insert(0, str(root))


class _RecvNode:
  """What GRPCServer needs from a Node on this path: process_tensor.  Records the arrival time of every hop."""

  def __init__(self):
    self.t_recv = {}

  async def process_tensor(self, shard, tensor, request_id=None, inference_state=None):
    self.t_recv[request_id] = time.monotonic_ns()
    return None


def _server_proc(port, ready, done, out_q):
  _import_reference()
  from xotorch.networking.grpc.grpc_server import GRPCServer

  async def main():
    node = _RecvNode()
    srv = GRPCServer(node, "127.0.0.1", port)
    await srv.start()
    ready.set()
    while not done.is_set():
      await asyncio.sleep(0.02)
    out_q.put(node.t_recv)
    await srv.stop()
  asyncio.run(main())


def _free_port():
  s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def measure_reference_hop(dim: int = 4096, n_bare: int = 300, n_state: int = 6, ctx_len: int = 128 + 1024) -> dict:
  ctx = mp.get_context("spawn")
  port = _free_port()
  ready, done, out_q = ctx.Event(), ctx.Event(), ctx.Queue()
  p = ctx.Process(target=_server_proc, args=(port, ready, done, out_q), daemon=True)
  p.start()
  try:
    if not ready.wait(60):
      raise RuntimeError("reference GRPCServer did not start")
    _import_reference()
    from xotorch.networking.grpc.grpc_peer_handle import GRPCPeerHandle
    from xotorch.inference.shard import Shard
    from xotorch.topology.device_capabilities import DeviceCapabilities, DeviceFlops
    caps = DeviceCapabilities(model="B200", chip="B200", memory=183359, flops=DeviceFlops(fp32=0, fp16=0, int8=0))
    shard = Shard("llama-3-8b", 16, 31, 32)
    x = np.random.default_rng(0).standard_normal((1, 1, dim)).astype(np.float32)
    # the state the reference's torch engine attaches to EVERY hop (ShardInferenceState.to_dict, llm_utils.py:617-623)
    big = {"tokens": [list(range(129))], "input_pos": [list(range(ctx_len))],
           "mask": np.tril(np.ones((ctx_len, ctx_len), dtype=bool))[None].tolist(), "curr_pos": 129}
    sent, rtt = {}, {}

    async def main():
      h = GRPCPeerHandle("peer", f"127.0.0.1:{port}", "ref hop bench", caps)
      await h.connect()
      for i in range(20):                                   # warm the channel
        await h.send_tensor(shard, x, inference_state=None, request_id=f"w{i}")
      for i in range(n_bare):
        t0 = time.monotonic_ns()
        await h.send_tensor(shard, x, inference_state=None, request_id=f"b{i}")
        sent[f"b{i}"], rtt[f"b{i}"] = t0, time.monotonic_ns() - t0
      for i in range(n_state):
        t0 = time.monotonic_ns()
        await h.send_tensor(shard, x, inference_state=big, request_id=f"s{i}")
        sent[f"s{i}"], rtt[f"s{i}"] = t0, time.monotonic_ns() - t0
      await h.disconnect()
    asyncio.run(main())
    done.set()
    t_recv = out_q.get(timeout=30)
  finally:
    done.set()
    p.join(timeout=20)
    if p.is_alive():
      p.kill()                                              # the exact process started above

  def stats(prefix):
    ow = np.array([t_recv[k] - sent[k] for k in sent if k.startswith(prefix) and k in t_recv], dtype=np.float64) / 1e3
    rt = np.array([rtt[k] for k in rtt if k.startswith(prefix)], dtype=np.float64) / 1e3
    return {"n": int(len(ow)), "one_way_p50_us": round(float(np.median(ow)), 1), "one_way_p90_us": round(float(np.percentile(ow, 90)), 1),
            "rtt_p50_us": round(float(np.median(rt)), 1)}
  return {"what": "reference GRPCPeerHandle.send_tensor -> reference GRPCServer.SendTensor, 2 processes, loopback, same host (unmodified reference code)",
          "payload_bytes": int(x.nbytes), "bare": stats("b"), "with_reference_inference_state": {**stats("s"), "state_ctx_len": ctx_len},
          "p50_us": stats("b")["one_way_p50_us"], "cores": len(os.sched_getaffinity(0))}


if __name__ == "__main__":
  print(json.dumps(measure_reference_hop(int(sys.argv[1]) if len(sys.argv) > 1 else 4096)))
