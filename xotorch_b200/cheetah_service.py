"""Cheetah-socket façade: a Unix-socket service that speaks the wire protocol of the reference's
`CheetahInferenceEngine.run_cheetah` (/root/reference/xotorch/inference/cheetah/sharded_inference_engine.py:331-457) and
answers from libxotb200 — so upstream's Cheetah client can drive the B200 engine without any change on its side.  The
reference ships only the CLIENT of that protocol (the C++ "Cheetah" service it talks to is not in the repository), so the
server side here is written from the client's send/receive code:

  request  (one per connection, `:343-395`)
      4 bytes  big-endian length of the JSON header — `struct.pack("!I", n)`; for a floating-point hidden state the
               reference packs the length with "!f" (`:388`, a float32 holding the integer) — both are accepted
      header   JSON: node_id, model, model_path, layer_start, layer_end, layer_total, session_id, has_hidden_state and
               either  input_ids_shape / mask_shape / input_pos_shape + dtype_input_ids / dtype_mask / dtype_input_pos
               or      hidden_state_shape / hidden_state_dtype                                  (`:321-329,357-371`)
      payload  tokens bytes + mask as uint8 + input_pos bytes   |   hidden state as float32     (`:377-395`)
  response (`:401-441`)
      4 bytes "!I" header length, JSON {"shape": [...], "dtype": "float32" | "bfloat16"}, raw tensor
      (bfloat16 travels as uint16 and is widened by the client; the client takes `out[:, -1]` of a token request, `:454`).

What the service keeps per (model_path, layer range): one engine; per session_id: the position — a hidden-state request
carries no position at all, so the service counts the tokens each session has pushed through this shard (a token request
states it in `input_pos`).  The mask is read and dropped: attention here touches only the valid prefix.
Last shard: the reply is the last position's logits as [1, 1, V] (the client's `[:, -1]` yields the [1, V] it passes to
`sample`); other shards: the hidden state [1, S, D].

Parity of this façade is pinned only against the client's own code path restated in tests/test_gpu_cheetah.py: the
reference client cannot be imported here (it needs torchtune), and no Cheetah server exists to compare with."""
from __future__ import annotations

import asyncio
import json
import os
import struct
from pathlib import Path
from typing import Callable, Dict, Optional, Tuple

import numpy as np

DEFAULT_SOCKET = "/tmp/cheetah_infra"            # sharded_inference_engine.py:346
_NP = {"int32": np.int32, "int64": np.int64, "int16": np.int16, "uint8": np.uint8, "bool": np.bool_, "float32": np.float32,
       "float16": np.float16}


def _header_len(raw: bytes) -> int:
  """4 bytes: "!I", or the reference's "!f" slip for floating-point hidden states (a float32 whose value is the length)"""
  n = struct.unpack("!I", raw)[0]
  if n > (1 << 24):                                # no JSON header is 16 MB long: it was packed as a float
    f = struct.unpack("!f", raw)[0]
    if not (0 < f < (1 << 24)) or f != int(f):
      raise ValueError(f"bad header length field {raw!r}")
    n = int(f)
  return n


class CheetahService:
  def __init__(self, socket_path: str = DEFAULT_SOCKET, device: int = 0, native_factory: Optional[Callable] = None,
               reply_dtype: str = "float32"):
    """native_factory(hf_config, device) -> NativeShard-like (tests inject a stand-in); reply_dtype "float32" or "bfloat16"
    (the client accepts both, `:426-434`; bfloat16 halves the bytes and is exact for a bf16 model)."""
    assert reply_dtype in ("float32", "bfloat16")
    self.socket_path, self.device, self.reply_dtype = socket_path, device, reply_dtype
    self._factory = native_factory
    self.engines: Dict[Tuple[str, int, int], object] = {}
    self.sessions: Dict[Tuple[str, Tuple[str, int, int]], int] = {}      # (session, shard key) -> tokens seen = next position
    self.server: Optional[asyncio.AbstractServer] = None
    self._lock = asyncio.Lock()                     # one request at a time per service, like the reference's 1-thread executor
    self.requests = 0

  # ---- engines ---------------------------------------------------------------------------------------------------
  def _engine(self, hdr: dict):
    key = (str(hdr["model_path"]), int(hdr["layer_start"]), int(hdr["layer_end"]))
    eng = self.engines.get(key)
    if eng is None:
      cfg = json.loads((Path(key[0]) / "config.json").read_text())
      if int(hdr["layer_total"]) != cfg["num_hidden_layers"]:
        raise ValueError(f"layer_total {hdr['layer_total']} but the model has {cfg['num_hidden_layers']} layers")
      if self._factory is not None:
        eng = self._factory(cfg, self.device)
      else:
        from xotorch_b200.native import NativeShard            # raises if libxotb200.so is missing: no CPU fallback
        eng = NativeShard(cfg, self.device)
      eng.load_dir(key[0], key[1], key[2])
      eng._cheetah_cfg = cfg
      self.engines[key] = eng
    return key, eng

  # ---- one request ---------------------------------------------------------------------------------------------------
  def _run(self, hdr: dict, payload: bytes) -> Tuple[np.ndarray, str]:
    key, eng = self._engine(hdr)
    cfg = eng._cheetah_cfg
    skey = (str(hdr.get("session_id")), key)
    if hdr.get("has_hidden_state"):
      shape = [int(x) for x in hdr["hidden_state_shape"]]
      x = np.frombuffer(payload, dtype=np.float32).reshape(shape)            # floats are always sent widened (`:386-389`)
      n = shape[1]
      pos0 = self.sessions.get(skey, 0)
      data = x
    else:
      ishape = [int(x) for x in hdr["input_ids_shape"]]
      mshape = [int(x) for x in hdr["mask_shape"]]
      pshape = [int(x) for x in hdr["input_pos_shape"]]
      it = np.dtype(_NP[hdr["dtype_input_ids"]])
      pt = np.dtype(_NP[hdr["dtype_input_pos"]])
      n_ids, n_mask, n_pos = int(np.prod(ishape)), int(np.prod(mshape)), int(np.prod(pshape)) if pshape else 1
      if len(payload) != n_ids * it.itemsize + n_mask + n_pos * pt.itemsize:
        raise ValueError("payload size does not match the shapes in the header")
      ids = np.frombuffer(payload, dtype=it, count=n_ids).reshape(ishape)
      pos = np.frombuffer(payload, dtype=pt, count=n_pos, offset=n_ids * it.itemsize + n_mask)
      n = ishape[-1]
      pos0 = int(pos.reshape(-1)[0])
      data = ids.reshape(1, n).astype(np.int32)
    if pos0 == 0:                                     # a new request on this shard: KV of prompt + 1024 like the reference
      need = min(n + 1024, int(cfg["max_position_embeddings"]))
      eng.begin_request(need)
    out = eng.forward(data, pos0)
    self.sessions[skey] = pos0 + n
    if out.ndim == 2:                                 # last shard: [1, V] -> [1, 1, V]; the client slices [:, -1]
      out = out.reshape(1, 1, -1)
    if self.reply_dtype == "bfloat16":
      bits = (np.ascontiguousarray(out, dtype=np.float32).view(np.uint32) >> 16).astype(np.uint16)   # exact: values are 16-bit already
      return bits, "bfloat16"
    return np.ascontiguousarray(out, dtype=np.float32), "float32"

  async def _handle(self, reader: asyncio.StreamReader, writer: asyncio.StreamWriter):
    try:
      n = _header_len(await reader.readexactly(4))
      hdr = json.loads((await reader.readexactly(n)).decode("utf-8"))
      if hdr.get("has_hidden_state"):
        size = int(np.prod(hdr["hidden_state_shape"])) * 4
      else:
        size = (int(np.prod(hdr["input_ids_shape"])) * np.dtype(_NP[hdr["dtype_input_ids"]]).itemsize + int(np.prod(hdr["mask_shape"])) +
                max(1, int(np.prod(hdr["input_pos_shape"]))) * np.dtype(_NP[hdr["dtype_input_pos"]]).itemsize)
      payload = await reader.readexactly(size)
      async with self._lock:
        out, dt = await asyncio.get_running_loop().run_in_executor(None, self._run, hdr, payload)
      self.requests += 1
      rh = json.dumps({"shape": list(out.shape), "dtype": dt}).encode("utf-8")
      writer.write(struct.pack("!I", len(rh)) + rh + out.tobytes())
      await writer.drain()
    except Exception as ex:                           # the protocol has no error frame: close, the client raises ConnectionError
      print(f"cheetah_service: request failed: {ex!r}")
    finally:
      writer.close()

  async def start(self):
    if os.path.exists(self.socket_path):
      os.unlink(self.socket_path)
    self.server = await asyncio.start_unix_server(self._handle, path=self.socket_path)

  async def stop(self):
    if self.server is not None:
      self.server.close()
      await self.server.wait_closed()
    for e in self.engines.values():
      if hasattr(e, "close"):
        e.close()
    self.engines.clear()
    if os.path.exists(self.socket_path):
      os.unlink(self.socket_path)


def main():
  import argparse
  ap = argparse.ArgumentParser(description="Cheetah-socket service backed by libxotb200 (B200)")
  ap.add_argument("--socket", default=DEFAULT_SOCKET)
  ap.add_argument("--device", type=int, default=0)
  ap.add_argument("--reply-dtype", default="float32", choices=["float32", "bfloat16"])
  a = ap.parse_args()

  async def run():
    svc = CheetahService(a.socket, a.device, reply_dtype=a.reply_dtype)
    await svc.start()
    print(f"cheetah_service: listening on {a.socket}")
    await asyncio.Event().wait()
  asyncio.run(run())


if __name__ == "__main__":
  main()
