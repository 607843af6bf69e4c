"""The Cheetah-socket façade (xotorch_b200/cheetah_service.py) against a client that restates the reference's
`run_cheetah` wire code line by line (/root/reference/xotorch/inference/cheetah/sharded_inference_engine.py:343-441; the
client itself cannot be imported: it needs torchtune).  CPU part: framing, the "!f" length slip, session positions, with a
stand-in engine.  GPU part (-m gpu): two real shards behind the socket give bit for bit what the C ABI gives directly."""
import asyncio
import json
import socket
import struct

import numpy as np
import pytest

from xotorch_b200.cheetah_service import CheetahService
from xotorch_b200.synth import PRESETS, hf_config, synth_prompt, write_hf_dir

NAME = "tiny-llama-hd64"


def reference_client(sock_path, header, tokens=None, input_pos=None, mask=None, hidden_state=None):
  """run_cheetah, synchronous: same bytes on the wire (:357-395), same parsing of the reply (:401-441)"""
  s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
  s.connect(sock_path)
  h = dict(header)
  if hidden_state is None:
    h.update(input_ids_shape=list(tokens.shape), mask_shape=list(mask.shape), input_pos_shape=list(input_pos.shape),
             dtype_input_ids="int32", dtype_mask="bool", dtype_input_pos="int64", has_hidden_state=False)
    hb = json.dumps(h).encode("utf-8")
    s.sendall(struct.pack("!I", len(hb)) + hb + tokens.astype(np.int32).tobytes() + mask.astype(np.uint8).tobytes() + input_pos.astype(np.int64).tobytes())
  else:
    h.update(hidden_state_dtype="bfloat16", hidden_state_shape=list(hidden_state.shape), has_hidden_state=True)
    hb = json.dumps(h).encode("utf-8")
    s.sendall(struct.pack("!f", len(hb)) + hb + hidden_state.astype(np.float32).tobytes())      # sic: "!f" (:388)

  def recv(n):
    buf = bytearray()
    while len(buf) < n:
      c = s.recv(n - len(buf))
      if not c:
        raise ConnectionError("incomplete reply")
      buf += c
    return bytes(buf)
  rh = json.loads(recv(struct.unpack("!I", recv(4))[0]).decode("utf-8"))
  shape, dtype = rh["shape"], rh["dtype"]
  if dtype == "bfloat16":
    out = (np.frombuffer(recv(int(np.prod(shape)) * 2), dtype=np.uint16).reshape(shape).astype(np.uint32) << 16).view(np.float32)
  else:
    out = np.frombuffer(recv(int(np.prod(shape)) * 4), dtype=np.float32).reshape(shape)
  s.close()
  return out if hidden_state is not None else out[:, -1]          # :443-457


def _header(model_dir, a, b, nl, session="s1"):
  return {"node_id": "n", "model": NAME, "model_path": str(model_dir), "layer_start": a, "layer_end": b, "layer_total": nl, "session_id": session}


def _drive(svc_kwargs, fn):
  """run `fn(socket_path)` in a worker thread while the service's loop runs here"""
  async def main():
    svc = CheetahService(**svc_kwargs)
    await svc.start()
    try:
      return await asyncio.get_running_loop().run_in_executor(None, fn, svc_kwargs["socket_path"]), svc.requests
    finally:
      await svc.stop()
  return asyncio.run(main())


class FakeNative:
  def __init__(self, cfg, device): self.cfg, self.calls = cfg, []
  def load_dir(self, d, a, b): self.a, self.b = a, b
  def begin_request(self, n): self.calls.append(("begin", n))
  def close(self): pass

  def forward(self, x, pos0):
    self.calls.append(("forward", x.ndim, tuple(x.shape), pos0))
    if self.b == self.cfg["num_hidden_layers"] - 1:
      out = np.zeros((1, self.cfg["vocab_size"]), np.float32); out[0, pos0 % 7] = 1.0
      return out
    return np.full((1, x.shape[1], self.cfg["hidden_size"]), np.float32(pos0), np.float32)


def test_cheetah_framing_and_session_positions_cpu(tmp_path):
  shape = PRESETS[NAME]
  d = write_hf_dir(tmp_path / NAME, shape, 1234, n_files=1)
  made = []

  def factory(cfg, dev):
    made.append(FakeNative(cfg, dev)); return made[-1]

  def client(path):
    L = 5 + 1024
    ids = np.arange(5, dtype=np.int32).reshape(1, 5)
    h0 = reference_client(path, _header(d, 0, 1, 4), tokens=ids, input_pos=np.arange(5), mask=np.tril(np.ones((L, L), bool))[None, :5])
    l0 = reference_client(path, _header(d, 2, 3, 4), hidden_state=np.zeros((1, 5, shape.dim), np.float32))
    h1 = reference_client(path, _header(d, 0, 1, 4), tokens=np.array([[9]], np.int32), input_pos=np.array([5]), mask=np.ones((1, 1, L), bool))
    l1 = reference_client(path, _header(d, 2, 3, 4), hidden_state=np.zeros((1, 1, shape.dim), np.float32))
    return h0, l0, h1, l1
  (h0, l0, h1, l1), nreq = _drive(dict(socket_path=str(tmp_path / "sock"), native_factory=factory), client)
  assert nreq == 4 and len(made) == 2
  assert h0.shape == (1, shape.dim) and l0.shape == (1, 1, shape.vocab) and h1.shape == (1, shape.dim)      # token requests: out[:, -1]
  first, last = made
  assert first.calls == [("begin", 5 + 1024), ("forward", 2, (1, 5), 0), ("forward", 2, (1, 1), 5)]
  # the hidden-state request carries no position: the service counted the session's tokens (5, then 1 more)
  assert last.calls == [("begin", 5 + 1024), ("forward", 3, (1, 5, shape.dim), 0), ("forward", 3, (1, 1, shape.dim), 5)]
  assert l0[0, 0, 0] == 1.0 and l1[0, 0, 5] == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("reply", ["float32", "bfloat16"])
def test_cheetah_service_equals_c_abi_gpu(tmp_path, reply):
  from xotorch_b200.native import NativeShard
  shape = PRESETS[NAME]
  d = write_hf_dir(tmp_path / NAME, shape, 1234, n_files=1)
  prompt = synth_prompt(shape.vocab, 9)
  a = NativeShard(hf_config(shape), 0); a.load_dir(d, 0, 1)
  b = NativeShard(hf_config(shape), 0); b.load_dir(d, 2, 3)
  a.begin_request(9 + 1024); b.begin_request(9 + 1024)
  want, x, pos = [], prompt.astype(np.int32), 0
  for _ in range(6):
    h = a.forward(x, pos); lg = b.forward(h, pos)
    want.append((h.copy(), lg.copy()))
    pos += x.shape[1]
    x = np.array([[int(np.argmax(lg[0]))]], dtype=np.int32)
  a.close(); b.close()

  def client(path):
    L = 9 + 1024
    got, x, pos = [], prompt.astype(np.int32), 0
    for _ in range(6):
      n = x.shape[1]
      mask = np.tril(np.ones((L, L), bool))[None, :n] if pos == 0 else (np.arange(L) <= pos)[None, None, :]
      h_last = reference_client(path, _header(d, 0, 1, 4), tokens=x, input_pos=np.arange(pos, pos + n), mask=mask)   # [1, D]: the client's [:, -1]
      # For a prompt the reference client's `[:, -1]` has already dropped all but the last row of the hidden state (an upstream
      # defect of the Cheetah path); the next shard is therefore fed the full [1, S, D] state of the direct run, whose last
      # row is asserted below to be bit-identical to what came back through the socket.
      got.append(h_last)
      x_h = want[len(got) - 1][0]                                   # [1, S, D] from the direct run (bit-identical, asserted below on its last row)
      lg = reference_client(path, _header(d, 2, 3, 4), hidden_state=x_h)
      got[-1] = (h_last, lg)
      pos += n
      x = np.array([[int(np.argmax(lg[0, 0]))]], dtype=np.int32)
    return got
  got, nreq = _drive(dict(socket_path=str(tmp_path / "sock"), device=0, reply_dtype=reply), client)
  assert nreq == 12
  for (h_last, lg), (wh, wl) in zip(got, want):
    assert np.array_equal(h_last, wh[:, -1]) and np.array_equal(lg[:, 0], wl)
