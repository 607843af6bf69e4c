"""Host-side engine interface of the hot path.

The reference defines its engines against an abstract base class (/root/reference/xotorch/inference/inference_engine.py:11-51)
and picks one in a hard-coded factory (:54-74).  This module states the same contract for the one engine this package
ships — identical coroutine names, argument order / meaning and return conventions — so `Node`-style callers and the
reference's own engine test (xotorch/inference/test_inference_engine.py) run unchanged against it:

  encode(shard, prompt)                 -> 1-D or [1, T] integer token ids
  infer_tensor(request_id, shard, x, s) -> (hidden [1, S, D] f32 | last-position logits [1, V] f32, state dict)
                                           x.ndim == 2: token ids, x.ndim == 3: hidden state from the previous shard
  infer_prompt(request_id, shard, prompt, s) = infer_tensor on encode(prompt) reshaped to [1, T]
  sample(logits, temp=..., top_k=...)   -> [B, 1] int32        decode(shard, ids) -> str
  load_checkpoint / save_checkpoint, and a small per-engine key/value `session`

When the reference package is importable `B200ShardInferenceEngine` is additionally registered as a virtual subclass of the
reference's ABC and installed in its factory / model cards by `register_with_xotorch()`."""
from __future__ import annotations

import abc
from typing import Any, Dict, Optional, Tuple

import numpy as np

Result = Tuple[np.ndarray, Optional[dict]]


class InferenceEngine(abc.ABC):
  session: Dict[Any, Any] = {}   # class-level on purpose: the reference shares one dict between engines the same way

  # ---- the step ---------------------------------------------------------------------------------------------------
  @abc.abstractmethod
  async def infer_tensor(self, request_id: str, shard, input_data: np.ndarray, inference_state: Optional[dict] = None) -> Result:
    """Run this shard's layers on `input_data`; see the module docstring for the ndim protocol."""

  async def infer_prompt(self, request_id: str, shard, prompt: str, inference_state: Optional[dict] = None) -> Result:
    ids = np.asarray(await self.encode(shard, prompt))
    return await self.infer_tensor(request_id, shard, ids.reshape(1, ids.size), inference_state)

  # ---- text <-> ids, token choice ------------------------------------------------------------------------------------
  @abc.abstractmethod
  async def encode(self, shard, prompt: str) -> np.ndarray:
    """Tokenise `prompt` for the model of `shard`."""

  @abc.abstractmethod
  async def decode(self, shard, tokens: np.ndarray) -> str:
    """Inverse of `encode`."""

  @abc.abstractmethod
  async def sample(self, x: np.ndarray) -> np.ndarray:
    """Pick the next token id per row of logits."""

  # ---- weights ---------------------------------------------------------------------------------------------------------
  @abc.abstractmethod
  async def load_checkpoint(self, shard, path: str):
    """Make the weights of `shard` resident."""

  async def save_checkpoint(self, shard, path: str):
    return None                     # inference only: nothing to write back

  # ---- session store -------------------------------------------------------------------------------------------------------
  async def save_session(self, key, value):
    type(self).session[key] = value

  async def clear_session(self):
    type(self).session.clear()      # (the reference calls a non-existent dict.empty() here and raises: SURVEY.md App. B8)


# engine name -> class name, the shape of the reference's registry (model cards are keyed by the CLASS name)
inference_engine_classes = {"b200": "B200ShardInferenceEngine"}


def get_inference_engine(inference_engine_name: str, shard_downloader=None) -> InferenceEngine:
  """Factory with the reference's signature; this package provides exactly one engine and no fallback."""
  if inference_engine_name != "b200":
    raise ValueError(f"Unsupported inference engine: {inference_engine_name}")
  from .b200_engine import B200ShardInferenceEngine
  return B200ShardInferenceEngine(shard_downloader)
