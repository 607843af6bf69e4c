"""`InferenceEngine` — host-side mirror of the reference's plugin interface for the hot path
(/root/reference/xotorch/inference/inference_engine.py:11-51), same method names, argument meaning and error
behaviour, so the parity tests read like the reference's own (xotorch/inference/test_inference_engine.py).

When the reference package is importable, `B200ShardInferenceEngine` ALSO registers as a virtual subclass of the
reference's ABC and `register_with_xotorch()` installs it in the reference's hard-coded factory / model cards
(SURVEY.md §8b "Registration")."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional, Tuple

import numpy as np

from .shard import Shard


class InferenceEngine(ABC):
  session = {}

  @abstractmethod
  async def encode(self, shard: Shard, prompt: str) -> np.ndarray: ...

  @abstractmethod
  async def sample(self, x: np.ndarray) -> np.ndarray: ...

  @abstractmethod
  async def decode(self, shard: Shard, tokens: np.ndarray) -> str: ...

  @abstractmethod
  async def infer_tensor(self, request_id: str, shard: Shard, input_data: np.ndarray,
                         inference_state: Optional[dict] = None) -> Tuple[np.ndarray, Optional[dict]]: ...

  @abstractmethod
  async def load_checkpoint(self, shard: Shard, path: str): ...

  async def save_checkpoint(self, shard: Shard, path: str):
    pass

  async def save_session(self, key, value):
    self.session[key] = value

  async def clear_session(self):
    self.session.clear()          # the reference calls dict.empty() here and crashes (App. B8) — not replicated

  async def infer_prompt(self, request_id: str, shard: Shard, prompt: str,
                         inference_state: Optional[dict] = None) -> Tuple[np.ndarray, Optional[dict]]:
    tokens = await self.encode(shard, prompt)
    x = np.asarray(tokens).reshape(1, -1)
    return await self.infer_tensor(request_id, shard, x, inference_state)


inference_engine_classes = {"b200": "B200ShardInferenceEngine"}


def get_inference_engine(inference_engine_name: str, shard_downloader=None):
  """Mirror of the reference factory (inference_engine.py:61-74) for the one engine this package provides."""
  if inference_engine_name == "b200":
    from .b200_engine import B200ShardInferenceEngine
    return B200ShardInferenceEngine(shard_downloader)
  raise ValueError(f"Unsupported inference engine: {inference_engine_name}")
