"""Tokenizer resolution for a LOCAL model directory (no hub access on the B200 box).  Mirrors what the reference's
engine keeps in `.tokenizer` (/root/reference/xotorch/inference/tokenizers.py:11-63): `.encode`, `.decode`,
`.eos_token_id`.  Synthetic model directories have no tokenizer files -> `IdTokenizer`."""
from __future__ import annotations

from pathlib import Path

import numpy as np


class IdTokenizer:
  """Whitespace-separated integers <-> ids; eos never fires (-1), so synthetic runs generate exactly N tokens."""
  eos_token_id = -1

  def __init__(self, vocab_size: int):
    self.vocab_size = vocab_size

  def encode(self, text, return_tensors=None):
    ids = [int(t) % self.vocab_size for t in str(text).split() if t.lstrip("-").isdigit()] or [0]
    return np.asarray(ids, dtype=np.int64)

  def decode(self, tokens) -> str:
    return " ".join(str(int(t)) for t in np.asarray(tokens).reshape(-1))

  def apply_chat_template(self, conversation, tokenize=True, add_generation_prompt=True, tools=None, **kwargs):
    return " ".join(str(m.get("content", "")) for m in conversation)


def resolve_local_tokenizer(model_dir, vocab_size: int):
  model_dir = Path(model_dir)
  if any((model_dir / f).exists() for f in ("tokenizer.json", "tokenizer.model", "tokenizer_config.json")):
    from transformers import AutoTokenizer
    return AutoTokenizer.from_pretrained(str(model_dir), local_files_only=True)
  return IdTokenizer(vocab_size)
