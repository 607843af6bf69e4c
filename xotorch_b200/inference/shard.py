"""`Shard`: which contiguous, INCLUSIVE range of decoder layers one engine owns.

API-compatible with the value type the reference passes across its engine / peer boundary
(/root/reference/xotorch/inference/shard.py:4-39: constructor arguments, attribute names, the six helper methods and
`shards_overlap`), written here as an immutable tuple type.  The engine never requires THIS class: anything with the four
attributes works (duck typing), so the reference's own `Shard` objects can be handed to `B200ShardInferenceEngine` directly
(tests/test_engine_host_logic.py drives it from the reference's `Node`)."""
from typing import Any, Dict, NamedTuple


class Shard(NamedTuple):
  model_id: str       # model card key (picks the local model directory)
  start_layer: int    # first decoder layer held, 0-based
  end_layer: int      # last decoder layer held, inclusive
  n_layers: int       # decoder layers of the whole model

  # -- position in the pipeline ----------------------------------------------------------------------------------
  def is_first_layer(self) -> bool:
    """Holds layer 0, hence the token embedding: its input is token ids."""
    return not self.start_layer

  def is_last_layer(self) -> bool:
    """Holds the final layer, hence final norm + lm_head: its output is logits."""
    return self.end_layer + 1 == self.n_layers

  def get_layer_count(self) -> int:
    return len(self.layers)

  @property
  def layers(self) -> range:
    return range(self.start_layer, self.end_layer + 1)

  # -- wire form (the reference ships shards as plain dicts inside its RPC messages) --------------------------------
  def to_dict(self) -> Dict[str, Any]:
    return self._asdict()

  @classmethod
  def from_dict(cls, data: Dict[str, Any]) -> "Shard":
    return cls(model_id=data["model_id"], start_layer=int(data["start_layer"]), end_layer=int(data["end_layer"]),
               n_layers=int(data["n_layers"]))

  def overlaps(self, other) -> bool:
    return shards_overlap(self, other)


def shards_overlap(a, b) -> bool:
  """Two shards of the SAME model that share at least one layer."""
  if a.model_id != b.model_id:
    return False
  return a.start_layer <= b.end_layer and b.start_layer <= a.end_layer
