"""Ring partitioning by per-GPU HBM — same rule and same rounding as the reference so every node computes the same
layer ranges and the reference's golden vectors hold (tests/test_partitioning.py):

* `RingMemoryWeightedPartitioningStrategy.partition`
    /root/reference/xotorch/topology/ring_memory_weighted_partitioning_strategy.py:7-18
  nodes ordered by (memory, id) descending; each gets the fraction memory/total; running end rounded to 5 decimals.
* `map_partitions_to_shards`
    /root/reference/xotorch/topology/partitioning_strategy.py:24-42
  layers [int(start*n), int(end*n)-1]; the last partition is forced to end at n-1; empty ranges are dropped.

`HBMWeightedRingPartitioningStrategy` is the B200 flavour: `memory` is what the GPU can actually give to weights +
KV (free HBM from cudaMemGetInfo / NVML, in MB like DeviceCapabilities.memory), one logical node per GPU (the
reference reports a multi-GPU host as ONE node with summed memory, device_capabilities.py:215-257 — App. B10).
Works on the reference's `Topology` (duck-typed: `.all_nodes()` yielding `(node_id, caps)` with `caps.memory`)
or on a plain `{node_id: memory}` dict.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Dict, Iterable, List, Tuple, Union

from xotorch_b200.inference.shard import Shard


@dataclass
class Partition:
  node_id: str
  start: float
  end: float


class PartitioningStrategy(ABC):
  @abstractmethod
  def partition(self, topology) -> List[Partition]:
    ...


def _nodes(topology) -> List[Tuple[str, float]]:
  if isinstance(topology, dict):
    return [(k, float(getattr(v, "memory", v))) for k, v in topology.items()]
  return [(nid, float(caps.memory)) for nid, caps in topology.all_nodes()]


class RingMemoryWeightedPartitioningStrategy(PartitioningStrategy):
  def partition(self, topology) -> List[Partition]:
    nodes = sorted(_nodes(topology), key=lambda kv: (kv[1], kv[0]), reverse=True)
    total = sum(m for _, m in nodes)
    out, start = [], 0
    for nid, mem in nodes:
      end = round(start + (mem / total), 5)
      out.append(Partition(nid, start, end))
      start = end
    return out


class HBMWeightedRingPartitioningStrategy(RingMemoryWeightedPartitioningStrategy):
  """Same rule; `memory` = usable HBM per GPU.  `gpu_memory_mb()` probes the local devices."""

  @staticmethod
  def gpu_memory_mb() -> Dict[str, int]:
    import torch
    out = {}
    for i in range(torch.cuda.device_count()):
      free, _total = torch.cuda.mem_get_info(i)
      out[f"cuda:{i}"] = int(free // (1024 * 1024))
    return out


def model_stream_bytes(hf_config: dict) -> Tuple[int, int]:
  """(bytes one decoder layer streams per decoded token, bytes the last shard streams on top: final norm + lm_head), 16-bit
  weights, from an HF config.json dict — SURVEY.md §8d's per-token byte formula (sparse MoE: the active experts only)."""
  D, F = hf_config["hidden_size"], hf_config["intermediate_size"]
  Hq, Hkv = hf_config["num_attention_heads"], hf_config["num_key_value_heads"]
  hd = hf_config.get("head_dim") or D // Hq
  E, k = int(hf_config.get("num_local_experts") or 0), int(hf_config.get("num_experts_per_tok") or 0)
  mlp = (E * D + k * 3 * D * F) if E else 3 * D * F
  layer = 2 * (D * Hq * hd + 2 * D * Hkv * hd + Hq * hd * D + mlp + 2 * D)
  head = 2 * (hf_config["vocab_size"] * D + D)
  return layer, head


class ByteBalancedRingPartitioningStrategy(RingMemoryWeightedPartitioningStrategy):
  """Ring partition that balances the BYTES A SHARD STREAMS PER TOKEN (= its decode time on an HBM-bound engine) instead of
  the layer count: the last shard also streams final norm + lm_head (1.05 GB for Llama-3-8B = 2.4 decoder layers), so under
  the reference rule it sets the period of a ring with several requests in flight (8B on 8 GPUs: 2.79 GB vs
  1.74 GB elsewhere).  Same node order as the reference ((memory, id) descending) and the same `Partition` fractions ->
  `map_partitions_to_shards` pipeline, so every node still derives the same ranges; node i's share of the bytes is its share
  of the memory.  Boundaries are whole layers: cumulative byte target T_i -> the layer count nearest to it (each node keeps
  at least one layer); the fraction handed to map_partitions_to_shards is (layers + 0.5) / n so its floor is exact.
  8B / 8 equal GPUs: 4,5,4,4,5,4,4,2 layers (max 2.18 GB vs 2.79 GB); 8B / 2: 17 + 15; with head_bytes = 0 and a model whose
  layers divide evenly it reproduces the reference's ranges."""

  def __init__(self, n_layers: int, layer_bytes: int, head_bytes: int):
    self.n_layers, self.layer_bytes, self.head_bytes = int(n_layers), int(layer_bytes), int(head_bytes)

  @classmethod
  def for_model(cls, hf_config: dict) -> "ByteBalancedRingPartitioningStrategy":
    layer, head = model_stream_bytes(hf_config)
    return cls(hf_config["num_hidden_layers"], layer, head)

  def layer_counts(self, weights: List[float]) -> List[int]:
    n, N = self.n_layers, len(weights)
    if N > n:
      raise ValueError(f"{N} nodes but only {n} layers")
    total = n * self.layer_bytes + self.head_bytes
    wsum = float(sum(weights))
    bounds, cum, prev = [], 0.0, 0
    for i in range(N - 1):
      cum += weights[i] / wsum
      k = int(round(cum * total / self.layer_bytes))          # cumulative layers whose bytes are nearest to the target
      k = max(prev + 1, min(k, n - (N - 1 - i)))              # every node keeps at least one layer
      bounds.append(k)
      prev = k
    bounds.append(n)
    return [b - a for a, b in zip([0] + bounds[:-1], bounds)]

  def partition(self, topology) -> List[Partition]:
    nodes = sorted(_nodes(topology), key=lambda kv: (kv[1], kv[0]), reverse=True)
    counts = self.layer_counts([m for _, m in nodes])
    out, start, cum = [], 0, 0
    for i, ((nid, _), k) in enumerate(zip(nodes, counts)):
      cum += k
      end = 1.0 if i == len(nodes) - 1 else round((cum + 0.5) / self.n_layers, 5)
      out.append(Partition(nid, start, end))
      start = end
    return out


def map_partitions_to_shards(partitions: List[Partition], num_layers: int, model_id: str) -> List[Shard]:
  shards: List[Shard] = []
  last = len(partitions) - 1
  for i, p in enumerate(partitions):
    lo = int(p.start * num_layers)
    hi = num_layers - 1 if i == last else int(p.end * num_layers) - 1
    if lo <= hi:
      shards.append(Shard(model_id, lo, hi, num_layers))
  if shards and shards[-1].end_layer < num_layers - 1:
    tail = shards[-1]
    shards[-1] = Shard(model_id, tail.start_layer, num_layers - 1, num_layers)
  return shards


def shard_for(node_id: str, topology, num_layers: int, model_id: str, strategy: PartitioningStrategy | None = None) -> Shard:
  """What Node.get_current_shard does (/root/reference/xotorch/orchestration/node.py:455-460)."""
  parts = (strategy or RingMemoryWeightedPartitioningStrategy()).partition(topology)
  shards = map_partitions_to_shards(parts, num_layers, model_id)
  idx = [p.node_id for p in parts].index(node_id)
  if idx >= len(shards):
    raise ValueError(f"node {node_id} received no layers")
  return shards[idx]
