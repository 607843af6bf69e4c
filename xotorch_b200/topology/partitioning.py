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
