"""Multi-GPU partitioning of independent streams.

Streams never exchange data (a stream -- all its channels -- is the atom, coupled only internally through the
maximum-energy channel of the phase chain), so the N-GPU job is N disjoint stream ranges, one process per GPU, and
NO collective on the data path.  ``torch.distributed`` is used only for the start/stop barrier and to combine the
per-rank counters (audio-seconds, device time) at the end.
"""
from typing import List, Sequence, Tuple


def estimated_blocks(n_out_samples: int, interval_samples: int) -> int:
    """Blocks a stream of ``n_out_samples`` output samples needs: one per ``interval_samples`` of output, the first
    at output sample 0 (W#48: samplesSinceLast starts at -1)."""
    return 0 if n_out_samples <= 0 else (int(n_out_samples) + interval_samples - 1) // interval_samples


def partition_streams(costs: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous ranges [lo, hi) per rank, balanced by cumulative cost (block count).  Deterministic, every stream
    in exactly one range, ranges ordered by rank; ranks may be empty when there are fewer streams than ranks."""
    n = len(costs)
    total = float(sum(costs))
    bounds = [0]
    acc = 0.0
    i = 0
    for r in range(1, world_size):
        target = total * r / world_size
        while i < n and acc + costs[i] * 0.5 <= target:
            acc += costs[i]
            i += 1
        bounds.append(i)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def my_range(costs: Sequence[int], rank: int, world_size: int) -> Tuple[int, int]:
    return partition_streams(costs, world_size)[rank]


def combine_counters(values, group=None):
    """Sum per-rank float counters and take the max of per-rank times: returns (sums, maxima) as lists.
    ``values`` = (list_of_sums, list_of_maxima).  Works on gloo (CPU tensors) and nccl (CUDA tensors)."""
    import torch
    import torch.distributed as dist
    sums, maxima = values
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(sums), list(maxima)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    s = torch.tensor(list(sums), dtype=torch.float64, device=dev)
    m = torch.tensor(list(maxima), dtype=torch.float64, device=dev)
    if s.numel():
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    if m.numel():
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return s.tolist(), m.tolist()
