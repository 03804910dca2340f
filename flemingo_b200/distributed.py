"""Multi-GPU evaluation: one process per GPU, organisms sharded across ranks, sequences replicated.

(organism, sequence) pairs are independent and an organism's fitness needs all of ITS energies, so the
population is partitioned by organism (cost-balanced, largest-first greedy) and every rank reduces the fitness
of its own organisms locally; the only exchange is one all-gather of the [n_org, 5] statistics per generation
(SURVEY.md section 8e).  This replaces the reference's MPI scatter/gather of pickled organisms
(src/search_organisms.py:140-143, :268-277, fragment_population :1145-1198).
"""
from __future__ import annotations

import numpy as np


def shard_by_cost(costs, world_size: int) -> list:
    """Largest-processing-time-first greedy partition; returns world_size index arrays (each ascending).

    Ties are broken deterministically so every rank computes the same partition without communication.
    """
    costs = np.asarray(costs, dtype=np.float64)
    order = np.lexsort((np.arange(costs.size), -costs))
    loads = np.zeros(world_size)
    counts = np.zeros(world_size, dtype=np.int64)
    buckets = [[] for _ in range(world_size)]
    for idx in order:
        # least loaded rank; among equals the one with fewer organisms, then the lowest rank
        r = int(np.lexsort((np.arange(world_size), counts, loads))[0])
        buckets[r].append(int(idx))
        loads[r] += costs[idx]
        counts[r] += 1
    return [np.array(sorted(b), dtype=np.int64) for b in buckets]


def organism_costs(organisms, seq_len: int) -> np.ndarray:
    """(n-1) * P(P+1)/2 DP cells + P * sum(L) score adds per placement."""
    out = np.empty(len(organisms))
    for i, org in enumerate(organisms):
        n = len(org.recognizer_types)
        sl = int(org.sum_recognizer_lengths)
        P = max(seq_len - sl + 1, 1)
        out[i] = (n - 1) * P * (P + 1) / 2 + P * sl
    return out


def gather_stats(local_stats: np.ndarray, shards: list, rank: int, world_size: int, device=None, group=None) -> np.ndarray:
    """All-gathers per-organism statistics ([len(shard), k] per rank) into the full [n_org, k] array.

    One collective (all_gather_into_tensor of equally padded blocks); NCCL when `device` is a CUDA device,
    gloo on CPU.
    """
    import torch
    import torch.distributed as dist
    n_org = int(sum(len(s) for s in shards))
    k = local_stats.shape[1] if local_stats.ndim == 2 else 5
    if world_size == 1:
        full = np.empty((n_org, k))
        full[shards[0]] = local_stats
        return full
    width = max(len(s) for s in shards)
    dev = torch.device("cpu") if device is None else device
    block = torch.zeros((width, k), dtype=torch.float64, device=dev)
    if len(shards[rank]):
        block[: len(shards[rank])] = torch.from_numpy(np.ascontiguousarray(local_stats)).to(dev)
    gathered = torch.empty((world_size * width, k), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gathered, block, group=group)
    host = gathered.cpu().numpy().reshape(world_size, width, k)
    full = np.empty((n_org, k))
    for r, shard in enumerate(shards):
        full[shard] = host[r, : len(shard)]
    return full


def evaluate_sharded(engine, organisms, pos_sequences, neg_sequences, method: str = "Welch", gamma: float = 0.2,
                     rank: int = 0, world_size: int = 1, device=None, group=None, shards=None) -> np.ndarray:
    """Fitness statistics [n_org, 5] of the WHOLE population on every rank; this rank places only its shard."""
    if shards is None:
        seq_len = max((len(s) for s in pos_sequences), default=1)
        shards = shard_by_cost(organism_costs(organisms, seq_len), world_size)
    mine = [organisms[i] for i in shards[rank]]
    local = engine.evaluate(mine, pos_sequences, neg_sequences, method, gamma) if mine else np.zeros((0, 5))
    return gather_stats(local, shards, rank, world_size, device, group)
