"""Multi-GPU evaluation: one process per GPU, organisms sharded across ranks, sequences replicated.

(organism, sequence) pairs are independent and an organism's fitness needs all of ITS energies, so the
population is partitioned by organism (cost-balanced, largest-first greedy) and every rank reduces the fitness
of its own organisms locally; the only exchange is one all-gather of the [n_org, 5] statistics per generation
(SURVEY.md section 8e).  This replaces the reference's MPI scatter/gather of pickled organisms
(src/search_organisms.py:140-143, :268-277, fragment_population :1145-1198).
"""
from __future__ import annotations

import numpy as np

from .builders import fitness_finalize


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


class DeviceGather:
    """The one collective of the multi-GPU path, without host hops: every rank's [width, 5] block of statistics is
    all-gathered straight out of the library's device buffer (Engine.fitness_dev) on the engine's own stream, and the
    gathered [world, width, 5] array comes back to the host in ONE copy through a pinned buffer.

    Replaces the reference's comm.gather of pickled organisms / fitness lists (src/search_organisms.py:268-277).
    """

    def __init__(self, engine, shards: list, rank: int, world_size: int, device, group=None):
        import torch
        self.engine, self.shards, self.rank, self.world, self.group = engine, shards, rank, world_size, group
        self.device = device
        self.width = max(1, max(len(s) for s in shards))
        self.n_org = int(sum(len(s) for s in shards))
        self.stream = torch.cuda.ExternalStream(engine.stream_handle, device=device)
        self.gathered = torch.empty((world_size * self.width, 5), dtype=torch.float64, device=device)
        self.host = torch.empty((world_size * self.width, 5), dtype=torch.float64).pin_memory()
        self._view, self._view_key = None, None
        # scatter map: row of the gathered array -> organism index
        self._rows = np.concatenate([r * self.width + np.arange(len(s)) for r, s in enumerate(shards)]) if self.n_org else np.zeros(0, np.int64)
        self._orgs = np.concatenate([np.asarray(s, dtype=np.int64) for s in shards]) if self.n_org else np.zeros(0, np.int64)

    def __call__(self, pos, neg, method: str = "Welch", gamma: float = 0.2, empty: bool = False) -> np.ndarray:
        """fitness reduction of this rank's organisms + all-gather; returns the full [n_org, 5] array on every rank.
        empty=True: this rank owns no organism (more ranks than organisms) and contributes a block of zeros."""
        import torch
        import torch.distributed as dist
        stats = None if empty else self.engine.fitness_dev(pos, neg, method, gamma, pad_rows=self.width)
        with torch.cuda.stream(self.stream):
            if empty:
                block = torch.zeros((self.width, 5), dtype=torch.float64, device=self.device)
            else:
                key = (stats.ptr, stats.rows)
                if self._view_key != key:   # the library's buffer only moves when it grows
                    self._view, self._view_key = torch.as_tensor(stats, device=self.device), key
                block = self._view[: self.width]
            if self.world > 1:
                dist.all_gather_into_tensor(self.gathered, block, group=self.group)
                self.host.copy_(self.gathered, non_blocking=True)
            else:
                self.host[: self.width].copy_(block, non_blocking=True)
        self.stream.synchronize()
        self.engine.sync()          # surfaces device-side errors (NaN null bin) like fl_fitness does
        full = np.empty((self.n_org, 5))
        full[self._orgs] = self.host.numpy()[self._rows]
        return fitness_finalize(full)     # column 0 with the reference's host arithmetic, as fl_fitness does


class MultiDeviceEvaluator:
    """ONE process driving every GPU of the box: one Engine (CUDA context + stream) per device, organisms sharded by
    cost, sequences replicated and kept resident, placement launched asynchronously on all devices before the first
    result is collected.  This is the shape the GA driver needs -- the population lives in one Python process
    (src/search_organisms.py holds it in a list) -- and it replaces the reference's mpi4py scatter / gather of pickled
    organisms (:140-143, :268-277) without replaying the GA on several ranks.  No collective is involved: the host
    process is the gather point.  (bench.py's one-rank-per-GPU mode uses DeviceGather / NCCL instead.)

        evaluator = MultiDeviceEvaluator()            # all visible devices
        stats = evaluator(organisms, pos, neg, "Welch", 0.2)      # [n_org, 5], same doubles as a single engine
    """

    def __init__(self, devices=None, engines=None):
        """devices: CUDA device indices (default: all visible); engines: ready-made engines instead (tests)."""
        if engines is not None:
            self.engines = list(engines)
            return
        from . import _cabi
        from .engine import Engine
        if devices is None:
            devices = list(range(max(1, _cabi.load_library().fl_device_count())))
        self.engines = [Engine(d) for d in devices]

    def __call__(self, organisms, pos_sequences, neg_sequences, method: str = "Welch", gamma: float = 0.2) -> np.ndarray:
        n_dev = len(self.engines)
        seq_len = max((len(s) for s in pos_sequences), default=1)
        shards = shard_by_cost(organism_costs(organisms, seq_len), n_dev)
        full = np.empty((len(organisms), 5))
        busy = []
        joined_pos = self.engines[0].join_sequences(pos_sequences)
        joined_neg = self.engines[0].join_sequences(neg_sequences)
        for eng, shard in zip(self.engines, shards):      # issue: uploads + asynchronous placement, device by device
            if not len(shard):
                continue
            pos = eng.upload_sequences(pos_sequences, 0, reuse=True, joined=joined_pos)
            neg = eng.upload_sequences(neg_sequences, 1, reuse=True, joined=joined_neg)
            eng.upload_population([organisms[i] for i in shard])
            eng.place(pos, fetch=False)
            eng.place(neg, fetch=False)
            busy.append((eng, shard, pos, neg))
        for eng, shard, pos, neg in busy:                 # collect: the fitness reduction synchronises its device
            full[shard] = eng.fitness(pos, neg, method, gamma)
        return full

    def close(self):
        for eng in self.engines:
            eng.close()


def evaluate_sharded(engine, organisms, pos_sequences, neg_sequences, method: str = "Welch", gamma: float = 0.2,
                     rank: int = 0, world_size: int = 1, device=None, group=None, shards=None, reuse_sets: bool = True,
                     gather: "DeviceGather | None" = None) -> np.ndarray:
    """Fitness statistics [n_org, 5] of the WHOLE population on every rank; this rank places only its shard.

    On a CUDA device the statistics never visit the host before the collective (DeviceGather); with device=None
    (gloo, CPU tests with a stand-in engine) they go through gather_stats."""
    if shards is None:
        seq_len = max((len(s) for s in pos_sequences), default=1)
        shards = shard_by_cost(organism_costs(organisms, seq_len), world_size)
    mine = [organisms[i] for i in shards[rank]]
    on_gpu = device is not None and getattr(device, "type", "cpu") == "cuda"
    if not on_gpu:
        local = engine.evaluate(mine, pos_sequences, neg_sequences, method, gamma) if mine else np.zeros((0, 5))
        return gather_stats(local, shards, rank, world_size, device, group)
    if gather is None:
        gather = DeviceGather(engine, shards, rank, world_size, device, group)
    # sequences first: their copies and the packing kernel run while the host packs the population
    pos = engine.upload_sequences(pos_sequences, 0, reuse=reuse_sets)
    neg = engine.upload_sequences(neg_sequences, 1, reuse=reuse_sets)
    if mine:
        engine.upload_population(mine)
        engine.place(pos, fetch=False)
        engine.place(neg, fetch=False)
    return gather(pos, neg, method, gamma, empty=not mine)
