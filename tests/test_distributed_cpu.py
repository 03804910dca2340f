"""CPU (gloo, world_size 2): the N>1 host path -- cost-balanced sharding and the single all-gather."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from flemingo_b200.distributed import evaluate_sharded, gather_stats, organism_costs, shard_by_cost


def test_shard_by_cost_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 200000, 257).astype(float)
    for world in (1, 2, 4, 8):
        shards = shard_by_cost(costs, world)
        allidx = np.sort(np.concatenate(shards))
        assert np.array_equal(allidx, np.arange(costs.size))
        loads = np.array([costs[s].sum() for s in shards])
        assert loads.max() - loads.min() <= costs.max()
        assert shard_by_cost(costs, world)[0].tolist() == shards[0].tolist()  # deterministic


class FakeEngine:
    """Stands in for the GPU engine: statistic k of an organism is a deterministic function of its id."""

    def evaluate(self, organisms, pos, neg, method, gamma):
        return np.array([[float(o.sum_recognizer_lengths) * (k + 1) + len(pos) + 0.5 * len(neg) for k in range(5)] for o in organisms])


class Org:
    def __init__(self, n, sl):
        self.recognizer_types = "p" * n
        self.sum_recognizer_lengths = sl


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orgs = [Org(1 + i % 6, 5 + (7 * i) % 40) for i in range(37)]
    full = evaluate_sharded(FakeEngine(), orgs, ["a" * 100] * 3, ["c" * 100] * 4, rank=rank, world_size=world)
    out[rank] = full
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gathers_whole_population():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    orgs = [Org(1 + i % 6, 5 + (7 * i) % 40) for i in range(37)]
    want = FakeEngine().evaluate(orgs, ["a"] * 3, ["c"] * 4, "Welch", 0.2)
    assert np.array_equal(out[0], want) and np.array_equal(out[1], want)
    costs = organism_costs(orgs, 100)
    shards = shard_by_cost(costs, 2)
    assert abs(costs[shards[0]].sum() - costs[shards[1]].sum()) <= costs.max()


def test_gather_stats_single_rank_identity():
    stats = np.arange(15.0).reshape(3, 5)
    assert np.array_equal(gather_stats(stats, [np.arange(3)], 0, 1), stats)


class RecordingEngine:
    """Stand-in for one device of MultiDeviceEvaluator: records the call order, answers like FakeEngine."""

    def __init__(self, name, log):
        self.name, self.log, self.orgs, self.sets = name, log, None, {}

    @staticmethod
    def join_sequences(seqs):
        return "".join(seqs).encode(), np.array([len(s) for s in seqs])

    def upload_sequences(self, seqs, set_id, reuse=False, joined=None):
        assert joined is not None and reuse          # joined once by the evaluator, resident across generations
        self.sets[set_id] = list(seqs)
        return set_id

    def upload_population(self, organisms):
        self.orgs = list(organisms)
        self.log.append((self.name, "upload", len(self.orgs)))

    def place(self, sset, fetch=True):
        assert fetch is False                        # asynchronous: nothing is read back before every device is busy
        self.log.append((self.name, "place", sset))

    def fitness(self, pos, neg, method, gamma):
        self.log.append((self.name, "fitness"))
        return FakeEngine().evaluate(self.orgs, self.sets[pos], self.sets[neg], method, gamma)


def test_multi_device_evaluator_issues_everything_before_collecting():
    from flemingo_b200.distributed import MultiDeviceEvaluator
    log = []
    ev = MultiDeviceEvaluator(engines=[RecordingEngine(f"gpu{d}", log) for d in range(3)])
    orgs = [Org(1 + i % 6, 5 + (7 * i) % 40) for i in range(41)]
    pos, neg = ["a" * 100] * 3, ["c" * 100] * 4
    got = ev(orgs, pos, neg, "Welch", 0.2)
    assert np.array_equal(got, FakeEngine().evaluate(orgs, pos, neg, "Welch", 0.2))
    first_fitness = next(i for i, e in enumerate(log) if e[1] == "fitness")
    assert all(e[1] != "fitness" for e in log[:first_fitness]) and all(e[1] == "fitness" for e in log[first_fitness:])
    assert sorted(e[2] for e in log if e[1] == "upload") and sum(e[2] for e in log if e[1] == "upload") == 41
    # fewer organisms than devices: idle devices are skipped
    log.clear()
    assert np.array_equal(ev(orgs[:2], pos, neg), FakeEngine().evaluate(orgs[:2], pos, neg, "Welch", 0.2))
    assert len([e for e in log if e[1] == "upload"]) == 2
