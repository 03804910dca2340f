"""CPU: the batched generation loop resolves competitions exactly like the reference's sequential loop
(src/search_organisms.py:166-262), with a deterministic stand-in for the device fitness."""
import random

import numpy as np

from flemingo_b200 import generation


class Org:
    def __init__(self, gene):
        self.gene = gene
        self.fitness = None
        self.evaluations = 0

    def count_recognizers(self):
        return 1 + self.gene % 5

    def set_fitness(self, pos, neg, method, gamma=0.2):     # the sequential reference path
        self.evaluations += 1
        self.fitness = fitness_of(self.gene, len(pos), len(neg))


def fitness_of(gene, n_pos, n_neg):
    return ((gene * 2654435761) % 1000003) / 1000.0 + n_pos - n_neg


def sequential(pop, rng, pos, neg):
    fit, nrec = [], []
    for i in range(0, len(pop) - 1, 2):
        pairs = [(pop[i], Org(rng.randrange(10**9))), (pop[i + 1], Org(rng.randrange(10**9)))]
        if rng.random() < 0.3:
            pairs = pairs[::-1]
        for j, (parent, child) in enumerate(pairs):
            if parent.fitness is None:
                parent.set_fitness(pos, neg, "Welch", 0.2)
            child.set_fitness(pos, neg, "Welch", 0.2)
            w = parent if parent.fitness > child.fitness else child
            pop[i + j] = w
            fit.append(w.fitness)
            nrec.append(w.count_recognizers())
    return fit, nrec


def test_batched_generation_equals_sequential_loop(monkeypatch):
    calls = []

    def fake_set_population_fitness(orgs, pos, neg, method, gamma=0.2, engine=None):
        calls.append(len(orgs))
        for o in orgs:
            o.evaluations += 1
            o.fitness = fitness_of(o.gene, len(pos), len(neg))
        return np.zeros((len(orgs), 5))

    monkeypatch.setattr(generation, "set_population_fitness", fake_set_population_fitness)
    pos, neg = ["a"] * 7, ["c"] * 3
    for seed in range(5):
        pop_a = [Org(g) for g in random.Random(seed).sample(range(10**6), 21)]   # odd size: last organism idles
        pop_b = [Org(o.gene) for o in pop_a]
        for o in pop_a[:5] + pop_b[:5]:
            o.fitness = None
        for o, p in zip(pop_a[5:], pop_b[5:]):
            o.fitness = p.fitness = fitness_of(o.gene, 7, 3)
        rng_a, rng_b = random.Random(100 + seed), random.Random(100 + seed)
        for gen in range(4):
            fa, na = sequential(pop_a, rng_a, pos, neg)

            def make_children(p1, p2):
                pairs = [(p1, Org(rng_b.randrange(10**9))), (p2, Org(rng_b.randrange(10**9)))]
                return pairs[::-1] if rng_b.random() < 0.3 else pairs

            calls.clear()
            fb, nb = generation.crowding_generation(pop_b, make_children, pos, neg, "Welch", 0.2)
            assert len(calls) == 1                       # one batched evaluation per generation
            assert fa == fb and na == nb
            assert [o.gene for o in pop_a] == [o.gene for o in pop_b]
        assert max(o.evaluations for o in pop_b) == 1    # nobody is evaluated twice
