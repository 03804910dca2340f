"""Batched GA generation: the deterministic-crowding loop of the reference with ONE device call per generation.

Reference loop (src/search_organisms.py:166-262): for every pair of parents -- create two children (MLE drive,
recombination or clone + mutate; host logic, consumes the random stream), then for each (parent, child): set the
parent's fitness if it is None, set the child's fitness, keep the fitter one.  `get_placement` consumes no random
numbers and a child's creation never depends on another pair's fitness, so all children
of a generation can be created first, in the same order and with the same random stream, then every missing fitness
evaluated in one batched call, then the competitions resolved in order: the population, the fitness list and the
recognizer counts come out identical to the sequential loop.
"""
from __future__ import annotations

from .objects import place_population, set_population_fitness


def evaluate_missing_fitness(organisms, pos_set, neg_set, method, gamma=0.2, engine=None):
    """`org.set_fitness(...)` for every listed organism whose fitness is None, in one batched device call.
    Duplicates (the same object listed twice) are evaluated once."""
    todo, seen = [], set()
    for org in organisms:
        if org.fitness is None and id(org) not in seen:
            seen.add(id(org))
            todo.append(org)
    if todo:
        set_population_fitness(todo, pos_set, neg_set, method, gamma, engine=engine)
    return len(todo)


def crowding_generation(organism_population, make_children, pos_set, neg_set, method, gamma=0.2, engine=None,
                        reset_parent_fitness=False):
    """One generation of deterministic crowding (search_organisms.py:166-262), batched.

    make_children(parent1, parent2) -> [(parent, child), (parent, child)] is the reference's host-side child creation
    for one pair of parents (`two_parent_child_pairs`, :180-222); it is called pair by pair, in population order.
    reset_parent_fitness mirrors :172-175 (datasets were reshuffled).

    The population list is updated in place; returns (pop_fitness_list, pop_n_recogs_list) as the reference
    accumulates them (:253-262).
    """
    all_pairs = []
    for i in range(0, len(organism_population) - 1, 2):
        parent1, parent2 = organism_population[i], organism_population[i + 1]
        if reset_parent_fitness:
            parent1.fitness = None
            parent2.fitness = None
        pairs = make_children(parent1, parent2)
        all_pairs.append((i, pairs))
    # children always get a fresh fitness (:244-247); parents only if they have none (:239-242)
    for _, pairs in all_pairs:
        for _, child in pairs:
            child.fitness = None
    evaluate_missing_fitness([org for _, pairs in all_pairs for pair in pairs for org in pair], pos_set, neg_set, method,
                             gamma, engine=engine)
    pop_fitness_list, pop_n_recogs_list = [], []
    for i, pairs in all_pairs:
        for j, (parent, child) in enumerate(pairs):
            winner = parent if parent.fitness > child.fitness else child
            organism_population[i + j] = winner
            pop_fitness_list.append(winner.fitness)
            pop_n_recogs_list.append(winner.count_recognizers())
    return pop_fitness_list, pop_n_recogs_list


def parents_placements(parents, dna_set, engine=None):
    """`[[p.get_placement(seq) for seq in dna_set] for p in parents]` in one device call -- the placements the
    MLE drive (search_organisms.py:183,186) and the recombination bookkeeping (organism_factory.py:631-643) need."""
    return place_population(parents, dna_set, engine=engine)
