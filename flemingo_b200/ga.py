"""The reference's GA driver loop with ONE batched device call per generation.

`run(so)` takes the reference's own `search_organisms` module (imported from its src/ directory, after `set_up()`)
and executes what `search_organisms.main()` (src/search_organisms.py:70-363) executes -- same configuration globals,
same helper functions of that module (FASTA reading, population initialisation, shuffles, statistics, exports), same
consumption order of the `random` stream -- except that a generation is run in three phases instead of pair by pair:

    breed     for every pair of parents, in population order: MLE drive / recombination / clone + mutate, check_size
              (:166-222).  Host logic of the reference, untouched; it consumes the random stream.
    evaluate  every parent without a fitness and every child, in one batched placement + fitness reduction on the
              GPU(s) (replaces the per-organism set_fitness calls of :239-247).
    select    the parent/child competitions, in the original order (:250-262).

`get_placement` consumes no random numbers and a child's creation never looks at another pair's fitness, so the
population, the fitness lists and every export come out identical to the sequential loop (SURVEY.md 3.2;
tests/test_reference_ga_gpu.py compares complete runs).  The MLE generations need every parent placed on the positive
set (:183,:186): those placements are computed for all parents in one traced batch as well.
"""
from __future__ import annotations

import random
import time

import numpy as np

from .engine import get_engine
from .formats import gini_RSV
from .objects import build_placement


class GenerationTimes:
    """Wall-clock breakdown of the generations run so far (seconds)."""

    def __init__(self):
        self.breed = self.evaluate = self.select = self.report = 0.0
        self.generations = 0
        self.evaluated = 0

    def as_dict(self):
        return {"generations": self.generations, "organisms_evaluated": self.evaluated, "breed_s": self.breed,
                "evaluate_s": self.evaluate, "select_s": self.select, "report_s": self.report}


def place_parents(parents, dna_set, placement_cls, engine=None):
    """`[[p.get_placement(seq) for seq in dna_set] for p in parents]` in one traced device call, as objects of the
    caller's PlacementObject class (placement_object.py:9-52)."""
    eng = engine or get_engine()
    eng.upload_population(parents)
    sset = eng.upload_sequences(dna_set, 2)          # set slot 2: the fitness sets stay resident in slots 0 and 1
    res = eng.place(sset, trace=True)
    out = []
    for o, org in enumerate(parents):
        out.append([build_placement(org._id, seq, org.recognizer_types, org.recognizer_lengths, res.energies[o, s],
                                    res.rec_scores[o, s], res.con_scores[o, s], res.gaps[o, s], placement_cls)
                    for s, seq in enumerate(dna_set)])
    return out


def evaluate_fitness(organisms, pos_set, neg_set, method, gamma, evaluator=None) -> int:
    """`org.set_fitness(pos_set, neg_set, method, gamma)` for every listed organism (each object once), batched."""
    todo, seen = [], set()
    for org in organisms:
        if id(org) not in seen:
            seen.add(id(org))
            todo.append(org)
    if not todo:
        return 0
    if evaluator is None:
        stats = get_engine().evaluate(todo, pos_set, neg_set, method, gamma)
    else:
        stats = evaluator(todo, pos_set, neg_set, method, gamma)
    for org, row in zip(todo, stats):
        org.fitness = row[0]
    return len(todo)


def breed_pair(so, factory, parent1, parent2, generation, positive_dataset, placements_of=None):
    """Child creation for one pair of parents (search_organisms.py:176-222); returns the two (parent, child) pairs."""
    if (generation + 1) % factory.periodic_mle == 0:
        # MLE-based memetic drive (:180-189); the placements come from the generation's traced batch
        child1 = factory.mle_org(parent1, placements_of[id(parent1)])
        child2 = factory.mle_org(parent2, placements_of[id(parent2)])
        return [(parent1, child1), (parent2, child2)]
    if random.random() < factory.probability_recombination:
        sample = random.sample(positive_dataset, 3)
        child1, child2 = factory.get_children(parent1, parent2, sample[0], sample)
        pairs = so.pair_parents_and_children(parent1, parent2, child1, child2)
    else:
        child1 = factory.clone_organism(parent1)
        child2 = factory.clone_organism(parent2)
        child1.mutate(factory)
        child2.mutate(factory)
        pairs = [(parent1, child1), (parent2, child2)]
    child1.check_size(factory)
    child2.check_size(factory)
    return pairs


def run(so, evaluator=None, times: GenerationTimes | None = None, placer=None):
    """Batched equivalent of search_organisms.main(); `so` is the reference's module after set_up().

    evaluator(organisms, pos, neg, method, gamma) -> [n, 5] overrides the single-GPU engine (multi-GPU:
    flemingo_b200.distributed.MultiDeviceEvaluator); placer(parents, dna_set, placement_cls) -> placements[parent][seq]
    overrides place_parents (the CPU test of this loop plugs the reference's own methods into both)."""
    from objects.organism_factory import OrganismFactory
    from objects.placement_object import PlacementObject

    times = times if times is not None else GenerationTimes()
    so.single_print("Loading parameters...")
    positive_dataset = so.read_fasta_file(so.DATASET_BASE_PATH_DIR + so.POSITIVE_FILENAME)
    if so.NEGATIVE_FILENAME is not None:
        negative_dataset = so.read_fasta_file(so.DATASET_BASE_PATH_DIR + so.NEGATIVE_FILENAME)
    else:
        print("Generateing negative set...")
        negative_dataset = so.generate_negative_set(positive_dataset, so.GENERATED_NEG_SET_SIZE, so.GENERATED_NEG_SET_KMER_LEN)
    so.single_print("Instantiating population...")
    lengths = [len(s) for s in positive_dataset + negative_dataset]
    so.single_print("min_seq_length =", min(lengths))
    so.single_print("max_seq_length =", max(lengths))
    factory = OrganismFactory(so.configOrganism, so.configOrganismFactory, so.configConnector, so.configPssm, so.rank,
                              so.configShape, min(lengths), max(lengths))
    population = so.initialize_population(factory)
    generation, max_score = 0, float("-inf")
    best_org, best_org_fitness = None, -np.inf
    so.single_print("Starting execution...")

    while not so.is_finished(so.END_WHILE_METHOD, generation, max_score):
        random.shuffle(population)
        reshuffled = generation % so.PERIODIC_DATASETS_SHUFFLE == 0
        if reshuffled:
            if so.RANDOM_SHUFFLE_SAMPLING_POS:
                positive_dataset = so.shuffle_dataset(positive_dataset)
            if so.RANDOM_SHUFFLE_SAMPLING_NEG:
                negative_dataset = so.shuffle_dataset(negative_dataset)
        max_score = float("-inf")
        initial = time.time()
        pos_fit = positive_dataset[:so.MAX_SEQUENCES_TO_FIT_POS]
        neg_fit = negative_dataset[:so.MAX_SEQUENCES_TO_FIT_NEG]
        n_pairs = len(population) // 2

        # ---- breed ------------------------------------------------------------------------------------------------
        t0 = time.perf_counter()
        if reshuffled and (so.RANDOM_SHUFFLE_SAMPLING_POS or so.RANDOM_SHUFFLE_SAMPLING_NEG):
            for org in population[:2 * n_pairs]:
                org.fitness = None                      # the datasets behind the cached fitness changed (:172-175)
        placements_of = None
        if (generation + 1) % factory.periodic_mle == 0:
            parents = population[:2 * n_pairs]
            placed = (placer or place_parents)(parents, pos_fit, PlacementObject)
            placements_of = {id(p): rows for p, rows in zip(parents, placed)}
        contests = []                                   # (population slot, parent, child) in the reference's order
        for pair in range(n_pairs):
            i = 2 * pair
            for j, (parent, child) in enumerate(breed_pair(so, factory, population[i], population[i + 1], generation,
                                                           positive_dataset, placements_of)):
                contests.append((i + j, parent, child))
        t1 = time.perf_counter()

        # ---- evaluate: parents without a fitness, and every child (:239-247) --------------------------------------------
        todo = [parent for _, parent, _ in contests if parent.fitness is None] + [child for _, _, child in contests]
        times.evaluated += evaluate_fitness(todo, pos_fit, neg_fit, so.FITNESS_FUNCTION, so.GAMMA, evaluator)
        t2 = time.perf_counter()

        # ---- select (:250-262) -------------------------------------------------------------------------------------------
        pop_fitness_list, pop_n_recogs_list = [], []
        for slot, parent, child in contests:
            winner = parent if parent.fitness > child.fitness else child
            population[slot] = winner
            pop_fitness_list.append(winner.fitness)
            pop_n_recogs_list.append(winner.count_recognizers())
        t3 = time.perf_counter()

        # ---- report (:280-360): statistics, the two printed placements, periodic exports ---------------------------------
        mean_fitness = np.mean(pop_fitness_list)
        standard_dev_fitness = np.std(pop_fitness_list)
        # the reference's gini_RSV (:1095-1142) is a double loop over the population -- 16.8 M iterations at 4096
        # organisms, more than a generation's placements take on the GPU; same statistic in O(N log N)
        gini_fitness = gini_RSV(pop_fitness_list)
        mean_n_rec = np.mean(pop_n_recogs_list)
        idx_of_max_org = np.argmax(pop_fitness_list)
        max_org = population[idx_of_max_org]
        max_org_fitness = pop_fitness_list[idx_of_max_org]
        if max_org_fitness >= best_org_fitness:
            best_org, best_org_fitness = max_org, max_org_fitness
        _m, _s = divmod((time.time() - initial), 60)
        _h, _m = divmod(_m, 60)
        s_time = "{}h:{}m:{:.2f}s".format(int(_h), int(_m), _s)
        so.print_ln("\n" + "_" * max(lengths), filepath=None, to_stdout=so.GEN_INFO_TO_STDOUT)
        so.print_ln(("Generation {} | AF:{:.2f} SDF:{:.2f} GF:{:.2f} A#R:{:.2f} | MO: {} MF: {:.2f} M#R: {} | BO: {} BF: {:.2f} B#R: {}"
                     " | Time: {}").format(generation, mean_fitness, standard_dev_fitness, gini_fitness, mean_n_rec, max_org._id,
                                           max_org_fitness, max_org.count_recognizers(), best_org._id, best_org_fitness,
                                           best_org.count_recognizers(), s_time),
                    so.RESULT_BASE_PATH_DIR + so.OUTPUT_FILENAME, to_stdout=so.GEN_INFO_TO_STDOUT)
        max_org.get_placement(random.choice(positive_dataset)).print_placement(stdout=so.GEN_INFO_TO_STDOUT)
        max_org.get_placement(random.choice(negative_dataset)).print_placement(stdout=so.GEN_INFO_TO_STDOUT)
        pos_set_for_export = sorted(positive_dataset) if so.RANDOM_SHUFFLE_SAMPLING_POS else positive_dataset
        if generation % so.PERIODIC_ORG_EXPORT == 0:
            filename = "{}_{}".format(so.time.strftime("%Y-%m-%d--%H-%M-%S"), max_org._id)
            so.export_organism(max_org, pos_set_for_export, filename, factory)
        if generation % so.PERIODIC_POP_EXPORT == 0:
            seq_idx = random.randint(0, len(pos_set_for_export) - 1)
            so.export_population(population, pos_set_for_export, factory, generation, seq_idx)
            so.export_plots()
        t4 = time.perf_counter()
        times.breed += t1 - t0
        times.evaluate += t2 - t1
        times.select += t3 - t2
        times.report += t4 - t3
        times.generations += 1
        generation += 1
    return population, times
