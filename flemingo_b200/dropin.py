"""Route the REFERENCE's own Python objects through the GPU path.

    import flemingo_b200.dropin as dropin
    dropin.install()                      # before `import search_organisms` / `objects.organism_object`

install() registers flemingo_b200.multiplacement under the top-level name `_multiplacement`
(organism_object.py:15 does `import _multiplacement`), so every `get_placement` runs fl_calculate.
accelerate(OrganismObject) additionally replaces the per-sequence Python loops of
`get_binding_energies`, `get_M_and_SE_on_DNA_set` and `set_fitness` (organism_object.py:848-966) with one
batched device call each; signatures and return values are unchanged.
"""
from __future__ import annotations

import sys

from . import multiplacement
from .engine import get_engine


def install() -> None:
    sys.modules["_multiplacement"] = multiplacement


def accelerate(organism_cls) -> None:
    def get_binding_energies(self, dna_set):
        if len(dna_set) == 0:
            return []
        eng = get_engine()
        eng.upload_population([self])
        return [float(e) for e in eng.place(eng.upload_sequences(dna_set, 0)).energies[0]]

    def get_M_and_SE_on_DNA_set(self, dna_set, method, gamma):
        eng = get_engine()
        eng.upload_population([self])
        sset = eng.upload_sequences(dna_set, 0)
        eng.place(sset, fetch=False)
        stats = eng.fitness(sset, sset, method, gamma)[0]
        return stats[1], stats[2]

    def set_fitness(self, pos_set, neg_set, method, gamma=0.2):
        self.fitness = get_engine().evaluate([self], pos_set, neg_set, method, gamma)[0][0]

    organism_cls.get_binding_energies = get_binding_energies
    organism_cls.get_M_and_SE_on_DNA_set = get_M_and_SE_on_DNA_set
    organism_cls.set_fitness = set_fitness
