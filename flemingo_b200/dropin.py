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
from .flatten import flatten_organism


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


def accelerate_builders(connector_cls=None, pssm_cls=None, shape_cls=None) -> None:
    """Replaces the per-object table-building loops of the reference's classes with the library's builders
    (flemingo_b200.builders: threaded C++ up to the logarithm, numpy's logarithm on the batch).  The attributes they
    fill keep their types and hold bit-identical doubles:

        ConnectorObject.set_precomputed_pdfs_cdfs   connector_object.py:175-206   stored_pdfs / stored_cdfs (lists)
        PssmObject.recalculate_pssm                 pssm_object.py:368-400        pssm (array of {c,t,g,a} dicts)
        ShapeObject.set_alt_model                   shape_object.py:156-180       alt_model (list), _mu clamped
    """
    import numpy as np

    from . import builders

    if connector_cls is not None:
        def set_precomputed_pdfs_cdfs(self) -> None:
            pdfs, cdfs = builders.connector_tables([self._mu], [self._sigma], self.max_seq_length, self.pseudo_count, n_threads=1)
            self.stored_pdfs = builders.TableList.of(pdfs[0])    # lists, as in the reference; deep-copied in one step
            self.stored_cdfs = builders.TableList.of(cdfs[0])

        connector_cls.set_precomputed_pdfs_cdfs = set_precomputed_pdfs_cdfs

    if pssm_cls is not None:
        def recalculate_pssm(self) -> None:
            vals = builders.pssm_values(np.array([[col["c"], col["t"], col["g"], col["a"]] for col in self.pwm], dtype=np.float64),
                                        self.pseudo_count)
            self.pssm = np.array([{"c": float(r[0]), "t": float(r[1]), "g": float(r[2]), "a": float(r[3])} for r in vals])
            # the same values in flatten()'s order (a, g, c, t per column), valid as long as self.pssm is this object
            self._pssm_flat = np.ascontiguousarray(vals[:, [3, 2, 0, 1]]).ravel()
            self._pssm_src = self.pssm

        pssm_cls.recalculate_pssm = recalculate_pssm

    if shape_cls is not None:
        def set_alt_model(self):
            if self._mu < self.min_mu:
                self._mu = self.min_mu
            elif self._mu > self.max_mu:
                self._mu = self.max_mu
            self.alt_model = builders.TableList.of(builders.shape_alt_models([self._mu], [self._sigma], [self.edges], self.pseudo_count)[0])

        shape_cls.set_alt_model = set_alt_model


def accelerate_reference() -> None:
    """install() + accelerate() + accelerate_builders() on the reference's own modules, which must be importable as
    `objects.*` (the reference is run from its src/ directory: search_organisms.py:24)."""
    install()
    from objects.connector_object import ConnectorObject
    from objects.organism_object import OrganismObject
    from objects.pssm_object import PssmObject
    from objects.shape_object import ShapeObject
    accelerate(OrganismObject)
    accelerate_builders(ConnectorObject, PssmObject, ShapeObject)
    OrganismObject.flatten = flatten_organism
