"""Host-side mirror of the reference's domain objects, restricted to the placement / fitness path.

Same class names, constructor signatures, attribute names and return values as
/root/reference/src/objects/{organism,pssm,shape,connector,placement}_object.py, so that code (and tests)
written against the reference reads the same here.  What is mirrored is what feeds or consumes a placement:

    PssmObject.recalculate_pssm            pssm_object.py:368-400
    ShapeObject.set_alt_model / get_type   shape_object.py:156-180, :407-418
    ConnectorObject.set_precomputed_pdfs_cdfs / adjust_scores   connector_object.py:175-206, :454-487
    OrganismObject.flatten / get_placement / get_binding_energies / get_M_and_SE_on_DNA_set / set_fitness
                                           organism_object.py:1362-1504, :1263-1357, :958-966, :848-956
    PlacementObject                        placement_object.py:9-52

The GA operators (mutation, recombination, size checks, export) are host bookkeeping outside the hot path and
are not mirrored; use the reference's own classes with flemingo_b200.dropin.install() for those.

Everything that computes a placement or a fitness statistic goes to the GPU through flemingo_b200.engine.
"""
from __future__ import annotations

import math

import numpy as np

from . import builders
from . import multiplacement as _multiplacement
from .flatten import flatten_organism
from .engine import get_engine
from .population import rescale_connector_block

# Shape null models, {type: {length: {"frequencies": array, "bins": array}}} -- same global the reference
# keeps in shape_object.null_models after OrganismFactory.__init__ (organism_factory.py:80-82).
null_models: dict = {}


def set_null_models(models: dict) -> None:
    global null_models
    null_models = models


def norm_cdf(x, mu, sigma):
    """connector_object.py:18-30 / shape_object.py:8-19."""
    if sigma == 0:
        if x < mu:
            return 0
        elif x > mu:
            return 1
        return 0.5
    z = (x - mu) / abs(sigma)
    return (1.0 + math.erf(z / math.sqrt(2.0))) / 2.0


class PlacementObject:
    """Result record of one placement (placement_object.py:9-52)."""

    def __init__(self, organism_id, dna_sequence):
        self.organism_id = organism_id
        self.dna_sequence = dna_sequence
        self.energy = None
        self.recognizers_scores = []
        self.connectors_scores = []
        self.recognizers_positions = []
        self.connectors_positions = []
        self.recognizer_types = ""

    def set_energy(self, energy):
        self.energy = energy

    def set_recognizer_scores(self, recog_scores):
        self.recognizers_scores = recog_scores

    def set_connectors_scores(self, conn_scores):
        self.connectors_scores = conn_scores

    def set_recognizer_types(self, types):
        self.recognizer_types = types

    def append_recognizer_position(self, recog_position):
        self.recognizers_positions.append(recog_position)

    def append_connector_position(self, connector_position):
        self.connectors_positions.append(connector_position)


class PssmObject:
    """PSSM recognizer: pwm (array of {a,g,c,t} dicts) -> pssm of round(log2(4p + pseudo), 2)."""

    def __init__(self, pwm, config: dict) -> None:
        self.length = len(pwm)
        self.pwm = pwm
        self.pssm = None
        self.type = 'pssm'
        self.pseudo_count = config["PSEUDO_COUNT"]
        self.recalculate_pssm()

    def recalculate_pssm(self) -> None:
        # pssm_object.py:368-400: float(np.log2(4 p + pc)) rounded to 2 decimals with Python's round, per base
        vals = builders.pssm_values(np.array([[column[b] for b in ("c", "t", "g", "a")] for column in self.pwm], dtype=np.float64),
                                    self.pseudo_count)
        self.pssm = np.array([{"c": float(r[0]), "t": float(r[1]), "g": float(r[2]), "a": float(r[3])} for r in vals])
        self._pssm_flat = np.ascontiguousarray(vals[:, [3, 2, 0, 1]]).ravel()   # flatten()'s order: a, g, c, t per column
        self._pssm_src = self.pssm

    def get_type(self):
        return 'p'

    def is_connector(self) -> bool:
        return False

    def is_pssm(self) -> bool:
        return True

    def is_shape(self) -> bool:
        return False


class ShapeObject:
    """Shape recognizer (mgw / prot / roll / helt): binned null model + Gaussian alternative model."""

    _TYPE_CHAR = {"mgw": 'm', "prot": 't', "helt": 'h', "roll": 'r'}

    def __init__(self, rec_type, rec_size, config, mu=None, sigma=None):
        self.type = rec_type
        self.length = rec_size
        self.null_model = []
        self.edges = []
        self.min_mu = None
        self.max_mu = None
        self.set_null_model()
        self._mu = mu
        self._sigma = sigma
        if mu is None:
            self._mu = np.random.uniform(self.min_mu, self.max_mu)                      # shape_object.py:116
        if sigma is None:
            self._sigma = np.random.default_rng().exponential((self.edges[-1] - self.edges[0]) / 12)  # :134-138
        self.pseudo_count = config["PSEUDO_COUNT"]
        self.num_bins = config["NUM_BINS"]
        self.alt_model = []
        self.set_alt_model()

    def set_null_model(self):
        model = null_models[self.type][self.length]
        self.null_model = model["frequencies"]
        self.edges = model["bins"]
        self.min_mu = self.edges[0]
        self.max_mu = self.edges[-1]

    def set_alt_model(self):
        # shape_object.py:156-180 (natural logs; the interval mass uses the module's norm_pf quirk for sigma == 0)
        if self._mu < self.min_mu:
            self._mu = self.min_mu
        elif self._mu > self.max_mu:
            self._mu = self.max_mu
        self.alt_model = builders.TableList.of(builders.shape_alt_models([self._mu], [self._sigma], [self.edges], self.pseudo_count)[0])

    def get_type(self):
        return self._TYPE_CHAR.get(self.type)

    def is_connector(self) -> bool:
        return False

    def is_pssm(self) -> bool:
        return False

    def is_shape(self) -> bool:
        return True


class ConnectorObject:
    """Connector: Gaussian(mu, sigma) on the gap length, with log2 pdf/cdf tables of length max_seq_length."""

    def __init__(self, config, max_seq_length, mu=None, sigma=None):
        self.avg_mu = config["AVG_MU"]
        self.max_seq_length = max_seq_length
        self.pseudo_count = config["PSEUDO_COUNT"]
        self._mu = mu
        self._sigma = sigma
        if mu is None:
            self._mu = np.random.poisson(self.avg_mu)                                    # connector_object.py:150
        if sigma is None:
            self._sigma = np.random.default_rng().exponential(self.max_seq_length / 12)  # :169-173
        self.stored_pdfs = []
        self.stored_cdfs = []
        self.set_precomputed_pdfs_cdfs()

    def set_mu(self, _mu) -> None:
        self._mu = _mu

    def set_sigma(self, sigma) -> None:
        self._sigma = sigma

    def set_precomputed_pdfs_cdfs(self) -> None:
        # connector_object.py:175-206: pdf[d] = log2(Phi(d+.5) - Phi(d-.5) + pc),
        #                              cdf[d] = log2(Phi(d-.5) - Phi(-.5) + pc*(d+1))
        # Phi point by point with libm erf in the library (fl_build_connector_tables), log2 by numpy: bit-identical
        # to the reference's scalar loop (tests/test_builders_cpu.py, golden flat arrays)
        pdfs, cdfs = builders.connector_tables([self._mu], [self._sigma], self.max_seq_length, self.pseudo_count, n_threads=1)
        self.stored_pdfs = builders.TableList.of(pdfs[0])
        self.stored_cdfs = builders.TableList.of(cdfs[0])

    def adjust_scores(self, c_scores, sequence_length, sum_recognizer_lengths):
        rescale_connector_block(c_scores, self._mu, self._sigma, self.pseudo_count, self.max_seq_length,
                                sequence_length, sum_recognizer_lengths)

    def is_connector(self) -> bool:
        return True


class OrganismObject:
    """A chain of recognizers joined by connectors; placement and fitness run on the GPU."""

    def __init__(self, _id: str, conf: dict) -> None:
        self._id = _id
        self.recognizers = []
        self.connectors = []
        self.recognizers_flat = []
        self.connectors_scores_flat = []
        self.recognizer_lengths = []
        self.recognizer_types = ""
        self.recognizer_models = []
        self.recognizer_bin_edges = []
        self.recognizer_bin_nums = []
        self.sum_recognizer_lengths = 0
        self.is_precomputed = conf["PRECOMPUTE"]
        self.min_num_of_recognizers = conf.get("MIN_NUM_OF_RECOGNIZERS")
        self.max_num_of_recognizers = conf.get("MAX_NUM_OF_RECOGNIZERS")
        self.fitness = None

    # ---- flattening (organism_object.py:1362-1504) ----------------------------------------------------
    def set_sum_recognizer_lengths(self) -> None:
        self.sum_recognizer_lengths = sum([i.length for i in self.recognizers])

    def flatten(self) -> None:
        flatten_organism(self)

    def count_recognizers(self) -> int:
        return len(self.recognizers)

    def count_connectors(self) -> int:
        return len(self.connectors)

    # ---- single placement (organism_object.py:1263-1357) ------------------------------------------------
    def get_placement(self, sequence: str) -> PlacementObject:
        if self.sum_recognizer_lengths > len(sequence):
            raise ValueError(("This shouldn't have happened. The check_size " +
                              "function was supposed to prevent organisms from " +
                              "getting 'too big'. Something is not working."))
        n = len(self.recognizers)
        max_length = -1
        gaps = np.empty(n, dtype=np.dtype('i'))
        gap_scores = np.empty(n - 1, dtype=np.dtype('d'))
        recognizer_scores = np.empty(n + 1, dtype=np.dtype('d'))
        c_scores = np.copy(self.connectors_scores_flat)
        if n > 1:
            max_length = self.connectors[0].max_seq_length - 1
            offset = 2
            for con in self.connectors:
                if float(len(sequence)) + 0.5 < con._mu:
                    idx = len(sequence) - self.sum_recognizer_lengths
                    if con.stored_pdfs[idx] <= np.log(1E-15):
                        con.adjust_scores(c_scores[offset:], len(sequence), self.sum_recognizer_lengths)
                offset += con.max_seq_length * 2 + 2
        _multiplacement.calculate(bytes(sequence, "ASCII"), bytes(self.recognizer_types, "ASCII"),
                                  self.recognizers_flat, self.recognizer_lengths, c_scores, recognizer_scores,
                                  gap_scores, gaps, max_length, self.recognizer_models, self.recognizer_bin_edges,
                                  self.recognizer_bin_nums)
        return build_placement(self._id, sequence, self.recognizer_types, self.recognizer_lengths,
                               float(recognizer_scores[-1]), recognizer_scores[:-1], gap_scores, gaps)

    # ---- energies and fitness (organism_object.py:848-966) ----------------------------------------------
    def get_binding_energies(self, dna_set):
        """List of binding energies on every sequence of dna_set -- one batched device call."""
        if len(dna_set) == 0:
            return []
        eng = get_engine()
        eng.upload_population([self])
        sset = eng.upload_sequences(dna_set, 0)
        return [float(e) for e in eng.place(sset).energies[0]]

    def get_M_and_SE_on_DNA_set(self, dna_set, method, gamma):
        """Mean and standard error of the energies on dna_set for the chosen fitness variant.

        Computed by the device reduction (fl_fitness) on the pair (dna_set, dna_set); the statistics of a
        set do not depend on the other set.
        """
        eng = get_engine()
        eng.upload_population([self])
        sset = eng.upload_sequences(dna_set, 0)
        eng.place(sset, fetch=False)
        stats = eng.fitness(sset, sset, method, gamma)[0]
        return stats[1], stats[2]

    def set_fitness(self, pos_set, neg_set, method, gamma=0.2):
        stats = get_engine().evaluate([self], pos_set, neg_set, method, gamma)[0]
        self.fitness = stats[0]


def build_placement(org_id, sequence, rec_types, rec_lengths, energy, rec_scores, con_scores, gaps, placement_cls=None):
    """Fills a PlacementObject from the kernel outputs exactly as organism_object.py:1337-1355 does.
    placement_cls: the class to instantiate (the reference's own PlacementObject when its driver is running)."""
    n = len(rec_types)
    placement = (placement_cls or PlacementObject)(org_id, sequence)
    placement.set_energy(float(energy))
    placement.set_recognizer_scores([float(s) for s in rec_scores[:n]])
    placement.set_connectors_scores([float(s) for s in con_scores[:n - 1]])
    placement.set_recognizer_types(rec_types)
    current = gaps[0]
    for i in range(n):
        stop = current + rec_lengths[i]
        placement.append_recognizer_position([int(current), int(stop)])
        current += rec_lengths[i]
        if i < n - 1:
            stop = current + gaps[i + 1]
            placement.append_connector_position([int(current), int(stop)])
            current += gaps[i + 1]
    return placement


def place_population(organisms, dna_set, engine=None):
    """Batched `[[org.get_placement(seq) for seq in dna_set] for org in organisms]` in one device call."""
    eng = engine or get_engine()
    packed = eng.upload_population(organisms)
    sset = eng.upload_sequences(dna_set, 0)
    res = eng.place(sset, trace=True)
    out = []
    for o, org in enumerate(organisms):
        row = []
        for s, seq in enumerate(dna_set):
            row.append(build_placement(getattr(org, "_id", str(o)), seq, org.recognizer_types, org.recognizer_lengths,
                                       res.energies[o, s], res.rec_scores[o, s], res.con_scores[o, s], res.gaps[o, s]))
        out.append(row)
    return out


def set_population_fitness(organisms, pos_set, neg_set, method, gamma=0.2, engine=None, reuse_sets=True) -> np.ndarray:
    """Batched `for org in organisms: org.set_fitness(pos_set, neg_set, method, gamma)`; returns the
    [n_org, 5] statistics (fitness, M_p, SE_p, M_n, SE_n)."""
    eng = engine or get_engine()
    stats = eng.evaluate(organisms, pos_set, neg_set, method, gamma, reuse_sets=reuse_sets)
    for org, row in zip(organisms, stats):
        org.fitness = row[0]
    return stats
