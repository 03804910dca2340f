#!/usr/bin/env python3
"""bench.py -- organism x sequence placements per second (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one fitness evaluation of the whole population: every organism placed on every positive and
negative sequence (K1+K2 kernels) followed by the per-organism fitness reduction (K3) and, on N > 1 ranks, the one
all-gather of the [n_org, 5] statistics.  The default workload is BASELINE config 3 -- 2000 mixed PSSM + shape
organisms x 10 000 sequences of 300 bp, the configuration north_star quotes "at 1/2/4/8 GPUs" -- under STRONG
scaling: the same population for every N, sharded by organism cost.  At N = 1 the line also carries `config2`, the
same measurement on BASELINE config 2 (256 organisms x 2000 sequences).  `--scaling weak` gives every rank its own
population of the workload.

  value  device-timed (CUDA events on the engine's stream), inputs resident in HBM
  e2e    wall-clocked through the public host API (evaluate_sharded on Python organisms and str sequences): host
         packing, H2D of population + sequences, kernels, all-gather, D2H of the statistics
  roofline       DP cells per second vs the rate of the DPX instruction the DP stage is made of (VIADDMNMX), MEASURED
                 in this run on this device (fl_measure_dpx_rate): the integer ALU pipe is the bound, not HBM and not
                 the tensor cores
  roofline_survey_alu   SURVEY.md section 8d's line (3 flop-equivalents per cell vs SMs x 128 lanes x max clock)
  roofline_hbm   the HBM line (algorithmic bytes vs measured copy bandwidth) showing the path is not memory bound
  hand_back_rate share of the pairs the integer filter pass handed to the exact kernel (fl_path_stats)
  cpu_baseline   the unmodified reference C extension (oracle/_ref) on the host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "organism x sequence placements/sec"
UNIT = "placements/s"
LANES_PER_SM = 128


def load_traffic(workload: str):
    """dram__bytes_read + dram__bytes_write of one launch of the dominant kernel, from the committed ncu capture."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        return json.loads(p.read_text()).get(workload)
    return None


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured"
    return 6650.0, 1965.0, "fallback"


# ------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------
def workload_spec(name: str):
    from flemingo_b200.synthetic import WORKLOADS
    return dict(WORKLOADS[name])


def build_organisms(name: str, n_org: int, seed: int):
    from flemingo_b200 import synthetic
    w = workload_spec(name)
    if w["kind"] == "pssm":
        return synthetic.pssm_organisms(n_org, w["seq_len"], seed, mu_range=w.get("mu_range"))
    return synthetic.mixed_organisms(n_org, w["seq_len"], seed)


def cells_per_placement(orgs, seq_len: int) -> float:
    """Mean DP cells W = sum_{i>=1} P(P+1)/2 and score adds per placement."""
    cells = adds = 0.0
    for o in orgs:
        n = len(o.recognizer_types)
        P = seq_len - o.sum_recognizer_lengths + 1
        cells += (n - 1) * P * (P + 1) / 2
        adds += P * o.sum_recognizer_lengths
    return cells / len(orgs), adds / len(orgs)


# ------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi at 20 ms while the bench runs.  Started BEFORE the warm-up steps (nvidia-smi needs ~0.1-0.3 s to
    deliver its first line, longer than a short multi-GPU timed region); `stop(t0, t1)` reports the samples that fell
    inside the timed region [t0, t1] and, if the region was too short to catch any, the ones of the whole run."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def wait_first(self, timeout: float = 3.0):
        t0 = time.time()
        while self.proc and not self.samples and time.time() - t0 < timeout:
            time.sleep(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0: float | None = None, t1: float | None = None) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.06)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(rows):
            sm, mx, reasons = [], [], set()
            for _, ln in rows:
                parts = [x.strip() for x in ln.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            return sm, mx, reasons

        inside = [r for r in self.samples if t0 is not None and t0 - 0.02 <= r[0] <= t1 + 0.04]
        sm, mx, reasons = digest(inside)
        window = "timed region"
        if not sm:
            sm, mx, reasons = digest(self.samples)
            window = "warm-up + timed region (timed region shorter than the sampling period)"
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_total": len(self.samples), "window": window}


# ------------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference C extension on the host cores
# ------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    org_slice, seqs = args
    from oracle import oracle
    t0 = time.perf_counter()
    n = 0
    for org in org_slice:
        for s in seqs:
            oracle.place(org, s, _ref_worker.impl)
            n += 1
    return n, time.perf_counter() - t0


def cpu_reference_rate(orgs, seqs, cores: int):
    """placements/s of the reference path on `cores` forked workers over disjoint organism slices (the
    reference's own MPI decomposition, search_organisms.py:1145-1179).  Uses oracle/_ref when it was built,
    otherwise the C restatement ('port')."""
    import multiprocessing as mp
    from oracle import oracle
    kind = "reference" if oracle.reference_module() is not None else "port"
    _ref_worker.impl = "reference" if kind == "reference" else "restated"
    slices = [orgs[i::cores] for i in range(cores)]
    slices = [s for s in slices if s]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(slices)) as pool:
        res = pool.map(_ref_worker, [(s, seqs) for s in slices])
    wall = time.perf_counter() - t0
    n = sum(r[0] for r in res)
    return n / wall, kind, len(slices), n, wall


def reference_sample(workload: str, cores: int, per_core: int):
    """A bounded sample of the workload for the CPU arm: organisms evenly spread over the population (cost grows
    with the number of recognizers, so the first few would not be representative) x the first sequences."""
    from flemingo_b200 import synthetic
    w = workload_spec(workload)
    everyone = build_organisms(workload, w["n_org"], 3)
    n_org = min(len(everyone), max(4 * cores, 64))
    pick = np.linspace(0, len(everyone) - 1, n_org).astype(int)
    long_seqs = w["seq_len"] > 300
    n_seq = max(2 if long_seqs else 8, int(per_core * cores / n_org / (60 if long_seqs else 1)))
    seqs = synthetic.random_sequences(n_seq, w["seq_len"], 1)
    return w, [everyone[i] for i in pick], seqs


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # ~2-4 s of CPU work per step: ~1-1.5 ms per config-3 placement per core
    w, orgs, seqs = reference_sample(args.workload, cores, 2500)
    rates = []
    for step in range(args.warmup + args.steps):
        rate, kind, used, n, wall = cpu_reference_rate(orgs, seqs, cores)
        if step >= args.warmup:
            rates.append((rate, wall))
    value = float(np.mean([r[0] for r in rates]))
    ms = float(np.mean([r[1] for r in rates]) * 1e3)
    sample = (f"{len(orgs)} organisms (evenly spread over the population) x {len(seqs)} sequences of {w['seq_len']} bp per step "
              f"({n} placements), {used} forked workers over disjoint organism slices; C part = the unmodified reference extension "
              "(oracle/_ref), Python marshalling = the oracle's restatement of OrganismObject.get_placement (no PlacementObject)")
    world_eff = max(1, args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.workload, w, w["n_org"] * (world_eff if args.scaling == "weak" else 1), world_eff, args.scaling),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    if args.scaling == "weak" and world_eff > 1:
        line["scaling"] = "weak"
    print(json.dumps(line), flush=True)


def bench_config(workload: str, w, n_org_total: int, world: int, scaling: str):
    return {"workload": f"BASELINE config {workload[1:]} ({workload}): {n_org_total} organisms x "
                        f"({w['n_pos']} pos + {w['n_neg']} neg) sequences of {w['seq_len']} bp",
            "organisms": n_org_total, "organisms_per_gpu": n_org_total / max(1, world), "sequences": w["n_pos"] + w["n_neg"],
            "seq_len": w["seq_len"], "fitness": "Welch", "offsets": False,
            "parallelism": f"organism-sharded x{max(1, world)} ({scaling} scaling), sequences replicated, one all-gather of [n_org,5]",
            "l2": "flushed between timed steps (256 MiB write)"}


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def measure_workload(args, name: str, scaling: str, rank: int, world: int, device, eng, flush, peak: dict, sampler=None):
    """Device-timed and end-to-end measurement of one workload; returns the JSON fields (meaningful on rank 0)."""
    import torch
    import torch.distributed as dist
    from flemingo_b200 import synthetic
    from flemingo_b200.distributed import DeviceGather, evaluate_sharded, organism_costs, shard_by_cost
    from flemingo_b200.population import pack_population

    w = workload_spec(name)
    pos = synthetic.random_sequences(w["n_pos"], w["seq_len"], 1)
    neg = synthetic.random_sequences(w["n_neg"], w["seq_len"], 2)
    n_seq = len(pos) + len(neg)
    if scaling == "strong" or world == 1:
        # the named workload, sharded by organism cost (largest first); every rank builds the same population
        everyone = build_organisms(name, w["n_org"], 3)
        shards = shard_by_cost(organism_costs(everyone, w["seq_len"]), world)
    else:
        # weak scaling: every rank owns one full population of the workload (seed 3 + rank)
        mine_only = build_organisms(name, w["n_org"], 3 + rank)
        everyone = [None] * (w["n_org"] * world)
        everyone[rank * w["n_org"]:(rank + 1) * w["n_org"]] = mine_only
        shards = [np.arange(r * w["n_org"], (r + 1) * w["n_org"]) for r in range(world)]
    orgs = [everyone[i] for i in shards[rank]]
    n_total_org, n_local = len(everyone), len(orgs)

    stream = torch.cuda.ExternalStream(eng.stream_handle, device=device)
    packed = pack_population(orgs)
    eng.upload_population(packed)
    sp, sn = eng.upload_sequences(pos, 0), eng.upload_sequences(neg, 1)
    gather = DeviceGather(eng, shards, rank, world, device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stats = None
    for _ in range(args.warmup):
        eng.place(sp, fetch=False)
        eng.place(sn, fetch=False)
        stats = gather(sp, sn, "Welch", 0.2)
    eng.path_stats()   # reset the hand-back counters
    barrier()
    launches0 = eng.launch_count
    dev_ms, place_ms = 0.0, 0.0
    t_wall0 = time.time()
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(1)   # evict L2 between timed steps
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            eng.place(sp, fetch=False)
            eng.place(sn, fetch=False)
            e1.record(stream)
            stats = gather(sp, sn, "Welch", 0.2)     # fitness kernel, all-gather out of device memory, one D2H
            e2.record(stream)
        torch.cuda.synchronize()
        dev_ms += e0.elapsed_time(e2)
        place_ms += e0.elapsed_time(e1)
    barrier()
    t_wall1 = time.time()
    wall_ms = (t_wall1 - t_wall0) * 1e3
    launches = eng.launch_count - launches0
    fast_pairs, handed_back, second_chance = eng.path_stats(detail=True)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler is not None else None

    # sharded == unsharded: a seeded sample of organisms from every shard, recomputed on this rank alone, must
    # reproduce the gathered rows bit for bit (same kernels, other batch composition, other GPU)
    check = None
    if scaling == "strong" or world == 1:
        rng = np.random.default_rng(7)
        sample = np.sort(rng.choice(n_total_org, size=min(24, n_total_org), replace=False))
        eng.upload_population([everyone[i] for i in sample])
        eng.place(sp, fetch=False)
        eng.place(sn, fetch=False)
        local = eng.fitness(sp, sn, "Welch", 0.2)
        same = bool(np.array_equal(local, stats[sample], equal_nan=True))
        check = {"organisms": int(sample.size), "bit_identical": same}
        if not same:
            raise AssertionError(f"rank {rank}: gathered statistics differ from the local recomputation")

    # end to end through the public host API (Python organisms + str sequences in, statistics out): host packing, H2D of
    # the population and both DNA sets EVERY step, kernels, all-gather, D2H.  e2e_resident_sets keeps the DNA sets --
    # which a GA run never changes between generations -- in HBM.
    e2e_steps = max(1, min(args.steps, 10))
    e2e_warmup = max(3, min(args.warmup, 5))

    def e2e_loop(reuse):
        for _ in range(e2e_warmup):
            evaluate_sharded(eng, everyone, pos, neg, "Welch", 0.2, rank, world, device, shards=shards, reuse_sets=reuse, gather=gather)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            evaluate_sharded(eng, everyone, pos, neg, "Welch", 0.2, rank, world, device, shards=shards, reuse_sets=reuse, gather=gather)
        barrier()
        return (time.perf_counter() - t0) / e2e_steps

    e2e_s = e2e_loop(False)
    e2e_res_s = e2e_loop(True)
    pop_bytes = int(sum(a.nbytes for a in (packed.rec_begin, packed.rec_type, packed.rec_len, packed.rec_nb, packed.rec_data,
                                           packed.pool_begin, packed.rec_pool, packed.con_M, packed.con_pool, packed.sum_len,
                                           packed.rec_class, packed.cls_kind, packed.cls_len, packed.cls_nb, packed.cls_edges,
                                           packed.edges_pool)) + 2 * packed.n_org * 12)
    seq_bytes = int(n_seq * (w["seq_len"] + 5 * 4 + 8))   # raw bytes + per-sequence offsets / lengths
    d2h = int(world * gather.width * 5 * 8)

    # local work of this rank (cells), then max / sum over ranks
    cells_mean, adds_mean = cells_per_placement(orgs, w["seq_len"]) if orgs else (0.0, 0.0)
    # DP cells of the organisms the host routes to the packed 16-bit pass (PSSM-only, >= 2 recognizers; the value-range
    # test of fits_int16 practically never turns such an organism down)
    packed_orgs = [o for o in orgs if len(o.recognizer_types) >= 2 and set(o.recognizer_types) == {"p"}]
    cells16 = (cells_per_placement(packed_orgs, w["seq_len"])[0] * len(packed_orgs) * n_seq) if packed_orgs else 0.0
    t = torch.tensor([dev_ms, place_ms, e2e_s, wall_ms, e2e_res_s], dtype=torch.float64, device=device)
    sums = torch.tensor([float(fast_pairs), float(handed_back), cells_mean * n_local * n_seq, adds_mean * n_local * n_seq,
                         float(launches), float(second_chance), cells16], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms, place_ms, e2e_s, wall_ms, e2e_res_s = (float(x) for x in t.cpu())
    fast_all, fb_all, cells_all, adds_all, launches_all, retry_all, cells16_all = (float(x) for x in sums.cpu())

    placements = n_total_org * n_seq
    ms_per_step = dev_ms / args.steps
    hbm_peak, sm_max_mhz, how = load_peaks()
    # dominant kernel = the placement kernels (integer DPX filter pass + exact refinement): all launches of a step,
    # max over ranks; algorithmic cells of the whole job / that time = whole-job cell rate, per GPU = / world
    k_ms = place_ms / args.steps
    cells_per_s_gpu = cells_all / (k_ms * 1e-3) / world
    # peak: the issue rate of the DPX instruction each organism's DP stage is made of, measured in this run -- int32
    # VIADDMNMX (one cell per instruction) or VIADDMNMX.S16x2 (two) -- blended by the DP cells that go through each
    share16 = cells16_all / cells_all if cells_all else 0.0
    dpx_peak = 1.0 / ((1.0 - share16) / peak["cells_per_s_i32"] + share16 / peak["cells_per_s_i16x2"])
    alu_peak = peak["n_sm"] * LANES_PER_SM * sm_max_mhz * 1e6 / 1e12
    flop_eq = (3.0 * cells_all + adds_all) / (k_ms * 1e-3) / world / 1e12
    bytes_alg = placements * ((w["seq_len"] + 3) // 4 + 8) / world
    hbm_achieved = bytes_alg / (k_ms * 1e-3) / 1e9
    line = {"metric": METRIC, "value": placements / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if (scaling == "strong" or world == 1) and world > 1 else ("weak" if world > 1 else "strong"),
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(name, w, n_total_org, world, scaling if world > 1 else "strong"),
            "clocks": clocks,
            "e2e": {"value": placements / e2e_s, "unit": UNIT, "h2d_bytes_per_step": pop_bytes + seq_bytes, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "warmup": e2e_warmup},
            "e2e_resident_sets": {"value": placements / e2e_res_s, "unit": UNIT, "h2d_bytes_per_step": pop_bytes,
                                  "d2h_bytes_per_step": d2h, "ms_per_step": e2e_res_s * 1e3,
                                  "note": "population uploaded every step; the two DNA sets stay resident in HBM"},
            "gpu_launches": int(launches_all),
            "hand_back_rate": (fb_all / fast_all) if fast_all else 0.0,
            "hand_back": {"filter_pairs": int(fast_all), "handed_to_exact_kernel": int(fb_all),
                          "int32_second_chance_after_packed_16bit_pass": int(retry_all)},
            "sharded_equals_unsharded": check,
            # primary line: the integer ALU pipe the DP runs on.  One cell = one VIADDMNMX = 2 integer ops (add, max);
            # the peak is the rate of that instruction MEASURED in this run on this device, x 2 ops
            "roofline": {"bound": "alu", "kernel": "place_fast_kernel (all placement launches of a step)",
                         "achieved": 2.0 * cells_per_s_gpu / 1e12, "peak": 2.0 * dpx_peak / 1e12,
                         "unit": "TFLOP/s per GPU (integer op equivalents: add + max per DP cell)",
                         "frac": cells_per_s_gpu / dpx_peak, "traffic": load_traffic(name),
                         "peak_source": f"measured in this run (fl_measure_dpx_rate, {peak['n_sm']} SMs): int32 VIADDMNMX "
                                        f"{peak['cells_per_s_i32'] / 1e12:.2f} Tcell/s, packed VIADDMNMX.S16x2 "
                                        f"{peak['cells_per_s_i16x2'] / 1e12:.2f} Tcell/s; blended by the share of the DP cells on the "
                                        f"packed pass ({share16:.1%})",
                         "frac_vs_int32_dpx": cells_per_s_gpu / peak["cells_per_s_i32"],
                         "cells_per_s": cells_per_s_gpu, "kernel_ms": k_ms},
            # SURVEY.md section 8d's line (3 flop-equivalents per cell vs SMs x 128 lanes x clock).  DPX retires add+max
            # in one instruction, so this fraction can exceed 1; kept for continuity
            "roofline_survey_alu": {"bound": "alu", "achieved": flop_eq, "peak": alu_peak,
                                    "unit": "TFLOP/s per GPU (non-tensor flop-equivalents: 3 per DP cell + 1 per score add)",
                                    "frac": flop_eq / alu_peak,
                                    "peak_source": f"{peak['n_sm']} SMs x {LANES_PER_SM} lanes x {sm_max_mhz:.0f} MHz ({how} max SM clock)"},
            "roofline_hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                             "traffic": None, "peak_source": how},
            "wall_ms_per_step": wall_ms / args.steps}
    return line, orgs, pos, neg


def generation_wall(eng, orgs, pos, neg, reps: int = 3) -> dict:
    """Wall clock of ONE GA generation's fitness work when every organism is new (the worst case: every child mutated):
    rebuild every node's tables from its parameters (PSSM log-odds, connector pdf/cdf tables, shape alt models -- the
    library's builders behind the reference's object API) and flatten, pack the population, then upload + place +
    reduce.  The GA operators themselves (clone, mutate, recombine) are the reference's own Python and not included."""
    from flemingo_b200 import objects, synthetic
    from flemingo_b200.population import pack_population

    def rebuild(org):
        recs = []
        for r in org.recognizers:
            if r.get_type() == "p":
                recs.append(objects.PssmObject(r.pwm, synthetic.CONF_PSSM))
            else:
                recs.append(objects.ShapeObject(r.type, r.length, synthetic.CONF_SHAPE, r._mu, r._sigma))
        cons = [objects.ConnectorObject(synthetic.CONF_CON, c.max_seq_length, c._mu, c._sigma) for c in org.connectors]
        return synthetic.assemble(org._id, recs, cons)

    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        fresh = [rebuild(o) for o in orgs]
        t1 = time.perf_counter()
        packed = pack_population(fresh)
        t2 = time.perf_counter()
        sp, sn = eng.upload_sequences(pos, 0, reuse=True), eng.upload_sequences(neg, 1, reuse=True)
        eng.upload_population(packed)
        eng.place(sp, fetch=False)
        eng.place(sn, fetch=False)
        eng.fitness(sp, sn, "Welch", 0.2)
        t3 = time.perf_counter()
        cur = (t1 - t0, t2 - t1, t3 - t2)
        if best is None or sum(cur) < sum(best):
            best = cur
    n = len(orgs)
    return {"organisms": n, "rebuild_tables_and_flatten_ms": best[0] * 1e3, "pack_ms": best[1] * 1e3,
            "upload_place_reduce_ms": best[2] * 1e3, "host_us_per_organism": (best[0] + best[1]) / n * 1e6,
            "note": "every organism rebuilt from its parameters (worst case: all children mutated); DNA sets resident; "
                    "best of %d" % reps}


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from flemingo_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()                       # before the warm-up: see ClockSampler
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    eng = Engine(local_rank)
    peak = eng.measure_dpx_rate()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    sampler.wait_first()
    line, orgs, pos, neg = measure_workload(args, args.workload, args.scaling, rank, world, device, eng, flush, peak, sampler)
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        w, ref_orgs, ref_seqs = reference_sample(args.workload, cores, 2500)
        rate, kind, used, n, wall = cpu_reference_rate(ref_orgs, ref_seqs, cores)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": used, "kind": kind,
                                "sample": f"{len(ref_orgs)} organisms (evenly spread over the population) x {len(ref_seqs)} sequences "
                                          f"({n} placements) in {wall:.1f} s; C part = the unmodified reference extension, Python "
                                          "marshalling = the oracle's restatement of get_placement"}
    if world == 1:
        line["generation_wall"] = generation_wall(eng, orgs, pos, neg)
    if world == 1 and args.workload != "c2" and not args.no_config2:
        # BASELINE config 2 (the single-GPU configuration) on the same box, same method
        c2, _, _, _ = measure_workload(args, "c2", "strong", rank, world, device, eng, flush, peak, None)
        line["config2"] = {k: c2[k] for k in ("value", "unit", "ms_per_step", "config", "e2e", "e2e_resident_sets", "gpu_launches",
                                              "hand_back_rate", "sharded_equals_unsharded", "roofline", "roofline_survey_alu")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE config; default c3 = 2000 mixed organisms x 10 000 sequences of 300 bp (north_star: at 1/2/4/8 GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config2", action="store_true", help="skip the additional config-2 measurement at N=1")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): the named workload is sharded over the ranks by organism cost (north_star: 2k x 10k "
                         "in < 100 ms on 8 GPUs); weak: every rank owns one full population of the workload")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
