#!/usr/bin/env python3
"""bench.py -- organism x sequence placements per second (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one fitness evaluation of the whole population: every organism placed on every positive and
negative sequence (K1+K2 kernel) followed by the per-organism fitness reduction (K3).  At N=1 the workload is
BASELINE config 2 (256 organisms, 3 PSSM recognizers of length 12, 1000 + 1000 sequences of 300 bp); with N
ranks every rank owns 256 organisms (weak scaling), sequences replicated, one all-gather of the statistics.

  value  device-timed (CUDA events on the engine's stream), inputs resident in HBM
  e2e    wall-clocked through the public host API (set_population_fitness on Python organisms and str
         sequences): host packing, H2D of population + sequences, kernels, D2H of the statistics
  roofline       DP cells per second (x 2 integer ops) vs the MEASURED rate of the instruction the DP uses (VIADDMNMX,
                 tools/microbench.cu): the integer ALU pipe is the bound, not HBM and not the tensor cores
  roofline_survey_alu   SURVEY.md section 8d's line (3 flop-equivalents per cell vs SMs x 128 lanes x max clock)
  roofline_hbm   the HBM line (algorithmic bytes vs measured copy bandwidth) showing the path is not memory bound
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
SM_COUNT, LANES_PER_SM = 148, 128
DPX_PEAK_TCELLS = 18.4   # measured on this pool's B200: one VIADDMNMX per cell, 64 lanes/clk/SM (tools/microbench.cu)


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

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.samples:
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    w = workload_spec(args.workload)
    cores = os.cpu_count() or 1
    # bounded sample of the same workload: ~10-20 s of CPU work per step at ~0.6 ms per placement per core
    n_org = min(w["n_org"], max(cores, 16))
    n_seq = max(8, int(24000 * cores / 8 / n_org)) if w["seq_len"] <= 300 else max(2, int(300 * cores / 8 / n_org))
    from flemingo_b200 import synthetic
    orgs = build_organisms(args.workload, n_org, 3)
    seqs = synthetic.random_sequences(n_seq, w["seq_len"], 1)
    rates = []
    for step in range(args.warmup + args.steps):
        rate, kind, used, n, wall = cpu_reference_rate(orgs, seqs, cores)
        if step >= args.warmup:
            rates.append((rate, wall))
    value = float(np.mean([r[0] for r in rates]))
    ms = float(np.mean([r[1] for r in rates]) * 1e3)
    sample = f"{n_org} organisms x {n_seq} sequences of {w['seq_len']} bp per step ({n} placements), {used} forked workers"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": bench_config(args, w, w["n_org"] * max(1, args.gpus)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bench_config(args, w, n_org_total):
    return {"workload": f"BASELINE config {args.workload[1:]} ({args.workload}): {n_org_total} organisms x "
                        f"({w['n_pos']} pos + {w['n_neg']} neg) sequences of {w['seq_len']} bp",
            "organisms": n_org_total, "organisms_per_gpu": n_org_total // max(1, args.gpus), "sequences": w["n_pos"] + w["n_neg"], "seq_len": w["seq_len"],
            "fitness": "Welch", "offsets": False, "parallelism": f"organism-sharded x{max(1, args.gpus)}",
            "l2": "flushed between timed steps (256 MiB write)"}


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from flemingo_b200 import synthetic
    from flemingo_b200.distributed import gather_stats
    from flemingo_b200.engine import Engine
    from flemingo_b200.objects import set_population_fitness
    from flemingo_b200.population import pack_population

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    w = workload_spec(args.workload)
    pos = synthetic.random_sequences(w["n_pos"], w["seq_len"], 1)
    neg = synthetic.random_sequences(w["n_neg"], w["seq_len"], 2)
    n_seq = len(pos) + len(neg)
    if args.scaling == "strong" and world > 1:
        # the named workload, sharded by organism cost (largest first); every rank builds the same population
        from flemingo_b200.distributed import organism_costs, shard_by_cost
        everyone = build_organisms(args.workload, w["n_org"], 3)
        shards = shard_by_cost(organism_costs(everyone, w["seq_len"]), world)
        orgs = [everyone[i] for i in shards[rank]]
        n_total_org = w["n_org"]
    else:
        # every rank draws its own population (seed 3 + rank): weak scaling, organism-sharded
        orgs = build_organisms(args.workload, w["n_org"], 3 + rank)
        shards = [np.arange(r * w["n_org"], (r + 1) * w["n_org"]) for r in range(world)]
        n_total_org = w["n_org"] * world
    n_local = len(orgs)

    eng = Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.stream_handle, device=device)
    packed = pack_population(orgs)
    eng.upload_population(packed)
    sp, sn = eng.upload_sequences(pos, 0), eng.upload_sequences(neg, 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def step_resident():
        eng.place(sp, fetch=False)
        eng.place(sn, fetch=False)
        local = eng.fitness(sp, sn, "Welch", 0.2)
        return gather_stats(local, shards, rank, world, device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        stats = step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count
    dev_ms, place_ms = 0.0, 0.0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(1)   # evict L2 between timed steps
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            eng.place(sp, fetch=False)
            eng.place(sn, fetch=False)
            e1.record(stream)
            local = eng.fitness(sp, sn, "Welch", 0.2)
            e2.record(stream)
        stats = gather_stats(local, shards, rank, world, device)
        torch.cuda.synchronize()
        dev_ms += e0.elapsed_time(e2)
        place_ms += e0.elapsed_time(e1)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = eng.launch_count - launches0
    clocks = sampler.stop()

    # end to end through the public host API (Python organisms + str sequences in, statistics out).
    # e2e: EVERYTHING is re-uploaded every step (population and both DNA sets); e2e_resident: the population is
    # uploaded every step, the DNA sets -- which a GA run never changes between generations -- stay in HBM.
    # The same W warm-up calls as the device-timed loop (at least 3): the first calls after start-up run 1-2 ms slower
    # (allocator growth, page faults on fresh host arrays), tools/e2e_iterations.py.
    e2e_steps = max(1, min(args.steps, 10))
    e2e_warmup = max(3, args.warmup)

    def e2e_loop(reuse):
        for _ in range(e2e_warmup):
            set_population_fitness(orgs, pos, neg, "Welch", 0.2, engine=eng, reuse_sets=reuse)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            loc = set_population_fitness(orgs, pos, neg, "Welch", 0.2, engine=eng, reuse_sets=reuse)
            gather_stats(loc, shards, rank, world, device)
        barrier()
        return (time.perf_counter() - t0) / e2e_steps

    e2e_s = e2e_loop(False)
    e2e_res_s = e2e_loop(True)
    pop_bytes = int(sum(a.nbytes for a in (packed.rec_begin, packed.rec_type, packed.rec_len, packed.rec_nb, packed.rec_data,
                                           packed.pool_begin, packed.rec_pool, packed.con_M, packed.con_pool, packed.sum_len,
                                           packed.rec_class, packed.cls_kind, packed.cls_len, packed.cls_nb, packed.cls_edges,
                                           packed.edges_pool)) + 2 * packed.n_org * 12)
    seq_bytes = int(n_seq * ((w["seq_len"] + 15) // 16 * 4 + (w["seq_len"] + 31) // 32 * 4 + 17))
    h2d = pop_bytes + seq_bytes
    d2h = int(packed.n_org * 5 * 8)

    # max over ranks
    t = torch.tensor([dev_ms, place_ms, e2e_s, wall_ms, e2e_res_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, place_ms, e2e_s, wall_ms, e2e_res_s = (float(x) for x in t.cpu())
    if rank == 0:
        n_org_total = n_total_org
        placements = n_org_total * n_seq
        ms_per_step = dev_ms / args.steps
        value = placements / (ms_per_step * 1e-3)
        cells, adds = cells_per_placement(orgs, w["seq_len"])
        hbm_peak, sm_max_mhz, how = load_peaks()
        # dominant kernel = place_fast_kernel (the integer DPX filter pass + exact refinement; one launch per
        # sequence set and tile bucket, two per step here), per GPU
        k_ms = place_ms / args.steps / 2
        k_pairs = n_local * n_seq / 2
        flops = k_pairs * (3.0 * cells + adds)
        alu_peak = SM_COUNT * LANES_PER_SM * sm_max_mhz * 1e6 / 1e12
        achieved = flops / (k_ms * 1e-3) / 1e12
        cells_per_s = k_pairs * cells / (k_ms * 1e-3)
        traffic = load_traffic(args.workload)
        bytes_alg = k_pairs * ((w["seq_len"] + 3) // 4 + 8)
        hbm_achieved = bytes_alg / (k_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": bench_config(args, w, n_org_total), "clocks": clocks,
                "e2e": {"value": placements / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "warmup": e2e_warmup},
                "e2e_resident_sets": {"value": placements / e2e_res_s, "unit": UNIT, "h2d_bytes_per_step": pop_bytes,
                                      "d2h_bytes_per_step": d2h, "ms_per_step": e2e_res_s * 1e3,
                                      "note": "population uploaded every step; the two DNA sets stay resident in HBM"},
                "gpu_launches": int(launches),
                # primary line: the integer ALU pipe the DP runs on.  One cell = one VIADDMNMX = 2 integer ops (add, max);
                # the peak is the MEASURED rate of that instruction on this pool's B200 (64 lanes/clk/SM), x 2 ops
                "roofline": {"bound": "alu", "kernel": "place_fast_kernel", "achieved": 2.0 * cells_per_s / 1e12,
                             "peak": 2.0 * DPX_PEAK_TCELLS, "unit": "TFLOP/s (integer op equivalents: add + max per DP cell)",
                             "frac": cells_per_s / 1e12 / DPX_PEAK_TCELLS, "traffic": traffic,
                             "peak_source": "measured: int32 VIADDMNMX (DPX) rate 18.4 Tcell/s, profiles/r1_microbench_pipe_rates.txt",
                             "cells_per_s": cells_per_s, "kernel_ms": k_ms},
                # SURVEY.md section 8d's line (3 flop-equivalents per cell vs SMs x 128 lanes x clock).  DPX retires add+max
                # in one instruction, so this fraction can exceed 1; it is kept for continuity with the first bench lines
                "roofline_survey_alu": {"bound": "alu", "achieved": achieved, "peak": alu_peak,
                                        "unit": "TFLOP/s (non-tensor flop-equivalents: 3 per DP cell + 1 per score add)",
                                        "frac": achieved / alu_peak,
                                        "peak_source": f"{SM_COUNT} SMs x {LANES_PER_SM} lanes x {sm_max_mhz:.0f} MHz ({how} max SM clock)"},
                "roofline_hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_achieved / hbm_peak, "traffic": None, "peak_source": how},
                "wall_ms_per_step": wall_ms / args.steps}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_o = min(len(orgs), max(cores, 16))
            n_s = max(8, int(24000 * cores / 8 / n_o)) if w["seq_len"] <= 300 else max(2, int(200 * cores / 8 / n_o))
            rate, kind, used, n, wall = cpu_reference_rate(orgs[:n_o], (pos + neg)[:n_s], cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": used, "kind": kind,
                                    "sample": f"{n_o} organisms x {n_s} sequences ({n} placements) in {wall:.1f} s"}
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
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank owns one full population of the workload (default); strong: the named workload is "
                         "sharded over the ranks by organism cost (north_star: 2k x 10k in < 100 ms on 8 GPUs)")
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
