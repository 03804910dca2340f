#!/usr/bin/env python3
"""Runs the REFERENCE's own GA driver (src/search_organisms.py, unmodified copy under oracle/_ref/src) for a few
generations, seeded, either on its stock C extension or over the GPU path -- TEST INFRASTRUCTURE (BASELINE config 1).

    python oracle/run_search_organisms.py --mode stock|dropin|accelerated|batched --workdir DIR --out result.json
                                          [--pop 16] [--gens 3] [--seed 11] [--mle 0] [--recombination 0.05]

modes
  stock        the reference as it is: its Python objects on its own `_multiplacement` extension (oracle/_ref/*.so)
  dropin       flemingo_b200.dropin.install(): `import _multiplacement` resolves to the GPU shim, every get_placement
               is one fl_calculate; nothing else changes
  accelerated  install() + accelerate(): get_binding_energies / get_M_and_SE_on_DNA_set / set_fitness of the
               reference's OrganismObject become one batched device call each, table builders run in the library
  batched      accelerated + the driver's generation loop replaced by flemingo_b200.ga.run (all children first, ONE
               device call per generation, competitions afterwards)
  batched-cpu  flemingo_b200.ga.run on the STOCK reference (its own C extension and fitness code plugged in as the
               evaluator): checks, without a GPU, that the three-phase loop reproduces the sequential one

What makes two runs comparable: `random` and `numpy.random` are seeded, `numpy.random.default_rng()` (called without a
seed by connector_object.py:169-173 and shape_object.py:134-138) is replaced by a seeded sequence of generators, and
the driver's `time.strftime` (used only in file names) counts calls instead of reading the clock.  The result file
holds the lines of output.txt without their "Time:" field and a sha256 per exported file.

The reference imports Bio, matplotlib and Markov_DNA, none of which exist in this image: oracle/shims/ stands in for
the handful of calls it makes (see oracle/shims/README.md).
"""
import argparse
import hashlib
import json
import os
import random
import re
import shutil
import sys
import time as _time
from pathlib import Path

HERE = Path(__file__).resolve().parent
REPO = HERE.parent


class _CountingTime:
    """`time` as search_organisms sees it: strftime numbers its calls (file names), the clock stays real."""

    def __init__(self):
        self.calls = 0

    def strftime(self, fmt, *args):
        self.calls += 1
        return f"t{self.calls:06d}"

    def time(self):
        return _time.time()

    def __getattr__(self, name):
        return getattr(_time, name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", required=True, choices=["stock", "dropin", "accelerated", "batched", "batched-cpu"])
    ap.add_argument("--workdir", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--pop", type=int, default=16)
    ap.add_argument("--gens", type=int, default=3)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--mle", type=int, default=0, help="PERIODIC_MLE (0: keep the config's 100, i.e. no MLE generation in a short run)")
    ap.add_argument("--recombination", type=float, default=0.05)
    ap.add_argument("--fitness", default="Welch")
    ap.add_argument("--max-pos", type=int, default=100)
    ap.add_argument("--max-neg", type=int, default=100)
    ap.add_argument("--synthetic", default=None, metavar="N_POS,N_NEG,LEN",
                    help="replace the bundled FASTA files by random-nucleotide datasets of this shape (seeds 1 and 2)")
    ap.add_argument("--devices", type=int, default=1, help="batched mode: GPUs driven by this one process (MultiDeviceEvaluator)")
    ap.add_argument("--no-exports", action="store_true", help="no periodic organism / population exports (long runs)")
    ap.add_argument("--max-recognizers", type=int, default=0, help="organism.MAX_NUM_OF_RECOGNIZERS (0: keep the config's 5)")
    ap.add_argument("--models", default=None, help="a models/models pickle written by an earlier run (skips the ~10 s null-model generation)")
    ap.add_argument("--save-models", default=None, help="copy this run's models/models pickle here")
    args = ap.parse_args()

    src_master = HERE / "_ref" / "src"
    if not (src_master / "search_organisms.py").exists():
        sys.exit("oracle/_ref/src is missing: run `make -C oracle refsrc` where /root/reference exists")
    args.out = str(Path(args.out).resolve())
    if args.models:
        args.models = str(Path(args.models).resolve())
    if args.save_models:
        args.save_models = str(Path(args.save_models).resolve())
    work = Path(args.workdir).resolve()
    if work.exists():
        shutil.rmtree(work)
    shutil.copytree(src_master, work / "src")
    src = work / "src"

    config = json.load(open(src / "config.json"))
    config["main"].update({
        "RUN_MODE": "serial", "POPULATION_SIZE": args.pop,
        "POSITIVE_FILENAME": "positive_dataset_7_100.fas", "NEGATIVE_FILENAME": "positive_dataset_7_100pc5.fas",
        "MIN_ITERATIONS": args.gens, "END_WHILE_METHOD": "iterations", "PERIODIC_ORG_EXPORT": 1, "PERIODIC_POP_EXPORT": 1,
        "GEN_INFO_TO_STDOUT": False, "FITNESS_FUNCTION": args.fitness,
        "MAX_SEQUENCES_TO_FIT_POS": args.max_pos, "MAX_SEQUENCES_TO_FIT_NEG": args.max_neg})
    if args.synthetic:
        n_pos, n_neg, seq_len = (int(x) for x in args.synthetic.split(","))
        import numpy as _np
        for name, n, seed in (("synthetic_pos.fas", n_pos, 1), ("synthetic_neg.fas", n_neg, 2)):
            rng = _np.random.default_rng(seed)
            letters = _np.frombuffer(b"ACGT", dtype=_np.uint8)
            with open(src / "datasets" / name, "w") as fh:
                for k, row in enumerate(rng.integers(0, 4, size=(n, seq_len))):
                    fh.write(f">s{k}\n{letters[row].tobytes().decode()}\n")
        config["main"].update({"POSITIVE_FILENAME": "synthetic_pos.fas", "NEGATIVE_FILENAME": "synthetic_neg.fas"})
    if args.no_exports:
        config["main"].update({"PERIODIC_ORG_EXPORT": 10 ** 9, "PERIODIC_POP_EXPORT": 10 ** 9})
    if args.max_recognizers:
        config["organism"]["MAX_NUM_OF_RECOGNIZERS"] = args.max_recognizers
    config["organismFactory"]["PROBABILITY_RECOMBINATION"] = args.recombination
    if args.mle:
        config["organismFactory"]["PERIODIC_MLE"] = args.mle
    json.dump(config, open(src / "config.json", "w"), indent=2)

    if args.models and Path(args.models).exists():
        shutil.copy(args.models, src / "models" / "models")
    sys.path.insert(0, str(HERE / "shims"))
    sys.path.insert(0, str(src))
    if str(REPO) not in sys.path:
        sys.path.insert(0, str(REPO))
    if args.mode in ("stock", "batched-cpu"):
        sys.path.insert(0, str(HERE / "_ref"))           # the compiled reference `_multiplacement`
    else:
        from flemingo_b200 import dropin
        dropin.install()                                  # `import _multiplacement` -> the GPU shim
    os.chdir(src)                                         # config.json, datasets/, models/models are relative paths

    import numpy as np
    random.seed(args.seed)
    np.random.seed(args.seed)
    counter = {"n": 0}
    real_default_rng = np.random.default_rng

    def seeded_default_rng(seed=None):
        if seed is not None:
            return real_default_rng(seed)
        counter["n"] += 1
        return real_default_rng([args.seed, counter["n"]])

    np.random.default_rng = seeded_default_rng

    import search_organisms as so
    so.time = _CountingTime()
    if args.mode in ("accelerated", "batched"):
        from flemingo_b200 import dropin
        dropin.accelerate_reference()
    if args.mode in ("dropin", "accelerated", "batched"):
        from flemingo_b200.engine import get_engine
        get_engine()          # CUDA context + module load (seconds on a fresh box) outside the timed run
    t0 = _time.time()
    so.set_up()
    ga_times = None
    if args.mode == "batched-cpu":
        # the batched generation loop on the STOCK reference (CPU): its three phases must reproduce the sequential loop
        import numpy as _np
        from flemingo_b200 import ga

        def cpu_evaluator(organisms, pos, neg, method, gamma):
            rows = _np.zeros((len(organisms), 5))
            for k, org in enumerate(organisms):
                org.set_fitness(pos, neg, method, gamma)
                rows[k, 0] = org.fitness
            return rows

        def cpu_placer(parents, dna_set, placement_cls):
            return [[p.get_placement(seq) for seq in dna_set] for p in parents]

        ga_times = ga.run(so, evaluator=cpu_evaluator, placer=cpu_placer)[1].as_dict()
    elif args.mode == "batched":
        from flemingo_b200 import ga
        evaluator = None
        if args.devices > 1:
            from flemingo_b200.distributed import MultiDeviceEvaluator
            evaluator = MultiDeviceEvaluator(list(range(args.devices)))
        ga_times = ga.run(so, evaluator=evaluator)[1].as_dict()
    else:
        so.main()
    wall = _time.time() - t0

    if args.save_models:
        shutil.copy(src / "models" / "models", args.save_models)
    result_dirs = sorted((src / "results").iterdir())
    assert len(result_dirs) == 1, result_dirs
    rdir = result_dirs[0]
    lines = [re.sub(r" \| Time: .*$", "", ln.rstrip("\n")) for ln in open(rdir / "output.txt")]
    files = {}
    for path in sorted(rdir.rglob("*")):
        if path.is_file() and path.name not in ("output.txt", "parameters.txt", "parameters.json"):
            files[str(path.relative_to(rdir))] = hashlib.sha256(path.read_bytes()).hexdigest()
    summary = {"mode": args.mode, "wall_s": wall, "output": lines, "files": files, "default_rng_calls": counter["n"]}
    if ga_times is not None:
        summary["generation_times"] = ga_times
    if args.mode not in ("stock", "batched-cpu"):
        from flemingo_b200.engine import get_engine
        summary["gpu_launches"] = get_engine().launch_count if args.devices <= 1 or args.mode != "batched" else \
            sum(e.launch_count for e in evaluator.engines) + get_engine().launch_count
    json.dump(summary, open(args.out, "w"), indent=1)
    print(f"{args.mode}: {len(lines)} generations, {len(files)} exported files, {wall:.1f} s")


if __name__ == "__main__":
    main()
