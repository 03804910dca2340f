"""BASELINE config 1 end to end: the REFERENCE's own driver (`search_organisms.py`, unmodified copy under oracle/_ref/src,
serial mode, bundled FASTA files) run for a few seeded generations

  * on its stock `_multiplacement` C extension (CPU), and
  * over the GPU path -- `dropin.install()` alone, `install() + accelerate()`, and the batched generation loop
    `flemingo_b200.ga.run` --

must write the same output.txt (fitness trajectory, best organisms) and byte-identical export files (best-organism
JSON / text / placements, population JSON / text / placements).  oracle/_ref/src is produced by `make -C oracle refsrc`
where /root/reference exists and travels to the GPU box as a git-ignored directory.
"""
import json
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
RUNNER = REPO / "oracle" / "run_search_organisms.py"
HAVE_REF = (REPO / "oracle" / "_ref" / "src" / "search_organisms.py").exists() and list((REPO / "oracle" / "_ref").glob("_multiplacement*.so"))

pytestmark = pytest.mark.skipif(not HAVE_REF, reason="needs oracle/_ref (make -C oracle ref refsrc)")


def _run(mode, tmp_path, tag, extra=()):
    out = tmp_path / f"{tag}_{mode}.json"
    models = tmp_path / "models.pickle"
    cmd = [sys.executable, str(RUNNER), "--mode", mode, "--workdir", str(tmp_path / f"work_{tag}_{mode}"), "--out", str(out),
           "--models", str(models), "--save-models", str(models), *extra]
    run = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, cwd=str(REPO))
    assert run.returncode == 0, (mode, run.stdout[-2000:], run.stderr[-3000:])
    return json.loads(out.read_text())


def _assert_same_run(a, b):
    assert a["output"] == b["output"], (a["mode"], b["mode"])
    assert sorted(a["files"]) == sorted(b["files"])
    diff = [k for k in a["files"] if a["files"][k] != b["files"][k]]
    assert not diff, (a["mode"], b["mode"], diff)
    assert a["default_rng_calls"] == b["default_rng_calls"]


def test_stock_runs_are_reproducible(tmp_path):
    """The harness itself: two seeded runs of the stock reference agree (CPU only, no GPU involved)."""
    a = _run("stock", tmp_path, "a")
    b = _run("stock", tmp_path, "b")
    assert len(a["output"]) == 3 and a["output"][0].startswith("Generation 0 | AF:")
    assert any(k.endswith("_organism.json") for k in a["files"])
    _assert_same_run(a, b)


def test_batched_generation_loop_equals_sequential_loop_on_cpu(tmp_path):
    """flemingo_b200.ga.run (breed everything, evaluate everything, then select) with the reference's own CPU code as
    evaluator and placer reproduces search_organisms.main() -- MLE generations and recombination included."""
    extra = ("--gens", "4", "--pop", "12", "--mle", "2", "--recombination", "0.5", "--seed", "5")
    _assert_same_run(_run("stock", tmp_path, "cpu", extra), _run("batched-cpu", tmp_path, "cpu", extra))


@pytest.mark.gpu
@pytest.mark.parametrize("scenario,extra", [
    ("plain", ("--gens", "3", "--pop", "16")),
    ("mle_recomb", ("--gens", "4", "--pop", "12", "--mle", "2", "--recombination", "0.5", "--seed", "5")),
    ("yuen", ("--gens", "2", "--pop", "8", "--fitness", "Yuen", "--seed", "9", "--max-pos", "60", "--max-neg", "80")),
])
def test_reference_driver_over_the_gpu_path_matches_stock(tmp_path, scenario, extra):
    stock = _run("stock", tmp_path, scenario, extra)
    for mode in ("dropin", "accelerated", "batched"):
        got = _run(mode, tmp_path, scenario, extra)
        _assert_same_run(stock, got)
        assert got["gpu_launches"] > 0
