"""CPU: the tiling / folding logic of the DP kernel (dp_stage, tile_step in csrc/place_kernel.cuh) compiled for
the host and run lane by lane against the naive O(P^2) recurrence, for every tile size and awkward P."""
import re
import subprocess
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]


def test_tiled_dp_stage_equals_naive(tmp_path):
    src = (REPO / "flemingo_b200" / "csrc" / "place_kernel.cuh").read_text()
    start = src.index("struct OpF64 {")
    end = src.index("// ------------------------------------------------------------------------------------------------\n// Shared pieces of the placement kernels")
    (tmp_path / "tile_extract.h").write_text(src[start:end])
    exe = tmp_path / "test_tile"
    subprocess.run(["g++", "-O2", "-std=c++17", f'-DTILE_EXTRACT="{tmp_path / "tile_extract.h"}"', "-o", str(exe),
                    str(REPO / "tests" / "cpp" / "test_tile.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "total bad 0" in out.stdout
