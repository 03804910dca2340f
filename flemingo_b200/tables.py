"""Data tables of the placement path (reference: extensions/multiplacement/src/_constants.h:7-11).

`data/tables.bin` is produced by tools/extract_tables.py from the reference header; layout (float64, LE):
MGW_SCORES[1024] PROT_SCORES[1024] ROLL_SCORES[2048] HELT_SCORES[2048] DEN_EXPANSION[10000].
"""
from __future__ import annotations

from functools import lru_cache
from pathlib import Path

import numpy as np

N_TABLE_DOUBLES = 16144
_OFFSETS = {"mgw": (0, 1024), "prot": (1024, 1024), "roll": (2048, 2048), "helt": (4096, 2048), "den": (6144, 10000)}
TABLES_PATH = Path(__file__).resolve().parent / "data" / "tables.bin"


@lru_cache(maxsize=1)
def load_tables() -> np.ndarray:
    """The whole blob as one read-only float64 array (what fl_ctx_create uploads)."""
    arr = np.fromfile(TABLES_PATH, dtype="<f8")
    if arr.size != N_TABLE_DOUBLES:
        raise RuntimeError(f"{TABLES_PATH} holds {arr.size} doubles, expected {N_TABLE_DOUBLES}")
    arr.setflags(write=False)
    return arr


def table(name: str) -> np.ndarray:
    """One table by name: 'mgw', 'prot', 'roll', 'helt' (pentamer shape scores) or 'den' (DEN_EXPANSION)."""
    start, size = _OFFSETS[name]
    return load_tables()[start:start + size]
