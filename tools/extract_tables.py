#!/usr/bin/env python3
"""Extract the reference's DATA tables into flemingo_b200/data/tables.bin.

The four pentamer shape tables and DEN_EXPANSION are data that the reference ships inside
`extensions/multiplacement/src/_constants.h:3-11`.  DEN_EXPANSION in particular is NOT exactly
round(log2(n!), 2), so it can only be reproduced by loading it.  This script parses
the initialisers as decimal literals and stores them as little-endian float64, in the order

    MGW_SCORES[1024] PROT_SCORES[1024] ROLL_SCORES[2048] HELT_SCORES[2048] DEN_EXPANSION[10000]

(the C array is declared [10000] with 9999 initialisers, so the last element is 0.0, as in C).
It cross-checks the shape tables against the reference's Python copy (src/models/constants.py).
Run once in the build container (where /root/reference exists); the .bin is committed.
"""
import re
import sys
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent.parent / "flemingo_b200" / "data" / "tables.bin"
SIZES = [("MGW_SCORES", 1024), ("PROT_SCORES", 1024), ("ROLL_SCORES", 2048), ("HELT_SCORES", 2048),
         ("DEN_EXPANSION", 10000)]


def main():
    text = (REF / "extensions/multiplacement/src/_constants.h").read_text()
    arrays = {}
    for name, size in SIZES:
        m = re.search(r"%s\[(\d+)\]\s*=\s*\{([^}]*)\}" % name, text)
        assert m and int(m.group(1)) == size, name
        vals = [float(tok) for tok in m.group(2).replace("\n", " ").split(",") if tok.strip()]
        assert len(vals) <= size, (name, len(vals))
        arr = np.zeros(size, dtype="<f8")
        arr[: len(vals)] = vals
        arrays[name] = arr
        print(f"{name}: declared {size}, initialisers {len(vals)}")
    ns = {}
    exec((REF / "src/models/constants.py").read_text(), ns)
    for name, _ in SIZES[:4]:
        assert np.array_equal(np.asarray(ns[name], dtype="<f8"), arrays[name]), name
    OUT.parent.mkdir(parents=True, exist_ok=True)
    with open(OUT, "wb") as fh:
        for name, _ in SIZES:
            fh.write(arrays[name].tobytes())
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
