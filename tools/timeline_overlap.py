"""Reads a CTA-0 phase timeline (tools/phase_clocks.py) and reports how the warps of each scheduler overlap."""
import sys, collections
import numpy as np
names = ["prologue", "plan", "unpack", "scores", "dp", "refine", "misc", "claim"]
ev = collections.defaultdict(list)
for line in open(sys.argv[1]):
    w, t, ph = line.split()
    ev[int(w)].append((int(t), int(ph)))
# keep the last launch only: split at gaps > 1e6 cycles
t_all = sorted(t for w in ev for t, _ in ev[w])
cut = 0
for a, b in zip(t_all, t_all[1:]):
    if b - a > 1_000_000:
        cut = b
iv = {}
for w in ev:
    e = [(t, ph) for t, ph in ev[w] if t >= cut]
    iv[w] = [(e[i - 1][0], e[i][0], e[i][1]) for i in range(1, len(e))]
t0 = min(a for w in iv for a, _, _ in iv[w]); t1 = max(b for w in iv for _, b, _ in iv[w])
print(f"launch span {t1 - t0} cycles, warps {len(iv)}")
grid = np.arange(t0, t1, 200)
phase_at = np.full((16, len(grid)), -1, dtype=np.int8)
for w in iv:
    for a, b, ph in iv[w]:
        phase_at[w, np.searchsorted(grid, a):np.searchsorted(grid, b)] = ph
for sched in range(4):
    ws = [w for w in iv if w % 4 == sched]
    in_dp = sum((phase_at[w] == 4).astype(int) for w in ws)
    in_ref = sum(((phase_at[w] == 5) | ((phase_at[w] >= 8) & (phase_at[w] <= 11))).astype(int) for w in ws)
    hist = np.bincount(in_dp, minlength=5) / len(grid)
    histr = np.bincount(in_ref, minlength=5) / len(grid)
    print(f"scheduler {sched}: P(#warps in dp = 0..4) = {np.round(hist, 3)}   P(#warps in refine = 0..4) = {np.round(histr, 3)}")
tot = np.bincount(sum((phase_at[w] == 4).astype(int) for w in iv), minlength=17) / len(grid)
print("whole CTA: P(#warps in dp = k):", np.round(tot, 3))
# a stretch of the timeline as text: one char per 2000 cycles
chars = "pPusDr.c1234????"
for w in sorted(iv):
    row = "".join(chars[phase_at[w, i]] if phase_at[w, i] >= 0 else " " for i in range(0, min(len(grid), 1500), 10))
    print(f"w{w:02d} {row}")
