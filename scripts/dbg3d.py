#!/usr/bin/env python
"""Diagnostics for the 3-D fused engine: a few cases against the oracle; on a mismatch, where it is."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import xinvert_b200 as xb  # noqa: E402
from tests import cases  # noqa: E402

ok = True
for shape, bcy, bcx, mx, seed in [((3, 3, 4), "fixed", "fixed", 0, 1), ((5, 9, 12), "fixed", "fixed", 0, 2),
                                  ((5, 9, 12), "fixed", "fixed", 3, 2), ((7, 12, 64), "fixed", "periodic", 0, 3),
                                  ((7, 12, 64), "fixed", "periodic", 4, 3), ((6, 29, 62), "extend", "fixed", 0, 4),
                                  ((6, 29, 62), "extend", "periodic", 3, 4), ((9, 40, 122), "extend", "periodic", 4, 5),
                                  ((37, 180, 360), "fixed", "periodic", 3, 6)]:
    c = cases.random_std3d(*shape, seed=seed)
    S_o, f_o = cases.run_std3d(oracle, c, bcy, bcx, mx, -1.0, ordering="colour")
    try:
        S_g, f_g = cases.run_std3d(xb, c, bcy, bcx, mx, -1.0, engine="fused")
    except Exception as e:
        print(shape, bcy, bcx, mx, "EXC", e)
        ok = False
        continue
    st = xb.default_context().stats()
    bad = np.argwhere(S_g != S_o)
    print(shape, bcy, bcx, "sweeps", mx + 1, "engine", st["engine"], "flags", f_g, f_o, "mismatches", len(bad), flush=True)
    if len(bad):
        ok = False
        print("  first", bad[:8].tolist())
        for ax, name in enumerate("kji"):
            print("  ", name, "values with mismatches:", sorted(set(bad[:, ax].tolist()))[:40])
        print("  max abs diff", np.abs(S_g - S_o).max(), " changed-in-oracle-but-not-gpu",
              int(((S_o != c['S0']) & (S_g == c['S0'])).sum()))
print("ALL OK" if ok else "FAILED")
