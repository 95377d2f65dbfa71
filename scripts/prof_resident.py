#!/usr/bin/env python
"""A 37 x 73 nine-point (invert_Eliassen-sized) solve with a fixed number of sweeps on the resident engine, for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xinvert_b200 as xb  # noqa: E402
from tests import cases  # noqa: E402
sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
c = cases.random_std2d(37, 73, with_B=True, seed=1)
for rep in range(3):
    S, fl = cases.run_std2d(xb, c, "fixed", "fixed", sweeps - 1, -1.0, omega=1.2)
    st = xb.default_context().stats()
    print(st["engine"], "us/sweep %.3f" % (st["solve_ms"] * 1e3 / sweeps), flush=True)
