#!/usr/bin/env python
"""C4-sized solves (invert_GillMatsuno 720x360 general form; one 1440x720 Poisson slice) with a fixed number of sweeps:
sweep time of the marching engine per kernel variant (XINV_FUSED_GEN_VARIANT / XINV_FUSED_RC_VARIANT)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import synthetic
import xinvert_b200 as xb
from tests import cases
SW = 1000
c = synthetic.gill_matsuno_beta(360, 720); p = c["p"]
for _ in range(2):
    S = c["S0"].copy()
    _, st = xb.solve_general_2D(S, c["A"], None, c["C"], c["D"], c["E"], c["F"], c["G"], "fixed", "periodic", p["del1"], p["del1Sqr"],
                                p["ratio"], p["ratioQtr"], p["ratioSqr"], 1.4, undef=cases.UNDEF, tolerance=-1.0, mxLoop=SW - 1)
print("c4 gen 720x360", st["engine"], "us/sweep %.3f" % (st["solve_ms"] * 1e3 / SW))
c = cases.poisson_latlon(720, 1440, land=True, noise=1e-6, seed=0); p = c["p"]
for _ in range(2):
    S = c["S0"].copy()
    _, st = xb.solve_standard_2D(S, c["A"], None, c["C"], c["F"], "extend", "periodic", p["del1Sqr"], p["ratioQtr"], p["ratioSqr"], p["optArg"],
                                 undef=cases.UNDEF, tolerance=-1.0, mxLoop=SW - 1)
print("poisson 1440x720", st["engine"], "us/sweep %.3f" % (st["solve_ms"] * 1e3 / SW))
