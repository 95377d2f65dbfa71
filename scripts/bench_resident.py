#!/usr/bin/env python
"""Small 2-D slices (the z-lat sections of invert_Eliassen, 37 x 73): sweep rate of the resident engine
(one CTA per slice, operands in shared memory) against the colour engine (3-6 launches per sweep) and,
for the 5-point stencil, the fused engine.   python scripts/bench_resident.py"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xinvert_b200 as xb
from xinvert_b200 import solvers
from tests import cases

SWEEPS = 2000
for ny, nx in ((37, 73), (61, 91)):
    for with_B in (True, False):
        for batch in (1, 148, 2048):
            c = cases.random_std2d(ny, nx, with_B=with_B, seed=1, batch=batch)
            p = c["p"]
            for engine in ("resident", "colour") + (() if with_B else ("fused",)):
                row = {"grid": [ny, nx], "stencil": "9-point" if with_B else "5-point", "slices": batch, "engine": engine}
                try:
                    for rep in range(2):
                        S = c["S0"].copy()
                        fl, st = solvers.solve_standard_2D(S, c["A"], c["B"], c["C"], c["F"], "fixed", "fixed", p["del1Sqr"],
                                                           p["ratioQtr"], p["ratioSqr"], 1.2, cases.UNDEF, (0.0, 1.0, 0.0), SWEEPS - 1, -1.0,
                                                           engine=engine)
                    row.update(engine_ran=st["engine"], sweep_loop_ms=st["solve_ms"], us_per_sweep=st["solve_ms"] * 1e3 / SWEEPS,
                               cell_updates_per_s=st["cell_updates"] / (st["solve_ms"] * 1e-3), launches=st["kernel_launches"])
                except xb.XinvError as e:
                    row["error"] = str(e)
                print(json.dumps(row), flush=True)
