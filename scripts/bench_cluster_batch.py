#!/usr/bin/env python
"""Batches of small lat-lon Poisson slices (fixed sweeps): the cluster engine against the marching engine.
python scripts/bench_cluster_batch.py"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xinvert_b200 as xb
from xinvert_b200 import solvers
from tests import cases

SWEEPS = 1000
for ny, nx in ((73, 144), (180, 360)):
    for batch in (1, 4, 9, 32, 148, 1024):
        if ny * nx * batch > 40e6:
            continue
        c = cases.poisson_latlon(ny, nx, land=True, noise=1e-6, seed=0, batch=batch)
        p = c["p"]
        for engine in ("cluster", "fused"):
            for rep in range(2):
                S = c["S0"].copy()
                fl, st = solvers.solve_standard_2D(S, c["A"], None, c["C"], c["F"], "extend", "periodic", p["del1Sqr"], p["ratioQtr"],
                                                   p["ratioSqr"], 1.4, cases.UNDEF, (0.0, 1.0, 0.0), SWEEPS - 1, -1.0, engine=engine)
            print(json.dumps({"grid": [ny, nx], "slices": batch, "engine": st["engine"], "sweep_loop_ms": round(st["solve_ms"], 3),
                              "us_per_sweep": round(st["solve_ms"] * 1e3 / SWEEPS, 3),
                              "cell_updates_per_s": st["cell_updates"] / (st["solve_ms"] * 1e-3)}), flush=True)
