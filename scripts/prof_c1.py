#!/usr/bin/env python
"""A C1-sized (360 x 180 lat-lon Poisson, fixed/periodic) solve with a fixed number of sweeps, for ncu
captures and quick timings of the cluster engine.  args: [sweeps] [ny nx] [bcy]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xinvert_b200 as xb  # noqa: E402
from tests import cases  # noqa: E402

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
ny, nx = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (180, 360)
bcy = sys.argv[4] if len(sys.argv) > 4 else "fixed"
c = cases.poisson_latlon(ny, nx, land=(bcy == "extend"), noise=1e-6, seed=0)
for rep in range(3):
    S, fl = cases.run_std2d(xb, c, bcy, "periodic", sweeps - 1, -1.0, omega=1.4)
    st = xb.default_context().stats()
    print((ny, nx), bcy, st["engine"], "us/sweep %.3f" % (st["solve_ms"] * 1e3 / sweeps), "launches", st["kernel_launches"], flush=True)
