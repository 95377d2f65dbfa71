#!/usr/bin/env python
"""A short C3-sized invert_standard_3D solve for ncu captures (ndarray level, device-resident operands
are not needed: the capture looks at the sweep kernel only).  args: [sweeps] [nz ny nx]
PROF_ROWS=1: A, B and C constant along x (the row-value kernels)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xinvert_b200 as xb  # noqa: E402
from tests import cases  # noqa: E402

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
shape = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (37, 180, 360)
c = cases.random_std3d(*shape, seed=3, land=0.0)
if os.environ.get("PROF_ROWS"):
    for k in ("A", "B", "C"):
        c[k] = np.ascontiguousarray(np.broadcast_to(c[k][..., :1], c[k].shape))
for rep in range(2):
    S, fl = cases.run_std3d(xb, c, "fixed", "periodic", sweeps - 1, -1.0)
    st = xb.default_context().stats()
    print(shape, st["engine"], "row_coeffs", st["row_coeffs"], "us/sweep", st["solve_ms"] * 1e3 / sweeps, "launches", st["kernel_launches"], flush=True)
