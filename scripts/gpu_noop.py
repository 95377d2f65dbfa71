"""Cost of a pass that finds nothing to do (all slices frozen) = launch overhead of the fused engine."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xinvert_b200 as xb
from tests import cases
for shape in ((1800, 3600), (180, 360)):
    c = cases.poisson_latlon(*shape, land=True, noise=1e-6, seed=0)
    p = c["p"]; dev = torch.device("cuda", 0)
    S = torch.zeros(c["S0"].shape, dtype=torch.float64, device=dev)
    A, Cc, F = (torch.from_numpy(c[k]).to(dev) for k in ("A", "C", "F"))
    torch.cuda.synchronize()
    for mx, ce in ((1, 256), (511, 256)):
        S.zero_(); torch.cuda.synchronize()
        fl, st = xb.solve_standard_2D(S, A, None, Cc, F, "extend", "periodic", p["del1Sqr"], p["ratioQtr"], p["ratioSqr"],
                                      p["optArg"], mxLoop=mx, tolerance=-1.0, check_every=ce)
        print(shape, "mxLoop", mx, "launches", st["kernel_launches"], "solve_ms %.3f" % st["solve_ms"],
              "us per launch %.2f" % (st["solve_ms"] * 1e3 / max(1, st["sweeps_launched"])), "sweeps_launched", st["sweeps_launched"])
