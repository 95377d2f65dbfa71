#!/usr/bin/env python
"""Many small slices -- the shape of reanalysis workloads (time x level x 73 x 144): invert_Poisson over a
batch of B slices through the facade; rate and loop counts.   python scripts/bench_many_small.py [B]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xinvert_b200 as xb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ny, nx = 73, 144
lat, lon = np.linspace(-90, 90, ny), np.linspace(0, 357.5, nx)
lam, phi = np.deg2rad(lon)[None, None, :], np.deg2rad(lat)[None, :, None]
t = np.arange(B)[:, None, None]
rng = np.random.default_rng(0)
z = 1e-5 * np.sin(3 * lam + 0.01 * t) * np.cos(phi) ** 2 * np.sin(2 * phi) + 1e-6 * rng.standard_normal((B, ny, nx))
F = xb.DataArray(z, ['time', 'lat', 'lon'], {'time': np.arange(B), 'lat': lat, 'lon': lon})
for ip in ({'BCs': ['extend', 'periodic'], 'tolerance': -1.0, 'mxLoop': 999, 'printInfo': False},
           {'BCs': ['extend', 'periodic'], 'tolerance': 1e-9, 'mxLoop': 5000, 'printInfo': False}):
    xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=dict(ip))
    t0 = time.perf_counter()
    ipp = dict(ip)
    S = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=ipp)
    wall = time.perf_counter() - t0
    st = xb.default_context().stats()
    fl = ipp.get('flags_all')
    print(json.dumps({"slices": B, "grid": [ny, nx], "tolerance": ip['tolerance'], "engine": st["engine"], "rows": st["row_coeffs"],
                      "cell_updates": st["cell_updates"], "sweep_loop_ms": st["solve_ms"], "wall_ms": wall * 1e3,
                      "gpu_cell_updates_per_s": st["cell_updates"] / (st["solve_ms"] * 1e-3),
                      "api_cell_updates_per_s": st["cell_updates"] / wall,
                      "loops_min_max": [float(fl[:, 2].min()), float(fl[:, 2].max())] if fl is not None else None}))
