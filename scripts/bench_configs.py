#!/usr/bin/env python
"""Secondary measurements (not the driver's bench.py contract): the BASELINE.json configs that
are parity cases -- C1 (invert_Poisson 360x180), C3 (invert_omega 360x180x37), C4
(invert_GillMatsuno 720x360 beta-plane) -- through the xarray-style facade, fixed sweep counts.

    python scripts/bench_configs.py [--cpu]      one JSON line per config

GPU rate = cell-updates / device time of the sweep loop (stats.solve_ms); "api" = wall time of
the whole invert_* call (host coefficient building + H2D + solve + D2H).  --cpu adds the C port
of the reference (lexicographic, one core) through the same facade.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xinvert_b200 as xb  # noqa: E402
from xinvert_b200 import core  # noqa: E402

DA = xb.DataArray


def c1():
    ny, nx = 180, 360
    lat, lon = -89.5 + np.arange(ny), 1.0 * np.arange(nx)
    lam, phi = np.deg2rad(lon)[None, :], np.deg2rad(lat)[:, None]
    zeta = 1e-5 * np.sin(3 * lam) * np.cos(phi) ** 2 * np.sin(2 * phi)
    F = DA(zeta, ['lat', 'lon'], {'lat': lat, 'lon': lon})
    ip = {'BCs': ['fixed', 'periodic'], 'optArg': 1.4, 'tolerance': -1.0, 'mxLoop': 1999, 'printInfo': False}
    return "c1 invert_Poisson 360x180", xb.invert_Poisson, (F,), dict(dims=['lat', 'lon'], iParams=ip), ny * nx


def c3():
    nz, ny, nx = 37, 180, 360
    lev = 100000.0 - 2500.0 * np.arange(nz)
    lat, lon = -89.5 + np.arange(ny), 1.0 * np.arange(nx)
    rng = np.random.default_rng(1)
    coords = {'LEV': lev, 'lat': lat, 'lon': lon}
    N2 = DA(1e-6 * (1 + 0.5 * rng.random((nz, ny, nx))), ['LEV', 'lat', 'lon'], coords)
    F = DA(1e-17 * rng.standard_normal((nz, ny, nx)), ['LEV', 'lat', 'lon'], coords)
    ip = {'BCs': ['fixed', 'fixed', 'periodic'], 'tolerance': -1.0, 'mxLoop': 199, 'printInfo': False}
    return ("c3 invert_omega 360x180x37", xb.invert_omega, (F,),
            dict(dims=['LEV', 'lat', 'lon'], iParams=ip, mParams={'N2': N2}), nz * ny * nx)


def c4():
    ny, nx = 360, 720
    y, x = np.linspace(-5e6, 5e6, ny), np.linspace(0, 4e7, nx, endpoint=False)
    yy, xx = np.meshgrid(y, x, indexing="ij")
    Q = DA(0.05 * np.exp(-((yy / 1e6) ** 2 + ((xx - 2e7) / 2e6) ** 2)), ['y', 'x'], {'y': y, 'x': x})
    ip = {'BCs': ['fixed', 'periodic'], 'optArg': 1.4, 'tolerance': -1.0, 'mxLoop': 999, 'printInfo': False}
    mp = {'f0': 0.0, 'beta': 2e-11, 'epsilon': 1e-5, 'Phi': 5000}
    return ("c4 invert_GillMatsuno 720x360 beta-plane", xb.invert_GillMatsuno, (Q,),
            dict(dims=['y', 'x'], coords='cartesian', iParams=ip, mParams=mp), ny * nx)


def notebook11():
    """The only timing the reference publishes (docs/source/notebooks/11_Omega_equation.ipynb:525-551):
    invert_omega on a 601 x 300 x 300 grid (x, y, levels), mxLoop = 500 (501 sweeps), fixed/fixed/extend
    BCs; 4 such solves took 2920 s there (about 730 s each, I/O and coefficient building included)."""
    nz, ny, nx = 300, 300, 601
    lev = 100000.0 - 300.0 * np.arange(nz)
    lat, lon = 20.0 + 0.1 * np.arange(ny), 140.0 + 0.1 * np.arange(nx)
    rng = np.random.default_rng(11)
    coords = {'LEV': lev, 'lat': lat, 'lon': lon}
    N2 = DA(1e-5 * (1 + 0.5 * rng.random((nz, ny, nx))), ['LEV', 'lat', 'lon'], coords)
    F = DA(1e-17 * rng.standard_normal((nz, ny, nx)), ['LEV', 'lat', 'lon'], coords)
    ip = {'BCs': ['fixed', 'fixed', 'extend'], 'tolerance': -1.0, 'mxLoop': 500, 'printInfo': False}
    return ("notebook-11 invert_omega 601x300x300, 501 sweeps", xb.invert_omega, (F,),
            dict(dims=['LEV', 'lat', 'lon'], iParams=ip, mParams={'N2': N2}), nz * ny * nx)


def main():
    cpu = "--cpu" in sys.argv
    ctx = xb.default_context(0)
    for make in ((notebook11,) if "--notebook" in sys.argv else (c1, c3, c4)):
        name, fn, a, kw, N = make()
        sweeps = kw["iParams"]["mxLoop"] + 1
        fn(*a, **kw)                                   # warm-up (allocations, first launches)
        t0 = time.perf_counter()
        fn(*a, **kw)
        wall = time.perf_counter() - t0
        st = ctx.stats()
        line = {"config": name, "sweeps": sweeps, "cells": N, "engine": st["engine"], "colours": st["ncolours"],
                "gpu_cell_updates_per_s": sweeps * N / (st["solve_ms"] * 1e-3), "sweep_loop_ms": st["solve_ms"],
                "us_per_sweep": st["solve_ms"] * 1e3 / sweeps, "kernel_launches": st["kernel_launches"],
                "api_wall_ms": wall * 1e3, "api_cell_updates_per_s": sweeps * N / wall}
        if cpu:
            from tests import oracle_backend
            kw2 = dict(kw, iParams=dict(kw["iParams"], mxLoop=max(1, min(sweeps, int(2e8 / N))) - 1,
                                        ordering="lexicographic"))
            saved, core.solvers = core.solvers, oracle_backend
            try:
                t0 = time.perf_counter()
                fn(*a, **kw2)
                dt = time.perf_counter() - t0
            finally:
                core.solvers = saved
            line["cpu_port_cell_updates_per_s_api"] = (kw2["iParams"]["mxLoop"] + 1) * N / dt
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
