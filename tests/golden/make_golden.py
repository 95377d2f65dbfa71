#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference kernels.

    python tests/golden/make_golden.py

Runs ONLY in the authoring container: it imports
``/root/reference/xinvert/numbas.py`` by file path (``oracle/ref_loader.py``)
and needs numba.  For every case it stores the seeded inputs, the scalar
parameters, and what the reference's ``invert_standard_2D`` /
``invert_general_2D`` / ``invert_standard_3D`` return (field + flags).  The
fixtures travel to the GPU box; the reference does not.  They pin

* the C oracle in lexicographic order (``tests/test_oracle_golden.py``, CPU), and
* the CUDA path in XINV_ORDER_LEX (``tests/test_gpu_golden.py``, GPU),

bit for bit.  A second family ("bridge") stores a red-black trajectory produced
by the reference's OWN arithmetic: alternating one-sweep reference calls
(``mxLoop=0``) with the other colour's forcing masked by ``undef``
(SURVEY.md 8c), which pins the colour ordering of oracle and CUDA path.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tests import cases  # noqa: E402

UNDEF = cases.UNDEF
BCS = [("fixed", "fixed"), ("fixed", "periodic"), ("extend", "fixed"), ("extend", "periodic")]


def _zeros_like_B(c):
    """The reference always needs a B array; B == 0 is the 5-point case."""
    return np.zeros_like(c["A"]) if c.get("B") is None else c["B"]


def ref_std2d(ref, c, bcy, bcx, mxLoop, tol, omega):
    p = c["p"]
    S = c["S0"].copy()
    fl = np.array([0.0, 1.0, 0.0])
    ref.invert_standard_2D(S, c["A"], _zeros_like_B(c), c["C"], c["F"], p["gc2"], p["gc1"], p["del2"], p["del1"],
                           bcy, bcx, p["del1Sqr"], p["ratioQtr"], p["ratioSqr"], omega, UNDEF, fl, mxLoop, tol)
    return S, fl


def ref_gen2d(ref, c, bcy, bcx, mxLoop, tol, omega):
    p = c["p"]
    S = c["S0"].copy()
    fl = np.array([0.0, 1.0, 0.0])
    ref.invert_general_2D(S, c["A"], _zeros_like_B(c), c["C"], c["D"], c["E"], c["F"], c["G"], p["gc2"], p["gc1"],
                          p["del2"], p["del1"], bcy, bcx, p["del1Sqr"], p["ratio"], p["ratioQtr"], p["ratioSqr"],
                          omega, UNDEF, fl, mxLoop, tol)
    return S, fl


def ref_std3d(ref, c, bcy, bcx, mxLoop, tol, omega):
    p = c["p"]
    S = c["S0"].copy()
    fl = np.array([0.0, 1.0, 0.0])
    ref.invert_standard_3D(S, c["A"], c["B"], c["C"], c["F"], p["gc3"], p["gc2"], p["gc1"], p["del3"], p["del2"],
                           p["del1"], "fixed", bcy, bcx, p["del1Sqr"], p["ratio2Sqr"], p["ratio1Sqr"],
                           omega, UNDEF, fl, mxLoop, tol)
    return S, fl


def bridge_redblack_std2d(ref, c, bcx, iters, omega):
    """Red-black iterations done by the reference's own code: per colour, mask
    the other colour's forcing with undef and run ONE reference sweep
    (mxLoop=0).  BCy is 'fixed' here (the extend copy would run per half-sweep)."""
    p = c["p"]
    ny, nx = c["F"].shape
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    S = c["S0"].copy()
    B = _zeros_like_B(c)
    for _ in range(iters):
        for colour in (0, 1):
            Fm = c["F"].copy()
            Fm[((ii + jj) & 1) != colour] = UNDEF
            fl = np.array([0.0, 1.0, 0.0])
            ref.invert_standard_2D(S, c["A"], B, c["C"], Fm, ny, nx, p["del2"], p["del1"], "fixed", bcx,
                                   p["del1Sqr"], p["ratioQtr"], p["ratioSqr"], omega, UNDEF, fl, 0, -1.0)
    return S


def bridge_redblack_gen2d(ref, c, bcx, iters, omega):
    p = c["p"]
    ny, nx = c["G"].shape
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    S = c["S0"].copy()
    B = _zeros_like_B(c)
    for _ in range(iters):
        for colour in (0, 1):
            Gm = c["G"].copy()
            Gm[((ii + jj) & 1) != colour] = UNDEF
            fl = np.array([0.0, 1.0, 0.0])
            ref.invert_general_2D(S, c["A"], B, c["C"], c["D"], c["E"], c["F"], Gm, ny, nx, p["del2"], p["del1"],
                                  "fixed", bcx, p["del1Sqr"], p["ratio"], p["ratioQtr"], p["ratioSqr"], omega,
                                  UNDEF, fl, 0, -1.0)
    return S


def bridge_redblack_std3d(ref, c, bcx, iters, omega):
    p = c["p"]
    nz, ny, nx = c["F"].shape
    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    S = c["S0"].copy()
    for _ in range(iters):
        for colour in (0, 1):
            Fm = c["F"].copy()
            Fm[((ii + jj + kk) & 1) != colour] = UNDEF
            fl = np.array([0.0, 1.0, 0.0])
            ref.invert_standard_3D(S, c["A"], c["B"], c["C"], Fm, nz, ny, nx, p["del3"], p["del2"], p["del1"],
                                   "fixed", "fixed", bcx, p["del1Sqr"], p["ratio2Sqr"], p["ratio1Sqr"], omega,
                                   UNDEF, fl, 0, -1.0)
    return S


def bridge_fourcolour_std2d(ref, c, bcx, iters, omega):
    """9-point stencil (B != 0): four masks, colour = 2*(j&1) + (i&1)."""
    p = c["p"]
    ny, nx = c["F"].shape
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    S = c["S0"].copy()
    for _ in range(iters):
        for colour in range(4):
            Fm = c["F"].copy()
            Fm[(2 * (jj & 1) + (ii & 1)) != colour] = UNDEF
            fl = np.array([0.0, 1.0, 0.0])
            ref.invert_standard_2D(S, c["A"], c["B"], c["C"], Fm, ny, nx, p["del2"], p["del1"], "fixed", bcx,
                                   p["del1Sqr"], p["ratioQtr"], p["ratioSqr"], omega, UNDEF, fl, 0, -1.0)
    return S


def extend_rows_numpy(P, bcx):
    """The y-'extend' copy of numbas.py:284-310 / :87-115 on one 2-D level, in numpy: row 0 takes row 1 and
    row -1 takes row -2 wherever the source is not undef; without periodic-x the corner cells copy their
    diagonal neighbour."""
    ny, nx = P.shape
    cols = np.arange(nx) if bcx == "periodic" else np.arange(1, nx - 1)
    r1, rm2 = P[1].copy(), P[-2].copy()
    for dst, src in ((0, r1), (ny - 1, rm2)):
        ok = src[cols] != UNDEF
        P[dst, cols[ok]] = src[cols[ok]]
        if bcx != "periodic":
            if src[1] != UNDEF:
                P[dst, 0] = src[1]
            if src[-2] != UNDEF:
                P[dst, -1] = src[-2]


def bridge_redblack_std2d_extend(ref, c, bcx, iters, omega):
    """BCy = 'extend' through the bridge: the reference's row copy runs at every CALL, i.e. it would run
    before each half-sweep; here it is applied once per iteration in numpy (its own statement, no arithmetic)
    and the reference is called with BCy = 'fixed' (last paragraph)."""
    p = c["p"]
    ny, nx = c["F"].shape
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    S = c["S0"].copy()
    B = _zeros_like_B(c)
    for _ in range(iters):
        extend_rows_numpy(S, bcx)
        for colour in (0, 1):
            Fm = c["F"].copy()
            Fm[((ii + jj) & 1) != colour] = UNDEF
            fl = np.array([0.0, 1.0, 0.0])
            ref.invert_standard_2D(S, c["A"], B, c["C"], Fm, ny, nx, p["del2"], p["del1"], "fixed", bcx,
                                   p["del1Sqr"], p["ratioQtr"], p["ratioSqr"], omega, UNDEF, fl, 0, -1.0)
    return S


def bridge_redblack_std3d_extend(ref, c, bcx, iters, omega):
    p = c["p"]
    nz, ny, nx = c["F"].shape
    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    S = c["S0"].copy()
    for _ in range(iters):
        for k in range(1, nz - 1):                    # numbas.py:87-115: levels 1..zc-2 only
            extend_rows_numpy(S[k], bcx)
        for colour in (0, 1):
            Fm = c["F"].copy()
            Fm[((ii + jj + kk) & 1) != colour] = UNDEF
            fl = np.array([0.0, 1.0, 0.0])
            ref.invert_standard_3D(S, c["A"], c["B"], c["C"], Fm, nz, ny, nx, p["del3"], p["del2"], p["del1"],
                                   "fixed", "fixed", bcx, p["del1Sqr"], p["ratio2Sqr"], p["ratio1Sqr"], omega,
                                   UNDEF, fl, 0, -1.0)
    return S


def _pack(c):
    d = {k: v for k, v in c.items() if isinstance(v, np.ndarray)}
    for k, v in c["p"].items():
        d["p_" + k] = np.float64(v)
    return d


def main():
    if not ref_loader.available():
        raise SystemExit("needs /root/reference and numba (authoring container only)")
    ref = ref_loader.ref_numbas()
    out = {}

    # ---- lexicographic: the reference's own trajectory --------------------
    for with_B in (False, True):
        c = cases.random_std2d(24, 36, with_B=with_B, seed=101 + with_B)
        tag = f"std2d_B{int(with_B)}"
        out[tag] = _pack(c)
        for bcy, bcx in BCS:
            for sweeps in (0, 9):
                S, fl = ref_std2d(ref, c, bcy, bcx, sweeps, -1.0, 1.4 if not with_B else 1.2)
                out[tag][f"S_{bcy}_{bcx}_{sweeps}"] = S
                out[tag][f"fl_{bcy}_{bcx}_{sweeps}"] = fl
    for with_B in (False, True):
        c = cases.random_gen2d(24, 36, with_B=with_B, seed=111 + with_B)
        tag = f"gen2d_B{int(with_B)}"
        out[tag] = _pack(c)
        for bcy, bcx in BCS:
            S, fl = ref_gen2d(ref, c, bcy, bcx, 9, -1.0, 1.3)
            out[tag][f"S_{bcy}_{bcx}_9"] = S
            out[tag][f"fl_{bcy}_{bcx}_9"] = fl
    c = cases.random_std3d(6, 12, 16, seed=121)
    out["std3d"] = _pack(c)
    for bcy, bcx in BCS:
        S, fl = ref_std3d(ref, c, bcy, bcx, 7, -1.0, 1.3)
        out["std3d"][f"S_{bcy}_{bcx}_7"] = S
        out["std3d"][f"fl_{bcy}_{bcx}_7"] = fl

    # ---- to tolerance: loop counts and flags of a real solve ---------------
    c = cases.poisson_latlon(36, 72, land=True, noise=1e-6, seed=7)
    out["poisson_tol"] = _pack(c)
    for bcy, bcx in (("fixed", "periodic"), ("extend", "periodic")):
        S, fl = ref_std2d(ref, c, bcy, bcx, 5000, 1e-8, 1.4)
        out["poisson_tol"][f"S_{bcy}_{bcx}"] = S
        out["poisson_tol"][f"fl_{bcy}_{bcx}"] = fl

    # ---- overflow: omega far outside (0, 2) --------------------------------
    c = cases.random_std2d(20, 28, with_B=False, seed=131)
    out["overflow"] = _pack(c)
    S, fl = ref_std2d(ref, c, "fixed", "fixed", 5000, 1e-12, 7.0)
    out["overflow"]["S"] = S
    out["overflow"]["fl"] = fl

    # ---- bridge: colour orderings in the reference's own arithmetic --------
    c = cases.random_std2d(22, 30, with_B=False, seed=141)
    out["bridge_std2d"] = _pack(c)
    for bcx in ("fixed", "periodic"):
        out["bridge_std2d"][f"S_{bcx}"] = bridge_redblack_std2d(ref, c, bcx, 6, 1.4)
    c = cases.random_gen2d(22, 30, with_B=False, seed=142)
    out["bridge_gen2d"] = _pack(c)
    for bcx in ("fixed", "periodic"):
        out["bridge_gen2d"][f"S_{bcx}"] = bridge_redblack_gen2d(ref, c, bcx, 6, 1.3)
    c = cases.random_std3d(6, 10, 14, seed=143)
    out["bridge_std3d"] = _pack(c)
    for bcx in ("fixed", "periodic"):
        out["bridge_std3d"][f"S_{bcx}"] = bridge_redblack_std3d(ref, c, bcx, 5, 1.3)
    c = cases.random_std2d(22, 30, with_B=True, seed=144)
    out["bridge_std2d_9pt"] = _pack(c)
    # periodic-x 9-point: the reference's west-column quirk (numbas.py:327-328) is part
    # of the arithmetic either way, so both BCs are pinned
    for bcx in ("fixed", "periodic"):
        out["bridge_std2d_9pt"][f"S_{bcx}"] = bridge_fourcolour_std2d(ref, c, bcx, 5, 1.2)

    # ---- bridge with BCy = 'extend' (row copy applied outside the calls), grids of several strips / tiles ----
    c = cases.random_std2d(40, 70, with_B=False, seed=151)
    c["S0"][1, 5:9] = UNDEF                           # sources the copy must skip
    out["bridge_std2d_extend"] = _pack(c)
    for bcx in ("fixed", "periodic"):
        out["bridge_std2d_extend"][f"S_{bcx}"] = bridge_redblack_std2d_extend(ref, c, bcx, 6, 1.4)
    c = cases.random_std3d(6, 18, 68, seed=152)
    c["S0"][2, -2, 20:24] = UNDEF
    out["bridge_std3d_extend"] = _pack(c)
    for bcx in ("fixed", "periodic"):
        out["bridge_std3d_extend"][f"S_{bcx}"] = bridge_redblack_std3d_extend(ref, c, bcx, 5, 1.3)

    # ---- SURVEY 8f #3 kernels, lexicographic: cases.run_* call the reference module directly (same signatures) ----
    c = cases.random_std2dt(24, 36, seed=161)
    out["std2dt"] = _pack(c)
    for bcy, bcx in BCS:
        S, fl = cases.run_std2dt(ref, c, bcy, bcx, 9, -1.0, omega=1.2)
        out["std2dt"][f"S_{bcy}_{bcx}_9"] = S
        out["std2dt"][f"fl_{bcy}_{bcx}_9"] = fl
    c = cases.random_gen3d(6, 12, 16, seed=162)
    out["gen3d"] = _pack(c)
    for bcy, bcx in BCS:
        S, fl = cases.run_gen3d(ref, c, bcy, bcx, 7, -1.0, omega=1.3)
        out["gen3d"][f"S_{bcy}_{bcx}_7"] = S
        out["gen3d"][f"fl_{bcy}_{bcx}_7"] = fl
    c = cases.random_std1d(41, seed=163)
    out["std1d"] = _pack(c)
    for bcx in ("fixed", "extend", "periodic"):
        S, fl = cases.run_std1d(ref, c, bcx, 25, -1.0, omega=1.5)
        out["std1d"][f"S_{bcx}_25"] = S
        out["std1d"][f"fl_{bcx}_25"] = fl
    S, fl = cases.run_std1d(ref, c, "fixed", 5000, 1e-9, omega=1.5)
    out["std1d"]["S_fixed_tol"] = S
    out["std1d"]["fl_fixed_tol"] = fl

    c = cases.random_bih(21, 27, seed=164)
    out["bih2d"] = _pack(c)
    for bcy, bcx in BCS:
        S, fl = cases.run_bih(ref, c, bcy, bcx, 6, -1.0)
        out["bih2d"][f"S_{bcy}_{bcx}_6"] = S
        out["bih2d"][f"fl_{bcy}_{bcx}_6"] = fl

    only = set(sys.argv[1:])
    for tag, d in out.items():
        if only and tag not in only:
            continue
        path = os.path.join(HERE, tag + ".npz")
        np.savez_compressed(path, **d)
        print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(d)} arrays")


if __name__ == "__main__":
    main()
