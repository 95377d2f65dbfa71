"""Seeded synthetic problems for the parity tests (ndarray level).

The coefficient recipes follow the reference's builders
(apps.py:1401-1409 lat-lon Poisson, :2245-2313 grid parameters) restated in
numpy, plus fully random variable-coefficient problems that exercise every
operand of the stencils.
"""
import numpy as np

from synthetic import (REARTH, UNDEF, gill_matsuno_beta, omega_latlon, params2d, params3d,  # noqa: F401
                       poisson_latlon, poisson_latlon_user)


def random_std2d(ny, nx, with_B, seed, land=0.1, batch=None):
    rng = np.random.default_rng(seed)
    shape = (ny, nx) if batch is None else (batch, ny, nx)
    A = 1.0 + 0.3 * rng.random(shape)
    C = 1.0 + 0.3 * rng.random(shape)
    B = 0.1 * rng.standard_normal(shape) if with_B else None
    F = 1e-9 * rng.standard_normal(shape)
    F[rng.random(shape) < land] = UNDEF
    S0 = rng.standard_normal(shape)
    p = params2d(ny, nx, 1.1e5, 0.9e5)
    return dict(A=A, B=B, C=C, F=F, S0=S0, p=p)


def random_std2d_rowcoef(ny, nx, seed, land=0.1, batch=None, undef_rows=False):
    """Like random_std2d (B == 0) but A and C vary with y only -- the structure every
    Poisson-type problem has -- which selects the fused engine's RC kernels."""
    c = random_std2d(ny, nx, False, seed, land=land, batch=batch)
    rng = np.random.default_rng(seed + 12345)
    a = 1.0 + 0.3 * rng.random(ny)
    cc = 1.0 + 0.3 * rng.random(ny)
    if undef_rows and ny > 8:
        a[ny // 3] = UNDEF                   # a whole row of undef A: rows ny//3-1 and ny//3 are never updated
        cc[2 * ny // 3] = UNDEF
    c["A"] = np.ascontiguousarray(np.broadcast_to(a[:, None], (ny, nx)))
    c["C"] = np.ascontiguousarray(np.broadcast_to(cc[:, None], (ny, nx)))
    return c


def random_gen2d(ny, nx, with_B, seed, land=0.1, batch=None):
    rng = np.random.default_rng(seed)
    shape = (ny, nx) if batch is None else (batch, ny, nx)
    A = -(1.0 + 0.3 * rng.random(shape))
    C = -(1.0 + 0.3 * rng.random(shape))
    B = 0.1 * rng.standard_normal(shape) if with_B else None
    D = 1e-6 * rng.standard_normal(shape)
    E = 1e-6 * rng.standard_normal(shape)
    F = 1e-12 * rng.random(shape)
    G = 1e-9 * rng.standard_normal(shape)
    G[rng.random(shape) < land] = UNDEF
    S0 = rng.standard_normal(shape)
    p = params2d(ny, nx, 1.1e5, 0.9e5)
    return dict(A=A, B=B, C=C, D=D, E=E, F=F, G=G, S0=S0, p=p)


def random_gen2d_rowcoef(ny, nx, seed, land=0.1, batch=None, undef_rows=False):
    """random_gen2d (B == 0) with A, C, D, E, F varying with y only -- the structure of the
    Gill-Matsuno and Stommel problems -- which the fused engine handles."""
    c = random_gen2d(ny, nx, False, seed, land=land, batch=batch)
    rng = np.random.default_rng(seed + 54321)
    vals = dict(A=-(1.0 + 0.3 * rng.random(ny)), C=-(1.0 + 0.3 * rng.random(ny)), D=1e-6 * rng.standard_normal(ny),
                E=1e-6 * rng.standard_normal(ny), F=1e-12 * rng.random(ny))
    if undef_rows and ny > 8:
        vals["D"][ny // 3] = UNDEF
        vals["A"][ny // 2] = UNDEF
    for k, v in vals.items():
        c[k] = np.ascontiguousarray(np.broadcast_to(v[:, None], (ny, nx)))
    return c


def random_std3d(nz, ny, nx, seed, land=0.1, batch=None):
    rng = np.random.default_rng(seed)
    shape = (nz, ny, nx) if batch is None else (batch, nz, ny, nx)
    A = 1.0 + 0.3 * rng.random(shape)
    B = 1.0 + 0.3 * rng.random(shape)
    C = 1.0 + 0.3 * rng.random(shape)
    F = 1e-9 * rng.standard_normal(shape)
    F[rng.random(shape) < land] = UNDEF
    S0 = rng.standard_normal(shape)
    delz, dely, delx = 2500.0, 1.1e5, 0.9e5
    p = dict(gc3=nz, gc2=ny, gc1=nx, del3=delz, del2=dely, del1=delx, del1Sqr=delx ** 2,
             ratio2Sqr=(delx / delz) ** 2 * 1e-3, ratio1Sqr=(delx / dely) ** 2, optArg=1.3)
    return dict(A=A, B=B, C=C, F=F, S0=S0, p=p)


def run_std2d(mod, c, bcy, bcx, mxLoop, tol, omega=None, **kw):
    """Call ``mod.invert_standard_2D`` (oracle, reference or CUDA shim) on case c."""
    p = c["p"]
    S = c["S0"].copy()
    fl = np.array([0.0, 1.0, 0.0])
    mod.invert_standard_2D(S, c["A"], c.get("B"), c["C"], c["F"], p["gc2"], p["gc1"], p["del2"], p["del1"],
                           bcy, bcx, p["del1Sqr"], p["ratioQtr"], p["ratioSqr"],
                           p["optArg"] if omega is None else omega, UNDEF, fl, mxLoop, tol, **kw)
    return S, fl


def run_gen2d(mod, c, bcy, bcx, mxLoop, tol, omega=None, **kw):
    p = c["p"]
    S = c["S0"].copy()
    fl = np.array([0.0, 1.0, 0.0])
    mod.invert_general_2D(S, c["A"], c.get("B"), c["C"], c["D"], c["E"], c["F"], c["G"], p["gc2"], p["gc1"],
                          p["del2"], p["del1"], bcy, bcx, p["del1Sqr"], p["ratio"], p["ratioQtr"],
                          p["ratioSqr"], p["optArg"] if omega is None else omega, UNDEF, fl, mxLoop, tol, **kw)
    return S, fl


def run_std3d(mod, c, bcy, bcx, mxLoop, tol, omega=None, **kw):
    p = c["p"]
    S = c["S0"].copy()
    fl = np.array([0.0, 1.0, 0.0])
    mod.invert_standard_3D(S, c["A"], c["B"], c["C"], c["F"], p["gc3"], p["gc2"], p["gc1"], p["del3"],
                           p["del2"], p["del1"], "fixed", bcy, bcx, p["del1Sqr"], p["ratio2Sqr"],
                           p["ratio1Sqr"], p["optArg"] if omega is None else omega, UNDEF, fl, mxLoop, tol, **kw)
    return S, fl


# ---- the remaining kernels ---------------------------------------------------------------------------
def random_std2dt(ny, nx, seed, land=0.1):
    """invert_standard_2D_test (numbas.py:420-629): A, D positive, B, C cross terms, E a small linear term."""
    rng = np.random.default_rng(seed)
    shape = (ny, nx)
    c = dict(A=1.0 + 0.3 * rng.random(shape), B=0.1 * rng.standard_normal(shape), C=0.1 * rng.standard_normal(shape),
             D=1.0 + 0.3 * rng.random(shape), E=-1e-11 * rng.random(shape), F=1e-9 * rng.standard_normal(shape),
             S0=rng.standard_normal(shape), p=params2d(ny, nx, 1.1e5, 0.9e5))
    c["F"][rng.random(shape) < land] = UNDEF
    c["B"][rng.random(shape) < 0.02] = UNDEF
    return c


def random_gen3d(nz, ny, nx, seed, land=0.1):
    """invert_general_3D (numbas.py:745-984)."""
    rng = np.random.default_rng(seed)
    shape = (nz, ny, nx)
    c = {k: -(1.0 + 0.3 * rng.random(shape)) for k in "ABC"}
    for k in "DEF":
        c[k] = 1e-6 * rng.standard_normal(shape)
    c["G"] = 1e-12 * rng.random(shape)
    c["H"] = 1e-9 * rng.standard_normal(shape)
    c["H"][rng.random(shape) < land] = UNDEF
    c["E"][rng.random(shape) < 0.02] = UNDEF
    c["S0"] = rng.standard_normal(shape)
    delz, dely, delx = 2500.0, 1.1e5, 0.9e5
    c["p"] = dict(gc3=nz, gc2=ny, gc1=nx, del3=delz, del2=dely, del1=delx, del1Sqr=delx ** 2, ratio2=delx / delz * 1e-2,
                  ratio1=delx / dely, ratio2Sqr=(delx / delz) ** 2 * 1e-3, ratio1Sqr=(delx / dely) ** 2, optArg=1.3)
    return c


def random_std1d(nx, seed, land=0.05, batch=None):
    """invert_standard_1D (numbas.py:632-742)."""
    rng = np.random.default_rng(seed)
    shape = (nx,) if batch is None else (batch, nx)
    c = dict(A=1.0 + 0.3 * rng.random(shape), B=-1e-11 * rng.random(shape), F=1e-9 * rng.standard_normal(shape),
             S0=rng.standard_normal(shape), p=dict(gc1=nx, del1=0.9e5, del1Sqr=0.9e5 ** 2, optArg=1.5))
    c["F"][rng.random(shape) < land] = UNDEF
    return c


def run_std2dt(mod, c, bcy, bcx, mxLoop, tol, omega=None, **kw):
    p = c["p"]
    S = c["S0"].copy()
    fl = np.array([0.0, 1.0, 0.0])
    mod.invert_standard_2D_test(S, c["A"], c["B"], c["C"], c["D"], c["E"], c["F"], p["gc2"], p["gc1"], p["del2"], p["del1"],
                                bcy, bcx, p["del1Sqr"], p["ratioQtr"], p["ratioSqr"], p["optArg"] if omega is None else omega,
                                UNDEF, fl, mxLoop, tol, **kw)
    return S, fl


def run_gen3d(mod, c, bcy, bcx, mxLoop, tol, omega=None, **kw):
    p = c["p"]
    S = c["S0"].copy()
    fl = np.array([0.0, 1.0, 0.0])
    mod.invert_general_3D(S, *[c[k] for k in "ABCDEFGH"], p["gc3"], p["gc2"], p["gc1"], p["del3"], p["del2"], p["del1"],
                          "fixed", bcy, bcx, p["del1Sqr"], p["ratio2"], p["ratio1"], p["ratio2Sqr"], p["ratio1Sqr"],
                          p["optArg"] if omega is None else omega, UNDEF, fl, mxLoop, tol, **kw)
    return S, fl


def run_std1d(mod, c, bcx, mxLoop, tol, omega=None, **kw):
    p = c["p"]
    S = c["S0"].copy()
    fl = np.array([0.0, 1.0, 0.0])
    mod.invert_standard_1D(S, c["A"], c["B"], c["F"], p["gc1"], p["del1"], bcx, p["del1Sqr"],
                           p["optArg"] if omega is None else omega, UNDEF, fl, mxLoop, tol, **kw)
    return S, fl


def random_bih(ny, nx, seed, land=0.08):
    """invert_general_bih_2D (numbas.py:1204-1586): A, C (4th-order terms) and B of one sign, the lower-order
    coefficients small against them at these grid spacings."""
    rng = np.random.default_rng(seed)
    shape = (ny, nx)
    dely, delx = 1.1e5, 0.9e5
    c = dict(A=1.0 + 0.3 * rng.random(shape), B=2.0 + 0.3 * rng.random(shape), C=1.0 + 0.3 * rng.random(shape),
             D=-1e-11 * rng.random(shape), E=1e-12 * rng.standard_normal(shape), F=-1e-11 * rng.random(shape),
             G=1e-17 * rng.standard_normal(shape), H=1e-17 * rng.standard_normal(shape), I=1e-22 * rng.random(shape),
             J=1e-19 * rng.standard_normal(shape), S0=rng.standard_normal(shape))
    c["J"][rng.random(shape) < land] = UNDEF
    c["E"][rng.random(shape) < 0.02] = UNDEF
    ratio = delx / dely
    c["p"] = dict(gc2=ny, gc1=nx, del2=dely, del1=delx, delxSSr=delx ** 4, delxTr=delx ** 3, del1Sqr=delx ** 2, ratio=ratio,
                  ratioSSr=ratio ** 4, ratioQtr=ratio / 4.0, ratioSqr=ratio ** 2, optArg=1.2)
    return c


def run_bih(mod, c, bcy, bcx, mxLoop, tol, omega=None, **kw):
    p = c["p"]
    S = c["S0"].copy()
    fl = np.array([0.0, 1.0, 0.0])
    mod.invert_general_bih_2D(S, *[c[k] for k in "ABCDEFGHIJ"], p["gc2"], p["gc1"], p["del2"], p["del1"], bcy, bcx,
                              p["delxSSr"], p["delxTr"], p["del1Sqr"], p["ratio"], p["ratioSSr"], p["ratioQtr"], p["ratioSqr"],
                              p["optArg"] if omega is None else omega, UNDEF, fl, mxLoop, tol, **kw)
    return S, fl
