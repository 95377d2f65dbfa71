"""Seeded synthetic problems for the parity tests (ndarray level).

The coefficient recipes follow the reference's builders
(apps.py:1401-1409 lat-lon Poisson, :2245-2313 grid parameters) restated in
numpy, plus fully random variable-coefficient problems that exercise every
operand of the stencils.
"""
import numpy as np

UNDEF = -9.99e8
REARTH = 6371200.0


def params2d(ny, nx, del2, del1):
    """numpy restatement of apps.__cal_params2D (apps.py:2282-2291)."""
    ratio = del1 / del2
    eps = np.sin(np.pi / (2.0 * nx + 2.0)) ** 2 + np.sin(np.pi / (2.0 * ny + 2.0)) ** 2
    return dict(gc2=ny, gc1=nx, del2=del2, del1=del1, ratio=ratio, ratioSqr=ratio ** 2.0,
                ratioQtr=ratio / 4.0, del1Sqr=del1 ** 2.0,
                optArg=2.0 / (1.0 + np.sqrt((2.0 - eps) * eps)))


def poisson_latlon(ny, nx, land=True, noise=1e-6, seed=0, batch=None, phase=0.0):
    """lat-lon Poisson problem: A=cosH, C=1/cosG, F=zeta*cosG with an optional
    land mask (apps.py:1401-1409); zeta is the SURVEY.md 8d formula."""
    dlat, dlon = 180.0 / ny, 360.0 / nx
    lat = -90.0 + dlat / 2 + dlat * np.arange(ny)
    lon = dlon * np.arange(nx)
    lats = np.deg2rad(lat)
    cosG = np.cos(lats)
    latm = np.empty(ny)
    latm[0] = np.nan
    latm[1:] = lats[:-1]
    cosH = np.cos((lats + latm) / 2.0)
    lam = np.deg2rad(lon)[None, :]
    phi = lats[:, None]
    rng = np.random.default_rng(seed)
    shape = (ny, nx) if batch is None else (batch, ny, nx)
    ph = phase if batch is None else (2 * np.pi * np.arange(batch) / batch)[:, None, None]
    zeta = 1e-5 * np.sin(3 * lam + ph) * np.cos(phi) ** 2 * np.sin(2 * phi)
    zeta = np.broadcast_to(zeta, shape) + noise * rng.standard_normal(shape)
    zeta = np.ascontiguousarray(zeta)
    F = zeta * cosG[:, None]
    if land:
        mask = np.sin(5 * lam) * np.cos(3 * phi) > 0.6
        F[..., mask] = UNDEF
    A = np.ascontiguousarray(np.broadcast_to(cosH[:, None], (ny, nx)))
    C = np.ascontiguousarray(np.broadcast_to(1.0 / cosG[:, None], (ny, nx)))
    p = params2d(ny, nx, np.deg2rad(dlat) * REARTH, np.deg2rad(dlon) * REARTH)
    return dict(A=A, C=C, F=F, p=p, S0=np.zeros(shape))


def poisson_latlon_user(ny, nx, land=True, noise=1e-6, seed=0, phase=0.0):
    """The same problem as poisson_latlon the way a user of invert_Poisson holds it: the raw
    vorticity (NaN on land) and the lat / lon coordinates."""
    dlat, dlon = 180.0 / ny, 360.0 / nx
    lat = -90.0 + dlat / 2 + dlat * np.arange(ny)
    lon = dlon * np.arange(nx)
    lam, phi = np.deg2rad(lon)[None, :], np.deg2rad(lat)[:, None]
    rng = np.random.default_rng(seed)
    zeta = 1e-5 * np.sin(3 * lam + phase) * np.cos(phi) ** 2 * np.sin(2 * phi) + noise * rng.standard_normal((ny, nx))
    if land:
        zeta[np.sin(5 * lam) * np.cos(3 * phi) > 0.6] = np.nan
    return zeta, lat, lon


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
