"""TEST INFRASTRUCTURE: a stand-in for ``xinvert_b200.solvers`` that runs the C
oracle slice by slice.  CPU tests monkeypatch it into ``xinvert_b200.core`` to
check the HOST logic of the facade (masking, coefficient builders, grid
parameters, batching, de-masking) without a GPU.  The product never imports it."""
import numpy as np

import oracle


def _slices(S, core_ndim):
    core = S.shape[-core_ndim:]
    return S.reshape((-1,) + core)


def _pick(a, b, core):
    if a is None:
        return None
    a = np.asarray(a)
    return a if a.shape == core else a.reshape((-1,) + core)[b]


def _ord(ordering, kw):
    """iParams['accel'] = 'chebyshev' -> the oracle's colour ordering with the same schedule of the relaxation factor."""
    if kw.get("accel") == "chebyshev" and ordering in ("colour", "color", "redblack"):
        return "chebyshev"
    return ordering


def _b_or_none(B):
    return None if (B is None or not np.any(B)) else B


def solve_standard_2D(S, A, B, C_, F, BCy, BCx, delxSqr, ratioQtr, ratioSqr, optArg, undef=-9.99e8,
                      flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8, ordering="colour", **kw):
    B = _b_or_none(B)
    Sv = _slices(S, 2)
    core = Sv.shape[1:]
    out = np.zeros((Sv.shape[0], 3))
    for b in range(Sv.shape[0]):
        fl = np.array(flags, dtype=np.float64).reshape(-1)[:3].copy()
        oracle.invert_standard_2D(Sv[b], _pick(A, b, core), _pick(B, b, core), _pick(C_, b, core), _pick(F, b, core),
                                  core[0], core[1], 0.0, 0.0, BCy, BCx, delxSqr, ratioQtr, ratioSqr, optArg, undef,
                                  fl, mxLoop, tolerance, ordering=_ord(ordering, kw))
        out[b] = fl
    return out, {"engine": "oracle"}


def solve_general_2D(S, A, B, C_, D, E, F, G, BCy, BCx, delx, delxSqr, ratio, ratioQtr, ratioSqr, optArg,
                     undef=-9.99e8, flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8, ordering="colour", **kw):
    B = _b_or_none(B)
    Sv = _slices(S, 2)
    core = Sv.shape[1:]
    out = np.zeros((Sv.shape[0], 3))
    for b in range(Sv.shape[0]):
        fl = np.array(flags, dtype=np.float64).reshape(-1)[:3].copy()
        oracle.invert_general_2D(Sv[b], *[_pick(x, b, core) for x in (A, B, C_, D, E, F, G)], core[0], core[1],
                                 0.0, delx, BCy, BCx, delxSqr, ratio, ratioQtr, ratioSqr, optArg, undef, fl,
                                 mxLoop, tolerance, ordering=_ord(ordering, kw))
        out[b] = fl
    return out, {"engine": "oracle"}


def solve_standard_3D(S, A, B, C_, F, BCz, BCy, BCx, delxSqr, ratio2Sqr, ratio1Sqr, optArg, undef=-9.99e8,
                      flags=(0.0, 1.0, 0.0), mxLoop=5000, tolerance=1e-8, ordering="colour", **kw):
    Sv = _slices(S, 3)
    core = Sv.shape[1:]
    out = np.zeros((Sv.shape[0], 3))
    for b in range(Sv.shape[0]):
        fl = np.array(flags, dtype=np.float64).reshape(-1)[:3].copy()
        oracle.invert_standard_3D(Sv[b], *[_pick(x, b, core) for x in (A, B, C_, F)], core[0], core[1], core[2],
                                  0.0, 0.0, 0.0, BCz, BCy, BCx, delxSqr, ratio2Sqr, ratio1Sqr, optArg, undef, fl,
                                  mxLoop, tolerance, ordering=_ord(ordering, kw))
        out[b] = fl
    return out, {"engine": "oracle"}
