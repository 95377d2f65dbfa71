"""CPU oracle for the SOR hot path -- TEST INFRASTRUCTURE, not product code.

Python face of ``oracle/sor_oracle.c`` (a plain-C restatement of
``/root/reference/xinvert/numbas.py``).  The functions below carry the exact
positional signatures of the reference's numba kernels
(``numbas.py:216-219``, ``:988-991``, ``:16-19``) plus one trailing keyword,
``ordering``:

* ``'lexicographic'`` -- the reference's own in-place Gauss-Seidel order;
* ``'colour'``        -- red-black / 4-colour order around the same per-cell
  expressions (the ordering-matched oracle).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this package.  ``xinvert_b200`` never does.
"""
import ctypes as C

import numpy as np

from .build import build as _build

_BC = {"fixed": 0, "extend": 1, "periodic": 2}
_ORD = {"lexicographic": 0, "lex": 0, "colour": 1, "color": 1, "redblack": 1,
        "chebyshev": 2}      # colour ordering + the Chebyshev schedule of the relaxation factor (xinv.h, xinv_opts.accel)

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build())
        dp, i64, dbl, ci = C.c_void_p, C.c_longlong, C.c_double, C.c_int
        L.xo_invert_standard_2D.argtypes = [dp] * 5 + [i64, i64, ci, ci] + [dbl] * 5 + [dp, i64, dbl, ci]
        L.xo_invert_standard_2D.restype = None
        L.xo_invert_general_2D.argtypes = [dp] * 8 + [i64, i64, ci, ci] + [dbl] * 7 + [dp, i64, dbl, ci]
        L.xo_invert_general_2D.restype = None
        L.xo_invert_standard_3D.argtypes = [dp] * 5 + [i64, i64, i64, ci, ci, ci] + [dbl] * 5 + [dp, i64, dbl, ci]
        L.xo_invert_standard_3D.restype = None
        L.xo_invert_standard_2D_test.argtypes = [dp] * 7 + [i64, i64, ci, ci] + [dbl] * 5 + [dp, i64, dbl, ci]
        L.xo_invert_standard_2D_test.restype = None
        L.xo_invert_general_3D.argtypes = [dp] * 9 + [i64, i64, i64, ci, ci, ci] + [dbl] * 8 + [dp, i64, dbl, ci]
        L.xo_invert_general_3D.restype = None
        L.xo_invert_standard_1D.argtypes = [dp] * 4 + [i64, ci] + [dbl] * 3 + [dp, i64, dbl, ci]
        L.xo_invert_standard_1D.restype = None
        L.xo_invert_general_bih_2D.argtypes = [dp] * 11 + [i64, i64, ci, ci] + [dbl] * 9 + [dp, i64, dbl, ci]
        L.xo_invert_general_bih_2D.restype = None
        L.xo_colour_sweep_std2d.argtypes = [dp] * 5 + [i64, i64, ci] + [dbl] * 5 + [ci]
        L.xo_colour_sweep_std2d.restype = None
        L.xo_colour_of.argtypes = [ci, ci, i64, i64, i64]
        L.xo_colour_of.restype = ci
        L.xo_num_colours.argtypes = [ci, ci, i64]
        L.xo_num_colours.restype = ci
        L.xo_abs_norm.argtypes = [dp, i64, dbl]
        L.xo_abs_norm.restype = dbl
        _lib = L
    return _lib


def _p(a, shape=None, allow_none=False):
    if a is None:
        if allow_none:
            return None
        raise ValueError("array required")
    if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
        raise ValueError("oracle needs C-contiguous float64 arrays")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"shape {a.shape} != {shape}")
    return a.ctypes.data


def invert_standard_2D(S, A, B, C_, F, yc, xc, dely, delx, BCy, BCx, delxSqr,
                       ratioQtr, ratioSqr, optArg, undef, flags, mxLoop,
                       tolerance, ordering="lexicographic"):
    """numbas.invert_standard_2D (numbas.py:215-416).  ``B=None`` means B==0."""
    sh = (yc, xc)
    lib().xo_invert_standard_2D(
        _p(S, sh), _p(A, sh), _p(B, sh, True), _p(C_, sh), _p(F, sh),
        yc, xc, _BC[BCy], _BC[BCx], delxSqr, ratioQtr, ratioSqr, optArg, undef,
        _p(flags, (3,)), int(mxLoop), float(tolerance), _ORD[ordering])
    return S


def invert_general_2D(S, A, B, C_, D, E, F, G, yc, xc, dely, delx, BCy, BCx,
                      delxSqr, ratio, ratioQtr, ratioSqr, optArg, undef, flags,
                      mxLoop, tolerance, ordering="lexicographic"):
    """numbas.invert_general_2D (numbas.py:987-1201).  ``B=None`` means B==0."""
    sh = (yc, xc)
    lib().xo_invert_general_2D(
        _p(S, sh), _p(A, sh), _p(B, sh, True), _p(C_, sh), _p(D, sh), _p(E, sh),
        _p(F, sh), _p(G, sh), yc, xc, _BC[BCy], _BC[BCx],
        float(delx), delxSqr, ratio, ratioQtr, ratioSqr, optArg, undef,
        _p(flags, (3,)), int(mxLoop), float(tolerance), _ORD[ordering])
    return S


def invert_standard_3D(S, A, B, C_, F, zc, yc, xc, delz, dely, delx, BCz, BCy,
                       BCx, delxSqr, ratio2Sqr, ratio1Sqr, optArg, undef, flags,
                       mxLoop, tolerance, ordering="lexicographic"):
    """numbas.invert_standard_3D (numbas.py:15-212)."""
    sh = (zc, yc, xc)
    lib().xo_invert_standard_3D(
        _p(S, sh), _p(A, sh), _p(B, sh), _p(C_, sh), _p(F, sh),
        zc, yc, xc, _BC[BCz], _BC[BCy], _BC[BCx],
        delxSqr, ratio2Sqr, ratio1Sqr, optArg, undef,
        _p(flags, (3,)), int(mxLoop), float(tolerance), _ORD[ordering])
    return S


def colour_sweep_std2d(S, A, B, C_, F, BCx, delxSqr, ratioQtr, ratioSqr, optArg,
                       undef, colour):
    yc, xc = S.shape
    lib().xo_colour_sweep_std2d(_p(S), _p(A), _p(B, None, True), _p(C_), _p(F),
                                yc, xc, _BC[BCx], delxSqr, ratioQtr, ratioSqr,
                                optArg, undef, int(colour))
    return S


def colour_map_2d(scheme, BCx, yc, xc):
    """Colour index of every cell (int array [yc, xc]) for ``scheme`` in {2, 4}."""
    L = lib()
    out = np.empty((yc, xc), dtype=np.int32)
    for j in range(yc):
        for i in range(xc):
            out[j, i] = L.xo_colour_of(scheme, _BC[BCx], xc, j, i)
    return out


def num_colours(scheme, BCx, xc):
    return lib().xo_num_colours(scheme, _BC[BCx], xc)


def abs_norm(S, undef):
    """numbas.absNorm2D/3D (numbas.py:1689-1728)."""
    S = np.ascontiguousarray(S, dtype=np.float64)
    return lib().xo_abs_norm(S.ctypes.data, S.size, undef)


def invert_standard_2D_test(S, A, B, C_, D, E, F, yc, xc, dely, delx, BCy, BCx, delxSqr,
                            ratioQtr, ratioSqr, optArg, undef, flags, mxLoop, tolerance, ordering="lexicographic"):
    """numbas.invert_standard_2D_test (numbas.py:420-629)."""
    sh = (yc, xc)
    lib().xo_invert_standard_2D_test(
        _p(S, sh), _p(A, sh), _p(B, sh), _p(C_, sh), _p(D, sh), _p(E, sh), _p(F, sh),
        yc, xc, _BC[BCy], _BC[BCx], delxSqr, ratioQtr, ratioSqr, optArg, undef,
        _p(flags, (3,)), int(mxLoop), float(tolerance), _ORD[ordering])
    return S


def invert_general_3D(S, A, B, C_, D, E, F, G, H, zc, yc, xc, delz, dely, delx, BCz, BCy, BCx, delxSqr,
                      ratio2, ratio1, ratio2Sqr, ratio1Sqr, optArg, undef, flags, mxLoop, tolerance,
                      ordering="lexicographic"):
    """numbas.invert_general_3D (numbas.py:745-984)."""
    sh = (zc, yc, xc)
    lib().xo_invert_general_3D(
        _p(S, sh), _p(A, sh), _p(B, sh), _p(C_, sh), _p(D, sh), _p(E, sh), _p(F, sh), _p(G, sh), _p(H, sh),
        zc, yc, xc, _BC[BCz], _BC[BCy], _BC[BCx], float(delx), delxSqr, ratio2, ratio1, ratio2Sqr, ratio1Sqr,
        optArg, undef, _p(flags, (3,)), int(mxLoop), float(tolerance), _ORD[ordering])
    return S


def invert_standard_1D(S, A, B, F, xc, delx, BCx, delxSqr, optArg, undef, flags, mxLoop, tolerance,
                       ordering="lexicographic"):
    """numbas.invert_standard_1D (numbas.py:632-742)."""
    sh = (xc,)
    lib().xo_invert_standard_1D(
        _p(S, sh), _p(A, sh), _p(B, sh), _p(F, sh), xc, _BC[BCx], delxSqr, optArg, undef,
        _p(flags, (3,)), int(mxLoop), float(tolerance), _ORD[ordering])
    return S


def invert_general_bih_2D(S, A, B, C_, D, E, F, G, H, I, J, yc, xc, dely, delx, BCy, BCx, delxSSr, delxTr, delxSqr,
                          ratio, ratioSSr, ratioQtr, ratioSqr, optArg, undef, flags, mxLoop, tolerance,
                          ordering="lexicographic"):
    """numbas.invert_general_bih_2D (numbas.py:1204-1586)."""
    sh = (yc, xc)
    lib().xo_invert_general_bih_2D(
        _p(S, sh), *[_p(a, sh) for a in (A, B, C_, D, E, F, G, H, I, J)], yc, xc, _BC[BCy], _BC[BCx],
        delxSSr, delxTr, delxSqr, ratio, ratioSSr, ratioQtr, ratioSqr, optArg, undef,
        _p(flags, (3,)), int(mxLoop), float(tolerance), _ORD[ordering])
    return S
