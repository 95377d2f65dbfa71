"""Drop-in for ``xinvert/core.py`` on the hot path: ``inv_standard2D``,
``inv_general2D`` and ``inv_standard3D`` with the reference's positional
signatures (core.py:88, :374, :20), taking xarray-style arrays and returning
``S`` (mutated in place, like the reference).

What changes underneath: the reference walks the non-core dims (time, level,
ensemble ...) in a serial Python loop and calls one numba kernel per slice
(core.py:129-139, :418-428, :59-69).  Here all non-core dims are flattened into
the ``batch`` argument of ONE C-ABI call (include/xinv.h); every slice still
stops on its own test, so per-slice results and loop counts are those of the
serial loop.  There is no CPU implementation behind these functions.

Extra ``iParams`` keys (all optional; defaults reproduce the reference's
semantics up to the documented ordering difference):
  ``ordering``  'colour' (red-black / 4-colour, the fast path; default) or
                'lexicographic' (the reference's own update order, bit-identical
                trajectory, much slower on a GPU)
  ``engine``    'auto' | 'colour' | 'fused'
  ``ctx``       an ``xinvert_b200.Context`` (device / stream); default: device 0
  ``devices``   list of GPU indices: the flattened batch is cut into contiguous blocks
                (distributed.shard_bounds), one context and one thread per GPU from this process;
                per-slice results are those of a single GPU (slices are independent solves)
After the call ``iParams['flags']`` holds the flags of the LAST slice (what the
reference's shared flags array ends up with) and ``iParams['flags_all']`` the
``[batch, 3]`` array of every slice.
"""
import itertools

import numpy as np

from . import solvers
from .xrshim import coord_values

_undeftmp = -9.99e8


def _layout(F, dims):
    """Axis bookkeeping: non-core dims first (in order), then the core dims."""
    all_dims = list(F.dims)
    for d in dims:
        if d not in all_dims:
            raise Exception(f"dimension {d!r} not in {tuple(all_dims)}")
    noncore = [d for d in all_dims if d not in dims]
    core = [d for d in all_dims if d in dims]
    if core != list(dims):
        raise Exception(f"core dims {dims} must appear in this order in the array dims {tuple(all_dims)}")
    order = [all_dims.index(d) for d in noncore + core]
    return all_dims, noncore, order


def _as_batched(X, F_dims, order, dims, name):
    """float64 C-contiguous [noncore..., core...] copy/view of X.values.  Arrays that
    carry only the core dims are passed through as one slice shared by the batch."""
    v = np.asarray(X.values if hasattr(X, "values") else X)
    xd = list(getattr(X, "dims", F_dims))
    if xd == list(F_dims):
        v = np.transpose(v, order)
    elif xd == list(dims):
        pass                                # one slice shared by the whole batch
    else:
        raise Exception(f"{name} has dims {tuple(xd)}; expected {tuple(F_dims)} or {tuple(dims)}")
    return np.ascontiguousarray(v, dtype=np.float64)


def _slice_labels(F, noncore):
    labels = [coord_values(F, d) for d in noncore]
    if not noncore:
        return [{}]
    return [dict(zip(noncore, combo)) for combo in itertools.product(*labels)]


def _report(iParams, labels, flags):
    """One line per slice, in the reference's format (core.py:141-153)."""
    if not iParams.get("printInfo", True):
        return
    for sel, fl in zip(labels, flags):
        info = (str(sel).replace("numpy.datetime64(", "").replace("np.datetime64(", "")
                .replace("numpy.timedelta64(", "").replace("np.timedelta64(", "")
                .replace("np.float64(", "").replace("np.int64(", "")
                .replace(")", "").replace("'", "").replace(".000000000", ""))
        tail = " (overflows!)" if fl[0] else ""
        print(info + " loops {0:4.0f} and tolerance is {1:e}".format(fl[2], fl[1]) + tail)


def _finish(S, result, order, iParams, labels, flags, stats=None):
    inv = np.argsort(order)
    out = np.transpose(result, inv)
    sv = S.values
    same = (isinstance(sv, np.ndarray) and out.shape == sv.shape and out.strides == sv.strides and out.dtype == sv.dtype
            and out.__array_interface__['data'][0] == sv.__array_interface__['data'][0])
    if not same:                                # (the solve ran in place on S.values when no re-layout was needed)
        sv[...] = out.astype(sv.dtype, copy=False)
    iParams["flags_all"] = flags
    if stats is not None:
        iParams["stats"] = stats
    fl = iParams.get("flags")
    if isinstance(fl, np.ndarray) and fl.shape == (3,):
        fl[:] = flags[-1]
    else:
        iParams["flags"] = flags[-1].copy()
    _report(iParams, labels, flags)
    return S


def _engine_kw(iParams):
    return dict(ordering=iParams.get("ordering", "colour"), engine=iParams.get("engine", "auto"), accel=iParams.get("accel"),
                ctx=iParams.get("ctx"), devices=iParams.get("devices"))


def _flags_in(iParams):
    fl = iParams.get("flags")
    return np.array([0.0, 1.0, 0.0]) if fl is None else np.asarray(fl, dtype=np.float64)


def inv_standard2D(A, B, C, F, S, dims, iParams):
    r"""Invert :math:`\partial_y(A\psi_y + B\psi_x) + \partial_x(B\psi_y + C\psi_x) = F`
    by SOR on every 2-D slice spanned by ``dims`` (core.py:88-155)."""
    if len(dims) != 2:
        raise Exception("2 dimensions are needed for inversion")
    all_dims, noncore, order = _layout(F, dims)
    arrs = [_as_batched(X, all_dims, order, dims, n) for X, n in ((S, "S"), (A, "A"), (B, "B"), (C, "C"), (F, "F"))]
    Sv, Av, Bv, Cv, Fv = arrs
    flags, stats = solvers.solve_standard_2D(
        Sv, Av, Bv, Cv, Fv, iParams["BCs"][0], iParams["BCs"][1], iParams["del1Sqr"], iParams["ratioQtr"],
        iParams["ratioSqr"], iParams["optArg"], _undeftmp, _flags_in(iParams), iParams["mxLoop"],
        iParams["tolerance"], **_engine_kw(iParams))
    return _finish(S, Sv, order, iParams, _slice_labels(F, noncore), flags, stats)


def inv_general2D(A, B, C, D, E, F, G, S, dims, iParams):
    r"""Invert :math:`A\psi_{yy} + B\psi_{yx} + C\psi_{xx} + D\psi_y + E\psi_x + F\psi = G`
    by SOR on every 2-D slice spanned by ``dims`` (core.py:374-444)."""
    if len(dims) != 2:
        raise Exception("2 dimensions are needed for inversion")
    all_dims, noncore, order = _layout(G, dims)
    names = ("S", "A", "B", "C", "D", "E", "F", "G")
    Sv, Av, Bv, Cv, Dv, Ev, Fv, Gv = [_as_batched(X, all_dims, order, dims, n)
                                      for X, n in zip((S, A, B, C, D, E, F, G), names)]
    flags, stats = solvers.solve_general_2D(
        Sv, Av, Bv, Cv, Dv, Ev, Fv, Gv, iParams["BCs"][0], iParams["BCs"][1], iParams["del1"], iParams["del1Sqr"],
        iParams["ratio"], iParams["ratioQtr"], iParams["ratioSqr"], iParams["optArg"], _undeftmp,
        _flags_in(iParams), iParams["mxLoop"], iParams["tolerance"], **_engine_kw(iParams))
    return _finish(S, Sv, order, iParams, _slice_labels(G, noncore), flags, stats)


def inv_standard3D(A, B, C, F, S, dims, iParams):
    r"""Invert :math:`\partial_z(A\omega_z) + \partial_y(B\omega_y) + \partial_x(C\omega_x) = F`
    by SOR on every 3-D volume spanned by ``dims`` (core.py:20-85)."""
    if len(dims) != 3:
        raise Exception("3 dimensions are needed for inversion")
    all_dims, noncore, order = _layout(F, dims)
    Sv, Av, Bv, Cv, Fv = [_as_batched(X, all_dims, order, dims, n)
                          for X, n in ((S, "S"), (A, "A"), (B, "B"), (C, "C"), (F, "F"))]
    flags, stats = solvers.solve_standard_3D(
        Sv, Av, Bv, Cv, Fv, iParams["BCs"][0], iParams["BCs"][1], iParams["BCs"][2], iParams["del1Sqr"],
        iParams["ratio2Sqr"], iParams["ratio1Sqr"], iParams["optArg"], _undeftmp, _flags_in(iParams),
        iParams["mxLoop"], iParams["tolerance"], **_engine_kw(iParams))
    return _finish(S, Sv, order, iParams, _slice_labels(F, noncore), flags, stats)
