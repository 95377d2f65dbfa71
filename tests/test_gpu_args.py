"""GPU: the edges of the input domain at the C-ABI and at the ndarray level -- empty batches, the smallest grids,
wrong shapes / dtypes / codes, the begin / step / end protocol -- through the same entries the parity tests use.

The reference raises Python exceptions for shape and argument errors (core.py:126-127, apps.py:1358-1359) and returns
normally for everything numerical (overflow is a flag); the library mirrors that: negative return code -> XinvError
(with .code), never a crash, never a silent fallback."""
import ctypes as C

import numpy as np
import pytest

import oracle
import xinvert_b200 as xb
from tests import cases
from xinvert_b200 import _lib, solvers

pytestmark = pytest.mark.gpu


def test_empty_batch_is_a_no_op(gpu_ctx):
    """batch = 0 (an empty selection along the non-core dims): nothing to solve, nothing written, no error."""
    S = np.zeros((0, 12, 16))
    A = np.ones((12, 16))
    fl, st = solvers.solve_standard_2D(S, A, None, A, np.zeros((0, 12, 16)), "fixed", "fixed", 1.0, 0.25, 1.0, 1.4)
    assert fl.shape == (0, 3) and st["cell_updates"] == 0
    fl, st = solvers.solve_standard_3D(np.zeros((0, 5, 6, 8)), *(np.ones((5, 6, 8)),) * 3, np.zeros((0, 5, 6, 8)), "fixed", "fixed",
                                       "fixed", 1.0, 1.0, 1.0, 1.3)
    assert fl.shape == (0, 3)
    F = xb.DataArray(np.zeros((0, 12, 16)), ['time', 'lat', 'lon'],
                     {'time': np.arange(0), 'lat': np.linspace(-50, 50, 12), 'lon': np.linspace(0, 337.5, 16)})
    out = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams={'BCs': ['fixed', 'periodic'], 'printInfo': False})
    assert out.values.shape == (0, 12, 16)


@pytest.mark.parametrize("shape", [(3, 3), (3, 4), (4, 3)])
def test_smallest_grids_every_engine_that_takes_them(gpu_ctx, shape):
    """One interior row / column: engines that cannot take the grid refuse it when forced, 'auto' always solves it."""
    c = cases.random_std2d_rowcoef(*shape, seed=3, land=0.0)
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "fixed", 5, -1.0, omega=1.2, ordering="colour")
    for engine in ("auto", "colour", "resident"):
        S_g, f_g = cases.run_std2d(xb, c, "fixed", "fixed", 5, -1.0, omega=1.2, engine=engine)
        assert np.array_equal(S_g, S_o) and f_g[2] == f_o[2]
    for engine in ("fused", "cluster"):
        try:
            S_g, f_g = cases.run_std2d(xb, c, "fixed", "fixed", 5, -1.0, omega=1.2, engine=engine)
            assert np.array_equal(S_g, S_o)
        except xb.XinvError as e:
            assert e.code == _lib.E_UNSUPPORTED


def test_shape_dtype_and_code_errors(gpu_ctx):
    c = cases.random_std2d(12, 16, with_B=False, seed=1)
    p = c["p"]
    S = c["S0"].copy()
    with pytest.raises(ValueError):                                   # S.shape != (yc, xc): the shim checks what numba would index
        xb.invert_standard_2D(S, c["A"], None, c["C"], c["F"], 13, 16, 1.0, 1.0, "fixed", "fixed", p["del1Sqr"], p["ratioQtr"],
                              p["ratioSqr"], 1.4, cases.UNDEF, np.zeros(3), 10, 1e-8)
    with pytest.raises(ValueError):                                   # a coefficient of another shape
        solvers.solve_standard_2D(S, c["A"][:-1], None, c["C"], c["F"], "fixed", "fixed", 1.0, 0.25, 1.0, 1.4)
    with pytest.raises(KeyError):                                     # an unknown boundary condition
        solvers.solve_standard_2D(S, c["A"], None, c["C"], c["F"], "open", "fixed", 1.0, 0.25, 1.0, 1.4)
    with pytest.raises(TypeError):                                    # S is updated in place: it must be an array
        solvers.solve_standard_2D(S.tolist(), c["A"], None, c["C"], c["F"], "fixed", "fixed", 1.0, 0.25, 1.0, 1.4)
    with pytest.raises(ValueError):                                   # flags of the wrong length
        solvers.solve_standard_2D(S, c["A"], None, c["C"], c["F"], "fixed", "fixed", 1.0, 0.25, 1.0, 1.4, flags=np.zeros(4))
    # float32 S: solved in float64, written back into the caller's float32 array (in place, as the reference does)
    S32 = c["S0"].astype(np.float32)
    solvers.solve_standard_2D(S32, c["A"], None, c["C"], c["F"], "fixed", "fixed", p["del1Sqr"], p["ratioQtr"], p["ratioSqr"], 1.4,
                              mxLoop=5, tolerance=-1.0)
    S_o, _ = cases.run_std2d(oracle, dict(c, S0=c["S0"].astype(np.float32).astype(np.float64)), "fixed", "fixed", 5, -1.0, omega=1.4,
                             ordering="colour")
    assert S32.dtype == np.float32 and np.array_equal(S32, S_o.astype(np.float32))


def test_c_abi_argument_errors_and_protocol(gpu_ctx):
    L = _lib.load()
    c = cases.random_std2d(12, 16, with_B=False, seed=2)
    p = c["p"]
    S = c["S0"].copy()
    fl = np.array([[0.0, 1.0, 0.0]])
    vp = lambda a: C.c_void_p(a.ctypes.data)
    tail = (1, 12, 16, 0, 0, p["del1Sqr"], p["ratioQtr"], p["ratioSqr"], 1.4, cases.UNDEF, vp(fl), 10, -1.0, None)

    def code(rc):
        with pytest.raises(xb.XinvError) as ei:
            _lib.check(rc)
        return ei.value.code

    assert code(L.xinv_std2d(gpu_ctx.handle, None, vp(c["A"]), None, vp(c["C"]), vp(c["F"]), *tail)) == _lib.E_ARG        # S == NULL
    assert code(L.xinv_std2d(gpu_ctx.handle, vp(S), None, None, vp(c["C"]), vp(c["F"]), *tail)) == _lib.E_ARG            # A == NULL
    bad_bc = (1, 12, 16, 7, 0) + tail[5:]
    assert code(L.xinv_std2d(gpu_ctx.handle, vp(S), vp(c["A"]), None, vp(c["C"]), vp(c["F"]), *bad_bc)) == _lib.E_ARG
    neg = (1, 12, 16, 0, 0) + tail[5:11] + (-1, -1.0, None)
    assert code(L.xinv_std2d(gpu_ctx.handle, vp(S), vp(c["A"]), None, vp(c["C"]), vp(c["F"]), *neg)) == _lib.E_ARG       # mxLoop < 0
    opts = _lib.make_opts()
    opts.struct_size = 8
    assert code(L.xinv_std2d(gpu_ctx.handle, vp(S), vp(c["A"]), None, vp(c["C"]), vp(c["F"]), *tail[:-1], C.byref(opts))) == _lib.E_ARG
    # protocol: step / end without begin; begin twice
    n = C.c_int64(0)
    assert code(L.xinv_step(gpu_ctx.handle, 1, C.byref(n))) == _lib.E_STATE
    assert code(L.xinv_end(gpu_ctx.handle)) == _lib.E_STATE
    _lib.check(L.xinv_std2d_begin(gpu_ctx.handle, vp(S), vp(c["A"]), None, vp(c["C"]), vp(c["F"]), *tail))
    assert code(L.xinv_std2d_begin(gpu_ctx.handle, vp(S), vp(c["A"]), None, vp(c["C"]), vp(c["F"]), *tail)) == _lib.E_STATE
    while True:
        _lib.check(L.xinv_step(gpu_ctx.handle, 3, C.byref(n)))
        if n.value == 0:
            break
    _lib.check(L.xinv_end(gpu_ctx.handle))
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "fixed", 10, -1.0, omega=1.4, ordering="colour")
    assert np.array_equal(S, S_o) and fl[0, 2] == f_o[2]
    # the context is usable again after every refused call
    S2, f2 = cases.run_std2d(xb, c, "fixed", "fixed", 10, -1.0, omega=1.4)
    assert np.array_equal(S2, S_o)
