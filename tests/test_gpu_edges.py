"""Edges of the loop control and of the input domain on the GPU, against the oracle (and, where the
reference itself defines the behaviour, against what numbas.py does):

* the ``norm == 0`` exit of invert_standard_2D (numbas.py:410): a zero forcing on a zero guess stops after
  the first sweep; the general form and the 3-D form have no such exit (:1196, :206);
* a slice that is ``undef`` everywhere: the norm's count is 0 -> NaN -> overflow flag (numbas.py:1724-1727,
  :403-405), also inside a batch whose other slices go on;
* P3 at BASELINE configs[0] size: the converged red-black field equals the converged lexicographic field of
  the reference order to <= 1e-10 relative at omega_opt;
* float32 inputs: the reference iterates in float32, this library promotes to float64 -- the size of that
  documented deviation.
"""
import numpy as np
import pytest

import oracle
import xinvert_b200 as xb
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kw", [dict(engine="fused"), dict(engine="colour"), dict(ordering="lexicographic")])
@pytest.mark.parametrize("bcy,bcx", [("fixed", "periodic"), ("extend", "fixed")])
def test_norm_zero_exit_std2d(gpu_ctx, kw, bcy, bcx):
    c = cases.poisson_latlon(40, 72, land=True, noise=0.0, seed=0)
    c["F"][c["F"] != cases.UNDEF] = 0.0                  # zero forcing, zero initial guess: nothing ever changes
    order = "lexicographic" if "ordering" in kw else "colour"
    S_o, f_o = cases.run_std2d(oracle, c, bcy, bcx, 500, 1e-8, omega=1.4, ordering=order)
    S_g, f_g = cases.run_std2d(xb, c, bcy, bcx, 500, 1e-8, omega=1.4, **kw)
    assert f_o[2] == 0.0 and f_o[0] == 0.0 and f_o[1] == 1.0          # one sweep, |0 - DBL_MAX| / DBL_MAX
    assert np.array_equal(f_g, f_o)
    assert np.array_equal(S_g, S_o) and not S_g.any()


def test_no_norm_zero_exit_in_general_and_3d_forms(gpu_ctx):
    """numbas.py:1196 / :206: with norm == 0 the relative change is 0/0 = NaN from the second sweep on,
    never below the tolerance: the loop runs to mxLoop (the oracle and the GPU agree; numba's default
    error model would raise ZeroDivisionError there)."""
    c = cases.random_gen2d_rowcoef(30, 64, seed=3, land=0.1)
    c["G"][c["G"] != cases.UNDEF] = 0.0
    c["S0"][...] = 0.0
    S_o, f_o = cases.run_gen2d(oracle, c, "fixed", "periodic", 7, 1e-8, ordering="colour")
    S_g, f_g = cases.run_gen2d(xb, c, "fixed", "periodic", 7, 1e-8)
    assert f_o[2] == 7.0 and f_g[2] == 7.0 and f_g[0] == f_o[0]
    assert np.isnan(f_g[1]) and np.isnan(f_o[1])
    assert np.array_equal(S_g, S_o)
    c3 = cases.random_std3d(6, 20, 64, seed=4)
    c3["F"][c3["F"] != cases.UNDEF] = 0.0
    c3["S0"][...] = 0.0
    for engine in ("fused", "colour"):
        S_o, f_o = cases.run_std3d(oracle, c3, "fixed", "periodic", 5, 1e-8, ordering="colour")
        S_g, f_g = cases.run_std3d(xb, c3, "fixed", "periodic", 5, 1e-8, engine=engine)
        assert f_o[2] == 5.0 and f_g[2] == 5.0 and np.isnan(f_g[1])
        assert np.array_equal(S_g, S_o)


@pytest.mark.parametrize("engine", ["fused", "colour"])
def test_all_undef_slice_sets_the_overflow_flag(gpu_ctx, engine):
    """count == 0 -> norm = NaN -> flags[0] = 1 and the loop breaks (numbas.py:1724-1727, :403-405); the
    incoming flags[1], flags[2] are left as they were.  In a batch, the other slices go on."""
    c = cases.random_std2d_rowcoef(30, 64, seed=5, batch=3)
    c["S0"][1] = cases.UNDEF
    c["F"][1] = cases.UNDEF                              # (no cell of that slice is ever updated: psi stays undef)
    p = c["p"]
    S = c["S0"].copy()
    fl, st = xb.solve_standard_2D(S, c["A"], None, c["C"], c["F"], "fixed", "periodic", p["del1Sqr"], p["ratioQtr"],
                                  p["ratioSqr"], 1.4, cases.UNDEF, mxLoop=6, tolerance=-1.0, engine=engine)
    for t in range(3):
        ct = dict(A=c["A"], C=c["C"], F=np.ascontiguousarray(c["F"][t]), S0=c["S0"][t], p=p)   # A, C: one slice shared by the batch
        S_o, f_o = cases.run_std2d(oracle, ct, "fixed", "periodic", 6, -1.0, omega=1.4, ordering="colour")
        assert np.array_equal(S[t], S_o), t
        assert fl[t, 0] == f_o[0] and fl[t, 2] == f_o[2] and np.isclose(fl[t, 1], f_o[1], rtol=1e-6), (t, fl[t], f_o)
    assert fl[1, 0] == 1.0 and fl[1, 2] == 0.0 and fl[0, 0] == 0.0 and fl[0, 2] == 6.0
    assert (S[1] == cases.UNDEF).all()


def test_all_undef_volume_3d(gpu_ctx):
    c = cases.random_std3d(5, 20, 64, seed=6)
    c["S0"][...] = cases.UNDEF
    c["F"][...] = cases.UNDEF
    for engine in ("fused", "colour"):
        S_o, f_o = cases.run_std3d(oracle, c, "extend", "periodic", 6, -1.0, ordering="colour")
        S_g, f_g = cases.run_std3d(xb, c, "extend", "periodic", 6, -1.0, engine=engine)
        assert f_o[0] == 1.0 and np.array_equal(f_g, f_o)
        assert np.array_equal(S_g, S_o)


def test_p3_converged_redblack_equals_lexicographic_at_c1_size(gpu_ctx):
    """BASELINE configs[0] grid (360x180 lat-lon, fixed/periodic) at omega_opt, both orderings iterated until the
    stop test's relative change is exactly 0 or 20000 sweeps: the fields agree to <= 1e-10 relative
    (6e-11 measured with the reference's own kernel); the sweep counts differ."""
    c = cases.poisson_latlon(180, 360, land=False, noise=0.0)
    S_lex, f_lex = cases.run_std2d(oracle, c, "fixed", "periodic", 20000, 0.0, ordering="lexicographic")
    S_rb, f_rb = cases.run_std2d(xb, c, "fixed", "periodic", 20000, 0.0, engine="fused")
    S_rbo, f_rbo = cases.run_std2d(oracle, c, "fixed", "periodic", 20000, 0.0, ordering="colour")
    assert np.array_equal(S_rb, S_rbo) and f_rb[2] == f_rbo[2]          # P1 on the way
    rel = np.abs(S_rb - S_lex).max() / np.abs(S_lex).max()
    print(f"P3 at C1 size: sweeps lex {int(f_lex[2]) + 1}, red-black {int(f_rb[2]) + 1}, max rel diff {rel:.3e}")
    assert rel <= 1e-10


def test_float32_inputs_promotion_deviation(gpu_ctx):
    """float32 operands (the reference's real-data tests feed NetCDF float32): numba then iterates with float32
    stores; this library promotes to float64, solves, and casts the result back.  The deviation from a float32
    iteration -- emulated here by the oracle run on the float32 values with a float32 round trip of psi after
    every sweep -- stays at float32 round-off level of the field's scale."""
    c = cases.poisson_latlon(45, 90, land=True, noise=1e-6, seed=2)
    c32 = {k: (v.astype(np.float32) if isinstance(v, np.ndarray) else v) for k, v in c.items()}
    S = c32["S0"].copy()
    p = c["p"]
    fl, _ = xb.solve_standard_2D(S, c32["A"], None, c32["C"], c32["F"], "fixed", "periodic", p["del1Sqr"], p["ratioQtr"],
                                 p["ratioSqr"], 1.4, cases.UNDEF, mxLoop=299, tolerance=-1.0)
    assert S.dtype == np.float32                                      # in place, cast back
    # float32-iterate emulation: one oracle sweep at a time, psi rounded to float32 in between
    c64 = {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in c32.items()}
    Se = c64["S0"].copy()
    for _ in range(300):
        ce = dict(c64, S0=Se)
        Se, _ = cases.run_std2d(oracle, ce, "fixed", "periodic", 0, -1.0, omega=1.4, ordering="colour")
        Se = Se.astype(np.float32).astype(np.float64)
    scale = np.abs(Se).max()
    dev = np.abs(S.astype(np.float64) - Se).max() / scale
    print(f"float32 inputs: promoted-to-float64 solve vs float32-iterate emulation, max rel deviation {dev:.3e}")
    assert dev < 5e-5                                                 # ~300 sweeps of float32 round-off (eps = 6e-8)


@pytest.mark.parametrize("pinned", [False, True])
def test_float32_io_of_the_poisson_front_end(gpu_ctx, pinned):
    """xinv_opts.io_f32: a float32 forcing goes to the device as float32 and the result comes back as
    float32 -- half the PCIe bytes -- widened / narrowed on the device around the same float64 solve: bit-equal to the
    host-built path (which promotes on the host and casts the result back), land mask and batch included."""
    import xinvert_b200 as xb
    from tests.test_apps_host import _c1_zeta
    ny, nx, T = 60, 120, 5
    zeta, co = _c1_zeta(ny, nx)
    rng = np.random.default_rng(4)
    z = np.stack([zeta * (1 + 0.3 * t) + 1e-6 * rng.standard_normal((ny, nx)) for t in range(T)]).astype(np.float32)
    lam, phi = np.deg2rad(co['lon'])[None, :], np.deg2rad(co['lat'])[:, None]
    z[:, np.sin(5 * lam) * np.cos(3 * phi) > 0.6] = np.nan
    if pinned:
        zp = xb.pinned_empty(z.shape, np.float32)
        zp[...] = z
        z = zp
    F = xb.DataArray(z, ['time', 'lat', 'lon'], dict(co, time=np.arange(T)))
    ip = {'BCs': ['extend', 'periodic'], 'tolerance': 1e-8, 'mxLoop': 3000, 'printInfo': False}
    ip1 = dict(ip)
    s1 = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=ip1)                      # device front end, float32 I/O
    st = ip1['stats']
    assert s1.values.dtype == np.float32
    assert st['h2d_bytes'] < 4 * z.size + 4096 and st['d2h_bytes'] == 4 * z.size     # float32 both ways (+ the row vectors)
    s2 = xb.invert_Poisson(F, dims=['lat', 'lon'], iParams=dict(ip, engine='colour'))   # host-built path, promoted
    assert s2.values.dtype == np.float32
    assert np.array_equal(s1.values, s2.values, equal_nan=True)
    assert np.isnan(s1.values).sum() == np.isnan(z).sum()


def test_float32_io_of_the_omega_and_gill_matsuno_front_ends(gpu_ctx, monkeypatch):
    """The same for xinv_std3d_rows (invert_omega) and xinv_gen2d_rows (invert_GillMatsuno): float32 in, float32 out,
    bit-equal to the host-built path on the promoted values."""
    import xinvert_b200 as xb
    from xinvert_b200 import apps
    DA = xb.DataArray
    nz, ny, nx, T = 9, 30, 64, 2
    lev = 100000.0 - 10000.0 * np.arange(nz)
    lat, lon = -58.0 + 4.0 * np.arange(ny), 5.625 * np.arange(nx)
    rng = np.random.default_rng(5)
    co = {'time': np.arange(T), 'LEV': lev, 'lat': lat, 'lon': lon}
    Fv = (1e-17 * rng.standard_normal((T, nz, ny, nx))).astype(np.float32)
    Fv[:, 2:5, 8:12, 10:20] = np.nan
    F = DA(Fv, ['time', 'LEV', 'lat', 'lon'], co)
    N2 = DA(1e-6 * (1 + 0.5 * rng.random(nz)), ['LEV'], {'LEV': lev})
    ip = {'BCs': ['fixed', 'fixed', 'periodic'], 'tolerance': 1e-9, 'mxLoop': 40, 'printInfo': False}
    kw = dict(dims=['LEV', 'lat', 'lon'], coords='lat-lon', mParams={'N2': N2})
    ip_f = dict(ip)
    w_f = xb.invert_omega(F, iParams=ip_f, **kw)
    assert w_f.values.dtype == np.float32 and ip_f['stats']['d2h_bytes'] == 4 * Fv.size
    with monkeypatch.context() as m:
        m.setattr(apps, "_omega_device_front", lambda *a, **k: None)
        w_h = xb.invert_omega(F, iParams=dict(ip), **kw)
    assert w_h.values.dtype == np.float32
    assert np.array_equal(w_f.values, w_h.values, equal_nan=True)
    # Gill-Matsuno on a beta plane (general form)
    ny, nx = 48, 96
    y, x = np.linspace(-5e6, 5e6, ny), np.linspace(0, 4e7, nx, endpoint=False)
    Q = (0.05 * np.exp(-((y[:, None] / 1e6) ** 2 + ((x[None, :] - 2e7) / 2e6) ** 2))).astype(np.float32)
    Qd = DA(Q, ['y', 'x'], {'y': y, 'x': x})
    ipg = {'BCs': ['fixed', 'periodic'], 'tolerance': 1e-9, 'mxLoop': 300, 'optArg': 1.4, 'printInfo': False}
    kwg = dict(dims=['y', 'x'], coords='cartesian', mParams={'f0': 0.0, 'beta': 2e-11, 'epsilon': 1e-5, 'Phi': 5000.0})
    ip_f = dict(ipg)
    h_f = xb.invert_GillMatsuno(Qd, iParams=ip_f, **kwg)
    assert h_f.values.dtype == np.float32 and ip_f['stats']['d2h_bytes'] == 4 * Q.size
    with monkeypatch.context() as m:
        m.setattr(apps, "_general_device_front", lambda *a, **k: None)
        h_h = xb.invert_GillMatsuno(Qd, iParams=dict(ipg), **kwg)
    assert np.array_equal(h_f.values, h_h.values, equal_nan=True)
