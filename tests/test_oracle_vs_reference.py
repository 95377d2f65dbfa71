"""CPU, authoring container only: the C oracle against the UNMODIFIED reference
numba kernels imported from /root/reference (skipped where that tree or numba
is absent, e.g. on the GPU box -- the committed fixtures of
tests/test_oracle_golden.py cover that case).  Bar: BIT-EXACT."""
import numpy as np
import pytest

import oracle
from oracle import ref_loader
from tests import cases
from tests.golden import make_golden as mg

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference or numba not available")

BCS = [("fixed", "fixed"), ("fixed", "periodic"), ("extend", "fixed"), ("extend", "periodic")]


def _ub(bcy, bcx, shape):
    """numbas.py:297-301: with BCy='extend' and non-periodic x the reference's second
    copy loop runs i over range(1, yc-1) while indexing the x axis; for yc-1 > xc that
    is an out-of-bounds access in nopython mode (no bounds check: it reads/writes the
    neighbouring rows or segfaults).  No defined behaviour to compare with; the oracle
    and the CUDA path clip the loop to the row."""
    return bcy == "extend" and bcx != "periodic" and shape[-2] - 1 > shape[-1]


@pytest.fixture(scope="module")
def ref():
    return ref_loader.ref_numbas()


@pytest.mark.parametrize("bcy,bcx", BCS)
def test_std2d_matches_reference(ref, bcy, bcx):
    for with_B, shape, seed in [(False, (31, 44), 1), (True, (31, 44), 2), (False, (50, 37), 3), (True, (17, 23), 4)]:
        if _ub(bcy, bcx, shape):
            continue
        c = cases.random_std2d(*shape, with_B=with_B, seed=seed)
        S_r, f_r = mg.ref_std2d(ref, c, bcy, bcx, 11, -1.0, 1.3)
        S_o, f_o = cases.run_std2d(oracle, c, bcy, bcx, 11, -1.0, omega=1.3)
        assert np.array_equal(S_o, S_r)
        assert np.array_equal(f_o, f_r)


@pytest.mark.parametrize("bcy,bcx", BCS)
def test_gen2d_matches_reference(ref, bcy, bcx):
    for with_B, shape, seed in [(False, (31, 44), 5), (True, (31, 44), 6), (False, (50, 37), 7)]:
        if _ub(bcy, bcx, shape):
            continue
        c = cases.random_gen2d(*shape, with_B=with_B, seed=seed)
        S_r, f_r = mg.ref_gen2d(ref, c, bcy, bcx, 11, -1.0, 1.3)
        S_o, f_o = cases.run_gen2d(oracle, c, bcy, bcx, 11, -1.0, omega=1.3)
        assert np.array_equal(S_o, S_r)
        assert np.array_equal(f_o, f_r)


@pytest.mark.parametrize("bcy,bcx", BCS)
def test_std3d_matches_reference(ref, bcy, bcx):
    for shape, seed in [((6, 14, 19), 8), ((9, 11, 12), 9)]:
        c = cases.random_std3d(*shape, seed=seed)
        S_r, f_r = mg.ref_std3d(ref, c, bcy, bcx, 8, -1.0, 1.3)
        S_o, f_o = cases.run_std3d(oracle, c, bcy, bcx, 8, -1.0, omega=1.3)
        assert np.array_equal(S_o, S_r)
        assert np.array_equal(f_o, f_r)


@pytest.mark.parametrize("bcy,bcx", BCS)
def test_new_kernels_match_reference(ref, bcy, bcx):
    """invert_standard_2D_test, invert_general_3D (incl. the west-column condition that skips H,
    numbas.py:869) and invert_standard_1D, lexicographic order, bit for bit."""
    for shape, seed in [((31, 44), 31), ((17, 23), 32)]:
        c = cases.random_std2dt(*shape, seed=seed)
        S_r, f_r = cases.run_std2dt(ref, c, bcy, bcx, 11, -1.0, omega=1.2)
        S_o, f_o = cases.run_std2dt(oracle, c, bcy, bcx, 11, -1.0, omega=1.2)
        assert np.array_equal(S_o, S_r) and np.array_equal(f_o, f_r)
    for shape, seed in [((6, 14, 19), 33), ((9, 11, 12), 34)]:
        c = cases.random_gen3d(*shape, seed=seed)
        c["H"][:, :, 0] = cases.UNDEF                      # periodic-x: the west column is updated all the same
        S_r, f_r = cases.run_gen3d(ref, c, bcy, bcx, 8, -1.0, omega=1.3)
        S_o, f_o = cases.run_gen3d(oracle, c, bcy, bcx, 8, -1.0, omega=1.3)
        assert np.array_equal(S_o, S_r) and np.array_equal(f_o, f_r)
    c = cases.random_std1d(57, seed=35)
    for b in (bcy, bcx):
        S_r, f_r = cases.run_std1d(ref, c, b, 30, -1.0)
        S_o, f_o = cases.run_std1d(oracle, c, b, 30, -1.0)
        assert np.array_equal(S_o, S_r) and np.array_equal(f_o, f_r)
    S_r, f_r = cases.run_std1d(ref, c, "fixed", 5000, 1e-9)
    S_o, f_o = cases.run_std1d(oracle, c, "fixed", 5000, 1e-9)
    assert np.array_equal(S_o, S_r) and np.array_equal(f_o, f_r) and f_r[2] > 20


@pytest.mark.parametrize("bcy,bcx", BCS)
def test_biharmonic_matches_reference(ref, bcy, bcx):
    """invert_general_bih_2D (numbas.py:1204-1586), lexicographic order, bit for bit -- including the two-row extend
    condition, the edge columns' own operation order in the G term and the stale inner-loop index in the B term of the
    two east columns (numbas.py:1495-1497, :1540-1542), for nx a multiple of 3 and not."""
    for shape, seed in [((21, 27), 1), ((17, 23), 2), ((30, 33), 3), ((19, 12), 4), ((9, 8), 5)]:
        if _ub(bcy, bcx, shape):
            continue
        c = cases.random_bih(*shape, seed=seed)
        for sweeps in (0, 6):
            S_r, f_r = cases.run_bih(ref, c, bcy, bcx, sweeps, -1.0)
            S_o, f_o = cases.run_bih(oracle, c, bcy, bcx, sweeps, -1.0)
            assert np.isfinite(S_r[S_r != cases.UNDEF]).all()
            assert np.array_equal(S_o, S_r) and np.array_equal(f_o, f_r)


def test_c1_known_answer(ref):
    """SURVEY.md 8c KAT (6): 360x180 lat-lon Poisson, fixed/periodic, omega 1.4,
    tol 1e-8: the reference stops at loop 2380 with max|psi| = 13182413.993245527;
    the oracle reproduces field and flags bit for bit."""
    c = cases.poisson_latlon(180, 360, land=False, noise=0.0, seed=0)
    S_r, f_r = mg.ref_std2d(ref, c, "fixed", "periodic", 5000, 1e-8, 1.4)
    S_o, f_o = cases.run_std2d(oracle, c, "fixed", "periodic", 5000, 1e-8, omega=1.4)
    assert f_r[2] == 2380 and np.isclose(np.abs(S_r).max(), 13182413.993245527, rtol=1e-12)
    assert np.array_equal(S_o, S_r) and np.array_equal(f_o, f_r)


def test_colour_orderings_match_masked_reference_bridge(ref):
    """Red-black / 4-colour oracle == alternating masked one-sweep reference calls."""
    c = cases.random_std2d(26, 34, with_B=False, seed=21)
    for bcx in ("fixed", "periodic"):
        S_b = mg.bridge_redblack_std2d(ref, c, bcx, 4, 1.4)
        S_o, _ = cases.run_std2d(oracle, c, "fixed", bcx, 3, -1.0, omega=1.4, ordering="colour")
        assert np.array_equal(S_o, S_b)
    c = cases.random_std2d(26, 34, with_B=True, seed=22)
    for bcx in ("fixed", "periodic"):
        S_b = mg.bridge_fourcolour_std2d(ref, c, bcx, 4, 1.2)
        S_o, _ = cases.run_std2d(oracle, c, "fixed", bcx, 3, -1.0, omega=1.2, ordering="colour")
        assert np.array_equal(S_o, S_b)
    c = cases.random_std3d(6, 10, 12, seed=23)
    S_b = mg.bridge_redblack_std3d(ref, c, "periodic", 4, 1.3)
    S_o, _ = cases.run_std3d(oracle, c, "fixed", "periodic", 3, -1.0, omega=1.3, ordering="colour")
    assert np.array_equal(S_o, S_b)
