"""Pin the oracle: golden vectors of the reference, reference npts.c == C port == SciPy banded LU."""
import os

import numpy as np
import pytest

from oracle import cfd_oracle as O
from tests.conftest import GOLD, convergence_ratios

KNOWN = {32: "0.0000293338", 64: "0.0000010363", 128: "0.0000000420", 256: "0.0000000019"}


def relinf(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("n", sorted(KNOWN))
def test_known_answer_port(n):
    """d/dx sin on [0, 2pi]: mean |cos - df| printed by the reference's test_npts (test_npts.c:146-157)."""
    x = np.arange(n) * (2 * np.pi / (n - 1))
    f = np.broadcast_to(np.sin(x), (2, 3, n)).copy()
    df = O.derivative(f, 0, 2 * np.pi / (n - 1))
    assert "%.10f" % np.mean(np.abs(np.cos(x) - df)) == KNOWN[n]


def test_known_answer_fixture_matches_reference_output():
    ka = np.load(os.path.join(GOLD, "known_answer.npz"))
    for n, v in KNOWN.items():
        assert str(ka[str(n)]) == "Average absolute error: " + v


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("n", [32, 64])
def test_known_answer_reference_binary(n):
    assert O.ref_known_answer(n) == "Average absolute error: " + KNOWN[n]


def test_port_matches_reference_fixture():
    """Committed output of the reference's own npts.c (one rank) on seeded RHS."""
    g = np.load(os.path.join(GOLD, "npts_ref.npz"))
    for n in (8, 32, 48, 64, 100, 256, 1024):
        r, u = g[f"r_{n}"], g[f"u_{n}"]
        beta, gam = O.npts_beta_gam(n)
        assert np.array_equal(beta, g[f"beta_{n}"]) and np.array_equal(gam, g[f"gam_{n}"])
        assert relinf(O.npts_solve(r, 0), u) < 1e-15
        assert relinf(O.near_toeplitz_solve(r, O.PADE, 0), u) < 1e-14
        assert relinf(O.scipy_solve_axis(r, O.PADE, 0), u) < 1e-14
        if n & (n - 1) == 0:
            assert relinf(O.cr_solve(r, O.PADE), u) < 1e-14      # the reference GPU algorithm, restated


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built on this box")
def test_port_matches_live_reference():
    rng = np.random.default_rng(1)
    for n in (16, 128, 512):
        r = rng.random((3, 4, n))
        u, beta, gam = O.ref_npts_solve(r)
        assert relinf(O.npts_solve(r, 0), u) < 1e-15


def test_derivative_fixture_all_axes():
    g = np.load(os.path.join(GOLD, "derivative.npz"))
    f = g["f"]
    for axis in range(3):
        h = float(g[f"h_{axis}"])
        assert relinf(O.derivative(f, axis, h), g[f"df_{axis}"]) < 1e-14
        assert relinf(O.scipy_derivative(f, axis, h), g[f"df_{axis}"]) < 1e-14


def test_rhs_formulas():
    """code/cuda/kernels.cu:34,38,44 written out by hand for one line."""
    rng = np.random.default_rng(2)
    f = rng.random((1, 1, 9))
    h = 0.3
    r = O.rhs(f, 0, h)[0, 0]
    l = f[0, 0]
    assert r[0] == (1. / (2 * h)) * (-5 * l[0] + 4 * l[1] + l[2])
    assert r[8] == -(1. / (2 * h)) * (-5 * l[8] + 4 * l[7] + l[6])
    for i in range(1, 8):
        assert r[i] == (3. / (4 * h)) * (l[i + 1] - l[i - 1])


def test_general_coefficients_vs_lapack():
    """code/ocl/test/test_near_toeplitz.py:31-48: (1,2,3,4,5,6,7), n = 32, random RHS, rtol 1e-7."""
    rng = np.random.default_rng(3)
    co = (1., 2., 3., 4., 5., 6., 7.)
    d = rng.random((1, 1, 32))
    want = O.scipy_solve_axis(d, co, 0)
    np.testing.assert_allclose(O.near_toeplitz_solve(d, co), want, rtol=1e-7)
    np.testing.assert_allclose(O.cr_solve(d, co), want, rtol=1e-7)


def test_pthomas_vs_lapack():
    """code/cuda/test/test_kernels.py:29-53: strided systems d[32,2,2], random a, b, c."""
    rng = np.random.default_rng(4)
    n = 32
    a, b, c = rng.random(n), rng.random(n) + 2, rng.random(n)
    d = rng.random((n, 2, 2))
    want = O.scipy_solve_banded(a, b, c, d.reshape(n, -1)).reshape(d.shape)
    np.testing.assert_allclose(O.pthomas(a, b, c, d), want, rtol=1e-7)


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("P", [2, 4])
def test_partition_algebra(axis, P):
    """The reference's multi-rank method (compact.py:65-154) reproduces the one-rank derivative."""
    rng = np.random.default_rng(5)
    f = rng.random((16, 24, 32))
    assert relinf(O.partition_derivative(f, axis, 0.1, P), O.scipy_derivative(f, axis, 0.1)) < 1e-14


def test_analytic_fields():
    """Known-answer fields of the reference tests (code/ocl/test/test_compact.py:15-73), decimal=2."""
    n = 32
    z, y, x = np.meshgrid(*(np.linspace(0, 2 * np.pi, n),) * 3, indexing="ij")
    h = 2 * np.pi / (n - 1)
    np.testing.assert_almost_equal(O.derivative(np.sin(x), 0, h), np.cos(x), decimal=2)
    np.testing.assert_almost_equal(O.derivative(x * y * z, 0, h), y * z, decimal=2)
    np.testing.assert_almost_equal(O.derivative(np.sin(y), 1, h), np.cos(y), decimal=2)
    np.testing.assert_almost_equal(O.derivative(x * y * z, 1, h), x * z, decimal=2)
    np.testing.assert_almost_equal(O.derivative(x * y * z ** 2, 2, h), 2 * x * y * z, decimal=2)


# ---------------------------------------------------------------------------------------------------
# the reference's DISTRIBUTED npts solve (emulated ranks) -- lanl-implementation/python/test_npts.py:13-54
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("npx,shape", [(3, (12, 6, 36)), (1, (2, 3, 32)), (2, (2, 3, 32)), (4, (3, 2, 64)), (8, (1, 2, 64)),
                                       (6, (2, 2, 24))])
def test_distributed_npts_restatement(npx, shape):
    """The reference's own check: random right-hand side, every x line against the banded LU (its 3x3x3-rank case is
    36 x 18 x 36 with 12 points per rank along the line); here to 1e-14 instead of rtol 1e-7, and against the
    one-rank npts solve of the port."""
    rng = np.random.default_rng(npx)
    r = rng.random(shape)
    got = O.npts_distributed_solve(r, npx)
    assert relinf(got, O.scipy_solve_axis(r, O.PADE, 0)) < 1e-14
    assert relinf(got, O.npts_solve(r, 0)) < 1e-14
    b1, g1 = O.npts_beta_gam(shape[2])
    bd, gd = O.npts_distributed_beta_gam(shape[2] // npx, npx)
    assert np.array_equal(bd.ravel(), b1) and np.array_equal(gd.ravel(), g1)    # the hand-off reproduces the serial pivots


def test_distributed_npts_c_twin_defect():
    """Why the restatement follows python/npts.py and not npts.c: with `product_1 = 0.0` (npts.c:525; 1.0 in
    python/npts.py:365) the R->L combination loses the ranks more than one block away -- exact for npx <= 2, wrong
    for short blocks at npx >= 3 (the "high error for the last elements in each block" of lanl-implementation/README.md)."""
    rng = np.random.default_rng(1)
    r = rng.random((2, 2, 16))
    want = O.scipy_solve_axis(r, O.PADE, 0)
    assert relinf(O.npts_distributed_solve(r, 2, c_twin_product_zero=True), want) < 1e-14
    bad = O.npts_distributed_solve(r, 4, c_twin_product_zero=True)
    assert relinf(bad, want) > 1e-6
    blk = np.abs(bad - want).reshape(2, 2, 4, 4).max(axis=(0, 1))      # [rank, position in block]
    assert blk[0].argmax() == 3 and blk[1].argmax() == 3               # worst at the last element of a block
    assert relinf(O.npts_distributed_solve(r, 4), want) < 1e-14


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_order_of_convergence_oracle(axis):
    """code/cuda/test/test_convergence.py (prints, no assert): 4th-order interior -> mean error falls ~16x per
    doubling; 3rd-order closures dominate the max error -> ~8x."""
    O.port().oracle_set_num_threads(os.cpu_count() or 1)
    mean_r, max_r = convergence_ratios(O.derivative, axis)
    assert 14.0 < mean_r[-1] < 17.5, mean_r
    assert 6.5 < max_r[-1] < 9.0, max_r
    assert all(r > 8.0 for r in mean_r)


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("n", [4, 7, 33, 130])
def test_oracle_structural_properties(axis, n):
    """Size-independent properties of the scheme the reference implements (code/cuda/kernels.cu:34-44 RHS,
    near_toeplitz.py:36-50 matrix): the derivative of a constant is zero, polynomials up to degree 3 are differentiated
    exactly (4th-order interior rows, 3rd-order closure rows), the operator is linear, and mirroring the line mirrors the
    derivative with the opposite sign (the closures at the two ends are mirror images of each other)."""
    rng = np.random.default_rng(n + axis)
    shape = [3, 4, 6]
    shape[2 - axis] = n
    h = 0.37
    x = np.arange(n) * h - 0.4 * n * h
    idx = [None, None, None]
    idx[2 - axis] = slice(None)
    xs = x[tuple(idx)] * np.ones(shape)
    assert np.abs(O.derivative(np.full(shape, 3.25), axis, h)).max() <= 1e-13
    c = rng.random(4)
    poly = c[0] + c[1] * xs + c[2] * xs ** 2 + c[3] * xs ** 3
    dpoly = c[1] + 2 * c[2] * xs + 3 * c[3] * xs ** 2
    scale = np.abs(poly).max() / h
    assert np.abs(O.derivative(poly, axis, h) - dpoly).max() <= 2e-13 * scale
    f, g = rng.random(shape), rng.random(shape)
    lin = O.derivative(2.5 * f - 1.5 * g, axis, h) - (2.5 * O.derivative(f, axis, h) - 1.5 * O.derivative(g, axis, h))
    assert np.abs(lin).max() <= 1e-12 * np.abs(O.derivative(f, axis, h)).max()
    mirrored = np.flip(O.derivative(np.flip(f, 2 - axis).copy(), axis, h), 2 - axis)
    assert np.abs(mirrored + O.derivative(f, axis, h)).max() <= 1e-12 * np.abs(O.derivative(f, axis, h)).max()
    # a quartic is NOT exact at the closures: the test above is sharp
    if n >= 7:
        assert np.abs(O.derivative(xs ** 4, axis, h) - 4 * xs ** 3).max() > 1e-6 * np.abs(4 * xs ** 3).max()
