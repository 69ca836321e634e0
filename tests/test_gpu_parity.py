"""
Parity of the CUDA path (through the C ABI, via the Python mirror of the reference classes) against the oracle.
Tolerance: <= 1e-12 relative L-infinity (BASELINE.json north_star), written in TOL below.
"""
import os

import numpy as np
import pytest

from oracle import cfd_oracle as O
from tests.conftest import GOLD, convergence_ratios

pytestmark = pytest.mark.gpu
TOL = 1e-12


def relinf(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.fixture(scope="module")
def C():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import compact_finite_differences_b200 as pkg
    return pkg


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def smooth(shape):
    nz, ny, nx = shape
    z, y, x = np.meshgrid(np.linspace(0, 2 * np.pi, nz), np.linspace(0, 2 * np.pi, ny),
                          np.linspace(0, 2 * np.pi, nx), indexing="ij")
    return x, y, z


# ---------------------------------------------------------------------------------------------------
# golden fixtures generated from the reference itself (oracle/make_golden.py)
# ---------------------------------------------------------------------------------------------------
def test_golden_derivative_fixture(C):
    g = np.load(os.path.join(GOLD, "derivative.npz"))
    f = g["f"]
    fd = dev(f)
    for axis in range(3):
        s = C.CompactFiniteDifferenceSolver(f.shape, float(g[f"h_{axis}"]), axis)
        got = s(fd).cpu().numpy()
        assert relinf(got, g[f"df_{axis}"]) <= TOL


def test_golden_npts_fixture(C):
    g = np.load(os.path.join(GOLD, "npts_ref.npz"))
    for n in (8, 32, 48, 64, 100, 256, 1024):
        r, u = g[f"r_{n}"], g[f"u_{n}"]
        d = dev(r)
        C.NearToeplitzSolver(r.shape, O.PADE).solve(d)
        assert relinf(d.cpu().numpy(), u) <= TOL


@pytest.mark.parametrize("n", [32, 64, 128, 256])
def test_known_answer(C, n):
    """BASELINE configs[0]: d/dx sin, 'Average absolute error' of the reference's test_npts."""
    known = {32: "0.0000293338", 64: "0.0000010363", 128: "0.0000000420", 256: "0.0000000019"}
    m = min(n, 64)
    x = np.arange(n) * (2 * np.pi / (n - 1))
    f = np.broadcast_to(np.sin(x), (m, m, n)).copy()
    s = C.CompactFiniteDifferenceSolver(f.shape, 2 * np.pi / (n - 1), 0)
    df = s(dev(f)).cpu().numpy()
    assert "%.10f" % np.mean(np.abs(np.cos(x) - df)) == known[n]


# ---------------------------------------------------------------------------------------------------
# derivative vs oracle on seeded random fields: ragged shapes, every axis
# ---------------------------------------------------------------------------------------------------
SHAPES = [(8, 8, 8), (8, 32, 16), (4, 6, 10), (5, 7, 34), (33, 31, 32), (16, 65, 66), (64, 64, 64), (70, 40, 96),
          (3, 5, 130), (129, 4, 6), (6, 200, 8), (32, 32, 256), (40, 130, 34)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_derivative_random(C, shape, axis):
    if shape[2 - axis] < 4:
        pytest.skip("line shorter than 4")
    rng = np.random.default_rng(hash((shape, axis)) % 2 ** 32)
    f = rng.random(shape)
    h = 0.01 + 0.1 * axis
    want = O.derivative(f, axis, h)
    got = C.CompactFiniteDifferenceSolver(shape, h, axis)(dev(f)).cpu().numpy()
    assert relinf(got, want) <= TOL


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_order_of_convergence(C, axis):
    """code/cuda/test/test_convergence.py:25-52 on the CUDA path: mean error ~16x per doubling (4th-order interior),
    max error ~8x (3rd-order closures); the same bounds the oracle meets in tests/test_oracle.py."""
    def derivative(f, ax, h):
        return C.CompactFiniteDifferenceSolver(f.shape, h, ax)(dev(f)).cpu().numpy()
    mean_r, max_r = convergence_ratios(derivative, axis)
    assert 14.0 < mean_r[-1] < 17.5, mean_r
    assert 6.5 < max_r[-1] < 9.0, max_r
    assert all(r > 8.0 for r in mean_r)


def test_reference_spellings_and_analytic(C):
    """code/ocl/test/test_compact.py:15-73 (decimal=2) through dfdx / dfdy / dfdz."""
    shape = (16, 32, 32)
    x, y, z = smooth(shape)
    s = C.CompactFiniteDifferenceSolver(shape)
    dx, dy, dz = x[0, 0, 1] - x[0, 0, 0], y[0, 1, 0] - y[0, 0, 0], z[1, 0, 0] - z[0, 0, 0]
    np.testing.assert_almost_equal(s.dfdx(dev(np.sin(x)), dx).cpu().numpy(), np.cos(x), decimal=2)
    np.testing.assert_almost_equal(s.dfdx(dev(x * y * z), dx).cpu().numpy(), y * z, decimal=2)
    np.testing.assert_almost_equal(s.dfdy(dev(np.sin(y)), dy).cpu().numpy(), np.cos(y), decimal=2)
    np.testing.assert_almost_equal(s.dfdy(dev(x * y * z), dy).cpu().numpy(), x * z, decimal=2)
    np.testing.assert_almost_equal(s.dfdz(dev(x * y * z ** 2), dz).cpu().numpy(), 2 * x * y * z, decimal=1)


def test_host_buffer_path(C):
    rng = np.random.default_rng(11)
    f = rng.random((12, 20, 64))
    s = C.CompactFiniteDifferenceSolver(f.shape, 0.05, 1)
    assert relinf(s(f), O.derivative(f, 1, 0.05)) <= TOL


def test_256_cubed_all_axes(C):
    """BASELINE configs[1]: 256^3, x, y, z, smooth field, full-field parity + analytic derivative."""
    import torch
    n = 256
    x, y, z = smooth((n, n, n))
    f = np.sin(x) * np.cos(y) * np.sin(z)
    h = 2 * np.pi / (n - 1)
    fd = dev(f)
    exact = [np.cos(x) * np.cos(y) * np.sin(z), -np.sin(x) * np.sin(y) * np.sin(z), np.sin(x) * np.cos(y) * np.cos(z)]
    O.port().oracle_set_num_threads(os.cpu_count() or 1)
    for axis in range(3):
        got = C.CompactFiniteDifferenceSolver((n, n, n), h, axis)(fd).cpu().numpy()
        assert relinf(got, O.derivative(f, axis, h)) <= TOL
        assert np.abs(got - exact[axis]).max() < 1e-5      # 3rd-order closure rows dominate: 2.7e-6 at N = 256
    torch.cuda.synchronize()


def test_512_cubed_properties(C):
    """BASELINE configs[2] at full size through size-independent properties: sampled-line parity with the
    oracle (4096 random lines + the boundary lines per axis), linearity, analytic derivative."""
    import torch
    n = 512
    h = 2 * np.pi / (n - 1)
    t = torch.linspace(0, 2 * np.pi, n, dtype=torch.float64, device="cuda")
    z, y, x = t[:, None, None], t[None, :, None], t[None, None, :]
    f = (torch.sin(x) * torch.cos(y) * torch.sin(z)).contiguous()
    g = (x * torch.cos(x * y) + y * torch.sin(z)).contiguous()          # reference demo field, run.py:29-30
    rng = np.random.default_rng(512)
    for axis in range(3):
        s = C.CompactFiniteDifferenceSolver((n, n, n), h, axis)
        df, dg = s(f), s(g)
        comb = s((2.0 * f - 0.5 * g).contiguous())
        lin = (comb - (2.0 * df - 0.5 * dg)).abs().max().item() / comb.abs().max().item()
        assert lin < 1e-13                                              # linearity of the operator
        # sampled lines against the oracle
        ax = 2 - axis
        fl = torch.movedim(g, ax, 2).reshape(-1, n)
        dl = torch.movedim(dg, ax, 2).reshape(-1, n)
        idx = np.unique(np.concatenate([rng.integers(0, n * n, 4096), [0, n - 1, n * n - n, n * n - 1]]))
        it = torch.from_numpy(idx).cuda()
        lines = fl[it].cpu().numpy()
        want = O.derivative(lines.reshape(1, -1, n), 0, h).reshape(-1, n)
        assert relinf(dl[it].cpu().numpy(), want) <= TOL
        exact = [torch.cos(x) * torch.cos(y) * torch.sin(z), -torch.sin(x) * torch.sin(y) * torch.sin(z),
                 torch.sin(x) * torch.cos(y) * torch.cos(z)][axis]
        assert (df - exact).abs().max().item() < 1e-6       # ~3.4e-7 at N = 512 (closure rows)
        del df, dg, comb


# ---------------------------------------------------------------------------------------------------
# d/dx and d/dy in one launch (cfd_apply_xy) and the gradient built on it
# ---------------------------------------------------------------------------------------------------
XY_SHAPES = [(1, 32, 4), (3, 32, 34), (5, 64, 64), (2, 96, 130), (17, 128, 66), (9, 32, 256), (4, 256, 32), (70, 64, 96),
             (6, 40, 34)]      # the last one has ny % 32 != 0: two-launch fallback inside the same call


@pytest.mark.parametrize("shape", XY_SHAPES)
def test_fused_xy_random(C, shape):
    rng = np.random.default_rng(hash(shape) % 2 ** 32)
    f = rng.random(shape)
    hx, hy = 0.013, 0.21
    s = C.CompactFiniteDifferenceSolver(shape)
    gx, gy = s.dfdxy(dev(f), hx, hy)
    assert relinf(gx.cpu().numpy(), O.derivative(f, 0, hx)) <= TOL
    assert relinf(gy.cpu().numpy(), O.derivative(f, 1, hy)) <= TOL


@pytest.mark.parametrize("warps,slots", [(1, 3), (8, 3), (2, 4), (7, 4)])
def test_fused_xy_launch_shapes(C, warps, slots):
    """Every ring / warp configuration of stream_kernel_xy, and the plain plane-by-plane draw order."""
    shape = (11, 96, 160)
    rng = np.random.default_rng(7)
    f = rng.random(shape)
    want = O.derivative(f, 0, 0.1), O.derivative(f, 1, 0.2)
    s = C.CompactFiniteDifferenceSolver(shape)
    try:
        C.lib().cfd_set_launch(warps, 0, slots)
        for active in ("", "0", "3"):
            if active:
                os.environ["CFD_XY_ACTIVE"] = active
            gx, gy = s.dfdxy(dev(f), 0.1, 0.2)
            assert relinf(gx.cpu().numpy(), want[0]) <= TOL and relinf(gy.cpu().numpy(), want[1]) <= TOL
    finally:
        os.environ.pop("CFD_XY_ACTIVE", None)
        C.lib().cfd_set_launch(0, 0, 0)


@pytest.mark.parametrize("sub", [1, 2, 3, 16, 2 | (4 << 8), 3 | (1 << 8), 16 | (32 << 8), 40 | (2 << 8)])
@pytest.mark.parametrize("shape", [(5, 160, 192), (3, 96, 258), (2, 64, 1056), (2, 1056, 64), (3, 1024, 1024)])
def test_fused_xy_sub_plane_wavefronts(C, sub, shape):
    """CFD_XY_SUB: lines cut into segments of `sub` chunks (warm-up chunk in front, look-ahead chunk behind), squares
    of sub x sub tiles as wavefronts of their own.  Every segment length, ragged last chunks, a line whose last
    segment is a single chunk (33 chunks, sub 16), and the whole-line case (sub >= chunks).  sub = x | y << 8 cuts the
    x lines and the y lines differently (rectangles; 16 | 32 << 8 is the library default for lines of >= 32 tiles)."""
    rng = np.random.default_rng(hash((sub, shape)) % 2 ** 32)
    f = rng.random(shape)
    want = O.derivative(f, 0, 0.1), O.derivative(f, 1, 0.2)
    try:
        os.environ["CFD_XY_SUB"] = str(sub)
        s = C.CompactFiniteDifferenceSolver(shape)
        gx, gy = s.dfdxy(dev(f), 0.1, 0.2)
        assert relinf(gx.cpu().numpy(), want[0]) <= TOL and relinf(gy.cpu().numpy(), want[1]) <= TOL
    finally:
        os.environ.pop("CFD_XY_SUB", None)
    gx, gy = s.dfdxy(dev(f), 0.1, 0.2)                  # the same plan returns to its default cut when the knob goes
    assert relinf(gx.cpu().numpy(), want[0]) <= TOL and relinf(gy.cpu().numpy(), want[1]) <= TOL


def test_fused_xy_per_plan_warps(C):
    """cfd_plan_set_xy_warps: the per-plan knob a caller uses to leave room for kernels running beside the launch."""
    shape = (37, 64, 130)
    rng = np.random.default_rng(3)
    f = rng.random(shape)
    s = C.CompactFiniteDifferenceSolver(shape)
    for warps in (5, 2, 0, 7):
        gx, gy = s.dfdxy(dev(f), 0.1, 0.2, warps=warps)
        assert relinf(gx.cpu().numpy(), O.derivative(f, 0, 0.1)) <= TOL
        assert relinf(gy.cpu().numpy(), O.derivative(f, 1, 0.2)) <= TOL
    assert C.lib().cfd_plan_set_xy_warps(s._plan(1, 0.2).handle, 5) == -1       # CFD_EINVAL: not an axis-0 plan
    assert C.lib().cfd_plan_set_xy_warps(s._plan(0, 0.1).handle, 9) == -1


def test_gradient_256_cubed(C):
    """BASELINE configs[1] through gradient(): one fused d/dx + d/dy launch and one d/dz launch, full-field parity."""
    n = 256
    x, y, z = smooth((n, n, n))
    f = x * np.cos(x * y) + y * np.sin(z)                  # reference demo field, perf-test/multi-GPU/PyCUDA/run.py:29-30
    h = 2 * np.pi / (n - 1)
    O.port().oracle_set_num_threads(os.cpu_count() or 1)
    g = C.CompactFiniteDifferenceSolver((n, n, n)).gradient(dev(f), (h, h, h))
    for axis in range(3):
        assert relinf(g[axis].cpu().numpy(), O.derivative(f, axis, h)) <= TOL


def test_fused_xy_512_cubed_matches_separate_launches(C):
    """At BASELINE's full size the fused launch runs the same arithmetic per line as dfdx / dfdy: bit-identical."""
    import torch
    n = 512
    f = torch.rand((n, n, n), dtype=torch.float64, device="cuda")
    s = C.CompactFiniteDifferenceSolver((n, n, n))
    gx, gy = s.dfdxy(f, 0.1, 0.2)
    assert torch.equal(gx, s.dfdx(f, 0.1))
    assert torch.equal(gy, s.dfdy(f, 0.2))


def test_gradient_two_streams_is_bit_equal_and_capturable(C, monkeypatch):
    """gradient() = cfd_apply_xyz: d/dz on a side stream of the library, forked / joined with events; same bits as the
    one-stream sequence, correct when called back to back on changing inputs, and capturable into a CUDA graph."""
    import torch
    rng = np.random.default_rng(77)
    shape = (40, 64, 96)
    hs = (0.1, 0.2, 0.3)
    s = C.CompactFiniteDifferenceSolver(shape)
    fs = [dev(rng.random(shape)) for _ in range(3)]
    monkeypatch.setenv("CFD_XYZ_SERIAL", "1")
    refs = [[o.clone() for o in s.gradient(f, hs)] for f in fs]
    monkeypatch.delenv("CFD_XYZ_SERIAL")
    out = [torch.empty(shape, dtype=torch.float64, device="cuda") for _ in range(3)]
    for rep in range(3):
        for f, ref in zip(fs, refs):
            s.gradient(f, hs, out)
            chk = [o.clone() for o in out]            # ordered on the current stream: the join event has been waited on
            assert all(torch.equal(a, b) for a, b in zip(chk, ref))
    for a in range(3):
        assert relinf(refs[0][a].cpu().numpy(), O.derivative(fs[0].cpu().numpy(), a, hs[a])) <= TOL
    st = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=st):
        s.gradient(fs[1], hs, out)
    for o in out:
        o.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(out, refs[1]))


def test_fused_xy_rejects_bad_arguments(C):
    import torch
    f = torch.rand((4, 32, 32), dtype=torch.float64, device="cuda")
    s = C.CompactFiniteDifferenceSolver((4, 32, 32))
    px, py, pz = s._plan(0, 0.1), s._plan(1, 0.1), s._plan(2, 0.1)
    o1, o2 = torch.empty_like(f), torch.empty_like(f)
    L = C.lib()                                   # -1 = CFD_EINVAL
    assert L.cfd_apply_xy(py.handle, px.handle, f.data_ptr(), o1.data_ptr(), o2.data_ptr(), None) == -1
    assert L.cfd_apply_xy(px.handle, pz.handle, f.data_ptr(), o1.data_ptr(), o2.data_ptr(), None) == -1
    assert L.cfd_apply_xy(px.handle, py.handle, f.data_ptr(), o1.data_ptr(), o1.data_ptr(), None) == -1
    assert L.cfd_apply_xy(px.handle, py.handle, f.data_ptr(), f.data_ptr(), o2.data_ptr(), None) == -1
    assert L.cfd_apply_xy(px.handle, py.handle, None, o1.data_ptr(), o2.data_ptr(), None) == -1
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------------------------------
# solver-only API
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [32, 64, 128, 512, 1024, 4096])
def test_near_toeplitz_pade(C, n):
    rng = np.random.default_rng(n)
    d = rng.random((4, 37, n))
    t = dev(d)
    C.NearToeplitzSolver(d.shape, O.PADE).solve(t)
    assert relinf(t.cpu().numpy(), O.near_toeplitz_solve(d, O.PADE)) <= TOL


@pytest.mark.parametrize("axis", [1, 2])
def test_near_toeplitz_other_axes(C, axis):
    rng = np.random.default_rng(axis)
    d = rng.random((96, 130, 40))
    t = dev(d)
    C.NearToeplitzSolver(d.shape, O.PADE, axis=axis).solve(t)
    assert relinf(t.cpu().numpy(), O.near_toeplitz_solve(d, O.PADE, axis)) <= TOL


def test_near_toeplitz_general_coefficients(C):
    """code/ocl/test/test_near_toeplitz.py:31-48: (1,2,3,4,5,6,7), shape (1,1,32), vs LAPACK, rtol 1e-7."""
    rng = np.random.default_rng(0)
    co = (1., 2., 3., 4., 5., 6., 7.)
    d = rng.random((1, 1, 32))
    t = dev(d)
    C.NearToeplitzSolver((1, 1, 32), co).solve(t)
    np.testing.assert_allclose(t.cpu().numpy(), O.scipy_solve_axis(d, co, 0), rtol=1e-7)
    assert relinf(t.cpu().numpy(), O.near_toeplitz_solve(d, co)) <= TOL


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("coeffs,tol", [((1., 2., 1. / 3, 1., 1. / 3, 2., 1.), 1e-12),      # 6th-order Pade, LA2 switched off
                                        ((1., 2., .45, 1., .45, 2., 1.), 1e-12),             # |g| = 0.63: 0.63^64 = 1e-13
                                        ((1., 2., 3., 4., 5., 6., 7.), 1e-9)])               # not dominant at all
def test_near_toeplitz_exact_two_pass(C, axis, coeffs, tol, monkeypatch):
    """Matrices the one-pass kernels refuse (coupling above 0.563 per row / non-convergent pivots) go through the exact
    two-pass LU (and so does alpha = 1/3 when the two-chunk look-ahead is switched off: the round-1 path)."""
    monkeypatch.setenv("CFD_NO_LA2", "1")
    rng = np.random.default_rng(axis)
    shape = [6, 10, 40]
    shape[2 - axis] = 200
    d = rng.random(shape)
    t = dev(d)
    s = C.NearToeplitzSolver(shape, coeffs, axis=axis)
    assert s.two_pass
    s.solve(t)
    want = O.near_toeplitz_solve(d, coeffs, axis)
    assert relinf(t.cpu().numpy(), want) <= tol


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("n", [96, 97, 128, 161, 1000])
def test_near_toeplitz_two_chunk_lookahead(C, axis, n):
    """alpha = 1/3 (the 6th-order scheme's matrix, |g| = 0.382): the look-ahead is derived from the coefficients -- two
    chunks, 0.382^64 = 2e-27 -- and the solve stays ONE pass (round 1: exact two-pass, 32 B / unknown)."""
    rng = np.random.default_rng(n + axis)
    co = (1., 2., 1. / 3, 1., 1. / 3, 2., 1.)
    shape = [6, 40, 34]
    n += (axis == 0 and n % 2)                     # nx is even (TMA row pitch)
    shape[2 - axis] = n
    d = rng.random(shape) - 0.5
    s = C.NearToeplitzSolver(d.shape, co, axis=axis)
    assert not s.two_pass and C.lib().nt_lookahead(s._handle) == 2
    t = dev(d)
    s.solve(t)
    assert relinf(t.cpu().numpy(), O.scipy_solve_axis(d, co, axis)) <= TOL


def test_lookahead_is_derived_from_the_coefficients(C):
    L = C.lib()
    for co, n, la, two_pass in (((1., 2., .25, 1., .25, 2., 1.), 512, 1, False),        # 0.268^32 = 5e-19
                                ((1., 2., .3, 1., .3, 2., 1.), 512, 2, False),           # |g| = 0.333: 5e-16 / 3e-31
                                ((1., 2., 1. / 3, 1., 1. / 3, 2., 1.), 512, 2, False),
                                ((1., 2., 1. / 3, 1., 1. / 3, 2., 1.), 64, 1, False),    # two chunks: exact anyway
                                ((1., 2., .45, 1., .45, 2., 1.), 512, 0, True),          # |g| = 0.63: two-pass
                                ((1., 2., 3., 4., 5., 6., 7.), 512, 0, True)):
        s = C.NearToeplitzSolver((2, 4, n), co)
        assert L.nt_lookahead(s._handle) == la and s.two_pass == two_pass, (co, n)
        d = np.random.default_rng(1).random((2, 4, n))
        t = dev(d)
        s.solve(t)
        assert relinf(t.cpu().numpy(), O.scipy_solve_axis(d, co, 0)) <= 1e-11


@pytest.mark.parametrize("scheme", ["compact6", "pade4-d2"])
@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("n", [6, 7, 31, 32, 33, 34, 64, 65, 66, 96, 97, 127, 129, 200, 512])
def test_scheme_derivative_random(C, scheme, axis, n):
    """6th-order first derivative (two chunks of look-ahead) and 4th-order second derivative through the general
    one-pass kernel, against the published schemes assembled with NumPy + banded LU; every chunk-count / ragged-tail
    case (closure rows n-2, n-1 falling into the last or the second-to-last chunk)."""
    rng = np.random.default_rng(7 * n + axis)
    shape = [5, 36, 38]
    n += (axis == 0 and n % 2)                     # nx is even (TMA row pitch): odd extents are covered along y and z
    shape[2 - axis] = n
    f = rng.random(shape)
    h = 0.37
    s = C.CompactFiniteDifferenceSolver(tuple(shape), h, axis, scheme=scheme)
    got = s(dev(f)).cpu().numpy()
    want = O.scheme_derivative(f, axis, h, scheme)
    assert relinf(got, want) <= TOL
    la = C.lib().cfd_plan_lookahead(s._plan(axis, h).handle)
    assert la == (2 if (scheme == "compact6" and n > 64) else 1)


@pytest.mark.parametrize("scheme,interior_order", [("compact6", 6), ("pade4-d2", 4)])
def test_scheme_order_of_accuracy(C, scheme, interior_order):
    """Analytic check: sin(x) on [0, 3]; away from the 3rd-order closures the error falls by 2^order per doubling."""
    errs = []
    for n in (33, 65, 129):
        x = np.linspace(0, 3.0, n)
        h = x[1] - x[0]
        f = np.broadcast_to(np.sin(x)[:, None], (2, n, 32)).copy()
        d = C.CompactFiniteDifferenceSolver(f.shape, h, 1, scheme=scheme)(dev(f)).cpu().numpy()[0, :, 0]
        true = np.cos(x) if scheme == "compact6" else -np.sin(x)
        errs.append(np.abs(d - true)[n // 4:3 * n // 4].max())
    for a, b in zip(errs[:-1], errs[1:]):
        assert a / b > 0.75 * 2 ** interior_order, errs


def test_scheme_256_cubed_and_gradient_fallback(C):
    """Full-size parity (256^3, all axes, both schemes) and the x/y pair of a non-Pade scheme (two launches inside
    cfd_apply_xy)."""
    N = 256
    zz, yy, xx = smooth((N, N, N))
    f = np.sin(xx) * np.cos(yy) + np.sin(zz) * xx
    h = 2 * np.pi / (N - 1)
    fd = dev(f)
    for scheme in ("compact6", "pade4-d2"):
        s = C.CompactFiniteDifferenceSolver((N, N, N), scheme=scheme)
        # The 3rd-order closure of the second derivative, (-27 d1 + 15 d2 - d3) / h^2 with d_k = f_k - f_0 ~ k h f', has
        # condition number ~ 50 |f'| / (h |f''|): two correct implementations that round in a different order (FMA
        # contraction) differ by that times eps -- 2e-12 here -- so its tolerance carries the conditioning explicitly.
        tol = TOL if scheme == "compact6" else max(TOL, 16 * 50 * 1.1e-16 * 2.0 / h)
        for axis, fn in enumerate((s.dfdx, s.dfdy, s.dfdz)):
            assert relinf(fn(fd, h).cpu().numpy(), O.scheme_derivative(f, axis, h, scheme)) <= tol
        gx, gy = s.dfdxy(fd, h, h)
        assert relinf(gx.cpu().numpy(), O.scheme_derivative(f, 0, h, scheme)) <= tol
        assert relinf(gy.cpu().numpy(), O.scheme_derivative(f, 1, h, scheme)) <= tol


def test_one_pass_is_kept_for_pade(C):
    assert not C.NearToeplitzSolver((4, 8, 512), O.PADE).two_pass
    assert not C.NearToeplitzSolver((1, 1, 32), (1., 2., 3., 4., 5., 6., 7.)).two_pass     # <= 2 chunks: exact anyway


def test_near_toeplitz_round_trip(C):
    """A x == d for the solved x (matrix applied with torch on the device), large batch."""
    import torch
    n, batch = 2048, 4096
    d = torch.rand((1, batch, n), dtype=torch.float64, device="cuda")
    x = d.clone()
    C.NearToeplitzSolver((1, batch, n), O.PADE).solve(x)
    ax = x.clone()
    ax[..., 1:-1] = 0.25 * x[..., :-2] + x[..., 1:-1] + 0.25 * x[..., 2:]
    ax[..., 0] = x[..., 0] + 2 * x[..., 1]
    ax[..., -1] = 2 * x[..., -2] + x[..., -1]
    assert (ax - d).abs().max().item() < 1e-13


def test_pthomas(C):
    """code/cuda/test/test_kernels.py:29-53."""
    rng = np.random.default_rng(4)
    n = 32
    a, b, c = rng.random(n), rng.random(n) + 2, rng.random(n)
    d = rng.random((n, 2, 2))
    t = dev(d)
    C.ReducedSolver((n, 2, 2)).solve(a, b, c, None, t)
    np.testing.assert_allclose(t.cpu().numpy(), O.pthomas(a, b, c, d), rtol=1e-12)


# ---------------------------------------------------------------------------------------------------
# partitioned line on ONE device: all ranks' blocks processed in turn through the multi-rank entry points
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("axis,P,shape", [(2, 2, (64, 24, 40)), (2, 4, (128, 16, 34)), (2, 8, (256, 8, 32)),
                                           (0, 4, (6, 10, 128)), (1, 2, (5, 80, 36)), (2, 2, (132, 20, 40)),
                                           (2, 3, (300, 6, 34)), (0, 2, (4, 6, 200)), (1, 4, (3, 512, 10))])
def test_partitioned_line_emulated(C, axis, P, shape):
    import torch
    rng = np.random.default_rng(P + axis)
    f = rng.random(shape)
    h = 0.13
    want = O.derivative(f, axis, h)
    ax = 2 - axis
    n = shape[ax] // P
    blocks = [np.ascontiguousarray(np.take(f, range(r * n, (r + 1) * n), axis=ax)) for r in range(P)]
    lshape = blocks[0].shape
    solvers = [C.CompactFiniteDifferenceSolver(lshape, h, axis, part=(r, P)) for r in range(P)]
    plane = blocks[0].size // n
    outs, faces = [], torch.empty((2 * P, plane), dtype=torch.float64, device="cuda")
    for r in range(P):
        fb = dev(blocks[r])
        lo = dev(np.take(blocks[r - 1], n - 1, axis=ax)) if r > 0 else None
        hi = dev(np.take(blocks[r + 1], 0, axis=ax)) if r < P - 1 else None
        out = solvers[r].apply_local(fb, None, lo, hi)
        solvers[r].interface_pack(out, faces[2 * r:2 * r + 2])
        outs.append(out)
    for r in range(P):
        solvers[r].reduced_correct(outs[r], faces)
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=ax)
    assert relinf(got, want) <= TOL
    assert relinf(got, O.partition_derivative(f, axis, h, P)) <= TOL
    if n >= 66:
        # fused path: interface planes from the block ends, reduced solve folded into the one kernel
        faces2 = torch.zeros_like(faces)
        halos = []
        for r in range(P):
            lo = dev(np.take(blocks[r - 1], n - 1, axis=ax)) if r > 0 else None
            hi = dev(np.take(blocks[r + 1], 0, axis=ax)) if r < P - 1 else None
            halos.append((lo, hi))
            solvers[r].edge_faces(dev(blocks[r]), faces2[2 * r:2 * r + 2], lo, hi)
        assert (faces2 - faces).abs().max().item() <= 1e-13 * max(1.0, faces.abs().max().item())
        ab = torch.empty((2, plane), dtype=torch.float64, device="cuda")
        outs2 = []
        for r in range(P):
            solvers[r].reduced_unknowns(faces2, ab)
            outs2.append(solvers[r].apply_coupled(dev(blocks[r]), None, halos[r][0], halos[r][1], ab))
        got2 = np.concatenate([o.cpu().numpy() for o in outs2], axis=ax)
        assert relinf(got2, want) <= TOL
        # neighbour-only exchange: each rank sees its own faces and one plane from each neighbour
        outs3 = []
        for r in range(P):
            pv, own = solvers[r].nb_layout()
            nbuf = torch.zeros((2 * pv, plane), dtype=torch.float64, device="cuda")
            nbuf[2 * own:2 * own + 2] = faces2[2 * r:2 * r + 2]
            if r > 0:
                nbuf[2 * own - 1] = faces2[2 * r - 1]
            if r < P - 1:
                nbuf[2 * own + 2] = faces2[2 * r + 2]
            solvers[r].reduced_unknowns(nbuf, ab, neighbours_only=True)
            outs3.append(solvers[r].apply_coupled(dev(blocks[r]), None, halos[r][0], halos[r][1], ab))
        got3 = np.concatenate([o.cpu().numpy() for o in outs3], axis=ax)
        assert relinf(got3, want) <= TOL


def test_partition_nccl_two_gpus():
    """The real multi-process path (one rank per GPU, NCCL) when the box has >= 2 GPUs."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(root, "scripts", "check_partition_nccl.py"), "256"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("P,shape", [(2, (132, 12, 34)), (4, (4 * 70, 6, 32))])
def test_peer_memory_protocol_one_device(C, P, shape):
    """comm="nvlink" kernels (cfd_push_planes, cfd_edge_faces_p2p, cfd_wait_flags, cfd_reduced_unknowns) with all P
    ranks' buffers on one device: "peer" addresses are plain local addresses, every producer is enqueued before
    any consumer waits, so the single stream cannot deadlock.  Two consecutive calls exercise both parities."""
    import ctypes
    import torch
    from compact_finite_differences_b200._lib import check, lib
    L = lib()
    rng = np.random.default_rng(P)
    h = 0.17
    n = shape[0] // P
    plane = shape[1] * shape[2]
    solvers = [C.CompactFiniteDifferenceSolver((n,) + shape[1:], h, 2, part=(r, P)) for r in range(P)]
    bufs = [torch.zeros(16 * plane + 16, dtype=torch.float64, device="cuda") for _ in range(P)]
    base = [b.data_ptr() for b in bufs]
    halo = lambda r, par, s: base[r] + 8 * ((par * 2 + s) * plane)              # noqa: E731
    faces = lambda r, par, i: base[r] + 8 * (4 * plane + (par * 6 + i) * plane)  # noqa: E731
    flag = lambda r, k: base[r] + 8 * (16 * plane + k)                          # noqa: E731
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for seq in (1, 2, 3):
        f = rng.random(shape)
        want = O.derivative(f, 2, h)
        blocks = [dev(f[r * n:(r + 1) * n]) for r in range(P)]
        par = seq & 1
        for r in range(P):
            lf, rt = (r - 1 if r > 0 else None), (r + 1 if r < P - 1 else None)
            check(L.cfd_push_planes(blocks[r][0].data_ptr() if lf is not None else None,
                                    halo(lf, par, 1) if lf is not None else None,
                                    blocks[r][-1].data_ptr() if rt is not None else None,
                                    halo(rt, par, 0) if rt is not None else None, plane,
                                    flag(lf, 1) if lf is not None else None, flag(rt, 0) if rt is not None else None,
                                    seq, st))
        for r in range(P):
            lf, rt = (r - 1 if r > 0 else None), (r + 1 if r < P - 1 else None)
            pv, own = solvers[r].nb_layout()
            check(L.cfd_wait_flags(flag(r, 0) if lf is not None else None, flag(r, 1) if rt is not None else None,
                                   seq, st))
            own_left = 1 if (lf is not None and lf > 0) else 0
            plan = solvers[r]._plan(2, h)
            check(L.cfd_edge_faces_p2p(plan.handle, blocks[r].data_ptr(),
                                       halo(r, par, 0) if lf is not None else None,
                                       halo(r, par, 1) if rt is not None else None,
                                       faces(r, par, 2 * own),
                                       faces(lf, par, 2 * own_left + 2) if lf is not None else None,
                                       faces(rt, par, 1) if rt is not None else None,
                                       flag(lf, 3) if lf is not None else None, flag(rt, 2) if rt is not None else None,
                                       seq, st))
        outs = []
        ab = torch.empty((2, plane), dtype=torch.float64, device="cuda")
        for r in range(P):
            lf, rt = (r - 1 if r > 0 else None), (r + 1 if r < P - 1 else None)
            out = torch.empty_like(blocks[r])
            plan = solvers[r]._plan(2, h)
            check(L.cfd_reduced_unknowns(plan.handle, faces(r, par, 0), 1, ab.data_ptr(),
                                         flag(r, 2) if lf is not None else None, flag(r, 3) if rt is not None else None,
                                         seq, st))
            check(L.cfd_apply_coupled(plan.handle, blocks[r].data_ptr(), out.data_ptr(),
                                      halo(r, par, 0) if lf is not None else None,
                                      halo(r, par, 1) if rt is not None else None, ab.data_ptr(), st))
            outs.append(out)
        got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
        assert relinf(got, want) <= TOL


@pytest.mark.parametrize("P,shape", [(2, (132, 12, 34)), (3, (3 * 66, 8, 32)), (4, (4 * 70, 6, 32))])
def test_one_launch_exchange_one_device(C, P, shape):
    """comm="nvlink" as ZPartitionedDerivative drives it now: cfd_edge_faces_push (faces without the neighbour points,
    own boundary rows pushed into the neighbours' halo slots, one flag per side) and cfd_reduced_unknowns_deferred
    (halo terms folded in with the plan's weights), all P ranks' buffers on one device, three calls (both parities)."""
    import ctypes
    import torch
    from compact_finite_differences_b200._lib import check, lib
    L = lib()
    rng = np.random.default_rng(10 + P)
    h = 0.23
    n = shape[0] // P
    plane = shape[1] * shape[2]
    solvers = [C.CompactFiniteDifferenceSolver((n,) + shape[1:], h, 2, part=(r, P)) for r in range(P)]
    bufs = [torch.zeros(16 * plane + 16, dtype=torch.float64, device="cuda") for _ in range(P)]
    base = [b.data_ptr() for b in bufs]
    halo = lambda r, par, s: base[r] + 8 * ((par * 2 + s) * plane)              # noqa: E731
    faces = lambda r, par, i: base[r] + 8 * (4 * plane + (par * 6 + i) * plane)  # noqa: E731
    flag = lambda r, k: base[r] + 8 * (16 * plane + k)                          # noqa: E731
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    zz, yy, xx = smooth(shape)
    for seq in (1, 2, 3, 4):
        # three random fields, then a smooth one at large amplitude: the faces must not lose digits to the guessed
        # neighbour points (a zero guess would: the face would be a difference of numbers 1/h times larger)
        f = rng.random(shape) if seq < 4 else 1e3 * (np.sin(zz) * np.cos(yy) + xx)
        want = O.derivative(f, 2, h)
        blocks = [dev(f[r * n:(r + 1) * n]) for r in range(P)]
        par = seq & 1
        for r in range(P):                         # producers: no wait, no halo needed
            lf, rt = (r - 1 if r > 0 else None), (r + 1 if r < P - 1 else None)
            pv, own = solvers[r].nb_layout()
            own_left = 1 if (lf is not None and lf > 0) else 0
            check(L.cfd_edge_faces_push(solvers[r]._plan(2, h).handle, blocks[r].data_ptr(), faces(r, par, 2 * own),
                                        faces(lf, par, 2 * own_left + 2) if lf is not None else None,
                                        faces(rt, par, 1) if rt is not None else None,
                                        halo(lf, par, 1) if lf is not None else None,
                                        halo(rt, par, 0) if rt is not None else None,
                                        flag(lf, 3) if lf is not None else None, flag(rt, 2) if rt is not None else None,
                                        seq, st))
        outs = []
        ab = torch.empty((2, plane), dtype=torch.float64, device="cuda")
        for r in range(P):                         # consumers
            lf, rt = (r - 1 if r > 0 else None), (r + 1 if r < P - 1 else None)
            out = torch.empty_like(blocks[r])
            plan = solvers[r]._plan(2, h)
            hl = halo(r, par, 0) if lf is not None else None
            hh = halo(r, par, 1) if rt is not None else None
            check(L.cfd_reduced_unknowns_deferred(plan.handle, faces(r, par, 0), hl, hh, blocks[r].data_ptr(),
                                                  ab.data_ptr(), flag(r, 2) if lf is not None else None,
                                                  flag(r, 3) if rt is not None else None, seq, st))
            check(L.cfd_apply_coupled(plan.handle, blocks[r].data_ptr(), out.data_ptr(), hl, hh, ab.data_ptr(), st))
            outs.append(out)
        torch.cuda.synchronize()
        for r in range(1, P):                      # the pushed halos are the neighbours' boundary rows, bit for bit
            got_lo = bufs[r][(par * 2 + 0) * plane:(par * 2 + 1) * plane].cpu().numpy()
            assert np.array_equal(got_lo, f[r * n - 1].ravel())
            got_hi = bufs[r - 1][(par * 2 + 1) * plane:(par * 2 + 2) * plane].cpu().numpy()
            assert np.array_equal(got_hi, f[r * n].ravel())
        got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
        assert relinf(got, want) <= TOL
        if seq == 4:                               # ... and no worse than the one-rank kernel on the whole line
            whole = C.CompactFiniteDifferenceSolver(shape, h, 2)(dev(f)).cpu().numpy()
            assert relinf(got, want) <= 4 * relinf(whole, want) + 1e-15


@pytest.mark.parametrize("P,shape", [(2, (132, 64, 96)), (3, (3 * 66, 32, 40)), (4, (4 * 70, 32, 64)), (2, (140, 12, 34)),
                                     (3, (3 * 128, 64, 256))])
def test_zpart_c_abi_one_process(C, P, shape, monkeypatch):
    """cfd_zpart_* (the C-side driver of the partitioned d/dz) with all P ranks in ONE process on one device: buffers
    wired with cfd_zpart_connect_ptr, one stream per rank, CTA counts capped so that the P persistent kernels are
    co-resident (a rank's phase B polls for what its neighbours' phase A, on another stream, stores).  Default path
    (one-kernel d/dz, kernels_zx.cuh) and the three-launch path (CFD_NO_ZX: edge items inside the x/y kernel or a
    separate edge launch, reduce, coupled kernel) against the oracle and against each other."""
    import ctypes
    import torch
    from compact_finite_differences_b200._lib import check, lib
    L = lib()
    rng = np.random.default_rng(40 + P)
    hs = (0.19, 0.07, 0.23)
    n = shape[0] // P
    lshape = (n,) + tuple(shape[1:])
    zsol = [C.CompactFiniteDifferenceSolver(lshape, hs[2], 2, part=(r, P)) for r in range(P)]
    xy = C.CompactFiniteDifferenceSolver(lshape)
    px, py = xy._plan(0, hs[0]), xy._plan(1, hs[1])
    zps = []
    for r in range(P):
        h = ctypes.c_void_p()
        check(L.cfd_zpart_create(ctypes.byref(h), zsol[r]._plan(2, hs[2]).handle))
        check(L.cfd_zpart_set_ctas(h, 148 // P))
        zps.append(h)
    bufs = [L.cfd_zpart_buffer(z) for z in zps]
    for r in range(P):
        check(L.cfd_zpart_connect_ptr(zps[r], bufs[r - 1] if r > 0 else None, bufs[r + 1] if r < P - 1 else None))
    assert L.cfd_zpart_connect_ptr(zps[0], bufs[1], bufs[1]) != 0        # rank 0 has no left neighbour
    streams = [torch.cuda.Stream() for _ in range(P)]
    zz, yy, xx = smooth(shape)

    def run(kind, blocks, outs):
        torch.cuda.synchronize()
        for r in range(P):
            sp = ctypes.c_void_p(streams[r].cuda_stream)
            if kind == "begin+apply":
                check(L.cfd_zpart_begin(zps[r], blocks[r].data_ptr(), sp))
            if kind in ("apply", "begin+apply"):
                check(L.cfd_zpart_apply(zps[r], blocks[r].data_ptr(), outs[r][2].data_ptr(), sp))
            else:
                check(L.cfd_zpart_apply_xyz(zps[r], px.handle, py.handle, blocks[r].data_ptr(), outs[r][0].data_ptr(),
                                            outs[r][1].data_ptr(), outs[r][2].data_ptr(), sp))
        torch.cuda.synchronize()
        assert L.cfd_async_status() == 0

    try:
        for it in range(4):
            f = rng.random(shape) if it < 3 else 1e3 * (np.sin(zz) * np.cos(yy) + xx)
            want = [O.derivative(f, a, hs[a]) for a in range(3)]
            blocks = [dev(f[r * n:(r + 1) * n]) for r in range(P)]
            results = {}
            for name, env, kind in (("zx apply", None, "apply"), ("zx xyz", None, "xyz"),
                                    ("3-launch apply", "1", "apply"), ("3-launch xyz", "1", "xyz"),
                                    ("begin + apply", None, "begin+apply")):
                if env:
                    monkeypatch.setenv("CFD_NO_ZX", env)
                else:
                    monkeypatch.delenv("CFD_NO_ZX", raising=False)
                outs = [[torch.zeros_like(b) for _ in range(3)] for b in blocks]
                run(kind, blocks, outs)
                gz = np.concatenate([o[2].cpu().numpy() for o in outs], axis=0)
                assert relinf(gz, want[2]) <= TOL, name
                results[name] = gz
                if kind == "xyz":
                    for a in (0, 1):
                        ga = np.concatenate([o[a].cpu().numpy() for o in outs], axis=0)
                        assert relinf(ga, want[a]) <= TOL
                    for r in range(P):
                        ex, ey = xy.dfdxy(blocks[r], hs[0], hs[1])
                        assert torch.equal(ex, outs[r][0]) and torch.equal(ey, outs[r][1])
            monkeypatch.delenv("CFD_NO_ZX", raising=False)
            assert np.array_equal(results["zx apply"], results["zx xyz"])
            assert np.array_equal(results["3-launch apply"], results["3-launch xyz"])
            assert np.array_equal(results["3-launch apply"], results["begin + apply"])
            assert relinf(results["zx apply"], results["3-launch apply"]) <= 1e-14
    finally:
        torch.cuda.synchronize()
        for z in zps:
            L.cfd_zpart_destroy(z)


@pytest.mark.parametrize("P,shape", [(2, (132, 12, 34)), (3, (3 * 66, 32, 40)), (4, (4 * 70, 8, 64)), (2, (2 * 200, 6, 32))])
def test_distributed_npts_one_process(C, P, shape):
    """The LANL distributed npts method on the GPU (cfd_create_npts + cfd_zpart_apply_npts: one LU of the whole line,
    u~ to the right, x~ to the left, both sweeps in one coupled pass), all P ranks in one process, one stream per
    rank: against the one-rank derivative AND against the oracle's restatement of the reference's distributed
    algorithm (python/npts.py:228-382: full phi / psi sweeps, all-gathers, prefix combinations)."""
    import ctypes
    import torch
    from compact_finite_differences_b200._lib import check, lib
    from compact_finite_differences_b200.compact import _Plan
    L = lib()
    rng = np.random.default_rng(60 + P)
    h = 0.21
    n = shape[0] // P
    lshape = (n,) + tuple(shape[1:])
    plans = [_Plan(lshape, 2, h, r, P, npts=True) for r in range(P)]
    zps = []
    for r in range(P):
        z = ctypes.c_void_p()
        check(L.cfd_zpart_create(ctypes.byref(z), plans[r].handle))
        zps.append(z)
    bufs = [L.cfd_zpart_buffer(z) for z in zps]
    for r in range(P):
        check(L.cfd_zpart_connect_ptr(zps[r], bufs[r - 1] if r > 0 else None, bufs[r + 1] if r < P - 1 else None))
    streams = [torch.cuda.Stream() for _ in range(P)]
    zz, yy, xx = smooth(shape)
    try:
        for it in range(4):                      # both parities, twice
            f = rng.random(shape) if it < 3 else 1e3 * (np.sin(zz) * np.cos(yy) + xx)
            blocks = [dev(f[r * n:(r + 1) * n]) for r in range(P)]
            outs = [torch.zeros_like(b) for b in blocks]
            torch.cuda.synchronize()
            for r in range(P):
                check(L.cfd_zpart_apply_npts(zps[r], blocks[r].data_ptr(), outs[r].data_ptr(),
                                             ctypes.c_void_p(streams[r].cuda_stream)))
            torch.cuda.synchronize()
            assert L.cfd_async_status() == 0
            got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
            assert relinf(got, O.derivative(f, 2, h)) <= TOL
            rhs_x = np.ascontiguousarray(np.moveaxis(O.rhs(f, 2, h), 0, 2))           # z lines as the oracle's x lines
            ref = np.moveaxis(O.npts_distributed_solve(rhs_x, P), 2, 0)
            assert relinf(got, ref) <= TOL
    finally:
        torch.cuda.synchronize()
        for z in zps:
            L.cfd_zpart_destroy(z)


def test_zpart_rejects_bad_arguments(C):
    import ctypes
    from compact_finite_differences_b200._lib import CFD_EINVAL, CFD_EUNSUPPORTED, lib
    L = lib()
    h = ctypes.c_void_p()
    whole = C.CompactFiniteDifferenceSolver((70, 8, 8), 0.1, 2)
    assert L.cfd_zpart_create(ctypes.byref(h), whole._plan(2, 0.1).handle) == CFD_EINVAL          # part_size 1
    thin = C.CompactFiniteDifferenceSolver((40, 8, 8), 0.1, 2, part=(0, 2))
    assert L.cfd_zpart_create(ctypes.byref(h), thin._plan(2, 0.1).handle) == CFD_EUNSUPPORTED     # < 66 planes
    xline = C.CompactFiniteDifferenceSolver((8, 8, 128), 0.1, 0, part=(0, 2))
    assert L.cfd_zpart_create(ctypes.byref(h), xline._plan(0, 0.1).handle) == CFD_EUNSUPPORTED    # not a z line
    ok = C.CompactFiniteDifferenceSolver((70, 8, 8), 0.1, 2, part=(0, 2))
    assert L.cfd_zpart_create(ctypes.byref(h), ok._plan(2, 0.1).handle) == 0
    f = dev(np.zeros((70, 8, 8)))
    assert L.cfd_zpart_apply(h, f.data_ptr(), f.data_ptr() + 8, None) == CFD_EINVAL                # not connected
    L.cfd_zpart_destroy(h)


def test_flag_wait_times_out_instead_of_trapping(C):
    """A rank that never arrives: the waiting kernel gives up after the configured time-out, the context survives and
    the next status query reports CFD_ETIMEOUT (the round-1 kernels trapped after ~10 s)."""
    import ctypes
    import torch
    from compact_finite_differences_b200._lib import CFD_ETIMEOUT, lib
    L = lib()
    C.CompactFiniteDifferenceSolver((70, 8, 8), 0.1, 2, part=(0, 2))._plan(2, 0.1)     # a partitioned plan: error word exists
    flag = torch.zeros(2, dtype=torch.int64, device="cuda")
    assert L.cfd_set_wait_timeout_ms(50) == 0
    try:
        assert L.cfd_wait_flags(flag.data_ptr(), None, 7, None) == 0
        torch.cuda.synchronize()                     # returns: no trap, no hang
        assert L.cfd_async_status() == CFD_ETIMEOUT
        assert L.cfd_async_status() == 0             # reported once
        flag[0] = 7
        assert L.cfd_wait_flags(flag.data_ptr(), None, 7, None) == 0
        torch.cuda.synchronize()
        assert L.cfd_async_status() == 0
    finally:
        L.cfd_set_wait_timeout_ms(120000)


def test_pinned_host_path(C):
    """cfd_apply_host honours `pinned`: page-locked buffers take the slab-pipelined path (x, y) and give the same
    numbers; a false promise is refused."""
    import ctypes
    import torch
    from compact_finite_differences_b200._lib import CFD_EINVAL, lib
    L = lib()
    rng = np.random.default_rng(5)
    shape = (21, 40, 64)
    f = rng.random(shape)
    fp = torch.from_numpy(f).pin_memory()
    for axis in range(3):
        op = C.CompactFiniteDifferenceSolver(shape, 0.1, axis)
        plan = op._plan(axis, 0.1)
        out = torch.empty(shape, dtype=torch.float64).pin_memory()
        assert L.cfd_apply_host(plan.handle, fp.data_ptr(), out.data_ptr(), 1) == 0
        assert relinf(out.numpy(), O.derivative(f, axis, 0.1)) <= TOL
        plain = np.empty(shape)
        assert L.cfd_apply_host(plan.handle, f.ctypes.data, plain.ctypes.data, 0) == 0
        assert np.array_equal(plain, out.numpy())
        assert L.cfd_apply_host(plan.handle, f.ctypes.data, plain.ctypes.data, 1) == CFD_EINVAL


def test_graph_replay_beside_eager_launches(C):
    """A captured launch owns its work counters: replaying the graph on one stream while thousands of eager launches
    of the same plan run on another (enough to wrap the eager counter ring) never shares a counter pair."""
    import torch
    rng = np.random.default_rng(8)
    shape = (24, 64, 96)
    f = dev(rng.random(shape))
    op = C.CompactFiniteDifferenceSolver(shape, 0.1, 1)
    want = O.derivative(f.cpu().numpy(), 1, 0.1)
    g_out, e_out = torch.empty_like(f), torch.empty_like(f)
    op(f, g_out)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s1):
        for _ in range(4):
            op(f, g_out)
    for _ in range(40):
        with torch.cuda.stream(s1):
            graph.replay()
        with torch.cuda.stream(s2):
            for _ in range(120):                 # 40 x 120 > 4096 eager launches: the ring wraps
                op(f, e_out)
    torch.cuda.synchronize()
    assert relinf(g_out.cpu().numpy(), want) <= TOL and relinf(e_out.cpu().numpy(), want) <= TOL


def test_pthomas_plan_is_reused(C):
    import torch
    rng = np.random.default_rng(2)
    n = 12
    a, b, c = rng.random(n), rng.random(n) + 2.0, rng.random(n)
    rs = C.ReducedSolver((n, 3, 5))
    for _ in range(3):
        d = rng.random((n, 3, 5))
        t = dev(d)
        rs.solve(a, b, c, None, t)
        np.testing.assert_allclose(t.cpu().numpy(), O.pthomas(a, b, c, d), rtol=1e-12)
    assert len(rs._plans) == 1
    rs.solve(a, b + 1.0, c, None, dev(rng.random((n, 3, 5))))
    assert len(rs._plans) == 2


def test_no_out_of_bounds_writes(C):
    """compute-sanitizer is closed on the GPU pool, so memory safety of the stores gets a check of its own: every output
    of every kernel family lives in the middle of a larger allocation whose guard zones (4 KiB on each side, plus the
    tail of ragged tiles inside TMA boxes that cross the end of the tensor) must come back untouched, on ragged shapes
    where boxes / thread blocks overhang every extent."""
    import torch
    G = 512                                                    # guard doubles on each side (16-byte aligned offsets)
    SENT = 1234.5

    class Guarded:
        def __init__(self, shape, fill=None):
            n = int(np.prod(shape))
            self.raw = torch.full((n + 2 * G,), SENT, dtype=torch.float64, device="cuda")
            self.t = self.raw[G:G + n].view(shape)
            if fill is not None:
                self.t.copy_(torch.from_numpy(np.ascontiguousarray(fill)))

        def intact(self):
            return bool((self.raw[:G] == SENT).all() and (self.raw[-G:] == SENT).all())

    rng = np.random.default_rng(123)
    held = []

    def out(shape, fill=None):
        g = Guarded(shape, fill)
        held.append(g)
        return g.t

    for shape in [(5, 7, 34), (4, 33, 66), (35, 6, 38), (4, 64, 98)]:
        f = rng.random(shape)
        fd = dev(f)
        for axis in range(3):
            o = out(shape)
            C.CompactFiniteDifferenceSolver(shape, 0.1, axis)(fd, o)
            assert relinf(o.cpu().numpy(), O.derivative(f, axis, 0.1)) <= TOL
            t = out(shape, f)
            C.NearToeplitzSolver(shape, O.PADE, axis=axis).solve(t)                       # in place
            t2 = out(shape, f)
            C.NearToeplitzSolver(shape, (1., 2., 1. / 3, 1., 1. / 3, 2., 1.), axis=axis).solve(t2)   # general kernel
            for scheme in ("compact6", "pade4-d2"):
                if shape[2 - axis] >= 6:
                    o2 = out(shape)
                    C.CompactFiniteDifferenceSolver(shape, 0.2, axis, scheme=scheme)(fd, o2)
        if shape[1] % 32 == 0:
            gx, gy, gz = out(shape), out(shape), out(shape)
            C.CompactFiniteDifferenceSolver(shape).gradient(fd, (0.1, 0.2, 0.3), (gx, gy, gz))
    # starved in-place solver (side buffer + scatter), contiguous and strided
    for shape, axis in (((1, 40, 1000), 0), ((545, 2, 34), 2)):
        d = rng.random(shape)
        t = out(shape, d)
        C.NearToeplitzSolver(shape, O.PADE, axis=axis).solve(t)
        assert relinf(t.cpu().numpy(), O.scipy_solve_axis(d, O.PADE, axis)) <= TOL
    # multi-rank pieces with their plane-sized outputs
    P, shape, h = 2, (140, 6, 34), 0.2
    f = rng.random(shape)
    n = shape[0] // P
    plane = shape[1] * shape[2]
    blocks = [dev(f[r * n:(r + 1) * n]) for r in range(P)]
    sol = [C.CompactFiniteDifferenceSolver((n,) + shape[1:], h, 2, part=(r, P)) for r in range(P)]
    halos = [(None if r == 0 else blocks[r - 1][-1].contiguous(), None if r == P - 1 else blocks[r + 1][0].contiguous())
             for r in range(P)]
    faces = out((2 * P, plane), np.zeros((2 * P, plane)))
    ab = out((2, plane))
    res = []
    for r in range(P):
        sol[r].edge_faces(blocks[r], faces[2 * r:2 * r + 2], *halos[r])
    for r in range(P):
        sol[r].reduced_unknowns(faces, ab)
        o = out((n,) + shape[1:])
        sol[r].apply_coupled(blocks[r], o, halos[r][0], halos[r][1], ab)
        res.append(o)
    assert relinf(torch.cat(res).cpu().numpy(), O.derivative(f, 2, h)) <= TOL
    res = []
    faces2 = out((2 * P, plane), np.zeros((2 * P, plane)))
    for r in range(P):
        o = out((n,) + shape[1:])
        sol[r].apply_local(blocks[r], o, *halos[r])
        sol[r].interface_pack(o, faces2[2 * r:2 * r + 2])
        res.append(o)
    for r in range(P):
        sol[r].reduced_correct(res[r], faces2)
        rhs = out((n,) + shape[1:])
        sol[r].compute_RHS(blocks[r], h, rhs, None, halo_lo=halos[r][0], halo_hi=halos[r][1])
    assert relinf(torch.cat(res).cpu().numpy(), O.derivative(f, 2, h)) <= TOL
    a, b, c = rng.random(9), rng.random(9) + 2, rng.random(9)
    pt = out((9, 3, 5), rng.random((9, 3, 5)))
    C.ReducedSolver((9, 3, 5)).solve(a, b, c, None, pt)
    torch.cuda.synchronize()
    assert all(g.intact() for g in held), "a kernel wrote outside its output tensor"


def test_host_gradient_pipeline(C):
    """HostGradient (pinned host buffers, slab-pipelined copies) == oracle on every direction."""
    import torch
    rng = np.random.default_rng(21)
    shape = (37, 48, 64)
    f = rng.random(shape)
    hs = (0.11, 0.07, 0.05)
    hg = C.HostGradient(shape, hs, slabs=5)
    fh = torch.from_numpy(f).pin_memory()
    for _ in range(2):
        dx, dy, dz = hg(fh)
        for got, axis, h in ((dx, 0, hs[0]), (dy, 1, hs[1]), (dz, 2, hs[2])):
            assert relinf(got.numpy(), O.derivative(f, axis, h)) <= TOL
    outs = [np.empty(shape) for _ in range(3)]
    hg(f, outs)                                   # NumPy in, NumPy out
    assert relinf(outs[1], O.derivative(f, 1, hs[1])) <= TOL


@pytest.mark.parametrize("kseg", [1, 2, 8])
def test_segmented_lines(C, kseg):
    """Forced line segmentation (cfd_set_segments): every axis, derivative and solver, ragged sizes."""
    rng = np.random.default_rng(kseg)
    try:
        C.lib().cfd_set_segments(kseg)
        for shape in [(6, 10, 400), (300, 6, 34), (5, 530, 36)]:
            f = rng.random(shape)
            for axis in range(3):
                if shape[2 - axis] < 4:
                    continue
                got = C.CompactFiniteDifferenceSolver(shape, 0.3, axis)(dev(f)).cpu().numpy()
                assert relinf(got, O.derivative(f, axis, 0.3)) <= TOL
                t = dev(f)
                C.NearToeplitzSolver(shape, O.PADE, axis=axis).solve(t)
                assert relinf(t.cpu().numpy(), O.near_toeplitz_solve(f, O.PADE, axis)) <= TOL
    finally:
        C.lib().cfd_set_segments(0)


def test_auto_segmentation_small_batch(C):
    """32 bundles of 4096-long lines: the launcher cuts the lines automatically (the in-place solver through its side
    buffer, test_in_place_solver_on_starved_shapes); results unchanged."""
    rng = np.random.default_rng(5)
    d = rng.random((1, 1024, 4096))
    t = dev(d)
    C.NearToeplitzSolver(d.shape, O.PADE).solve(t)
    assert relinf(t.cpu().numpy(), O.near_toeplitz_solve(d, O.PADE)) <= TOL
    got = C.CompactFiniteDifferenceSolver(d.shape, 0.2, 0)(dev(d)).cpu().numpy()
    assert relinf(got, O.derivative(d, 0, 0.2)) <= TOL


@pytest.mark.parametrize("shape,axis", [((1, 64, 4096), 0), ((1, 1024, 4096), 0), ((2, 40, 1000), 0), ((2, 2048, 64), 1),
                                        ((1024, 4, 64), 2), ((545, 2, 96), 2), ((3, 100, 8192), 0)])
def test_in_place_solver_on_starved_shapes(C, shape, axis):
    """Long lines, few bundles: the solver cuts the lines into segments IN PLACE -- each segment's first and last
    result chunk (its neighbours' look-ahead / warm-up inputs) take a detour through a side buffer of 2 chunks per
    segment and a small scatter launch -- instead of round 1's field-sized scratch + device copy."""
    rng = np.random.default_rng(sum(shape) + axis)
    d = rng.random(shape) - 0.5
    for co in (O.PADE, (1.5, 0.3, 0.2, 1.0, 0.25, 0.1, 2.0)):
        s = C.NearToeplitzSolver(shape, co, axis=axis)
        t = dev(d)
        n0 = C.lib().cfd_launch_count()
        s.solve(t)
        launches = C.lib().cfd_launch_count() - n0
        assert relinf(t.cpu().numpy(), O.scipy_solve_axis(d, co, axis)) <= TOL
        assert launches == 2                   # segmented kernel + scatter (one launch would mean whole lines)
        s.solve(t)                             # and again on its own output: the side buffer is re-usable
        assert relinf(t.cpu().numpy(), O.scipy_solve_axis(O.scipy_solve_axis(d, co, axis), co, axis)) <= 10 * TOL


def test_very_long_lines(C):
    """n = 16384 (512 chunks per line; the reference caps nx at 2048): derivative and solver vs the oracle."""
    rng = np.random.default_rng(16384)
    f = rng.random((2, 40, 16384))
    got = C.CompactFiniteDifferenceSolver(f.shape, 0.01, 0)(dev(f)).cpu().numpy()
    assert relinf(got, O.derivative(f, 0, 0.01)) <= TOL
    t = dev(f)
    C.NearToeplitzSolver(f.shape, O.PADE).solve(t)
    assert relinf(t.cpu().numpy(), O.near_toeplitz_solve(f, O.PADE)) <= TOL
    g = np.ascontiguousarray(f.transpose(2, 1, 0))            # the same lines along z
    got = C.CompactFiniteDifferenceSolver(g.shape, 0.01, 2)(dev(g)).cpu().numpy()
    assert relinf(got, O.derivative(g, 2, 0.01)) <= TOL


def test_maximum_size_round_trip(C):
    """2^30 unknowns (8 GiB, the cap of BASELINE configs[4]): A x == d in place, checked on the device."""
    import torch
    n, batch = 1024, 2 ** 20
    d = torch.rand((1, batch, n), dtype=torch.float64, device="cuda")
    chk = d[0, ::4099].clone()                                 # keep a sample of the right-hand sides
    C.NearToeplitzSolver((1, batch, n), O.PADE).solve(d)
    x = d[0, ::4099]
    ax = x.clone()
    ax[:, 1:-1] = 0.25 * x[:, :-2] + x[:, 1:-1] + 0.25 * x[:, 2:]
    ax[:, 0] = x[:, 0] + 2 * x[:, 1]
    ax[:, -1] = 2 * x[:, -2] + x[:, -1]
    assert (ax - chk).abs().max().item() < 1e-13
    lines = x[:8].cpu().numpy()
    want = O.near_toeplitz_solve(chk[:8].cpu().numpy().reshape(1, 8, n), O.PADE).reshape(8, n)
    assert relinf(lines, want) <= TOL


def test_reentrant_on_two_streams(C):
    """Plans are immutable and every launch takes its own work-counter pair: concurrent calls on distinct streams
    (same plan, different fields) must not disturb each other."""
    import torch
    rng = np.random.default_rng(77)
    shape = (96, 128, 160)
    fa, fb = rng.random(shape), rng.random(shape)
    da, db = dev(fa), dev(fb)
    oa, ob = torch.empty_like(da), torch.empty_like(db)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for axis in range(3):
        op = C.CompactFiniteDifferenceSolver(shape, 0.1, axis)
        op2 = C.CompactFiniteDifferenceSolver(shape, 0.1, axis)
        for _ in range(5):
            with torch.cuda.stream(s1):
                op(da, oa)
            with torch.cuda.stream(s2):
                op2(db, ob)
        torch.cuda.synchronize()
        assert relinf(oa.cpu().numpy(), O.derivative(fa, axis, 0.1)) <= TOL
        assert relinf(ob.cpu().numpy(), O.derivative(fb, axis, 0.1)) <= TOL


@pytest.mark.parametrize("shape", [(4, 4, 4), (1, 1, 4), (2, 5, 6), (1, 4, 2), (4, 1, 2)])
def test_tiny_shapes(C, shape):
    rng = np.random.default_rng(sum(shape))
    f = rng.random(shape)
    for axis in range(3):
        if shape[2 - axis] < 4:
            continue
        got = C.CompactFiniteDifferenceSolver(shape, 0.5, axis)(dev(f)).cpu().numpy()
        assert relinf(got, O.derivative(f, axis, 0.5)) <= TOL


def test_error_codes(C):
    """Shape / argument errors come back as exceptions carrying the C error text, never as a crash."""
    import torch
    with pytest.raises(C.CfdError):
        C.CompactFiniteDifferenceSolver((8, 8, 7), 0.1, 0)            # nx odd
    with pytest.raises(C.CfdError):
        C.CompactFiniteDifferenceSolver((8, 8, 8), -1.0, 0)           # spacing
    with pytest.raises(C.CfdError):
        C.CompactFiniteDifferenceSolver((2, 8, 8), 0.1, 2)            # fewer than 3 rows along the axis
    with pytest.raises(C.CfdError):
        C.CompactFiniteDifferenceSolver((3, 8, 8), 0.1, 2)            # the 3-row closure matrix is singular
    with pytest.raises(C.CfdError):
        C.NearToeplitzSolver((4, 4, 8), (0., 1., 1., 1., 1., 1., 1.))  # zero pivot
    s = C.CompactFiniteDifferenceSolver((8, 8, 8), 0.1, 0)
    f = torch.zeros((8, 8, 8), dtype=torch.float64, device="cuda")
    with pytest.raises(AssertionError):
        s(f, out=f)                                                   # in-place derivative
    with pytest.raises(AssertionError):
        s(f.float())


def test_reference_call_shape(C):
    """The reference's own call sequence (code/cuda/test/test_compact.py:19-31): a line_da, dfdx(f_d, dx, x_d,
    f_local_d) filling x_d; decimal=2 against cos(x), plus oracle parity."""
    import torch
    shape = (32, 32, 32)
    x, y, z = smooth(shape)
    line_da = C.LineDA(shape)                              # one rank along the line
    cfd = C.CompactFiniteDifferenceSolver(line_da, solver='templated')
    f = np.sin(x)
    f_d = dev(f)
    x_d = torch.empty_like(f_d)
    f_local_d = torch.empty((34, 34, 34), dtype=torch.float64, device="cuda")     # the reference's ghosted scratch
    dx = x[0, 0, 1] - x[0, 0, 0]
    cfd.dfdx(f_d, dx, x_d, f_local_d)
    np.testing.assert_almost_equal(np.cos(x), x_d.cpu().numpy(), decimal=2)
    assert relinf(x_d.cpu().numpy(), O.derivative(f, 0, dx)) <= TOL


def test_cuda_graph_capture(C):
    """apply() makes no allocation and no synchronisation, so a gradient step can be captured into a CUDA graph
    (small grids are launch-bound: 64^3 is ~5 us of kernel per derivative); replays must reproduce the result."""
    import torch
    n = 64
    rng = np.random.default_rng(64)
    f = rng.random((n, n, n))
    fd = dev(f)
    outs = [torch.empty_like(fd) for _ in range(3)]
    ops = [C.CompactFiniteDifferenceSolver((n, n, n), 0.1, a) for a in range(3)]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for a in range(3):                       # warm-up outside capture (first-call attribute set-up)
            ops[a](fd, outs[a])
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for a in range(3):
            ops[a](fd, outs[a])
    for rep in range(3):
        fd.copy_(dev(f * (rep + 1)))
        for o in outs:
            o.zero_()
        g.replay()
        torch.cuda.synchronize()
        for a in range(3):
            assert relinf(outs[a].cpu().numpy(), O.derivative(f * (rep + 1), a, 0.1)) <= TOL


def test_cuda_graph_capture_gradient(C):
    """The same for gradient(): the fused d/dx + d/dy launch draws its work order from a table built at plan
    creation, so the first call inside a capture allocates nothing either."""
    import torch
    shape = (40, 64, 96)
    rng = np.random.default_rng(65)
    f = rng.random(shape)
    fd = dev(f)
    outs = [torch.empty_like(fd) for _ in range(3)]
    sol = C.CompactFiniteDifferenceSolver(shape)
    hs = (0.1, 0.2, 0.3)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        sol.gradient(fd, hs, outs)               # warm-up outside capture (first-call attribute set-up)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        sol.gradient(fd, hs, outs)
    for rep in range(2):
        fd.copy_(dev(f * (rep + 2)))
        for o in outs:
            o.zero_()
        g.replay()
        torch.cuda.synchronize()
        for a in range(3):
            assert relinf(outs[a].cpu().numpy(), O.derivative(f * (rep + 2), a, hs[a])) <= TOL


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_reference_stages_one_by_one(C, axis):
    """The reference's dfdx body (code/cuda/compact.py:40-44) stage by stage on a 3-rank line, every stage checked
    against the oracle's restatement of that stage: compute_RHS, solve_primary_system, solve_secondary_systems,
    the reduced system, sum_solutions."""
    import torch
    P, h = 3, 0.23
    shape = [10, 12, 14]
    shape[2 - axis] = 3 * 40
    rng = np.random.default_rng(axis + 30)
    f = rng.random(shape)
    ax = 2 - axis
    n = shape[ax] // P
    blocks = [np.ascontiguousarray(np.take(f, range(r * n, (r + 1) * n), axis=ax)) for r in range(P)]
    plane = blocks[0].size // n
    faces = torch.zeros((2 * P, plane), dtype=torch.float64, device="cuda")
    xr, cfds = [], []
    for r in range(P):
        lo = np.take(blocks[r - 1], n - 1, axis=ax) if r > 0 else None
        hi = np.take(blocks[r + 1], 0, axis=ax) if r < P - 1 else None
        cfd = C.CompactFiniteDifferenceSolver(C.LineDA(blocks[r].shape, r, P, axis), solver="templated")
        cfd.spacing = h
        x_d = cfd.compute_RHS(dev(blocks[r]), h, None, None, None if lo is None else dev(lo), None if hi is None else dev(hi))
        want_rhs = O.rhs(blocks[r], axis, h, halo_lo=lo, halo_hi=hi)
        assert relinf(x_d.cpu().numpy(), want_rhs) <= 1e-15                               # a1 computeRHS
        cfd.solve_primary_system(x_d)
        want_xr = O.scipy_solve_axis(want_rhs, O.partition_local_coeffs(r, P), axis)
        assert relinf(x_d.cpu().numpy(), want_xr) <= TOL                                  # a3 primary solve
        xu, xl = cfd.solve_secondary_systems()
        xu_w, xl_w = O.partition_secondary(n, r, P)
        np.testing.assert_allclose(xu.cpu().numpy(), xu_w, rtol=1e-13, atol=1e-300)       # a4 secondary systems
        np.testing.assert_allclose(xl.cpu().numpy(), xl_w, rtol=1e-13, atol=1e-300)
        cfd.interface_pack(x_d, faces[2 * r:2 * r + 2])                                   # a5 negateAndCopyFaces
        xr.append(x_d)
        cfds.append(cfd)
    ra, rb, rc = O.partition_reduced_matrix(n, P)
    sol = O.scipy_solve_banded(ra, rb, rc, faces.cpu().numpy())
    ab = torch.empty((2, plane), dtype=torch.float64, device="cuda")
    for r in range(P):
        cfds[r].reduced_unknowns(faces, ab)                                               # a6 reduced system
        np.testing.assert_allclose(ab.cpu().numpy(), sol[2 * r:2 * r + 2], rtol=1e-12, atol=1e-14)
        cfds[r].sum_solutions(xr[r], ab[0], ab[1])                                        # a7 sumSolutions
    got = np.concatenate([x.cpu().numpy() for x in xr], axis=ax)
    assert relinf(got, O.derivative(f, axis, h)) <= TOL                                   # a9 the whole dfdx
