"""CPU-side checks of the product: C-ABI surface, host tables, chunk schedule, loud failure without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import cfd_oracle as O
from tests.conftest import ROOT
from tests.emulate import PADE, edge_faces, stream_lines, tables


def relinf(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_library_exports_every_declared_symbol():
    from compact_finite_differences_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cfd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:cfd|nt)_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/cfd_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert _lib.lib().cfd_version() == 200


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import compact_finite_differences_b200 as C
    with pytest.raises(C.CfdError):
        C.CompactFiniteDifferenceSolver((8, 8, 8), 0.1, 0)
    with pytest.raises(C.CfdError):
        C.NearToeplitzSolver((8, 8, 8), PADE)


def test_product_does_not_use_oracle():
    """The product has no CPU path: nothing under the package may import, link or load oracle/."""
    pkg = os.path.join(ROOT, "compact_finite_differences_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                for bad in ("import oracle", "from oracle", "liboracle", "cfd_oracle", "scipy"):
                    assert bad not in src, f"{fn} references {bad!r}"


@pytest.mark.parametrize("n", [4, 5, 31, 32, 33, 63, 64, 65, 96, 100, 128, 257, 512, 1000, 2048])
def test_chunk_schedule_derivative(n):
    """Kernel schedule (HEAD/MID/TAIL tables, 32-row look-ahead) == oracle, for ragged and full chunk counts."""
    rng = np.random.default_rng(n)
    F = rng.random((4, n))
    h = 0.37
    want = O.derivative(F.reshape(1, 4, n), 0, h).reshape(4, n)
    assert relinf(stream_lines(F, PADE, h), want) < 1e-14


@pytest.mark.parametrize("n", [4, 32, 64, 65, 128, 4096])
def test_chunk_schedule_solve(n):
    rng = np.random.default_rng(n)
    F = rng.random((3, n))
    want = O.near_toeplitz_solve(F.reshape(1, 3, n), PADE).reshape(3, n)
    assert relinf(stream_lines(F, PADE), want) < 1e-14


def test_general_coefficients_small_and_rejected_large():
    co = (1., 2., 3., 4., 5., 6., 7.)
    rng = np.random.default_rng(7)
    for n in (32, 64):          # <= 2 chunks: exact for any matrix
        F = rng.random((2, n))
        want = O.near_toeplitz_solve(F.reshape(1, 2, n), co).reshape(2, n)
        assert relinf(stream_lines(F, co), want) < 1e-12
        assert tables(n, co, 1.0)["fast_ok"]
    assert not tables(128, co, 1.0)["fast_ok"]       # not diagonally dominant: streaming path must refuse
    assert tables(128, PADE, 1.0)["fast_ok"]
    assert tables(4096, (1., .25, .25, 1., .25, .25, 1.), 1.0)["fast_ok"]


@pytest.mark.parametrize("rank,size", [(0, 2), (1, 2), (0, 4), (2, 4), (3, 4)])
def test_block_local_solve_with_halo(rank, size):
    """Block-local x_R of a partitioned line: cut Toeplitz ends + halo points in the RHS."""
    n = 64
    rng = np.random.default_rng(rank * 10 + size)
    F = rng.random((3, size * n))
    h = 0.2
    blk = F[:, rank * n:(rank + 1) * n]
    lo = None if rank == 0 else F[:, rank * n - 1]
    hi = None if rank == size - 1 else F[:, (rank + 1) * n]
    co = O.partition_local_coeffs(rank, size)
    got = stream_lines(blk, co, h, lo_closure=rank == 0, hi_closure=rank == size - 1, halo_lo=lo, halo_hi=hi)
    rr = O.rhs(blk.reshape(1, 3, n), 0, h, halo_lo=lo, halo_hi=hi).reshape(3, n)
    a, b, c = O.banded_abc(n, co)
    assert relinf(got, O.scipy_solve_banded(a, b, c, rr.T).T) < 1e-14


@pytest.mark.parametrize("n,size", [(8, 2), (32, 4), (128, 8)])
def test_reduced_system_tables(n, size):
    """Secondary solutions and the 2P x 2P interface matrix built by the library == reference algebra."""
    from compact_finite_differences_b200._lib import check, lib
    dp = ctypes.POINTER(ctypes.c_double)
    ra_w, rb_w, rc_w = O.partition_reduced_matrix(n, size)
    for rank in range(size):
        xu, xl = np.zeros(n), np.zeros(n)
        ra, rb, rc = np.zeros(2 * size), np.zeros(2 * size), np.zeros(2 * size)
        check(lib().cfd_debug_secondary(n, rank, size, *(v.ctypes.data_as(dp) for v in (xu, xl, ra, rb, rc))))
        xu_w, xl_w = O.partition_secondary(n, rank, size)
        np.testing.assert_allclose(xu, xu_w, rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(xl, xl_w, rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(ra, ra_w, rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(rb, rb_w, rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(rc, rc_w, rtol=1e-13, atol=1e-300)


@pytest.mark.parametrize("n,size", [(66, 2), (96, 3), (128, 4), (100, 2)])
def test_fused_multirank_schedule(n, size):
    """cfd_edge_faces + all-gather + cfd_apply_coupled (emulated with the library's tables) == one-rank derivative:
    interface planes from the block ends only, reduced solve, interface unknowns folded into rows 0 / n-1."""
    from compact_finite_differences_b200._lib import check, lib
    dp = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(n + size)
    nl = 5
    F = rng.random((nl, size * n))
    h = 0.21
    want = O.derivative(F.reshape(1, nl, size * n), 0, h).reshape(nl, size * n)
    faces = np.zeros((2 * size, nl))
    blocks = []
    for r in range(size):
        blk = F[:, r * n:(r + 1) * n]
        lo = None if r == 0 else F[:, r * n - 1]
        hi = None if r == size - 1 else F[:, (r + 1) * n]
        co = O.partition_local_coeffs(r, size)
        blocks.append((blk, co, lo, hi))
        faces[2 * r], faces[2 * r + 1] = edge_faces(blk, co, h, r == 0, r == size - 1, lo, hi)
        # the edge values equal the faces of the full block-local solve
        xr = stream_lines(blk, co, h, lo_closure=r == 0, hi_closure=r == size - 1, halo_lo=lo, halo_hi=hi)
        if r > 0:
            np.testing.assert_allclose(faces[2 * r], -xr[:, 0], rtol=1e-13, atol=1e-15)
        if r < size - 1:
            np.testing.assert_allclose(faces[2 * r + 1], -xr[:, -1], rtol=1e-13, atol=1e-15)
    ra, rb, rc = (np.zeros(2 * size) for _ in range(3))
    check(lib().cfd_debug_secondary(n, 0, size, None, None, *(v.ctypes.data_as(dp) for v in (ra, rb, rc))))
    sol = O.scipy_solve_banded(ra, rb, rc, faces)
    for r in range(size):
        blk, co, lo, hi = blocks[r]
        got = stream_lines(blk, co, h, lo_closure=r == 0, hi_closure=r == size - 1, halo_lo=lo, halo_hi=hi,
                           alpha=sol[2 * r], beta=sol[2 * r + 1])
        assert relinf(got, want[:, r * n:(r + 1) * n]) < 1e-13


@pytest.mark.parametrize("n,size", [(66, 2), (64, 3), (128, 4), (96, 8)])
def test_neighbour_only_reduced_system(n, size):
    """Neighbour-only interface system (ranks r-1, r, r+1) gives the same two unknowns as the full 2P system."""
    from compact_finite_differences_b200._lib import check, lib
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    rng = np.random.default_rng(size)
    faces = rng.standard_normal((2 * size, 7))
    faces[0] = 0.0
    faces[-1] = 0.0
    ra, rb, rc = O.partition_reduced_matrix(n, size)
    full = O.scipy_solve_banded(ra, rb, rc, faces)
    for r in range(size):
        pv, own = ctypes.c_int(), ctypes.c_int()
        va, vb, vc = (np.zeros(6) for _ in range(3))
        check(lib().cfd_debug_neighbour(n, r, size, ctypes.byref(pv), ctypes.byref(own),
                                        *(v.ctypes.data_as(dp) for v in (va, vb, vc))))
        pv, own = pv.value, own.value
        lo = r - own
        assert pv == (3 if 0 < r < size - 1 else 2) and 0 <= lo and lo + pv <= size
        fv = faces[2 * lo:2 * (lo + pv)].copy()
        fv[0] = 0.0
        fv[-1] = 0.0
        sol = O.scipy_solve_banded(va[:2 * pv], vb[:2 * pv], vc[:2 * pv], fv)
        np.testing.assert_allclose(sol[2 * own:2 * own + 2], full[2 * r:2 * r + 2], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("n,kseg", [(200, 2), (256, 1), (512, 8), (1000, 8), (4096, 16), (97, 1), (64, 1)])
def test_segmented_schedule(n, kseg):
    """Lines cut into segments (work items of the streaming kernel for long lines / few bundles): a 32-row forward
    warm-up before and the usual look-ahead chunk after every cut keep the result exact to fp64."""
    rng = np.random.default_rng(n + kseg)
    F = rng.random((3, n))
    h = 0.4
    want = O.derivative(F.reshape(1, 3, n), 0, h).reshape(3, n)
    got = stream_lines(F, PADE, h, kseg=kseg)
    assert not np.isnan(got).any()
    assert relinf(got, want) < 1e-14
    want2 = O.near_toeplitz_solve(F.reshape(1, 3, n), PADE).reshape(3, n)
    assert relinf(stream_lines(F, PADE, kseg=kseg), want2) < 1e-14


def test_chunk_schedule_random_property():
    """Randomised sweep over line lengths, segmentations and block positions: the kernel schedule (emulated with the
    library's tables) stays within 1e-13 of the oracle."""
    rng = np.random.default_rng(2026)
    for _ in range(40):
        n = int(rng.integers(4, 700))
        kseg = int(rng.integers(0, 6))
        F = rng.standard_normal((2, n))
        h = float(rng.uniform(0.01, 2.0))
        want = O.derivative(F.reshape(1, 2, n), 0, h).reshape(2, n)
        got = stream_lines(F, PADE, h, kseg=kseg or None)
        assert relinf(got, want) < 1e-13, (n, kseg)


# ---------------------------------------------------------------------------------------------------
# cfd_apply_xy: the wavefront draw order (host table)
# ---------------------------------------------------------------------------------------------------
def _xy_order(nz, nxp, nyp, active, sub=0):
    from compact_finite_differences_b200._lib import lib
    n = lib().cfd_debug_xy_order(nz, nxp, nyp, float(active), sub, None, 0)
    assert n >= nz * (nxp + nyp)
    out = np.zeros(n, dtype=np.int32)
    assert lib().cfd_debug_xy_order(nz, nxp, nyp, float(active), sub, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n) == n
    return out >> 3, out & 7


@pytest.mark.parametrize("nz,nxp,nyp,active", [(1, 1, 1, 3.0), (7, 16, 16, 18.5), (40, 8, 8, 37.0), (5, 2, 32, 4.0),
                                               (9, 32, 1, 1.0), (33, 4, 5, 100.0), (12, 16, 16, 0.0), (3, 0, 4, 2.0)])
def test_xy_draw_order_is_a_wavefront(nz, nxp, nyp, active):
    """Every (plane, bundle) item exactly once; inside a plane the x and y bundles come in ascending order (bundle j
    starts j tile-times after bundle 0, so both readers of tile (j, k) reach it j + k tile-times in); bundle j of the
    two directions is drawn back to back; planes start in ascending order, `active` of them in flight."""
    order, seg = _xy_order(nz, nxp, nyp, active)
    ipp = nxp + nyp
    assert sorted(order.tolist()) == list(range(nz * ipp)) and not seg.any()
    if active == 0.0:
        assert order.tolist() == list(range(nz * ipp))
        return
    pos = np.empty(nz * ipp, dtype=np.int64)
    pos[order] = np.arange(nz * ipp)
    first = []
    for z in range(nz):
        px, py = pos[z * ipp:z * ipp + nxp], pos[z * ipp + nxp:(z + 1) * ipp]
        assert np.all(np.diff(px) > 0) and np.all(np.diff(py) > 0)
        for j in range(min(nxp, nyp)):
            assert py[j] == px[j] + 1                       # the pair (x_j, y_j) is adjacent in the draw
        first.append(min(px.min() if nxp else 1 << 60, py.min() if nyp else 1 << 60))
    assert np.all(np.diff(first) > 0)
    # planes in flight: between the first and the last draw of a plane, no more than ~active other planes start
    M = max(nxp, nyp)
    if nz > 2 * active + 2 and M > 1:
        z = nz // 2
        last = pos[z * ipp:(z + 1) * ipp].max()
        started = sum(1 for zz in range(z + 1, nz) if first[zz] < last)
        assert started <= int(np.ceil(active)) + 1


@pytest.mark.parametrize("nz,nxp,nyp,sub", [(3, 32, 32, 16), (2, 32, 8, 16), (2, 5, 40, 4), (1, 64, 64, 4), (4, 7, 7, 2),
                                            (2, 16, 16, 16), (2, 33, 17, 16), (3, 32, 32, 16 | (32 << 8)),
                                            (2, 32, 32, 32 | (16 << 8)), (2, 32, 32, 8 | (16 << 8))])
def test_xy_draw_order_sub_planes(nz, nxp, nyp, sub):
    """Sub-plane wavefronts: lines longer than `sub` tiles are cut into segments; every line has each of its
    segments exactly once, the segments tile the line, and an x segment is drawn next to the y segments that share
    its square of tiles (same plane, x bundle inside the y segment's chunk range and vice versa)."""
    order, seg = _xy_order(nz, nxp, nyp, 5.0, sub)
    ipp = nxp + nyp
    subx, suby = sub & 255, (sub >> 8) or (sub & 255)           # segment length of the x lines / of the y lines
    ksx = subx if subx * 7 >= nyp else (nyp + 6) // 7            # an x line is nyp tiles long; at most 7 segments
    ksy = suby if suby * 7 >= nxp else (nxp + 6) // 7
    ngx = -(-nyp // ksx) if nyp > ksx else 1                     # segments of an x line
    ngy = -(-nxp // ksy) if nxp > ksy else 1
    cut = ngx > 1 or ngy > 1
    assert seg.max() <= 7
    seen = {}
    for e, s in zip(order.tolist(), seg.tolist()):
        seen.setdefault(e, []).append(s)
    assert sorted(seen) == list(range(nz * ipp))
    for e, segs in seen.items():
        is_x = (e % ipp) < nxp
        ng = ngx if is_x else ngy
        assert sorted(segs) == ([0] if ng == 1 else list(range(1, ng + 1))), (e, segs)
    if not cut:
        return
    # neighbours in the draw belong to the same square: an (x, y) pair drawn back to back shares tiles
    sy, sx = (ksy if ngy > 1 else nxp), (ksx if ngx > 1 else nyp)
    pairs = 0
    for i in range(len(order) - 1):
        e0, e1 = order[i], order[i + 1]
        if e0 // ipp != e1 // ipp:
            continue
        r0, r1 = e0 % ipp, e1 % ipp
        if r0 < nxp <= r1:                                    # x item followed by a y item of the same plane
            j, k = r0, r1 - nxp
            b = seg[i] - 1 if ngx > 1 else 0                  # x segment index = x block of the square
            a = seg[i + 1] - 1 if ngy > 1 else 0              # y segment index = y block of the square
            if j // sy == a and k // sx == b:
                pairs += 1
    assert pairs >= nz * min(nxp, nyp) // 2
    # skewed starts: segment b + 1 of a line is drawn about one square (a square = ~2 * ks draws of its own, other
    # squares in between) after segment b, so that the tile the two share is asked for by both within a tile-time
    pos = {}
    for i, (e, sg) in enumerate(zip(order.tolist(), seg.tolist())):
        pos[(e, sg)] = i
    for (e, sg), i in pos.items():
        if sg >= 1 and (e, sg + 1) in pos:
            assert pos[(e, sg + 1)] - i >= (ksx if (e % ipp) < nxp else ksy), (e, sg)


# ---------------------------------------------------------------------------------------------------
# one-launch exchange: the faces are linear in the neighbour points of f, with the library's two weights
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,h", [(66, 0.2), (128, 0.013)])
def test_halo_weights_of_the_deferred_exchange(n, h):
    """cfd_edge_faces_push computes -x_R[0], -x_R[n-1] with guessed neighbour points and
    cfd_reduced_unknowns_deferred adds w * (halo - guess): check w_lo, w_hi against block-local solves by the oracle
    (interior block, and the end blocks' interior-type ends), and that the corrected faces equal the direct ones."""
    from compact_finite_differences_b200._lib import check, lib
    w_lo, w_hi = ctypes.c_double(), ctypes.c_double()
    check(lib().cfd_debug_halo_weights(n, h, ctypes.byref(w_lo), ctypes.byref(w_hi)))
    w_lo, w_hi = w_lo.value, w_hi.value
    rng = np.random.default_rng(n)
    blk = rng.random((4, n))
    for rank, size in ((1, 3), (0, 2), (1, 2)):
        co = O.partition_local_coeffs(rank, size)
        a, b, c = O.banded_abc(n, co)

        def faces(lo, hi):
            rr = O.rhs(blk.reshape(1, 4, n), 0, h, halo_lo=lo, halo_hi=hi).reshape(4, n)
            x = O.scipy_solve_banded(a, b, c, rr.T).T
            return -x[:, 0], -x[:, -1]

        lo_true = rng.random(4) if rank > 0 else None
        hi_true = rng.random(4) if rank < size - 1 else None
        guess_lo = blk[:, 0] if rank > 0 else None
        guess_hi = blk[:, -1] if rank < size - 1 else None
        f_lo, f_hi = faces(lo_true, hi_true)
        g_lo, g_hi = faces(guess_lo, guess_hi)
        scale = max(np.abs(f_lo).max(), np.abs(f_hi).max())
        if rank > 0:
            assert np.abs(g_lo + w_lo * (lo_true - guess_lo) - f_lo).max() <= 1e-13 * scale
        if rank < size - 1:
            assert np.abs(g_hi + w_hi * (hi_true - guess_hi) - f_hi).max() <= 1e-13 * scale


def test_lookahead_from_coefficients_host():
    """ceil(log(1.2e-16) / log|g| / 32) chunks of look-ahead, from the matrix alone (no device needed)."""
    from compact_finite_differences_b200._lib import lib
    dp = ctypes.POINTER(ctypes.c_double)

    def la(co, n):
        a = np.array(co, dtype=np.float64)
        return lib().cfd_debug_lookahead(n, a.ctypes.data_as(dp))
    assert la(PADE, 4096) == 1
    assert la((1, 2, 1 / 3, 1, 1 / 3, 2, 1), 4096) == 2
    assert la((1, 2, 1 / 3, 1, 1 / 3, 2, 1), 64) == 1            # at most two chunks: every sweep is exact
    assert la((1, 2, .45, 1, .45, 2, 1), 4096) == 0             # two-pass
    assert la((1, 2, 3, 4, 5, 6, 7), 4096) == 0
    for alpha in (0.05, 0.2, 0.3, 0.35, 0.42):
        g = (1 - np.sqrt(1 - 4 * alpha * alpha)) / (2 * alpha)   # interior coupling of (alpha, 1, alpha)
        want = 1 if g ** 32 <= 1.2e-16 else 2 if g ** 64 <= 1.2e-16 else 0
        assert la((1, 2, alpha, 1, alpha, 2, 1), 1024) == want, alpha


@pytest.mark.parametrize("scheme", ["pade4", "compact6", "pade4-d2"])
def test_scheme_definitions_match_the_published_schemes(scheme):
    """The library's scheme tables (matrix rows, interior stencil, closure rows), applied with NumPy, reproduce the
    oracle's independent assembly of the published formulas (Lele 1992) on random data."""
    from compact_finite_differences_b200._lib import check, lib
    dp = ctypes.POINTER(ctypes.c_double)
    n, h = 41, 0.23
    out = np.zeros(37)
    check(lib().cfd_debug_scheme(O.SCHEMES[scheme], n, h, out.ctypes.data_as(dp)))
    b1, c1, ai, bi, ci, an, bn, two, a2, b2, c2, am, bm, cm, c0, s1, s2, sgn, nsp = out[:19]
    q, p = out[19:27].reshape(2, 4), out[27:35].reshape(2, 4)
    a, b, c = np.full(n, ai), np.full(n, bi), np.full(n, ci)
    a[0], b[0], c[0] = 0, b1, c1
    a[-1], b[-1], c[-1] = an, bn, 0
    if two:
        a[1], b[1], c[1] = a2, b2, c2
        a[-2], b[-2], c[-2] = am, bm, cm
    wa, wb, wc = O.scheme_system(n, h, scheme)
    assert np.allclose(a, wa, rtol=1e-15) and np.allclose(b, wb, rtol=1e-15) and np.allclose(c, wc, rtol=1e-15)
    f = np.random.default_rng(3).random((5, n))
    fp = np.pad(f, ((0, 0), (2, 2)))
    r = c0 * f + s1 * (fp[:, 3:-1] + sgn * fp[:, 1:-3]) + s2 * (fp[:, 4:] + sgn * fp[:, :-4])
    for k in range(int(nsp)):
        r[:, k] = f[:, :4] @ q[k]
        r[:, n - 1 - k] = f[:, ::-1][:, :4] @ p[k]
    want = O.scheme_rhs(f, h, scheme)
    assert relinf(r, want) < 1e-14
    assert out[35] == 1                                         # 41 rows = two chunks: every sweep is exact
    check(lib().cfd_debug_scheme(O.SCHEMES[scheme], 400, h, out.ctypes.data_as(dp)))
    assert out[35] == (2 if scheme == "compact6" else 1) and out[36] ** (32 * out[35]) <= 1.2e-16


@pytest.mark.parametrize("nz", [1, 3, 4, 5, 7, 40, 48, 64, 128, 500, 512, 1000, 1024])
@pytest.mark.parametrize("slabs", [1, 5, 8])
@pytest.mark.parametrize("ramp", [False, True])
def test_host_gradient_slab_cuts(nz, slabs, ramp):
    """The transfer pipeline of HostGradient (host API of the OpenCL flavour, code/ocl/compact.py:26-61): its z-slabs
    tile [0, nz) without gaps or overlaps, no slab is thicker than nz / slabs, and with `ramp` the first slab -- the lag
    before the device->host copies can start -- is at most an eighth of a uniform slab (down to 4 planes)."""
    from compact_finite_differences_b200.host import HostGradient
    cuts = HostGradient.slab_cuts(nz, slabs, ramp)
    assert cuts[0][0] == 0 and cuts[-1][1] == nz
    assert all(a1 == b0 for (_, b0), (a1, _) in zip(cuts[:-1], cuts[1:]))
    assert all(b > a for a, b in cuts)
    full = max(1, nz // min(slabs, nz))
    assert max(b - a for a, b in cuts) <= full + min(4, full)         # the last slab may absorb a sliver
    if ramp and nz >= 64 * 4 and slabs == 8:
        assert cuts[0][1] == nz // 64
        sizes = [b - a for a, b in cuts]
        assert all(s2 <= 2 * s1 or s2 <= full for s1, s2 in zip(sizes[:-1], sizes[1:]))
    if not ramp:
        assert len(cuts) <= -(-nz // full)


@pytest.mark.parametrize("shape, warps, cut", [
    ((512, 512, 512), 6, False),          # the headline launch: 6 warps per SM (7: 0.545 vs 0.534 ms)
    ((256, 256, 256), 7, False),          # BASELINE configs[1]: short launch, a seventh warp (0.0808 -> 0.0756 ms)
    ((128, 128, 128), 7, False),
    ((64, 512, 512), 7, False),           # the e2e pipeline's slabs
    ((128, 1024, 1024), 6, True),         # the 8-GPU slab: 32-tile lines are cut, 6 warps
    ((512, 1024, 1024), 6, True),
    ((2, 64, 64), 1, False),              # fewer items than warps
])
def test_xy_launch_shape_rule(shape, warps, cut):
    """The launch shape of the fused d/dx + d/dy kernel is a host-side rule (api.cu xy_shape) measured in
    profiles/r2y_sweep_xy_warps.txt: pinned here for a 148-SM device so that a change of the rule shows up on CPU."""
    from compact_finite_differences_b200._lib import check, lib
    w, sub, act = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
    check(lib().cfd_debug_xy_shape(*shape, 148, ctypes.byref(w), ctypes.byref(act), ctypes.byref(sub)))
    assert w.value == warps
    assert (sub.value != 0) == cut
    if cut:
        assert sub.value == (16 | (32 << 8))
    assert act.value >= 1.0


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the driver's reference arm: the reference's own CPU path from oracle/_ref, or the C
    port when the reference did not compile) prints exactly one JSON line with the contract's keys and needs no GPU."""
    import json
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-seconds", "0.5"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["metric"].startswith("grid points/sec per derivative") and d["unit"] == "points/s"
    assert d["higher_is_better"] is True and d["value"] > 1e6 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "512" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["config"]["workload"].startswith("512^3")


@pytest.mark.parametrize("n", [33, 64, 130, 515])
def test_kernel_schedule_is_exact_on_cubics(n):
    """The kernel's schedule (chunked forward sweep, Horner warm-up, look-ahead back-substitution: tests/emulate.py follows
    kernels.cuh step by step) keeps the scheme's exactness: cubic polynomials are differentiated to round-off and a
    mirrored line gives the mirrored derivative with the opposite sign -- truncating the look-ahead at 32 rows costs
    nothing measurable."""
    h = 0.21
    x = np.arange(n) * h - 0.3 * n * h
    rng = np.random.default_rng(n)
    c = rng.random((5, 4))
    F = c[:, :1] + c[:, 1:2] * x + c[:, 2:3] * x ** 2 + c[:, 3:4] * x ** 3
    dF = c[:, 1:2] + 2 * c[:, 2:3] * x + 3 * c[:, 3:4] * x ** 2
    got = stream_lines(F, PADE, h)
    assert np.abs(got - dF).max() <= 2e-13 * np.abs(F).max() / h
    R = rng.random((5, n))
    a = stream_lines(R, PADE, h)
    b = stream_lines(R[:, ::-1].copy(), PADE, h)[:, ::-1]
    assert np.abs(a + b).max() <= 1e-12 * np.abs(a).max()
