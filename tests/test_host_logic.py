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
    assert _lib.lib().cfd_version() == 100


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
