"""
oracle/cfd_oracle.py -- Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/README.md).  The product package never does, and has no CPU path.

Three independent checkers live here:

  port   ctypes binding of oracle/cfd_oracle.c (our plain-C restatement; OpenMP over lines).
  ref    ctypes binding of oracle/_ref/libnpts_ref.so = the reference's own
         lanl-implementation/npts.c compiled UNMODIFIED against the one-rank mpi.h stand-in
         (oracle/Makefile `ref`).  Present only where it was built (this container; it travels to
         the GPU box as a git-ignored artefact).
  scipy  the banded-LU form every reference test compares with
         (code/cuda/compact.py:189-203 `scipy_solve_banded`).

plus a NumPy restatement of the reference's multi-rank partition algebra
(code/cuda/compact.py:65-154) used to check the z-partitioned path.

Parity of the oracle itself is pinned in tests/test_oracle.py (golden known-answer values of
lanl-implementation/test_npts.c, port == ref == scipy to 1e-14, committed fixtures).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "liboracle_port.so")
_REF_SO = os.path.join(_HERE, "_ref", "libnpts_ref.so")
REF_TEST_BIN = os.path.join(_HERE, "_ref", "test_npts.run")
REF_TIME_BIN = os.path.join(_HERE, "_ref", "time_npts.run")

_dp = ctypes.POINTER(ctypes.c_double)


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def build(force: bool = False) -> None:
    """Compile the C restatement (always) and the reference's npts (when /root/reference exists)."""
    if force or not os.path.exists(_PORT_SO) or os.path.getmtime(_PORT_SO) < os.path.getmtime(
            os.path.join(_HERE, "cfd_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "port"])
    if os.path.isdir("/root/reference/lanl-implementation") and (force or not os.path.exists(_REF_SO)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_port = None


def port():
    global _port
    if _port is None:
        build()
        lib = ctypes.CDLL(_PORT_SO)
        i, d = ctypes.c_int, ctypes.c_double
        lib.oracle_rhs.argtypes = [_dp, _dp, i, i, i, i, d, _dp, _dp]
        lib.oracle_npts_beta_gam.argtypes = [i, _dp, _dp]
        lib.oracle_npts_solve.argtypes = [_dp, _dp, i, i, i, i, _dp, _dp]
        lib.oracle_near_toeplitz_solve.argtypes = [_dp, i, i, i, i, _dp]
        lib.oracle_cr_solve.argtypes = [_dp, ctypes.c_long, i, _dp]
        lib.oracle_pthomas.argtypes = [_dp, _dp, _dp, _dp, i, ctypes.c_long]
        lib.oracle_derivative.argtypes = [_dp, _dp, i, i, i, i, d]
        lib.oracle_num_threads.restype = i
        lib.oracle_set_num_threads.argtypes = [i]
        for fn in ("oracle_rhs", "oracle_npts_beta_gam", "oracle_npts_solve", "oracle_near_toeplitz_solve",
                   "oracle_cr_solve", "oracle_pthomas", "oracle_derivative", "oracle_set_num_threads"):
            getattr(lib, fn).restype = None
        _port = lib
    return _port


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


_ref = None


def ref():
    """The reference's own npts.c (lanl-implementation/npts.h:6-9), one rank."""
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(_REF_SO)
        i = ctypes.c_int
        lib.precompute_beta_gam.argtypes = [i, i, i, i, _dp, _dp]
        lib.precompute_beta_gam.restype = None
        lib.nonperiodic_tridiagonal_solver.argtypes = [i, i, i, i, _dp, _dp, _dp, _dp, _dp, _dp]
        lib.nonperiodic_tridiagonal_solver.restype = None
        _ref = lib
    return _ref


# ------------------------------------------------------------------------------------------------
# port wrappers
# ------------------------------------------------------------------------------------------------
def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def rhs(f, axis, h, halo_lo=None, halo_hi=None):
    f = _c(f)
    nz, ny, nx = f.shape
    out = np.empty_like(f)
    hl = _c(halo_lo).ravel() if halo_lo is not None else None
    hh = _c(halo_hi).ravel() if halo_hi is not None else None
    port().oracle_rhs(_ptr(f), _ptr(out), nz, ny, nx, axis, float(h), _ptr(hl), _ptr(hh))
    return out


def npts_beta_gam(n):
    beta, gam = np.empty(n), np.empty(n)
    port().oracle_npts_beta_gam(n, _ptr(beta), _ptr(gam))
    return beta, gam


def npts_solve(r, axis):
    r = _c(r)
    nz, ny, nx = r.shape
    beta, gam = npts_beta_gam(r.shape[2 - axis])
    u = np.empty_like(r)
    port().oracle_npts_solve(_ptr(r), _ptr(u), nz, ny, nx, axis, _ptr(beta), _ptr(gam))
    return u


def derivative(f, axis, h):
    """RHS + npts solve on one rank: the reference CPU path (test_npts.c:86-97,126)."""
    f = _c(f)
    nz, ny, nx = f.shape
    df = np.empty_like(f)
    port().oracle_derivative(_ptr(f), _ptr(df), nz, ny, nx, axis, float(h))
    return df


def near_toeplitz_solve(d, coeffs, axis=0):
    d = _c(d).copy()
    nz, ny, nx = d.shape
    co = _c(coeffs)
    port().oracle_near_toeplitz_solve(_ptr(d), nz, ny, nx, axis, _ptr(co))
    return d


def cr_solve(d, coeffs):
    """The reference GPU solver's own cyclic-reduction algorithm, x lines only."""
    d = _c(d).copy()
    n = d.shape[-1]
    co = _c(coeffs)
    port().oracle_cr_solve(_ptr(d), d.size // n, n, _ptr(co))
    return d


def pthomas(a, b, c, d):
    """d is [n, ...]: systems strided by prod(d.shape[1:]) (code/cuda/kernels.cu:115-145)."""
    d = _c(d).copy()
    n = d.shape[0]
    a, b, c = _c(a), _c(b), _c(c)
    port().oracle_pthomas(_ptr(a), _ptr(b), _ptr(c), _ptr(d), n, d.size // n)
    return d


# ------------------------------------------------------------------------------------------------
# reference (unmodified npts.c) wrappers
# ------------------------------------------------------------------------------------------------
def ref_npts_solve(r):
    """x-lines of r[nz,ny,nx] through the reference's nonperiodic_tridiagonal_solver, 1 rank."""
    r = _c(r)
    nz, ny, nx = r.shape
    beta, gam = np.zeros(nx), np.zeros(nx)
    ref().precompute_beta_gam(0, nx, ny, nz, _ptr(beta), _ptr(gam))
    u, phi, psi = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
    # PRINT_TIMINGS (options.h:1) makes the solver print eight timing lines to fd 1; silence them.
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        ref().nonperiodic_tridiagonal_solver(0, nx, ny, nz, _ptr(beta), _ptr(gam), _ptr(r), _ptr(u),
                                             _ptr(phi), _ptr(psi))
        ctypes.CDLL(None).fflush(None)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    return u, beta, gam


def ref_known_answer(n):
    """`Average absolute error` printed by the reference's test_npts (test_npts.c:146-157)."""
    out = subprocess.check_output([REF_TEST_BIN, str(n), str(n), str(n), "1", "1", "1"], text=True)
    return out.strip().splitlines()[-1]


# ------------------------------------------------------------------------------------------------
# scipy: the ground truth of the reference's own tests
# ------------------------------------------------------------------------------------------------
def banded_abc(n, coeffs):
    b1, c1, ai, bi, ci, an, bn = coeffs
    a = np.full(n, float(ai)); b = np.full(n, float(bi)); c = np.full(n, float(ci))
    b[0], c[0], a[-1], b[-1] = b1, c1, an, bn
    return a, b, c


def scipy_solve_banded(a, b, c, rhs_):
    """Same banded layout as code/cuda/compact.py:189-203; rhs_ may be [n] or [n, m]."""
    from scipy.linalg import solve_banded
    ab = np.vstack([np.append(0, c[:-1]), b, np.append(a[1:], 0)])
    return solve_banded((1, 1), ab, rhs_)


PADE = (1., 2., 1. / 4, 1., 1. / 4, 2., 1.)


def scipy_solve_axis(d, coeffs, axis):
    """Solve along `axis` (0=x,1=y,2=z) of d[nz,ny,nx] with LAPACK's banded LU."""
    d = _c(d)
    ax = 2 - axis
    n = d.shape[ax]
    a, b, c = banded_abc(n, coeffs)
    m = np.moveaxis(d, ax, 0).reshape(n, -1)
    x = scipy_solve_banded(a, b, c, m)
    return np.ascontiguousarray(np.moveaxis(x.reshape(np.moveaxis(d, ax, 0).shape), 0, ax))


def scipy_derivative(f, axis, h):
    return scipy_solve_axis(rhs(f, axis, h), PADE, axis)


# ------------------------------------------------------------------------------------------------
# multi-rank partition method of the reference, restated in NumPy (serial emulation of P ranks)
# ------------------------------------------------------------------------------------------------
def partition_local_coeffs(rank, size):
    """code/cuda/compact.py:159-166: interior blocks are pure Toeplitz, ends carry the closures."""
    co = [1., 1. / 4, 1. / 4, 1., 1. / 4, 1. / 4, 1.]
    if rank == 0:
        co[1] = 2.
    if rank == size - 1:
        co[5] = 2.
    return co


def partition_secondary(n, rank, size):
    """Unit responses x_UH, x_LH of the local matrix (code/cuda/compact.py:128-154)."""
    a = np.full(n, .25); b = np.ones(n); c = np.full(n, .25)
    if rank == 0:
        c[0], a[0] = 2.0, 0.0
    if rank == size - 1:
        a[-1], c[-1] = 2.0, 0.0
    r_uh, r_lh = np.zeros(n), np.zeros(n)
    r_uh[0] = -a[0]
    r_lh[-1] = -c[-1]
    return scipy_solve_banded(a, b, c, r_uh), scipy_solve_banded(a, b, c, r_lh)


def partition_reduced_matrix(n_local, size):
    """a, b, c of the 2P x 2P interface system (code/cuda/compact.py:96-111)."""
    uh = np.zeros(2 * size); lh = np.zeros(2 * size)
    for r in range(size):
        x_uh, x_lh = partition_secondary(n_local, r, size)
        uh[2 * r], uh[2 * r + 1] = x_uh[0], x_uh[-1]
        lh[2 * r], lh[2 * r + 1] = x_lh[0], x_lh[-1]
    a = np.zeros(2 * size); b = np.zeros(2 * size); c = np.zeros(2 * size)
    a[0::2] = -1.; a[1::2] = uh[1::2]
    b[0::2] = uh[0::2]; b[1::2] = lh[1::2]
    c[0::2] = lh[0::2]; c[1::2] = -1.
    a[0] = c[0] = 0.; b[0] = 1.
    a[-1] = c[-1] = 0.; b[-1] = 1.
    a[1] = 0.; c[-2] = 0.
    return a, b, c


def partition_derivative(f, axis, h, size):
    """
    Emulate the reference's P-rank derivative along `axis` on one process
    (code/cuda/compact.py:29-44): halo -> local RHS -> local solve -> interface system -> sum.
    Returns the assembled global derivative; equals scipy_derivative(f) up to round-off.
    """
    f = _c(f)
    ax = 2 - axis
    fm = np.moveaxis(f, ax, 0)                       # [N, ...]
    N = fm.shape[0]
    assert N % size == 0
    n = N // size
    rest = fm.shape[1:]
    x_r, faces = [], np.zeros((2 * size,) + rest)
    for r in range(size):
        blk = fm[r * n:(r + 1) * n]
        # RHS with ghost points (code/cuda/kernels.cu:34-44)
        rr = np.empty_like(blk)
        lo = fm[r * n - 1] if r > 0 else None
        hi = fm[(r + 1) * n] if r < size - 1 else None
        rr[1:-1] = (3. / (4 * h)) * (blk[2:] - blk[:-2])
        rr[0] = (3. / (4 * h)) * (blk[1] - lo) if r > 0 else (1. / (2 * h)) * (-5 * blk[0] + 4 * blk[1] + blk[2])
        rr[-1] = (3. / (4 * h)) * (hi - blk[-2]) if r < size - 1 else \
            -(1. / (2 * h)) * (-5 * blk[-1] + 4 * blk[-2] + blk[-3])
        a, b, c = banded_abc(n, partition_local_coeffs(r, size))
        xr = scipy_solve_banded(a, b, c, rr.reshape(n, -1)).reshape(rr.shape)
        x_r.append(xr)
        # negateAndCopyFaces (code/cuda/kernels.cu:76-113)
        faces[2 * r] = 0. if r == 0 else -xr[0]
        faces[2 * r + 1] = 0. if r == size - 1 else -xr[-1]
    ra, rb, rc = partition_reduced_matrix(n, size)
    sol = scipy_solve_banded(ra, rb, rc, faces.reshape(2 * size, -1)).reshape(faces.shape)
    out = np.empty_like(fm)
    for r in range(size):
        x_uh, x_lh = partition_secondary(n, r, size)
        sh = (n,) + (1,) * len(rest)
        out[r * n:(r + 1) * n] = x_r[r] + sol[2 * r] * x_uh.reshape(sh) + sol[2 * r + 1] * x_lh.reshape(sh)
    return np.ascontiguousarray(np.moveaxis(out, 0, ax))


# ------------------------------------------------------------------------------------------------
# the reference's DISTRIBUTED npts solve (lanl-implementation/python/npts.py:172-382; C twin npts.c:275-655),
# restated in NumPy as a serial emulation of the npx ranks of one x line.  The C version has `product_1 = 0.0`
# where the Python twin has 1.0 (npts.c:525 vs python/npts.py:365) and gives wrong last elements for npx > 2
# (lanl-implementation/README.md:3-12), so the restatement follows the Python twin.
# ------------------------------------------------------------------------------------------------
def npts_distributed_beta_gam(nx_local, npx):
    """precompute_beta_gam_dfdx (python/npts.py:172-225): LU pivots handed from rank to rank along the line.
    Returns beta[npx][nx_local], gam[npx][nx_local]."""
    beta = np.zeros((npx, nx_local))
    gam = np.zeros((npx, nx_local))
    last_beta = 0.0
    for r in range(npx):
        if r == 0:
            beta[r, 0], gam[r, 0] = 1.0, 0.0
        else:
            beta[r, 0] = 1. / (1. - (1. / 4) * last_beta * (1. / 4))
            gam[r, 0] = last_beta * (1. / 4)
        for i in range(1, nx_local):
            gam[r, i] = beta[r, i - 1] * (2. if (r == 0 and i == 1) else 1. / 4)
            if (r == npx - 1 and i == nx_local - 1) or (r == 0 and i == 1):
                beta[r, i] = 1. / (1. - 2.0 * beta[r, i - 1] * (1. / 4))
            else:
                beta[r, i] = 1. / (1. - (1. / 4) * beta[r, i - 1] * (1. / 4))
        last_beta = beta[r, -1]
    return beta, gam


def npts_distributed_solve(r, npx, c_twin_product_zero=False):
    """dfdx_parallel (python/npts.py:228-382) for the x lines of r[nz, ny, NX], NX split over npx emulated ranks:
    L->R sweep of (phi, psi) on every rank at once, all-gather of the last faces, prefix combination u_tilda,
    u = phi + u_tilda psi; R->L sweep likewise with the first faces; x = phi + x_tilda psi.
    c_twin_product_zero=True reproduces npts.c:525 (`product_1 = 0.0` in the R->L combination), which drops the
    contributions of ranks further than one block away."""
    r = _c(r)
    nz, ny, NX = r.shape
    assert NX % npx == 0
    nx = NX // npx
    beta, gam = npts_distributed_beta_gam(nx, npx)
    blocks = [r[:, :, m * nx:(m + 1) * nx] for m in range(npx)]
    phi = [np.zeros((nz, ny, nx)) for _ in range(npx)]
    psi = [np.zeros((nz, ny, nx)) for _ in range(npx)]
    u = [np.zeros((nz, ny, nx)) for _ in range(npx)]
    x = [np.zeros((nz, ny, nx)) for _ in range(npx)]

    # ---- L-R sweep (python/npts.py:252-272)
    for m in range(npx):
        b, rr = beta[m], blocks[m]
        if m == 0:
            phi[m][:, :, 0] = 0.0
            psi[m][:, :, 0] = 1.0
        else:
            phi[m][:, :, 0] = b[0] * rr[:, :, 0]
            psi[m][:, :, 0] = -(1. / 4) * b[0]
        for i in range(1, nx):
            phi[m][:, :, i] = b[i] * (rr[:, :, i] - (1. / 4) * phi[m][:, :, i - 1])
            psi[m][:, :, i] = -(1. / 4) * b[i] * psi[m][:, :, i - 1]
        if m == npx - 1:
            i = nx - 1
            phi[m][:, :, i] = b[i] * (rr[:, :, i] - 2 * phi[m][:, :, i - 1])
            psi[m][:, :, i] = -2 * b[i] * psi[m][:, :, i - 1]
    phi_lasts = [p[:, :, -1].copy() for p in phi]            # line_allgather_faces(..., 'last'), :277-281
    psi_lasts = [p[:, :, -1].copy() for p in psi]
    u_first = beta[0][0] * blocks[0][:, :, 0]                 # :293-300, broadcast from the line root
    for m in range(npx):
        if m == 0:
            u_tilda = u_first
        else:                                                 # :302-311
            u_tilda = np.zeros((nz, ny))
            product_2 = np.ones((nz, ny))
            for i in range(m):
                product_1 = np.ones((nz, ny))
                for j in range(i + 1, m):
                    product_1 = product_1 * psi_lasts[j]
                u_tilda = u_tilda + phi_lasts[i] * product_1
                product_2 = product_2 * psi_lasts[i]
            u_tilda = u_tilda + u_first * product_2
        u[m] = phi[m] + u_tilda[:, :, None] * psi[m]          # :316-317

    # ---- R-L sweep (:325-343)
    gam_firsts = [gam[m][0] for m in range(npx)]
    for m in range(npx):
        g = gam[m]
        if m == npx - 1:
            phi[m][:, :, -1] = 0.0
            psi[m][:, :, -1] = 1.0
        else:
            phi[m][:, :, -1] = u[m][:, :, -1]
            psi[m][:, :, -1] = -gam_firsts[m + 1]
        for i in range(1, nx):
            phi[m][:, :, -1 - i] = u[m][:, :, -1 - i] - g[-1 - i + 1] * phi[m][:, :, -1 - i + 1]
            psi[m][:, :, -1 - i] = -g[-1 - i + 1] * psi[m][:, :, -1 - i + 1]
    phi_firsts = [p[:, :, 0].copy() for p in phi]             # :349-353
    psi_firsts = [p[:, :, 0].copy() for p in psi]
    x_last = u[npx - 1][:, :, -1]                             # :361-368
    for m in range(npx):
        if m == npx - 1:
            x_tilda = x_last
        else:                                                 # :370-382
            x_tilda = np.zeros((nz, ny))
            for i in range(m + 2, npx):
                product_1 = np.zeros((nz, ny)) if c_twin_product_zero else np.ones((nz, ny))
                for j in range(m + 1, i):
                    product_1 = product_1 * psi_firsts[j]
                x_tilda = x_tilda + phi_firsts[i] * product_1
            product_2 = np.ones((nz, ny))
            for i in range(m + 1, npx):
                product_2 = product_2 * psi_firsts[i]
            x_tilda = x_tilda + phi_firsts[m + 1] + x_last * product_2
        x[m] = phi[m] + x_tilda[:, :, None] * psi[m]
    return np.concatenate(x, axis=2)


# ------------------------------------------------------------------------------------------------
# Compact schemes beyond the reference's Pade-4 first derivative.  The reference has no
# implementation of them -- only its solver accepts their matrices (near_toeplitz.py:49-50) -- so the ground truth is
# the published scheme (S. K. Lele, J. Comput. Phys. 103 (1992) 16-42: eqs. 2.1, 2.2, 4.1.4, 4.3.4), assembled here
# with plain NumPy slices and solved with the banded LU the reference's own tests use (compact.py:189-203).
# ------------------------------------------------------------------------------------------------
SCHEMES = {"pade4": 0, "compact6": 1, "pade4-d2": 2}


def scheme_system(n, h, scheme):
    """(a, b, c) diagonals of the scheme's matrix for a line of n points."""
    a, b, c = np.zeros(n), np.ones(n), np.zeros(n)
    if scheme == "pade4":
        a[:], c[:] = 0.25, 0.25
        c[0], a[-1] = 2.0, 2.0
    elif scheme == "compact6":
        a[:], c[:] = 1.0 / 3, 1.0 / 3
        a[1] = c[1] = a[-2] = c[-2] = 0.25              # 4th-order Pade rows next to the closures
        c[0], a[-1] = 2.0, 2.0                          # 3rd-order closure rows (the reference's)
    elif scheme == "pade4-d2":
        a[:], c[:] = 0.1, 0.1
        c[0], a[-1] = 11.0, 11.0
    else:
        raise ValueError(scheme)
    a[0] = 0.0
    c[-1] = 0.0
    return a, b, c


def scheme_rhs(f, h, scheme):
    """Right-hand side along the LAST axis of f.  The new schemes are written over differences of neighbouring points
    (every derivative stencil annihilates constants: 13 f0 - 27 f1 + 15 f2 - f3 = -27 (f1 - f0) + 15 (f2 - f0) - (f3 - f0)):
    the same numbers in exact arithmetic, but round-off of order eps |h f'| / h^m instead of eps |f| / h^m -- for the
    second derivative at 256^3 that is the difference between 1e-14 and 3e-11.  'pade4' keeps the reference's literal
    operand order (code/cuda/kernels.cu:34-44)."""
    f = np.asarray(f, dtype=np.float64)
    r = np.empty_like(f)
    if scheme == "pade4":
        r[..., 1:-1] = (3. / (4 * h)) * (f[..., 2:] - f[..., :-2])
        r[..., 0] = (1. / (2 * h)) * (-5 * f[..., 0] + 4 * f[..., 1] + f[..., 2])
        r[..., -1] = -(1. / (2 * h)) * (-5 * f[..., -1] + 4 * f[..., -2] + f[..., -3])
    elif scheme == "compact6":
        r[..., 2:-2] = (14. / 9) * (f[..., 3:-1] - f[..., 1:-3]) / (2 * h) + (1. / 9) * (f[..., 4:] - f[..., :-4]) / (4 * h)
        r[..., 1] = (3. / 4) * (f[..., 2] - f[..., 0]) / h
        r[..., -2] = (3. / 4) * (f[..., -1] - f[..., -3]) / h
        r[..., 0] = (2 * (f[..., 1] - f[..., 0]) + 0.5 * (f[..., 2] - f[..., 0])) / h
        r[..., -1] = -(2 * (f[..., -2] - f[..., -1]) + 0.5 * (f[..., -3] - f[..., -1])) / h
    elif scheme == "pade4-d2":
        c = f[..., 1:-1]
        r[..., 1:-1] = (6. / 5) * ((f[..., 2:] - c) + (f[..., :-2] - c)) / h ** 2
        f0, fn = f[..., 0], f[..., -1]
        r[..., 0] = (-27 * (f[..., 1] - f0) + 15 * (f[..., 2] - f0) - (f[..., 3] - f0)) / h ** 2
        r[..., -1] = (-27 * (f[..., -2] - fn) + 15 * (f[..., -3] - fn) - (f[..., -4] - fn)) / h ** 2
    else:
        raise ValueError(scheme)
    return r


def scheme_derivative(f, axis, h, scheme):
    """Derivative (first, or second for 'pade4-d2') of f[nz,ny,nx] along axis 0 = x, 1 = y, 2 = z."""
    f = _c(f)
    ax = 2 - axis
    g = np.moveaxis(f, ax, -1)
    n = g.shape[-1]
    a, b, c = scheme_system(n, h, scheme)
    r = scheme_rhs(g, h, scheme)
    x = scipy_solve_banded(a, b, c, r.reshape(-1, n).T).T.reshape(g.shape)
    return np.ascontiguousarray(np.moveaxis(x, -1, ax))
