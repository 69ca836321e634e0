"""
Host-side logic of the z-partitioned path on CPU: two processes, gloo backend.  The product's own exchange
functions (partition.exchange_halo_planes / gather_interface_planes) carry the planes; block-local solves come from
the ORACLE (this is a test: the product has no CPU compute path), and the reduced-system tables come from the
library's host-only inspection entry.  Checks that halo planes, interface all-gather and correction assemble
the one-rank derivative.
"""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cfd_oracle as O
        from compact_finite_differences_b200 import (exchange_halo_planes, exchange_interface_planes,
                                                     gather_interface_planes)
        from compact_finite_differences_b200._lib import check, lib

        rng = np.random.default_rng(1234)                  # same field on every rank
        NZ, ny, nx = 16 * world, 6, 10
        f = rng.random((NZ, ny, nx))
        h = 0.3
        n = NZ // world
        blk = torch.from_numpy(f[rank * n:(rank + 1) * n].copy())

        # (1) halo exchange through the product's plumbing
        lo, hi = exchange_halo_planes(blk[0].contiguous(), blk[-1].contiguous(), rank, world)
        if rank > 0:
            assert np.array_equal(lo.numpy(), f[rank * n - 1])
        else:
            assert lo is None
        if rank < world - 1:
            assert np.array_equal(hi.numpy(), f[(rank + 1) * n])
        else:
            assert hi is None

        # (2) block-local solve (oracle stands in for the CUDA kernel here)
        rr = O.rhs(blk.numpy(), 2, h, halo_lo=None if lo is None else lo.numpy(),
                   halo_hi=None if hi is None else hi.numpy())
        a, b, c = O.banded_abc(n, O.partition_local_coeffs(rank, world))
        xr = O.scipy_solve_banded(a, b, c, rr.reshape(n, -1)).reshape(rr.shape)

        # (3) interface planes, all-gathered
        faces = torch.zeros((2, ny, nx), dtype=torch.float64)
        if rank > 0:
            faces[0] = torch.from_numpy(-xr[0])
        if rank < world - 1:
            faces[1] = torch.from_numpy(-xr[-1])
        allf = gather_interface_planes(faces, world).numpy()
        assert allf.shape == (2 * world, ny, nx)
        assert np.array_equal(allf[2 * rank:2 * rank + 2], faces.numpy())

        # (4) reduced system + correction with the LIBRARY's tables
        dp = ctypes.POINTER(ctypes.c_double)
        xu, xl = np.zeros(n), np.zeros(n)
        ra, rb, rc = (np.zeros(2 * world) for _ in range(3))
        check(lib().cfd_debug_secondary(n, rank, world, *(v.ctypes.data_as(dp) for v in (xu, xl, ra, rb, rc))))
        sol = O.scipy_solve_banded(ra, rb, rc, allf.reshape(2 * world, -1)).reshape(allf.shape)
        out = xr + sol[2 * rank] * xu[:, None, None] + sol[2 * rank + 1] * xl[:, None, None]

        want = O.derivative(f, 2, h)[rank * n:(rank + 1) * n]
        err = np.abs(out - want).max() / np.abs(want).max()

        # (5) neighbour-only exchange carries exactly the planes the virtual layout expects
        ip = ctypes.POINTER(ctypes.c_int)
        pv, own = ctypes.c_int(), ctypes.c_int()
        check(lib().cfd_debug_neighbour(n, rank, world, ctypes.byref(pv), ctypes.byref(own), None, None, None))
        pv, own = pv.value, own.value
        nb = torch.zeros((2 * pv, ny, nx), dtype=torch.float64)
        nb[2 * own:2 * own + 2] = faces
        exchange_interface_planes(nb, own, pv, rank, world)
        lo_r = rank - own
        expect = allf[2 * lo_r:2 * (lo_r + pv)].copy()
        expect[0] = 0.0
        expect[-1] = 0.0
        assert np.array_equal(nb.numpy(), expect)
        ret[rank] = err
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_z_partition_plumbing_gloo(world):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] < 1e-13, f"rank {r}: rel L-inf {ret[r]}"


class _OracleBlockSolver:
    """CPU stand-in for CompactFiniteDifferenceSolver's multi-rank methods, built on the ORACLE, so that the
    orchestration of ZPartitionedDerivative (mode / comm switching, buffers, call order) can run under gloo.
    Test-only: the product class always talks to libcfd_b200."""

    def __init__(self, local_shape, h, rank, size):
        from oracle import cfd_oracle as O
        self.O, self.shape, self.h, self.rank, self.size = O, tuple(local_shape), h, rank, size
        self.n = local_shape[0]
        self.co = O.partition_local_coeffs(rank, size)
        self.direction, self.spacing = 2, h

    def nb_layout(self):
        lo = self.rank - 1 if self.rank > 0 else self.rank
        hi = self.rank + 1 if self.rank < self.size - 1 else self.rank
        return hi - lo + 1, self.rank - lo

    def _local(self, f, lo, hi):
        O = self.O
        rr = O.rhs(f.numpy(), 2, self.h, halo_lo=None if lo is None else lo.numpy(),
                   halo_hi=None if hi is None else hi.numpy())
        a, b, c = O.banded_abc(self.n, self.co)
        return rr, O.scipy_solve_banded(a, b, c, rr.reshape(self.n, -1)).reshape(rr.shape)

    def apply_local(self, f, out, lo, hi):
        out = torch.empty_like(f) if out is None else out
        out.copy_(torch.from_numpy(self._local(f, lo, hi)[1]))
        return out

    def edge_faces(self, f, faces, lo=None, hi=None):
        xr = self._local(f, lo, hi)[1]
        faces[0] = 0.0 if self.rank == 0 else torch.from_numpy(-xr[0])
        faces[1] = 0.0 if self.rank == self.size - 1 else torch.from_numpy(-xr[-1])
        return faces

    def interface_pack(self, df, faces):
        faces[0] = 0.0 if self.rank == 0 else -df[0]
        faces[1] = 0.0 if self.rank == self.size - 1 else -df[-1]
        return faces

    def _unknowns(self, faces, neighbours_only):
        O = self.O
        ra, rb, rc = O.partition_reduced_matrix(self.n, self.size)
        fa = faces.numpy()
        if neighbours_only:
            pv, own = self.nb_layout()
            lo = self.rank - own
            a, b, c = ra[2 * lo:2 * (lo + pv)].copy(), rb[2 * lo:2 * (lo + pv)].copy(), rc[2 * lo:2 * (lo + pv)].copy()
            a[0] = c[0] = 0.0; b[0] = 1.0; a[-1] = c[-1] = 0.0; b[-1] = 1.0; a[1] = 0.0; c[-2] = 0.0
            fa = fa.copy(); fa[0] = 0.0; fa[-1] = 0.0
            sol = O.scipy_solve_banded(a, b, c, fa.reshape(2 * pv, -1)).reshape(fa.shape)
            return sol[2 * own], sol[2 * own + 1]
        sol = O.scipy_solve_banded(ra, rb, rc, fa.reshape(2 * self.size, -1)).reshape(fa.shape)
        return sol[2 * self.rank], sol[2 * self.rank + 1]

    def reduced_unknowns(self, faces, ab, neighbours_only=False, **kw):
        al, be = self._unknowns(faces, neighbours_only)
        ab[0] = torch.from_numpy(al)
        ab[1] = torch.from_numpy(be)
        return ab

    def apply_coupled(self, f, out, lo, hi, ab):
        O = self.O
        xu, xl = O.partition_secondary(self.n, self.rank, self.size)
        xr = self._local(f, lo, hi)[1]
        res = xr + ab[0].numpy() * xu[:, None, None] + ab[1].numpy() * xl[:, None, None]
        out = torch.empty_like(f) if out is None else out
        out.copy_(torch.from_numpy(res))
        return out

    def reduced_correct(self, df, faces_all):
        O = self.O
        al, be = self._unknowns(faces_all, False)
        xu, xl = O.partition_secondary(self.n, self.rank, self.size)
        df += torch.from_numpy(al * xu[:, None, None] + be * xl[:, None, None])
        return df


def _worker_class(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cfd_oracle as O
        from compact_finite_differences_b200.partition import ZPartitionedDerivative

        rng = np.random.default_rng(99)
        n, ny, nx = 70, 4, 6
        f = rng.random((n * world, ny, nx))
        h = 0.17
        want = O.derivative(f, 2, h)[rank * n:(rank + 1) * n]
        slab = torch.from_numpy(f[rank * n:(rank + 1) * n].copy())
        worst = 0.0
        for mode, comm in (("fused", "allgather"), ("fused", "pairwise"), ("reference", "allgather")):
            op = ZPartitionedDerivative.__new__(ZPartitionedDerivative)          # no CUDA plan: inject the stand-in
            op.group, op.rank, op.size, op.direction = None, rank, world, 2
            op.local_shape = (n, ny, nx)
            op.solver = _OracleBlockSolver(op.local_shape, h, rank, world)
            op.mode, op.comm = mode, comm
            op._buf, op._side, op._pending, op._peer = None, None, None, None
            got = op(slab).numpy()
            worst = max(worst, np.abs(got - want).max() / np.abs(want).max())
        ret[rank] = worst
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_derivative_orchestration_gloo(world):
    """ZPartitionedDerivative's own __call__ / _exchange logic (fused + all-gather, fused + pairwise, reference order)
    under gloo, with the block kernels replaced by the oracle."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_class, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r] < 1e-13, f"rank {r}: rel L-inf {ret[r]}"
