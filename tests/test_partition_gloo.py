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

    def __init__(self, local_shape, h, rank, size, direction=2):
        from oracle import cfd_oracle as O
        self.O, self.shape, self.h, self.rank, self.size = O, tuple(local_shape), h, rank, size
        self.direction, self.spacing = direction, h
        self.ax = 2 - direction                    # tensor dimension of the line
        self.n = local_shape[self.ax]
        self.co = O.partition_local_coeffs(rank, size)

    def _first(self, a):
        return np.take(a, 0, axis=self.ax)

    def _last(self, a):
        return np.take(a, self.n - 1, axis=self.ax)

    def _along(self, v):
        """A per-row vector shaped to broadcast along the line axis of the block."""
        sh = [1, 1, 1]
        sh[self.ax] = self.n
        return v.reshape(sh)

    def _planes(self, p):
        """A per-line plane shaped to broadcast over the rows of the block."""
        return np.expand_dims(np.asarray(p), self.ax)

    def nb_layout(self):
        lo = self.rank - 1 if self.rank > 0 else self.rank
        hi = self.rank + 1 if self.rank < self.size - 1 else self.rank
        return hi - lo + 1, self.rank - lo

    def _local(self, f, lo, hi):
        O = self.O
        rr = O.rhs(f.numpy(), self.direction, self.h, halo_lo=None if lo is None else lo.numpy(),
                   halo_hi=None if hi is None else hi.numpy())
        a, b, c = O.banded_abc(self.n, self.co)
        rows_first = np.moveaxis(rr, self.ax, 0)
        x = O.scipy_solve_banded(a, b, c, rows_first.reshape(self.n, -1)).reshape(rows_first.shape)
        return rr, np.ascontiguousarray(np.moveaxis(x, 0, self.ax))

    def apply_local(self, f, out, lo, hi):
        out = torch.empty_like(f) if out is None else out
        out.copy_(torch.from_numpy(self._local(f, lo, hi)[1]))
        return out

    def edge_faces(self, f, faces, lo=None, hi=None):
        xr = self._local(f, lo, hi)[1]
        faces[0] = 0.0 if self.rank == 0 else torch.from_numpy(-self._first(xr))
        faces[1] = 0.0 if self.rank == self.size - 1 else torch.from_numpy(-self._last(xr))
        return faces

    def interface_pack(self, df, faces):
        faces[0] = 0.0 if self.rank == 0 else torch.from_numpy(-self._first(df.numpy()))
        faces[1] = 0.0 if self.rank == self.size - 1 else torch.from_numpy(-self._last(df.numpy()))
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
        res = xr + self._planes(ab[0].numpy()) * self._along(xu) + self._planes(ab[1].numpy()) * self._along(xl)
        out = torch.empty_like(f) if out is None else out
        out.copy_(torch.from_numpy(res))
        return out

    def reduced_correct(self, df, faces_all):
        O = self.O
        al, be = self._unknowns(faces_all, False)
        xu, xl = O.partition_secondary(self.n, self.rank, self.size)
        df += torch.from_numpy(self._planes(al) * self._along(xu) + self._planes(be) * self._along(xl))
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
            op._buf, op._side, op._pending, op._zp = None, None, None, None
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


# ---------------------------------------------------------------------------------------------------
# Cartesian process grids (grid.DA): line groups, DA_arange, block gather / scatter, derivatives along x and y lines
# ---------------------------------------------------------------------------------------------------
def _worker_grid(rank, world, port, proc_sizes, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cfd_oracle as O
        from compact_finite_differences_b200.grid import DA, DA_arange, DA_gather_blocks, DA_scatter_blocks
        from compact_finite_differences_b200.partition import PartitionedDerivative

        local = (70, 68, 66)
        da = DA(None, local, proc_sizes)
        npz, npy, npx = proc_sizes
        assert (da.mz, da.my, da.mx) == np.unravel_index(rank, proc_sizes)          # MPI_Cart row-major coordinates
        for direction, m, n in ((0, da.mx, npx), (1, da.my, npy), (2, da.mz, npz)):
            lda = da.get_line_DA(direction)
            assert (lda.rank, lda.size, lda.direction) == (m, n, direction)
            g, r, sz = da.line(direction)
            if sz > 1:                                         # the line group really is the line: sum of coordinates
                t = torch.tensor([float(rank)])
                dist.all_reduce(t, group=g)
                idx = [da.mz, da.my, da.mx]
                want = 0
                for k in range(sz):
                    idx[2 - direction] = k
                    want += int(np.ravel_multi_index(idx, proc_sizes))
                assert int(t.item()) == want

        # DA_arange: the blocks tile the global linspace grid (gpuDA.py:402-432; test_misc.py:15-28)
        x, y, z = DA_arange(da, (0, 1), (0, 2), (0, 3))
        NZ, NY, NX = da.global_dims
        gz, gy, gx = np.meshgrid(np.linspace(0, 3, NZ), np.linspace(0, 2, NY), np.linspace(0, 1, NX), indexing="ij")
        z0, y0, x0 = da.block_start
        sl = (slice(z0, z0 + local[0]), slice(y0, y0 + local[1]), slice(x0, x0 + local[2]))
        assert np.allclose(x, gx[sl], atol=1e-14) and np.allclose(y, gy[sl], atol=1e-14) and np.allclose(z, gz[sl], atol=1e-14)
        xt, yt, zt = DA_arange(da, (0, 1), (0, 2), (0, 3), device="cpu")
        assert np.allclose(xt.numpy(), x, atol=1e-14) and np.allclose(zt.numpy(), z, atol=1e-14)

        # scatter / gather round trip of a global field held by rank 0
        rng = np.random.default_rng(5)
        f = rng.random(da.global_dims)                      # same on every rank (seeded)
        blk = torch.empty(local, dtype=torch.float64)
        DA_scatter_blocks(da, torch.from_numpy(f) if rank == 0 else None, blk)
        assert np.array_equal(blk.numpy(), f[sl])
        back = DA_gather_blocks(da, blk)
        if rank == 0:
            assert np.array_equal(back.numpy(), f)
        else:
            assert back is None

        # global_to_local / local_to_global (code/cuda/test/test_gpuDA/test_3d.py:17-66): fill with the rank, ghosts = neighbours
        for sw in (1, 2):
            da_sw = da if sw == 1 else DA(None, local, proc_sizes, stencil_width=2)
            a = da_sw.create_global_vector()
            a.fill_(float(rank))
            b = da_sw.create_local_vector()
            b.fill_(-1.0)
            da_sw.global_to_local(a, b)
            assert torch.all(b[sw:-sw, sw:-sw, sw:-sw] == rank)
            for dim in range(3):
                for step, ghost in ((-1, slice(0, sw)), (+1, slice(-sw, None))):
                    idx = [slice(sw, -sw)] * 3
                    idx[dim] = ghost
                    nb = da_sw._neighbour(dim, step)
                    want_v = -1.0 if nb is None else float(nb)          # physical boundaries stay untouched
                    assert torch.all(b[tuple(idx)] == want_v), (rank, dim, step)
                    if nb is not None:
                        c = list(np.unravel_index(rank, proc_sizes))
                        c[dim] += step
                        assert nb == int(np.ravel_multi_index(c, proc_sizes))
            assert torch.all(b[:sw, :sw, :] == -1.0)                       # edges / corners are not exchanged
            back_g = torch.empty(local, dtype=torch.float64)
            da_sw.local_to_global(b, back_g)
            assert torch.all(back_g == rank)

        # derivative along every direction through the line groups, block kernels replaced by the oracle
        worst = 0.0
        for direction in range(3):
            h = 0.1 + 0.03 * direction
            g, r, sz = da.line(direction)
            want = O.derivative(f, direction, h)[sl]
            if sz == 1:
                continue                                     # a local line: CompactFiniteDifferenceSolver, GPU tests
            for mode, comm in (("fused", "pairwise"), ("fused", "allgather"), ("reference", "allgather")):
                op = PartitionedDerivative.__new__(PartitionedDerivative)
                op.group, op.rank, op.size, op.direction = g, r, sz, direction
                op.local_shape = local
                op.solver = _OracleBlockSolver(local, h, r, sz, direction)
                op.mode, op.comm = mode, comm
                op._buf, op._side, op._pending, op._zp = None, None, None, None
                got = op(blk).numpy()
                worst = max(worst, np.abs(got - want).max() / np.abs(want).max())
        ret[rank] = worst
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("proc_sizes", [(1, 2, 2), (2, 1, 2)])
def test_process_grid_gloo(proc_sizes):
    """grid.DA on a 2-D process grid (4 ranks): coordinates, line groups, DA_arange, scatter / gather of blocks and
    PartitionedDerivative along the partitioned directions (the reference runs 2x2x2, code/cuda/test/Makefile:9)."""
    world = int(np.prod(proc_sizes))
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_grid, args=(world, port, proc_sizes, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r] < 1e-13, f"rank {r}: rel L-inf {ret[r]}"
