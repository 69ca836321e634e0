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
