"""
Cartesian process grids for the partitioned derivative -- what the derivative path uses of the reference's
distributed-array layer (code/cuda/gpuDA.py): the (npz, npy, npx) decomposition, the line sub-communicators
(`get_line_DA`, gpuDA.py:154-180), the coordinate generator `DA_arange` (gpuDA.py:402-432) and the block
gather / scatter helpers (gpuDA.py:434-488).  One process per GPU, torch.distributed process groups (NCCL on GPUs,
gloo in the CPU tests) instead of mpi4py communicators.

The derivative itself never uses a ghosted array: it reads ONE boundary plane per line neighbour, straight from the
un-ghosted block (partition.PartitionedDerivative).  `global_to_local` / `local_to_global` (gpuDA.py:61-141, six-face
halo exchange into a [nz+2sw, ny+2sw, nx+2sw] array) are provided for callers of the reference that use the DA for
their own stencils; they are torch copies + point-to-point messages, no kernels of this library.

    da  = DA(None, (nz, ny, nx), (npz, npy, npx))            # every rank, same arguments (collective)
    x, y, z = DA_arange(da, (0, 2 * pi), (0, 2 * pi), (0, 2 * pi), device="cuda")
    ddx = da.derivative(0, dx)                                # PartitionedDerivative on this rank's x line
    dfdx = ddx(f_local)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .compact import LineDA


class DA:
    def __init__(self, comm, local_dims, proc_sizes, stencil_width=1):
        """
        Reference signature (gpuDA.py:9): DA(comm, local_dims, proc_sizes, stencil_width).

        :param comm: torch.distributed process group of all ranks of the grid (None = the default group)
        :param local_dims: (nz, ny, nx) of the block every rank owns
        :param proc_sizes: (npz, npy, npx); ranks are laid out row-major like MPI_Cart_create:
                           rank = (mz * npy + my) * npx + mx
        :param stencil_width: ghost width of global_to_local (the derivative itself always exchanges one plane)
        """
        assert dist.is_initialized(), "torch.distributed must be initialised (one process per GPU)"
        self.comm = comm
        self.local_dims = tuple(int(s) for s in local_dims)
        self.proc_sizes = tuple(int(s) for s in proc_sizes)
        self.stencil_width = int(stencil_width)          # used by global_to_local only; the derivative exchanges 1 plane
        assert self.stencil_width >= 1
        self.rank = dist.get_rank(comm)
        self.size = dist.get_world_size(comm)
        self.nz, self.ny, self.nx = self.local_dims
        self.npz, self.npy, self.npx = self.proc_sizes
        assert self.size == self.npz * self.npy * self.npx, \
            f"{self.size} ranks cannot form a {self.proc_sizes} process grid"           # gpuDA.py:38
        self.mz, rem = divmod(self.rank, self.npy * self.npx)
        self.my, self.mx = divmod(rem, self.npx)
        # Line groups.  new_group is collective over ALL ranks, so every rank creates every line's group in the
        # same order and keeps the ones it belongs to.
        ranks = np.arange(self.size).reshape(self.proc_sizes)                            # gpuDA.py:165
        to_global = (lambda r: int(r)) if comm is None else (lambda r: dist.get_global_rank(comm, int(r)))
        self._line = {}
        for direction, axis in ((0, 2), (1, 1), (2, 0)):
            if self.proc_sizes[axis] == 1:
                self._line[direction] = (None, 0, 1)
                continue
            lines = np.moveaxis(ranks, axis, -1).reshape(-1, self.proc_sizes[axis])
            for line in lines:
                g = dist.new_group([to_global(r) for r in line])
                if self.rank in line:
                    self._line[direction] = (g, int(np.where(line == self.rank)[0][0]), len(line))

    # -- the reference's accessors -------------------------------------------------------------------------
    def line(self, direction):
        """(process group, rank in the line, line size) of this rank's line along direction 0 = x, 1 = y, 2 = z."""
        return self._line[int(direction)]

    def get_line_DA(self, direction):
        """gpuDA.py:154-180.  The reference permutes local_dims so that the line axis comes last (its y / z
        derivatives work on host-transposed copies); here the block stays [nz, ny, nx] and `direction` travels
        with the line object."""
        g, r, n = self.line(direction)
        da = LineDA(self.local_dims, r, n, int(direction))
        da.group = g
        return da

    def derivative(self, direction, spacing, mode="fused", comm="pairwise"):
        """The derivative operator of this rank's block along `direction`, partitioned over the line group."""
        from .partition import PartitionedDerivative
        g, r, n = self.line(direction)
        if n == 1:
            from .compact import CompactFiniteDifferenceSolver
            return CompactFiniteDifferenceSolver(self.local_dims, spacing, int(direction))
        return PartitionedDerivative(self.local_dims, spacing, int(direction), group=g, mode=mode, comm=comm)

    def create_local_vector(self, device=None):
        """gpuDA.py:51-59: the ghosted array [nz + 2 sw, ny + 2 sw, nx + 2 sw], zero-filled."""
        sw = self.stencil_width
        return torch.zeros((self.nz + 2 * sw, self.ny + 2 * sw, self.nx + 2 * sw), dtype=torch.float64, device=device)

    def _neighbour(self, dim, step):
        """Rank (in `comm`) of the neighbour one block away along tensor dimension dim (0 = z), or None at the edge."""
        m = [self.mz, self.my, self.mx]
        m[dim] += step
        if m[dim] < 0 or m[dim] >= self.proc_sizes[dim]:
            return None
        return (m[0] * self.npy + m[1]) * self.npx + m[2]

    def global_to_local(self, global_array, local_array):
        """gpuDA.py:61-132: copy the block into the interior of the ghosted array and fill its six ghost faces
        (width stencil_width) from the face neighbours; ghost cells at physical boundaries, edges and corners are
        left as they were."""
        sw = self.stencil_width
        g, l = global_array, local_array
        assert tuple(g.shape) == self.local_dims and tuple(l.shape) == tuple(s + 2 * sw for s in self.local_dims)
        inner = (slice(sw, sw + self.nz), slice(sw, sw + self.ny), slice(sw, sw + self.nx))
        l[inner] = g
        ops, recvs = [], []
        peer = (lambda r: r) if self.comm is None else (lambda r: dist.get_global_rank(self.comm, r))
        for dim in range(3):
            n = self.local_dims[dim]
            for step, send_lo, ghost_lo in ((-1, 0, 0), (+1, n - sw, sw + n)):
                nb = self._neighbour(dim, step)
                if nb is None:
                    continue
                send = g.narrow(dim, send_lo, sw).contiguous()
                recv = torch.empty_like(send)
                ops.append(dist.P2POp(dist.isend, send, peer(nb), self.comm))
                ops.append(dist.P2POp(dist.irecv, recv, peer(nb), self.comm))
                idx = list(inner)
                idx[dim] = slice(ghost_lo, ghost_lo + sw)
                recvs.append((tuple(idx), recv))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for idx, recv in recvs:
            l[idx] = recv
        return l

    def local_to_global(self, local_array, global_array):
        """gpuDA.py:134-141: strip the ghost cells."""
        sw = self.stencil_width
        global_array.copy_(local_array[sw:sw + self.nz, sw:sw + self.ny, sw:sw + self.nx])
        return global_array

    def create_global_vector(self, device=None):
        """gpuDA.py:41-49: this rank's block, zero-filled (the reference's name for the un-ghosted array)."""
        return torch.zeros(self.local_dims, dtype=torch.float64, device=device)

    @property
    def global_dims(self):
        return (self.nz * self.npz, self.ny * self.npy, self.nx * self.npx)

    @property
    def block_start(self):
        return (self.mz * self.nz, self.my * self.ny, self.mx * self.nx)


def DA_arange(da, x_range, y_range, z_range, device=None):
    """Coordinates of this rank's block for a grid spanning the given ranges (gpuDA.py:402-432).
    Returns x, y, z of shape (nz, ny, nx): NumPy arrays, or tensors on `device` if one is given (the block is
    then generated on the GPU -- fields too large for one host array never exist on the host)."""
    nz, ny, nx = da.nz, da.ny, da.nx
    NZ, NY, NX = da.global_dims
    dx = float(x_range[-1] - x_range[0]) / (NX - 1)
    dy = float(y_range[-1] - y_range[0]) / (NY - 1)
    dz = float(z_range[-1] - z_range[0]) / (NZ - 1)
    x0 = x_range[0] + da.mx * nx * dx
    y0 = y_range[0] + da.my * ny * dy
    z0 = z_range[0] + da.mz * nz * dz
    if device is None:
        z, y, x = np.meshgrid(np.linspace(z0, z0 + (nz - 1) * dz, nz), np.linspace(y0, y0 + (ny - 1) * dy, ny),
                              np.linspace(x0, x0 + (nx - 1) * dx, nx), indexing="ij")
        return x, y, z
    mk = lambda a, n, d: a + d * torch.arange(n, dtype=torch.float64, device=device)  # noqa: E731
    x = mk(x0, nx, dx)[None, None, :].expand(nz, ny, nx)
    y = mk(y0, ny, dy)[None, :, None].expand(nz, ny, nx)
    z = mk(z0, nz, dz)[:, None, None].expand(nz, ny, nx)
    return x, y, z


def DA_gather_blocks(da, x_local, x_global=None, root=0):
    """Assemble the global (NZ, NY, NX) array on `root` from every rank's block (gpuDA.py:462-488).
    Returns x_global on root, None elsewhere."""
    assert tuple(x_local.shape) == da.local_dims
    blocks = None
    if da.rank == root:
        blocks = [torch.empty_like(x_local) for _ in range(da.size)]
    dst = root if da.comm is None else dist.get_global_rank(da.comm, root)
    dist.gather(x_local.contiguous(), blocks, dst=dst, group=da.comm)
    if da.rank != root:
        return None
    if x_global is None:
        x_global = torch.empty(da.global_dims, dtype=x_local.dtype, device=x_local.device)
    nz, ny, nx = da.local_dims
    for r, b in enumerate(blocks):
        mz, rem = divmod(r, da.npy * da.npx)
        my, mx = divmod(rem, da.npx)
        x_global[mz * nz:(mz + 1) * nz, my * ny:(my + 1) * ny, mx * nx:(mx + 1) * nx] = b
    return x_global


def DA_scatter_blocks(da, x_global, x_local, root=0):
    """Hand every rank its block of the global array held by `root` (gpuDA.py:434-460)."""
    assert tuple(x_local.shape) == da.local_dims
    blocks = None
    if da.rank == root:
        nz, ny, nx = da.local_dims
        blocks = []
        for r in range(da.size):
            mz, rem = divmod(r, da.npy * da.npx)
            my, mx = divmod(rem, da.npx)
            blocks.append(x_global[mz * nz:(mz + 1) * nz, my * ny:(my + 1) * ny, mx * nx:(mx + 1) * nx].contiguous())
    src = root if da.comm is None else dist.get_global_rank(da.comm, root)
    dist.scatter(x_local, blocks, src=src, group=da.comm)
    return x_local
