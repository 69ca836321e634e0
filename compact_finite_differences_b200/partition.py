"""
Multi-GPU derivative on a z-partition: one process per GPU, torch.distributed (NCCL over NVLink) for the two
exchanges the path really has, hand-written kernels for everything else.

The reference's multi-rank dfdx (code/cuda/compact.py:29-44) is
    halo exchange (gpuDA.global_to_local) -> RHS -> local solve -> gather interface faces to the line root ->
    reduced solve on the root -> scatter -> sum.
Here, for a grid split into P slabs along z:
    d/dx, d/dy : no communication at all (lines never leave the slab);
    d/dz       : (1) send/recv ONE boundary plane of f with each z-neighbour;
                 (2) fused RHS + local solve (one kernel, cfd_apply);
                 (3) pack the two interface planes (cfd_interface_pack) and ALL-GATHER them -- every rank then
                     solves the identical 2P-unknown reduced system redundantly (no root, no scatter);
                 (4) correction x += alpha*x_UH + beta*x_LH on the few planes next to the interfaces where the
                     secondary solutions are above fp64 round-off (cfd_reduced_correct).
The two collectives below are plain functions over torch tensors so that the host-side logic is testable on
CPU with the gloo backend (tests/test_partition_gloo.py); compute always goes through libcfd_b200.so.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .compact import CompactFiniteDifferenceSolver


def _peer(group, r):
    return r if group is None else dist.get_global_rank(group, r)


def exchange_halo_planes(first_plane, last_plane, rank, size, group=None, halo_lo=None, halo_hi=None):
    """
    One-plane halo exchange along the partitioned axis (the z faces of DA.global_to_local,
    code/cuda/gpuDA.py:86-113, without the other four faces the derivative never reads).
    Returns (halo_lo, halo_hi): the last plane of rank-1 and the first plane of rank+1 (None at the ends).
    """
    ops = []
    if rank > 0:
        if halo_lo is None:
            halo_lo = torch.empty_like(first_plane)
        ops.append(dist.P2POp(dist.isend, first_plane, _peer(group, rank - 1), group))
        ops.append(dist.P2POp(dist.irecv, halo_lo, _peer(group, rank - 1), group))
    else:
        halo_lo = None
    if rank < size - 1:
        if halo_hi is None:
            halo_hi = torch.empty_like(last_plane)
        ops.append(dist.P2POp(dist.isend, last_plane, _peer(group, rank + 1), group))
        ops.append(dist.P2POp(dist.irecv, halo_hi, _peer(group, rank + 1), group))
    else:
        halo_hi = None
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return halo_lo, halo_hi


def gather_interface_planes(faces, size, group=None, out=None):
    """faces: [2, plane] of this rank -> [2*size, plane] of the whole line, rank-major
    (replaces Gather-to-root + Scatter of code/cuda/compact.py:93-94,121-122)."""
    if out is None:
        out = torch.empty((2 * size,) + tuple(faces.shape[1:]), dtype=faces.dtype, device=faces.device)
    dist.all_gather_into_tensor(out, faces, group=group)
    return out


class ZPartitionedDerivative:
    """Derivative of a z-partitioned field; `local_shape` is this rank's slab [nz/P, ny, nx]."""

    def __init__(self, local_shape, spacing, direction, group=None):
        assert dist.is_initialized(), "torch.distributed must be initialised (one process per GPU)"
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.direction = int(direction)
        self.local_shape = tuple(int(s) for s in local_shape)
        part = (self.rank, self.size) if self.direction == 2 else (0, 1)
        self.solver = CompactFiniteDifferenceSolver(self.local_shape, spacing, self.direction, part=part)
        self._buf = None

    def _buffers(self, f):
        if self._buf is None or self._buf[0].device != f.device:
            nz, ny, nx = self.local_shape
            mk = lambda *s: torch.empty(s, dtype=torch.float64, device=f.device)  # noqa: E731
            self._buf = (mk(ny, nx), mk(ny, nx), mk(2, ny, nx), mk(2 * self.size, ny, nx))
        return self._buf

    def __call__(self, f, out=None):
        if self.direction != 2 or self.size == 1:
            return self.solver(f, out)
        lo_buf, hi_buf, faces, faces_all = self._buffers(f)
        halo_lo, halo_hi = exchange_halo_planes(f[0], f[-1], self.rank, self.size, self.group, lo_buf, hi_buf)
        out = self.solver.apply_local(f, out, halo_lo, halo_hi)
        self.solver.interface_pack(out, faces)
        gather_interface_planes(faces, self.size, self.group, faces_all)
        self.solver.reduced_correct(out, faces_all)
        return out
