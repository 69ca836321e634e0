"""
Multi-GPU derivative on a z-partition: one process per GPU, torch.distributed (NCCL over NVLink) for the two
exchanges the path really has, hand-written kernels for everything else.

The reference's multi-rank dfdx (code/cuda/compact.py:29-44) is
    halo exchange (gpuDA.global_to_local) -> RHS -> local solve -> gather interface faces to the line root ->
    reduced solve on the root -> scatter -> sum.
Here, for a grid split into P slabs along z:
    d/dx, d/dy : no communication at all (lines never leave the slab);
    d/dz       : (1) send/recv ONE boundary plane of f with each z-neighbour;
                 (2) interface planes -x_R[first], -x_R[last] from the 33 + 34 planes next to the slab ends
                     (cfd_edge_faces: they do not depend on planes further away, to 5e-19);
                 (3) ALL-GATHER them (or, comm="pairwise"/"nvlink", one plane from each z-neighbour) -- every
                     rank then solves the reduced system redundantly for its own two unknowns per line
                     (cfd_reduced_unknowns; no root, no scatter);
                 (4) ONE fused kernel (cfd_apply_coupled): RHS + solve with the two interface unknowns folded
                     into rows 0 and n-1 -- the final derivative, no correction pass.
                 mode="reference" keeps the reference's order instead: local solve (cfd_apply) ->
                 cfd_interface_pack -> all-gather -> cfd_reduced_correct (correction sweep on the planes next
                 to the interfaces); it also serves slabs thinner than 66 planes.
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


def exchange_interface_planes(faces_nb, own, pv, rank, size, group=None):
    """
    Neighbour-only replacement of the all-gather: faces_nb is [2*pv, plane] over the virtual ranks
    (rank-1 if any, rank, rank+1 if any) with this rank's own faces already at planes 2*own, 2*own+1.
    Sends faces[0] (= -x_R[first]) to rank-1 and faces[1] (= -x_R[last]) to rank+1; receives the left neighbour's
    faces[1] into plane 2*own-1 and the right neighbour's faces[0] into plane 2*own+2.
    """
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, faces_nb[2 * own], _peer(group, rank - 1), group))
        ops.append(dist.P2POp(dist.irecv, faces_nb[2 * own - 1], _peer(group, rank - 1), group))
    if rank < size - 1:
        ops.append(dist.P2POp(dist.isend, faces_nb[2 * own + 1], _peer(group, rank + 1), group))
        ops.append(dist.P2POp(dist.irecv, faces_nb[2 * own + 2], _peer(group, rank + 1), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return faces_nb


class PeerExchange:
    """
    NVLink peer-memory buffers of one rank for the d/dz exchange (comm="nvlink"): no NCCL on the data path.
    One symmetric allocation per rank (torch.distributed._symmetric_memory), laid out in doubles as
        halo  [2 parities][2][plane]   slot 0: last plane of rank-1, slot 1: first plane of rank+1
        faces [2 parities][6][plane]   the neighbour-only interface buffer of cfd_apply_coupled_nb
        flags [16]                     uint64 arrival counters: 0/1 halo from left/right, 2/3 face from left/right
    Neighbours write into it with plain stores from libcfd_b200's kernels (cfd_push_planes, cfd_edge_faces_p2p)
    and raise the flags to the call number; consumers wait with cfd_wait_flags.  Two parities + the monotone
    call number make buffer re-use safe without any global barrier.
    """

    def __init__(self, plane, rank, size, group, device):
        import torch.distributed._symmetric_memory as symm
        self.plane, self.rank, self.size = int(plane), rank, size
        total = 4 * self.plane + 12 * self.plane + 16
        self.buf = symm.empty(total, dtype=torch.float64, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        dist.barrier(group)                      # every rank's flags are zero before anybody pushes
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.seq = 0

    # byte addresses inside rank r's buffer
    def halo(self, r, parity, slot):
        return self.ptrs[r] + 8 * ((parity * 2 + slot) * self.plane)

    def faces(self, r, parity, i):
        return self.ptrs[r] + 8 * (4 * self.plane + (parity * 6 + i) * self.plane)

    def flag(self, r, k):
        return self.ptrs[r] + 8 * (16 * self.plane + k)

    def local_halo(self, parity, slot):
        o = (parity * 2 + slot) * self.plane
        return self.buf[o:o + self.plane]

    def local_faces(self, parity, n):
        o = 4 * self.plane + parity * 6 * self.plane
        return self.buf[o:o + n * self.plane]


class PartitionedDerivative:
    """
    Derivative along `direction` of one block of a field whose lines along `direction` are cut into `size` blocks,
    one per rank of `group` (the reference's line communicator, code/cuda/gpuDA.py:154-180 `get_line_DA`; its dfdx
    runs on exactly such a line, code/cuda/compact.py:18-27).  `local_shape` is this rank's block [nz, ny, nx]; the
    rank order inside `group` is the block order along the line.  With a Cartesian process grid (grid.DA) every
    direction gets its own line groups; ZPartitionedDerivative below is the z-slab special case of BASELINE configs[3].
    """
    _part_axis = None               # None: the partitioned axis is `direction` itself

    def __init__(self, local_shape, spacing, direction, group=None, mode="fused", comm="allgather"):
        """
        mode "fused"     : edge faces -> exchange -> ONE coupled kernel (final derivative, no correction pass)
             "reference" : local solve -> pack -> all-gather -> correction sweep (the reference's order)
        comm "allgather" : every rank receives all 2P interface planes (the reference's Gather+Scatter, rootless)
             "pairwise"  : one interface plane from each line neighbour by NCCL send/recv (exact in fp64 for blocks
                           >= 64 rows; fused mode only)
             "nvlink"    : the same neighbour-only data flow, but the kernels store halo and interface planes
                           straight into the neighbours' memory over NVLink/NVSwitch (symmetric memory) and
                           synchronise with flags -- no NCCL call on the data path (fused mode, z lines only:
                           their boundary planes are contiguous; other directions use "pairwise").  ONE producer
                           launch (cfd_edge_faces_push: faces without the neighbour points + own boundary rows into
                           the neighbours' buffers) and one consumer (cfd_reduced_unknowns_deferred: folds the halo
                           terms in); "nvlink-2step" keeps the first protocol (halo push, wait, edge faces, reduce)
        """
        assert dist.is_initialized(), "torch.distributed must be initialised (one process per GPU)"
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.direction = int(direction)
        self.local_shape = tuple(int(s) for s in local_shape)
        part = (self.rank, self.size) if self._partitioned else (0, 1)
        self.solver = CompactFiniteDifferenceSolver(self.local_shape, spacing, self.direction, part=part)
        assert mode in ("fused", "reference") and comm in ("allgather", "pairwise", "nvlink", "nvlink-2step")
        self._two_step = comm == "nvlink-2step"     # halo push + wait before the edge kernel (the first protocol)
        if self._two_step:
            comm = "nvlink"
        self._peer = None
        self.mode = mode if self.local_shape[self._dim] >= 66 else "reference"
        self.comm = comm if self.mode == "fused" else "allgather"
        if self.comm == "nvlink" and self._dim != 0:
            self.comm = "pairwise"
        self._buf = None
        self._side = None          # (stream, event) of an exchange started early by begin()
        self._pending = None

    # -- geometry of the line ------------------------------------------------------------------------------
    @property
    def _partitioned(self):
        axis = self.direction if self._part_axis is None else self._part_axis
        return axis == self.direction

    @property
    def _dim(self):
        """Tensor dimension of [nz, ny, nx] the derivative runs along (direction 0 = x = last dimension)."""
        return 2 - self.direction

    @property
    def plane_shape(self):
        """Shape of one boundary plane: the block with the derivative axis removed (one value per line)."""
        return tuple(s for d, s in enumerate(self.local_shape) if d != self._dim)

    def _ends(self, f):
        """First and last plane of the block along the line, contiguous (z planes are views, x / y planes are packed:
        the face packs of code/cuda/gpuDA.py:76-83)."""
        first, last = f.select(self._dim, 0), f.select(self._dim, f.shape[self._dim] - 1)
        return first.contiguous(), last.contiguous()

    def _buffers(self, f):
        if self._buf is None or self._buf[0].device != f.device:
            ps = self.plane_shape
            mk = lambda *s: torch.empty(s, dtype=torch.float64, device=f.device)  # noqa: E731
            pv, own = self.solver.nb_layout() if self.size > 1 and self._partitioned else (1, 0)
            faces_nb = torch.zeros((2 * pv,) + ps, dtype=torch.float64, device=f.device)
            self._buf = (mk(*ps), mk(*ps), mk(2, *ps), mk(2 * self.size, *ps), faces_nb, pv, own)
            self._ab = mk(2, *ps)
        return self._buf

    def _exchange_nvlink(self, f):
        """Halo push -> edge faces written into the neighbours' buffers -> flags; all stream-ordered kernels."""
        import ctypes
        from ._lib import check, lib
        nz, ny, nx = self.local_shape
        if self._peer is None:
            self._peer = PeerExchange(ny * nx, self.rank, self.size, self.group, f.device)
        px, r, P = self._peer, self.rank, self.size
        px.seq += 1
        seq, par = px.seq, px.seq & 1
        pv, own = self.solver.nb_layout()
        stream = ctypes.c_void_p(torch.cuda.current_stream(f.device).cuda_stream)
        left, right = (r - 1 if r > 0 else None), (r + 1 if r < P - 1 else None)
        L = lib()
        if not getattr(self, "_two_step", False):
            halo_lo = px.local_halo(par, 0) if left is not None else None
            halo_hi = px.local_halo(par, 1) if right is not None else None
            faces_nb = px.local_faces(par, 2 * pv)
            own_left = 1 if (left is not None and left > 0) else 0
            plan = self.solver._plan(self.solver.direction, self.solver.spacing)
            check(L.cfd_edge_faces_push(
                plan.handle, f.data_ptr(), faces_nb.data_ptr() + 8 * 2 * own * px.plane,
                px.faces(left, par, 2 * own_left + 2) if left is not None else None,
                px.faces(right, par, 1) if right is not None else None,
                px.halo(left, par, 1) if left is not None else None,
                px.halo(right, par, 0) if right is not None else None,
                px.flag(left, 3) if left is not None else None, px.flag(right, 2) if right is not None else None,
                seq, stream))
            self._buffers(f)
            check(L.cfd_reduced_unknowns_deferred(
                plan.handle, faces_nb.data_ptr(),
                halo_lo.data_ptr() if halo_lo is not None else None,
                halo_hi.data_ptr() if halo_hi is not None else None, f.data_ptr(), self._ab.data_ptr(),
                px.flag(r, 2) if left is not None else None, px.flag(r, 3) if right is not None else None,
                seq, stream))
            return halo_lo, halo_hi, self._ab
        check(L.cfd_push_planes(
            f[0].data_ptr() if left is not None else None, px.halo(left, par, 1) if left is not None else None,
            f[-1].data_ptr() if right is not None else None, px.halo(right, par, 0) if right is not None else None,
            ny * nx, px.flag(left, 1) if left is not None else None, px.flag(right, 0) if right is not None else None,
            seq, stream))
        check(L.cfd_wait_flags(px.flag(r, 0) if left is not None else None,
                               px.flag(r, 1) if right is not None else None, seq, stream))
        halo_lo = px.local_halo(par, 0) if left is not None else None
        halo_hi = px.local_halo(par, 1) if right is not None else None
        faces_nb = px.local_faces(par, 2 * pv)
        own_left = 1 if (left is not None and left > 0) else 0
        plan = self.solver._plan(self.solver.direction, self.solver.spacing)
        check(L.cfd_edge_faces_p2p(
            plan.handle, f.data_ptr(),
            halo_lo.data_ptr() if halo_lo is not None else None, halo_hi.data_ptr() if halo_hi is not None else None,
            faces_nb.data_ptr() + 8 * 2 * own * px.plane,
            px.faces(left, par, 2 * own_left + 2) if left is not None else None,
            px.faces(right, par, 1) if right is not None else None,
            px.flag(left, 3) if left is not None else None, px.flag(right, 2) if right is not None else None,
            seq, stream))
        self._buffers(f)
        # wait for the neighbours' interface planes inside the (tiny) reduced-solve kernel, then alpha / beta
        check(L.cfd_reduced_unknowns(plan.handle, faces_nb.data_ptr(), 1, self._ab.data_ptr(),
                                     px.flag(r, 2) if left is not None else None,
                                     px.flag(r, 3) if right is not None else None, seq, stream))
        return halo_lo, halo_hi, self._ab

    def _exchange(self, f):
        """Steps (1)-(3) of the fused path; returns (halo_lo, halo_hi, alpha/beta planes) for the coupled kernel."""
        if self.comm == "nvlink" and self._peer is None:
            # symmetric memory is a torch-internal API: if it cannot be set up on this system, every rank falls
            # back (collectively -- the outcome is agreed by all-reduce) to the NCCL send/recv exchange
            nz, ny, nx = self.local_shape
            ok = 1
            try:
                self._peer = PeerExchange(ny * nx, self.rank, self.size, self.group, f.device)
            except Exception as e:                                   # pragma: no cover - depends on the box
                import sys
                print(f"[cfd_b200] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL send/recv",
                      file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=f.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if flag.item() == 0:
                self._peer = None
                self.comm = "pairwise"
        if self.comm == "nvlink":
            return self._exchange_nvlink(f)
        lo_buf, hi_buf, faces, faces_all, faces_nb, pv, own = self._buffers(f)
        first, last = self._ends(f)
        halo_lo, halo_hi = exchange_halo_planes(first, last, self.rank, self.size, self.group, lo_buf, hi_buf)
        if self.comm == "pairwise":
            self.solver.edge_faces(f, faces_nb[2 * own:2 * own + 2], halo_lo, halo_hi)
            exchange_interface_planes(faces_nb, own, pv, self.rank, self.size, self.group)
            self.solver.reduced_unknowns(faces_nb, self._ab, neighbours_only=True)
            return halo_lo, halo_hi, self._ab
        self.solver.edge_faces(f, faces, halo_lo, halo_hi)
        gather_interface_planes(faces, self.size, self.group, faces_all)
        self.solver.reduced_unknowns(faces_all, self._ab)
        return halo_lo, halo_hi, self._ab

    def begin(self, f):
        """Start the halo / interface exchange on a side stream so that it overlaps whatever the caller launches
        next on the current stream (typically the derivatives of the same field along the other directions).  The
        next __call__ with the same f picks the result up.  No-op where there is nothing to exchange."""
        if not self._partitioned or self.size == 1 or self.mode != "fused":
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=f.device)
        cur = torch.cuda.current_stream(f.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            res = self._exchange(f)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._pending = (f.data_ptr(), res, ev)

    def __call__(self, f, out=None):
        if not self._partitioned or self.size == 1:
            return self.solver(f, out)
        if self.mode == "fused":
            if self._pending is not None and self._pending[0] == f.data_ptr():
                _, (halo_lo, halo_hi, planes), ev = self._pending
                torch.cuda.current_stream(f.device).wait_event(ev)
            else:
                halo_lo, halo_hi, planes = self._exchange(f)
            self._pending = None
            return self.solver.apply_coupled(f, out, halo_lo, halo_hi, planes)
        lo_buf, hi_buf, faces, faces_all = self._buffers(f)[:4]
        first, last = self._ends(f)
        halo_lo, halo_hi = exchange_halo_planes(first, last, self.rank, self.size, self.group, lo_buf, hi_buf)
        out = self.solver.apply_local(f, out, halo_lo, halo_hi)
        self.solver.interface_pack(out, faces)
        gather_interface_planes(faces, self.size, self.group, faces_all)
        self.solver.reduced_correct(out, faces_all)
        return out


class ZPartitionedDerivative(PartitionedDerivative):
    """Derivative of a z-partitioned field; `local_shape` is this rank's slab [nz/P, ny, nx].  d/dx and d/dy
    (direction 0, 1) never leave the slab; d/dz is the partitioned line."""
    _part_axis = 2
