"""
Multi-GPU derivative on a z-partition: one process per GPU, torch.distributed (NCCL over NVLink) for the two
exchanges the path really has, hand-written kernels for everything else.

The reference's multi-rank dfdx (code/cuda/compact.py:29-44) is
    halo exchange (gpuDA.global_to_local) -> RHS -> local solve -> gather interface faces to the line root ->
    reduced solve on the root -> scatter -> sum.
Here, for a grid split into P slabs along z:
    d/dx, d/dy : no communication at all (lines never leave the slab);
    d/dz       : (1) send/recv ONE boundary plane of f with each z-neighbour;
                 (2) interface planes -x_R[first], -x_R[last] from the first 32 and the last 32 planes of the slab
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


class ZPart:
    """
    One rank of a z-partitioned line with the host side in C (`cfd_zpart_*`, include/cfd_b200.h): the rank's receive
    buffer (halo planes, interface planes, arrival flags) is a plain cudaMalloc allocation of libcfd_b200, the two
    z-neighbours' buffers are mapped with CUDA IPC handles, and the kernels store into them over NVLink and
    synchronise with flags -- no NCCL, no torch-internal API on the data path.  torch.distributed only carries the
    64-byte handles once, at construction (any backend; the reference would use its MPI line communicator).
    """

    def __init__(self, plan, rank, size, group=None):
        import ctypes
        from ._lib import CFD_IPC_HANDLE_BYTES, check, lib
        self._lib, self._check = lib(), check
        self.handle = ctypes.c_void_p()
        check(self._lib.cfd_zpart_create(ctypes.byref(self.handle), plan.handle))
        self._plan = plan                       # keeps the cfd_plan alive
        buf = ctypes.create_string_buffer(CFD_IPC_HANDLE_BYTES)
        check(self._lib.cfd_zpart_export(self.handle, buf))
        handles = [None] * size
        dist.all_gather_object(handles, bytes(buf.raw), group=group)
        lo = ctypes.create_string_buffer(handles[rank - 1], CFD_IPC_HANDLE_BYTES) if rank > 0 else None
        hi = ctypes.create_string_buffer(handles[rank + 1], CFD_IPC_HANDLE_BYTES) if rank < size - 1 else None
        check(self._lib.cfd_zpart_connect(self.handle, lo, hi))
        self.group = group

    @staticmethod
    def _sp(f):
        import ctypes
        return ctypes.c_void_p(torch.cuda.current_stream(f.device).cuda_stream)

    def begin(self, f):
        self._check(self._lib.cfd_zpart_begin(self.handle, f.data_ptr(), self._sp(f)))

    def apply(self, f, out):
        self._check(self._lib.cfd_zpart_apply(self.handle, f.data_ptr(), out.data_ptr(), self._sp(f)))
        return out

    def apply_npts(self, f, out):
        self._check(self._lib.cfd_zpart_apply_npts(self.handle, f.data_ptr(), out.data_ptr(), self._sp(f)))
        return out

    def apply_xyz(self, plan_x, plan_y, f, out_x, out_y, out_z):
        self._check(self._lib.cfd_zpart_apply_xyz(self.handle, plan_x.handle, plan_y.handle, f.data_ptr(),
                                                  out_x.data_ptr(), out_y.data_ptr(), out_z.data_ptr(), self._sp(f)))
        return out_x, out_y, out_z

    def close(self):
        """Collective: no neighbour may still be storing into this rank's buffer when it is freed."""
        if self.handle:
            torch.cuda.synchronize()
            dist.barrier(self.group)
            self._lib.cfd_zpart_destroy(self.handle)
            self.handle = None


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
             "npts"      : the LANL distributed npts method instead of the reduced system (z lines, >= 66 planes per
                           slab): one LU of the whole line, u~ handed to the right and x~ to the left over NVLink,
                           both sweeps in one coupled pass (cfd_create_npts / cfd_zpart_apply_npts)
        comm "allgather" : every rank receives all 2P interface planes (the reference's Gather+Scatter, rootless)
             "pairwise"  : one interface plane from each line neighbour by NCCL send/recv (exact in fp64 for blocks
                           >= 64 rows; fused mode only)
             "nvlink"    : the same neighbour-only data flow, but the kernels store into the neighbours' memory over
                           NVLink/NVSwitch themselves (cfd_zpart: CUDA IPC mappings) -- no NCCL call on the data path
                           (fused mode, z lines only: their boundary planes are contiguous; other directions use
                           "pairwise").  The whole partitioned derivative is ONE launch (stream_kernel_zx: per bundle
                           edge faces, self-validating words into the neighbours' receive arrays, reduced system in
                           registers, coupled solve).  begin() keeps the three-launch form for callers that want the
                           exchange early on a side stream: cfd_edge_faces_push (faces without the neighbour points
                           + own boundary rows into the neighbours' buffers, one flag per side),
                           cfd_reduced_unknowns_deferred (folds the halo terms in), cfd_apply_coupled.  (The first,
                           two-step protocol lives on as C entry points: cfd_push_planes / cfd_wait_flags /
                           cfd_edge_faces_p2p.)
        """
        assert dist.is_initialized(), "torch.distributed must be initialised (one process per GPU)"
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.direction = int(direction)
        self.local_shape = tuple(int(s) for s in local_shape)
        part = (self.rank, self.size) if self._partitioned else (0, 1)
        self.solver = CompactFiniteDifferenceSolver(self.local_shape, spacing, self.direction, part=part)
        assert mode in ("fused", "reference", "npts") and comm in ("allgather", "pairwise", "nvlink")
        self._zp = None
        self._npts = None
        if mode == "npts":
            from .compact import _Plan
            assert self._partitioned and self.size > 1 and self._dim == 0 and self.local_shape[0] >= 66, \
                "mode 'npts' serves z-partitioned lines with >= 66 planes per slab"
            self._npts_plan = _Plan(self.local_shape, 2, spacing, self.rank, self.size, npts=True)
            self._npts = ZPart(self._npts_plan, self.rank, self.size, self.group)
            self.mode, self.comm = "npts", "nvlink"
            self._buf = self._side = self._pending = None
            return
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

    def _zpart(self, f):
        """The C-side exchange object of comm="nvlink", created collectively on first use.  CUDA IPC can be
        unavailable (containers without a shared PID / IPC namespace): the outcome is agreed by all-reduce and every
        rank of the line falls back to the NCCL send/recv exchange together."""
        if self._zp is None and self.comm == "nvlink":
            ok = 1
            try:
                plan = self.solver._plan(self.solver.direction, self.solver.spacing)
                self._zp = ZPart(plan, self.rank, self.size, self.group)
            except Exception as e:                                       # pragma: no cover - depends on the box
                import sys
                print(f"[cfd_b200] CUDA IPC peer mapping unavailable ({type(e).__name__}: {e}); using NCCL send/recv",
                      file=sys.stderr)
                ok = 0
            on_gpu = dist.get_backend(self.group) == "nccl"
            flag = torch.tensor([ok], dtype=torch.int32, device=f.device if on_gpu else "cpu")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if flag.item() == 0:
                if self._zp is not None:
                    self._zp.close()
                self._zp = None
                self.comm = "pairwise"
        return self._zp

    def _exchange(self, f):
        """Steps (1)-(3) of the fused path; returns (halo_lo, halo_hi, alpha/beta planes) for the coupled kernel."""
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
        if self._zpart(f) is not None:          # comm "nvlink": cfd_zpart orders the two streams with its own events
            if self._side is None:
                self._side = torch.cuda.Stream(device=f.device)
            self._side.wait_stream(torch.cuda.current_stream(f.device))     # f is ready on the caller's stream
            with torch.cuda.stream(self._side):
                self._zp.begin(f)
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
        if getattr(self, "_npts", None) is not None:
            if out is None:
                out = torch.empty_like(f)
            return self._npts.apply_npts(f, out)
        if not self._partitioned or self.size == 1:
            return self.solver(f, out)
        if self.mode == "fused" and self._zpart(f) is not None:
            if out is None:
                out = torch.empty_like(f)
            return self._zp.apply(f, out)
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

    def gradient(self, f, dx, dy, out=None):
        """(df/dx, df/dy, df/dz) of the slab.  With direction = 2 and comm = "nvlink": two launches
        (cfd_zpart_apply_xyz) -- the fused d/dx + d/dy kernel and the one-kernel partitioned d/dz.  Otherwise the
        exchange is started on a side stream and the same results come from separate calls."""
        assert self.direction == 2, "gradient() belongs to the d/dz operator of the slab"
        out = [torch.empty_like(f) if o is None else o for o in (out if out is not None else (None, None, None))]
        if getattr(self, "_xy", None) is None:
            self._xy = CompactFiniteDifferenceSolver(self.local_shape)
        if self.size > 1 and self.mode == "fused" and self._zpart(f) is not None:
            px, py = self._xy._plan(0, float(dx)), self._xy._plan(1, float(dy))
            self._zp.apply_xyz(px, py, f, out[0], out[1], out[2])
            return tuple(out)
        self.begin(f)
        self._xy.dfdxy(f, dx, dy, out[0], out[1])
        self(f, out[2])
        return tuple(out)

    def close(self):
        if self._zp is not None:
            self._zp.close()
            self._zp = None
        if getattr(self, "_npts", None) is not None:
            self._npts.close()
            self._npts = None
