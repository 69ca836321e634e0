"""
CompactFiniteDifferenceSolver -- the reference's derivative operator
(code/cuda/compact.py:16-44 `dfdx(f_d, dx, x_d, f_local_d)`; code/ocl/compact.py:26-61 `dfdx/dfdy/dfdz(f, d)`),
served by one fused sm_100a kernel per direction (RHS stencil + tridiagonal solve, f read once, f' written once).

    s  = CompactFiniteDifferenceSolver((nz, ny, nx), spacing=dx, direction=0)   # 0 = x, 1 = y, 2 = z
    df = s(f)                                  # f: CUDA float64 tensor [nz, ny, nx] (contiguous)
    s.dfdx(f, dx), s.dfdy(f, dy), s.dfdz(f, dz)   # reference spellings; plans are cached per (axis, spacing)

A NumPy array in gives a NumPy array out (the OpenCL flavour's contract) through the library's host-buffer
entry point: the arithmetic still runs on the GPU -- there is no CPU implementation in this package.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import check, lib

_AXIS_NAMES = {0: "x", 1: "y", 2: "z"}
#: compact schemes of cfd_create_scheme (include/cfd_b200.h): the reference's 4th-order Pade first derivative, the
#: 6th-order tridiagonal first derivative (two chunks of look-ahead), the 4th-order Pade SECOND derivative
SCHEMES = {"pade4": 0, "compact6": 1, "pade4-d2": 2}


def _stream_ptr(t):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class _Plan:
    """Owns one cfd_plan (cfd_create / cfd_destroy)."""

    def __init__(self, shape, axis, spacing, part_rank=0, part_size=1, scheme="pade4", npts=False):
        nz, ny, nx = (int(s) for s in shape)
        self.handle = ctypes.c_void_p()
        if npts:                        # this rank's slice of the LU of the whole line (distributed npts)
            assert scheme == "pade4"
            check(lib().cfd_create_npts(ctypes.byref(self.handle), nz, ny, nx, int(axis), float(spacing),
                                        int(part_rank), int(part_size)))
        elif scheme == "pade4":
            check(lib().cfd_create(ctypes.byref(self.handle), nz, ny, nx, int(axis), float(spacing),
                                   int(part_rank), int(part_size)))
        else:
            assert int(part_size) == 1, "schemes other than 'pade4' serve unpartitioned lines"
            check(lib().cfd_create_scheme(ctypes.byref(self.handle), nz, ny, nx, int(axis), float(spacing),
                                          SCHEMES[scheme]))
        self.scheme = scheme
        self.shape, self.axis, self.spacing = (nz, ny, nx), int(axis), float(spacing)
        self.part = (int(part_rank), int(part_size))
        self.plane_elems = lib().cfd_plane_elems(self.handle)

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                lib().cfd_destroy(h)
            except Exception:
                pass
            self.handle = None


class LineDA:
    """The handful of attributes of the reference's line distributed-array object that the derivative operator
    reads (code/cuda/gpuDA.py:154-180 `get_line_DA`; used at code/cuda/compact.py:49,159-166): local block shape and
    the position of the block along the derivative line."""

    def __init__(self, local_dims, rank=0, size=1, direction=0):
        self.nz, self.ny, self.nx = (int(s) for s in local_dims)
        self.rank, self.size, self.direction = int(rank), int(size), int(direction)
        self.mx, self.npx = self.rank, self.size            # names used by the reference kernels (kernels.cu:36,42)


class CompactFiniteDifferenceSolver:
    def __init__(self, shape, spacing=None, direction=None, part=(0, 1), solver=None, scheme="pade4"):
        """
        Reference spelling: ``CompactFiniteDifferenceSolver(line_da, solver='templated')`` (code/cuda/compact.py:18)
        with any object carrying ``nz, ny, nx, rank, size`` (e.g. :class:`LineDA`); `solver` is accepted and
        ignored (there is one kernel).  Native spelling:

        :param shape: (nz, ny, nx) of the (local) block, C order, x fastest
        :param spacing: grid spacing along `direction` (may instead be given per call to dfdx/dfdy/dfdz)
        :param direction: 0 = x, 1 = y, 2 = z (numbering of code/cuda/gpuDA.py:162)
        :param part: (rank, size) of this block along the derivative line -- the reference's
                     (line_da.rank, line_da.size), code/cuda/compact.py:159-166.  (0, 1) = whole line here.
        :param scheme: "pade4" (the reference's scheme, default), "compact6" (6th-order first derivative) or
                       "pade4-d2" (4th-order SECOND derivative: dfdx / dfdy / dfdz then return d2f/dx2 ...), see SCHEMES
        """
        assert scheme in SCHEMES, f"scheme is one of {sorted(SCHEMES)}"
        self.scheme = scheme
        self._group = None
        if hasattr(shape, "nz") and hasattr(shape, "rank"):            # a line_da of the reference
            da = shape
            part = (da.rank, da.size)
            if direction is None:
                direction = getattr(da, "direction", 0)
            self._group = getattr(da, "group", None)                   # DA.get_line_DA: the line's process group
            shape = (da.nz, da.ny, da.nx)
        assert len(shape) == 3, "shape is (nz, ny, nx)"
        self.shape = tuple(int(s) for s in shape)
        self.part = (int(part[0]), int(part[1]))
        self.direction = None if direction is None else int(direction)
        self.spacing = None if spacing is None else float(spacing)
        self._plans = {}
        self._last_h = None          # spacing of the last call that was given one (the reference passes dx per call)
        self._line_ops = {}          # (axis, spacing) -> PartitionedDerivative of a multi-rank line
        if self.direction is not None:
            assert self.direction in (0, 1, 2), "direction is 0 (x), 1 (y) or 2 (z)"
            if self.spacing is not None:
                self._plan(self.direction, self.spacing)          # fail early, like the reference constructor

    # -- plans ---------------------------------------------------------------------------------------
    def _plan(self, axis, spacing) -> _Plan:
        key = (int(axis), float(spacing))
        p = self._plans.get(key)
        if p is None:
            p = _Plan(self.shape, axis, spacing, *self.part, scheme=self.scheme)
            self._plans[key] = p
        return p

    # -- the hot path --------------------------------------------------------------------------------
    def _apply(self, axis, spacing, f, out=None, halo_lo=None, halo_hi=None, local=False):
        if self.part[1] > 1 and not local and axis == self.direction:
            # The reference's multi-rank call shape (code/cuda/compact.py:29-44): the block of a partitioned line in,
            # the derivative of the WHOLE line out -- halo exchange, reduced system and correction happen inside.
            return self._line_op(axis, spacing)(f, out)
        plan = self._plan(axis, spacing)
        if isinstance(f, np.ndarray):
            return self._apply_host(plan, f, out)
        import torch
        assert isinstance(f, torch.Tensor), "f is a CUDA float64 tensor (or a NumPy array for the host-buffer path)"
        assert f.is_cuda and f.dtype == torch.float64, "f must be a CUDA float64 tensor"
        assert tuple(f.shape) == self.shape, f"f has shape {tuple(f.shape)}, solver was built for {self.shape}"
        assert f.is_contiguous(), "f must be contiguous (C order, x fastest)"
        if out is None:
            out = torch.empty_like(f)
        else:
            assert out.is_cuda and out.dtype == torch.float64 and tuple(out.shape) == self.shape and out.is_contiguous()
            assert out.data_ptr() != f.data_ptr(), "the derivative is out of place"
        for h in (halo_lo, halo_hi):
            if h is not None:
                assert h.is_cuda and h.dtype == torch.float64 and h.is_contiguous() and h.numel() == plan.plane_elems
        check(lib().cfd_apply(plan.handle, f.data_ptr(), out.data_ptr(),
                              halo_lo.data_ptr() if halo_lo is not None else None,
                              halo_hi.data_ptr() if halo_hi is not None else None, _stream_ptr(f)))
        return out

    @staticmethod
    def _apply_host(plan, f, out):
        assert f.dtype == np.float64 and f.shape == plan.shape, "f must be float64 of the solver's shape"
        f = np.ascontiguousarray(f)
        if out is None:
            out = np.empty_like(f)
        assert out.dtype == np.float64 and out.shape == plan.shape and out.flags.c_contiguous
        check(lib().cfd_apply_host(plan.handle, f.ctypes.data, out.ctypes.data, 0))
        return out

    def __call__(self, f, out=None):
        assert self.direction is not None and self.spacing is not None, \
            "construct with spacing= and direction= to call the solver directly"
        return self._apply(self.direction, self.spacing, f, out)

    def apply_local(self, f, out=None, halo_lo=None, halo_hi=None):
        """Block-local solution x_R of a partitioned line (code/cuda/compact.py:46-64): needs the neighbour
        planes of f where this block does not own the physical end."""
        return self._apply(self.direction, self._h(None, self.direction), f, out, halo_lo, halo_hi, local=True)

    def _line_op(self, axis, spacing):
        """PartitionedDerivative over this block's line group (the line_da's, or the default group when it has
        exactly the line's size): the reference's multi-rank dfdx as one object, built on first use."""
        key = (int(axis), float(spacing))
        op = self._line_ops.get(key)
        if op is None:
            import torch.distributed as dist
            from .partition import PartitionedDerivative
            assert dist.is_initialized(), \
                "a block of a multi-rank line needs torch.distributed (one process per GPU) -- or call apply_local / " \
                "the stage methods with explicit halo planes"
            assert dist.get_world_size(self._group) == self.part[1] and dist.get_rank(self._group) == self.part[0], \
                "the line's process group does not match (rank, size) of the line_da"
            # z lines exchange over NVLink peer memory (cfd_zpart), x / y lines (strided boundary planes) over NCCL
            op = PartitionedDerivative(self.shape, spacing, axis, group=self._group, mode="fused",
                                       comm="nvlink" if axis == 2 else "pairwise")
            self._line_ops[key] = op
        return op

    def dfdxy(self, f, dx, dy, out_x=None, out_y=None, warps=None):
        """d/dx and d/dy of f in ONE launch (cfd_apply_xy): the two derivatives share the HBM reads of f through L2.
        warps: warps per SM of the launch (None = library default); 5 leaves room for kernels running beside it."""
        import torch
        px, py = self._plan(0, float(dx)), self._plan(1, float(dy))
        if warps is not None and getattr(px, "xy_warps", None) != warps:
            check(lib().cfd_plan_set_xy_warps(px.handle, int(warps)))
            px.xy_warps = warps
        assert f.is_cuda and f.dtype == torch.float64 and f.is_contiguous() and tuple(f.shape) == self.shape
        out_x = torch.empty_like(f) if out_x is None else out_x
        out_y = torch.empty_like(f) if out_y is None else out_y
        check(lib().cfd_apply_xy(px.handle, py.handle, f.data_ptr(), out_x.data_ptr(), out_y.data_ptr(), _stream_ptr(f)))
        return out_x, out_y

    def gradient(self, f, spacings, out=None):
        """(df/dx, df/dy, df/dz) with spacings = (dx, dy, dz): x and y in one launch, z in a second one on a side
        stream of the library (cfd_apply_xyz: forked and joined with events, ordered on the current stream)."""
        import torch
        dx, dy, dz = spacings
        assert f.is_cuda and f.dtype == torch.float64 and f.is_contiguous() and tuple(f.shape) == self.shape
        out = [torch.empty_like(f) if o is None else o for o in (out if out is not None else (None, None, None))]
        if self.part != (0, 1):
            gx, gy = self.dfdxy(f, dx, dy, out[0], out[1])
            return gx, gy, self.dfdz(f, dz, out[2])
        px, py, pz = self._plan(0, float(dx)), self._plan(1, float(dy)), self._plan(2, float(dz))
        check(lib().cfd_apply_xyz(px.handle, py.handle, pz.handle, f.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                                  out[2].data_ptr(), _stream_ptr(f)))
        return tuple(out)

    # reference spellings (code/ocl/compact.py:26,41,52; code/cuda/compact.py:29)
    def dfdx(self, f, dx=None, out=None, f_local=None):
        """f_local (the reference's ghosted scratch array) is accepted and ignored: no ghost copy is made."""
        return self._apply(0, self._h(dx, 0), f, out)

    def dfdy(self, f, dy=None, out=None, f_local=None):
        return self._apply(1, self._h(dy, 1), f, out)

    def dfdz(self, f, dz=None, out=None, f_local=None):
        return self._apply(2, self._h(dz, 2), f, out)

    def _h(self, h, axis):
        if h is not None:
            self._last_h = float(h)
            return float(h)
        if self.spacing is not None and self.direction == axis:
            return self.spacing
        assert self._last_h is not None, f"spacing along {_AXIS_NAMES[axis]} not given"
        return self._last_h

    def _stage_plan(self):
        """Plan for the stages that do not depend on the spacing (primary / secondary / reduced systems, sum): the
        reference constructor takes no spacing (code/cuda/compact.py:18), so neither do they."""
        axis = self._axis()
        h = self.spacing if self.spacing is not None else (self._last_h if self._last_h is not None else 1.0)
        return self._plan(axis, h)

    # -- the reference's stage methods (code/cuda/compact.py:46-154), each on its own kernel ------------------
    def compute_RHS(self, f, dx=None, x=None, f_local=None, halo_lo=None, halo_hi=None):
        """x = Pade right-hand side of f (compact.py:46-51).  f_local (ghosted scratch) is accepted and ignored;
        blocks that do not own a physical end take the neighbour planes as halo_lo / halo_hi."""
        import torch
        h = self._h(dx, self._axis())
        plan = self._plan(self._axis(), h)
        if x is None:
            x = torch.empty_like(f)
        check(lib().cfd_compute_rhs(plan.handle, f.data_ptr(), x.data_ptr(),
                                    halo_lo.data_ptr() if halo_lo is not None else None,
                                    halo_hi.data_ptr() if halo_hi is not None else None, _stream_ptr(f)))
        return x

    def solve_primary_system(self, x):
        """In-place block-local tridiagonal solve of the right-hand sides in x (compact.py:62-64)."""
        from .near_toeplitz import NearToeplitzSolver
        if getattr(self, "_primary", None) is None:
            co = (ctypes.c_double * 7)()
            check(lib().cfd_plan_coeffs(self._stage_plan().handle, co))
            self._primary = NearToeplitzSolver(self.shape, list(co), axis=self._axis())
        return self._primary.solve(x)

    def solve_secondary_systems(self):
        """(x_UH, x_LH): unit responses of the block matrix (compact.py:128-154), as CUDA tensors."""
        import torch
        n = self.shape[2 - self._axis()]
        dp = ctypes.POINTER(ctypes.c_double)
        xu, xl = np.zeros(n), np.zeros(n)
        check(lib().cfd_plan_secondary(self._stage_plan().handle, xu.ctypes.data_as(dp), xl.ctypes.data_as(dp),
                                       None, None, None))
        return torch.from_numpy(xu).cuda(), torch.from_numpy(xl).cuda()

    def solve_reduced_system(self, x_UH, x_LH, x_R, group=None):
        """(alpha, beta) planes of this block from the block-local solution x_R (compact.py:65-126): interface faces
        (negateAndCopyFaces) -> all ranks of the line (the reference gathers to the line root, solves there and
        scatters; here every rank receives the 2P planes and solves for its own two unknowns) -> reduced solve.
        x_UH / x_LH are accepted for the reference's signature; the reduced matrix was built from them at plan
        creation."""
        import torch
        import torch.distributed as dist
        from .partition import gather_interface_planes
        plan = self._stage_plan()
        group = self._group if group is None else group
        ps = tuple(s for d, s in enumerate(self.shape) if d != 2 - self._axis())
        faces = torch.empty((2,) + ps, dtype=torch.float64, device=x_R.device)
        check(lib().cfd_interface_pack(plan.handle, x_R.data_ptr(), faces.data_ptr(), _stream_ptr(x_R)))
        assert dist.is_initialized() and dist.get_world_size(group) == self.part[1], \
            "solve_reduced_system gathers over the line's process group"
        faces_all = gather_interface_planes(faces, self.part[1], group)
        ab = torch.empty((2,) + ps, dtype=torch.float64, device=x_R.device)
        check(lib().cfd_reduced_unknowns(plan.handle, faces_all.data_ptr(), 0, ab.data_ptr(), None, None, 0,
                                         _stream_ptr(x_R)))
        return ab[0], ab[1]

    def sum_solutions(self, *args):
        """x_R += alpha * x_UH + beta * x_LH over the whole block (compact.py:52-61).  Reference signature
        ``sum_solutions(x_UH, x_LH, x_R, alpha, beta)``; the short form ``sum_solutions(x_R, alpha, beta)`` uses the
        plan's own secondary solutions (which is what the five-argument form receives from solve_secondary_systems)."""
        if len(args) == 5:
            _, _, x_R, alpha, beta = args
        else:
            x_R, alpha, beta = args
        alpha, beta = alpha.contiguous(), beta.contiguous()
        check(lib().cfd_sum_solutions(self._stage_plan().handle, x_R.data_ptr(), alpha.data_ptr(), beta.data_ptr(),
                                      _stream_ptr(x_R)))
        return x_R

    def _axis(self):
        assert self.direction is not None, "construct with direction= (or a line_da) to use the stage methods"
        return self.direction

    # -- multi-rank pieces (used by partition.ZPartitionedDerivative) -----------------------------------
    def interface_pack(self, df, faces):
        plan = self._plan(self.direction, self._h(None, self.direction))
        check(lib().cfd_interface_pack(plan.handle, df.data_ptr(), faces.data_ptr(), _stream_ptr(df)))
        return faces

    def edge_faces(self, f, faces, halo_lo=None, halo_hi=None):
        """Interface planes straight from f (no block solve): cfd_edge_faces."""
        plan = self._plan(self.direction, self._h(None, self.direction))
        check(lib().cfd_edge_faces(plan.handle, f.data_ptr(),
                                   halo_lo.data_ptr() if halo_lo is not None else None,
                                   halo_hi.data_ptr() if halo_hi is not None else None,
                                   faces.data_ptr(), _stream_ptr(f)))
        return faces

    def reduced_unknowns(self, faces, ab, neighbours_only=False, flag0=None, flag1=None, seq=0):
        """alpha / beta planes of this rank from the gathered interface planes: cfd_reduced_unknowns."""
        plan = self._plan(self.direction, self._h(None, self.direction))
        check(lib().cfd_reduced_unknowns(plan.handle, faces.data_ptr(), 1 if neighbours_only else 0, ab.data_ptr(),
                                         flag0, flag1, int(seq), _stream_ptr(faces)))
        return ab

    def apply_coupled(self, f, out, halo_lo, halo_hi, ab):
        """Final derivative of the block in one pass, interface unknowns folded in: cfd_apply_coupled."""
        import torch
        plan = self._plan(self.direction, self._h(None, self.direction))
        if out is None:
            out = torch.empty_like(f)
        check(lib().cfd_apply_coupled(plan.handle, f.data_ptr(), out.data_ptr(),
                                      halo_lo.data_ptr() if halo_lo is not None else None,
                                      halo_hi.data_ptr() if halo_hi is not None else None,
                                      ab.data_ptr(), _stream_ptr(f)))
        return out

    def nb_layout(self):
        """(virtual ranks V, own index) of the neighbour-only interface buffer [2V, plane]: cfd_nb_layout."""
        plan = self._plan(self.direction, self._h(None, self.direction))
        pv, own = ctypes.c_int(), ctypes.c_int()
        check(lib().cfd_nb_layout(plan.handle, ctypes.byref(pv), ctypes.byref(own)))
        return pv.value, own.value

    def reduced_correct(self, df, faces_all):
        plan = self._plan(self.direction, self._h(None, self.direction))
        check(lib().cfd_reduced_correct(plan.handle, df.data_ptr(), faces_all.data_ptr(), _stream_ptr(df)))
        return df
