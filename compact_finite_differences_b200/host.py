"""
HostGradient -- the OpenCL flavour's contract (host array in, host arrays out: code/ocl/compact.py:26-61
`dfdx/dfdy/dfdz(f, d) -> ndarray`) for all three directions at once, with the host<->device copies pipelined
against the kernels:

    z-slab s of f:  H2D (copy stream)  ->  d/dx, d/dy of the slab (compute stream)  ->  D2H of both (copy-back stream)
    after the last slab:                    d/dz of the whole field                 ->  D2H

x and y lines never leave a z-slab, so their kernels start as soon as the first slab has landed and the
device->host copies of the results overlap the remaining host->device copies (PCIe is full duplex).  Only the
z derivative has to wait for the whole field.  All arithmetic runs in libcfd_b200's kernels.
"""
from __future__ import annotations

import numpy as np
import torch

from .compact import CompactFiniteDifferenceSolver


class HostGradient:
    @staticmethod
    def slab_cuts(nz, slabs=8, ramp=True):
        """[(first plane, end plane)] of the transfer pipeline.  The device->host copies are the bound of the call
        (three results for one field) and cannot start before the first slab has landed and been differentiated, so
        with `ramp` the first slabs are thin (nz/64 planes, doubling up to nz/slabs): that lag is an eighth of a
        uniform slab's."""
        slabs = max(1, min(int(slabs), nz))
        full = max(1, nz // slabs)
        cuts, step = [0], (max(4, nz // 64) if ramp else full)
        while cuts[-1] < nz:
            step = min(step, full)
            cuts.append(min(nz, cuts[-1] + step))
            step *= 2
        if len(cuts) > 2 and cuts[-1] - cuts[-2] < min(4, full):     # no sliver at the end
            del cuts[-2]
        return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]

    def __init__(self, shape, spacings, slabs=8, device=None, ddz=None, ramp=True):
        """
        :param shape: (nz, ny, nx)
        :param spacings: (dx, dy, dz)
        :param slabs: nz / slabs = planes per z-slab of the transfer pipeline
        :param ramp: start with thin slabs (see below); False = uniform slabs
        :param ddz: optional d/dz operator `ddz(f, out)` for the resident block, e.g. a ZPartitionedDerivative when
                    this block is one rank's slab of a z-partitioned field (d/dx, d/dy never leave the slab)
        """
        self.shape = tuple(int(s) for s in shape)
        nz, ny, nx = self.shape
        self.dx, self.dy, self.dz = (float(h) for h in spacings)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.slabs = self.slab_cuts(nz, slabs, ramp)
        self._solvers = {}
        for a, b in self.slabs:
            if (b - a) not in self._solvers:
                s = CompactFiniteDifferenceSolver((b - a, ny, nx))
                self._solvers[b - a] = s
        self._z = ddz if ddz is not None else CompactFiniteDifferenceSolver(self.shape, self.dz, 2)
        with torch.cuda.device(self.device):
            self._f = torch.empty(self.shape, dtype=torch.float64, device=self.device)
            self._d = [torch.empty(self.shape, dtype=torch.float64, device=self.device) for _ in range(3)]
            self._s_in, self._s_comp, self._s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self.h2d_bytes = self._f.numel() * 8
        self.d2h_bytes = 3 * self._f.numel() * 8

    @staticmethod
    def _as_tensor(a):
        if isinstance(a, np.ndarray):
            assert a.dtype == np.float64 and a.flags.c_contiguous
            return torch.from_numpy(a)
        assert isinstance(a, torch.Tensor) and a.dtype == torch.float64 and not a.is_cuda and a.is_contiguous()
        return a

    def __call__(self, f_host, out=None):
        """f_host: float64 host array/tensor [nz, ny, nx] (pinned memory gives full PCIe speed).
        Returns (dfdx, dfdy, dfdz) as host tensors (pinned if allocated here, or `out` = three host buffers)."""
        fh = self._as_tensor(f_host)
        assert tuple(fh.shape) == self.shape
        if out is None:
            out = [torch.empty(self.shape, dtype=torch.float64, pin_memory=True) for _ in range(3)]
        oh = [self._as_tensor(o) for o in out]
        s_in, s_comp, s_out = self._s_in, self._s_comp, self._s_out
        cur = torch.cuda.current_stream(self.device)
        for s in (s_in, s_comp, s_out):
            s.wait_stream(cur)
        f, d = self._f, self._d
        for a, b in self.slabs:
            with torch.cuda.stream(s_in):
                f[a:b].copy_(fh[a:b], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(ev_in)
                sol = self._solvers[b - a]
                sol.dfdxy(f[a:b], self.dx, self.dy, d[0][a:b], d[1][a:b])     # one launch, f read from HBM once
                ev_c = torch.cuda.Event()
                ev_c.record(s_comp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_c)
                oh[0][a:b].copy_(d[0][a:b], non_blocking=True)
                oh[1][a:b].copy_(d[1][a:b], non_blocking=True)
        with torch.cuda.stream(s_comp):
            self._z(f, d[2])
            ev_z = torch.cuda.Event()
            ev_z.record(s_comp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_z)
            oh[2].copy_(d[2], non_blocking=True)
        s_out.synchronize()
        cur.wait_stream(s_out)
        return tuple(out)
