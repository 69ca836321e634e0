"""
ReducedSolver -- thread-parallel Thomas ("pThomas") over interleaved systems that share one general
tridiagonal matrix: the reference's `ReducedSolver(shape).solve(a_d, b_d, c_d, c2_d, x_d)`
(code/cuda/reduced.py:5-18, kernel code/cuda/kernels.cu:115-145).  shape = (n, ny', nx'): n unknowns per
system, ny'*nx' systems, elements of one system ny'*nx' apart.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import check, lib


class ReducedSolver:
    def __init__(self, shape):
        self.nz, self.ny, self.nx = (int(s) for s in shape)
        self._plans = {}            # (a, b, c) bytes -> pt_plan: the matrix is eliminated once, not per call

    def _plan(self, host):
        key = b"".join(v.tobytes() for v in host)
        h = self._plans.get(key)
        if h is None:
            if len(self._plans) >= 16:
                self.close()
            dp = ctypes.POINTER(ctypes.c_double)
            h = ctypes.c_void_p()
            check(lib().cfd_pthomas_create(ctypes.byref(h), host[0].ctypes.data_as(dp), host[1].ctypes.data_as(dp),
                                           host[2].ctypes.data_as(dp), self.nz))
            self._plans[key] = h
        return h

    def close(self):
        for h in self._plans.values():
            lib().cfd_pthomas_destroy(h)
        self._plans = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, a, b, c, c2, x_d):
        """a, b, c: length-n diagonals (NumPy or tensors; a[0], c[-1] ignored).  c2 is the reference's scratch
        array for the modified upper diagonal -- accepted and ignored (pivots are precomputed on the host).
        x_d: CUDA float64 tensor [n, ny', nx'], solved in place."""
        import torch
        n = self.nz
        host = []
        for v in (a, b, c):
            v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            v = np.ascontiguousarray(v, dtype=np.float64)
            assert v.shape == (n,)
            host.append(v)
        assert x_d.is_cuda and x_d.dtype == torch.float64 and x_d.is_contiguous()
        assert x_d.numel() == n * self.ny * self.nx
        stream = ctypes.c_void_p(torch.cuda.current_stream(x_d.device).cuda_stream)
        check(lib().cfd_pthomas_solve(self._plan(host), x_d.data_ptr(), self.ny * self.nx, stream))
        return x_d
