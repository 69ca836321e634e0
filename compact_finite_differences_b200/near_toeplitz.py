"""
NearToeplitzSolver -- batched solve of the near-Toeplitz tridiagonal system

    b1 c1
    ai bi ci
       ai bi ci
          ...
             an bn

for every line of a [nz, ny, nx] array, IN PLACE, exactly the reference's
`NearToeplitzSolver(shape, coeffs).solve(x_d)` (code/cuda/solvers/templated/near_toeplitz.py:36-107),
with `coeffs = [b1, c1, ai, bi, ci, an, bn]`.  The reference solves along x only, needs nx a power of two
<= 2048; here any axis (keyword `axis`, default 0 = x) and any n >= 3.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import check, lib


class NearToeplitzSolver:
    def __init__(self, shape, coeffs, axis=0):
        assert len(shape) == 3, "shape is (nz, ny, nx)"
        assert len(coeffs) == 7, "coeffs = [b1, c1, ai, bi, ci, an, bn]"
        self.nz, self.ny, self.nx = (int(s) for s in shape)
        self.shape = (self.nz, self.ny, self.nx)
        self.coeffs = [float(c) for c in coeffs]
        self.axis = int(axis)
        self._handle = ctypes.c_void_p()
        co = (ctypes.c_double * 7)(*self.coeffs)
        check(lib().nt_create(ctypes.byref(self._handle), self.nz, self.ny, self.nx, self.axis, co))
        #: True when the matrix is not diagonally dominant enough for the one-pass kernel (exact two-pass LU instead)
        self.two_pass = bool(lib().nt_is_exact_two_pass(self._handle))

    def solve(self, x_d):
        """Solve in place: on entry x_d holds the right-hand sides, on exit the solutions."""
        import torch
        if isinstance(x_d, np.ndarray):
            raise TypeError("NearToeplitzSolver.solve works in place on a CUDA float64 tensor")
        assert isinstance(x_d, torch.Tensor) and x_d.is_cuda and x_d.dtype == torch.float64
        assert tuple(x_d.shape) == self.shape and x_d.is_contiguous()
        stream = ctypes.c_void_p(torch.cuda.current_stream(x_d.device).cuda_stream)
        check(lib().nt_solve(self._handle, x_d.data_ptr(), stream))
        return x_d

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                lib().nt_destroy(h)
            except Exception:
                pass
            self._handle = None
