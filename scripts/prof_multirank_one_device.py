"""Profiling driver (single process, single GPU) for the multi-rank kernels: an interior rank's slab of a 1024^2-plane
partition, halo / interface buffers local.  Used under ncu to get per-kernel metrics of edge_faces, push_planes,
reduced_planes and the coupled stream kernel."""
import ctypes
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C
from compact_finite_differences_b200._lib import check, lib

nz, ny, nx = 128, 1024, 1024
P, r = 8, 3
f = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda")
out = torch.empty_like(f)
lo, hi = torch.rand((ny, nx), dtype=torch.float64, device="cuda"), torch.rand((ny, nx), dtype=torch.float64, device="cuda")
s = C.CompactFiniteDifferenceSolver((nz, ny, nx), 0.01, 2, part=(r, P))
pv, own = s.nb_layout()
faces_nb = torch.zeros((2 * pv, ny, nx), dtype=torch.float64, device="cuda")
ab = torch.empty((2, ny, nx), dtype=torch.float64, device="cuda")
flags = torch.zeros(8, dtype=torch.int64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
plan = s._plan(2, 0.01)
for seq in range(1, 4):
    dst0, dst1 = torch.empty_like(lo), torch.empty_like(hi)
    check(lib().cfd_push_planes(f[0].data_ptr(), dst0.data_ptr(), f[-1].data_ptr(), dst1.data_ptr(), ny * nx,
                                flags.data_ptr(), flags.data_ptr() + 8, seq, st))
    check(lib().cfd_wait_flags(flags.data_ptr(), flags.data_ptr() + 8, seq, st))
    check(lib().cfd_edge_faces_p2p(plan.handle, f.data_ptr(), lo.data_ptr(), hi.data_ptr(),
                                   faces_nb.data_ptr() + 8 * 2 * own * ny * nx,
                                   faces_nb[2 * own - 1].data_ptr(), faces_nb[2 * own + 2].data_ptr(),
                                   flags.data_ptr() + 16, flags.data_ptr() + 24, seq, st))
    check(lib().cfd_reduced_unknowns(plan.handle, faces_nb.data_ptr(), 1, ab.data_ptr(),
                                     flags.data_ptr() + 16, flags.data_ptr() + 24, seq, st))
    s.apply_coupled(f, out, lo, hi, ab)
torch.cuda.synchronize()
print("ok")
