"""Profiling driver: N^3 field, `reps` gradient steps (one stream_kernel_xy launch + one d/dz stream_kernel launch each).
Used under ncu; prints the CUDA-event time per launch when run plain."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
h = 2 * np.pi / (N - 1)
t = torch.arange(N, dtype=torch.float64, device="cuda") * h
f = (torch.sin(t)[None, None, :] * torch.cos(t)[None, :, None] * torch.sin(t)[:, None, None]).contiguous()
out = [torch.empty_like(f) for _ in range(3)]
sol = C.CompactFiniteDifferenceSolver((N, N, N))
for r in range(reps):
    sol.gradient(f, (h, h, h), out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
for _ in range(10):
    sol.dfdxy(f, h, h, out[0], out[1])
ev[1].record()
for _ in range(10):
    sol.dfdz(f, h, out[2])
ev[2].record()
torch.cuda.synchronize()
print("ms per launch xy, z:", [ev[a].elapsed_time(ev[a + 1]) / 10 for a in range(2)])
