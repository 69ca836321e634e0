"""Profiling driver: N^3 field, `reps` derivative launches per axis (x, y, z in turn).  Used under ncu."""
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if len(sys.argv) > 4:
    C.lib().cfd_set_launch(int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]) if len(sys.argv) > 5 else 0)
h = 2 * np.pi / (N - 1)
t = torch.arange(N, dtype=torch.float64, device="cuda") * h
f = (torch.sin(t)[None, None, :] * torch.cos(t)[None, :, None] * torch.sin(t)[:, None, None]).contiguous()
df = torch.empty_like(f)
ops = [C.CompactFiniteDifferenceSolver((N, N, N), h, a) for a in range(3)]
for r in range(reps):
    for a in range(3):
        ops[a](f, df)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for a in range(3):
    ev[a].record()
    for _ in range(10):
        ops[a](f, df)
ev[3].record()
torch.cuda.synchronize()
print("ms per launch x,y,z:", [ev[a].elapsed_time(ev[a + 1]) / 10 for a in range(3)])
