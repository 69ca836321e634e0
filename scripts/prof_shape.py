"""Profiling driver for one shape: a few launches per axis."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C
shape = tuple(int(a) for a in sys.argv[1:4])
f = torch.rand(shape, dtype=torch.float64, device="cuda")
df = torch.empty_like(f)
for r in range(3):
    for a in range(3):
        C.CompactFiniteDifferenceSolver(shape, 0.1, a)(f, df)
torch.cuda.synchronize()
