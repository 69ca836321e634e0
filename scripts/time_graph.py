"""Eager vs CUDA-graph replay of a gradient step (3 launches) on small grids."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C

for n in (32, 64, 128, 256):
    f = torch.rand((n, n, n), dtype=torch.float64, device="cuda")
    outs = [torch.empty_like(f) for _ in range(3)]
    ops = [C.CompactFiniteDifferenceSolver((n, n, n), 0.1, a) for a in range(3)]

    def step():
        for a in range(3):
            ops[a](f, outs[a])
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()

    def timeit(fn, reps=200):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    te, tg = timeit(step), timeit(g.replay)
    print(f"{n}^3 gradient step: eager {te:.1f} us, graph replay {tg:.1f} us "
          f"({3 * n ** 3 / tg * 1e6:.3e} pts/s per derivative)", flush=True)
