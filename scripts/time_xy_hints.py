"""ms of the fused d/dx + d/dy launch on slab shapes with the L2 eviction hints of the SEG variants (CFD_XY_HINTS:
bit 0 evict_last on tile loads, bit 1 evict_first on result stores), interleaved rounds.
usage: time_xy_hints.py [nz ny nx] ..."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C


def timeit(fn, reps=30):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


a = [int(v) for v in sys.argv[1:]]
shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)] or [(128, 1024, 1024), (256, 1024, 1024), (512, 512, 512)]
for shape in shapes:
    f = torch.rand(shape, dtype=torch.float64, device="cuda")
    ox, oy = torch.empty_like(f), torch.empty_like(f)
    s = C.CompactFiniteDifferenceSolver(shape)
    os.environ.pop("CFD_XY_HINTS", None)
    rx, ry = [o.clone() for o in s.dfdxy(f, 0.1, 0.1)]
    res = {h: [] for h in (0, 1, 2, 3)}
    same = {}
    for rnd in range(5):
        for h in res:
            os.environ["CFD_XY_HINTS"] = str(h)
            res[h].append(timeit(lambda: s.dfdxy(f, 0.1, 0.1, ox, oy)))
            torch.cuda.synchronize()
            same[h] = torch.equal(ox, rx) and torch.equal(oy, ry)
    os.environ.pop("CFD_XY_HINTS", None)
    n = f.numel()
    for h, r in res.items():
        r = sorted(r)
        print(f"{shape} hints={h}: " + " ".join(f"{v:.4f}" for v in r) + f" ms  median {r[2]:.4f} "
              f"({24 * n / r[2] / 1e6:.0f} GB/s algorithmic)  bit-equal={same[h]}", flush=True)
    del f, ox, oy
