"""ms of the fused d/dx + d/dy launch on slab shapes under the values of one experiment switch that the library reads
at every launch (CFD_XY_HINTS, CFD_XY_SUB, CFD_XY_TAU, ...), interleaved rounds, bit-equality against the default.
usage: time_xy_env.py VAR v1,v2,... [nz ny nx] ..."""
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


VAR, VALUES = sys.argv[1], sys.argv[2].split(",")
a = [int(v) for v in sys.argv[3:]]
shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)] or [(128, 1024, 1024), (256, 1024, 1024), (512, 512, 512)]
for shape in shapes:
    f = torch.rand(shape, dtype=torch.float64, device="cuda")
    ox, oy = torch.empty_like(f), torch.empty_like(f)
    s = C.CompactFiniteDifferenceSolver(shape)
    os.environ.pop(VAR, None)
    rx, ry = [o.clone() for o in s.dfdxy(f, 0.1, 0.1)]
    res = {h: [] for h in VALUES}
    same = {}
    for rnd in range(5):
        for h in res:
            os.environ[VAR] = str(h)
            res[h].append(timeit(lambda: s.dfdxy(f, 0.1, 0.1, ox, oy)))
            torch.cuda.synchronize()
            same[h] = torch.equal(ox, rx) and torch.equal(oy, ry)
    os.environ.pop(VAR, None)
    n = f.numel()
    for h, r in res.items():
        r = sorted(r)
        print(f"{shape} {VAR}={h}: " + " ".join(f"{v:.4f}" for v in r) + f" ms  median {r[2]:.4f} "
              f"({24 * n / r[2] / 1e6:.0f} GB/s algorithmic)  bit-equal={same[h]}", flush=True)
    del f, ox, oy
