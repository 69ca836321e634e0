"""The fused x/y launch for the library's choice of warps per SM (0) and 5 / 6 / 7 forced, on cubes and slabs of several
sizes (rotating over input / output sets so that nothing is L2-resident), interleaved rounds.
usage: sweep_xy_warps.py [N | nz,ny,nx ...]"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C


def timeit(fn, reps):
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


args = sys.argv[1:] or ["128", "160", "192", "224", "256", "288", "320", "384", "512", "64,512,512", "32,512,512",
                        "128,256,256", "512,128,128", "100,320,200"]
for a in args:
    shape = tuple(int(v) for v in a.split(",")) if "," in a else (int(a),) * 3
    n = shape[0] * shape[1] * shape[2]
    nset = max(1, min(8, (1 << 30) // (n * 8)))
    fs = [torch.rand(shape, dtype=torch.float64, device="cuda") for _ in range(nset)]
    os_ = [[torch.empty_like(fs[0]) for _ in range(2)] for _ in range(nset)]
    s = C.CompactFiniteDifferenceSolver(shape)
    k = [0]
    res = {w: [] for w in (0, 5, 6, 7)}
    for rnd in range(3):
        for w in res:
            def run():
                i = k[0] % nset
                k[0] += 1
                s.dfdxy(fs[i], 0.1, 0.1, os_[i][0], os_[i][1], warps=w)
            res[w].append(timeit(run, 40 if n <= 256 ** 3 else 15))
    best = min(res, key=lambda w: min(res[w]) if w else 1e9)
    print(f"{str(shape):18s} ({nset} sets): " + "   ".join(
        (f"{w} warps " if w else "library ") + " ".join(f"{t:.4f}" for t in ts) + f" (min {min(ts):.4f})"
        for w, ts in res.items()) + f"   library / best forced = {min(res[0]) / min(res[best]):.3f}", flush=True)
    del fs, os_
