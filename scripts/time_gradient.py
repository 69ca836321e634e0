"""ms per gradient (d/dx, d/dy, d/dz) of an N^3 field: one stream (x/y launch, then z launch) against cfd_apply_xyz
(z on a side stream: the second kernel's CTAs take over as the first one's retire).  usage: time_gradient.py N [N ...]"""
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


for N in [int(a) for a in sys.argv[1:]] or [512]:
    nset = 4 if N <= 256 else 1                         # rotate small fields: nothing L2-resident
    fs = [torch.rand((N, N, N), dtype=torch.float64, device="cuda") for _ in range(nset)]
    outs = [[torch.empty_like(fs[0]) for _ in range(3)] for _ in range(nset)]
    s = C.CompactFiniteDifferenceSolver((N, N, N))
    h = (0.1, 0.1, 0.1)
    k = [0]

    def grad():
        i = k[0] % nset
        k[0] += 1
        s.gradient(fs[i], h, outs[i])
    ref = [o.clone() for o in s.gradient(fs[0], h)]
    cfgs = (("one stream", {"CFD_XYZ_SERIAL": "1"}), ("two streams, x/y first", {"CFD_XYZ_XY_FIRST": "1"}),
            ("two streams, z first (library default)", {}))
    res = {name: [] for name, _ in cfgs}
    same = {}
    for rnd in range(4):                                # interleaved rounds: clock / power drift hits every variant alike
        for name, env in cfgs:
            for kk in ("CFD_XYZ_SERIAL", "CFD_XYZ_XY_FIRST", "CFD_XYZ_TWO_STREAMS"):
                os.environ.pop(kk, None)
            os.environ.update(env)
            res[name].append(timeit(grad, reps=20))
            got = s.gradient(fs[0], h)
            torch.cuda.synchronize()
            same[name] = all(torch.equal(a, b) for a, b in zip(ref, got))
    for name, _ in cfgs:
        r = res[name]
        print(f"N = {N:4d}  {name:40s} " + " ".join(f"{v:.4f}" for v in r) + f" ms  (best {min(r):.4f}: "
              f"{3 * N ** 3 / min(r) / 1e6:.0f} Mpts/s per derivative)  bit-equal={same[name]}", flush=True)
    for kk in ("CFD_XYZ_SERIAL", "CFD_XYZ_XY_FIRST", "CFD_XYZ_TWO_STREAMS"):
        os.environ.pop(kk, None)
    del fs, outs
