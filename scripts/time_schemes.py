"""ms and algorithmic GB/s (16 B / point) of the general one-pass kernel at N^3: the 6th-order first derivative
(two chunks of look-ahead), the 4th-order second derivative, and the solver-only call with alpha = 1/3 one-pass vs the
exact two-pass LU it replaced (CFD_NO_LA2=1).  usage: time_schemes.py [N]"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
shape = (N, N, N)
f = torch.rand(shape, dtype=torch.float64, device="cuda")
df = torch.empty_like(f)
pts = f.numel()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"N = {N}")
if os.environ.get("SCHEME_SWEEP"):           # launch shape of the general kernel: (warps per SM, ring slots)
    co = (1., 2., 1. / 3, 1., 1. / 3, 2., 1.)
    for w, ns in ((0, 0), (7, 0), (6, 0), (5, 0), (4, 0)):
        C.lib().cfd_set_launch(w, 0, ns)
        line = f"general kernel warps={w or 'default'} slots={ns or 'default'}:"
        for scheme in ("compact6", "pade4-d2"):
            res = []
            for a in range(3):
                op = C.CompactFiniteDifferenceSolver(shape, 0.1, a, scheme=scheme)
                res.append(timeit(lambda: op(f, df)))
            line += f"  {scheme} " + "/".join(f"{r:.4f}" for r in res)
        res = []
        for a in range(3):
            s = C.NearToeplitzSolver(shape, co, axis=a)
            res.append(timeit(lambda: s.solve(f)))
        line += "  solve(1/3) " + "/".join(f"{r:.4f}" for r in res)
        print(line, flush=True)
    C.lib().cfd_set_launch(0, 0, 0)
for scheme in ("pade4", "compact6", "pade4-d2"):
    res = []
    for a in range(3):
        op = C.CompactFiniteDifferenceSolver(shape, 0.1, a, scheme=scheme)
        res.append(timeit(lambda: op(f, df)))
    print(f"derivative {scheme:9s}: " + "  ".join(f"{'xyz'[a]} {res[a]:.4f} ms ({16 * pts / res[a] / 1e6:.0f} GB/s)" for a in range(3)), flush=True)
for co, name in (((1., 2., .25, 1., .25, 2., 1.), "alpha 1/4 (Pade, LA 1, register kernel)"),
                 ((1., 2., 1. / 3, 1., 1. / 3, 2., 1.), "alpha 1/3 (LA 2, general kernel)"),
                 ((1., 2., .3, 1., .3, 2., 1.), "alpha 0.3 (LA 2)")):
    for env in ("", "1"):
        if env and "1/4" in name:
            continue
        if env:
            os.environ["CFD_NO_LA2"] = "1"
        else:
            os.environ.pop("CFD_NO_LA2", None)
        res = []
        for a in range(3):
            s = C.NearToeplitzSolver(shape, co, axis=a)
            res.append(timeit(lambda: s.solve(f)))
            kind = "two-pass" if s.two_pass else f"one-pass LA {C.lib().nt_lookahead(s._handle)}"
        print(f"solve {name:40s} [{kind}]: " + "  ".join(f"{'xyz'[a]} {res[a]:.4f} ms ({16 * pts / res[a] / 1e6:.0f} GB/s)" for a in range(3)), flush=True)
os.environ.pop("CFD_NO_LA2", None)
