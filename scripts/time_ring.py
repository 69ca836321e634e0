"""Experiment: single-direction launches through the ring-staged kernel (CFD_RING_STAGING=1) against the default
stream_kernel, for several shapes and warps/SM.  usage: time_ring.py nz ny nx [nz ny nx ...]"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C


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


args = [int(a) for a in sys.argv[1:]] or [512, 512, 512]
shapes = [tuple(args[i:i + 3]) for i in range(0, len(args), 3)]
for shape in shapes:
    f = torch.rand(shape, dtype=torch.float64, device="cuda")
    out = torch.empty_like(f)
    pts = f.numel()
    for axis in range(3):
        op = C.CompactFiniteDifferenceSolver(shape, 0.1, axis)
        os.environ.pop("CFD_RING_STAGING", None)
        C.lib().cfd_set_launch(0, 0, 0)
        ref = op(f).clone()
        t0 = timeit(lambda: op(f, out))
        line = f"{shape} d/d{'xyz'[axis]}: default {t0:.4f} ms ({16 * pts / t0 / 1e6:.0f} GB/s) | ring-staged"
        os.environ["CFD_RING_STAGING"] = "1"
        for w in (4, 5, 6, 7):
            C.lib().cfd_set_launch(w, 0, 0)
            out.zero_()
            op(f, out)
            ok = torch.equal(out, ref) or float((out - ref).abs().max() / ref.abs().max()) < 1e-14
            t = timeit(lambda: op(f, out))
            line += f" w{w}: {t:.4f}{'' if ok else ' WRONG'}"
        print(line, flush=True)
    os.environ.pop("CFD_RING_STAGING", None)
    C.lib().cfd_set_launch(0, 0, 0)
    del f, out
