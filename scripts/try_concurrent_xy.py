"""Experiment: d/dx and d/dy of the same field launched concurrently on two streams so that the second reader of
each z-plane hits L2.  Prints ms for x+y sequential vs concurrent, for several warps/CTA settings."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
h = 2 * np.pi / (N - 1)
t = torch.arange(N, dtype=torch.float64, device="cuda") * h
f = (torch.sin(t)[None, None, :] * torch.cos(t)[None, :, None] * torch.sin(t)[:, None, None]).contiguous()
d = [torch.empty_like(f) for _ in range(3)]
ops = [C.CompactFiniteDifferenceSolver((N, N, N), h, a) for a in range(3)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def seq():
    ops[0](f, d[0]); ops[1](f, d[1])


def conc():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        ops[0](f, d[0])
    with torch.cuda.stream(s2):
        ops[1](f, d[1])
    cur.wait_stream(s1); cur.wait_stream(s2)


def conc3():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        ops[0](f, d[0])
    with torch.cuda.stream(s2):
        ops[1](f, d[1])
    ops[2](f, d[2])
    cur.wait_stream(s1); cur.wait_stream(s2)


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


for w in (4, 3, 2, 1):
    C.lib().cfd_set_launch(w, 0, 0)
    print(f"warps/CTA={w}: x+y sequential {timeit(seq):.4f} ms, concurrent {timeit(conc):.4f} ms, "
          f"x|y|z all concurrent {timeit(conc3):.4f} ms", flush=True)
ref = [o.clone() for o in d]
C.lib().cfd_set_launch(0, 0, 0)
seq(); ops[2](f, d[2]); torch.cuda.synchronize()
print("max diff vs default:", [float((a - b).abs().max()) for a, b in zip(ref, d)])
