"""Launch-shape sweep: ms per 512^3 (or N^3) derivative launch for warps/CTA x CTAs/SM x ring slots."""
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
df = torch.empty_like(f)
ops = [C.CompactFiniteDifferenceSolver((N, N, N), h, a) for a in range(3)]


def timeit(a, reps=20):
    for _ in range(3):
        ops[a](f, df)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops[a](f, df)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cfgs = []
for ns in (3, 4, 5):
    for ctas in (1, 2):
        for w in (2, 3, 4, 5, 6, 7, 8):
            cfgs.append((w, ctas, ns))
print("N =", N)
for (w, ctas, ns) in cfgs:
    C.lib().cfd_set_launch(w, ctas, ns)
    try:
        r = [timeit(a) for a in range(3)]
    except Exception as e:
        print(w, ctas, ns, "ERR", e)
        continue
    print(f"warps={w} ctas={ctas} ns={ns}  x={r[0]:.4f} y={r[1]:.4f} z={r[2]:.4f}")
