"""Per-axis ms of N^3 derivative launches for warps/SM x CTA count (CFD_CTAS experiment switch of stream_kernel):
how much of the small-field time is the tail of the dynamic draw.  usage: sweep_256.py [N]"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fs = [torch.rand((N, N, N), dtype=torch.float64, device="cuda") for _ in range(4)]      # rotate: inputs never in L2
ds = [torch.empty_like(fs[0]) for _ in range(4)]
ops = [C.CompactFiniteDifferenceSolver((N, N, N), 0.1, a) for a in range(3)]


def timeit(a, reps=40):
    for i in range(4):
        ops[a](fs[i & 3], ds[i & 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops[a](fs[i & 3], ds[i & 3])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"N = {N}: 16 B/pt -> {16 * N ** 3 / 1e6:.0f} MB per launch; 70 % of 8 TB/s = {16 * N ** 3 / 5.6e9:.4f} ms")
for w in (4, 5, 3):
    for ctas in ("", "144", "140", "137", "132", "128", "120", "114", "103"):
        C.lib().cfd_set_launch(w, 0, 0)
        if ctas:
            os.environ["CFD_CTAS"] = ctas
        else:
            os.environ.pop("CFD_CTAS", None)
        r = [timeit(a) for a in range(3)]
        print(f"warps={w} ctas={ctas or '148':>4}  x={r[0]:.4f} y={r[1]:.4f} z={r[2]:.4f} ms   "
              f"({16 * N ** 3 / min(r) / 1e6:.0f} GB/s best)", flush=True)
os.environ.pop("CFD_CTAS", None)
C.lib().cfd_set_launch(0, 0, 0)
