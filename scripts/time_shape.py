"""ms per derivative launch along x, y, z for arbitrary (nz, ny, nx); optional launch knobs after the shape."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C

if os.environ.get("CFD_KSEG"):
    C.lib().cfd_set_segments(int(os.environ["CFD_KSEG"]))
shapes = []
args = [int(a) for a in sys.argv[1:]]
while len(args) >= 3:
    shapes.append(tuple(args[:3]))
    args = args[3:]
for shape in shapes:
    f = torch.rand(shape, dtype=torch.float64, device="cuda")
    df = torch.empty_like(f)
    res = []
    for a in range(3):
        if shape[2 - a] < 4:
            res.append(float("nan"))
            continue
        op = C.CompactFiniteDifferenceSolver(shape, 0.1, a)
        for _ in range(3):
            op(f, df)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            op(f, df)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20)
    pts = shape[0] * shape[1] * shape[2]
    print(shape, " ".join(f"{'xyz'[a]}={res[a]:.4f}ms({16 * pts / res[a] / 1e6:.0f}GB/s)" for a in range(3)), flush=True)
    del f, df
