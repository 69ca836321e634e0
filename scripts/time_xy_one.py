import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C
shape = tuple(int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 30
f = torch.rand(shape, dtype=torch.float64, device="cuda")
ox, oy = torch.empty_like(f), torch.empty_like(f)
s = C.CompactFiniteDifferenceSolver(shape)
for _ in range(3):
    s.dfdxy(f, 0.1, 0.1, ox, oy)
torch.cuda.synchronize()
ts = []
for rnd in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        s.dfdxy(f, 0.1, 0.1, ox, oy)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / reps)
print(shape, "skew=" + os.environ.get("CFD_XY_SKEW", "default"), " ".join(f"{t:.4f}" for t in sorted(ts)), "checksum", float(ox.double().sum() + oy.double().sum()))
