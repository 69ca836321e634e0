"""
Solver-only roofline sweep (BASELINE configs[4]): NearToeplitzSolver.solve, in place, contiguous lines,
system size n in 32..4096 x batch in 2^10..2^20 (capped at 2^30 unknowns), Pade matrix [1,2,1/4,1,1/4,2,1].
Prints a table of ms and algorithmic GB/s (16 B per unknown) and writes JSON lines to gpurun_out/.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C

PADE = [1., 2., .25, 1., .25, 2., 1.]
out = open(os.path.join("gpurun_out", "solver_sweep.jsonl"), "w") if os.path.isdir("gpurun_out") else None
sizes = [32, 64, 128, 256, 512, 1024, 2048, 4096]
batches = [2 ** e for e in range(10, 21, 2)]
print("rows: system size n; columns: batch (number of systems); cell: ms / GB/s", flush=True)
print("n".rjust(6) + "".join(str(b).rjust(18) for b in batches))
for n in sizes:
    line = str(n).rjust(6)
    for b in batches:
        if n * b > 2 ** 30:
            line += "-".rjust(18)
            continue
        d = torch.rand((1, b, n), dtype=torch.float64, device="cuda")
        s = C.NearToeplitzSolver((1, b, n), PADE)
        for _ in range(3):
            s.solve(d)
        reps = 20 if n * b >= 2 ** 24 else 100
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            s.solve(d)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = 16 * n * b / ms / 1e6
        line += f"{ms:9.4f}/{gbs:7.0f} "
        if out:
            out.write(json.dumps({"n": n, "batch": b, "ms": ms, "GBps": gbs, "unknowns_per_s": n * b / ms * 1e3}) + "\n")
        del d, s
    print(line, flush=True)
# strided layouts at one size
for axis in (1, 2):
    shape = (512, 512, 512)
    d = torch.rand(shape, dtype=torch.float64, device="cuda")
    s = C.NearToeplitzSolver(shape, PADE, axis=axis)
    for _ in range(3):
        s.solve(d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        s.solve(d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"512^3 solve along axis {axis}: {ms:.4f} ms, {16 * 512 ** 3 / ms / 1e6:.0f} GB/s", flush=True)
    if out:
        out.write(json.dumps({"shape": shape, "axis": axis, "ms": ms, "GBps": 16 * 512 ** 3 / ms / 1e6}) + "\n")
