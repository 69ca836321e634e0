"""
CPU baselines of SURVEY.md section 8(d), timed on the host cores of the box this runs on:
 (i)   the reference's own time_npts / test_npts binaries (oracle/_ref, unmodified npts.c), one core, N = 64, 256;
 (ii)  all-core variant: one independent npts solver call per host thread over a y-z sub-batch (+ port RHS);
 (iii) LAPACK dgtsv (stand-in for perf-test/CPU/intel-MKL/main.cpp:116) on [nx, ny*nz] right-hand sides.
Writes gpurun_out/cpu_baselines.json when that directory exists.
"""
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cfd_oracle as O  # noqa: E402
import bench  # noqa: E402

out = {"nproc": os.cpu_count()}
O.build()
if O.have_ref():
    for n in (64, 256):
        r = subprocess.check_output([O.REF_TIME_BIN, str(n), str(n), str(n), "1", "1", "1"], text=True)
        times = [float(x) for x in re.findall(r"^Time: ([0-9.]+)", r, flags=re.M)]
        med = float(np.median(times[2:]))
        out[f"time_npts_{n}"] = {"median_s": med, "pts_per_s": n ** 3 / med, "cores": 1, "reps": len(times),
                                 "what": "reference time_npts.run N N N 1 1 1 (solve only, RHS = 1)"}
        t0 = time.perf_counter()
        line = O.ref_known_answer(n)
        out[f"test_npts_{n}"] = {"wall_s": time.perf_counter() - t0, "line": line,
                                 "what": "reference test_npts.run (RHS + solve of d/dx sin, incl. process start-up)"}
    for n in (256, 512):
        rate, kind, cores, sample = bench.cpu_reference_rate(seconds_budget=8.0, n=n)
        out[f"allcore_{n}"] = {"pts_per_s": rate, "kind": kind, "cores": cores, "sample": sample}
try:
    from scipy.linalg import lapack
    for n in (64, 256):
        rng = np.random.default_rng(0)
        B = np.asfortranarray(rng.random((n, n * n)))
        dl = np.full(n - 1, 0.25); d = np.ones(n); du = np.full(n - 1, 0.25)
        du[0] = 2.0; dl[-1] = 2.0
        lapack.dgtsv(dl.copy(), d.copy(), du.copy(), B.copy())
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            lapack.dgtsv(dl.copy(), d.copy(), du.copy(), B.copy())
        dt = (time.perf_counter() - t0) / reps
        out[f"dgtsv_{n}"] = {"s": dt, "pts_per_s": n ** 3 / dt, "what": "scipy.linalg.lapack.dgtsv, Pade matrix, "
                             "[n, n*n] Fortran RHS (includes the RHS copy)"}
except Exception as e:  # pragma: no cover
    out["dgtsv_error"] = str(e)
print(json.dumps(out, indent=1))
if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cpu_baselines.json"), "w"), indent=1)
