#!/usr/bin/env python
"""
bench.py -- headline benchmark of the compact-derivative path (BASELINE.json metric:
"grid points/sec per derivative (fp64), 1/2/4/8 B200, % of HBM roofline").

    python bench.py --gpus 1 --steps K --warmup W          512^3 fp64, derivative along x, y and z (configs[2])
    torchrun ... bench.py --gpus N ...                     1024^3 fp64 z-partitioned over N ranks (configs[3])
    python bench.py --impl reference ...                   the reference's CPU path (npts.c) on the host cores

A "step" = the three derivatives d/dx, d/dy, d/dz of one resident field: one launch of stream_kernel_xy (d/dx and
d/dy, f read from HBM once) and one of stream_kernel (d/dz); at N > 1 d/dz adds the halo / interface exchange
kernels.  --separate runs d/dx and d/dy as two launches (the round-1 step).
`value` = (grid points x 3 derivatives x steps) / time: grid points per second PER DERIVATIVE, whole job.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid points/sec per derivative (fp64)"
UNIT = "points/s"
# The contract is ONE JSON line on stdout.  Libraries underneath (NCCL prints "NCCL version ..." on some boxes) write
# to fd 1 as well, so fd 1 is pointed at stderr for the whole run and the result line goes to the real stdout.
RESULT = sys.stdout


def _reserve_stdout():
    global RESULT
    sys.stdout.flush()
    RESULT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


BYTES_PER_POINT = 16            # algorithmic: read f (8 B) + write f' (8 B), SURVEY.md section 8(d)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, 2 ms period; nvidia-smi query
    of the same fields as a fallback -- B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.sm_max = index, [], set(), None
        self._stop, self._t, self.n = threading.Event(), None, 0
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            # honour CUDA_VISIBLE_DEVICES: map the CUDA ordinal to the NVML handle through the PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self._h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hnd = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hnd).bus == bus:
                        self._h = hnd
            if self._h is None:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv = self._nvml
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for bit, name in ((nv.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"),
                          (nv.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                          (nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                          (nv.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap")):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits"], text=True, timeout=5)
        s = [v.strip() for v in out.strip().split(",")]
        self.sm.append(float(s[0]))
        self.sm_max = float(s[1])
        for v, name in zip(s[2:], ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if v.lower().startswith("active"):
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                (self._sample_nvml if self._nvml else self._sample_smi)()
                self.n += 1
            except Exception:
                pass
            self._stop.wait(0.002 if self._nvml else 0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": self.n,
                "source": "nvml" if self._nvml else "nvidia-smi"}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the reference's own npts.c (oracle/_ref) when it was built, else the C port
# ----------------------------------------------------------------------------------------------------
def cpu_reference_rate(seconds_budget=12.0, n=512, threads=None):
    """
    d/dx (RHS + tridiagonal solve) of a [planes, n, n] sample of the n^3 workload on the host cores.
    kind "reference": RHS by the C port (the reference computes it inside test_npts.c's main(), :86-97, not as a
    callable), solve by the UNMODIFIED lanl-implementation/npts.c (oracle/_ref), one independent solver call per
    thread over a y-z sub-batch (= the reference's npz x npy decomposition with npx = 1, which needs no
    communication).  kind "port": oracle/cfd_oracle.c for both.
    Returns (points_per_s, kind, cores, sample_text).
    """
    from oracle import cfd_oracle as O
    O.build()
    cores = threads or (os.cpu_count() or 1)
    kind = "reference" if O.have_ref() else "port"
    planes = 4 * max(cores, 16)
    rng = np.random.default_rng(0)
    f = rng.random((planes, n, n))
    h = 2 * np.pi / (n - 1)
    O.port().oracle_set_num_threads(cores)

    if kind == "reference":
        lib = O.ref()
        dp = ctypes.POINTER(ctypes.c_double)
        beta, gam = np.zeros(n), np.zeros(n)
        lib.precompute_beta_gam(0, n, n, 1, beta.ctypes.data_as(dp), gam.ctypes.data_as(dp))
        per = planes // cores
        slabs = [(t * per, (t + 1) * per if t < cores - 1 else planes) for t in range(cores)]
        u = np.zeros_like(f)
        phi, psi = np.zeros_like(f), np.zeros_like(f)

        def one_pass():
            r = O.rhs(f, 0, h)

            def work(lo, hi):
                lib.nonperiodic_tridiagonal_solver(0, n, n, hi - lo, beta.ctypes.data_as(dp), gam.ctypes.data_as(dp),
                                                   r[lo:hi].ctypes.data_as(dp), u[lo:hi].ctypes.data_as(dp),
                                                   phi[lo:hi].ctypes.data_as(dp), psi[lo:hi].ctypes.data_as(dp))
            ts = [threading.Thread(target=work, args=s) for s in slabs if s[1] > s[0]]
            [t.start() for t in ts]
            [t.join() for t in ts]
    else:
        def one_pass():
            O.derivative(f, 0, h)

    # npts.c prints timing lines (options.h PRINT_TIMINGS): silence fd 1 around the timed region
    sys.stdout.flush()
    saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        one_pass()                                   # warm-up
        t0, reps = time.perf_counter(), 0
        while True:
            one_pass()
            reps += 1
            dt = time.perf_counter() - t0
            if dt > seconds_budget or reps >= 2000:
                break
        ctypes.CDLL(None).fflush(None)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    rate = f.size * reps / dt
    sample = f"d/dx of a [{planes},{n},{n}] fp64 slab of the {n}^3 field, {reps} passes in {dt:.1f} s"
    return rate, kind, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 512 if args.gpus == 1 else 1024
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    if args.ref_seconds:
        per_step = float(args.ref_seconds)
    rates = []
    kind = cores = sample = None
    for i in range(args.warmup + args.steps):
        r, kind, cores, sample = cpu_reference_rate(seconds_budget=per_step, n=n)
        if i >= args.warmup:
            rates.append(r)
    value = float(np.mean(rates))
    pts = n ** 3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * 3 * pts / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT, flush=True)


def workload_config(n_gpus, comm="nvlink", overlap=True, fused_edge=False, zx=False):
    if n_gpus == 1:
        return {"workload": "512^3 fp64 field, derivative along x, y and z on 1 B200 (BASELINE configs[2])",
                "grid": [512, 512, 512], "derivatives_per_step": 3, "partition": "none",
                "step": "one fused d/dx + d/dy launch (f read from HBM once) + one d/dz launch",
                "l2": "inputs (1 GiB field) larger than the 126 MB L2; no flush needed"}
    return {"workload": f"1024^3 fp64 field z-partitioned over {n_gpus} B200, derivative along x, y and z "
                        "(d/dz: one-plane halo + interface exchange with the z-neighbours + coupled solve; BASELINE configs[3])",
            "grid": [1024, 1024, 1024], "derivatives_per_step": 3, "partition": f"z/{n_gpus}",
            "step": ("per rank two launches (= cfd_zpart_apply_xyz): fused d/dx + d/dy kernel, then the one-kernel partitioned "
                     "d/dz (per bundle: edge faces, self-validating words stored into the z-neighbours' memory over "
                     "NVLink, reduced system in registers, coupled solve)") if zx else
                    ("per rank three launches (cfd_zpart_apply_xyz, CFD_NO_ZX): fused d/dx + d/dy kernel whose first work items "
                     "are the edge faces of d/dz (pushed to the z-neighbours over NVLink), reduced solve, coupled d/dz kernel")
            if fused_edge else "per rank: one fused d/dx + d/dy launch on the slab + the partitioned d/dz",
            "ddz": ("one kernel, exchange inside it" if zx else f"fused (edge faces -> {comm} exchange -> coupled kernel)"),
            "overlap": ("exchange inside the d/dz launch, one bundle ahead of its use" if zx else
                        "exchange inside the x/y launch" if fused_edge else
                        "d/dz exchange started before d/dx, d/dy" if overlap else "none"),
            "l2": "inputs (slab >= 1 GiB) larger than the 126 MB L2; no flush needed"}



# ----------------------------------------------------------------------------------------------------
# in-run correctness checks (outside the timed region) and the extra driver-visible configurations
# ----------------------------------------------------------------------------------------------------
def check_derivatives(f, df, h, N, rank, world, nz_loc, dev, dist, nsample=256):
    """All three derivatives of the timed step against (a) the analytic derivative of the synthetic field, whole
    field, every rank, and (b) the CPU oracle on `nsample` sampled lines per direction -- z lines gathered across the
    ranks, so the partitioned d/dz (the only part that communicates) is checked end to end.  The oracle is the checker
    here, never the thing measured."""
    import torch
    from oracle import cfd_oracle as O
    t1 = torch.arange(N, dtype=torch.float64, device=dev) * h
    zz = t1[rank * nz_loc:(rank + 1) * nz_loc]
    sx, cx = torch.sin(t1)[None, None, :], torch.cos(t1)[None, None, :]
    sy, cy = torch.sin(t1)[None, :, None], torch.cos(t1)[None, :, None]
    sz, cz = torch.sin(zz)[:, None, None], torch.cos(zz)[:, None, None]
    errs = []
    for a, ex in enumerate((lambda: cx * cy * sz, lambda: -sx * sy * sz, lambda: sx * cy * cz)):
        e = ex()
        errs.append((df[a] - e).abs().max().item())
        del e
    g = torch.Generator().manual_seed(1234)                      # the same sample on every rank
    iz = torch.randint(0, nz_loc, (nsample,), generator=g)
    iy = torch.randint(0, N, (nsample,), generator=g)
    ix = torch.randint(0, N, (nsample,), generator=g)
    iy[:4] = torch.tensor([0, 0, N - 1, N - 1]); ix[:4] = torch.tensor([0, N - 1, 0, N - 1])   # corner lines too
    izd, iyd, ixd = iz.to(dev), iy.to(dev), ix.to(dev)
    rel = []
    # x lines and y lines live inside the slab
    for a, (lines_f, lines_d) in enumerate(((f[izd, iyd, :], df[0][izd, iyd, :]),
                                            (f[izd, :, ixd], df[1][izd, :, ixd]))):
        want = O.derivative(np.ascontiguousarray(lines_f.cpu().numpy()[None]), 0, h)[0]
        rel.append(float(np.abs(lines_d.cpu().numpy() - want).max() / np.abs(want).max()))
    # z lines: every rank contributes its part of the sampled columns
    part_f, part_d = f[:, iyd, ixd].contiguous(), df[2][:, iyd, ixd].contiguous()          # [nz_loc, nsample]
    if world > 1:
        all_f = [torch.empty_like(part_f) for _ in range(world)]
        all_d = [torch.empty_like(part_d) for _ in range(world)]
        dist.all_gather(all_f, part_f)
        dist.all_gather(all_d, part_d)
        part_f, part_d = torch.cat(all_f, 0), torch.cat(all_d, 0)
    zl_f = np.ascontiguousarray(part_f.cpu().numpy().T)                                     # [nsample, N]
    want = O.derivative(zl_f[None], 0, h)[0]
    rel.append(float(np.abs(part_d.cpu().numpy().T - want).max() / np.abs(want).max()))
    t = torch.tensor(errs + rel, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    v = t.tolist()
    out = {"ddx_max_abs_err_vs_analytic": v[0], "ddy_max_abs_err_vs_analytic": v[1], "ddz_max_abs_err_vs_analytic": v[2],
           "ddx_rel_linf_vs_oracle": v[3], "ddy_rel_linf_vs_oracle": v[4], "ddz_rel_linf_vs_oracle": v[5],
           "oracle_lines_per_direction": nsample, "ranks_checked": world}
    assert max(v[:3]) < 1e-6, f"analytic check failed: {out}"
    assert max(v[3:]) <= 1e-12, f"oracle parity failed: {out}"
    return out


def _time_calls(fn, reps=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def extra_configs(C, dev, peak):
    """BASELINE configs[1] (256^3 along x, y, z), four corners of configs[4] (solver-only sweep) and the 1024^3
    gradient on ONE GPU (the like-for-like point of the multi-GPU curve), each timed with CUDA events over 20
    back-to-back calls after 3 warm-ups.  256^3 (128 MiB in + 128 MiB out) is of the order of the 126 MB L2: the
    rotating set below (4 field pairs, 1 GiB) keeps every call's input out of L2."""
    import torch
    out = {}
    N = 256
    h = 2 * np.pi / (N - 1)
    t1 = torch.arange(N, dtype=torch.float64, device=dev) * h
    fs = [(torch.sin(t1 + 0.1 * i)[None, None, :] * torch.cos(t1)[None, :, None] * torch.sin(t1)[:, None, None]).contiguous()
          for i in range(4)]
    ds = [torch.empty_like(fs[0]) for _ in range(4)]
    per_axis = {}
    for a in range(3):
        op = C.CompactFiniteDifferenceSolver((N, N, N), h, a)
        k = [0]

        def call():
            op(fs[k[0] & 3], ds[k[0] & 3])
            k[0] += 1
        ms = _time_calls(call)
        gbs = 16 * N ** 3 / (ms * 1e-3) / 1e9
        per_axis["xyz"[a]] = {"ms": ms, "GBps": gbs, "frac_of_measured": gbs / peak, "frac_of_8TBps": gbs / 8000.0}
    g = C.CompactFiniteDifferenceSolver((N, N, N))
    k = [0]

    def grad():
        g.gradient(fs[k[0] & 3], (h, h, h), (ds[0], ds[1], ds[2]))
        k[0] += 1
    per_axis["gradient_ms"] = _time_calls(grad)            # cfd_apply_xyz: d/dz on the library's side stream
    os.environ["CFD_XYZ_SERIAL"] = "1"
    per_axis["gradient_one_stream_ms"] = _time_calls(grad)
    os.environ.pop("CFD_XYZ_SERIAL")
    out["256^3 (configs[1])"] = per_axis
    del fs, ds
    # the 512^3 step of the headline as gradient() issues it (x/y launch and d/dz launch on two streams)
    N = 512
    h = 2 * np.pi / (N - 1)
    f = torch.rand((N, N, N), dtype=torch.float64, device=dev)
    d3 = [torch.empty_like(f) for _ in range(3)]
    g = C.CompactFiniteDifferenceSolver((N, N, N))
    ms = _time_calls(lambda: g.gradient(f, (h, h, h), d3))
    out["512^3 gradient() with d/dz on the library's side stream (cfd_apply_xyz)"] = {
        "ms_per_step": ms, "points_per_s_per_derivative": 3 * N ** 3 / (ms * 1e-3),
        "note": "the timed step of `value` issues the two launches on ONE stream so that each has its own CUDA-event duration"}
    del f, d3, g
    sweep = []
    for n, batch in ((32, 1 << 10), (32, 1 << 20), (4096, 1 << 10), (4096, 1 << 18)):      # 2^30-unknown cap
        d = torch.rand((1, batch, n), dtype=torch.float64, device=dev)
        sol = C.NearToeplitzSolver(d.shape, [1., 2., 0.25, 1., 0.25, 2., 1.])
        ms = _time_calls(lambda: sol.solve(d))
        sweep.append({"n": n, "batch": batch, "ms": ms, "GBps": 16 * n * batch / (ms * 1e-3) / 1e9,
                      "in_L2": bool(8 * n * batch < 100e6)})
        del d, sol
    out["solver_sweep_corners (configs[4])"] = sweep
    N = 1024
    h = 2 * np.pi / (N - 1)
    t1 = torch.arange(N, dtype=torch.float64, device=dev) * h
    f = (torch.sin(t1)[None, None, :] * torch.cos(t1)[None, :, None] * torch.sin(t1)[:, None, None]).contiguous()
    df = [torch.empty_like(f) for _ in range(3)]
    g = C.CompactFiniteDifferenceSolver((N, N, N))
    ms = _time_calls(lambda: g.gradient(f, (h, h, h), df), reps=10)
    out["1024^3 gradient on 1 GPU (like-for-like point of the multi-GPU curve)"] = {
        "ms_per_step": ms, "points_per_s_per_derivative": 3 * N ** 3 / (ms * 1e-3)}
    del f, df
    torch.cuda.empty_cache()
    return out


def bind_to_gpu_numa_node(local):
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off, so that the pinned host buffers of the
    e2e leg are allocated (first touch) in memory local to the GPU's PCIe root.  Returns (node, previous affinity)."""
    import torch
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None, None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node, prev
    except Exception:
        pass
    return None, None


def copy_ceiling(f, out_host, f_host, fence, reduce_max, reps=2):
    """Pinned-copy ceiling of the e2e step: H2D of the field and D2H of three results, both directions at once, no
    kernel, ALL ranks copying at the same time -- every repetition starts behind a barrier (`fence`) and counts as the
    slowest rank's time (`reduce_max`); the best repetition is returned, in seconds per step.  (Round 2 found the
    barrier missing: ranks drifted apart between repetitions, the slow ranks' best repetition ran while the fast ones
    were already idle, and the 8-GPU ceiling read 203 ms where the copies of all ranks together take 299 ms --
    profiles/r2z_e2e_variants_n8.txt.)"""
    import torch
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    tmp = torch.empty_like(f)
    best = None
    for _ in range(reps + 1):
        fence()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_in):
            tmp.copy_(f_host, non_blocking=True)
        with torch.cuda.stream(s_out):
            for o in out_host:
                o.copy_(f, non_blocking=True)
        torch.cuda.synchronize()
        dt = reduce_max(time.perf_counter() - t0)
        best = dt if best is None else min(best, dt)
    return best

# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import compact_finite_differences_b200 as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N = 512 if world == 1 else 1024
    nz_loc = N // world
    h = 2 * np.pi / (N - 1)
    t1 = torch.arange(N, dtype=torch.float64, device=dev) * h
    zz = t1[rank * nz_loc:(rank + 1) * nz_loc]
    f = (torch.sin(t1)[None, None, :] * torch.cos(t1)[None, :, None] * torch.sin(zz)[:, None, None]).contiguous()
    df = [torch.empty_like(f) for _ in range(3)]
    # d/dx and d/dy never cross a z-slab: one fused launch (cfd_apply_xy) on every rank; d/dz is the partitioned one
    xy = C.CompactFiniteDifferenceSolver((nz_loc, N, N))
    ddz = C.ZPartitionedDerivative((nz_loc, N, N), h, 2, mode="fused", comm=args.comm) if world > 1 else \
        C.CompactFiniteDifferenceSolver((nz_loc, N, N), h, 2)
    pts_local = f.numel()
    # N > 1, default: two launches per rank and step -- the fused d/dx + d/dy kernel and the ONE-kernel partitioned d/dz
    # (stream_kernel_zx: edge faces, NVLink exchange, reduced system and coupled solve per bundle) -- issued exactly
    # as cfd_zpart_apply_xyz issues them (cfd_apply_xy, then cfd_zpart_apply), so that each has its own CUDA events.
    # --three-launch = the first half of round 2 (edge items inside the x/y kernel, reduce, coupled d/dz; one C call);
    # --chain = the round-1 step (exchange chain on a side stream beside the x/y launch, 5 xy warps on thin slabs).
    zx = world > 1 and args.comm == "nvlink" and not args.chain and not args.three_launch
    fused_edge = world > 1 and args.comm == "nvlink" and args.three_launch and not args.separate
    if fused_edge:
        os.environ["CFD_NO_ZX"] = "1"
    xy_warps = 5 if (world > 1 and args.chain and not args.no_overlap and nz_loc <= 128) else None

    def gradient(src, events=None):
        if fused_edge:
            if events is not None:
                events[0][0].record()
            ddz.gradient(src, h, h, df)
            if events is not None:
                events[1][1].record()
            return
        if world > 1 and args.chain and not args.no_overlap:
            ddz.begin(src)           # halo + interface exchange of d/dz overlaps the d/dx + d/dy kernel
        if events is not None:
            events[0][0].record()
        if args.separate:
            xy.dfdx(src, h, df[0])
            xy.dfdy(src, h, df[1])
        else:
            xy.dfdxy(src, h, h, df[0], df[1], warps=xy_warps)
        if events is not None:
            events[0][1].record()
            events[1][0].record()
        ddz(src, df[2])
        if events is not None:
            events[1][1].record()

    def step(events=None):
        gradient(f, events)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    fence()
    if zx or fused_edge:
        assert ddz.comm == "nvlink", "CUDA IPC unavailable: the step fell back; rerun with --chain --comm pairwise"

    # in-run correctness of the timed path, all three derivatives, every rank (cheap, outside the timed region)
    check = check_derivatives(f, df, h, N, rank, world, nz_loc, dev, dist)

    launches0 = C.lib().cfd_launch_count()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2)]
          for _ in range(args.steps)]
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        fence()
        t_beg.record()
        for s in range(args.steps):
            step(ev[s])
        t_end.record()
        fence()
    launches = C.lib().cfd_launch_count() - launches0
    ms = t_beg.elapsed_time(t_end)
    if fused_edge:
        # the three launches are issued by ONE C call: per-launch durations come from a second, untimed pass through
        # the same entry points one at a time (begin = edge + reduce, cfd_apply_xy, coupled d/dz) -- they explain the
        # step, they are not part of `value`
        span = float(np.mean([ev[s][0][0].elapsed_time(ev[s][1][1]) for s in range(args.steps)]))
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = np.zeros(3)
        for _ in range(5):
            fence()
            e[0].record(); ddz.begin(f); e[1].record(); xy.dfdxy(f, h, h, df[0], df[1]); e[2].record()
            ddz(f, df[2]); e[3].record()
            torch.cuda.synchronize()
            acc += [e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])]
        acc /= 5
        per_launch = [span - float(acc[2]), float(acc[2])]      # [fused x/y + edge items (+ reduce), coupled d/dz]
        pieces = {"edge_plus_reduce_standalone_ms": float(acc[0]), "xy_standalone_ms": float(acc[1]),
                  "z_coupled_standalone_ms": float(acc[2]), "step_span_ms": span}
    else:
        per_launch = [float(np.mean([ev[s][a][0].elapsed_time(ev[s][a][1]) for s in range(args.steps)])) for a in range(2)]
        pieces = None
    if world > 1:
        t = torch.tensor([ms] + per_launch, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, per_launch = t[0].item(), t[1:].tolist()
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    value = (N ** 3) * 3 * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers in pinned memory through the public host API, copies inside the timed region
    e2e_steps = max(1, min(args.steps, 5))
    numa_node, prev_aff = bind_to_gpu_numa_node(local) if world > 1 else (None, None)   # NUMA-local pinned buffers
    f_host = torch.empty(f.shape, dtype=torch.float64, pin_memory=True)
    f_host.copy_(f)
    out_host = [torch.empty(f.shape, dtype=torch.float64, pin_memory=True) for _ in range(3)]
    for o in out_host:
        o.zero_()                      # first touch here, on the bound cores
    # "pipelined" = HostGradient: H2D in z-slabs, fused d/dx + d/dy per slab, D2H overlapping the remaining H2D, then
    # d/dz + D2H (at N > 1 each rank pipelines its own slab and d/dz is the partitioned operator).  "sequential" =
    # H2D, gradient, D2H one after the other.  Measured (round 2, profiles/r2e_bench_n2*.json): pipelined reaches
    # 98 % (N = 1) and 99.7 % (N = 2) of the pinned-copy ceiling, sequential 86 % at N = 2 -> pipelined at every N.
    e2e_mode = args.e2e if args.e2e != "auto" else "pipelined"
    if e2e_mode == "pipelined":
        hg = C.HostGradient((nz_loc, N, N), (h, h, h), slabs=8, ddz=ddz if world > 1 else None)

        def e2e_step():
            hg(f_host, out_host)
    else:
        f_in = torch.empty_like(f)

        def e2e_step():
            f_in.copy_(f_host, non_blocking=True)
            gradient(f_in)
            for a in range(3):
                out_host[a].copy_(df[a], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    e2e_step()
    fence()
    # every step is timed on its own (host clock, fence on both sides) and the MEDIAN step is reported: the leg is a
    # few hundred milliseconds of host-driven copies, and one scheduling hiccup on a busy host would otherwise decide it
    e2e_times = []
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        e2e_step()
        fence()
        e2e_times.append(time.perf_counter() - t0)
    e2e_s = float(np.median(e2e_times)) * e2e_steps
    step()                     # refresh df with the device-resident result for the comparison below
    torch.cuda.synchronize()
    e2e_ok = all(float((out_host[a] - df[a].cpu()).abs().max()) == 0.0 for a in range(3))
    # pinned-copy ceiling of the same step: all ranks copy at once (they share the host's memory system), no kernels
    fence()
    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    ceil_s = copy_ceiling(f, out_host, f_host, fence, reduce_max)
    e2e_s = reduce_max(e2e_s)
    e2e_value = (N ** 3) * 3 * e2e_steps / e2e_s
    ceil_value = (N ** 3) * 3 / ceil_s
    if prev_aff:
        os.sched_setaffinity(0, prev_aff)
    del f_host, out_host

    # ---- roofline, live CUDA-event durations.  The fused d/dx + d/dy launch is the longest one of the step; the bytes
    # it cannot avoid are f once + two results = 24 B/point (the two derivatives as separate passes are 2 x 16 B/point
    # -- that credit is reported next to it, not used for `frac`).  At N > 1 the launch also carries the edge-face
    # items of d/dz (67 of nz_loc planes re-read): they are overhead, not credited.
    peak, peak_src = measured_peak()
    xy_bytes = (2 * BYTES_PER_POINT if args.separate else 24) * pts_local
    z_bytes = BYTES_PER_POINT * pts_local
    gbps = [xy_bytes / (per_launch[0] * 1e-3) / 1e9, z_bytes / (per_launch[1] * 1e-3) / 1e9]
    dom = int(np.argmax(per_launch))
    names = ["stream_kernel d/dx, d/dy (two launches)" if args.separate else
             ("stream_kernel_xy (d/dx + d/dy + edge-face items of d/dz, one launch) + reduced_planes" if fused_edge else
              "stream_kernel_xy (d/dx + d/dy, one launch)"),
             "stream_kernel_zx (partitioned d/dz in one launch: edge faces, NVLink exchange, reduced system, coupled solve)"
             if zx else ("stream_kernel d/dz (coupled)" if world > 1 else "stream_kernel d/dz")]
    # dram bytes (read + write) per launch from `ncu --set full`, keyed by kernel and slab shape; null when this shape
    # has not been profiled (profiles/traffic.json names the capture each figure comes from)
    shape_key = f"{nz_loc}x{N}x{N}"
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    tkeys = [("xy_edge@" if fused_edge else "xy@") + shape_key,
             ("zx@" if zx else "z_coupled@" if world > 1 else "z@") + shape_key]
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": gbps[dom], "peak": peak,
                "unit": "GB/s", "frac": gbps[dom] / peak, "traffic": traffic.get(tkeys[dom]), "peak_source": peak_src,
                "frac_of_8TBps_nominal": gbps[dom] / 8000.0,
                "launches": {
                    "xy": {"kernel": names[0], "ms": per_launch[0], "algorithmic_bytes": xy_bytes, "achieved": gbps[0],
                           "frac": gbps[0] / peak, "traffic": traffic.get(tkeys[0])},
                    "z": {"kernel": names[1], "ms": per_launch[1], "algorithmic_bytes": z_bytes, "achieved": gbps[1],
                          "frac": gbps[1] / peak, "traffic": traffic.get(tkeys[1])}},
                "algorithmic_bytes_per_launch": [xy_bytes, z_bytes][dom],
                "step_GBps_at_16B_per_point_per_derivative":
                    3 * BYTES_PER_POINT * pts_local * args.steps / (ms * 1e-3) / 1e9}
    if pieces:
        roofline["pieces"] = pieces
    # Streaming yardsticks, live, on this rank's arrays: a plain grid-stride kernel with the launches' read : write
    # mixes (cfd_debug_stream: copy = a single derivative's mix; one read + two writes = the fused x/y launch's).
    # DRAM streams the 1 : 2 mix several percent below the 1 : 1 copy figure `peak` holds; `frac_of_yardstick` says
    # how far each launch is from a kernel that does nothing but move its bytes.  Purely additive: never fails the line.
    try:
        L = C.lib()
        sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

        def stream_ms(third):
            def call():
                rc = L.cfd_debug_stream(f.data_ptr(), df[0].data_ptr(), df[1].data_ptr() if third else None,
                                        f.numel(), sp)
                if rc != 0:
                    raise RuntimeError(f"cfd_debug_stream: {rc}")
            return _time_calls(call, reps=10, warm=2)
        y_copy = 16 * pts_local / (stream_ms(False) * 1e-3) / 1e9
        y_r1w2 = 24 * pts_local / (stream_ms(True) * 1e-3) / 1e9
        roofline["yardstick"] = {"what": "plain streaming kernels timed live on the same arrays (cfd_debug_stream)",
                                 "copy_1r1w_GBps": y_copy, "one_read_two_writes_GBps": y_r1w2}
        roofline["launches"]["xy"]["frac_of_yardstick"] = gbps[0] / (y_copy if args.separate else y_r1w2)
        roofline["launches"]["z"]["frac_of_yardstick"] = gbps[1] / y_copy
    except Exception as e:         # noqa: BLE001  (df[0], df[1] now hold the yardstick's copies; nothing reads them after this)
        roofline["yardstick"] = {"error": repr(e)}

    extra = None
    if world == 1 and not args.no_extra:
        del df, f
        torch.cuda.empty_cache()
        extra = extra_configs(C, dev, peak)

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu:
            r, kind, cores, sample = cpu_reference_rate(seconds_budget=12.0, n=N)
            cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, args.comm, not args.no_overlap, fused_edge, zx),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pts_local * 8 * world,
                    "d2h_bytes_per_step": 3 * pts_local * 8 * world, "steps": e2e_steps, "verified": e2e_ok,
                    "mode": e2e_mode, "timing": "median of the per-step wall times",
                    "step_ms": [round(1e3 * t, 2) for t in e2e_times],
                    "pinned_copy_ceiling": ceil_value, "frac_of_copy_ceiling": e2e_value / ceil_value,
                    "numa_node_of_rank0_buffers": numa_node,
                    "ceiling": "H2D of f + D2H of three results, both directions at once, all ranks at the same time (barrier before every repetition, slowest rank counts), no kernels"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "check": check,
        }
        if extra:
            line["extra"] = extra
        print(json.dumps(line), file=RESULT, flush=True)
    if world > 1:
        dist.barrier()
        ddz.close()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ref-seconds", type=float, default=0.0,
                    help="--impl reference: CPU seconds per step (default: 60 s spread over the steps, 2 .. 20 s each)")
    ap.add_argument("--comm", default="nvlink", choices=["nvlink", "pairwise", "allgather"],
                    help="exchange of the partitioned d/dz: NVLink peer-memory stores from our kernels, NCCL send/recv per "
                         "z-neighbour, or NCCL all-gather")
    ap.add_argument("--no-overlap", action="store_true", help="do not start the d/dz exchange before d/dx, d/dy")
    ap.add_argument("--e2e", default="auto", choices=["auto", "pipelined", "sequential"],
                    help="host-buffer leg: HostGradient slab pipeline, or H2D -> gradient -> D2H in sequence")
    ap.add_argument("--separate", action="store_true", help="d/dx and d/dy as two launches instead of cfd_apply_xy")
    ap.add_argument("--chain", action="store_true",
                    help="N > 1: the round-1 step (exchange chain on a side stream beside the x/y launch)")
    ap.add_argument("--three-launch", action="store_true",
                    help="N > 1: edge items inside the x/y kernel + reduce + coupled d/dz (first half of round 2) instead of "
                         "the x/y launch + one-kernel d/dz")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the extra configurations (256^3, solver sweep, 1024^3)")
    args = ap.parse_args()
    _reserve_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
