"""torchrun --nproc-per-node P scripts/e2e_variants_mp.py
The e2e step of the z-partitioned 1024^3 gradient (pinned host slab -> H2D -> three derivatives -> D2H) on P GPUs in
several schedules, with per-rank milestones (ms after the step's start, max over ranks): when the host->device copies,
the x/y results' device->host copies and the d/dz kernel finish.  Shows where the pipeline loses against the bare
copies ("ceiling": the same copies with no kernels and no dependencies) on boxes whose GPUs share PCIe uplinks.
    ceiling       H2D and the three D2H independent of each other
    pipelined     HostGradient as bench.py runs it (thin first slabs)
    uniform       HostGradient, 8 uniform slabs
    slabs4/16     coarser / finer uniform slabs
    sequential    H2D, gradient, D2H one after the other
    h2d-first     all H2D slabs issued before any D2H may start (x/y kernels overlap the H2D, D2H afterwards)
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nz = N // world
h = 2 * np.pi / (N - 1)
shape = (nz, N, N)
ddz = C.ZPartitionedDerivative(shape, h, 2, mode="fused", comm="nvlink")
xy = C.CompactFiniteDifferenceSolver(shape)
f = torch.rand(shape, dtype=torch.float64, device=dev)
d = [torch.empty_like(f) for _ in range(3)]
fh = torch.empty(shape, dtype=torch.float64, pin_memory=True)
fh.copy_(f)
oh = [torch.empty(shape, dtype=torch.float64, pin_memory=True) for _ in range(3)]
for o in oh:
    o.zero_()
s_in, s_comp, s_out, s_out2 = (torch.cuda.Stream() for _ in range(4))


def fence():
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()


def ev(stream):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream)
    return e


def pipeline(cuts, hold_d2h=False, two_out=False):
    """HostGradient.__call__ with milestones.  Returns events (start, h2d done, xy d2h done, z kernel done, end)."""
    solvers = {}
    for a, b in cuts:
        solvers.setdefault(b - a, C.CompactFiniteDifferenceSolver((b - a, N, N)))

    def run():
        cur = torch.cuda.current_stream()
        e0 = ev(cur)
        for s in (s_in, s_comp, s_out, s_out2):
            s.wait_stream(cur)
        comp_done = []
        for a, b in cuts:
            with torch.cuda.stream(s_in):
                f[a:b].copy_(fh[a:b], non_blocking=True)
                e_in = ev(s_in)
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(e_in)
                solvers[b - a].dfdxy(f[a:b], h, h, d[0][a:b], d[1][a:b])
                comp_done.append(ev(s_comp))
            if not hold_d2h:
                with torch.cuda.stream(s_out):
                    s_out.wait_event(comp_done[-1])
                    oh[0][a:b].copy_(d[0][a:b], non_blocking=True)
                so = s_out2 if two_out else s_out
                with torch.cuda.stream(so):
                    so.wait_event(comp_done[-1])
                    oh[1][a:b].copy_(d[1][a:b], non_blocking=True)
        e_h2d = ev(s_in)
        with torch.cuda.stream(s_comp):
            ddz(f, d[2])
            e_z = ev(s_comp)
        if hold_d2h:
            with torch.cuda.stream(s_out):
                s_out.wait_event(e_h2d)
                for (a, b), e_c in zip(cuts, comp_done):
                    s_out.wait_event(e_c)
                    oh[0][a:b].copy_(d[0][a:b], non_blocking=True)
                    oh[1][a:b].copy_(d[1][a:b], non_blocking=True)
        e_xy = ev(s_out)
        with torch.cuda.stream(s_out):
            s_out.wait_event(e_z)
            if two_out:
                s_out.wait_stream(s_out2)
            oh[2].copy_(d[2], non_blocking=True)
        e1 = ev(s_out)
        s_out.synchronize()
        cur.wait_stream(s_out)
        return e0, e_h2d, e_xy, e_z, e1
    return run


def ceiling():
    cur = torch.cuda.current_stream()
    e0 = ev(cur)
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    with torch.cuda.stream(s_in):
        f.copy_(fh, non_blocking=True)
        e_h2d = ev(s_in)
    with torch.cuda.stream(s_out):
        oh[0].copy_(d[0], non_blocking=True)
        oh[1].copy_(d[1], non_blocking=True)
        e_xy = ev(s_out)
        oh[2].copy_(d[2], non_blocking=True)
        e1 = ev(s_out)
    torch.cuda.synchronize()
    return e0, e_h2d, e_xy, e_xy, e1


def sequential():
    cur = torch.cuda.current_stream()
    e0 = ev(cur)
    f.copy_(fh, non_blocking=True)
    e_h2d = ev(cur)
    ddz.gradient(f, h, h, d)
    e_z = ev(cur)
    oh[0].copy_(d[0], non_blocking=True)
    oh[1].copy_(d[1], non_blocking=True)
    e_xy = ev(cur)
    oh[2].copy_(d[2], non_blocking=True)
    e1 = ev(cur)
    cur.synchronize()
    return e0, e_h2d, e_xy, e_z, e1


HG = C.HostGradient
variants = [("ceiling", ceiling),
            ("pipelined (thin first slabs)", pipeline(HG.slab_cuts(nz, 8, True))),
            ("uniform 8 slabs", pipeline(HG.slab_cuts(nz, 8, False))),
            ("uniform 4 slabs", pipeline(HG.slab_cuts(nz, 4, False))),
            ("uniform 16 slabs", pipeline(HG.slab_cuts(nz, 16, False))),
            ("thin first, x and y results on two D2H streams", pipeline(HG.slab_cuts(nz, 8, True), two_out=True)),
            ("h2d-first (D2H held until the H2D is done)", pipeline(HG.slab_cuts(nz, 8, True), hold_d2h=True)),
            ("sequential", sequential)]
if rank == 0:
    print(f"{world} GPUs, slab {shape} per rank: {f.numel() * 8 / 2**30:.1f} GiB in, {3 * f.numel() * 8 / 2**30:.1f} GiB out "
          f"per rank and step; wall ms (max over ranks) | milestones: h2d done, x/y d2h done, d/dz kernel done, end", flush=True)
ddz(f, d[2])
fence()
for name, fn in variants:
    fn()
    fence()
    rows = []
    for _ in range(reps):
        fence()
        t0 = time.perf_counter()
        evs = fn()
        fence()
        wall = (time.perf_counter() - t0) * 1e3
        ms = [evs[0].elapsed_time(e) for e in evs[1:]]
        t = torch.tensor([wall] + ms, dtype=torch.float64, device=dev)
        tmin = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        rows.append((t.tolist(), tmin.tolist()))
    if rank == 0:
        best = min(rows, key=lambda r: r[0][0])
        mx, mn = best
        print(f"{name:50s} wall {' '.join(f'{r[0][0]:6.1f}' for r in rows)} | max over ranks: "
              + " ".join(f"{v:6.1f}" for v in mx[1:]) + " | min over ranks: " + " ".join(f"{v:6.1f}" for v in mn[1:])
              + f" | {3 * N ** 3 / (mx[0] * 1e-3) / 1e9:.2f} Gpts/s per derivative", flush=True)
ddz.close()
dist.destroy_process_group()
