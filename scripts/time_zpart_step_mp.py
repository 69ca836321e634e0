"""torchrun --nproc-per-node P scripts/time_zpart_step_mp.py [nz_local] [N] [reps]
The gradient step of a z-partitioned field ([nz_local, N, N] per rank, P ranks, real NVLink exchange through
cfd_zpart / CUDA IPC), three ways:
    fused   ZPartitionedDerivative.gradient -> cfd_zpart_apply_xyz (edge items inside the x/y kernel; 3 launches)
    chain   exchange chain on a side stream beside the x/y launch (5 / 6 xy warps; the round-1 step)
    serial  edge + reduce, x/y, coupled d/dz on one stream
ms per step = max over ranks of the CUDA-event time of `reps` back-to-back steps (barrier + sync on both sides)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nzl = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
h = 2 * np.pi / (nzl * world - 1)
f = torch.rand((nzl, N, N), dtype=torch.float64, device=dev)
out = [torch.empty_like(f) for _ in range(3)]
xy = C.CompactFiniteDifferenceSolver((nzl, N, N))
ddz = C.ZPartitionedDerivative((nzl, N, N), h, 2, mode="fused", comm="nvlink")


def zx():
    ddz.gradient(f, h, h, out)


def zx2():
    os.environ["CFD_ZX_TWO_STREAMS"] = "1"
    ddz.gradient(f, h, h, out)
    os.environ.pop("CFD_ZX_TWO_STREAMS")


def fused():
    os.environ["CFD_NO_ZX"] = "1"
    ddz.gradient(f, h, h, out)
    os.environ.pop("CFD_NO_ZX")


def chain(warps):
    def run():
        ddz.begin(f)
        xy.dfdxy(f, h, h, out[0], out[1], warps=warps)
        ddz(f, out[2])
    return run


def serial():
    ddz._zp.begin(f)
    xy.dfdxy(f, h, h, out[0], out[1], warps=6)
    ddz(f, out[2])


def timeit(name, fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    lo = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{name:52s} {t.item():.4f} ms/step (fastest rank {lo.item():.4f})  "
              f"{3 * nzl * world * N * N / t.item() / 1e6:.0f} Mpts/s per derivative, whole job", flush=True)


zx()
assert ddz.comm == "nvlink", ddz.comm
if rank == 0:
    print(f"P = {world}, slab [{nzl}, {N}, {N}] per rank, {reps} reps")
ref = None
for rnd in range(2):
    for name, fn in (("zx: x/y launch + one-kernel d/dz (default)", zx), ("zx, d/dz on a side stream (CFD_ZX_TWO_STREAMS)", zx2),
                     ("fused: edge items in the x/y kernel, 3 launches", fused),
                     ("chain beside xy, 5 warps (round-1 step)", chain(5)),
                     ("chain beside xy, 6 warps", chain(6)), ("serial (edge+reduce, xy, z)", serial)):
        timeit(name, fn)
        torch.cuda.synchronize()
        if ref is None:
            ref = [o.clone() for o in out]
        else:
            worst = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(out, ref))
            t = torch.tensor([worst], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0 and t.item() > 1e-13:
                print(f"   RESULTS DIFFER from the first variant: rel {t.item():.2e}", flush=True)
timeit("x/y launch alone (no exchange)", lambda: xy.dfdxy(f, h, h, out[0], out[1], warps=6))
timeit("partitioned d/dz alone (cfd_zpart_apply)", lambda: ddz(f, out[2]))
assert C.lib().cfd_async_status() == 0
dist.barrier()
ddz.close()
dist.destroy_process_group()
