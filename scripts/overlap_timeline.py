"""torchrun --nproc-per-node P scripts/overlap_timeline.py nz_local [N]
Timeline of one gradient step on a z-partitioned field ([nz_local, N, N] per rank): when the d/dz exchange chain
(side stream), the fused d/dx + d/dy launch and the coupled d/dz launch finish, relative to the start of the step.
Variants: xy warps/SM, chain started before / after the xy launch, side-stream priority."""
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
h = 0.01
f = torch.rand((nzl, N, N), dtype=torch.float64, device=dev)
out = [torch.empty_like(f) for _ in range(3)]
xy = C.CompactFiniteDifferenceSolver((nzl, N, N))
ddz = C.ZPartitionedDerivative((nzl, N, N), h, 2, mode="fused", comm="nvlink")
ddz(f, out[2])
torch.cuda.synchronize()
dist.barrier()


def run(label, warps, order, prio, reps=20):
    C.lib().cfd_set_launch(warps, 0, 0)
    if prio is not None:
        ddz._side = torch.cuda.Stream(device=dev, priority=prio)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    acc = np.zeros(4)
    for it in range(reps + 3):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e_chain, e_xy, e_z = ev(), ev(), ev(), ev()
        e0.record()
        if order == "chain-first":
            ddz.begin(f)
            e_chain.record(ddz._side)
            xy.dfdxy(f, h, h, out[0], out[1])
        elif order == "xy-first":
            xy.dfdxy(f, h, h, out[0], out[1])
            ddz.begin(f)
            e_chain.record(ddz._side)
        else:                           # no overlap: the chain runs inside ddz() after xy
            xy.dfdxy(f, h, h, out[0], out[1])
            e_chain.record()
        e_xy.record()
        ddz(f, out[2])
        e_z.record()
        torch.cuda.synchronize()
        if it >= 3:
            acc += [e0.elapsed_time(e_chain), e0.elapsed_time(e_xy), e0.elapsed_time(e_z), e_xy.elapsed_time(e_z)]
    t = torch.tensor(acc / reps, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label:46s} chain done {t[0].item():.3f}  xy done {t[1].item():.3f}  step done {t[2].item():.3f}  "
              f"(z after xy {t[3].item():.3f}) ms", flush=True)
    C.lib().cfd_set_launch(0, 0, 0)


for bs in os.environ.get("EDGE_BS", "128").split(","):
    os.environ["CFD_EDGE_BS"] = bs
    for warps in (6, 5):
        run(f"edge bs={bs}, xy w{warps}, chain first", warps, "chain-first", 0)
        run(f"edge bs={bs}, xy w{warps}, chain first, high priority", warps, "chain-first", -1)
run("xy w6, no overlap", 6, "none", 0)
dist.barrier()
dist.destroy_process_group()
