"""torchrun --nproc-per-node P scripts/partition_breakdown.py [N] : per-stage times of the partitioned d/dz."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C
from compact_finite_differences_b200.partition import (exchange_halo_planes, exchange_interface_planes,
                                                       gather_interface_planes)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nl = N // world
h = 2 * np.pi / (N - 1)
f = torch.rand((nl, N, N), dtype=torch.float64, device=dev)
out = torch.empty_like(f)


def timed(name, fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name:48s} {t.item():8.4f} ms", flush=True)


plain = C.CompactFiniteDifferenceSolver((nl, N, N), h, 2)
timed("single-rank kernel on the slab (no coupling)", lambda: plain(f, out))
opn = C.ZPartitionedDerivative((nl, N, N), h, 2, mode="fused", comm="nvlink")
opn(f, out)
timed("[nvlink] exchange chain only (cfd_zpart_begin)", lambda: opn._zp.begin(f))
timed("[nvlink] full d/dz (cfd_zpart_apply)", lambda: opn(f, out))
g3 = [torch.empty_like(f) for _ in range(3)]
timed("[nvlink] slab gradient, three launches (cfd_zpart_apply_xyz)", lambda: opn.gradient(f, h, h, g3))
del g3
for comm in ("pairwise", "allgather"):
    op = C.ZPartitionedDerivative((nl, N, N), h, 2, mode="fused", comm=comm)
    op(f, out)
    res = op._exchange(f)
    timed(f"[{comm}] exchange chain only", lambda: op._exchange(f))
    timed(f"[{comm}] coupled kernel only", lambda: op.solver.apply_coupled(f, out, res[0], res[1], res[2]))
    timed(f"[{comm}] full d/dz", lambda: op(f, out))
op = C.ZPartitionedDerivative((nl, N, N), h, 2, mode="fused", comm="pairwise")
lo_buf, hi_buf, faces, faces_all, faces_nb, pv, own = op._buffers(f)
timed("NCCL halo send/recv (1 plane each way)", lambda: exchange_halo_planes(f[0], f[-1], rank, world, None, lo_buf, hi_buf))
hl, hh = exchange_halo_planes(f[0], f[-1], rank, world, None, lo_buf, hi_buf)
timed("edge_faces kernel", lambda: op.solver.edge_faces(f, faces, hl, hh))
timed("NCCL interface send/recv", lambda: exchange_interface_planes(faces_nb, own, pv, rank, world))
timed("NCCL interface all-gather", lambda: gather_interface_planes(faces, world, None, faces_all))
timed("reduced_unknowns kernel (neighbour-only)", lambda: op.solver.reduced_unknowns(faces_nb, op._ab, neighbours_only=True))
timed("reduced_unknowns kernel (all 2P planes)", lambda: op.solver.reduced_unknowns(faces_all, op._ab))
dist.barrier()
opn.close()
dist.destroy_process_group()
