"""
torchrun --nproc-per-node P scripts/check_partition_nccl.py [N]
z-partitioned derivative on P GPUs (NCCL) against (a) the oracle on a small random field, (b) the single-GPU
kernel on the full N^3 smooth field (each rank recomputes the whole line set on its own GPU and compares its slab).
Prints one line per check on rank 0; exits non-zero on a parity failure.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C
from oracle import cfd_oracle as O

TOL = 1e-12
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ok = True


def report(name, err):
    global ok
    t = torch.tensor([err], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    good = t.item() <= TOL
    ok = ok and good
    if rank == 0:
        print(f"{name}: max rel L-inf over ranks = {t.item():.3e}  {'OK' if good else 'FAIL'}", flush=True)


# (a) small random field vs the oracle, all three directions
rng = np.random.default_rng(7)
shape = (32 * world, 24, 40)
f = rng.random(shape)
n = shape[0] // world
fl = torch.from_numpy(f[rank * n:(rank + 1) * n].copy()).to(dev)
for axis in range(3):
    h = 0.1 + 0.05 * axis
    got = C.ZPartitionedDerivative((n, shape[1], shape[2]), h, axis)(fl).cpu().numpy()
    want = O.derivative(f, axis, h)[rank * n:(rank + 1) * n]
    report(f"random {shape} axis {axis} P={world} vs oracle", np.abs(got - want).max() / np.abs(want).max())

# (b) N^3 smooth field: partitioned d/dz vs the single-GPU kernel on the whole field
h = 2 * np.pi / (N - 1)
t1 = torch.arange(N, dtype=torch.float64, device=dev) * h
full = (torch.sin(t1)[None, None, :] * torch.cos(t1)[None, :, None] * torch.sin(t1)[:, None, None]
        + t1[None, None, :] * torch.cos(t1[None, None, :] * t1[None, :, None])).contiguous()
nl = N // world
ref = C.CompactFiniteDifferenceSolver((N, N, N), h, 2)(full)[rank * nl:(rank + 1) * nl]
slab = full[rank * nl:(rank + 1) * nl].contiguous()
COMMS = os.environ.get("CFD_COMMS", "nvlink,pairwise,allgather").split(",")
for mode, comm in [("fused", c) for c in COMMS] + [("reference", "allgather")]:
    op = C.ZPartitionedDerivative((nl, N, N), h, 2, mode=mode, comm=comm)
    got = op(slab)
    report(f"smooth {N}^3 d/dz P={world} mode={op.mode}/{op.comm} vs single-GPU",
           ((got - ref).abs().max() / ref.abs().max()).item())
    # timing of the partitioned d/dz
    for _ in range(3):
        op(slab, got)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        op(slab, got)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"partitioned d/dz {N}^3 on {world} GPUs, mode={op.mode}/{op.comm}: {ms.item():.3f} ms -> "
              f"{N ** 3 / ms.item() * 1e3:.3e} pts/s", flush=True)

# overlap: exchange of d/dz started before d/dx, d/dy of the same field
opx = C.ZPartitionedDerivative((nl, N, N), h, 0)
opy = C.ZPartitionedDerivative((nl, N, N), h, 1)
opz = C.ZPartitionedDerivative((nl, N, N), h, 2, mode="fused", comm=COMMS[0])
o3 = [torch.empty_like(slab) for _ in range(3)]
for overlap in (False, True):
    def step():
        if overlap:
            opz.begin(slab)
        opx(slab, o3[0]); opy(slab, o3[1]); opz(slab, o3[2])
    for _ in range(3):
        step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    report(f"gradient step overlap={overlap}: d/dz parity", ((o3[2] - ref).abs().max() / ref.abs().max()).item())
    if rank == 0:
        print(f"gradient step (x, y, z) {N}^3 on {world} GPUs, overlap={overlap}: {ms.item():.3f} ms -> "
              f"{3 * N ** 3 / ms.item() * 1e3:.3e} pts/s per derivative", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
