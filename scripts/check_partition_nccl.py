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
for mode, comm in [("fused", c) for c in COMMS] + [("reference", "allgather"), ("npts", "nvlink")]:
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

# the whole gradient of the slab in three launches (cfd_zpart_apply_xyz: edge-face items inside the fused x/y kernel)
refx = C.CompactFiniteDifferenceSolver((N, N, N), h, 0)(full)[rank * nl:(rank + 1) * nl]
refy = C.CompactFiniteDifferenceSolver((N, N, N), h, 1)(full)[rank * nl:(rank + 1) * nl]
opg = C.ZPartitionedDerivative((nl, N, N), h, 2, mode="fused", comm="nvlink")
for it in range(3):
    g3 = opg.gradient(slab, h, h, o3)
torch.cuda.synchronize()
report(f"fused gradient (zpart_apply_xyz, comm={opg.comm}) d/dx vs single-GPU", ((g3[0] - refx).abs().max() / refx.abs().max()).item())
report(f"fused gradient (zpart_apply_xyz, comm={opg.comm}) d/dy vs single-GPU", ((g3[1] - refy).abs().max() / refy.abs().max()).item())
report(f"fused gradient (zpart_apply_xyz, comm={opg.comm}) d/dz vs single-GPU", ((g3[2] - ref).abs().max() / ref.abs().max()).item())
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    opg.gradient(slab, h, h, o3)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"fused gradient step {N}^3 on {world} GPUs (three launches): {ms.item():.3f} ms -> "
          f"{3 * N ** 3 / ms.item() * 1e3:.3e} pts/s per derivative", flush=True)
del refx, refy, full

# (b2) the reference's own call shapes on a multi-rank line (code/cuda/compact.py:18-44): the operator built from a
#      line_da, (i) one dfdz call that does everything inside, (ii) the five stages one by one
lda = C.LineDA((nl, N, N), rank, world, direction=2)
cfd = C.CompactFiniteDifferenceSolver(lda)
x_d = torch.empty_like(slab)
cfd.dfdz(slab, h, x_d, None)                                          # dfdx(f_d, dx, x_d, f_local_d) of the reference
report("reference call shape: CompactFiniteDifferenceSolver(line_da).dfdz(f, dz, x, f_local) vs single-GPU",
       ((x_d - ref).abs().max() / ref.abs().max()).item())
halo_lo, halo_hi = C.exchange_halo_planes(slab[0].contiguous(), slab[-1].contiguous(), rank, world)
cfd2 = C.CompactFiniteDifferenceSolver(C.LineDA((nl, N, N), rank, world, direction=2))
cfd2.compute_RHS(slab, h, x_d, None, halo_lo=halo_lo, halo_hi=halo_hi)
x_UH, x_LH = cfd2.solve_secondary_systems()
cfd2.solve_primary_system(x_d)
alpha, beta = cfd2.solve_reduced_system(x_UH, x_LH, x_d)
cfd2.sum_solutions(x_UH, x_LH, x_d, alpha, beta)
report("reference stages: compute_RHS, solve_secondary_systems, solve_primary_system, solve_reduced_system, "
       "sum_solutions vs single-GPU", ((x_d - ref).abs().max() / ref.abs().max()).item())
del x_d

# (c) Cartesian process grids (grid.DA): x-, y- and z-partitioned lines through PartitionedDerivative, field
#     generated per block on the device with DA_arange, result gathered with DA_gather_blocks and checked on rank 0
#     against the oracle on the global field (the reference's 2x2x2 test layout, code/cuda/test/test_compact.py:19-57)
GRIDS = {2: [(1, 1, 2), (1, 2, 1), (2, 1, 1)], 4: [(1, 2, 2), (2, 2, 1), (2, 1, 2)], 8: [(2, 2, 2)]}.get(world, [])
local = (68, 72, 80)
for proc_sizes in GRIDS:
    da = C.DA(None, local, proc_sizes)
    x, y, z = C.DA_arange(da, (0.0, 2 * np.pi), (0.0, 2 * np.pi), (0.0, 2 * np.pi), device=dev)
    fb = (x * torch.cos(x * y) + y * torch.sin(z)).contiguous()          # reference demo field, run.py:29-30
    NZ, NY, NX = da.global_dims
    hs = (2 * np.pi / (NX - 1), 2 * np.pi / (NY - 1), 2 * np.pi / (NZ - 1))
    fg = C.DA_gather_blocks(da, fb)
    for direction in range(3):
        for mode, comm in (("fused", "pairwise"), ("fused", "allgather"), ("reference", "allgather")):
            if da.line(direction)[2] == 1 and (mode, comm) != ("fused", "pairwise"):
                continue
            op = da.derivative(direction, hs[direction], mode=mode, comm=comm)
            got = C.DA_gather_blocks(da, op(fb))
            err = 0.0
            if rank == 0:
                want = O.derivative(fg.cpu().numpy(), direction, hs[direction])
                err = float(np.abs(got.cpu().numpy() - want).max() / np.abs(want).max())
            report(f"process grid {proc_sizes} local {local} d/d{'xyz'[direction]} {mode}/{comm} vs oracle", err)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
