"""
The multi-process d/dz path as it runs on a multi-GPU box -- one process per rank, neighbours' receive buffers mapped
with CUDA IPC handles (cfd_zpart_*), kernels of different PROCESSES synchronising through system-scope flags -- on
whatever GPUs the box has: with fewer GPUs than ranks the ranks share devices (IPC works between processes on one
device too, the contexts are time-sliced), so the real protocol is exercised on the driver's 1-GPU test box.
torch.distributed (gloo) only carries the 64-byte handles and the final barrier.
"""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, shape, h, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(rank % ndev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import compact_finite_differences_b200 as C
        from oracle import cfd_oracle as O
        C.lib().cfd_set_wait_timeout_ms(60000)
        rng = np.random.default_rng(99)
        n = shape[0] // world
        op = C.ZPartitionedDerivative((n,) + tuple(shape[1:]), h, 2, mode="fused", comm="nvlink")
        errs = []
        for it in range(4):                                   # both parities, twice
            f = rng.random(shape) if it < 3 else 1e3 * np.sin(np.arange(np.prod(shape)).reshape(shape) * 1e-3)
            fl = torch.from_numpy(f[rank * n:(rank + 1) * n].copy()).cuda()
            want = [O.derivative(f, a, (0.7 * h, 1.3 * h, h)[a])[rank * n:(rank + 1) * n] for a in range(3)]
            if it % 2 == 0:
                got = op(fl)                                  # cfd_zpart_apply: edge push -> reduce -> coupled
                errs.append(float(np.abs(got.cpu().numpy() - want[2]).max() / np.abs(want[2]).max()))
            else:
                if it == 3:
                    op.begin(fl)                              # exchange started early on a side stream, then ignored
                gx, gy, gz = op.gradient(fl, 0.7 * h, 1.3 * h)    # cfd_zpart_apply_xyz: edge items inside the x/y kernel
                for g, w in zip((gx, gy, gz), want):
                    errs.append(float(np.abs(g.cpu().numpy() - w).max() / np.abs(w).max()))
        assert op.comm == "nvlink", f"fell back to {op.comm}: the IPC path was not exercised"
        # the distributed npts method over the same IPC plumbing
        op2 = C.ZPartitionedDerivative((n,) + tuple(shape[1:]), h, 2, mode="npts")
        for it in range(2):
            f = rng.random(shape)
            fl = torch.from_numpy(f[rank * n:(rank + 1) * n].copy()).cuda()
            w = O.derivative(f, 2, h)[rank * n:(rank + 1) * n]
            errs.append(float(np.abs(op2(fl).cpu().numpy() - w).max() / np.abs(w).max()))
        assert C.lib().cfd_async_status() == 0
        ret[rank] = max(errs)
        op2.close()
        op.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (2 * 70, 64, 96)), (3, (3 * 66, 32, 40))])
def test_zpart_ipc_processes(world, shape):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, 0.11, ret)) for r in range(world)]
    [p.start() for p in procs]
    for p in procs:
        p.join(timeout=600)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("worker did not finish")
        assert p.exitcode == 0
    assert len(ret) == world
    assert max(ret.values()) <= TOL, dict(ret)
