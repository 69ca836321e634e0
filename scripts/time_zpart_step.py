"""One rank's step of the z-partitioned gradient on ONE GPU: the middle rank of a 3-way split ([nz_local, N, N] slab,
interior-type ends on both sides) wired to ITSELF as both neighbours (what it stores for its neighbours lands in its own
receive arrays: wrong numbers, right timing) and with its arrival flags pre-released, so that the per-rank work of the
8-GPU run can be timed (and profiled with ncu) without 8 GPUs:
    zx      cfd_zpart_apply_xyz: x/y launch + the one-kernel d/dz (kernels_zx.cuh)                      (2 launches)
    fused   CFD_NO_ZX=1: x/y kernel with the edge items first, reduced solve, coupled d/dz              (3 launches)
    chain   cfd_zpart_begin on a side stream, cfd_apply_xy (5 or 6 warps), cfd_zpart_apply              (round-1 step)
    serial  the same launches on one stream
usage: time_zpart_step.py [nz_local] [N] [reps]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C
from compact_finite_differences_b200._lib import check, lib

nzl = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
only = os.environ.get("ZSTEP_ONLY", "")
if os.environ.get("ZX_WARPS"):
    lib().cfd_set_launch(int(os.environ["ZX_WARPS"]), 0, 0)
L = lib()
h = 0.01
shape = (nzl, N, N)
f = torch.rand(shape, dtype=torch.float64, device="cuda")
out = [torch.empty_like(f) for _ in range(3)]
zs = [C.CompactFiniteDifferenceSolver(shape, h, 2, part=(r, 3)) for r in range(3)]
xy = C.CompactFiniteDifferenceSolver(shape)
px, py = xy._plan(0, h), xy._plan(1, h)
zps = []
z = ctypes.c_void_p()
check(L.cfd_zpart_create(ctypes.byref(z), zs[1]._plan(2, h).handle))
zps.append(z)
bufs = [None, L.cfd_zpart_buffer(z)]
check(L.cfd_zpart_connect_ptr(z, bufs[1], bufs[1]))
# release the middle rank's arrival flags for good: every wait passes at once
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaMemset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]
plane = N * N
assert cudart.cudaMemset(ctypes.c_void_p(bufs[1] + 8 * 16 * plane), 0x7f, 16 * 8) == 0
torch.cuda.synchronize()
zp = zps[0]
side = torch.cuda.Stream()


def sp(s=None):
    return ctypes.c_void_p((s or torch.cuda.current_stream()).cuda_stream)


def xyz():
    check(L.cfd_zpart_apply_xyz(zp, px.handle, py.handle, f.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                                out[2].data_ptr(), sp()))


def xyz2():
    os.environ["CFD_ZX_TWO_STREAMS"] = "1"
    xyz()
    os.environ.pop("CFD_ZX_TWO_STREAMS")


def fused():
    os.environ["CFD_NO_ZX"] = "1"
    xyz()
    os.environ.pop("CFD_NO_ZX")


def chain_begin(stream):
    check(L.cfd_zpart_begin(zp, f.data_ptr(), sp(stream)))


def chain(warps):
    def run():
        side.wait_stream(torch.cuda.current_stream())
        check(L.cfd_zpart_begin(zp, f.data_ptr(), sp(side)))
        xy.dfdxy(f, h, h, out[0], out[1], warps=warps)
        check(L.cfd_zpart_apply(zp, f.data_ptr(), out[2].data_ptr(), sp()))
    return run


def serial():
    check(L.cfd_zpart_begin(zp, f.data_ptr(), sp()))
    xy.dfdxy(f, h, h, out[0], out[1], warps=6)
    check(L.cfd_zpart_apply(zp, f.data_ptr(), out[2].data_ptr(), sp()))


def timeit(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:54s} {ms:.4f} ms/step  ({3 * f.numel() / ms / 1e6:.0f} Mpts/s per derivative)", flush=True)
    return ms


print(f"slab {shape}, middle rank of 3, {reps} reps")
cases = [("zx (zpart_apply_xyz: x/y launch + one-kernel d/dz)", xyz),
         ("zx two streams (CFD_ZX_TWO_STREAMS)", xyz2),
         ("zx (again)", xyz),
         ("zx two streams (again)", xyz2),
         ("zx d/dz alone (cfd_zpart_apply, one launch)", lambda: check(L.cfd_zpart_apply(zp, f.data_ptr(), out[2].data_ptr(), sp()))),
         ("fused (CFD_NO_ZX: edge items in the x/y kernel)", fused),
         ("chain on side stream, xy 5 warps (round-1 step)", chain(5)),
         ("chain on side stream, xy 6 warps", chain(6)),
         ("serial: edge+reduce, xy, coupled z", serial),
         ("xy alone (cfd_apply_xy, 6 warps)", lambda: xy.dfdxy(f, h, h, out[0], out[1], warps=6)),
         ("edge + reduce alone (cfd_zpart_begin)", lambda: check(L.cfd_zpart_begin(zp, f.data_ptr(), sp()))),
         ("plain d/dz of the slab (no coupling)", lambda: xy.dfdz(f, h, out[2]))]
ref = None
for name, fn in cases:
    if only and only not in name:
        continue
    timeit(name, fn)
    if name.startswith("fused") or name.startswith("serial"):
        torch.cuda.synchronize()
        cur = [o.clone() for o in out]
        if ref is None:
            ref = cur
        else:
            print("   bit-equal to the fused step:", all(torch.equal(a, b) for a, b in zip(ref, cur)))
    if name.startswith("zx (") :
        torch.cuda.synchronize()
        zxres = [o.clone() for o in out]
    if name.startswith("fused"):
        print("   zx vs fused d/dz rel diff:", float((zxres[2] - out[2]).abs().max() / out[2].abs().max()))
assert L.cfd_async_status() == 0
for z in zps:
    L.cfd_zpart_destroy(z)
