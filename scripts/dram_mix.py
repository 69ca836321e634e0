"""Streaming yardsticks for the read : write mixes of the library's launches (scripts/dram_mix.cu): a plain copy kernel
(8 B read + 8 B written per point, the mix of a single derivative) and a one-read-two-writes kernel (the mix of the
fused x/y launch) at 512^3 and 256^3 doubles, next to torch's copy_ (what MEASURED_PEAKS.json times) and the library's
launches on the same box.  usage: dram_mix.py"""
import ctypes
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import compact_finite_differences_b200 as C

SO = os.path.join(ROOT, "scripts", "_ab", "libdrammix.so")
if not os.path.exists(SO):                                  # build here (nvcc cross-compiles), the .so travels with gpurun
    import subprocess
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared",
                           "-Xcompiler", "-fPIC", "-o", SO, os.path.join(ROOT, "scripts", "dram_mix.cu")])
M = ctypes.CDLL(SO)
vp = ctypes.c_void_p
M.mix_copy.argtypes = [vp, vp, ctypes.c_long, ctypes.c_int, ctypes.c_int, vp]
M.mix_r1w2.argtypes = [vp, vp, vp, ctypes.c_long, ctypes.c_int, ctypes.c_int, vp]


def timeit(fn, reps=30):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for N in (512, 256):
    nset = 4 if N <= 256 else 1
    n = N ** 3
    fs = [torch.rand((N, N, N), dtype=torch.float64, device="cuda") for _ in range(nset)]
    os_ = [[torch.empty_like(fs[0]) for _ in range(2)] for _ in range(nset)]
    k = [0]

    def rot():
        k[0] += 1
        return k[0] % nset
    st = lambda: vp(torch.cuda.current_stream().cuda_stream)
    rows = []

    def tcopy():
        i = rot(); os_[i][0].copy_(fs[i])
    rows.append(("torch copy_ (1r : 1w)", timeit(tcopy), 16))
    for blocks, threads in ((148 * 8, 256), (148 * 16, 256), (148 * 4, 512), (148 * 32, 128)):
        def kc():
            i = rot(); assert M.mix_copy(fs[i].data_ptr(), os_[i][0].data_ptr(), n, blocks, threads, st()) == 0
        rows.append((f"copy kernel {blocks}x{threads} (1r : 1w)", timeit(kc), 16))
    for blocks, threads in ((148 * 8, 256), (148 * 16, 256), (148 * 4, 512), (148 * 32, 128)):
        def kr():
            i = rot(); assert M.mix_r1w2(fs[i].data_ptr(), os_[i][0].data_ptr(), os_[i][1].data_ptr(), n, blocks, threads, st()) == 0
        rows.append((f"1 read + 2 writes kernel {blocks}x{threads}", timeit(kr), 24))
    s = C.CompactFiniteDifferenceSolver((N, N, N))
    h = 0.01
    for a, nm in enumerate("xyz"):
        fn = (s.dfdx, s.dfdy, s.dfdz)[a]
        def kd():
            i = rot(); fn(fs[i], h, os_[i][0])
        rows.append((f"library d/d{nm}", timeit(kd), 16))
    def kxy():
        i = rot(); s.dfdxy(fs[i], h, h, os_[i][0], os_[i][1])
    rows.append(("library fused d/dx + d/dy", timeit(kxy), 24))
    print(f"--- {N}^3 doubles")
    for name, ms, bpp in rows:
        print(f"{name:44s} {ms:.4f} ms  {bpp * n / ms / 1e6:8.0f} GB/s", flush=True)
    del fs, os_
