"""A/B timing of differently compiled libcfd_b200 builds in ONE gpurun call (boxes differ by 3-5 %, so variants are only
comparable inside a call): every variant runs in its own child process, rounds interleaved over the variants.
    parent: ab_libs.py lib_a.so lib_b.so ...          (paths relative to the repo root; the first one is the baseline)
    child:  AB_LIB=<path> ab_libs.py --child <tag>
Cases: 512^3 d/dx, d/dy, d/dz, the fused x/y launch; the 8-GPU slab shape [128,1024,1024]: x/y launch, the one-kernel
partitioned d/dz (middle rank of 3 wired to itself, flags released: right timing, wrong neighbours).
Each child stores a sample of its results; the parent reports the largest relative deviation from the baseline's."""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(tag):
    import torch
    import compact_finite_differences_b200._lib as _lib
    _lib.LIB_PATH = os.path.join(ROOT, os.environ["AB_LIB"])
    import compact_finite_differences_b200 as C
    from compact_finite_differences_b200._lib import check, lib
    L = lib()
    reps = int(os.environ.get("AB_REPS", "30"))

    def timeit(fn):
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

    res, samples = {}, {}
    h = 0.01
    torch.manual_seed(1)
    N = 512
    f = torch.rand((N, N, N), dtype=torch.float64, device="cuda")
    o = [torch.empty_like(f) for _ in range(3)]
    s = C.CompactFiniteDifferenceSolver((N, N, N))
    for a, nm in enumerate("xyz"):
        fn = (s.dfdx, s.dfdy, s.dfdz)[a]
        res[f"512 d/d{nm}"] = timeit(lambda: fn(f, h, o[a]))
        samples[f"512 d/d{nm}"] = o[a][::61, ::7, :].clone().cpu()
    res["512 xy"] = timeit(lambda: s.dfdxy(f, h, h, o[0], o[1]))
    samples["512 xy.x"] = o[0][::61, ::7, :].clone().cpu()
    samples["512 xy.y"] = o[1][::61, ::7, :].clone().cpu()
    res["512 xy+z"] = timeit(lambda: (s.dfdxy(f, h, h, o[0], o[1]), s.dfdz(f, h, o[2])))
    del f, o, s
    shape = (128, 1024, 1024)
    f = torch.rand(shape, dtype=torch.float64, device="cuda")
    o = [torch.empty_like(f) for _ in range(3)]
    s = C.CompactFiniteDifferenceSolver(shape)
    res["slab xy"] = timeit(lambda: s.dfdxy(f, h, h, o[0], o[1]))
    res["slab z plain"] = timeit(lambda: s.dfdz(f, h, o[2]))
    zs = C.CompactFiniteDifferenceSolver(shape, h, 2, part=(1, 3))
    z = ctypes.c_void_p()
    check(L.cfd_zpart_create(ctypes.byref(z), zs._plan(2, h).handle))
    buf = L.cfd_zpart_buffer(z)
    check(L.cfd_zpart_connect_ptr(z, buf, buf))
    cudart = ctypes.CDLL("libcudart.so.12")
    cudart.cudaMemset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]
    assert cudart.cudaMemset(ctypes.c_void_p(buf + 8 * 16 * shape[1] * shape[2]), 0x7f, 16 * 8) == 0
    torch.cuda.synchronize()
    sp = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    res["slab zx"] = timeit(lambda: check(L.cfd_zpart_apply(z, f.data_ptr(), o[2].data_ptr(), sp())))
    px, py = s._plan(0, h), s._plan(1, h)
    res["slab step (xy + zx)"] = timeit(lambda: check(L.cfd_zpart_apply_xyz(
        z, px.handle, py.handle, f.data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), sp())))
    assert L.cfd_async_status() == 0
    L.cfd_zpart_destroy(z)
    del f, o
    # small fields, rotating over four input / output sets so that nothing is L2-resident
    for N in (256, 128):
        fs = [torch.rand((N, N, N), dtype=torch.float64, device="cuda") for _ in range(4)]
        os_ = [[torch.empty_like(fs[0]) for _ in range(3)] for _ in range(4)]
        s = C.CompactFiniteDifferenceSolver((N, N, N))
        k = [0]

        def rot():
            k[0] += 1
            return k[0] % 4
        for a, nm in enumerate("xyz"):
            fn = (s.dfdx, s.dfdy, s.dfdz)[a]

            def kd():
                i = rot()
                fn(fs[i], h, os_[i][a])
            res[f"{N} d/d{nm}"] = timeit(kd)

        def kxy():
            i = rot()
            s.dfdxy(fs[i], h, h, os_[i][0], os_[i][1])
        res[f"{N} xy"] = timeit(kxy)

        def kg():
            i = rot()
            s.gradient(fs[i], (h, h, h), os_[i])
        res[f"{N} gradient()"] = timeit(kg)
        del fs, os_, s
    # the general kernel (6th-order scheme)
    N = 512
    f = torch.rand((N, N, N), dtype=torch.float64, device="cuda")
    out = torch.empty_like(f)
    s6 = C.CompactFiniteDifferenceSolver((N, N, N), h, 2, scheme="compact6")
    res["512 compact6 d/dz"] = timeit(lambda: s6(f, out))
    torch.save(samples, f"/tmp/ab_{tag}.pt")
    print("AB_JSON " + json.dumps(res), flush=True)


def main():
    libs = sys.argv[1:]
    rounds = int(os.environ.get("AB_ROUNDS", "3"))
    table = {l: [] for l in libs}
    for r in range(rounds):
        for i, l in enumerate(libs):
            env = dict(os.environ, AB_LIB=l)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(i)], env=env,
                               capture_output=True, text=True, timeout=600)
            line = [x for x in p.stdout.splitlines() if x.startswith("AB_JSON ")]
            if not line:
                print(f"{l}: child failed\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}", flush=True)
                continue
            table[l].append(json.loads(line[0][8:]))
    cases = list(table[libs[0]][0].keys())
    print(f"{'case':24s} " + " ".join(f"{os.path.basename(l):>34s}" for l in libs))
    for c in cases:
        row = []
        for l in libs:
            v = [t[c] for t in table[l]]
            row.append(" ".join(f"{x:.4f}" for x in v) + f" | med {sorted(v)[len(v) // 2]:.4f}")
        print(f"{c:24s} " + "   ".join(f"{x:>34s}" for x in row), flush=True)
    import torch
    base = torch.load("/tmp/ab_0.pt")
    for i, l in enumerate(libs[1:], 1):
        other = torch.load(f"/tmp/ab_{i}.pt")
        for k in base:
            d = float((other[k] - base[k]).abs().max() / base[k].abs().max())
            print(f"{os.path.basename(l)} vs baseline, {k}: rel L-inf {d:.2e}")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        main()
