"""cfd_apply_xy (d/dx and d/dy in one launch, wavefront draw order) against the two separate launches:
bit-equality of the results and ms per pair, for several settings of the planes-in-flight knob.
usage: time_xy.py [nz ny nx]..."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


args = [int(a) for a in sys.argv[1:]] or [512, 512, 512]
shapes = [tuple(args[i:i + 3]) for i in range(0, len(args), 3)]
actives = os.environ.get("XY_ACTIVES", "default,0,4,9,14,18.5,24,32,48").split(",")
taus = os.environ.get("XY_TAUS", "").split(",")                 # start-up stagger per slot in ns ("" = off)
ctas_list = os.environ.get("XY_CTAS", "").split(",")            # "" = all SMs; "auto" = full last round; or a count
warps_list = [int(w) for w in os.environ.get("XY_WARPS", "0").split(",")]      # 0 = library default
slots_list = [int(w) for w in os.environ.get("XY_SLOTS", "0").split(",")]
for shape in shapes:
    f = torch.rand(shape, dtype=torch.float64, device="cuda")
    op = C.CompactFiniteDifferenceSolver(shape, 0.1, 0)
    gx, gy = torch.empty_like(f), torch.empty_like(f)
    C.lib().cfd_set_launch(0, 0, 0)
    rx, ry = op.dfdx(f, 0.1), op.dfdy(f, 0.2)
    t_sep = timeit(lambda: (op.dfdx(f, 0.1, gx), op.dfdy(f, 0.2, gy)))
    pts = f.numel()
    print(f"{shape}: separate x+y {t_sep:.4f} ms ({32 * pts / t_sep / 1e6:.0f} GB/s algorithmic)", flush=True)
    for w, ns, ct, tau in [(w, ns, ct, tau) for ns in slots_list for w in warps_list for ct in ctas_list for tau in taus]:
        if tau:
            os.environ["CFD_XY_TAU"] = tau
        else:
            os.environ.pop("CFD_XY_TAU", None)
        if ct:
            os.environ["CFD_XY_CTAS"] = ct
        else:
            os.environ.pop("CFD_XY_CTAS", None)
        for a in actives:
            C.lib().cfd_set_launch(w, 0, ns)
            if a == "default":
                os.environ.pop("CFD_XY_ACTIVE", None)
            else:
                os.environ["CFD_XY_ACTIVE"] = a
            gx.zero_(); gy.zero_()
            C.lib().cfd_set_launch(w, 0, ns)
            op.dfdxy(f, 0.1, 0.2, gx, gy)
            torch.cuda.synchronize()
            same = bool(torch.equal(gx, rx) and torch.equal(gy, ry))
            err = max(float((gx - rx).abs().max() / rx.abs().max()), float((gy - ry).abs().max() / ry.abs().max()))
            t = timeit(lambda: op.dfdxy(f, 0.1, 0.2, gx, gy))
            print(f"  warps={w} slots={ns} ctas={ct or 'all':>4} tau={tau or '0':>5} active={a:>7}: {t:.4f} ms ({32 * pts / t / 1e6:.0f} GB/s algorithmic, "
                  f"{t_sep / t:.3f}x) bit-equal={same} rel-diff={err:.1e}", flush=True)
    C.lib().cfd_set_launch(0, 0, 0)
    os.environ.pop("CFD_XY_ACTIVE", None)
    os.environ.pop("CFD_XY_CTAS", None)
    os.environ.pop("CFD_XY_TAU", None)
    del f, gx, gy, rx, ry
