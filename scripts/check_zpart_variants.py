"""All P ranks of a z-partitioned line in ONE process on one device (cfd_zpart_connect_ptr, one stream per rank, CTA counts
capped so that the persistent kernels are co-resident): the one-kernel d/dz and the three-launch form alternately,
status word and error against the oracle after every call.  A 3 s wait time-out makes a protocol hang show up as
status -4 instead of blocking.  usage: check_zpart_variants.py P nz ny nx [ctas]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import compact_finite_differences_b200 as C
from compact_finite_differences_b200._lib import check, lib
from oracle import cfd_oracle as O
L = lib()
L.cfd_set_wait_timeout_ms(3000)
P, shape = int(sys.argv[1]), tuple(int(a) for a in sys.argv[2:5])
cta = int(sys.argv[5]) if len(sys.argv) > 5 else 148 // P
rng = np.random.default_rng(1)
hs = (0.19, 0.07, 0.23)
n = shape[0] // P
lshape = (n,) + shape[1:]
zsol = [C.CompactFiniteDifferenceSolver(lshape, hs[2], 2, part=(r, P)) for r in range(P)]
zps = []
for r in range(P):
    h = ctypes.c_void_p()
    check(L.cfd_zpart_create(ctypes.byref(h), zsol[r]._plan(2, hs[2]).handle))
    check(L.cfd_zpart_set_ctas(h, cta))
    zps.append(h)
bufs = [L.cfd_zpart_buffer(z) for z in zps]
for r in range(P):
    check(L.cfd_zpart_connect_ptr(zps[r], bufs[r - 1] if r > 0 else None, bufs[r + 1] if r < P - 1 else None))
streams = [torch.cuda.Stream() for _ in range(P)]
f = rng.random(shape)
want = O.derivative(f, 2, hs[2])
blocks = [torch.from_numpy(f[r * n:(r + 1) * n].copy()).cuda() for r in range(P)]
for name, env in (("zx", None), ("zx", None), ("old", "1"), ("zx", None), ("old", "1")):
    if env: os.environ["CFD_NO_ZX"] = env
    else: os.environ.pop("CFD_NO_ZX", None)
    outs = [torch.zeros_like(b) for b in blocks]
    torch.cuda.synchronize()
    for r in range(P):
        check(L.cfd_zpart_apply(zps[r], blocks[r].data_ptr(), outs[r].data_ptr(), ctypes.c_void_p(streams[r].cuda_stream)))
    torch.cuda.synchronize()
    st = L.cfd_async_status()
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    err = np.abs(got - want).max() / np.abs(want).max()
    bad = np.argwhere(np.abs(got - want) > 1e-9 * np.abs(want).max())
    print(name, "status", st, "rel err", err, "bad points", len(bad), bad[:3].tolist() if len(bad) else "", flush=True)
