"""Small end-to-end case for compute-sanitizer: every kernel of the library once, on ragged shapes."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compact_finite_differences_b200 as C
from oracle import cfd_oracle as O

rng = np.random.default_rng(0)
worst = 0.0
for shape in [(5, 7, 34), (40, 36, 70)]:
    f = rng.random(shape)
    fd = torch.from_numpy(f).cuda()
    for axis in range(3):
        got = C.CompactFiniteDifferenceSolver(shape, 0.1, axis)(fd).cpu().numpy()
        want = O.derivative(f, axis, 0.1)
        worst = max(worst, np.abs(got - want).max() / np.abs(want).max())
        t = fd.clone()
        C.NearToeplitzSolver(shape, O.PADE, axis=axis).solve(t)
        want = O.near_toeplitz_solve(f, O.PADE, axis)
        worst = max(worst, np.abs(t.cpu().numpy() - want).max() / np.abs(want).max())
# two-pass solver
d = rng.random((3, 4, 130))
t = torch.from_numpy(d).cuda()
C.NearToeplitzSolver(d.shape, (1., 2., 1 / 3, 1., 1 / 3, 2., 1.)).solve(t)
want = O.near_toeplitz_solve(d, (1., 2., 1 / 3, 1., 1 / 3, 2., 1.))
worst = max(worst, np.abs(t.cpu().numpy() - want).max() / np.abs(want).max())
# partitioned line, both paths, all ranks on this device
P, shape, h = 2, (140, 6, 34), 0.2
f = rng.random(shape)
want = O.derivative(f, 2, h)
n = shape[0] // P
plane = shape[1] * shape[2]
blocks = [torch.from_numpy(f[r * n:(r + 1) * n].copy()).cuda() for r in range(P)]
sol = [C.CompactFiniteDifferenceSolver((n,) + shape[1:], h, 2, part=(r, P)) for r in range(P)]
faces = torch.zeros((2 * P, plane), dtype=torch.float64, device="cuda")
halos = [(None if r == 0 else blocks[r - 1][-1].contiguous(), None if r == P - 1 else blocks[r + 1][0].contiguous())
         for r in range(P)]
outs = []
for r in range(P):
    o = sol[r].apply_local(blocks[r], None, *halos[r])
    sol[r].interface_pack(o, faces[2 * r:2 * r + 2])
    outs.append(o)
for r in range(P):
    sol[r].reduced_correct(outs[r], faces)
got = torch.cat(outs).cpu().numpy()
worst = max(worst, np.abs(got - want).max() / np.abs(want).max())
faces2 = torch.zeros_like(faces)
ab = torch.empty((2, plane), dtype=torch.float64, device="cuda")
outs = []
for r in range(P):
    sol[r].edge_faces(blocks[r], faces2[2 * r:2 * r + 2], *halos[r])
for r in range(P):
    sol[r].reduced_unknowns(faces2, ab)
    outs.append(sol[r].apply_coupled(blocks[r], None, halos[r][0], halos[r][1], ab))
got = torch.cat(outs).cpu().numpy()
worst = max(worst, np.abs(got - want).max() / np.abs(want).max())
# round 2: fused x/y launch, gradient on two streams, general kernel (both look-aheads, both stencils), starved
# in-place solver with its side buffer, and the C-side zpart driver incl. the fused x/y + edge launch
shape = (6, 64, 96)
g = rng.random(shape)
hs = (0.1, 0.2, 0.3)
got3 = C.CompactFiniteDifferenceSolver(shape).gradient(torch.from_numpy(g).cuda(), hs)
for a in range(3):
    want = O.derivative(g, a, hs[a])
    worst = max(worst, np.abs(got3[a].cpu().numpy() - want).max() / np.abs(want).max())
for scheme in ("compact6", "pade4-d2"):
    for axis, shp in ((0, (3, 5, 130)), (1, (2, 97, 34)), (2, (161, 3, 34))):
        q = rng.random(shp)
        got = C.CompactFiniteDifferenceSolver(shp, 0.3, axis, scheme=scheme)(torch.from_numpy(q).cuda()).cpu().numpy()
        want = O.scheme_derivative(q, axis, 0.3, scheme)
        worst = max(worst, np.abs(got - want).max() / np.abs(want).max())
d = rng.random((1, 40, 1000))
t = torch.from_numpy(d).cuda()
C.NearToeplitzSolver(d.shape, O.PADE).solve(t)
want = O.near_toeplitz_solve(d, O.PADE)
worst = max(worst, np.abs(t.cpu().numpy() - want).max() / np.abs(want).max())
import ctypes
from compact_finite_differences_b200._lib import check, lib
L = lib()
P, shape, h = 3, (3 * 66, 32, 64), 0.2
f = rng.random(shape)
n = shape[0] // P
lshape = (n,) + shape[1:]
zs = [C.CompactFiniteDifferenceSolver(lshape, h, 2, part=(r, P)) for r in range(P)]
xy = C.CompactFiniteDifferenceSolver(lshape)
px, py = xy._plan(0, 0.1), xy._plan(1, 0.15)
zps = []
for r in range(P):
    z = ctypes.c_void_p()
    check(L.cfd_zpart_create(ctypes.byref(z), zs[r]._plan(2, h).handle))
    check(L.cfd_zpart_set_ctas(z, 148 // P))
    zps.append(z)
bufs = [L.cfd_zpart_buffer(z) for z in zps]
for r in range(P):
    check(L.cfd_zpart_connect_ptr(zps[r], bufs[r - 1] if r > 0 else None, bufs[r + 1] if r < P - 1 else None))
blocks = [torch.from_numpy(f[r * n:(r + 1) * n].copy()).cuda() for r in range(P)]
outs = [[torch.empty_like(b) for _ in range(3)] for b in blocks]
streams = [torch.cuda.Stream() for _ in range(P)]
torch.cuda.synchronize()
for it in range(2):
    for r in range(P):
        sp = ctypes.c_void_p(streams[r].cuda_stream)
        if it == 0:
            check(L.cfd_zpart_apply(zps[r], blocks[r].data_ptr(), outs[r][2].data_ptr(), sp))
        else:
            check(L.cfd_zpart_apply_xyz(zps[r], px.handle, py.handle, blocks[r].data_ptr(), outs[r][0].data_ptr(),
                                        outs[r][1].data_ptr(), outs[r][2].data_ptr(), sp))
    torch.cuda.synchronize()
    got = np.concatenate([o[2].cpu().numpy() for o in outs], axis=0)
    want = O.derivative(f, 2, h)
    worst = max(worst, np.abs(got - want).max() / np.abs(want).max())
for z in zps:
    L.cfd_zpart_destroy(z)
torch.cuda.synchronize()
print("sanitize case: worst rel L-inf", worst)
assert worst <= 1e-12
