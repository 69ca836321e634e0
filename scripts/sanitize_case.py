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
torch.cuda.synchronize()
print("sanitize case: worst rel L-inf", worst)
assert worst <= 1e-12
