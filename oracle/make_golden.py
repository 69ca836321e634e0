"""
Regenerate tests/golden/*.npz from the reference itself (run in the build container, where /root/reference
exists):  python -m oracle.make_golden

  npts_ref.npz      seeded random RHS blocks -> output of the reference's own nonperiodic_tridiagonal_solver
                    (lanl-implementation/npts.c, compiled unmodified, one rank) + its beta/gamma tables
  known_answer.npz  the `Average absolute error` lines printed by the reference's test_npts for
                    NX = 32, 64, 128, 256 (test_npts.c:146-157)
  derivative.npz    small smooth fields and their derivative along x, y, z computed as
                    reference-npts(RHS) with the RHS formulas of test_npts.c:86-97 / kernels.cu:34-44
"""
import os

import numpy as np

from . import cfd_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    assert O.have_ref(), "oracle/_ref is not built (make -C oracle ref)"
    os.makedirs(GOLD, exist_ok=True)
    rng = np.random.default_rng(20261018)
    d = {}
    for n in (8, 32, 48, 64, 100, 256, 1024):
        r = rng.random((2, 3, n))
        u, beta, gam = O.ref_npts_solve(r)
        d[f"r_{n}"], d[f"u_{n}"], d[f"beta_{n}"], d[f"gam_{n}"] = r, u, beta, gam
    np.savez_compressed(os.path.join(GOLD, "npts_ref.npz"), **d)

    ka = {str(n): O.ref_known_answer(n) for n in (32, 64, 128, 256)}
    np.savez(os.path.join(GOLD, "known_answer.npz"), **ka)

    # derivative fixtures: RHS by the port (pinned separately against the formulas), solve by the REFERENCE
    dd = {}
    nz, ny, nx = 12, 20, 40
    z, y, x = np.meshgrid(np.linspace(0, 2 * np.pi, nz), np.linspace(0, 2 * np.pi, ny),
                          np.linspace(0, 2 * np.pi, nx), indexing="ij")
    f = np.sin(x) * np.cos(y) * np.sin(z) + x * np.cos(x * y) + y * np.sin(z)
    dd["f"] = f
    for axis, h in ((0, x[0, 0, 1] - x[0, 0, 0]), (1, y[0, 1, 0] - y[0, 0, 0]), (2, z[1, 0, 0] - z[0, 0, 0])):
        rhs = O.rhs(f, axis, h)
        # bring the derivative axis last, run the reference's x-line solver, move it back
        ax = 2 - axis
        rt = np.ascontiguousarray(np.moveaxis(rhs, ax, 2))
        u, _, _ = O.ref_npts_solve(rt)
        dd[f"df_{axis}"] = np.ascontiguousarray(np.moveaxis(u, 2, ax))
        dd[f"h_{axis}"] = h
    np.savez_compressed(os.path.join(GOLD, "derivative.npz"), **dd)
    print("golden fixtures written to", os.path.normpath(GOLD))


if __name__ == "__main__":
    main()
