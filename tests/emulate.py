"""
NumPy emulation of the streaming kernel's chunk schedule (csrc/kernels.cuh: fwd_chunk / bwd_chunk / the tile
loop), driven by the SAME host tables the CUDA kernel receives (cfd_debug_tables).  Test infrastructure: it lets
the CPU suite verify the table construction, the head/mid/tail mode logic and the 32-row look-ahead
back-substitution against the oracle without a GPU.  It is not a product path.
"""
import ctypes

import numpy as np

from compact_finite_differences_b200._lib import check, lib

CH = 32
PADE = (1., 2., .25, 1., .25, 2., 1.)


def tables(n, coeffs, scale):
    out = np.zeros(6 * CH + 8)
    co = (ctypes.c_double * 7)(*[float(c) for c in coeffs])
    check(lib().cfd_debug_tables(int(n), co, float(scale), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    head = dict(sk=out[0:32], l=out[32:64], g=out[64:96])
    tail = dict(sk=out[96:128], l=out[128:160], g=out[160:192])
    s = out[192:]
    return dict(head=head, tail=tail, sk_mid=s[0], l_mid=s[1], g_mid=s[2], beta0=s[3], betan=s[4],
                K=int(s[5]), jl=int(s[6]), fast_ok=bool(s[7]))


def stream_lines(F, coeffs, h=None, lo_closure=True, hi_closure=True, halo_lo=None, halo_hi=None,
                 alpha=None, beta=None, kseg=None):
    """
    F: [nlines, n] (each row one line).  h given -> derivative (Pade RHS fused), else plain solve of F.
    Follows the kernel step by step: per chunk forward elimination with HEAD / MID / TAIL coefficients,
    32-row warm-up back-substitution from the next chunk, exact sweep at the end of the line.
    """
    F = np.asarray(F, dtype=np.float64)
    nl, n = F.shape
    deriv = h is not None
    T = tables(n, coeffs, 3. / (4 * h) if deriv else 1.0)
    K, jl = T["K"], T["jl"]
    s0c = T["beta0"] / (2 * h) if deriv else 0.0
    snc = T["betan"] / (2 * h) if deriv else 0.0
    Fp = np.zeros((nl, K * CH))
    Fp[:, :n] = F                                   # TMA zero-fills out-of-bounds rows
    out = np.full((nl, K * CH), np.nan)
    # coupled multi-rank solve (cfd_apply_coupled): interface unknowns as Dirichlet data of the block
    coupled = alpha is not None
    if coupled and not lo_closure:
        T["head"]["l"][0] = coeffs[2] * T["beta0"]                  # a_i * beta_0, as cfd_create does
    snb = T["betan"] * coeffs[4] if (coupled and not hi_closure) else 0.0
    bval = np.asarray(beta, dtype=np.float64) if coupled else np.zeros(nl)
    state = {}

    def reset_state():
        state["eprev"] = np.asarray(alpha, dtype=np.float64).copy() if (coupled and not lo_closure) else np.zeros(nl)
        state["fm1"] = np.zeros(nl) if (lo_closure or not deriv) else np.asarray(halo_lo, dtype=np.float64).copy()
        state["fm2"] = np.zeros(nl)

    reset_state()
    hval = np.zeros(nl) if (hi_closure or not deriv) else np.asarray(halo_hi, dtype=np.float64)
    eA = None

    def fwd(k, kend):
        eprev, fm1, fm2 = state["eprev"], state["fm1"], state["fm2"]
        Fc = Fp[:, k * CH:(k + 1) * CH]
        last = (k == K - 1)
        peek = Fp[:, (k + 1) * CH] if k < kend else np.zeros(nl)     # the item's next tile, if it has one
        e = np.zeros((nl, CH))
        mode = 1 if k == 0 else (2 if last else 0)
        tab = T["head"] if mode == 1 else T["tail"]
        jlast = jl if last else -1
        for j in range(CH):
            if mode == 0:
                if deriv:
                    nxt = Fc[:, j + 1] if j < CH - 1 else peek
                    r = T["sk_mid"] * (nxt - fm1)
                    fm2, fm1 = fm1, Fc[:, j]
                else:
                    r = T["sk_mid"] * Fc[:, j]
                eprev = -T["l_mid"] * eprev + r
            else:
                if deriv:
                    nxt = Fc[:, j + 1] if j < CH - 1 else peek
                    if j == jlast and not hi_closure:
                        nxt = hval
                    r = tab["sk"][j] * (nxt - fm1)
                    if mode == 1 and j == 0 and lo_closure:
                        r = s0c * (-5. * Fc[:, 0] + 4. * Fc[:, 1] + Fc[:, 2])
                    if j == jlast and hi_closure:
                        r = snc * (5. * Fc[:, j] - 4. * fm1 - fm2)
                    if j == jlast and not hi_closure:
                        r = r - snb * bval
                    fm2, fm1 = fm1, Fc[:, j]
                else:
                    r = tab["sk"][j] * Fc[:, j]
                eprev = -tab["l"][j] * eprev + r
            e[:, j] = eprev
        state["eprev"], state["fm1"], state["fm2"] = eprev, fm1, fm2
        return e

    def bwd(e, x, mode, k_out):
        tab = T["head"] if mode == 1 else T["tail"]
        if mode == 0 and k_out is None:
            # warm-up sweep (kernels.cuh bwd_chunk<0, false>): only the end value is used, evaluated as four Horner
            # chains in g^4 -- x_out = sum_j (-g)^j e_j + (-g)^32 x_in
            ng = -T["g_mid"]
            g4 = (ng * ng) * (ng * ng)
            c = [g4 * x + e[:, CH - 4], e[:, CH - 3].copy(), e[:, CH - 2].copy(), e[:, CH - 1].copy()]
            for j in range(CH - 8, -1, -4):
                for m in range(4):
                    c[m] = g4 * c[m] + e[:, j + m]
            return ng * (ng * (ng * c[3] + c[2]) + c[1]) + c[0]
        for j in range(CH - 1, -1, -1):
            g = T["g_mid"] if mode == 0 else tab["g"][j]
            x = -g * x + e[:, j]
            if k_out is not None:
                out[:, k_out * CH + j] = x
        return x

    kseg = K if not kseg else min(int(kseg), K)
    nseg = (K + kseg - 1) // kseg
    for seg in range(nseg):                       # work items of one bundle (kernel: item_range)
        c0 = seg * kseg
        c1 = min(c0 + kseg, K)
        kbeg, kend, kout = max(c0 - 1, 0), min(c1, K - 1), c0
        reset_state()
        eA = None
        for k in range(kbeg, kend + 1):
            eB = fwd(k, kend)
            last = (k == K - 1)
            x = np.zeros(nl)
            if last:
                x = bwd(eB, x, 1 if k == 0 else 2, k)
                if k > kbeg and k - 1 >= kout:
                    x = bwd(eA, x, 1 if k == 1 else 0, k - 1)
            elif k > kbeg:
                x = bwd(eB, x, 0, None)
                if k - 1 >= kout:
                    x = bwd(eA, x, 1 if k == 1 else 0, k - 1)
            eA = eB
    return out[:, :n]


def edge_faces(F, coeffs, h, lo_closure, hi_closure, halo_lo, halo_hi):
    """Emulation of edge_faces_kernel: interface values from the first 32 and the last 32 rows of the block."""
    F = np.asarray(F, dtype=np.float64)
    nl, n = F.shape
    assert n >= 2 * CH + 2
    T = tables(n, coeffs, 3. / (4 * h))
    lo_face, hi_face = np.zeros(nl), np.zeros(nl)
    if not lo_closure:
        fm1 = np.asarray(halo_lo, dtype=np.float64)
        eprev = np.zeros(nl)
        e = np.zeros((nl, CH - 1))
        for j in range(CH - 1):
            eprev = -T["head"]["l"][j] * eprev + T["head"]["sk"][j] * (F[:, j + 1] - fm1)
            e[:, j] = eprev
            fm1 = F[:, j]
        x = np.zeros(nl)
        for j in range(CH - 2, -1, -1):
            x = -T["head"]["g"][j] * x + e[:, j]
        lo_face = -x
    if not hi_closure:
        Ft = F[:, n - CH:]
        eprev = np.zeros(nl)
        for j in range(1, CH - 1):
            eprev = -T["l_mid"] * eprev + T["sk_mid"] * (Ft[:, j + 1] - Ft[:, j - 1])
        jl = T["jl"]
        eprev = -T["tail"]["l"][jl] * eprev + T["tail"]["sk"][jl] * (np.asarray(halo_hi) - Ft[:, CH - 2])
        hi_face = -eprev
    return lo_face, hi_face
