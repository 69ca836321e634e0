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


def stream_lines(F, coeffs, h=None, lo_closure=True, hi_closure=True, halo_lo=None, halo_hi=None):
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
    eprev = np.zeros(nl)
    fm1 = np.zeros(nl) if (lo_closure or not deriv) else np.asarray(halo_lo, dtype=np.float64).copy()
    fm2 = np.zeros(nl)
    hval = np.zeros(nl) if (hi_closure or not deriv) else np.asarray(halo_hi, dtype=np.float64)
    eA = None

    def fwd(k):
        nonlocal eprev, fm1, fm2
        Fc = Fp[:, k * CH:(k + 1) * CH]
        last = (k == K - 1)
        peek = np.zeros(nl) if last else Fp[:, (k + 1) * CH]
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
                    fm2, fm1 = fm1, Fc[:, j]
                else:
                    r = tab["sk"][j] * Fc[:, j]
                eprev = -tab["l"][j] * eprev + r
            e[:, j] = eprev
        return e

    def bwd(e, x, mode, k_out):
        tab = T["head"] if mode == 1 else T["tail"]
        for j in range(CH - 1, -1, -1):
            g = T["g_mid"] if mode == 0 else tab["g"][j]
            x = -g * x + e[:, j]
            if k_out is not None:
                out[:, k_out * CH + j] = x
        return x

    for k in range(K):
        eB = fwd(k)
        last = (k == K - 1)
        x = np.zeros(nl)
        if last:
            x = bwd(eB, x, 1 if k == 0 else 2, k)
            if k > 0:
                x = bwd(eA, x, 1 if k == 1 else 0, k - 1)
        elif k > 0:
            x = bwd(eB, x, 0, None)
            x = bwd(eA, x, 1 if k == 1 else 0, k - 1)
        eA = eB
    return out[:, :n]
