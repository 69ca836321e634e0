// kernels_general.cuh -- the one-pass streaming solve for everything the Pade-4 fast path (kernels.cuh) does not cover:
//
//   * WIDER LOOK-AHEAD.  The backward sweep of a chunk is started LA chunks to the right; the neglected coupling is
//     |g|^(32 LA).  LA = 1 needs |g| <= 0.316 (the reference's Pade matrix: 0.268); LA = 2 serves |g| <= 0.563, e.g. the
//     6th-order tridiagonal scheme (alpha = 1/3, |g| = 0.382: 0.382^64 = 2e-27) -- still read-once / write-once, where
//     round 1 fell back to the exact two-pass kernel (32 B / unknown).  The look-ahead width is derived from the
//     coefficients on the host (api.cu lookahead_chunks).
//   * GENERAL RIGHT-HAND SIDES.  Interior: STENCIL = 1 (first derivatives) r_i = c1 (f_{i+1} - f_{i-1}) + c2 (f_{i+2} - f_{i-2});
//     STENCIL = 2 (second derivatives) r_i = c1 ((f_{i+1} - f_i) + (f_{i-1} - f_i)) + c2 ((f_{i+2} - f_i) + (f_{i-2} - f_i)).
//     Up to two closure rows per end with 4-point one-sided stencils, r = sum_k q_k (f_k - f_0): every derivative
//     stencil annihilates constants, and written over DIFFERENCES of neighbouring points it loses eps |h f'| / h^m to
//     round-off instead of eps |f| / h^m (two digits at 256^3 for the second derivative).  The matrix may differ from
//     the Toeplitz interior in its first two and last two rows.  Serves the 6th-order first derivative and the
//     4th-order second derivative; the reference solver's general [b1, c1, ai, bi, ci, an, bn] matrices are the
//     STENCIL = 0 case (right-hand side = the tile itself).
//
// Same warp-autonomous TMA ring, dynamic bundle draw and TMA stores as stream_kernel; whole lines only (no
// segmentation).  Where the forward values of the waiting chunks live is described at the kernel.
#pragma once
#include "kernels.cuh"

namespace cfd {

struct GParams {
    int n, K, jl;
    int inner, inner_tiles;
    long nb, rows;
    unsigned long long *counter;
    // matrix: per-row tables for the first chunk and the last TWO chunks (a special row n-2 can fall into chunk K-2),
    // constants in between.  sk multiplies the right-hand side (beta_i, times the caller's scale for STENCIL = false).
    double sk_mid, l_mid, g_mid;
    RowTab head, tail2, tail;
    // right-hand side stencil (STENCIL != 0)
    double c1, c2;
    int nspecial;                 // closure rows per end: 1 or 2
    double q[2][4];               // row 0, row 1:     r = sum_{k>=1} q[.][k] (f[k] - f[0])          (q[.][0] = -sum of the rest)
    double p[2][4];               // row n-1, row n-2: r = sum_{k>=1} p[.][k] (f[n-1-k] - f[n-1])
};

template <bool CONTIG>
__device__ __forceinline__ void store_chunk(unsigned char *slot, int lane, const double (&E)[CH])
{
    if constexpr (CONTIG) {
        unsigned char *row = slot + lane * 128;
        const int sw = (lane & 7) << 4;
#pragma unroll
        for (int m = 0; m < 16; m++)
            *reinterpret_cast<double2 *>(row + (m >> 3) * 4096 + (((m & 7) << 4) ^ sw)) = make_double2(E[2 * m], E[2 * m + 1]);
    } else {
        double *col = reinterpret_cast<double *>(slot) + lane;
#pragma unroll
        for (int j = 0; j < CH; j++) col[j * CH] = E[j];
    }
}

template <bool CONTIG>
__device__ __forceinline__ double load_row(const unsigned char *slot, int lane, int r)   // r = 0 or 1
{
    if constexpr (CONTIG) return *reinterpret_cast<const double *>(slot + lane * 128 + ((lane & 7) << 4) + 8 * r);
    else return reinterpret_cast<const double *>(slot)[r * CH + lane];
}

// Forward elimination of chunk k.  TAB: 0 = constants, 1 = table chunk (head / tail2 / tail; rows located at run time).
// History: h1, h2, h3 = f[i-1], f[i-2], f[i-3] on entry; pk0, pk1 = first two rows of the next tile (0 past the end).
template <int STENCIL, int TAB>
__device__ __forceinline__ void gfwd_chunk(const GParams &p, const RowTab *T, int k, const double (&F)[CH], double pk0,
                                           double pk1, double (&e)[CH], double &eprev, double &h1, double &h2, double &h3)
{
    const int n = p.n;
    const int row0 = k * CH;
#pragma unroll
    for (int j = 0; j < CH; j++) {
        double r;
        if constexpr (STENCIL != 0) {
            const double a1 = (j >= 1) ? F[j - 1] : h1, a2 = (j >= 2) ? F[j - 2] : (j == 1 ? h1 : h2);
            const double b1 = (j < CH - 1) ? F[j + 1] : pk0, b2 = (j < CH - 2) ? F[j + 2] : (j == CH - 2 ? pk0 : pk1);
            if constexpr (STENCIL == 1) r = fma(p.c2, b2 - a2, p.c1 * (b1 - a1));
            else                        r = fma(p.c2, (b2 - F[j]) + (a2 - F[j]), p.c1 * ((b1 - F[j]) + (a1 - F[j])));
            if constexpr (TAB == 1) {
                const int row = row0 + j;
                if (k == 0 && j < 2 && j < p.nspecial) {          // closure rows 0 (, 1): forward one-sided, inside chunk 0
                    r = fma(p.q[j][3], F[3] - F[0], fma(p.q[j][2], F[2] - F[0], p.q[j][1] * (F[1] - F[0])));
                }
                const int back = n - 1 - row;                     // 0: row n-1, 1: row n-2
                if (back >= 0 && back < p.nspecial) {
                    const double a3 = (j >= 3) ? F[j - 3] : (j == 2 ? h1 : (j == 1 ? h2 : h3));
                    // differences from f[n-1]: this row (back = 0) or the next one (back = 1)
                    if (back == 0) r = fma(p.p[0][3], a3 - F[j], fma(p.p[0][2], a2 - F[j], p.p[0][1] * (a1 - F[j])));
                    else           r = fma(p.p[1][3], a2 - b1, fma(p.p[1][2], a1 - b1, p.p[1][1] * (F[j] - b1)));
                }
            }
        } else {
            r = F[j];
        }
        if constexpr (TAB == 1) eprev = fma(-T->l[j], eprev, T->sk[j] * r);
        else                    eprev = fma(-p.l_mid, eprev, p.sk_mid * r);
        e[j] = eprev;
    }
    if constexpr (STENCIL != 0) { h3 = F[CH - 3]; h2 = F[CH - 2]; h1 = F[CH - 1]; }
}

// Backward sweep over one chunk held in registers; OUT writes x over the values of `slot` (tile layout).
template <bool CONTIG, int TAB, bool OUT>
__device__ __forceinline__ void gbwd_chunk(const GParams &p, const RowTab *T, const double (&e)[CH], double &x,
                                           unsigned char *slot, int lane)
{
    double X[CH];
#pragma unroll
    for (int j = CH - 1; j >= 0; j--) {
        const double ng = (TAB == 1) ? -T->g[j] : -p.g_mid;
        x = fma(ng, x, e[j]);
        X[j] = x;
    }
    if constexpr (OUT) store_chunk<CONTIG>(slot, lane, X);
}

// Register / shared-memory budget (second version; the first kept every waiting chunk's e-values in shared memory,
// 40-48 KiB per warp, 4 warps per SM: latency-bound on the recurrence chains, issue slots 16-18 % busy,
// profiles/r2i_ncu_full_raw_general_kernel_512_first_version.csv).  Now the two newest chunks' e-values stay in
// registers (eB = chunk k, eA = chunk k-1, as in stream_kernel) and shared memory is the NS-slot tile ring and nothing
// else, as in stream_kernel_xy: the slot a tile was consumed from doubles as
//   LA = 1: the staging slot of result chunk k-1, shipped in the same step, refilled one step later;
//   LA = 2: the parking slot of e(k-1), which waits one step there for its backward sweep (chunk k-2 is swept, in
//           place, in the slot parked the step before), shipped then, refilled two steps after it was consumed.
// 32 KiB per warp with NS = 4 -> 6 warps per SM.
template <bool CONTIG, int STENCIL, int LA, int NS>
__global__ void __launch_bounds__(224, 1)
stream_kernel_g(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                const __grid_constant__ GParams p)
{
    extern __shared__ unsigned char smem_raw[];
    constexpr int PER_WARP = NS * SLOT_BYTES;
    constexpr int CTRL = NS * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *wbase = base + warp * PER_WARP;
    unsigned char *ctrl = base + nwarps * PER_WARP + warp * CTRL;
    const uint32_t bar0 = smem_u32(ctrl);
    volatile long long *tag = reinterpret_cast<volatile long long *>(ctrl + NS * 8);
    const int K = p.K;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        fence_mbar_init();
    }
    __syncwarp();
    pdl_launch_dependents();
    pdl_wait();

    long ib = 0;
    int ik = 0, islot = 0;
    bool dry = false;
    auto issue = [&]() {
        if (ik == 0 && !dry) {
            ib = (long)atomicAdd(p.counter, 1ULL);
            dry = ib >= p.nb;
        }
        if (dry) {
            tag[islot] = -1;
        } else {
            tag[islot] = ib;
            const uint32_t bar = bar0 + 8 * islot;
            const uint32_t dst = smem_u32(wbase + islot * SLOT_BYTES);
            mbar_expect_tx(bar, SLOT_BYTES);
            if constexpr (CONTIG) {
                tma_load_2d(dst, &tm_in, bar, ik * CH, (int)(ib * CH));
                tma_load_2d(dst + 4096, &tm_in, bar, ik * CH + 16, (int)(ib * CH));
            } else {
                tma_load_3d(dst, &tm_in, bar, (int)(ib % p.inner_tiles) * CH, ik * CH, (int)(ib / p.inner_tiles));
            }
            if (++ik == K) ik = 0;
        }
        if (++islot == NS) islot = 0;
    };
    if (lane == 0) {
#pragma unroll 1
        for (int s = 0; s < NS; s++) issue();
    }
    __syncwarp();

    auto table_of = [&](int c) -> const RowTab * {            // nullptr = constants
        if (c == 0) return &p.head;
        if (c == K - 1) return &p.tail;
        if (c == K - 2) return &p.tail2;
        return nullptr;
    };
    auto ship = [&](int c, long b, unsigned char *slot) {     // TMA store of result chunk c from `slot`
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            if constexpr (CONTIG) {
                tma_store_2d(&tm_out, smem_u32(slot), c * CH, (int)(b * CH));
                tma_store_2d(&tm_out, smem_u32(slot) + 4096, c * CH + 16, (int)(b * CH));
            } else {
                tma_store_3d(&tm_out, smem_u32(slot), (int)(b % p.inner_tiles) * CH, c * CH, (int)(b / p.inner_tiles));
            }
            tma_commit();
        }
    };
    // backward sweep over a chunk held in registers; `out`: results staged in `slot` and shipped
    auto sweep_regs = [&](const double (&E)[CH], int c, long b, double &x, bool out, unsigned char *slot) {
        const RowTab *T = table_of(c);
        if (out) {
            if (T) gbwd_chunk<CONTIG, 1, true>(p, T, E, x, slot, lane);
            else   gbwd_chunk<CONTIG, 0, true>(p, T, E, x, slot, lane);
            ship(c, b, slot);
        } else {
            if (T) gbwd_chunk<CONTIG, 1, false>(p, T, E, x, slot, lane);
            else   gbwd_chunk<CONTIG, 0, false>(p, T, E, x, slot, lane);
        }
    };

    double F[CH], eA[CH], eB[CH];
    double eprev = 0.0, h1 = 0.0, h2 = 0.0, h3 = 0.0;
    long b = 0;
    int k = 0, slot = 0;
    int skip = LA;                                // refills lag the consumer by LA steps (see above)
    unsigned char *parked = nullptr;              // LA = 2: the slot holding e(k-2)
    uint32_t phase = 0;
    for (;;) {
        if (k == 0) {
            b = tag[slot];
            if (b < 0) break;
            eprev = 0.0; h1 = 0.0; h2 = 0.0; h3 = 0.0;
        }
        const bool last = (k == K - 1);
        unsigned char *cur = wbase + slot * SLOT_BYTES;
        mbar_wait(bar0 + 8 * slot, phase);
        load_chunk<CONTIG>(cur, lane, F);
        double pk0 = 0.0, pk1 = 0.0;
        if constexpr (STENCIL != 0) {
            if (!last) {
                const int s1 = (slot + 1 == NS) ? 0 : slot + 1;
                const uint32_t ph1 = (slot + 1 == NS) ? (phase ^ 1u) : phase;
                mbar_wait(bar0 + 8 * s1, ph1);
                pk0 = load_row<CONTIG>(wbase + s1 * SLOT_BYTES, lane, 0);
                pk1 = load_row<CONTIG>(wbase + s1 * SLOT_BYTES, lane, 1);
            }
        }
        {
            const RowTab *T = table_of(k);
            if (T) gfwd_chunk<STENCIL, 1>(p, T, k, F, pk0, pk1, eB, eprev, h1, h2, h3);
            else   gfwd_chunk<STENCIL, 0>(p, T, k, F, pk0, pk1, eB, eprev, h1, h2, h3);
        }
        // refill the slot that was shipped in the previous step (consumed LA steps ago)
        __syncwarp();
        if (lane == 0) {
            if (skip == 0) { tma_wait_read0(); issue(); }
        }
        if (skip > 0) --skip;
        __syncwarp();

        double x = 0.0;
        if (last) {                                   // exact sweep from the true end of the line
            sweep_regs(eB, k, b, x, true, cur);
            if (k >= 1) {
                if (lane == 0) tma_wait_read0();      // chunks K-1 and K-2 share the staging slot
                __syncwarp();
                sweep_regs(eA, k - 1, b, x, true, cur);
            }
            if constexpr (LA == 2) {
                if (k >= 2) {
                    double E[CH];
                    load_chunk<CONTIG>(parked, lane, E);
                    sweep_regs(E, k - 2, b, x, true, parked);
                }
            }
        } else if constexpr (LA == 1) {
            if (k >= 1) {
                sweep_regs(eB, k, b, x, false, cur);  // warm-up through chunk k
                sweep_regs(eA, k - 1, b, x, true, cur);
            }
        } else {
            if (k >= 2) {
                sweep_regs(eB, k, b, x, false, cur);  // warm-up through chunks k and k-1
                sweep_regs(eA, k - 1, b, x, false, cur);
            }
            unsigned char *old = parked;
            if (k >= 1) {                             // e(k-1) waits one step in the slot tile k came from
                store_chunk<CONTIG>(cur, lane, eA);
                parked = cur;
            }
            if (k >= 2) {                             // chunk k-2: swept in place in the slot parked the step before
                double E[CH];
                load_chunk<CONTIG>(old, lane, E);
                sweep_regs(E, k - 2, b, x, true, old);
            }
        }
#pragma unroll
        for (int j = 0; j < CH; j++) eA[j] = eB[j];
        if (++k == K) k = 0;
        if (++slot == NS) { slot = 0; phase ^= 1u; }
    }
    if (lane == 0) {
        tma_wait_all0();
        __threadfence();
        const unsigned long long total = (unsigned long long)gridDim.x * nwarps;
        if (atomicAdd(p.counter + 1, 1ULL) == total - 1) {
            p.counter[0] = 0ULL;
            p.counter[1] = 0ULL;
            __threadfence();
        }
    }
}

}  // namespace cfd
