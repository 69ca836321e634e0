// kernels_zx.cuh -- the partitioned d/dz of a z-slab in ONE kernel: edge faces, neighbour exchange over NVLink, reduced
// system and coupled solve, bundle by bundle.
//
// The multi-rank derivative needs, per line, the two interface unknowns alpha / beta, which depend on the first and last
// 32 rows of this slab AND of the two neighbour slabs (reference: halo exchange, local solve, Gather of the interface
// faces, reduced solve on the line root, Scatter, sumSolutions -- code/cuda/compact.py:29-126).  Round 1 and the first
// half of round 2 computed the faces in a pass of their own (edge_faces_kernel, then edge items inside the x/y kernel):
// 64 of the slab's planes read a second time from HBM -- 0.53 GB and 0.09 ms per step whatever the slab thickness, 9 %
// of the 8-GPU step -- plus a reduce launch and two alpha / beta planes.  Here a warp does, for the bundles it draws,
//
//     A(b2)  B(b1)  A(b3)  B(b2)  A(b4)  B(b3) ...
//
//   A(b): the head tile (rows 0..31) and the tail tile (rows n-32..n-1) of bundle b through the TMA ring; the two
//         interface faces with guessed neighbour points (arithmetic of edge_faces_kernel, defer form); the faces and the
//         slab's first / last row are stored straight into the z-neighbours' receive arrays (NVLink) as self-validating
//         words -- no fence, no flag;
//   B(b): polls the four values of every line that the neighbours' A(b) stored here, solves the neighbour-only reduced
//         system for alpha / beta in registers, and runs the coupled one-pass solve of bundle b (kernels.cuh chunk
//         primitives, results staged in the ring as in stream_kernel_xy).  The two tiles A(b) read are read again one
//         bundle-time (~8 us) later: a ~10 MB working set for the whole GPU, served by L2.
// A never waits, and B(b) only waits for work the neighbours' A phases do without waiting, in the same draw order: no
// deadlock (the rank that has drawn the fewest bundles always finds its data).  One bundle of lag hides the NVLink
// latency.  DRAM traffic of the partitioned d/dz = the unpartitioned one's; launches per step and rank: the x/y launch
// and this one.
//
// Exchange words ("LL": NCCL's low-latency idea for fp64): a double travels as two 64-bit words, each = 32 data bits |
// 32-bit call number, written by one 16-byte vector store.  Whatever the fabric does to that store, an 8-byte word is
// atomic, so a receiver that sees the current call number in BOTH words has the value of this call -- stale data of the
// call two steps back (same parity buffer) carries another number.
#pragma once
#include "kernels.cuh"

namespace cfd {

struct ZXParams {
    // receive arrays of this rank (local memory, written by the neighbours): [plane] ulonglong2 each
    const ulonglong2 *in_face_lo, *in_halo_lo;       // from the left neighbour: its hi face, its last row
    const ulonglong2 *in_face_hi, *in_halo_hi;       // from the right neighbour: its lo face, its first row
    // the neighbours' receive arrays (peer addresses)
    ulonglong2 *out_face_lo, *out_halo_lo;           // left neighbour's "from the right" arrays: our lo face, our first row
    ulonglong2 *out_face_hi, *out_halo_hi;           // right neighbour's "from the left" arrays: our hi face, our last row
    unsigned int tag;                                // call number (low 32 bits), never 0
    int pv, own;                                     // neighbour-only reduced system: virtual ranks, own index
    double w_lo, w_hi;                               // d(lo face) / d f[-1], d(hi face) / d f[n]
    double sk_last, l_last;
    double lu[36];                                   // elimination table [6][2 pv]
    WaitP wait;
    int hints;                                       // bit 0: evict_last on phase-A tiles, bit 1: evict_first on the rest
                                                     // of the loads, bit 2: evict_first on the result stores
};

__device__ __forceinline__ void ll_store(ulonglong2 *p, double v, unsigned int tag)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long w0 = (bits & 0xffffffffULL) | ((unsigned long long)tag << 32);
    const unsigned long long w1 = (bits >> 32) | ((unsigned long long)tag << 32);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}

__device__ __forceinline__ bool ll_try_load(const ulonglong2 *p, unsigned int tag, double &v)
{
    unsigned long long w0, w1;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
    if ((unsigned int)(w0 >> 32) != tag || (unsigned int)(w1 >> 32) != tag) return false;
    v = __longlong_as_double((long long)((w0 & 0xffffffffULL) | (w1 << 32)));
    return true;
}

__device__ __forceinline__ ulonglong2 ll_load_raw(const ulonglong2 *p)
{
    ulonglong2 w;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w.x), "=l"(w.y) : "l"(p) : "memory");
    return w;
}

__device__ __forceinline__ bool ll_decode(const ulonglong2 &w, unsigned int tag, double &v)
{
    if ((unsigned int)(w.x >> 32) != tag || (unsigned int)(w.y >> 32) != tag) return false;
    v = __longlong_as_double((long long)((w.x & 0xffffffffULL) | (w.y << 32)));
    return true;
}

// what phase A leaves for phase B of the same bundle
struct ZXPending {
    double lo_face, hi_face, row0, rowl;
    long b;
};

// The neighbours' words of a bundle, requested one phase early (before the phase A that precedes its phase B) so that
// their global-load latency is not exposed in front of the forward recurrence; zx_interface polls only what was not
// there yet.
struct ZXWords {
    ulonglong2 w[4];
    long b;                              // bundle they belong to, -1 = none requested
};

__device__ __forceinline__ void zx_request(const KParams &p, const ZXParams &z, long b, int lane, ZXWords &q)
{
    const long col = (b % p.inner_tiles) * CH + lane;
    q.b = b;
    if (col < p.inner) {
        if (!p.lo_closure) { q.w[0] = ll_load_raw(z.in_face_lo + col); q.w[1] = ll_load_raw(z.in_halo_lo + col); }
        if (!p.hi_closure) { q.w[2] = ll_load_raw(z.in_face_hi + col); q.w[3] = ll_load_raw(z.in_halo_hi + col); }
    }
}

template <int NS, class Issue>
__device__ __forceinline__ void zx_phase_a(const KParams &p, const ZXParams &z, long b, unsigned char *wbase, uint32_t bar0,
                                           int lane, int &slot, uint32_t &phase, bool &first_step, Issue &issue,
                                           ZXPending &out)
{
    const long col = (b % p.inner_tiles) * CH + lane;       // outer == 1 for z lines: the bundle is 32 columns of the plane
    const bool ok = col < p.inner;
    auto take_tile = [&](double (&F)[CH]) {
        mbar_wait(bar0 + 8 * slot, phase);
        load_chunk<false>(wbase + slot * SLOT_BYTES, lane, F);
        __syncwarp();
        if (lane == 0 && !first_step) {      // ring discipline of stream_kernel_xy: refill the previous step's slot
            tma_wait_read0();
            issue();
        }
        first_step = false;
        __syncwarp();
        if (++slot == NS) { slot = 0; phase ^= 1u; }
    };
    out.lo_face = out.hi_face = out.row0 = out.rowl = 0.0;
    out.b = b;
    if (!p.lo_closure) {
        double F[CH], e[CH - 1];
        take_tile(F);
        double fm1 = F[0], eprev = 0.0;      // guess f[-1] := f[0]
#pragma unroll
        for (int j = 0; j < CH - 1; j++) {
            eprev = fma(-p.head.l[j], eprev, p.head.sk[j] * (F[j + 1] - fm1));
            e[j] = eprev;
            fm1 = F[j];
        }
        double x = 0.0;
#pragma unroll
        for (int j = CH - 2; j >= 0; j--) x = fma(-p.head.g[j], x, e[j]);
        out.lo_face = -x;
        out.row0 = F[0];
        if (ok) {
            ll_store(z.out_face_lo + col, out.lo_face, z.tag);
            ll_store(z.out_halo_lo + col, out.row0, z.tag);
        }
    }
    if (!p.hi_closure) {
        double F[CH];                        // rows n-32 .. n-1
        take_tile(F);
        double eprev = 0.0;
#pragma unroll
        for (int j = 1; j < CH - 1; j++)                                    // rows n-31 .. n-2 from a zero state
            eprev = fma(-p.l_mid, eprev, p.sk_mid * (F[j + 1] - F[j - 1]));
        eprev = fma(-z.l_last, eprev, z.sk_last * (F[CH - 1] - F[CH - 2])); // row n-1, guess f[n] := f[n-1]
        out.hi_face = -eprev;
        out.rowl = F[CH - 1];
        if (ok) {
            ll_store(z.out_face_hi + col, out.hi_face, z.tag);
            ll_store(z.out_halo_hi + col, out.rowl, z.tag);
        }
    }
}

// Phase B, first half: wait for the neighbours' words of this bundle, fold the halo terms into the faces, solve the
// reduced system.  Returns halo_lo, halo_hi, alpha, beta of this lane's line.
__device__ __forceinline__ void zx_interface(const KParams &p, const ZXParams &z, long b, int lane, const ZXPending &pd,
                                             const ZXWords &pre, double &halo_lo, double &halo_hi, double &alpha, double &beta)
{
    const long col = (b % p.inner_tiles) * CH + lane;
    const bool ok = col < p.inner;
    double nb_face_lo = 0.0, nb_face_hi = 0.0;
    halo_lo = halo_hi = 0.0;
    if (ok) {
        bool g0 = p.lo_closure, g1 = p.lo_closure, g2 = p.hi_closure, g3 = p.hi_closure;
        if (pre.b == b) {                                    // what the early request already brought
            if (!g0) g0 = ll_decode(pre.w[0], z.tag, nb_face_lo);
            if (!g1) g1 = ll_decode(pre.w[1], z.tag, halo_lo);
            if (!g2) g2 = ll_decode(pre.w[2], z.tag, nb_face_hi);
            if (!g3) g3 = ll_decode(pre.w[3], z.tag, halo_hi);
        }
        unsigned long long t0 = 0;
        unsigned ns = 32;
        while (!(g0 && g1 && g2 && g3)) {
            if (!g0) g0 = ll_try_load(z.in_face_lo + col, z.tag, nb_face_lo);
            if (!g1) g1 = ll_try_load(z.in_halo_lo + col, z.tag, halo_lo);
            if (!g2) g2 = ll_try_load(z.in_face_hi + col, z.tag, nb_face_hi);
            if (!g3) g3 = ll_try_load(z.in_halo_hi + col, z.tag, halo_hi);
            if (g0 && g1 && g2 && g3) break;
            if (t0 == 0) { t0 = global_timer_ns(); continue; }      // first miss: look again at once
            __nanosleep(ns);
            if (ns < 2048) ns <<= 1;
            if (global_timer_ns() - t0 > z.wait.timeout_ns) {
                if (z.wait.err) { *(volatile int *)z.wait.err = -4; __threadfence_system(); }
                break;
            }
        }
    }
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int o = z.own;
    v[2 * o] = pd.lo_face;
    v[2 * o + 1] = pd.hi_face;
    if (!p.lo_closure) {                                     // there is a left neighbour
        const double d = halo_lo - pd.row0;
        v[2 * o] += z.w_lo * d;
        v[2 * o - 1] = nb_face_lo - z.w_hi * d;
    }
    if (!p.hi_closure) {                                     // there is a right neighbour
        const double d = halo_hi - pd.rowl;
        v[2 * o + 1] += z.w_hi * d;
        v[2 * o + 2] = nb_face_hi - z.w_lo * d;
    }
    reduced_unknowns<false>(v, z.lu, 1, 0, z.pv, z.own, alpha, beta);
}

// Phase B, second half: the coupled one-pass solve of bundle b (whole line, strided layout), results staged in the
// ring slot just consumed -- stream_kernel_xy's xy_run_item with the interface data as initial state.
template <int NS, class Issue>
__device__ __forceinline__ void zx_phase_b(const KParams &p, const CUtensorMap *tm_out, long b, unsigned char *wbase,
                                           uint32_t bar0, int lane, int &slot, uint32_t &phase, bool &first_step, Issue &issue,
                                           double halo_lo, double halo_hi, double alpha, double beta, bool store_hint,
                                           unsigned long long pol_first)
{
    const int K = p.K;
    const int oc0 = (int)(b % p.inner_tiles) * CH, oc2 = (int)(b / p.inner_tiles);
    double eA[CH], eB[CH], F[CH];
    double eprev = alpha, fm1 = halo_lo, fm2 = 0.0;          // row 0 sees x_{-1} = alpha through l_0 = a_i beta_0
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const bool last = (k == K - 1);
        unsigned char *cur = wbase + slot * SLOT_BYTES;
        mbar_wait(bar0 + 8 * slot, phase);
        load_chunk<false>(cur, lane, F);
        double peek = 0.0;
        if (!last) {
            const int s1 = (slot + 1 == NS) ? 0 : slot + 1;
            const uint32_t ph1 = (slot + 1 == NS) ? (phase ^ 1u) : phase;
            mbar_wait(bar0 + 8 * s1, ph1);
            peek = load_first<false>(wbase + s1 * SLOT_BYTES, lane);
        }
        if (k == 0) {
            fwd_chunk<1, true, -1>(p, F, peek, halo_hi, beta, eB, eprev, fm1, fm2);      // K >= 3: never the last chunk
        } else if (last) {
            if (p.jl == CH - 1) fwd_chunk<2, true, CH - 1>(p, F, peek, halo_hi, beta, eB, eprev, fm1, fm2);
            else                fwd_chunk<2, true, -2>(p, F, peek, halo_hi, beta, eB, eprev, fm1, fm2);
        } else {
            fwd_chunk<0, true, -1>(p, F, peek, halo_hi, beta, eB, eprev, fm1, fm2);
        }
        __syncwarp();
        if (lane == 0 && !first_step) {
            tma_wait_read0();
            issue();
        }
        first_step = false;
        __syncwarp();
        auto flush = [&](int kc) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (store_hint) tma_store_3d_hint(tm_out, smem_u32(cur), oc0, kc * CH, oc2, pol_first);
                else            tma_store_3d(tm_out, smem_u32(cur), oc0, kc * CH, oc2);
                tma_commit();
            }
        };
        double x = 0.0;
        if (last) {
            bwd_chunk<2, true, false>(p, eB, x, cur, lane);
            flush(k);
            if (lane == 0) tma_wait_read0();      // the line's last two tiles share the slot
            __syncwarp();
            if (k == 1) bwd_chunk<1, true, false>(p, eA, x, cur, lane);
            else        bwd_chunk<0, true, false>(p, eA, x, cur, lane);
            flush(k - 1);
        } else if (k > 0) {
            bwd_chunk<0, false, false>(p, eB, x, cur, lane);
            if (k == 1) bwd_chunk<1, true, false>(p, eA, x, cur, lane);
            else        bwd_chunk<0, true, false>(p, eA, x, cur, lane);
            flush(k - 1);
        }
#pragma unroll
        for (int j = 0; j < CH; j++) eA[j] = eB[j];
        if (++slot == NS) { slot = 0; phase ^= 1u; }
    }
}

template <int NS>
__global__ void __launch_bounds__(224, 1)
stream_kernel_zx(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                 const __grid_constant__ KParams p, const __grid_constant__ ZXParams z)
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
    const int na = (p.lo_closure ? 0 : 1) + (p.hi_closure ? 0 : 1);      // tiles of a phase A (1 or 2)

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        fence_mbar_init();
    }
    __syncwarp();
    pdl_launch_dependents();
    pdl_wait();

    // ---- producer (lane 0).  Tile sequence of the warp: A(b1) A(b2) B(b1) A(b3) B(b2) ... A(bm) B(bm-1) B(bm): a bundle
    // is drawn when its phase A starts; at most two bundles wait between their A and their B (FIFO q0, q1).
    // tag of a phase's first tile = (bundle << 1) | kind (0 = A, 1 = B); -2 inside a phase; -1 = no more work.
    const unsigned long long pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
    long q0 = -1, q1 = -1;              // bundles that have had their A, oldest first
    int qn = 0;
    long cb = -1;                       // bundle of the phase being issued
    int ct = 0;                         // its next tile
    int mode = 0;                       // 0: decide, 1: issuing a phase A, 2: issuing a phase B, 3: end
    bool dry = false;
    int islot = 0;
    auto issue = [&]() {
        if (mode == 0) {
            if (!dry && qn < 2) {
                const long w = (long)atomicAdd(p.counter, 1ULL);
                if (w < p.nb) { cb = w; ct = 0; mode = 1; }
                else dry = true;
            }
            if (mode == 0) {
                if (qn > 0) { cb = q0; ct = 0; mode = 2; }       // (here qn == 2, or the draw ran dry)
                else mode = 3;
            }
        }
        if (mode == 3) {
            tag[islot] = -1;
        } else {
            const uint32_t bar = bar0 + 8 * islot;
            const uint32_t dst = smem_u32(wbase + islot * SLOT_BYTES);
            const int c0 = (int)(cb % p.inner_tiles) * CH, c2 = (int)(cb / p.inner_tiles);
            mbar_expect_tx(bar, SLOT_BYTES);
            if (mode == 1) {
                tag[islot] = (ct == 0) ? (cb << 1) : -2;
                const int row = (!p.lo_closure && ct == 0) ? 0 : p.n - CH;      // head tile, then tail tile
                // these two tiles come back in phase B, one bundle-time later: ask L2 to hold on to them
                if (z.hints & 1) tma_load_3d_hint(dst, &tm_in, bar, c0, row, c2, pol_last);
                else             tma_load_3d(dst, &tm_in, bar, c0, row, c2);
                if (++ct == na) {
                    if (qn == 0) q0 = cb; else q1 = cb;
                    ++qn;
                    mode = 0;
                }
            } else {
                tag[islot] = (ct == 0) ? ((cb << 1) | 1) : -2;
                if (z.hints & 2) tma_load_3d_hint(dst, &tm_in, bar, c0, ct * CH, c2, pol_first);     // last use of every tile
                else             tma_load_3d(dst, &tm_in, bar, c0, ct * CH, c2);
                if (++ct == K) {
                    q0 = q1; q1 = -1;
                    --qn;
                    mode = 0;
                }
            }
        }
        if (++islot == NS) islot = 0;
    };
    if (lane == 0) {
#pragma unroll 1
        for (int s = 0; s < NS; s++) issue();
    }
    __syncwarp();

    // ---- consumer: the same sequence, read off the tags; what a phase A leaves waits in a two-deep FIFO for its phase B
    int slot = 0;
    uint32_t phase = 0;
    bool first_step = true;
    ZXPending pend0 = {0.0, 0.0, 0.0, 0.0, -1}, pend1 = {0.0, 0.0, 0.0, 0.0, -1};
    ZXWords pre;
    pre.b = -1;
    int npend = 0;
    for (;;) {
        const long w = tag[slot];
        if (w < 0) break;                                    // (-1; a phase never starts on a -2 tile)
        const long b = w >> 1;
        if ((w & 1) == 0) {
            if (npend == 0) {
                zx_phase_a<NS>(p, z, b, wbase, bar0, lane, slot, phase, first_step, issue, pend0);
            } else {
                zx_request(p, z, pend0.b, lane, pre);        // the phase B that follows this phase A: ask for its words now
                zx_phase_a<NS>(p, z, b, wbase, bar0, lane, slot, phase, first_step, issue, pend1);
            }
            ++npend;
        } else {
            double halo_lo, halo_hi, alpha, beta;
            zx_interface(p, z, b, lane, pend0, pre, halo_lo, halo_hi, alpha, beta);
            pend0 = pend1;
            --npend;
            zx_phase_b<NS>(p, &tm_out, b, wbase, bar0, lane, slot, phase, first_step, issue, halo_lo, halo_hi, alpha, beta,
                           (z.hints & 4) != 0, pol_first);
        }
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
