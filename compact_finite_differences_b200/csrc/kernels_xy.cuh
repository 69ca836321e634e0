// kernels_xy.cuh -- d/dx and d/dy of the same field in ONE launch, with the work interleaved plane by plane.
//
// Both derivatives are still the one-pass streaming solve of kernels.cuh (same chunk primitives, same tables), and
// each still writes its own 8 B/point.  What changes is the ORDER of the work: the item list is
//     plane 0: x-bundles (32 rows of the plane each), y-bundles (32 columns each); plane 1: ...; ...
// and warps draw items from it dynamically.  At any time the ~600 resident warps work on a window of ~20 planes,
// so every tile of f is fetched from HBM by whichever of its two readers comes first and served from L2 to the
// other a few microseconds later: 24 B/point of DRAM traffic for the two derivatives instead of 32.
// The reference computes the directions by separate passes (and host transposes, code/ocl/compact.py:41-61).
#pragma once
#include "kernels.cuh"

namespace cfd {

struct XYParams {
    long nitems;        // entries of the draw order
    long nxy;           // nz * (nxp + nyp): item ids below are x / y bundles, ids from here on are edge items
    int nxp, nyp;       // x-bundles (ny/32) and y-bundles (ceil(nx/32)) per plane
    unsigned long long *counter;
    const int *order;   // draw position -> (item << 3) | segment, item = plane * (nxp + nyp) + index in plane;
                        // segment 0 = the whole line, s >= 1 = output chunks [(s-1) kseg, s kseg); nullptr = identity
    int kseg, ksegy;    // chunks per segment of the x lines / of the y lines (sub-plane wavefronts, see xy_order; 0 = whole)
    // Start-up stagger (off when tau_ns == 0): a warp whose FIRST item sits in slot s of the draw order starts it
    // s * tau_ns late, so that the wavefront exists from the first generation of items on instead of all resident
    // warps starting together (worth 3 % on lines of >= 32 tiles, nothing on shorter ones).
    float tau_ns, slot_items;
    // SEG variants only (long lines, sub-plane wavefronts): L2 eviction hints -- bit 0: evict_last on the tile loads
    // (every tile has a second reader a few microseconds away), bit 1: evict_first on the result stores (never read)
    int hints;
};

// One item = all chunks of one bundle, x (CONTIG) or y (STRIDED).  Ring state is shared across items.
//
// Shared memory per warp is the NS-slot tile ring and nothing else: the slot a tile was just consumed from (into
// registers) doubles as the staging slot of the result tile this step emits, and is refilled ONE STEP LATER, once
// the TMA store has finished reading it.  24 KiB per warp instead of 40 puts 8 warps on an SM (two per scheduler):
// with one warp per scheduler every fixed-latency dependency of the recurrences is exposed and the pair of
// derivatives is latency-bound (0.825 / 0.641 / 0.553 ms with 3 / 4 / 5 warps, time x warps = const).
template <bool CONTIG, int NS, bool SEG, class Issue>
__device__ __forceinline__ void xy_run_item(const KParams &p, const CUtensorMap *tm_out, long b, int oc0, int oc2,
                                            unsigned char *wbase, uint32_t bar0, int lane, int &slot, uint32_t &phase,
                                            bool &first_step, Issue &issue, int kbeg_, int kend_, int kout_, int kstop_,
                                            int hints = 0)
{
    // Tiles kbeg .. kend of the line, results for chunks kout .. kstop-1: the whole line (0, K-1, 0, K), or a
    // segment with one warm-up chunk in front (forward sweep from a zero state, exact to
    // 0.268^32 like the backward look-ahead) and the look-ahead chunk behind -- the same cut as kernels.cuh makes.
    // SEG = false (whole lines only) folds the range to (0, K-1, 0, K) at compile time: the kernel of the common
    // case carries no segment state in its 255 registers.
    const int K = p.K;
    const int kbeg = SEG ? kbeg_ : 0, kend = SEG ? kend_ : K - 1, kout = SEG ? kout_ : 0, kstop = SEG ? kstop_ : K;
    double eA[CH], eB[CH], F[CH];
    double eprev = 0.0, fm1 = 0.0, fm2 = 0.0;
#pragma unroll 1
    for (int k = kbeg; k <= kend; ++k) {
        const bool last = (k == K - 1);
        unsigned char *cur = wbase + slot * SLOT_BYTES;
        mbar_wait(bar0 + 8 * slot, phase);
        load_chunk<CONTIG>(cur, lane, F);
        double peek = 0.0;
        if (k < kend) {                  // the item's next tile is in the ring: peek at its first row
            const int s1 = (slot + 1 == NS) ? 0 : slot + 1;
            const uint32_t ph1 = (slot + 1 == NS) ? (phase ^ 1u) : phase;
            mbar_wait(bar0 + 8 * s1, ph1);
            peek = load_first<CONTIG>(wbase + s1 * SLOT_BYTES, lane);
        }
        if (k == 0) {
            if (last) fwd_chunk<1, true, -2>(p, F, peek, 0.0, 0.0, eB, eprev, fm1, fm2);
            else      fwd_chunk<1, true, -1>(p, F, peek, 0.0, 0.0, eB, eprev, fm1, fm2);
        } else if (last) {
            if (p.jl == CH - 1) fwd_chunk<2, true, CH - 1>(p, F, peek, 0.0, 0.0, eB, eprev, fm1, fm2);
            else                fwd_chunk<2, true, -2>(p, F, peek, 0.0, 0.0, eB, eprev, fm1, fm2);
        } else {
            fwd_chunk<0, true, -1>(p, F, peek, 0.0, 0.0, eB, eprev, fm1, fm2);
        }
        // refill the slot of the PREVIOUS step: its result tile (if any) has had a forward sweep's time to leave
        __syncwarp();
        if (lane == 0 && !first_step) {
            tma_wait_read0();
            issue();
        }
        first_step = false;
        __syncwarp();

        auto flush = [&](int kc) {       // ship the result tile staged in the current slot
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bool hinted = false;
                if constexpr (SEG) {
                    if (hints & 2) {
                        const unsigned long long pol = l2_policy_evict_first();
                        if constexpr (CONTIG) {
                            tma_store_2d_hint(tm_out, smem_u32(cur), kc * CH, (int)(b * CH), pol);
                            tma_store_2d_hint(tm_out, smem_u32(cur) + 4096, kc * CH + 16, (int)(b * CH), pol);
                        } else {
                            tma_store_3d_hint(tm_out, smem_u32(cur), oc0, kc * CH, oc2, pol);
                        }
                        hinted = true;
                    }
                }
                if (!hinted) {
                    if constexpr (CONTIG) {
                        tma_store_2d(tm_out, smem_u32(cur), kc * CH, (int)(b * CH));
                        tma_store_2d(tm_out, smem_u32(cur) + 4096, kc * CH + 16, (int)(b * CH));
                    } else {
                        tma_store_3d(tm_out, smem_u32(cur), oc0, kc * CH, oc2);
                    }
                }
                tma_commit();
            }
        };
        double x = 0.0;
        if (last) {
            if (k >= kstop) {                         // the line's last chunk is only this segment's look-ahead
                bwd_chunk<2, false, CONTIG>(p, eB, x, cur, lane);
            } else {
                if (k == 0) bwd_chunk<1, true, CONTIG>(p, eB, x, cur, lane);
                else        bwd_chunk<2, true, CONTIG>(p, eB, x, cur, lane);
                flush(k);
            }
            if (k > kbeg && k - 1 >= kout) {
                if (lane == 0) tma_wait_read0();      // the line's last two tiles share the slot
                __syncwarp();
                if (k == 1) bwd_chunk<1, true, CONTIG>(p, eA, x, cur, lane);
                else        bwd_chunk<0, true, CONTIG>(p, eA, x, cur, lane);
                flush(k - 1);
            }
        } else if (k > kbeg) {
            bwd_chunk<0, false, CONTIG>(p, eB, x, cur, lane);
            if (k - 1 >= kout) {                      // (chunk kbeg of a later segment is warm-up only)
                if (k == 1) bwd_chunk<1, true, CONTIG>(p, eA, x, cur, lane);
                else        bwd_chunk<0, true, CONTIG>(p, eA, x, cur, lane);
                flush(k - 1);
            }
        }
#pragma unroll
        for (int j = 0; j < CH; j++) eA[j] = eB[j];
        if (++slot == NS) { slot = 0; phase ^= 1u; }
    }
}

// ------------------------------------------------------------------------------------------------
// Edge items (EDGE = true, cfd_zpart_apply_xyz): the interface-face pass of a z-partitioned d/dz -- the work of
// edge_faces_kernel in its one-launch ("defer") form -- as work items of this persistent kernel instead of a launch
// beside it (a separate launch serialises with a persistent kernel whatever the stream priority: DESIGN.md section 6).
// Item t = 32 consecutive columns of the flattened [ny*nx] plane: the head tile (planes 0..31) and the tail tile
// (planes n-32..n-1) arrive through the same TMA ring (z-direction tensor map); nothing else is read.
// The faces go to the local buffer and straight into the neighbours' (NVLink stores), this slab's first / last plane
// into the neighbours' halo slots; edge items are drawn FIRST, every warp counts the ones it finished and reports them
// when it turns to its first x / y item, and the warp that completes the count raises the neighbours' flags.
// Same arithmetic, in the same order, as edge_faces_kernel: the faces are bit-identical to cfd_edge_faces_push.
// ------------------------------------------------------------------------------------------------
struct EdgeX {
    long nedge;                              // edge items = ceil(ny*nx / 32); 0 = none
    long nlines;                             // ny * nx
    int n;                                   // planes of the slab (>= 66)
    int has_lo, has_hi;                      // a left / right neighbour exists
    double sk_mid, l_mid, sk_last, l_last;
    const double *f;
    double *faces;                           // local [2][plane]: -x_R[first], -x_R[last] (with guessed neighbour points)
    double *peer_face_lo, *peer_face_hi;     // the neighbours' slots for them
    double *push_lo, *push_hi;               // the neighbours' halo slots for our first / last plane
    unsigned long long *flag_lo, *flag_hi;   // the neighbours' arrival flags
    unsigned long long *done;                // finished edge items (zero on entry, zero on exit)
    unsigned long long seq;
    RowTab head;
};

template <int NS, class Issue>
__device__ __forceinline__ void xy_run_edge(const EdgeX &ex, long t, unsigned char *wbase, uint32_t bar0, int lane,
                                            int &slot, uint32_t &phase, bool &first_step, Issue &issue)
{
    const long col = t * CH + lane;
    const bool ok = col < ex.nlines;
    auto take_tile = [&](double (&F)[CH]) {
        mbar_wait(bar0 + 8 * slot, phase);
        load_chunk<false>(wbase + slot * SLOT_BYTES, lane, F);
        __syncwarp();
        if (lane == 0 && !first_step) {      // same ring discipline as xy_run_item: refill the previous step's slot
            tma_wait_read0();
            issue();
        }
        first_step = false;
        __syncwarp();
        if (++slot == NS) { slot = 0; phase ^= 1u; }
    };
    double lo_face = 0.0, hi_face = 0.0, row0 = 0.0, rowl = 0.0;
    if (ex.has_lo) {
        double F[CH], e[CH - 1];
        take_tile(F);
        double fm1 = F[0], eprev = 0.0;      // guess f[-1] := f[0]
#pragma unroll
        for (int j = 0; j < CH - 1; j++) {
            eprev = fma(-ex.head.l[j], eprev, ex.head.sk[j] * (F[j + 1] - fm1));
            e[j] = eprev;
            fm1 = F[j];
        }
        double x = 0.0;
#pragma unroll
        for (int j = CH - 2; j >= 0; j--) x = fma(-ex.head.g[j], x, e[j]);
        lo_face = -x;
        row0 = F[0];
    }
    if (ex.has_hi) {
        double F[CH];                        // rows n-32 .. n-1
        take_tile(F);
        double eprev = 0.0;
#pragma unroll
        for (int j = 1; j < CH - 1; j++)                                    // rows n-31 .. n-2 from a zero state
            eprev = fma(-ex.l_mid, eprev, ex.sk_mid * (F[j + 1] - F[j - 1]));
        eprev = fma(-ex.l_last, eprev, ex.sk_last * (F[CH - 1] - F[CH - 2]));   // row n-1, guess f[n] := f[n-1]
        hi_face = -eprev;
        rowl = F[CH - 1];
    }
    if (ok) {
        ex.faces[col] = lo_face;
        ex.faces[ex.nlines + col] = hi_face;
        if (ex.peer_face_lo) ex.peer_face_lo[col] = lo_face;
        if (ex.peer_face_hi) ex.peer_face_hi[col] = hi_face;
        if (ex.push_lo) ex.push_lo[col] = row0;
        if (ex.push_hi) ex.push_hi[col] = rowl;
    }
}

template <int NS, bool SEG, bool EDGE>
__global__ void __launch_bounds__(256, 1)
stream_kernel_xy(const __grid_constant__ CUtensorMap tmx_in, const __grid_constant__ CUtensorMap tmx_out,
                 const __grid_constant__ CUtensorMap tmy_in, const __grid_constant__ CUtensorMap tmy_out,
                 const __grid_constant__ KParams px, const __grid_constant__ KParams py,
                 const __grid_constant__ XYParams q, const __grid_constant__ CUtensorMap tmz_in,
                 const __grid_constant__ EdgeX ex)
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
    const int ipp = q.nxp + q.nyp;       // items per plane

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        fence_mbar_init();
    }
    __syncwarp();
    pdl_launch_dependents();
    pdl_wait();

    // draw-table entry -> (direction, bundle, plane, tile range)
    // (edge items, EDGE only: entries beyond the nz * ipp x / y items; b = edge item, tiles 0 .. ke)
    const long nxy = q.nxy;
    auto decode = [&](long e, bool &contig, long &b, int &c0, int &c2, int &kb, int &ke, int &ko, int &kp) -> bool {
        const int seg = (int)(e & 7);
        const long w = e >> 3;
        if constexpr (EDGE) {
            if (w >= nxy) {
                contig = false; b = w - nxy; c0 = 0; c2 = 0; kb = 0; ko = 0;
                ke = ex.has_lo + ex.has_hi - 1; kp = ke + 1;
                return true;
            }
        }
        const long z = w / ipp;
        const int r = (int)(w - z * ipp);
        contig = r < q.nxp;
        if (contig) { b = z * q.nxp + r; c0 = 0; c2 = 0; }
        else        { b = z * q.nyp + (r - q.nxp); c0 = (r - q.nxp) * CH; c2 = (int)z; }
        const int K = contig ? px.K : py.K;
        if (!SEG || seg == 0) { kb = 0; ke = K - 1; ko = 0; kp = K; }
        else {
            const int ks = contig ? q.kseg : q.ksegy;
            const int s0 = (seg - 1) * ks;
            const int s1 = (s0 + ks < K) ? s0 + ks : K;
            ko = s0;
            kp = s1;
            kb = s0 > 0 ? s0 - 1 : 0;
            ke = s1 < K ? s1 : K - 1;
        }
        return false;
    };

    // ---- producer (lane 0): one call per tile position; after the initial fill it runs NS - 1 positions ahead
    long iw = 0, ib = 0;
    int ik = 1, ikend = 0, iko = 0, ikp = 0, ic0 = 0, ic2 = 0, islot = 0;
    bool icontig = true, dry = false, first_draw = true, iedge = false;
    auto issue = [&]() {
        if (ik > ikend && !dry) {
            iw = (long)atomicAdd(q.counter, 1ULL);
            dry = iw >= q.nitems;
            if (!dry) {
                const long pos = iw - (EDGE ? ex.nedge : 0);      // position among the x / y items (edge items come first)
                if (first_draw && pos >= 0 && q.tau_ns > 0.f && q.slot_items > 0.f) {
                    const unsigned long long t0 = global_timer_ns();
                    const unsigned long long wait = (unsigned long long)(floorf((float)pos / q.slot_items) * q.tau_ns);
                    while (global_timer_ns() - t0 < wait) __nanosleep(256);
                }
                if (pos >= 0) first_draw = false;
                iw = q.order ? (long)q.order[iw] : (iw << 3);
                iedge = decode(iw, icontig, ib, ic0, ic2, ik, ikend, iko, ikp);
            }
        }
        if (dry) {
            tag[islot] = -1;
        } else {
            tag[islot] = iw;
            const uint32_t bar = bar0 + 8 * islot;
            const uint32_t dst = smem_u32(wbase + islot * SLOT_BYTES);
            mbar_expect_tx(bar, SLOT_BYTES);
            if (EDGE && iedge) {             // head tile = planes 0..31, tail tile = planes n-32..n-1
                const int row = (ex.has_lo && ik == 0) ? 0 : ex.n - CH;
                tma_load_3d(dst, &tmz_in, bar, (int)(ib * CH), row, 0);
            } else if (SEG && (q.hints & 1)) {
                const unsigned long long pol = l2_policy_evict_last();
                if (icontig) {
                    tma_load_2d_hint(dst, &tmx_in, bar, ik * CH, (int)(ib * CH), pol);
                    tma_load_2d_hint(dst + 4096, &tmx_in, bar, ik * CH + 16, (int)(ib * CH), pol);
                } else {
                    tma_load_3d_hint(dst, &tmy_in, bar, ic0, ik * CH, ic2, pol);
                }
            } else if (icontig) {
                tma_load_2d(dst, &tmx_in, bar, ik * CH, (int)(ib * CH));
                tma_load_2d(dst + 4096, &tmx_in, bar, ik * CH + 16, (int)(ib * CH));
            } else {
                tma_load_3d(dst, &tmy_in, bar, ic0, ik * CH, ic2);
            }
            ++ik;
        }
        if (++islot == NS) islot = 0;
    };
    if (lane == 0) {
#pragma unroll 1
        for (int s = 0; s < NS; s++) issue();
    }
    __syncwarp();

    // ---- consumer
    int slot = 0;
    uint32_t phase = 0;
    bool first_step = true;
    unsigned long long edge_pending = 0;     // EDGE: edge items this warp has finished and not yet reported
    auto edge_report = [&]() {
        __threadfence_system();              // every lane: its face / halo stores are visible system-wide
        __syncwarp();
        if (lane == 0) {
            const unsigned long long prev = atomicAdd(ex.done, edge_pending);
            if (prev + edge_pending == (unsigned long long)ex.nedge) {     // all edge items of the launch are done
                *ex.done = 0ULL;
                if (ex.flag_lo) flag_raise(ex.flag_lo, ex.seq);
                if (ex.flag_hi) flag_raise(ex.flag_hi, ex.seq);
            }
        }
        edge_pending = 0;
    };
    for (;;) {
        const long w = tag[slot];
        if (w < 0) break;
        bool contig;
        long b;
        int c0, c2, kb, ke, ko, kp;
        const bool edge = decode(w, contig, b, c0, c2, kb, ke, ko, kp);
        if constexpr (EDGE) {
            if (edge) {
                xy_run_edge<NS>(ex, b, wbase, bar0, lane, slot, phase, first_step, issue);
                ++edge_pending;
                continue;
            }
            if (edge_pending) edge_report();
        }
        if (contig) xy_run_item<true, NS, SEG>(px, &tmx_out, b, c0, c2, wbase, bar0, lane, slot, phase, first_step, issue, kb, ke, ko, kp, q.hints);
        else        xy_run_item<false, NS, SEG>(py, &tmy_out, b, c0, c2, wbase, bar0, lane, slot, phase, first_step, issue, kb, ke, ko, kp, q.hints);
    }
    if constexpr (EDGE) {
        if (edge_pending) edge_report();
    }
    if (lane == 0) {
        tma_wait_all0();
        __threadfence();
        const unsigned long long total = (unsigned long long)gridDim.x * nwarps;
        if (atomicAdd(q.counter + 1, 1ULL) == total - 1) {
            q.counter[0] = 0ULL;
            q.counter[1] = 0ULL;
            __threadfence();
        }
    }
}

}  // namespace cfd
