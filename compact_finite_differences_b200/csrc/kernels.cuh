// kernels.cuh -- sm_100a device code of the compact-derivative path.
//
// One fused kernel per direction: Pade right-hand side (reference computeRHS,
// code/cuda/kernels.cu:4-47) + batched near-Toeplitz tridiagonal solve (reference
// sharedMemCyclicReduction, code/cuda/solvers/templated/kernels.jinja2:2-123) in ONE pass:
// every point of f is read from HBM once and every point of f' written once (16 B / point).
//
// Design (nothing of the reference's one-block-per-line cyclic reduction survives):
//
//  * WARP-AUTONOMOUS STREAMING.  A warp owns a bundle of 32 lines (lane = line) and marches along
//    them in chunks of CH = 32 rows.  Tiles of 32 rows x 32 lines x 8 B = 8 KiB arrive through
//    TMA (cp.async.bulk.tensor) into a private 3-slot shared-memory ring guarded by the warp's own
//    mbarriers; no CTA-wide barrier exists anywhere, so warps drift freely and hide each other's
//    latencies.  Warps are persistent and draw their bundles from a global counter; the flat tile
//    sequence (bundle, chunk) is prefetched NS tiles ahead, across bundle boundaries.
//  * LU WITH CONSTANT PIVOTS + LOOK-AHEAD BACK-SUBSTITUTION.  The forward elimination
//    e_i = s*r_i - l*e_{i-1} is carried exactly along the whole line in registers.  The backward
//    sweep x_i = e_i - g*x_{i+1} of chunk k-1 is started one chunk to the right, at the end of
//    chunk k, from x = e: the neglected term decays by |g| = 2 - sqrt(3) = 0.268 per row, i.e. by
//    0.268^32 = 5e-19 over the 32 warm-up rows -- below fp64 round-off.  At the true end of the
//    line the sweep is exact.  Forward-eliminated values of two chunks (64 doubles) live in
//    registers; nothing is spilled to memory between the two sweeps, which is what makes the
//    kernel read-once / write-once for any line length (the reference caps nx at 2048).
//  * LAYOUTS.  CONTIG (derivative along x): lines are rows of the [nz*ny, nx] matrix; the tile is two
//    TMA boxes of 32 rows x 16 doubles with 128-byte swizzle, so that lane r reading its own row with
//    16-byte LDS.128 is bank-conflict free (the "tiled transpose" happens inside the TMA unit); the
//    result tile goes back through swizzled shared memory and a TMA store.  STRIDED (y, z): lanes map
//    to 32 consecutive x columns, the tile is one dense [32 rows][32 cols] box, LDS.64 / STS.64 with no
//    conflicts, and the result tile leaves through the same staging slot + TMA store (which also clips
//    ragged rows / columns, so the inner loops carry no predicates and no 64-bit address arithmetic).
//
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tables.h"

namespace cfd {

constexpr int SLOT_BYTES = CH * CH * 8;   // 8 KiB

struct KParams {
    int n;             // rows per line
    int K;             // chunks per line = ceil(n / 32)
    int jl;            // position of row n-1 inside the last chunk
    int kseg, nseg;    // line segmentation: chunks per segment, segments per line (nseg == 1: whole lines)
    int aux_on;        // in-place solve with segments: the first and last result chunk of every segment go to tm_aux
    int inner;         // STRIDED: extent of the contiguous dimension the lanes map to
    int inner_tiles;   // STRIDED: ceil(inner / 32)
    int outer;         // STRIDED: number of outer slices
    int lo_closure;    // this block owns the physical left end  -> one-sided RHS row 0
    int hi_closure;    // this block owns the physical right end -> one-sided RHS row n-1
    long nb;           // number of 32-line bundles
    long rows;         // CONTIG: number of lines
    double sk_mid, l_mid, g_mid;
    double s0c, snc;   // beta_0/(2h), beta_{n-1}/(2h) for the closure rows
    const double *halo_lo, *halo_hi;   // neighbour planes of f (multi-rank), one value per line
    unsigned long long *counter;   // {next bundle, finished warps}, zero on entry, zero again on exit
    // coupled multi-rank solve: the two interface unknowns of every line (alpha = last point of the left
    // neighbour, beta = first point of the right neighbour) folded into rows 0 and n-1 as Dirichlet data
    const double *ab;              // [2][plane] = alpha plane, beta plane; nullptr = block-local solve
    long nlines;                   // plane size
    double snb;                    // beta_{n-1} * c_i: weight of beta in row n-1
    RowTab head, tail;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }

__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}

// Bounded wait: a descriptor mistake must fault, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000LL) __trap();
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// L2 eviction-priority hints for tiles with a known future: evict_last = will be read again soon, evict_first = never.
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2,
                                                 unsigned long long pol)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}

__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1,
                                                 unsigned long long pol)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "l"(pol) : "memory");
}

__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap *tm, uint32_t src, int c0, int c1, unsigned long long pol)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "l"(pol) : "memory");
}

__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap *tm, uint32_t src, int c0, int c1, int c2, unsigned long long pol)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, uint32_t src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, uint32_t src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{ asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// Raise an arrival flag (possibly in a peer GPU's memory) to `seq`, with release semantics at system scope.  A max,
// not a store: flags only ever grow, even if two producer launches of consecutive calls (issued on different streams)
// finish out of order -- a plain store of the older call number would then hide the newer one from the waiter.
__device__ __forceinline__ void flag_raise(unsigned long long *p, unsigned long long v)
{
    __threadfence_system();
    atomicMax_system(p, v);
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may be
// scheduled while its predecessor in the stream is still draining; it must not touch global memory before pdl_wait()
// (everything in front of it -- shared-memory carve-up, mbarrier init -- overlaps the predecessor's tail and the launch
// latency).  Both instructions are no-ops for a plain launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Cross-GPU arrival flags.  A waiter polls with ld.acquire.sys and backs off with __nanosleep (the spinning thread
// must not starve the SM's other warps); after `timeout_ns` it gives up, records CFD_ETIMEOUT in the host-mapped
// error word `err` (the host returns it from the next library call, cfd_async_status) and lets the kernel finish
// on whatever is in the buffers -- a rank that is minutes late (first-call set-up, a debugger, I/O) must not cost
// the waiting rank its CUDA context, which a __trap() would.
struct WaitP {
    unsigned long long timeout_ns;
    int *err;                        // host-mapped (zero-copy) error word, may be nullptr
};

__device__ __forceinline__ void wait_flag(const unsigned long long *flag, unsigned long long seq, const WaitP &w)
{
    if (ld_acquire_sys(flag) >= seq) return;
    const unsigned long long t0 = global_timer_ns();
    unsigned ns = 32;
    while (ld_acquire_sys(flag) < seq) {
        __nanosleep(ns);
        if (ns < 2048) ns <<= 1;
        if (global_timer_ns() - t0 > w.timeout_ns) {
            if (w.err) { *(volatile int *)w.err = -4; __threadfence_system(); }
            return;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Reduced (interface) system of one line, solved for this rank's two unknowns only:
//   alpha = the left neighbour's last point, beta = the right neighbour's first point.
// Rows: code/cuda/compact.py:96-111; same tridiagonal elimination as reducedSolverKernel
// (code/cuda/kernels.cu:115-145) with host-precomputed pivots, run from both ends towards rows (2r, 2r+1)
// and closed with a 2x2 solve, so it needs registers only.
//   lu = [6][2P]: a_i, c_i, 1/p_i, c_i/p_i (top-down pivots p), 1/q_i, a_i/q_i (bottom-up pivots q).
// ------------------------------------------------------------------------------------------------
// PEER: the planes are written by the neighbour GPUs while the calling kernel may already be running (it waits for
// their arrival flags first): they are read with ld.global.cg (L2, never the non-coherent path), see peer_ld().
__device__ __forceinline__ double peer_ld(const double *p) { return __ldcg(p); }

template <bool PEER>
__device__ __forceinline__ void reduced_unknowns(const double *faces_all, const double *__restrict__ lu,
                                                 long nlines, long line, int P, int rank, double &alpha, double &beta)
{
    const int m = 2 * P;
    const double *a = lu, *c = lu + m, *ip = lu + 2 * m, *cp = lu + 3 * m, *iq = lu + 4 * m, *aq = lu + 5 * m;
    const int r0 = 2 * rank, r1 = 2 * rank + 1;
    auto face = [&](int i) { return PEER ? peer_ld(faces_all + (long)i * nlines + line) : faces_all[(long)i * nlines + line]; };
    double t = face(0) * ip[0];
    for (int i = 1; i <= r0; i++) t = (face(i) - a[i] * t) * ip[i];
    double u = face(m - 1) * iq[m - 1];
    for (int i = m - 2; i >= r1; i--) u = (face(i) - c[i] * u) * iq[i];
    // x[r0] = t - cp[r0]*x[r1];  x[r1] = u - aq[r1]*x[r0]
    alpha = (t - cp[r0] * u) / (1.0 - cp[r0] * aq[r1]);
    beta = u - aq[r1] * alpha;
}

// ------------------------------------------------------------------------------------------------
// Tile access.  CONTIG: slot = 2 boxes [32 rows][16 doubles], 128B-swizzled: the 16-byte unit u of row r
// sits at r*128 + ((u ^ (r & 7)) << 4).  STRIDED: slot = [32 rows][32 cols] dense.
// ------------------------------------------------------------------------------------------------
template <bool CONTIG>
__device__ __forceinline__ void load_chunk(const unsigned char *slot, int lane, double (&F)[CH])
{
    if constexpr (CONTIG) {
        const unsigned char *row = slot + lane * 128;
        const int sw = (lane & 7) << 4;
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const double2 v = *reinterpret_cast<const double2 *>(row + (m >> 3) * 4096 + (((m & 7) << 4) ^ sw));
            F[2 * m] = v.x;
            F[2 * m + 1] = v.y;
        }
    } else {
        const double *col = reinterpret_cast<const double *>(slot) + lane;
#pragma unroll
        for (int j = 0; j < CH; j++) F[j] = col[j * CH];
    }
}

template <bool CONTIG>
__device__ __forceinline__ double load_first(const unsigned char *slot, int lane)
{
    if constexpr (CONTIG) return *reinterpret_cast<const double *>(slot + lane * 128 + ((lane & 7) << 4));
    else return reinterpret_cast<const double *>(slot)[lane];
}

// ------------------------------------------------------------------------------------------------
// Forward elimination of one chunk.
//   MODE 0: interior chunk, constant coefficients.  MODE 1: first chunk (HEAD table; also the only
//   chunk when K == 1).  MODE 2: last chunk (TAIL table).
//   DERIV: r_i is the Pade stencil of f (code/cuda/kernels.cu:34-44), else r_i = tile value.
// State carried along the line: eprev = e_{i-1}, fm1 = f_{i-1}, fm2 = f_{i-2}.
// ------------------------------------------------------------------------------------------------
//   TAILPOS: where row n-1 sits in this chunk, resolved at compile time when possible so that the common
//   shapes pay no per-row selects: -1 = not in this chunk, 31 = last row of the chunk (n a multiple of 32),
//   -2 = anywhere (run-time p.jl; ragged n, or a one-chunk line).
template <int MODE, bool DERIV, int TAILPOS>
__device__ __forceinline__ void fwd_chunk(const KParams &p, const double (&F)[CH], double peek, double hval, double bval,
                                          double (&e)[CH], double &eprev, double &fm1, double &fm2)
{
    if constexpr (MODE == 0) {
        const double sk = p.sk_mid, nl = -p.l_mid;
#pragma unroll
        for (int j = 0; j < CH; j++) {
            double r;
            if constexpr (DERIV) {
                const double nxt = (j < CH - 1) ? F[j + 1] : peek;
                r = sk * (nxt - fm1);
                fm1 = F[j];
            } else {
                r = sk * F[j];
            }
            eprev = fma(nl, eprev, r);
            e[j] = eprev;
        }
        if constexpr (DERIV) fm2 = F[CH - 2];
    } else {
        const RowTab &T = (MODE == 1) ? p.head : p.tail;
        const int jl = (TAILPOS == -2) ? p.jl : TAILPOS;
#pragma unroll
        for (int j = 0; j < CH; j++) {
            double r;
            if constexpr (DERIV) {
                const bool at_end = (TAILPOS == -2) ? (j == jl) : (j == TAILPOS);
                double nxt = (j < CH - 1) ? F[j + 1] : peek;
                if (at_end && !p.hi_closure) nxt = hval;
                r = T.sk[j] * (nxt - fm1);
                if (MODE == 1 && j == 0 && p.lo_closure) r = p.s0c * (-5.0 * F[0] + 4.0 * F[1] + F[2]);
                if (at_end) {
                    if (p.hi_closure) r = p.snc * (5.0 * F[j] - 4.0 * fm1 - fm2);
                    else              r = fma(-p.snb, bval, r);      // coupled: - beta_{n-1} c x_n
                }
                fm2 = fm1;
                fm1 = F[j];
            } else {
                r = T.sk[j] * F[j];
            }
            eprev = fma(-T.l[j], eprev, r);
            e[j] = eprev;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward substitution over one chunk, x carried in.  OUT = results of this chunk are final.
// ------------------------------------------------------------------------------------------------
template <int MODE, bool OUT, bool CONTIG>
__device__ __forceinline__ void bwd_chunk(const KParams &p, const double (&e)[CH], double &x, unsigned char *oslot, int lane)
{
    if constexpr (MODE == 0 && !OUT) {
        // Warm-up sweep: only its end value is used, x_out = sum_j (-g)^j e_j + (-g)^32 x_in.  Evaluated as four
        // Horner chains in g^4 (8 + 3 dependent FMAs instead of 32, same flop count): the sweep sits between the
        // forward elimination of this chunk and the output sweep of the previous one on the warp's critical path.
        const double ng = -p.g_mid, g2 = ng * ng, g4 = g2 * g2;
        double c0 = fma(g4, x, e[CH - 4]), c1 = e[CH - 3], c2 = e[CH - 2], c3 = e[CH - 1];
#pragma unroll
        for (int j = CH - 8; j >= 0; j -= 4) {
            c0 = fma(g4, c0, e[j]);
            c1 = fma(g4, c1, e[j + 1]);
            c2 = fma(g4, c2, e[j + 2]);
            c3 = fma(g4, c3, e[j + 3]);
        }
        x = fma(ng, fma(ng, fma(ng, c3, c2), c1), c0);
    } else {
        const RowTab &T = (MODE == 1) ? p.head : p.tail;
        double hold = 0.0;
#pragma unroll
        for (int j = CH - 1; j >= 0; j--) {
            const double ng = (MODE == 0) ? -p.g_mid : -T.g[j];
            x = fma(ng, x, e[j]);
            if constexpr (OUT) {
                if constexpr (CONTIG) {
                    if (j & 1) hold = x;
                    else {
                        const int m = j >> 1;
                        *reinterpret_cast<double2 *>(oslot + lane * 128 + (m >> 3) * 4096 +
                                                     (((m & 7) << 4) ^ ((lane & 7) << 4))) = make_double2(x, hold);
                    }
                } else {
                    reinterpret_cast<double *>(oslot)[j * CH + lane] = x;     // dense [row][col] staging tile
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The streaming kernel.
//   NS   = ring slots per warp (>= 3: tile k, tile k+1 for the stencil peek, the rest in flight).
//   Work distribution is DYNAMIC: a warp draws its next 32-line bundle from a global counter when its
//   producer lane starts prefetching it, so SMs that see less memory bandwidth (fuller GPCs, far L2
//   slices) simply take fewer bundles and all warps run dry together.  The counter pair
//   {next bundle, finished warps} is reset by the last warp to finish, so a launch needs no memset.
// ------------------------------------------------------------------------------------------------
template <bool CONTIG, bool DERIV, int NS>
__global__ void __launch_bounds__(224, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
              const __grid_constant__ KParams p, const __grid_constant__ CUtensorMap tm_aux)
{
    extern __shared__ unsigned char smem_raw[];
    constexpr int PER_WARP = (NS + 2) * SLOT_BYTES;     // ring + two result staging slots
    constexpr int CTRL = NS * 16;       // per warp: NS mbarriers + NS bundle tags
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

    // 1 KiB alignment in the shared window (128B swizzle atom = 8 rows x 128 B)
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *wbase = base + warp * PER_WARP;
    unsigned char *oslot = wbase + NS * SLOT_BYTES;      // staging slot in use; alternates with the one 8 KiB above/below
    int ocur = 0;
    unsigned char *ctrl = base + nwarps * PER_WARP + warp * CTRL;
    const uint32_t bar0 = smem_u32(ctrl);
    volatile long long *tag = reinterpret_cast<volatile long long *>(ctrl + NS * 8);

    const int K = p.K;
    const long nb = p.nb;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        fence_mbar_init();
    }
    __syncwarp();
    pdl_launch_dependents();
    pdl_wait();

    // A work item is (bundle, segment): output chunks [c0, c1) of the bundle's lines.  Its tile sequence runs
    // from chunk max(c0-1, 0) -- a 32-row forward warm-up from a zero state, exact to 0.268^32 like the backward
    // one -- to chunk min(c1, K-1), the look-ahead chunk of the last backward sweep.  With nseg == 1 an item is a
    // whole bundle (the common case); long lines with few bundles are cut so that every warp finds work.
    const int nseg = p.nseg, kseg = p.kseg;
    const long nitems = nb * nseg;
    auto item_range = [&](long w, long &bb, int &kb, int &ke, int &ko) {
        if (nseg == 1) { bb = w; kb = 0; ke = K - 1; ko = 0; return; }     // whole lines: no 64-bit division
        bb = w / nseg;
        const int c0 = (int)(w % nseg) * kseg;
        const int c1 = (c0 + kseg < K) ? c0 + kseg : K;
        ko = c0;
        kb = c0 > 0 ? c0 - 1 : 0;
        ke = c1 < K ? c1 : K - 1;
    };

    // ---- producer side (lane 0): one call per tile position, NS positions ahead of the consumer
    long iw = 0, ib = 0;  // item / bundle being prefetched
    int ik = 1, ikend = 0, iko = 0;   // its next chunk, its last chunk
    int islot = 0;
    bool dry = false;
    auto issue = [&]() {
        if (ik > ikend && !dry) {
            iw = (long)atomicAdd(p.counter, 1ULL);
            dry = iw >= nitems;
            if (!dry) item_range(iw, ib, ik, ikend, iko);
        }
        if (dry) {
            tag[islot] = -1;
        } else {
            tag[islot] = iw;
            const uint32_t bar = bar0 + 8 * islot;
            const uint32_t dst = smem_u32(wbase + islot * SLOT_BYTES);
            mbar_expect_tx(bar, SLOT_BYTES);
            if constexpr (CONTIG) {
                tma_load_2d(dst, &tm_in, bar, ik * CH, (int)(ib * CH));
                tma_load_2d(dst + 4096, &tm_in, bar, ik * CH + 16, (int)(ib * CH));
            } else {
                const int o = (int)(ib / p.inner_tiles), it = (int)(ib % p.inner_tiles);
                tma_load_3d(dst, &tm_in, bar, it * CH, ik * CH, o);
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

    // ---- consumer side (all lanes)
    double eA[CH], eB[CH], F[CH];
    double eprev = 0.0, fm1 = 0.0, fm2 = 0.0, hval = 0.0, bval = 0.0;
    long b = 0;
    int k = 0, kbeg = 0, kend = 0, kout = 0, slot = 0;
    bool fresh = true;            // the next tile opens a new work item
    uint32_t phase = 0;
    int oc0 = 0, oc2 = 0;         // STRIDED: first column / outer index of the bundle (TMA store coordinates)

    // Per-line boundary data of a partitioned block (neighbour points of f, interface unknowns) are fetched
    // one bundle ahead, while the previous bundle's last chunk is being computed, so that their global-load
    // latency never sits in front of the forward recurrence.
    long pf_id = -1;
    double pf_lo = 0.0, pf_hi = 0.0, pf_a = 0.0, pf_b = 0.0;
    auto prefetch_edge = [&](long nb_) {
        long line;
        bool lane_ok;
        if constexpr (CONTIG) {
            line = nb_ * CH + lane;
            lane_ok = line < p.rows;
        } else {
            const int c0 = (int)(nb_ % p.inner_tiles) * CH;
            lane_ok = c0 + lane < p.inner;
            line = (nb_ / p.inner_tiles) * (long)p.inner + c0 + lane;
        }
        pf_lo = pf_hi = pf_a = pf_b = 0.0;
        if (lane_ok) {
            if (!p.lo_closure) pf_lo = __ldg(p.halo_lo + line);
            if (!p.hi_closure) pf_hi = __ldg(p.halo_hi + line);
            if (p.ab != nullptr) {
                pf_a = __ldg(p.ab + line);
                pf_b = __ldg(p.ab + p.nlines + line);
            }
        }
        pf_id = nb_;
    };

    for (;;) {
        if (fresh) {
            const long w = tag[slot];
            if (w < 0) break;
            item_range(w, b, kbeg, kend, kout);
            k = kbeg;
            fresh = false;
            if constexpr (!CONTIG) {
                oc2 = (int)(b / p.inner_tiles);
                oc0 = (int)(b % p.inner_tiles) * CH;
            }
            eprev = 0.0; fm1 = 0.0; fm2 = 0.0; hval = 0.0; bval = 0.0;
            if constexpr (DERIV) {
                if (!p.lo_closure || !p.hi_closure) {      // block of a partitioned line
                    if (pf_id != b) prefetch_edge(b);       // first bundle of this warp: nothing was prefetched
                    fm1 = pf_lo;                            // f[-1]: neighbour plane of f
                    hval = pf_hi;                           // f[n]
                    eprev = pf_a;                           // coupled: row 0 sees x_{-1} = alpha through l_0 = a_i*beta_0
                    bval = pf_b;                            //          row n-1 sees x_n = beta through snb
                }
            }
        }
        const bool last = (k == K - 1);
        if constexpr (DERIV) {
            if (k == kend && (!p.lo_closure || !p.hi_closure)) {
                const long nxt = tag[(slot + 1 == NS) ? 0 : slot + 1];     // the next tile opens the next item
                if (nxt >= 0) prefetch_edge(nseg == 1 ? nxt : nxt / nseg);
            }
        }
        const unsigned char *st = wbase + slot * SLOT_BYTES;
        mbar_wait(bar0 + 8 * slot, phase);
        load_chunk<CONTIG>(st, lane, F);
        double peek = 0.0;
        if constexpr (DERIV) {
            if (k < kend) {               // the item's next tile is in the ring: peek at its first row
                const int s1 = (slot + 1 == NS) ? 0 : slot + 1;
                const uint32_t ph1 = (slot + 1 == NS) ? (phase ^ 1u) : phase;
                mbar_wait(bar0 + 8 * s1, ph1);
                peek = load_first<CONTIG>(wbase + s1 * SLOT_BYTES, lane);
            }
        }
        if (k == 0) {
            if (last) fwd_chunk<1, DERIV, -2>(p, F, peek, hval, bval, eB, eprev, fm1, fm2);     // one-chunk line
            else      fwd_chunk<1, DERIV, -1>(p, F, peek, hval, bval, eB, eprev, fm1, fm2);
        } else if (last) {
            if (p.jl == CH - 1) fwd_chunk<2, DERIV, CH - 1>(p, F, peek, hval, bval, eB, eprev, fm1, fm2);
            else                fwd_chunk<2, DERIV, -2>(p, F, peek, hval, bval, eB, eprev, fm1, fm2);
        } else {
            fwd_chunk<0, DERIV, -1>(p, F, peek, hval, bval, eB, eprev, fm1, fm2);
        }

        // the slot has been consumed into registers: refill it (tile position t + NS)
        __syncwarp();
        if (lane == 0) issue();
        __syncwarp();

        // ---- backward sweeps
        auto flush = [&](int kc) {       // ship the staged result tile of chunk kc (TMA clips rows >= n, cols >= extent)
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                // In-place solve of segmented lines (solver-only): a segment's first and last result chunk are its
                // neighbours' look-ahead / warm-up INPUT tiles, so they must not land in the field while any segment
                // may still read them.  They go to the side buffer tm_aux (a field of 2 chunks per segment: chunk
                // 2 seg = first, 2 seg + 1 = last) and are copied into place by scatter_aux_kernel afterwards.
                const CUtensorMap *tm = &tm_out;
                int row = kc * CH;
                if constexpr (!DERIV) {
                    if (p.aux_on) {
                        const int c1 = (kout + kseg < K) ? kout + kseg : K;
                        if (kc == kout || kc == c1 - 1) {
                            tm = &tm_aux;
                            row = (2 * (kout / kseg) + (kc == kout ? 0 : 1)) * CH;
                        }
                    }
                }
                if constexpr (CONTIG) {
                    tma_store_2d(tm, smem_u32(oslot), row, (int)(b * CH));
                    tma_store_2d(tm, smem_u32(oslot) + 4096, row + 16, (int)(b * CH));
                } else {
                    tma_store_3d(tm, smem_u32(oslot), oc0, row, oc2);
                }
                tma_commit();
            }
            ocur ^= 1;                   // next result tile goes to the other staging slot
            oslot = wbase + (NS + ocur) * SLOT_BYTES;
        };
        auto acquire_out = [&]() {       // the store issued from THIS slot two tiles ago must have drained it;
            if (lane == 0) tma_wait_read1();   // the most recent one (other slot) may still be in flight
            __syncwarp();
        };
        double x = 0.0;
        if (last) {
            acquire_out();
            if (k == 0) bwd_chunk<1, true, CONTIG>(p, eB, x, oslot, lane);
            else        bwd_chunk<2, true, CONTIG>(p, eB, x, oslot, lane);
            flush(k);
            if (k > kbeg && k - 1 >= kout) {
                acquire_out();
                if (k == 1) bwd_chunk<1, true, CONTIG>(p, eA, x, oslot, lane);
                else        bwd_chunk<0, true, CONTIG>(p, eA, x, oslot, lane);
                flush(k - 1);
            }
        } else if (k > kbeg) {
            bwd_chunk<0, false, CONTIG>(p, eB, x, oslot, lane);   // 32-row warm-up
            if (k - 1 >= kout) {                                  // (chunk kbeg of a later segment is warm-up only)
                acquire_out();
                if (k == 1) bwd_chunk<1, true, CONTIG>(p, eA, x, oslot, lane);
                else        bwd_chunk<0, true, CONTIG>(p, eA, x, oslot, lane);
                flush(k - 1);
            }
        }
#pragma unroll
        for (int j = 0; j < CH; j++) eA[j] = eB[j];

        if (k == kend) fresh = true; else ++k;
        if (++slot == NS) { slot = 0; phase ^= 1u; }
    }
    if (lane == 0) {
        tma_wait_all0();
        // last warp out re-arms the counters for the next launch that uses this pair
        __threadfence();
        const unsigned long long total = (unsigned long long)gridDim.x * nwarps;
        if (atomicAdd(p.counter + 1, 1ULL) == total - 1) {
            p.counter[0] = 0ULL;
            p.counter[1] = 0ULL;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Exact two-pass solver for matrices the one-pass kernel must refuse (pivots that do not converge, or an
// interior coupling |g|^32 above fp64 round-off, e.g. the 6th-order scheme with alpha = 1/3 or the
// non-dominant (1,2,3,4,5,6,7) test matrix at n > 64).  One launch per sweep of the LU solve, each a
// first-order recurrence along the line with per-row coefficients from global tables:
//     y_i = pc_i * v_i - qc_i * y_(i-1)    (REVERSE: y_(i+1)),
// forward:  pc = beta_i,  qc = a_i*beta_i  (v = right-hand side, y = e)
// backward: pc = 1,       qc = c_i*beta_i  (v = e,               y = x)
// Same warp-autonomous TMA ring, dynamic bundle draw and staged TMA store as stream_kernel; 32 B/unknown.
// ------------------------------------------------------------------------------------------------
struct RParams {
    int K, inner_tiles;
    long nb;
    const double *pc, *qc;         // [K*32] device tables, zero beyond row n-1
    unsigned long long *counter;
};

template <bool CONTIG, bool REVERSE>
__global__ void __launch_bounds__(224, 1)
recurrence_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                  const __grid_constant__ RParams p)
{
    extern __shared__ unsigned char smem_raw[];
    constexpr int NS = 3;
    constexpr int PER_WARP = (NS + 1) * SLOT_BYTES;
    constexpr int CTRL = NS * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *wbase = base + warp * PER_WARP;
    unsigned char *oslot = wbase + NS * SLOT_BYTES;
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
            const int kc = REVERSE ? K - 1 - ik : ik;
            mbar_expect_tx(bar, SLOT_BYTES);
            if constexpr (CONTIG) {
                tma_load_2d(dst, &tm_in, bar, kc * CH, (int)(ib * CH));
                tma_load_2d(dst + 4096, &tm_in, bar, kc * CH + 16, (int)(ib * CH));
            } else {
                tma_load_3d(dst, &tm_in, bar, (int)(ib % p.inner_tiles) * CH, kc * CH, (int)(ib / p.inner_tiles));
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

    double F[CH];
    double carry = 0.0;
    long b = 0;
    int k = 0, slot = 0;
    uint32_t phase = 0;
    for (;;) {
        if (k == 0) {
            b = tag[slot];
            if (b < 0) break;
            carry = 0.0;
        }
        const int kc = REVERSE ? K - 1 - k : k;
        mbar_wait(bar0 + 8 * slot, phase);
        load_chunk<CONTIG>(wbase + slot * SLOT_BYTES, lane, F);
        __syncwarp();
        if (lane == 0) issue();
        __syncwarp();
        const double *pc = p.pc + kc * CH, *qc = p.qc + kc * CH;
#pragma unroll
        for (int jj = 0; jj < CH; jj++) {
            const int j = REVERSE ? CH - 1 - jj : jj;
            carry = fma(-__ldg(qc + j), carry, __ldg(pc + j) * F[j]);
            F[j] = carry;
        }
        if (lane == 0) tma_wait_read0();
        __syncwarp();
        if constexpr (CONTIG) {
            const int sw = (lane & 7) << 4;
#pragma unroll
            for (int m = 0; m < 16; m++)
                *reinterpret_cast<double2 *>(oslot + lane * 128 + (m >> 3) * 4096 + (((m & 7) << 4) ^ sw)) =
                    make_double2(F[2 * m], F[2 * m + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < CH; j++) reinterpret_cast<double *>(oslot)[j * CH + lane] = F[j];
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            if constexpr (CONTIG) {
                tma_store_2d(&tm_out, smem_u32(oslot), kc * CH, (int)(b * CH));
                tma_store_2d(&tm_out, smem_u32(oslot) + 4096, kc * CH + 16, (int)(b * CH));
            } else {
                tma_store_3d(&tm_out, smem_u32(oslot), (int)(b % p.inner_tiles) * CH, kc * CH, (int)(b / p.inner_tiles));
            }
            tma_commit();
        }
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

// ------------------------------------------------------------------------------------------------
// Multi-rank helpers (z-partition): interface planes, reduced system, correction.
// Element (o, i, c) of a block lives at (o*n + i)*inner + c; its line index is o*inner + c.
// ------------------------------------------------------------------------------------------------
struct GeomP { long nlines; int n; long inner; };

// faces[0][line] = -x_R[first], faces[1][line] = -x_R[last]; zero at physical ends
// (reference negateAndCopyFaces, code/cuda/kernels.cu:76-113).
__global__ void interface_pack_kernel(const double *__restrict__ x, double *__restrict__ faces, GeomP g,
                                      int lo_phys, int hi_phys)
{
    const long line = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= g.nlines) return;
    const long o = line / g.inner, c = line % g.inner;
    const long first = (o * g.n) * g.inner + c;
    const long lastp = (o * g.n + (g.n - 1)) * g.inner + c;
    faces[line] = lo_phys ? 0.0 : -x[first];
    faces[g.nlines + line] = hi_phys ? 0.0 : -x[lastp];
}

// Reduced 2P-unknown system per line, solved redundantly by every rank after the all-gather
// (rows: code/cuda/compact.py:96-111; same tridiagonal elimination as reducedSolverKernel,
// code/cuda/kernels.cu:115-145, with the pivots precomputed on the host), then
// x += alpha*x_UH + beta*x_LH (sumSolutions, kernels.cu:49-74) on the `wc` rows next to each interface
// where the secondary solutions exceed fp64 round-off.
// A rank needs only its own two unknowns u[2r] = alpha, u[2r+1] = beta, so the elimination runs from
// both ends towards rows (2r, 2r+1) and finishes with a 2x2 solve -- registers only, no per-thread array.
//   lu = [6][2P]: a_i, c_i, 1/p_i, c_i/p_i (top-down pivots p), 1/q_i, a_i/q_i (bottom-up pivots q).
template <bool CONTIG>
__global__ void reduced_correct_kernel(double *__restrict__ x, const double *__restrict__ faces_all,
                                       const double *__restrict__ lu, const double *__restrict__ x_uh,
                                       const double *__restrict__ x_lh, GeomP g, int P, int rank, int wc)
{
    // CONTIG: one warp per line, lanes stride over the band rows (contiguous in memory).
    // STRIDED: one thread per line, loop over the band rows (coalesced across threads).
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long line = CONTIG ? (tid >> 5) : tid;
    const int lane = threadIdx.x & 31;
    if (line >= g.nlines) return;
    double alpha, beta;
    reduced_unknowns<false>(faces_all, lu, g.nlines, line, P, rank, alpha, beta);
    const long o = line / g.inner, col = line % g.inner;
    double *xl = x + (o * g.n) * g.inner + col;
    const int n = g.n;
    const bool all = 2 * wc >= n;
    if (CONTIG) {
        if (all) {
            for (int i = lane; i < n; i += 32) xl[i] += alpha * x_uh[i] + beta * x_lh[i];
        } else {
            for (int i = lane; i < wc; i += 32) xl[i] += alpha * x_uh[i] + beta * x_lh[i];
            for (int i = n - wc + lane; i < n; i += 32) xl[i] += alpha * x_uh[i] + beta * x_lh[i];
        }
    } else {
        if (all) {
            for (int i = 0; i < n; i++) xl[(long)i * g.inner] += alpha * x_uh[i] + beta * x_lh[i];
        } else {
            for (int i = 0; i < wc; i++) xl[(long)i * g.inner] += alpha * x_uh[i] + beta * x_lh[i];
            for (int i = n - wc; i < n; i++) xl[(long)i * g.inner] += alpha * x_uh[i] + beta * x_lh[i];
        }
    }
}

// Interface unknowns of every line of this rank: ab[0] = alpha plane, ab[1] = beta plane, from the gathered
// interface planes (all 2P of them, or the neighbour-only buffer with its own small elimination table).
// Optionally first waits (bounded spin) until the neighbours' planes of call `seq` have landed.
__global__ void __launch_bounds__(256)
reduced_planes_kernel(const double *faces, const double *__restrict__ lu, long nlines, int P, int rank,
                      double *ab, const unsigned long long *flag0, const unsigned long long *flag1,
                      unsigned long long seq, WaitP wp)
{
    // `faces` is peer-written while this kernel runs (no __restrict__, no const-cache loads: peer_ld)
    pdl_launch_dependents();
    pdl_wait();
    if (flag0 || flag1) {
        if (threadIdx.x == 0) {
            if (flag0) wait_flag(flag0, seq, wp);
            if (flag1) wait_flag(flag1, seq, wp);
        }
        __syncthreads();
    }
    const long line = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= nlines) return;
    double alpha, beta;
    reduced_unknowns<true>(faces, lu, nlines, line, P, rank, alpha, beta);
    ab[line] = alpha;
    ab[nlines + line] = beta;
}

// The same for the one-launch exchange (cfd_edge_faces_push + cfd_reduced_unknowns_deferred): the four faces of the
// neighbour-only system that involve a block boundary were computed without the neighbour point of f; add it here,
// from planes this rank owns (its own first / last row: the neighbours' missing points) or has received (its halos).
// Every face was computed with its missing point guessed as the block's own boundary row, which is exactly the
// plane the other side holds:
//   own lo face += w_lo * (halo_lo - our row 0)       left neighbour's hi face  += w_hi * (our row 0 - halo_lo)
//   own hi face += w_hi * (halo_hi - our row n-1)     right neighbour's lo face += w_lo * (our row n-1 - halo_hi)
__global__ void __launch_bounds__(256)
reduced_planes_deferred_kernel(const double *faces, const double *__restrict__ lu, long nlines, int pv,
                               int own, double *ab, const double *halo_lo, const double *halo_hi,
                               const double *__restrict__ f, long inner, int n,
                               double w_lo, double w_hi, const unsigned long long *flag0,
                               const unsigned long long *flag1, unsigned long long seq, WaitP wp)
{
    // faces / halo_lo / halo_hi are peer-written while this kernel runs: no __restrict__, loads through peer_ld
    pdl_launch_dependents();
    pdl_wait();
    if (flag0 || flag1) {
        if (threadIdx.x == 0) {
            if (flag0) wait_flag(flag0, seq, wp);
            if (flag1) wait_flag(flag1, seq, wp);
        }
        __syncthreads();
    }
    const long line = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= nlines) return;
    double v[6];
#pragma unroll
    for (int i = 0; i < 6; i++) v[i] = (i < 2 * pv) ? peer_ld(faces + (long)i * nlines + line) : 0.0;
    const long o = line / inner, col = line % inner;
    const double *fl = f + (o * n) * inner + col;
    if (halo_lo) {                                       // there is a left neighbour
        const double d = peer_ld(halo_lo + line) - fl[0];
        v[2 * own] += w_lo * d;
        v[2 * own - 1] -= w_hi * d;
    }
    if (halo_hi) {                                       // there is a right neighbour
        const double d = peer_ld(halo_hi + line) - fl[(long)(n - 1) * inner];
        v[2 * own + 1] += w_hi * d;
        v[2 * own + 2] -= w_lo * d;
    }
    double alpha, beta;
    reduced_unknowns<false>(v, lu, 1, 0, pv, own, alpha, beta);
    ab[line] = alpha;
    ab[nlines + line] = beta;
}

// ------------------------------------------------------------------------------------------------
// Interface planes straight from f, WITHOUT the block solve: faces[0] = -x_R[0], faces[1] = -x_R[n-1]
// (what negateAndCopyFaces, code/cuda/kernels.cu:76-113, extracts after the reference's full local solve).
// x_R[0] depends on the first rows only and x_R[n-1] on the last rows only, up to 0.268^31 = 1.9e-18:
//   head: forward rows 0..30 with the HEAD table (row 30 reads f[31]), back-substitute from x_30 = e_30 down to x_0;
//   tail: forward rows n-31..n-1 from a zero state with the MID constants and the last TAIL row; x_{n-1} = e_{n-1}.
// One thread per line; reads the first 32 and the last 32 rows of f -- exactly one 32-row tile per block end, which is
// what lets the same work ride in the fused x/y launch as two-tile items (kernels_xy.cuh xy_run_edge).  Needs n >= 66.
// ------------------------------------------------------------------------------------------------
struct EdgeP {
    long nlines, inner;
    int n, jl;
    int lo_closure, hi_closure;
    double sk_mid, l_mid, s0c, snc, sk_last, l_last;
    const double *halo_lo, *halo_hi;
    // NVLink peer-memory variant: the two faces are also stored straight into the neighbours' interface
    // buffers (peer addresses), and the last CTA to finish raises the neighbours' arrival flags to `seq`.
    double *peer_lo, *peer_hi;                 // left neighbour's slot for our faces[0], right neighbour's for faces[1]
    unsigned long long *flag_lo, *flag_hi;     // arrival flags in the neighbours' memory
    unsigned long long *done;                  // local CTA counter (zero on entry, zero on exit)
    unsigned long long seq;
    // One-launch exchange (cfd_edge_faces_push): the neighbour points of f are not known yet -- the faces are
    // computed with the guesses f[-1] := f[0], f[n] := f[n-1] (they enter LINEARLY, with plan-time weights: the
    // consumer adds w * (halo - guess) later, reduced_planes_deferred_kernel; a zero guess would make the face the
    // difference of two numbers 1/h times larger) and this rank's own first / last row is stored into the
    // neighbours' halo slots.
    int defer;
    double *push_lo, *push_hi;                 // left neighbour's slot for our row 0, right neighbour's for our row n-1
    RowTab head;
};

// Last CTA of a grid to arrive publishes `seq` to the (peer) flags; every CTA's stores were made visible
// system-wide by its own __threadfence_system() before it arrived.
__device__ __forceinline__ void publish_when_grid_done(unsigned long long *done, unsigned long long *flag_a,
                                                       unsigned long long *flag_b, unsigned long long seq)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(done, 1ULL) == (unsigned long long)gridDim.x - 1) {
            *done = 0ULL;
            if (flag_a) flag_raise(flag_a, seq);
            if (flag_b) flag_raise(flag_b, seq);
        }
    }
}

// The 64 rows are read once and must not push the tiles other kernels share through L2 out of it (the exchange chain
// runs beside the fused d/dx + d/dy launch): streaming (evict-first) loads.
#ifndef EDGE_LD
#define EDGE_LD __ldcs
#endif
__global__ void __launch_bounds__(128)
edge_faces_kernel(const double *__restrict__ f, double *__restrict__ faces, const __grid_constant__ EdgeP p)
{
    const long line = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = line < p.nlines;
    const long o = active ? line / p.inner : 0, col = active ? line % p.inner : 0;
    const double *fl = f + (o * p.n) * p.inner + col;
    const long st = p.inner;
    double lo_face = 0.0, hi_face = 0.0;
    if (active && !p.lo_closure) {
        double F[CH], e[CH - 1];
#pragma unroll
        for (int j = 0; j < CH; j++) F[j] = EDGE_LD(fl + (long)j * st);
        double fm1 = p.defer ? F[0] : __ldg(p.halo_lo + line), eprev = 0.0;
        if (p.push_lo) p.push_lo[line] = F[0];            // our first row = the left neighbour's f[n]
#pragma unroll
        for (int j = 0; j < CH - 1; j++) {
            eprev = fma(-p.head.l[j], eprev, p.head.sk[j] * (F[j + 1] - fm1));
            e[j] = eprev;
            fm1 = F[j];
        }
        double x = 0.0;
#pragma unroll
        for (int j = CH - 2; j >= 0; j--) x = fma(-p.head.g[j], x, e[j]);
        lo_face = -x;
    }
    if (active && !p.hi_closure) {
        const int n = p.n;
        const double *ft = fl + (long)(n - CH) * st;           // rows n-32 .. n-1
        double F[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) F[j] = EDGE_LD(ft + (long)j * st);
        const double hval = p.defer ? F[CH - 1] : __ldg(p.halo_hi + line);
        if (p.push_hi) p.push_hi[line] = F[CH - 1];       // our last row = the right neighbour's f[-1]
        double eprev = 0.0;
#pragma unroll
        for (int j = 1; j < CH - 1; j++)                        // rows n-31 .. n-2
            eprev = fma(-p.l_mid, eprev, p.sk_mid * (F[j + 1] - F[j - 1]));
        eprev = fma(-p.l_last, eprev, p.sk_last * (hval - F[CH - 2]));   // row n-1: neighbour point from the halo
        hi_face = -eprev;
    }
    if (active) {
        faces[line] = lo_face;
        faces[p.nlines + line] = hi_face;
        if (p.peer_lo) p.peer_lo[line] = lo_face;        // NVLink store into the left neighbour's buffer
        if (p.peer_hi) p.peer_hi[line] = hi_face;        // ... and the right neighbour's
    }
    if (p.done) publish_when_grid_done(p.done, p.flag_lo, p.flag_hi, p.seq);
}

// ------------------------------------------------------------------------------------------------
// Distributed npts (lanl-implementation/npts.c:329-562, python/npts.py:228-382), the LANL alternative
// to the reduced-system method: ONE LU of the whole line with pivots handed from rank to rank (precompute_beta_gam),
// every rank sweeping its block as  u = phi + u~ psi  (L->R) and  x = phi' + x~ psi'  (R->L), where u~ is the
// forward-eliminated value at the left neighbour's last row and x~ the solution at the right neighbour's first row.
// The reference runs both sweeps from zero states, materialises phi, psi, u (five full-size arrays), all-gathers the end
// faces and combines prefixes over all ranks.  Here psi is never formed: a sweep that STARTS from the incoming value is
// phi + u~ psi, so the coupled stream_kernel does both sweeps in one pass once u~ and x~ are known -- and since psi decays
// by 0.268 per row, u~ and x~ depend on the 32 rows next to the interface only (0.268^32 = 5e-19; the prefix
// combination over ranks further away multiplies by psi_last ~ 1e-37):
//   PHASE 0 (u~ for the right neighbour): rows n-32..n-1 forward from a zero state, GLOBAL pivots; stored into the right
//            neighbour's buffer;
//   PHASE 1 (x~ for the left neighbour): rows 0..31 forward from the received u~, back-substitution from x_31 = u_31;
//            stored into the left neighbour's buffer.
// One thread per line, the layout and tables of edge_faces_kernel (tables = this rank's slice of the global pivots).
template <int PHASE>
__global__ void __launch_bounds__(128)
npts_edge_kernel(const double *__restrict__ f, const __grid_constant__ EdgeP p, const double *u_in, double *peer_out,
                 unsigned long long *peer_flag)
{
    const long line = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = line < p.nlines;
    const long o = active ? line / p.inner : 0, col = active ? line % p.inner : 0;
    const double *fl = f + (o * p.n) * p.inner + col;
    const long st = p.inner;
    if (active) {
        if constexpr (PHASE == 0) {
            const double *ft = fl + (long)(p.n - CH) * st;         // rows n-32 .. n-1
            double F[CH];
#pragma unroll
            for (int j = 0; j < CH; j++) F[j] = EDGE_LD(ft + (long)j * st);
            const double hval = __ldg(p.halo_hi + line);
            double eprev = 0.0;
#pragma unroll
            for (int j = 1; j < CH - 1; j++)
                eprev = fma(-p.l_mid, eprev, p.sk_mid * (F[j + 1] - F[j - 1]));
            eprev = fma(-p.l_last, eprev, p.sk_last * (hval - F[CH - 2]));
            peer_out[line] = eprev;                                // u at this block's last row
        } else {
            double F[CH], e[CH - 1];
#pragma unroll
            for (int j = 0; j < CH; j++) F[j] = EDGE_LD(fl + (long)j * st);
            double fm1 = __ldg(p.halo_lo + line), eprev = peer_ld(u_in + line);
#pragma unroll
            for (int j = 0; j < CH - 1; j++) {
                eprev = fma(-p.head.l[j], eprev, p.head.sk[j] * (F[j + 1] - fm1));
                e[j] = eprev;
                fm1 = F[j];
            }
            double x = 0.0;
#pragma unroll
            for (int j = CH - 2; j >= 0; j--) x = fma(-p.head.g[j], x, e[j]);
            peer_out[line] = x;                                    // solution at this block's first row
        }
    }
    publish_when_grid_done(p.done, peer_flag, nullptr, p.seq);
}

// Halo push over NVLink: copy this rank's first / last plane of f into the neighbours' halo buffers (peer
// addresses) and raise their arrival flags (replaces the NCCL send/recv pair of the halo exchange).
__global__ void __launch_bounds__(256)
push_planes_kernel(const double2 *__restrict__ src0, double2 *__restrict__ dst0, const double2 *__restrict__ src1,
                   double2 *__restrict__ dst1, long n2, unsigned long long *flag0, unsigned long long *flag1,
                   unsigned long long *done, unsigned long long seq)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        if (dst0) dst0[i] = src0[i];
        if (dst1) dst1[i] = src1[i];
    }
    publish_when_grid_done(done, flag0, flag1, seq);
}

// Stream-ordered wait until the neighbours' data of call `seq` has landed in this rank's memory.
__global__ void wait_flags_kernel(const unsigned long long *flag0, const unsigned long long *flag1,
                                  unsigned long long seq, WaitP wp)
{
    if (flag0) wait_flag(flag0, seq, wp);
    if (flag1) wait_flag(flag1, seq, wp);
}

// ------------------------------------------------------------------------------------------------
// The reference's individual stages, for callers that drive the path stage by stage the way
// CompactFiniteDifferenceSolver.dfdx does (code/cuda/compact.py:40-44).  The fused kernels above do not use them.
// ------------------------------------------------------------------------------------------------
// computeRHS (code/cuda/kernels.cu:4-47) without the ghosted copy: one thread per point, neighbours `stride` apart,
// one-sided closures at physical ends, halo planes at block ends.  Memory order = thread order (coalesced).
__global__ void __launch_bounds__(256)
rhs_kernel(const double *__restrict__ f, double *__restrict__ rhs, long total, int n, long stride, double h,
           int lo_closure, int hi_closure, const double *__restrict__ halo_lo, const double *__restrict__ halo_hi)
{
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const long i = (p / stride) % n;                       // coordinate along the derivative axis
    const long line = (p / (stride * n)) * stride + (p % stride);
    const double k = 3. / (4 * h);
    double r;
    if (i == 0) {
        if (lo_closure) r = (1. / (2 * h)) * (-5 * f[p] + 4 * f[p + stride] + f[p + 2 * stride]);
        else            r = k * (f[p + stride] - halo_lo[line]);
    } else if (i == n - 1) {
        if (hi_closure) r = -(1. / (2 * h)) * (-5 * f[p] + 4 * f[p - stride] + f[p - 2 * stride]);
        else            r = k * (halo_hi[line] - f[p - stride]);
    } else {
        r = k * (f[p + stride] - f[p - stride]);
    }
    rhs[p] = r;
}

// sumSolutions (code/cuda/kernels.cu:49-74): x += alpha[line]*x_UH[i] + beta[line]*x_LH[i] over the WHOLE block.
__global__ void __launch_bounds__(256)
sum_solutions_kernel(double *__restrict__ x, const double *__restrict__ alpha, const double *__restrict__ beta,
                     const double *__restrict__ x_uh, const double *__restrict__ x_lh, long total, int n, long stride)
{
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const long i = (p / stride) % n;
    const long line = (p / (stride * n)) * stride + (p % stride);
    x[p] += alpha[line] * x_uh[i] + beta[line] * x_lh[i];
}

// Second half of the in-place solve of segmented lines: the first / last result chunk of every segment, parked in the
// side buffer by stream_kernel (aux_on), into their place in the field.  aux is a field with n_aux = 2 nseg 32 rows per
// line; element (line, r): chunk a = r / 32 -> segment a / 2, first (a even) or last (a odd) chunk of the segment.
__global__ void __launch_bounds__(256)
scatter_aux_kernel(double *__restrict__ d, const double *__restrict__ aux, long total, int n, int n_aux, long inner,
                   int kseg, int K)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long c = t % inner, rest = t / inner;
    const int r = (int)(rest % n_aux);
    const long o = rest / n_aux;
    const int a = r / CH, seg = a >> 1;
    const int c0 = seg * kseg, c1 = (c0 + kseg < K) ? c0 + kseg : K;
    if ((a & 1) && c1 - 1 == c0) return;                    // one-chunk segment: its only chunk was stored as "first"
    const int row = ((a & 1) ? c1 - 1 : c0) * CH + (r % CH);
    if (row < n) d[(o * n + row) * inner + c] = aux[t];
}

// Thread-parallel Thomas over interleaved systems sharing one matrix (reference reducedSolverKernel,
// code/cuda/kernels.cu:115-145).  lu = [3][n]: a_i, 1/pivot_i, c_i/pivot_i.
__global__ void pthomas_kernel(double *__restrict__ d, const double *__restrict__ lu, int n, long nsys)
{
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsys) return;
    const double *a = lu, *ip = lu + n, *cp = lu + 2 * n;
    double prev = d[s] * ip[0];
    d[s] = prev;
    for (int i = 1; i < n; i++) {
        prev = (d[s + (long)i * nsys] - a[i] * prev) * ip[i];
        d[s + (long)i * nsys] = prev;
    }
    for (int i = n - 2; i >= 0; i--) {
        prev = d[s + (long)i * nsys] - cp[i] * prev;
        d[s + (long)i * nsys] = prev;
    }
}

}  // namespace cfd
