// api.cu -- C ABI of libcfd_b200.so (see include/cfd_b200.h).  Host logic only: plan construction
// (coefficient tables, secondary systems, reduced matrix), TMA descriptor encoding, kernel launches.
// No field arithmetic happens on the host and there is no CPU fallback.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifndef CFD_PDL_DEFAULT
#define CFD_PDL_DEFAULT 0
#endif
#include <algorithm>
#include <utility>
#include "../../include/cfd_b200.h"
#include "kernels.cuh"
#include "kernels_xy.cuh"
#include "kernels_general.cuh"
#include "kernels_zx.cuh"

using namespace cfd;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long> g_launches{0};
static int g_warps = 0, g_ctas = 0, g_slots = 0, g_kseg = 0;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return fail(CFD_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

extern "C" const char *cfd_last_error(void) { return g_err.c_str(); }
extern "C" int cfd_version(void) { return CFD_B200_VERSION; }
extern "C" long cfd_launch_count(void) { return g_launches.load(); }
extern "C" int cfd_set_launch(int warps_per_cta, int ctas_per_sm, int ring_slots)
{
    if (warps_per_cta < 0 || warps_per_cta > 8 || ctas_per_sm < 0 || ctas_per_sm > 8 ||
        (ring_slots != 0 && (ring_slots < 3 || ring_slots > 5)))
        return fail(CFD_EINVAL, "cfd_set_launch(%d, %d, %d): out of range", warps_per_cta, ctas_per_sm, ring_slots);
    g_warps = warps_per_cta;
    g_ctas = ctas_per_sm;
    g_slots = ring_slots;
    return CFD_OK;
}

extern "C" int cfd_set_segments(int chunks_per_segment)
{
    if (chunks_per_segment < 0) return fail(CFD_EINVAL, "chunks_per_segment must be >= 0 (0 = automatic)");
    g_kseg = chunks_per_segment;
    return CFD_OK;
}

// ------------------------------------------------------------------------------------------------
// Kernel launch, optionally as a programmatic dependent launch (kernels.cuh pdl_wait): the streaming kernels, the
// reduced-system kernels and the fused x/y kernel all park at pdl_wait() before their first global access, so a
// launch may be scheduled while the previous kernel of the stream drains.  CFD_PDL=0 turns it off.
// ------------------------------------------------------------------------------------------------
static bool pdl_enabled()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("CFD_PDL"); v = e ? (atoi(e) != 0) : CFD_PDL_DEFAULT; }
    return v != 0;
}

template <class... KArgs, class... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    if (pdl_enabled()) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------------------------------------
// TMA descriptor encoding (driver entry point fetched through the runtime: no link-time libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode()
{
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    }
    return fn;
}

// ------------------------------------------------------------------------------------------------
// geometry shared by both plan kinds
// ------------------------------------------------------------------------------------------------
struct Geometry {
    int nz, ny, nx, axis;
    bool contig;
    int n;        // rows per line
    long inner;   // elements between consecutive rows of a line (1 for x)
    long outer;   // number of outer slices
    long nlines;
    long nb;      // 32-line bundles
    int K, jl;
    int inner_tiles;
};

static int make_geometry(Geometry &g, int nz, int ny, int nx, int axis)
{
    if (nz < 1 || ny < 1 || nx < 1) return fail(CFD_EINVAL, "shape (%d,%d,%d) must be positive", nz, ny, nx);
    if (axis < 0 || axis > 2) return fail(CFD_EINVAL, "axis %d not in {0,1,2}", axis);
    if (nx % 2) return fail(CFD_EINVAL, "nx = %d must be even (TMA needs 16-byte row pitch)", nx);
    g.nz = nz; g.ny = ny; g.nx = nx; g.axis = axis;
    g.contig = (axis == 0);
    if (axis == 0)      { g.n = nx; g.inner = 1;             g.outer = (long)nz * ny; }
    else if (axis == 1) { g.n = ny; g.inner = nx;            g.outer = nz; }
    else                { g.n = nz; g.inner = (long)ny * nx; g.outer = 1; }
    if (g.n < 3) return fail(CFD_EINVAL, "extent along axis %d is %d; need >= 3", axis, g.n);
    g.nlines = (long)nz * ny * nx / g.n;
    g.K = (g.n + CH - 1) / CH;
    g.jl = (g.n - 1) - CH * (g.K - 1);
    if (g.contig) {
        g.inner_tiles = 1;
        g.nb = (g.nlines + CH - 1) / CH;
    } else {
        if (g.inner > 0x7fffffffL) return fail(CFD_EINVAL, "plane too large");
        g.inner_tiles = (int)((g.inner + CH - 1) / CH);
        g.nb = g.outer * g.inner_tiles;
    }
    return CFD_OK;
}

static int encode_maps(const Geometry &g, const void *in, const void *out, CUtensorMap *tm_in, CUtensorMap *tm_out)
{
    encode_tiled_fn enc = get_encode();
    if (!enc) return fail(CFD_ECUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    if (((uintptr_t)in & 15) || ((uintptr_t)out & 15)) return fail(CFD_EINVAL, "field pointers must be 16-byte aligned");
    CUresult r;
    if (g.contig) {
        cuuint64_t dims[2] = {(cuuint64_t)g.nx, (cuuint64_t)g.nlines};
        cuuint64_t strides[1] = {(cuuint64_t)g.nx * 8};
        cuuint32_t box[2] = {16, CH};
        cuuint32_t es[2] = {1, 1};
        r = enc(tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = enc(tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[3] = {(cuuint64_t)g.inner, (cuuint64_t)g.n, (cuuint64_t)g.outer};
        cuuint64_t strides[2] = {(cuuint64_t)g.inner * 8, (cuuint64_t)g.inner * 8 * (cuuint64_t)g.n};
        cuuint32_t box[3] = {CH, CH, 1};
        cuuint32_t es[3] = {1, 1, 1};
        r = enc(tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = enc(tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(CFD_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CFD_OK;
}

// ------------------------------------------------------------------------------------------------
// launch of the streaming kernel
// ------------------------------------------------------------------------------------------------
struct DeviceInfo { int sms = 0; int max_smem = 0; bool ok = false; };

static int device_info(DeviceInfo &d)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int major = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major < 10) return fail(CFD_EUNSUPPORTED, "compute capability %d.x: this library is sm_100a only", major);
    d.ok = true;
    return CFD_OK;
}

// Work counters: {next bundle, finished warps} pairs, one per in-flight launch; each kernel leaves its pair zeroed.
//  * Eager launches draw from a ring of COUNTER_RING pairs per device (allocated by the first plan created on the
//    device): a pair is re-used 4096 launches later, long after its kernel has finished.
//  * A launch that is being CAPTURED into a CUDA graph bakes its pair into the graph and may replay at any later time,
//    concurrently with eager launches or other graphs.  It therefore gets a pair of its own from the owner's PairPool
//    -- never handed out again while the owner (plan) lives -- plus a memset node in front of the kernel node, so the
//    pair needs no history.  Pool chunks are allocated on demand under relaxed capture mode
//    (cudaThreadExchangeStreamCaptureMode), the documented way for a library to call cudaMalloc during a capture.
constexpr int COUNTER_RING = 4096;
constexpr int MAX_DEVICES = 64;
static unsigned long long *g_counters[MAX_DEVICES] = {nullptr};     // one ring per device of this process
static std::atomic<unsigned long> g_launch_seq{0};
static std::mutex g_counter_mu;

static int current_device(int &dev)
{
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return fail(CFD_EUNSUPPORTED, "device ordinal %d out of range", dev);
    return CFD_OK;
}

struct PairPool {
    static constexpr int CHUNK = 256;                 // pairs per chunk (4 KiB)
    std::mutex mu;
    std::vector<unsigned long long *> chunks;
    int used = CHUNK;                                 // pairs taken from the last chunk
    int fresh(unsigned long long **out)
    {
        std::lock_guard<std::mutex> lock(mu);
        if (used == CHUNK) {
            unsigned long long *p = nullptr;
            cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
            cudaThreadExchangeStreamCaptureMode(&mode);
            const cudaError_t e = cudaMalloc(&p, sizeof(unsigned long long) * 2 * CHUNK);
            cudaThreadExchangeStreamCaptureMode(&mode);
            if (e != cudaSuccess) return fail(CFD_ECUDA, "cudaMalloc of work counters: %s", cudaGetErrorString(e));
            chunks.push_back(p);
            used = 0;
        }
        *out = chunks.back() + 2 * used++;
        return CFD_OK;
    }
    ~PairPool() { for (auto *p : chunks) cudaFree(p); }
};
static PairPool g_planless_pool;                      // captured launches of entry points that take no plan

// Called by every plan constructor (never during a capture): the eager ring of the current device.
static int ensure_counters()
{
    int dev = 0;
    int rc = current_device(dev);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_counter_mu);
    if (!g_counters[dev]) {
        unsigned long long *p = nullptr;
        CUDA_TRY(cudaMalloc(&p, sizeof(unsigned long long) * 2 * COUNTER_RING));
        CUDA_TRY(cudaMemset(p, 0, sizeof(unsigned long long) * 2 * COUNTER_RING));
        g_counters[dev] = p;
    }
    return CFD_OK;
}

static int counter_pair(unsigned long long **out, cudaStream_t stream, PairPool *pool = nullptr)
{
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
    if (cs == cudaStreamCaptureStatusActive) {
        int rc = (pool ? pool : &g_planless_pool)->fresh(out);
        if (rc) return rc;
        CUDA_TRY(cudaMemsetAsync(*out, 0, 2 * sizeof(unsigned long long), stream));     // a node of the graph
        return CFD_OK;
    }
    int dev = 0;
    int rc = current_device(dev);
    if (rc) return rc;
    if (!g_counters[dev]) { rc = ensure_counters(); if (rc) return rc; }
    *out = g_counters[dev] + 2 * (g_launch_seq++ % COUNTER_RING);
    return CFD_OK;
}

// ------------------------------------------------------------------------------------------------
// Cross-GPU flag waits: time-out and the host-mapped error word (kernels.cuh wait_flag)
// ------------------------------------------------------------------------------------------------
static std::atomic<long long> g_wait_timeout_ns{120LL * 1000000000LL};
static int *g_err_host = nullptr, *g_err_dev = nullptr;
static std::mutex g_err_mu;

static int ensure_err_word()        // plan constructors of partitioned plans call this (never during a capture)
{
    std::lock_guard<std::mutex> lock(g_err_mu);
    if (g_err_host) return CFD_OK;
    int *h = nullptr, *d = nullptr;
    CUDA_TRY(cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    *h = 0;
    CUDA_TRY(cudaHostGetDevicePointer(&d, h, 0));
    g_err_host = h;
    g_err_dev = d;
    return CFD_OK;
}

static WaitP wait_params()
{
    WaitP w;
    w.timeout_ns = (unsigned long long)g_wait_timeout_ns.load();
    w.err = g_err_dev;
    return w;
}

extern "C" int cfd_set_wait_timeout_ms(long ms)
{
    if (ms < 1) return fail(CFD_EINVAL, "time-out must be >= 1 ms");
    g_wait_timeout_ns = (long long)ms * 1000000LL;
    return CFD_OK;
}

extern "C" int cfd_async_status(void)
{
    if (g_err_host && *(volatile int *)g_err_host != 0) {
        *(volatile int *)g_err_host = 0;
        return fail(CFD_ETIMEOUT, "a kernel gave up waiting for a neighbour's arrival flag (time-out %lld ms): results "
                                  "of that call are invalid", g_wait_timeout_ns.load() / 1000000LL);
    }
    return CFD_OK;
}

template <bool CONTIG, bool DERIV, int NSLOT>
static int launch_stream_ns(const Geometry &g, KParams kp, const CUtensorMap &tm_in, const CUtensorMap &tm_out,
                            cudaStream_t stream, bool in_place, PairPool *pool, int force_kseg = 0,
                            const CUtensorMap *tm_aux = nullptr)
{
    static DeviceInfo dinfo;
    if (!dinfo.ok) { int rc = device_info(dinfo); if (rc) return rc; }
    constexpr int per_warp = (NSLOT + 2) * SLOT_BYTES + NSLOT * 16;
    // Measured on B200 at 512^3 (scripts/sweep_launch.py, profiles/r1_sweep_launch_512.txt): 4 warps per SM is
    // fastest for both layouts (x 0.337, y 0.333, z 0.341 ms); 5-7 warps cost 1-2 %, 3 warps 8-13 %.
    // Shared memory (40 KiB per warp: 3 ring slots + 2 staging slots) caps a CTA at 5 warps.
    const int max_warps = 5;
    int warps = g_warps ? g_warps : 4;
    if (warps > max_warps) warps = max_warps;
    int ctas = g_ctas ? g_ctas : 1;
    // Long lines, few bundles (e.g. 1024 systems of 4096 unknowns = 32 bundles): cut the lines into segments of
    // >= 8 chunks so that there are ~4 work items per warp; each cut costs two extra tiles of reads (look-ahead).
    kp.kseg = g.K; kp.nseg = 1;
    {
        // In place (solver-only API) a segment's warm-up / look-ahead tiles belong to its neighbours' OUTPUT and
        // may already have been overwritten: segmentation is for the out-of-place derivative only.
        int kseg = in_place ? 0 : g_kseg;
        if (!in_place && !kseg && g.K >= 16 && g.nb < 2L * dinfo.sms * warps) {
            long want = (4L * dinfo.sms * warps + g.nb - 1) / g.nb;      // segments per line we would like
            if (want > g.K / 8) want = g.K / 8;
            if (want > 1) kseg = (int)((g.K + want - 1) / want);
        }
        // Contiguous lines of >= 1024 rows with fewer than ~8 bundles per warp: the dynamic draw gets coarse (a
        // warp's last bundle is 1/4..1/8 of its work).  Segments of >= 16 chunks shorten the tail; measured on
        // [128,1024,1024]: d/dx 0.362 -> 0.348 ms (strided lines did not gain and keep whole lines).
        if (CONTIG && !in_place && !kseg && g.K >= 32 && g.nb < 8L * dinfo.sms * warps) {
            long want = (8L * dinfo.sms * warps + g.nb - 1) / g.nb;
            if (want > g.K / 16) want = g.K / 16;
            if (want > 1) kseg = (int)((g.K + want - 1) / want);
        }
        if (force_kseg > 0) kseg = force_kseg;       // in place with the side buffer: the plan fixed the cut
        if (kseg > 0 && kseg < g.K) { kp.kseg = kseg; kp.nseg = (g.K + kseg - 1) / kseg; }
        kp.aux_on = (tm_aux != nullptr && kp.nseg > 1) ? 1 : 0;
    }
    const long nitems = g.nb * kp.nseg;
    if (!g_warps) {   // small problems: spread the work items over all SMs before stacking warps on one
        const long per_sm = (nitems + dinfo.sms - 1) / dinfo.sms;
        if (per_sm < warps) warps = (int)(per_sm < 1 ? 1 : per_sm);
    }
    auto smem_for = [&](int w) { return (size_t)w * per_warp + 1024; };
    while (warps > 1 && ((long)ctas * (long)(smem_for(warps) + 1024) > 233472L || (long)ctas * warps > 8)) warps--;
    const size_t smem = smem_for(warps);
    if (smem > 232448) return fail(CFD_EUNSUPPORTED, "ring of %d slots does not fit shared memory", NSLOT);
    auto kern = stream_kernel<CONTIG, DERIV, NSLOT>;
    static size_t configured[MAX_DEVICES] = {0};
    int dev = 0;
    { int rc = current_device(dev); if (rc) return rc; }
    if (configured[dev] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    int rc = counter_pair(&kp.counter, stream, pool);
    if (rc) return rc;
    long blocks = (nitems + warps - 1) / warps;
    const long cap = (long)dinfo.sms * ctas;
    if (blocks > cap) blocks = cap;
    if (const char *e = getenv("CFD_CTAS")) { if (atol(e) >= 1 && atol(e) < blocks) blocks = atol(e); }    // experiment
    CUDA_TRY(launch_k(kern, (unsigned)blocks, warps * 32, smem, stream, tm_in, tm_out, kp, tm_aux ? *tm_aux : tm_out));
    g_launches++;
    return CFD_OK;
}

template <bool CONTIG, bool DERIV>
static int launch_stream(const Geometry &g, const KParams &kp, const CUtensorMap &tm_in, const CUtensorMap &tm_out,
                         cudaStream_t stream, PairPool *pool, bool in_place = false, int force_kseg = 0,
                         const CUtensorMap *tm_aux = nullptr)
{
    switch (g_slots ? g_slots : 3) {
        case 3: return launch_stream_ns<CONTIG, DERIV, 3>(g, kp, tm_in, tm_out, stream, in_place, pool, force_kseg, tm_aux);
        case 4: return launch_stream_ns<CONTIG, DERIV, 4>(g, kp, tm_in, tm_out, stream, in_place, pool, force_kseg, tm_aux);
        case 5: return launch_stream_ns<CONTIG, DERIV, 5>(g, kp, tm_in, tm_out, stream, in_place, pool, force_kseg, tm_aux);
        default: return fail(CFD_EINVAL, "ring slots must be 3, 4 or 5");
    }
}

// ------------------------------------------------------------------------------------------------
// plans
// ------------------------------------------------------------------------------------------------
// Last encoded descriptor pair of a plan.  Guarded by a mutex and handed out BY VALUE, so concurrent calls on one
// plan from several host threads (distinct streams, distinct fields) never see a half-updated descriptor.
struct MapCache { std::mutex mu; const void *in = nullptr, *out = nullptr; CUtensorMap tm_in, tm_out; };
struct MapPair { CUtensorMap tm_in, tm_out; };

struct cfd_plan {
    Geometry g;
    double h;
    int rank, size;
    KParams kp;                       // host template (halo/out pointers filled per call)
    MapCache cache;
    PairPool pool;                    // work counters of this plan's CAPTURED launches (see counter_pair)
    int scheme = 0;                   // CFD_SCHEME_*; schemes other than PADE4 run the general kernel
    int la = 1;                       // look-ahead chunks of the general kernel
    bool second = false;              // the scheme is a second derivative (symmetric stencil)
    bool npts = false;                // tables are this rank's slice of the GLOBAL LU (cfd_create_npts)
    GParams gp;
    // multi-rank
    std::vector<double> x_uh, x_lh, ra, rb, rc, lu;
    std::vector<double> lu_nb;        // neighbour-only reduced system (ranks r-1, r, r+1)
    int nb_pv = 0, nb_own = 0;
    double *d_x_uh = nullptr, *d_x_lh = nullptr, *d_lu = nullptr, *d_lu_nb = nullptr;
    int wc = 0;
    // host staging
    double *d_f = nullptr, *d_df = nullptr;
    cudaStream_t hstream = nullptr, hcopy = nullptr, hback = nullptr;
    cudaEvent_t hev[8] = {nullptr};
    // cfd_apply_xy: draw order of the (plane, bundle) items (axis-0 plans only)
    std::mutex xy_mu;
    int *d_xy_order = nullptr;
    long xy_entries = 0;              // entries of the table (> items when lines are cut into segments)
    int xy_kseg = 0, xy_ksegy = 0, xy_sub = -1;     // segment lengths of the x / y lines in the draw table (0 = whole)
    long xy_nedge = 0;                // edge entries in front of the table (drawn by cfd_zpart_apply_xyz only)
    double xy_active = -1.0;
    int xy_warps = 0;                 // 0 = default; cfd_plan_set_xy_warps
    double w_lo = 0.0, w_hi = 0.0;    // d(lo face)/d f[-1], d(hi face)/d f[n]: cfd_reduced_unknowns_deferred
    // cfd_apply_xyz (axis-2 plans): side stream of the d/dz launch and the fork / join events
    cudaStream_t gside = nullptr;
    cudaEvent_t gev_fork = nullptr, gev_join = nullptr;
};

struct nt_plan {
    Geometry g;
    KParams kp;
    MapCache cache;
    PairPool pool;
    bool exact = false;               // two-pass exact solver (matrix refused by the one-pass kernels)
    int la = 0;                       // > 0: the general one-pass kernel with this many look-ahead chunks
    GParams gp;
    double *d_tab = nullptr;          // [4][K*32]: forward pc, qc; backward pc, qc
    // Starved shape (long lines, few bundles): the lines are cut into segments so that every warp finds work.  In
    // place that is only safe if no segment's result lands on a tile a neighbouring segment still has to read: the
    // first and last result chunk of every segment go to this side buffer (2 chunks per segment, <= 1/4 of the field)
    // and are copied into place by a second, small launch.
    double *d_aux = nullptr;
    int aux_kseg = 0, aux_nseg = 0;
    Geometry g_aux;
    MapCache aux_cache;
};

#define CFD_ESLOWPATH (-100)   /* internal: matrix needs the exact two-pass solver */

static int fill_tables(KParams &kp, const Geometry &g, const LineCoeffs &m, double scale, bool need_fast)
{
    Pivots pv = build_pivots(g.n, m);
    if (!pv.finite) return fail(CFD_EINVAL, "zero pivot: LU without pivoting breaks down for these coefficients");
    if (g.K > 2 && need_fast) {
        if (!pv.converged || std::pow(pv.decay, CH) > 1.2e-16)
            return fail(CFD_ESLOWPATH,
                        "coefficients (ai,bi,ci) = (%g,%g,%g): LU pivots do not converge / coupling %.3f^32 above fp64 "
                        "round-off; the streaming solver needs a diagonally dominant interior for n > 64",
                        m.ai, m.bi, m.ci, pv.decay);
    }
    memset(&kp, 0, sizeof kp);
    kp.n = g.n; kp.K = g.K; kp.jl = g.jl; kp.kseg = g.K; kp.nseg = 1;
    kp.inner = (int)g.inner; kp.inner_tiles = g.inner_tiles; kp.outer = (int)g.outer;
    kp.nb = g.nb; kp.rows = g.nlines;
    kp.head = chunk_table(pv, g.n, 0, scale);
    kp.tail = chunk_table(pv, g.n, g.K - 1, scale);
    if (g.K > 2) { kp.sk_mid = pv.beta[CH] * scale; kp.l_mid = pv.l[CH]; kp.g_mid = pv.g[CH]; }
    kp.s0c = pv.beta[0];        // scaled by the caller for the derivative
    kp.snc = pv.beta[g.n - 1];
    return CFD_OK;
}

static LineCoeffs pade_block(int rank, int size)
{   // code/cuda/compact.py:159-166
    LineCoeffs m = {1.0, 0.25, 0.25, 1.0, 0.25, 0.25, 1.0};
    if (rank == 0) m.c1 = 2.0;
    if (rank == size - 1) m.an = 2.0;
    return m;
}

static void secondary_systems(int n, int rank, int size, std::vector<double> &x_uh, std::vector<double> &x_lh)
{   // code/cuda/compact.py:128-154
    std::vector<double> a(n, 0.25), b(n, 1.0), c(n, 0.25);
    if (rank == 0) { c[0] = 2.0; a[0] = 0.0; }
    if (rank == size - 1) { a[n - 1] = 2.0; c[n - 1] = 0.0; }
    x_uh.assign(n, 0.0); x_lh.assign(n, 0.0);
    x_uh[0] = -a[0];
    x_lh[n - 1] = -c[n - 1];
    thomas_host(n, a, b, c, x_uh);
    thomas_host(n, a, b, c, x_lh);
}

// Reduced (interface) matrix of the whole line: every rank can build all rows, they depend on the block
// length and position only (the reference gathers two scalars per rank instead, compact.py:77-82).
static void reduced_matrix(int n, int P, std::vector<double> &ra, std::vector<double> &rb, std::vector<double> &rc)
{
    const int mm = 2 * P;
    std::vector<double> uh(mm), lh(mm);
    for (int r = 0; r < P; r++) {
        std::vector<double> u, l;
        secondary_systems(n, r, P, u, l);
        uh[2 * r] = u[0]; uh[2 * r + 1] = u[n - 1];
        lh[2 * r] = l[0]; lh[2 * r + 1] = l[n - 1];
    }
    ra.assign(mm, 0.0); rb.assign(mm, 0.0); rc.assign(mm, 0.0);
    for (int i = 0; i < mm; i += 2) { ra[i] = -1.0; rb[i] = uh[i]; rc[i] = lh[i]; }     // compact.py:100-105
    for (int i = 1; i < mm; i += 2) { ra[i] = uh[i]; rb[i] = lh[i]; rc[i] = -1.0; }
    ra[0] = 0.0; rc[0] = 0.0; rb[0] = 1.0;                                                // :106-107
    ra[mm - 1] = 0.0; rc[mm - 1] = 0.0; rb[mm - 1] = 1.0;                                 // :108-109
    ra[1] = 0.0; rc[mm - 2] = 0.0;                                                        // :110-111
}

// Two-sided elimination tables of a tridiagonal (a, b, c): [6][m] = a_i, c_i, 1/p_i, c_i/p_i, 1/q_i, a_i/q_i
// (p: pivots of the top-down sweep, q: pivots of the bottom-up sweep).
static std::vector<double> elimination_table(const std::vector<double> &ra, const std::vector<double> &rb,
                                             const std::vector<double> &rc)
{
    const int mm = (int)ra.size();
    std::vector<double> lu(6 * mm, 0.0);
    double *a = &lu[0], *c = &lu[mm], *ip = &lu[2 * mm], *cp = &lu[3 * mm], *iq = &lu[4 * mm], *aq = &lu[5 * mm];
    for (int i = 0; i < mm; i++) { a[i] = ra[i]; c[i] = rc[i]; }
    double piv = rb[0];
    ip[0] = 1.0 / piv; cp[0] = c[0] / piv;
    for (int i = 1; i < mm; i++) { piv = rb[i] - a[i] * cp[i - 1]; ip[i] = 1.0 / piv; cp[i] = c[i] / piv; }
    piv = rb[mm - 1];
    iq[mm - 1] = 1.0 / piv; aq[mm - 1] = a[mm - 1] / piv;
    for (int i = mm - 2; i >= 0; i--) { piv = rb[i] - c[i] * aq[i + 1]; iq[i] = 1.0 / piv; aq[i] = a[i] / piv; }
    return lu;
}

// Neighbour-only ("pairwise") reduced system of rank r: the unknowns of ranks r-1, r, r+1 only.  For blocks of
// >= 64 rows the couplings x_UH[-1], x_LH[0] that tie an interface to the next one are ~0.268^64 = 1e-37, so
// cutting the chain at the neighbours (identity rows, as at a physical end) changes nothing in fp64 and a rank
// needs ONE interface plane from each neighbour instead of the all-gathered 2P planes.
//   pv = number of virtual ranks (2 or 3), own = index of this rank among them.
static void neighbour_matrix(int n, int rank, int P, std::vector<double> &va, std::vector<double> &vb,
                             std::vector<double> &vc, int &pv, int &own)
{
    std::vector<double> ra, rb, rc;
    reduced_matrix(n, P, ra, rb, rc);
    const int lo = rank > 0 ? rank - 1 : rank, hi = rank < P - 1 ? rank + 1 : rank;
    pv = hi - lo + 1;
    own = rank - lo;
    const int mm = 2 * pv;
    va.assign(mm, 0.0); vb.assign(mm, 0.0); vc.assign(mm, 0.0);
    for (int i = 0; i < mm; i++) { va[i] = ra[2 * lo + i]; vb[i] = rb[2 * lo + i]; vc[i] = rc[2 * lo + i]; }
    va[0] = 0.0; vc[0] = 0.0; vb[0] = 1.0;                    // outer unknowns of the outermost virtual ranks:
    va[mm - 1] = 0.0; vc[mm - 1] = 0.0; vb[mm - 1] = 1.0;     // identity rows with zero right-hand side
    va[1] = 0.0; vc[mm - 2] = 0.0;
}

// Host-only inspection of the multi-rank tables (no device needed).
extern "C" int cfd_debug_secondary(int n, int part_rank, int part_size, double *x_uh, double *x_lh, double *ra,
                                   double *rb, double *rc)
{
    if (n < 3 || part_size < 2 || part_rank < 0 || part_rank >= part_size) return fail(CFD_EINVAL, "bad argument");
    std::vector<double> u, l, a, b, c;
    secondary_systems(n, part_rank, part_size, u, l);
    reduced_matrix(n, part_size, a, b, c);
    if (x_uh) memcpy(x_uh, u.data(), n * sizeof(double));
    if (x_lh) memcpy(x_lh, l.data(), n * sizeof(double));
    if (ra) memcpy(ra, a.data(), a.size() * sizeof(double));
    if (rb) memcpy(rb, b.data(), b.size() * sizeof(double));
    if (rc) memcpy(rc, c.data(), c.size() * sizeof(double));
    return CFD_OK;
}

extern "C" int cfd_debug_neighbour(int n, int part_rank, int part_size, int *virtual_ranks, int *own_index,
                                   double *va, double *vb, double *vc)
{
    if (n < 3 || part_size < 2 || part_rank < 0 || part_rank >= part_size) return fail(CFD_EINVAL, "bad argument");
    std::vector<double> a, b, c;
    int pv = 0, own = 0;
    neighbour_matrix(n, part_rank, part_size, a, b, c, pv, own);
    if (virtual_ranks) *virtual_ranks = pv;
    if (own_index) *own_index = own;
    if (va) memcpy(va, a.data(), a.size() * sizeof(double));
    if (vb) memcpy(vb, b.data(), b.size() * sizeof(double));
    if (vc) memcpy(vc, c.data(), c.size() * sizeof(double));
    return CFD_OK;
}

// The interface faces depend linearly on the neighbour points of f (edge_faces_kernel): lo face = -x_0 of the head
// solve, where f[-1] enters row 0 only (r_0 = sk_0 (f[1] - f[-1])); hi face = -e_{n-1}, where f[n] enters through
// sk_last.  Unit responses, for the exchange that folds the halos in afterwards (cfd_reduced_unknowns_deferred).  They
// are applied to the NEIGHBOURS' faces too, so they come from the tables of an interior block (a block end that has a
// neighbour is an interior-type end on every rank), not from this rank's own closures.
static int halo_weights(const Geometry &g, double h, double &w_lo, double &w_hi)
{
    KParams ki;
    int rc = fill_tables(ki, g, pade_block(1, 3), 3.0 / (4.0 * h), true);
    if (rc) return rc == CFD_ESLOWPATH ? CFD_EUNSUPPORTED : rc;
    double e[CH], ep = 0.0;                 // the same 31 rows as edge_faces_kernel
    for (int j = 0; j < CH - 1; j++) { ep = -ki.head.l[j] * ep + (j == 0 ? -ki.head.sk[0] : 0.0); e[j] = ep; }
    double x = 0.0;
    for (int j = CH - 2; j >= 0; j--) x = e[j] - ki.head.g[j] * x;
    w_lo = -x;
    w_hi = -ki.tail.sk[g.jl];
    return CFD_OK;
}

// Host-only inspection entries for the CPU tests (no device needed).
extern "C" int cfd_debug_halo_weights(int n, double h, double *w_lo, double *w_hi)
{
    if (!w_lo || !w_hi || n < 2 * CH + 2 || !(h > 0.0)) return fail(CFD_EINVAL, "cfd_debug_halo_weights: bad argument");
    Geometry g;
    int rc = make_geometry(g, n, 2, 2, 2);
    if (rc) return rc;
    return halo_weights(g, h, *w_lo, *w_hi);
}

// d/dx and d/dy of one field in one launch (kernels_xy.cuh).  A plane of f is a grid of 32 x 32 tiles (j, k):
// x-bundle j walks tiles (j, 0), (j, 1), ... and y-bundle k walks (0, k), (1, k), ...  When x-bundle j and y-bundle j
// both start j tile-times after their plane's first bundles, the two readers of EVERY tile (j, k) ask for it at the
// same moment, j + k tile-times in: one of them brings it from HBM, the other finds it in L2.  The draw order below
// produces that wavefront with dynamically scheduled warps: per slot (= one tile-time) every active plane
// contributes its next (x, y) bundle pair, and `active` planes are in flight so that a slot's worth of items is
// what all resident warps draw in one tile-time (active = warps / (2 K)).
// Sub-plane wavefronts (sub > 0): on long lines the data in flight between the two readers of a tile outgrows L2
// (it is proportional to the tiles per line), so a plane is cut into sub x sub-tile squares and the lines into
// segments of `sub` chunks (one warm-up chunk in front, the look-ahead chunk behind, both served by L2); every square
// is then a wavefront of its own, as short as a small plane's.  Entries are (item << 3) | segment, segment 0 = whole
// line.  Returns the effective segment length in `kseg` (0 = nothing was cut).
// `sub` = segment length of the x lines | segment length of the y lines << 8 (a zero high byte = the same for both;
// a length >= the line's tile count leaves that direction whole); ksegx / ksegy return what was cut (0 = whole lines).
static std::vector<int> xy_order(int nz, int nxp, int nyp, double active, int sub, int &ksegx, int &ksegy, long nedge = 0)
{
    const int ipp = nxp + nyp;
    ksegx = ksegy = 0;
    if (sub > 0) {
        int subx = sub & 255, suby = (sub >> 8) ? (sub >> 8) : (sub & 255);
        if (subx * 7 < nyp) subx = (nyp + 6) / 7;             // at most 7 segments fit the 3-bit code
        if (suby * 7 < nxp) suby = (nxp + 6) / 7;
        if (nyp > subx) ksegx = subx;                          // an x line is nyp tiles long
        if (nxp > suby) ksegy = suby;                          // a y line is nxp tiles long
    }
    const int ngx = ksegx ? (nyp + ksegx - 1) / ksegx : 1;     // segments of an x line
    const int ngy = ksegy ? (nxp + ksegy - 1) / ksegy : 1;     // segments of a y line
    const int sy = ngy > 1 ? ksegy : nxp, sx = ngx > 1 ? ksegx : nyp;       // x bundles / y bundles per rectangle
    std::vector<int> order;
    order.reserve((size_t)nz * (nxp * ngx + nyp * ngy) + (size_t)nedge);
    // edge items of a z-partitioned d/dz (ids from nz * ipp on) go FIRST: their faces have the whole launch to travel
    for (long t = 0; t < nedge; t++) order.push_back((int)(((long)nz * ipp + t) << 3));
    if (active <= 0.0) {                       // plain plane-by-plane order (whole lines)
        ksegx = ksegy = 0;
        for (long w = 0; w < (long)nz * ipp; w++) order.push_back((int)(w << 3));
        return order;
    }
    const long nv = (long)nz * ngy * ngx;      // squares ("virtual planes"), z-major, then y block, then x block
    const int M = sy > sx ? sy : sx;
    const double sigma = (double)M / active;   // slots between the starts of consecutive squares
    // Skew: rectangle (a, b) of a plane starts a * sy + b * sx slots after (0, 0), so that the warm-up / look-ahead
    // tile a segment shares with the neighbouring square is asked for when that square's own readers ask for it (the
    // plane-wide wavefront, tile (j, k) at j + k, kept although the lines are cut).  Without it the squares of a plane
    // start together and the shared tiles are read M tile-times apart: one of the two reads goes to DRAM.
    static const bool skew = !(getenv("CFD_XY_SKEW") && atoi(getenv("CFD_XY_SKEW")) == 0);
    auto emit = [&](long v, long j) {
        const int z = (int)(v / (ngy * ngx)), a = (int)((v / ngx) % ngy), b = (int)(v % ngx);
        const long xj = (long)a * sy + j, yk = (long)b * sx + j;
        if (j < sy && xj < nxp) order.push_back((int)((((long)z * ipp + xj) << 3) | (ngx > 1 ? b + 1 : 0)));
        if (j < sx && yk < nyp) order.push_back((int)((((long)z * ipp + nxp + yk) << 3) | (ngy > 1 ? a + 1 : 0)));
    };
    if (skew && ngx * ngy > 1) {
        std::vector<std::pair<long, long>> starts;             // (start slot, square)
        starts.reserve((size_t)nv);
        for (long v = 0; v < nv; v++) {
            const long z = v / (ngy * ngx), a = (v / ngx) % ngy, b = v % ngx;
            starts.emplace_back((long)std::floor((double)(z * ngy * ngx) * sigma) + a * sy + b * sx, v);
        }
        std::stable_sort(starts.begin(), starts.end());
        size_t lo = 0;
        for (long s = 0; lo < starts.size(); s++) {
            for (size_t i = lo; i < starts.size() && starts[i].first <= s; i++) {
                const long j = s - starts[i].first;
                if (j >= M) { if (i == lo) lo++; continue; }
                emit(starts[i].second, j);
            }
        }
        return order;
    }
    auto start = [&](long v) { return (long)std::floor(v * sigma); };
    long vlo = 0;
    for (long s = 0; vlo < nv; s++) {
        for (long v = vlo; v < nv && start(v) <= s; v++) {
            const long j = s - start(v);
            if (j >= M) { if (v == vlo) vlo++; continue; }
            emit(v, j);
        }
    }
    return order;
}

extern "C" long cfd_debug_xy_order(int nz, int nxp, int nyp, double active, int sub, int *out, long capacity)
{
    if (nz < 1 || nxp < 0 || nyp < 0 || nxp + nyp < 1 || sub < 0) return fail(CFD_EINVAL, "cfd_debug_xy_order: bad argument");
    int ksegx = 0, ksegy = 0;
    const std::vector<int> order = xy_order(nz, nxp, nyp, active, sub, ksegx, ksegy);
    if (out) {
        if ((long)order.size() > capacity) return fail(CFD_EINVAL, "cfd_debug_xy_order: %ld entries, capacity %ld", (long)order.size(), capacity);
        memcpy(out, order.data(), order.size() * sizeof(int));
    }
    return (long)order.size();
}

// Default launch shape of stream_kernel_xy and the planes-in-flight figure that goes with it.
static void xy_shape(const Geometry &gx, int sms, int nslot, int plan_warps, int &warps, double &active, int &sub)
{
    const int max_warps = nslot == 3 ? 8 : 7, def_warps = nslot == 3 ? 8 : 6;
    int Kx = gx.K, Ky = (gx.ny + CH - 1) / CH;
    // Sub-plane wavefronts for lines of >= 32 tiles ([128,1024,1024]: 0.578 -> 0.561 ms with squares of 16 x 16 tiles;
    // 512^3 with 8 x 8: 0.541 -> 0.582, so shorter lines stay whole).  CFD_XY_SUB overrides (0 = never cut).
    // Rectangles instead of squares (round 2): x lines cut at 16 tiles, y lines only beyond 32 -- every cut line walks
    // two extra tiles per segment, and with the skewed starts of xy_order the wavefront no longer needs both
    // directions cut: 128 / 256 / 512 planes of 1024^2: -1.5 / -4.5 / -2.4 % (profiles/r2s_xy_rectangles.txt).
    sub = (Kx >= 32 || Ky >= 32) ? (16 | (32 << 8)) : 0;
    if (const char *e = getenv("CFD_XY_SUB")) sub = atoi(e) > 0 ? atoi(e) : 0;
    if (sub > 0) {                                             // tiles an item walks when its line is cut
        int subx = sub & 255, suby = (sub >> 8) ? (sub >> 8) : (sub & 255);
        if (subx * 7 < Kx) subx = (Kx + 6) / 7;
        if (suby * 7 < Ky) suby = (Ky + 6) / 7;
        if (Kx > subx) Kx = subx + 2;
        if (Ky > suby) Ky = suby + 2;
    }
    const long nitems = (long)gx.nz * (gx.ny / CH + (gx.nx + CH - 1) / CH);
    warps = g_warps ? g_warps : (plan_warps ? plan_warps : def_warps);
    // Launches of few items per warp take a seventh warp per SM: the launch runs in about items / warps generations of
    // items and ends with a partly filled one, so finer generations pay (256^3: 0.0813 -> 0.0760 ms, [64,512,512]:
    // 0.0951 -> 0.0866; never slower than 6 warps up to 384^3), while long launches are better off with 6 warps and
    // fewer concurrent DRAM streams (512^3: 0.533 vs 0.545 ms).  profiles/r2y_sweep_xy_warps.txt.
    if (!g_warps && !plan_warps && nslot == 4 && sub == 0 && nitems <= 72L * sms) warps = 7;
    if (warps > max_warps) warps = max_warps;
    const long per_sm = (nitems + sms - 1) / sms;
    if (!g_warps && per_sm < warps) warps = (int)(per_sm < 1 ? 1 : per_sm);
    active = (double)sms * warps / (Kx + Ky);          // = warps / (2 K) for square planes
    if (const char *e = getenv("CFD_XY_ACTIVE")) active = atof(e);      // 0 = plain plane-by-plane order
    if (active > 0.0 && active < 1.0) active = 1.0;
}

// Measurement yardstick (bench.py `roofline.yardstick`, scripts/dram_mix.py): a plain grid-stride streaming kernel with
// the read : write mix of the library's launches -- c == nullptr: copy (8 B read + 8 B written per element, the mix of a
// single derivative); otherwise one read and two writes (the mix of the fused d/dx + d/dy launch).  DRAM streams the
// 1 : 2 mix about 7 % below the 1 : 1 copy figure (profiles/r2u_dram_mix_yardstick.txt), which is what a roofline
// fraction against the copy bandwidth cannot show.  No part of any derivative path.
__global__ void __launch_bounds__(256) stream_yardstick_kernel(const double2 *__restrict__ a, double2 *__restrict__ b,
                                                               double2 *__restrict__ c, long n2)
{
    const long stride = (long)gridDim.x * blockDim.x;
    if (c) {
        for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n2; i += stride) {
            const double2 v = a[i];
            b[i] = v;
            c[i] = make_double2(v.y, v.x);
        }
    } else {
        for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n2; i += stride) b[i] = a[i];
    }
}

extern "C" int cfd_debug_stream(const double *a, double *b, double *c, long n, void *stream)
{
    if (!a || !b || n < 2 || (n & 1) || ((uintptr_t)a & 15) || ((uintptr_t)b & 15) || ((uintptr_t)c & 15))
        return fail(CFD_EINVAL, "cfd_debug_stream: needs 16-byte aligned arrays of an even number of doubles");
    static DeviceInfo dinfo;
    if (!dinfo.ok) { int rc = device_info(dinfo); if (rc) return rc; }
    stream_yardstick_kernel<<<16 * dinfo.sms, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2 *>(a), reinterpret_cast<double2 *>(b), reinterpret_cast<double2 *>(c), n / 2);
    CUDA_TRY(cudaGetLastError());
    return CFD_OK;
}

// Host-only inspection entry for the CPU tests: the launch shape cfd_apply_xy would pick on a device of `sms` SMs.
extern "C" int cfd_debug_xy_shape(int nz, int ny, int nx, int sms, int *warps, double *active, int *sub)
{
    if (!warps || !active || !sub || sms < 1) return fail(CFD_EINVAL, "cfd_debug_xy_shape: bad argument");
    Geometry g;
    int rc = make_geometry(g, nz, ny, nx, 0);
    if (rc) return rc;
    xy_shape(g, sms, 4, 0, *warps, *active, *sub);
    return CFD_OK;
}

// (Re)build the draw-order table of an axis-0 plan.  Called from cfd_create with the default launch shape, so that
// cfd_apply_xy neither allocates nor synchronises; it only runs again if the launch knobs were changed afterwards.
static int xy_prepare(cfd_plan *px, double active, int sub)
{
    // The table always starts with the edge entries of the plane (one per 32 columns): a plain cfd_apply_xy launch
    // skips them (order + xy_nedge), the fused zpart launch draws them first -- so neither ever rebuilds the table.
    const long nedge = ((long)px->g.ny * px->g.nx + CH - 1) / CH;
    std::lock_guard<std::mutex> lock(px->xy_mu);
    if (px->d_xy_order && px->xy_active == active && px->xy_sub == sub) return CFD_OK;
    const int nxp = px->g.ny / CH, nyp = (px->g.nx + CH - 1) / CH;
    int kseg = 0, ksegy = 0;
    const std::vector<int> order = xy_order(px->g.nz, nxp, nyp, active, sub, kseg, ksegy, nedge);
    if ((long)order.size() < (long)px->g.nz * (nxp + nyp) + nedge) return fail(CFD_EINVAL, "internal: xy draw order is incomplete");
    if (px->d_xy_order && (long)order.size() > px->xy_entries) { cudaFree(px->d_xy_order); px->d_xy_order = nullptr; }
    if (!px->d_xy_order) CUDA_TRY(cudaMalloc(&px->d_xy_order, order.size() * sizeof(int)));
    CUDA_TRY(cudaMemcpy(px->d_xy_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
    px->xy_entries = (long)order.size();
    px->xy_kseg = kseg;
    px->xy_ksegy = ksegy;
    px->xy_active = active;
    px->xy_sub = sub;
    px->xy_nedge = nedge;
    return CFD_OK;
}

static bool xy_eligible(const Geometry &g)
{   // id << 3 must fit an int (edge items of a fused zpart launch included: one per 32 columns of the plane)
    return g.axis == 0 && g.ny % CH == 0 &&
           (long)g.nz * (g.ny / CH + (g.nx + CH - 1) / CH) + ((long)g.ny * g.nx + CH - 1) / CH <= 0x0fffffffL;
}

extern "C" int cfd_create(cfd_plan **out, int nz, int ny, int nx, int axis, double h, int part_rank, int part_size)
{
    if (!out) return fail(CFD_EINVAL, "plan pointer is NULL");
    *out = nullptr;
    if (!(h > 0.0) || !std::isfinite(h)) return fail(CFD_EINVAL, "spacing h = %g must be positive", h);
    if (part_size < 1 || part_rank < 0 || part_rank >= part_size || part_size > 64)
        return fail(CFD_EINVAL, "partition position %d of %d is invalid", part_rank, part_size);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(CFD_ECUDA, "no CUDA device: libcfd_b200 has no CPU path");
    }
    cfd_plan *p = new cfd_plan();
    int rc = make_geometry(p->g, nz, ny, nx, axis);
    if (!rc) rc = ensure_counters();
    if (!rc && part_size > 1) rc = ensure_err_word();
    if (rc) { delete p; return rc; }
    p->h = h; p->rank = part_rank; p->size = part_size;
    const LineCoeffs m = pade_block(part_rank, part_size);
    rc = fill_tables(p->kp, p->g, m, 3.0 / (4.0 * h), true);
    if (rc) { delete p; return rc == CFD_ESLOWPATH ? CFD_EUNSUPPORTED : rc; }
    p->kp.lo_closure = (part_rank == 0);
    p->kp.hi_closure = (part_rank == part_size - 1);
    // coupled multi-rank solve: known neighbour unknowns enter rows 0 / n-1 as Dirichlet data
    if (part_rank > 0) p->kp.head.l[0] = m.ai * p->kp.s0c;               // a_i * beta_0 (eprev starts at alpha)
    p->kp.snb = (part_rank < part_size - 1) ? p->kp.snc * m.ci : 0.0;     // beta_{n-1} * c_i
    p->kp.s0c *= 1.0 / (2.0 * h);
    p->kp.snc *= 1.0 / (2.0 * h);

    if (part_size > 1) {
        const int n = p->g.n, P = part_size;
        secondary_systems(n, part_rank, P, p->x_uh, p->x_lh);
        reduced_matrix(n, P, p->ra, p->rb, p->rc);
        p->lu = elimination_table(p->ra, p->rb, p->rc);
        {
            std::vector<double> va, vb, vc;
            neighbour_matrix(n, part_rank, P, va, vb, vc, p->nb_pv, p->nb_own);
            p->lu_nb = elimination_table(va, vb, vc);
        }
        // correction band: rows where the secondary solutions are above round-off
        int wc = 0;
        while (wc < n && (std::fabs(p->x_uh[wc]) > 1e-19 || std::fabs(p->x_lh[n - 1 - wc]) > 1e-19)) wc++;
        p->wc = wc;
        rc = halo_weights(p->g, h, p->w_lo, p->w_hi);
        if (rc) { cfd_destroy(p); return rc; }
        if (cudaMalloc(&p->d_x_uh, n * sizeof(double)) != cudaSuccess || cudaMalloc(&p->d_x_lh, n * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&p->d_lu, p->lu.size() * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&p->d_lu_nb, p->lu_nb.size() * sizeof(double)) != cudaSuccess) {
            cfd_destroy(p);
            return fail(CFD_ECUDA, "cudaMalloc of plan tables failed");
        }
        if (cudaMemcpy(p->d_x_uh, p->x_uh.data(), n * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(p->d_x_lh, p->x_lh.data(), n * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(p->d_lu, p->lu.data(), p->lu.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(p->d_lu_nb, p->lu_nb.data(), p->lu_nb.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            const cudaError_t e = cudaGetLastError();
            cfd_destroy(p);
            return fail(CFD_ECUDA, "upload of plan tables failed: %s", cudaGetErrorString(e));
        }
    }
    if (part_size == 1 && xy_eligible(p->g)) {        // draw order of cfd_apply_xy, default launch shape
        DeviceInfo di;
        int warps = 0, sub = 0;
        double active = 0.0;
        rc = device_info(di);
        if (!rc) { xy_shape(p->g, di.sms, g_slots == 3 ? 3 : 4, p->xy_warps, warps, active, sub); rc = xy_prepare(p, active, sub); }
        if (rc) { cfd_destroy(p); return rc; }
    }
    *out = p;
    return CFD_OK;
}

extern "C" void cfd_destroy(cfd_plan *p)
{
    if (!p) return;
    cudaFree(p->d_x_uh); cudaFree(p->d_x_lh); cudaFree(p->d_lu); cudaFree(p->d_lu_nb);
    cudaFree(p->d_f); cudaFree(p->d_df); cudaFree(p->d_xy_order);
    if (p->hstream) cudaStreamDestroy(p->hstream);
    if (p->gside) cudaStreamDestroy(p->gside);
    if (p->gev_fork) cudaEventDestroy(p->gev_fork);
    if (p->gev_join) cudaEventDestroy(p->gev_join);
    if (p->hcopy) cudaStreamDestroy(p->hcopy);
    if (p->hback) cudaStreamDestroy(p->hback);
    for (auto e : p->hev) if (e) cudaEventDestroy(e);
    delete p;
}

extern "C" long cfd_plane_elems(const cfd_plan *p) { return p ? p->g.nlines : 0; }

extern "C" int cfd_tables_size(void) { return 3 * CH * 2 + 8; }

extern "C" int cfd_plan_tables(const cfd_plan *p, double *out)
{
    if (!p || !out) return fail(CFD_EINVAL, "NULL argument");
    memcpy(out, &p->kp.head, sizeof(RowTab));
    memcpy(out + 3 * CH, &p->kp.tail, sizeof(RowTab));
    double *s = out + 6 * CH;
    s[0] = p->kp.sk_mid; s[1] = p->kp.l_mid; s[2] = p->kp.g_mid; s[3] = p->kp.s0c; s[4] = p->kp.snc;
    s[5] = p->kp.K; s[6] = p->kp.jl; s[7] = p->wc;
    return CFD_OK;
}

extern "C" int cfd_plan_secondary(const cfd_plan *p, double *x_uh, double *x_lh, double *ra, double *rb, double *rc)
{
    if (!p) return fail(CFD_EINVAL, "NULL plan");
    if (p->size < 2) return fail(CFD_EINVAL, "plan has part_size 1: no secondary systems");
    const int n = p->g.n, mm = 2 * p->size;
    if (x_uh) memcpy(x_uh, p->x_uh.data(), n * sizeof(double));
    if (x_lh) memcpy(x_lh, p->x_lh.data(), n * sizeof(double));
    if (ra) memcpy(ra, p->ra.data(), mm * sizeof(double));
    if (rb) memcpy(rb, p->rb.data(), mm * sizeof(double));
    if (rc) memcpy(rc, p->rc.data(), mm * sizeof(double));
    return CFD_OK;
}

// Host-only inspection (no device needed): the tables cfd_create / nt_create would build for a line of
// n rows.  out: head[96], tail[96], then sk_mid, l_mid, g_mid, beta_0, beta_{n-1}, K, jl, fast_ok.
extern "C" int cfd_debug_tables(int n, const double coeffs[7], double scale, double *out)
{
    if (n < 3 || !coeffs || !out) return fail(CFD_EINVAL, "bad argument");
    Geometry g;
    memset(&g, 0, sizeof g);
    g.n = n; g.K = (n + CH - 1) / CH; g.jl = (n - 1) - CH * (g.K - 1); g.inner = 1; g.inner_tiles = 1;
    KParams kp;
    const LineCoeffs m = {coeffs[0], coeffs[1], coeffs[2], coeffs[3], coeffs[4], coeffs[5], coeffs[6]};
    int rc = fill_tables(kp, g, m, scale, false);
    if (rc) return rc;
    Pivots pv = build_pivots(n, m);
    memcpy(out, &kp.head, sizeof(RowTab));
    memcpy(out + 3 * CH, &kp.tail, sizeof(RowTab));
    double *s = out + 6 * CH;
    s[0] = kp.sk_mid; s[1] = kp.l_mid; s[2] = kp.g_mid; s[3] = kp.s0c; s[4] = kp.snc;
    s[5] = kp.K; s[6] = kp.jl;
    s[7] = (g.K <= 2 || (pv.converged && std::pow(pv.decay, CH) <= 1.2e-16)) ? 1.0 : 0.0;
    return CFD_OK;
}

// ------------------------------------------------------------------------------------------------
// General one-pass kernel (kernels_general.cuh): look-ahead derived from the coefficients, compact schemes beyond Pade-4
// ------------------------------------------------------------------------------------------------
// Chunks of backward look-ahead a matrix needs so that the neglected coupling |g|^(32 LA) stays below fp64 round-off;
// 0 = not served one-pass (pivots that do not converge by row 32, or a coupling weaker than 0.563 per row).
static int lookahead_chunks(const Pivots &pv, int K)
{
    if (K <= 2) return 1;                                   // at most two chunks: every sweep starts at the true end
    if (!pv.converged) return 0;
    if (std::pow(pv.decay, CH) <= 1.2e-16) return 1;
    if (std::pow(pv.decay, 2 * CH) <= 1.2e-16) return 2;
    return 0;
}

static void fill_general(GParams &gp, const Geometry &g, const Pivots &pv, double scale)
{
    memset(&gp, 0, sizeof gp);
    gp.n = g.n; gp.K = g.K; gp.jl = g.jl;
    gp.inner = (int)g.inner; gp.inner_tiles = g.inner_tiles;
    gp.nb = g.nb; gp.rows = g.nlines;
    gp.head = chunk_table(pv, g.n, 0, scale);
    gp.tail = chunk_table(pv, g.n, g.K - 1, scale);
    gp.tail2 = chunk_table(pv, g.n, g.K >= 2 ? g.K - 2 : 0, scale);
    if (g.K > 3) { gp.sk_mid = pv.beta[CH] * scale; gp.l_mid = pv.l[CH]; gp.g_mid = pv.g[CH]; }
    gp.nspecial = 0;
}

template <bool CONTIG, int STENCIL, int LA, int NS>
static int launch_general_ns(const Geometry &g, GParams gp, const CUtensorMap &tm_in, const CUtensorMap &tm_out,
                             cudaStream_t stream, PairPool *pool)
{
    static DeviceInfo dinfo;
    if (!dinfo.ok) { int rc = device_info(dinfo); if (rc) return rc; }
    constexpr int per_warp = NS * SLOT_BYTES + NS * 16;
    constexpr int max_warps = (232448 - 1024) / per_warp < 7 ? (232448 - 1024) / per_warp : 7;
    int warps = g_warps ? g_warps : max_warps;
    if (warps > max_warps) warps = max_warps;
    const long per_sm = (g.nb + dinfo.sms - 1) / dinfo.sms;
    if (!g_warps && per_sm < warps) warps = (int)(per_sm < 1 ? 1 : per_sm);
    const size_t smem = (size_t)warps * per_warp + 1024;
    auto kern = stream_kernel_g<CONTIG, STENCIL, LA, NS>;
    static size_t configured[MAX_DEVICES] = {0};
    int dev = 0;
    { int rc = current_device(dev); if (rc) return rc; }
    if (configured[dev] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_warps * per_warp + 1024));
        configured[dev] = max_warps * per_warp + 1024;
    }
    int rc = counter_pair(&gp.counter, stream, pool);
    if (rc) return rc;
    long blocks = (g.nb + warps - 1) / warps;
    if (blocks > dinfo.sms) blocks = dinfo.sms;
    CUDA_TRY(launch_k(kern, (unsigned)blocks, warps * 32, smem, stream, tm_in, tm_out, gp));
    g_launches++;
    return CFD_OK;
}

template <bool CONTIG, int STENCIL, int LA>
static int launch_general_la(const Geometry &g, const GParams &gp, const CUtensorMap &tm_in, const CUtensorMap &tm_out,
                             cudaStream_t stream, PairPool *pool)
{
    // 4 ring slots per warp (32 KiB) and 7 warps per SM.  Measured at 512^3 (scripts/time_schemes.py,
    // profiles/r2j_time_schemes_512.txt): the kernel is latency-bound on its recurrence chains, time x warps ~ const --
    // 6th-order derivative 0.465 / 0.418 / 0.381 / 0.351 ms with 4 / 5 / 6 / 7 warps; 5 slots x 5 warps 0.379 ms.
    return launch_general_ns<CONTIG, STENCIL, LA, 4>(g, gp, tm_in, tm_out, stream, pool);
}

template <int STENCIL>
static int launch_general(const Geometry &g, const GParams &gp, int la, const CUtensorMap &tm_in, const CUtensorMap &tm_out,
                          cudaStream_t stream, PairPool *pool)
{
    if (g.contig) return la == 2 ? launch_general_la<true, STENCIL, 2>(g, gp, tm_in, tm_out, stream, pool)
                                 : launch_general_la<true, STENCIL, 1>(g, gp, tm_in, tm_out, stream, pool);
    return la == 2 ? launch_general_la<false, STENCIL, 2>(g, gp, tm_in, tm_out, stream, pool)
                   : launch_general_la<false, STENCIL, 1>(g, gp, tm_in, tm_out, stream, pool);
}

// Compact schemes beyond the reference's (Lele, "Compact finite difference schemes with spectral-like resolution",
// J. Comput. Phys. 103, 1992): matrix rows, interior stencil, closure rows.
struct SchemeDef {
    LineMatrix m;
    double c0, c1, c2, sgn;
    int nspecial, min_n;
    double q[2][4], p[2][4];
};

static int make_scheme(int scheme, double h, SchemeDef &d)
{
    memset(&d, 0, sizeof d);
    if (scheme == CFD_SCHEME_COMPACT6) {
        // interior (eq. 2.1, alpha = 1/3): f'_{i-1}/3 + f'_i + f'_{i+1}/3 = 14/9 (f_{i+1} - f_{i-1}) / 2h + 1/9 (f_{i+2} - f_{i-2}) / 4h
        // row 1 / n-2: the 4th-order Pade row  f'_{i-1}/4 + f'_i + f'_{i+1}/4 = 3/4 (f_{i+1} - f_{i-1}) / h
        // row 0 / n-1: the 3rd-order closure  f'_0 + 2 f'_1 = (-5/2 f_0 + 2 f_1 + 1/2 f_2) / h   (eq. 4.1.4; the reference's closure)
        d.m = {1.0, 2.0, 1.0 / 3, 1.0, 1.0 / 3, 2.0, 1.0, true, 0.25, 1.0, 0.25, 0.25, 1.0, 0.25};
        d.c0 = 0.0; d.c1 = 14.0 / 9 / (2 * h); d.c2 = 1.0 / 9 / (4 * h); d.sgn = -1.0;
        d.nspecial = 2; d.min_n = 6;
        const double q0[4] = {-2.5 / h, 2.0 / h, 0.5 / h, 0.0}, q1[4] = {-0.75 / h, 0.0, 0.75 / h, 0.0};
        for (int k = 0; k < 4; k++) { d.q[0][k] = q0[k]; d.q[1][k] = q1[k]; d.p[0][k] = -q0[k]; }
        d.p[1][0] = 0.75 / h; d.p[1][1] = 0.0; d.p[1][2] = -0.75 / h; d.p[1][3] = 0.0;     // 3/4 (f_{n-1} - f_{n-3}) / h
        return CFD_OK;
    }
    if (scheme == CFD_SCHEME_PADE4_D2) {
        // interior (eq. 2.2, alpha = 1/10): f''_{i-1}/10 + f''_i + f''_{i+1}/10 = 6/5 (f_{i+1} - 2 f_i + f_{i-1}) / h^2
        // row 0 / n-1 (eq. 4.3.4, 3rd order): f''_0 + 11 f''_1 = (13 f_0 - 27 f_1 + 15 f_2 - f_3) / h^2
        const double ih2 = 1.0 / (h * h);
        d.m = {1.0, 11.0, 0.1, 1.0, 0.1, 11.0, 1.0, false, 0, 0, 0, 0, 0, 0};
        d.c0 = -2.4 * ih2; d.c1 = 1.2 * ih2; d.c2 = 0.0; d.sgn = 1.0;
        d.nspecial = 1; d.min_n = 5;
        const double q0[4] = {13.0 * ih2, -27.0 * ih2, 15.0 * ih2, -1.0 * ih2};
        for (int k = 0; k < 4; k++) { d.q[0][k] = q0[k]; d.p[0][k] = q0[k]; }
        return CFD_OK;
    }
    return fail(CFD_EINVAL, "unknown scheme %d", scheme);
}

extern "C" int cfd_create_scheme(cfd_plan **out, int nz, int ny, int nx, int axis, double h, int scheme)
{
    if (scheme == CFD_SCHEME_PADE4) return cfd_create(out, nz, ny, nx, axis, h, 0, 1);
    if (!out) return fail(CFD_EINVAL, "plan pointer is NULL");
    *out = nullptr;
    if (!(h > 0.0) || !std::isfinite(h)) return fail(CFD_EINVAL, "spacing h = %g must be positive", h);
    SchemeDef d;
    int rc = make_scheme(scheme, h, d);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(CFD_ECUDA, "no CUDA device: libcfd_b200 has no CPU path");
    }
    cfd_plan *p = new cfd_plan();
    rc = make_geometry(p->g, nz, ny, nx, axis);
    if (!rc && p->g.n < d.min_n) rc = fail(CFD_EINVAL, "extent along axis %d is %d; scheme %d needs >= %d", axis, p->g.n, scheme, d.min_n);
    if (!rc) rc = ensure_counters();
    if (rc) { delete p; return rc; }
    p->h = h; p->rank = 0; p->size = 1; p->scheme = scheme;
    const Pivots pv = build_pivots(p->g.n, d.m);
    if (!pv.finite) { delete p; return fail(CFD_EINVAL, "internal: zero pivot in scheme %d", scheme); }
    p->la = lookahead_chunks(pv, p->g.K);
    if (p->la == 0) { delete p; return fail(CFD_EUNSUPPORTED, "internal: scheme %d is not diagonally dominant enough", scheme); }
    fill_general(p->gp, p->g, pv, 1.0);
    p->gp.c1 = d.c1; p->gp.c2 = d.c2; p->gp.nspecial = d.nspecial;
    p->second = d.sgn > 0.0;
    // the kernel evaluates every stencil over differences (kernels_general.cuh): that needs stencils that annihilate
    // constants -- true of any derivative stencil, checked here
    {
        bool ok = std::fabs(d.c0 + (d.sgn > 0.0 ? 2.0 * (d.c1 + d.c2) : 0.0)) <= 1e-12 * (std::fabs(d.c1) + std::fabs(d.c2));
        for (int r = 0; r < d.nspecial; r++) {
            double sq = 0, sp = 0, aq = 0, ap = 0;
            for (int k = 0; k < 4; k++) { sq += d.q[r][k]; sp += d.p[r][k]; aq += std::fabs(d.q[r][k]); ap += std::fabs(d.p[r][k]); }
            ok = ok && std::fabs(sq) <= 1e-12 * aq && std::fabs(sp) <= 1e-12 * ap;
        }
        if (!ok) { delete p; return fail(CFD_EINVAL, "internal: scheme %d has a stencil that does not annihilate constants", scheme); }
    }
    memcpy(p->gp.q, d.q, sizeof d.q);
    memcpy(p->gp.p, d.p, sizeof d.p);
    memset(&p->kp, 0, sizeof p->kp);
    p->kp.lo_closure = p->kp.hi_closure = 1;
    *out = p;
    return CFD_OK;
}

extern "C" int cfd_plan_lookahead(const cfd_plan *p) { return p ? p->la : 0; }

// Plan of one rank for the distributed npts solve: the LU of the WHOLE line, pivots handed from rank to rank
// (lanl-implementation/npts.c:580-655 precompute_beta_gam; here every rank computes the global sequence itself, it
// depends on the line length only) and this rank's rows cut out of it.
extern "C" int cfd_create_npts(cfd_plan **out, int nz, int ny, int nx, int axis, double h, int part_rank, int part_size)
{
    if (!out) return fail(CFD_EINVAL, "plan pointer is NULL");
    *out = nullptr;
    if (!(h > 0.0) || !std::isfinite(h)) return fail(CFD_EINVAL, "spacing h = %g must be positive", h);
    if (part_size < 2 || part_rank < 0 || part_rank >= part_size || part_size > 64)
        return fail(CFD_EINVAL, "partition position %d of %d is invalid (npts plans are for partitioned lines)", part_rank, part_size);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(CFD_ECUDA, "no CUDA device: libcfd_b200 has no CPU path");
    }
    cfd_plan *p = new cfd_plan();
    int rc = make_geometry(p->g, nz, ny, nx, axis);
    if (!rc && p->g.n < 2 * CH + 2) rc = fail(CFD_EUNSUPPORTED, "distributed npts needs >= %d rows per block (have %d)", 2 * CH + 2, p->g.n);
    if (!rc) rc = ensure_counters();
    if (!rc) rc = ensure_err_word();
    if (rc) { delete p; return rc; }
    p->h = h; p->rank = part_rank; p->size = part_size; p->npts = true;
    const int n = p->g.n;
    const LineCoeffs whole = {1.0, 2.0, 0.25, 1.0, 0.25, 2.0, 1.0};
    const Pivots gp = build_pivots(n * part_size, whole);
    Pivots pv;                                               // this rank's slice
    pv.beta.assign(gp.beta.begin() + (long)part_rank * n, gp.beta.begin() + (long)(part_rank + 1) * n);
    pv.l.assign(gp.l.begin() + (long)part_rank * n, gp.l.begin() + (long)(part_rank + 1) * n);
    pv.g.assign(gp.g.begin() + (long)part_rank * n, gp.g.begin() + (long)(part_rank + 1) * n);
    const double scale = 3.0 / (4.0 * h);
    KParams &kp = p->kp;
    memset(&kp, 0, sizeof kp);
    kp.n = n; kp.K = p->g.K; kp.jl = p->g.jl; kp.kseg = p->g.K; kp.nseg = 1;
    kp.inner = (int)p->g.inner; kp.inner_tiles = p->g.inner_tiles; kp.outer = (int)p->g.outer;
    kp.nb = p->g.nb; kp.rows = p->g.nlines;
    kp.head = chunk_table(pv, n, 0, scale);
    kp.tail = chunk_table(pv, n, p->g.K - 1, scale);
    kp.sk_mid = pv.beta[CH] * scale; kp.l_mid = pv.l[CH]; kp.g_mid = pv.g[CH];
    kp.lo_closure = (part_rank == 0);
    kp.hi_closure = (part_rank == part_size - 1);
    kp.s0c = pv.beta[0] / (2.0 * h);                         // closure rows of the global ends (ranks 0 and P-1 only)
    kp.snc = pv.beta[n - 1] / (2.0 * h);
    kp.snb = kp.hi_closure ? 0.0 : pv.g[n - 1];              // gamma of the next rank's first row = beta_{n-1} c_i
    // the backward sweep of the last chunk ends at row n-1 with x_{n-1} = e_{n-1} - snb x~ : no g beyond it
    kp.tail.g[p->g.jl] = 0.0;
    *out = p;
    return CFD_OK;
}

// Host-only inspection (no device): look-ahead chunks for a line of n rows of the matrix coeffs, and the definition
// of a scheme: out[0..6] = b1,c1,ai,bi,ci,an,bn; [7] = two special rows per end; [8..13] = a2,b2,c2,am,bm,cm;
// [14..18] = c0,c1,c2,sgn,closure rows per end; [19..26] = q[2][4]; [27..34] = p[2][4]; [35] = look-ahead chunks for n
// rows; [36] = interior coupling per row.
extern "C" int cfd_debug_lookahead(int n, const double coeffs[7])
{
    if (n < 3 || !coeffs) return fail(CFD_EINVAL, "bad argument");
    const LineCoeffs m = {coeffs[0], coeffs[1], coeffs[2], coeffs[3], coeffs[4], coeffs[5], coeffs[6]};
    const Pivots pv = build_pivots(n, as_matrix(m));
    if (!pv.finite) return fail(CFD_EINVAL, "zero pivot");
    return lookahead_chunks(pv, (n + CH - 1) / CH);
}

extern "C" int cfd_debug_scheme(int scheme, int n, double h, double *out)
{
    if (!out || n < 6 || !(h > 0.0)) return fail(CFD_EINVAL, "bad argument");
    SchemeDef d;
    if (scheme == CFD_SCHEME_PADE4) {
        memset(&d, 0, sizeof d);
        d.m = as_matrix(pade_block(0, 1));
        d.c1 = 3.0 / (4.0 * h); d.sgn = -1.0; d.nspecial = 1;
        const double q0[4] = {-2.5 / h, 2.0 / h, 0.5 / h, 0.0};
        for (int k = 0; k < 4; k++) { d.q[0][k] = q0[k]; d.p[0][k] = -q0[k]; }
    } else {
        int rc = make_scheme(scheme, h, d);
        if (rc) return rc;
    }
    const double v[19] = {d.m.b1, d.m.c1, d.m.ai, d.m.bi, d.m.ci, d.m.an, d.m.bn, d.m.two ? 1.0 : 0.0, d.m.a2, d.m.b2, d.m.c2,
                          d.m.am, d.m.bm, d.m.cm, d.c0, d.c1, d.c2, d.sgn, (double)d.nspecial};
    memcpy(out, v, sizeof v);
    memcpy(out + 19, d.q, sizeof d.q);
    memcpy(out + 27, d.p, sizeof d.p);
    const Pivots pv = build_pivots(n, d.m);
    out[35] = lookahead_chunks(pv, (n + CH - 1) / CH);
    out[36] = pv.decay;
    return CFD_OK;
}

static int get_maps(MapCache &c, const Geometry &g, const void *in, const void *out, MapPair &m)
{
    std::lock_guard<std::mutex> lock(c.mu);
    if (!(c.in == in && c.out == out)) {
        int rc = encode_maps(g, in, out, &c.tm_in, &c.tm_out);
        if (rc) { c.in = c.out = nullptr; return rc; }
        c.in = in; c.out = out;
    }
    m.tm_in = c.tm_in; m.tm_out = c.tm_out;
    return CFD_OK;
}


static int launch_one_direction_ring(cfd_plan *p, const MapPair &mp, cudaStream_t stream);

static int apply_impl(cfd_plan *p, const double *f, double *df, const double *halo_lo, const double *halo_hi,
                      const double *ab, void *stream)
{
    if (!p || !f || !df) return fail(CFD_EINVAL, "NULL argument");
    if (f == df) return fail(CFD_EINVAL, "the derivative is out of place: f and df must differ");
    if (!p->kp.lo_closure && !halo_lo) return fail(CFD_EINVAL, "rank %d of %d needs halo_lo", p->rank, p->size);
    if (!p->kp.hi_closure && !halo_hi) return fail(CFD_EINVAL, "rank %d of %d needs halo_hi", p->rank, p->size);
    MapPair mp;
    int rc = get_maps(p->cache, p->g, f, df, mp);
    if (rc) return rc;
    if (p->scheme != CFD_SCHEME_PADE4) {
        if (p->second) return launch_general<2>(p->g, p->gp, p->la, mp.tm_in, mp.tm_out, (cudaStream_t)stream, &p->pool);
        return launch_general<1>(p->g, p->gp, p->la, mp.tm_in, mp.tm_out, (cudaStream_t)stream, &p->pool);
    }
    // Experiment switch (CFD_RING_STAGING=1): whole, unpartitioned lines through the ring-staged kernel of
    // kernels_xy.cuh with one direction empty -- 32 KiB per warp, 6 warps per SM instead of 40 KiB and 4.
    if (p->size == 1 && !ab && getenv("CFD_RING_STAGING")) return launch_one_direction_ring(p, mp, (cudaStream_t)stream);
    KParams kp = p->kp;
    kp.halo_lo = halo_lo; kp.halo_hi = halo_hi;
    kp.ab = ab; kp.nlines = p->g.nlines;
    if (p->g.contig) return launch_stream<true, true>(p->g, kp, mp.tm_in, mp.tm_out, (cudaStream_t)stream, &p->pool);
    return launch_stream<false, true>(p->g, kp, mp.tm_in, mp.tm_out, (cudaStream_t)stream, &p->pool);
}

extern "C" int cfd_apply(cfd_plan *p, const double *f, double *df, const double *halo_lo, const double *halo_hi,
                         void *stream)
{
    return apply_impl(p, f, df, halo_lo, halo_hi, nullptr, stream);
}

extern "C" int cfd_apply_coupled(cfd_plan *p, const double *f, double *df, const double *halo_lo,
                                 const double *halo_hi, const double *ab, void *stream)
{
    if (p && p->size < 2) return fail(CFD_EINVAL, "plan has part_size 1: use cfd_apply");
    if (!ab) return fail(CFD_EINVAL, "ab (interface unknowns) is NULL");
    return apply_impl(p, f, df, halo_lo, halo_hi, ab, stream);
}

extern "C" int cfd_reduced_unknowns(cfd_plan *p, const double *faces, int neighbours_only, double *ab,
                                    const unsigned long long *flag0, const unsigned long long *flag1,
                                    unsigned long long seq, void *stream)
{
    if (!p || !faces || !ab) return fail(CFD_EINVAL, "NULL argument");
    if (p->size < 2) return fail(CFD_EINVAL, "plan has part_size 1: no interfaces");
    if (neighbours_only && p->g.n < 2 * CH)
        return fail(CFD_EUNSUPPORTED, "neighbour-only coupling needs >= %d rows per block", 2 * CH);
    const int bs = 256;
    CUDA_TRY(launch_k(reduced_planes_kernel, (unsigned)((p->g.nlines + bs - 1) / bs), bs, 0, (cudaStream_t)stream,
                      faces, (const double *)(neighbours_only ? p->d_lu_nb : p->d_lu), p->g.nlines,
                      neighbours_only ? p->nb_pv : p->size, neighbours_only ? p->nb_own : p->rank, ab, flag0, flag1, seq,
                      wait_params()));
    g_launches++;
    return CFD_OK;
}

extern "C" int cfd_nb_layout(const cfd_plan *p, int *virtual_ranks, int *own_index)
{
    if (!p || p->size < 2) return fail(CFD_EINVAL, "plan has part_size 1: no neighbours");
    if (virtual_ranks) *virtual_ranks = p->nb_pv;
    if (own_index) *own_index = p->nb_own;
    return CFD_OK;
}

static int edge_impl(cfd_plan *p, const double *f, const double *halo_lo, const double *halo_hi, double *faces,
                     double *peer_lo, double *peer_hi, unsigned long long *flag_lo, unsigned long long *flag_hi,
                     unsigned long long seq, bool p2p, void *stream, bool defer = false, double *push_lo = nullptr,
                     double *push_hi = nullptr);

template <int NSLOT>
static int launch_xy(cfd_plan *px, cfd_plan *py, const MapPair &mx, const MapPair &my, long nitems, cudaStream_t stream,
                     const EdgeX *edge = nullptr, const CUtensorMap *tmz = nullptr)
{
    static DeviceInfo dinfo;
    int rc;
    if (!dinfo.ok) { rc = device_info(dinfo); if (rc) return rc; }
    // The ring is all the shared memory a warp has (results are staged in the slot just consumed): 8 warps per SM
    // fit with 3 slots (also the limit of 255 registers per thread), 7 with 4.  Measured at 512^3 (scripts/time_xy.py,
    // profiles/r1k_time_xy.txt): 4 slots x 6 warps 0.532 ms, 4 x 7 0.546, 3 x 8 0.546, 4 x 5 0.564; the old
    // 3 + 2 staging slots x 5 warps 0.553; d/dx then d/dy as two launches 0.673.
    constexpr int per_warp = NSLOT * SLOT_BYTES + NSLOT * 16;
    XYParams q;
    q.nxp = px->g.ny / CH;
    q.nyp = py->g.inner_tiles;
    q.nitems = nitems;
    int warps = 0, sub = 0;
    double active = 0.0;
    xy_shape(px->g, dinfo.sms, NSLOT, px->xy_warps, warps, active, sub);
    rc = xy_prepare(px, active, sub);   // no-op unless the launch knobs changed since cfd_create
    if (rc) return rc;
    if (edge && edge->nedge != px->xy_nedge) return fail(CFD_EINVAL, "internal: edge item count does not match the plane");
    q.order = px->d_xy_order + (edge ? 0 : px->xy_nedge);
    q.nitems = px->xy_entries - (edge ? 0 : px->xy_nedge);
    q.nxy = nitems;
    q.kseg = px->xy_kseg;
    q.ksegy = px->xy_ksegy;
    q.slot_items = (float)(2.0 * active);
    // Start-up stagger: pays on long lines only (>= 32 tiles: [128,1024,1024] 0.596 -> 0.577 ms with 1 us per slot;
    // 512^3, 16 tiles: 0.531 -> 0.534), profiles/r1n_time_xy_startup_stagger.txt.  CFD_XY_TAU overrides (ns).
    q.hints = 1;                  // evict_last on the tile loads of the long-line (SEG) variants: -1.5 % on 128 x 1024^2
    if (const char *e = getenv("CFD_XY_HINTS")) q.hints = atoi(e);
    q.tau_ns = (px->g.K + py->g.K >= 64) ? 1000.f : 0.f;
    if (const char *e = getenv("CFD_XY_TAU")) q.tau_ns = (float)atof(e);
    const size_t smem = (size_t)warps * per_warp + 1024;
    const bool seg = q.kseg > 0 || q.ksegy > 0;          // whole lines only: the variant without segment state
    auto kern = seg ? stream_kernel_xy<NSLOT, true, false> : stream_kernel_xy<NSLOT, false, false>;
    if (edge) {
        if constexpr (NSLOT == 4) kern = seg ? stream_kernel_xy<4, true, true> : stream_kernel_xy<4, false, true>;
        else return fail(CFD_EUNSUPPORTED, "the fused x/y + edge launch is built for 4 ring slots");
    }
    static size_t configured[4][MAX_DEVICES] = {{0}, {0}, {0}, {0}};
    const int variant = (seg ? 1 : 0) + (edge ? 2 : 0);
    int dev = 0;
    rc = current_device(dev);
    if (rc) return rc;
    if (configured[variant][dev] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[variant][dev] = smem;
    }
    rc = counter_pair(&q.counter, stream, &px->pool);
    if (rc) return rc;
    KParams kx = px->kp, ky = py->kp;
    kx.ab = nullptr; ky.ab = nullptr;
    long blocks = (q.nitems + warps - 1) / warps;
    if (blocks > dinfo.sms) blocks = dinfo.sms;
    if (const char *e = getenv("CFD_XY_CTAS")) {      // experiment: "auto" = full last round, or an absolute CTA count
        if (!strcmp(e, "auto")) {
            const long rounds = (q.nitems + blocks * warps - 1) / (blocks * warps);
            const long need = (q.nitems + rounds - 1) / rounds;          // warps that make every round full
            const long b2 = (need + warps - 1) / warps;
            if (b2 >= 1 && b2 < blocks) blocks = b2;
        } else if (atol(e) >= 1 && atol(e) <= blocks) blocks = atol(e);
    }
    static const EdgeX no_edge = {};
    CUDA_TRY(launch_k(kern, (unsigned)blocks, warps * 32, smem, stream, mx.tm_in, mx.tm_out, my.tm_in, my.tm_out, kx, ky, q,
                      tmz ? *tmz : mx.tm_in, edge ? *edge : no_edge));
    g_launches++;
    return CFD_OK;
}

static int launch_one_direction_ring(cfd_plan *p, const MapPair &mp, cudaStream_t stream)
{
    static DeviceInfo dinfo;
    int rc;
    if (!dinfo.ok) { rc = device_info(dinfo); if (rc) return rc; }
    constexpr int NSLOT = 4;
    constexpr int per_warp = NSLOT * SLOT_BYTES + NSLOT * 16;
    if (p->g.nb > 0x7fffffffL) return fail(CFD_EUNSUPPORTED, "too many bundles");
    XYParams q;
    q.nxp = p->g.contig ? (int)p->g.nb : 0;
    q.nyp = p->g.contig ? 0 : p->g.inner_tiles;
    q.nitems = p->g.nb;
    q.nxy = p->g.nb;
    q.order = nullptr;
    q.kseg = 0; q.ksegy = 0;
    q.hints = 0;
    q.tau_ns = 0.f; q.slot_items = 0.f;
    int warps = g_warps ? g_warps : 6;
    if (warps > 7) warps = 7;
    const long per_sm = (q.nitems + dinfo.sms - 1) / dinfo.sms;
    if (!g_warps && per_sm < warps) warps = (int)(per_sm < 1 ? 1 : per_sm);
    const size_t smem = (size_t)warps * per_warp + 1024;
    auto kern = stream_kernel_xy<NSLOT, false, false>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * per_warp + 1024));
    rc = counter_pair(&q.counter, stream, &p->pool);
    if (rc) return rc;
    KParams k = p->kp;
    k.ab = nullptr;
    long blocks = (q.nitems + warps - 1) / warps;
    if (blocks > dinfo.sms) blocks = dinfo.sms;
    static const EdgeX no_edge = {};
    kern<<<(unsigned)blocks, warps * 32, smem, stream>>>(mp.tm_in, mp.tm_out, mp.tm_in, mp.tm_out, k, k, q, mp.tm_in, no_edge);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CFD_OK;
}

// Warps per SM of the fused x/y launch for this (axis-0) plan.  Callers that run other kernels BESIDE the launch --
// the exchange chain of a partitioned d/dz -- take 5 instead of the default 6 to leave registers for them
// (scripts/overlap_timeline.py: 1.108 vs 1.133 ms per step on [128,1024,1024] slabs).  The draw order is rebuilt here,
// so cfd_apply_xy itself still allocates nothing.
extern "C" int cfd_plan_set_xy_warps(cfd_plan *p, int warps_per_sm)
{
    if (!p || p->g.axis != 0) return fail(CFD_EINVAL, "cfd_plan_set_xy_warps needs an axis-0 plan");
    if (warps_per_sm < 0 || warps_per_sm > 8) return fail(CFD_EINVAL, "warps per SM must be 0 (default) .. 8");
    p->xy_warps = warps_per_sm;
    if (p->size == 1 && xy_eligible(p->g)) {
        DeviceInfo di;
        int warps = 0, sub = 0;
        double active = 0.0;
        int rc = device_info(di);
        if (rc) return rc;
        xy_shape(p->g, di.sms, g_slots == 3 ? 3 : 4, p->xy_warps, warps, active, sub);
        return xy_prepare(p, active, sub);
    }
    return CFD_OK;
}

extern "C" int cfd_apply_xy(cfd_plan *px, cfd_plan *py, const double *f, double *dfdx, double *dfdy, void *stream)
{
    if (!px || !py || !f || !dfdx || !dfdy) return fail(CFD_EINVAL, "NULL argument");
    if (px->g.axis != 0 || py->g.axis != 1) return fail(CFD_EINVAL, "plans must be for axis 0 (x) and axis 1 (y)");
    if (px->g.nz != py->g.nz || px->g.ny != py->g.ny || px->g.nx != py->g.nx)
        return fail(CFD_EINVAL, "plans are for different shapes");
    if (px->size != 1 || py->size != 1) return fail(CFD_EINVAL, "cfd_apply_xy serves unpartitioned x / y lines");
    if (px->scheme != CFD_SCHEME_PADE4 || py->scheme != CFD_SCHEME_PADE4) {     // other schemes: one launch each
        int rc2 = cfd_apply(px, f, dfdx, nullptr, nullptr, stream);
        return rc2 ? rc2 : cfd_apply(py, f, dfdy, nullptr, nullptr, stream);
    }
    if (f == dfdx || f == dfdy || dfdx == dfdy) return fail(CFD_EINVAL, "f, dfdx, dfdy must be three different fields");
    const long nitems = (long)px->g.nz * (px->g.ny / CH + py->g.inner_tiles);
    if (!xy_eligible(px->g) || getenv("CFD_NO_XY")) {
        int rc = cfd_apply(px, f, dfdx, nullptr, nullptr, stream);
        if (rc) return rc;
        return cfd_apply(py, f, dfdy, nullptr, nullptr, stream);
    }
    MapPair mx, my;
    int rc = get_maps(px->cache, px->g, f, dfdx, mx);
    if (rc) return rc;
    rc = get_maps(py->cache, py->g, f, dfdy, my);
    if (rc) return rc;
    switch (g_slots ? g_slots : 4) {
        case 3: return launch_xy<3>(px, py, mx, my, nitems, (cudaStream_t)stream);
        case 4: return launch_xy<4>(px, py, mx, my, nitems, (cudaStream_t)stream);
        default: return fail(CFD_EINVAL, "cfd_apply_xy: ring slots must be 3 or 4");
    }
}

// The whole gradient of an unpartitioned field: the fused d/dx + d/dy launch on the caller's stream and the d/dz launch
// on a plan-owned side stream, forked and joined with events.  The three results are independent (out of place, same
// input), and both kernels are persistent with one CTA per SM: run one after the other each pays its own ramp-up and
// the tail of its dynamic draw (warps running dry within one bundle-time of each other); on two streams the second
// kernel's CTAs take over the SMs one by one as the first one's CTAs retire.
extern "C" int cfd_apply_xyz(cfd_plan *px, cfd_plan *py, cfd_plan *pz, const double *f, double *dfdx, double *dfdy,
                             double *dfdz, void *stream)
{
    if (!px || !py || !pz || !f || !dfdx || !dfdy || !dfdz) return fail(CFD_EINVAL, "NULL argument");
    if (pz->g.axis != 2 || pz->size != 1) return fail(CFD_EINVAL, "plan_z must be an unpartitioned axis-2 plan");
    if (pz->g.nz != px->g.nz || pz->g.ny != px->g.ny || pz->g.nx != px->g.nx) return fail(CFD_EINVAL, "plans are for different shapes");
    if (dfdz == dfdx || dfdz == dfdy || dfdz == f) return fail(CFD_EINVAL, "f and the three derivatives must be four different fields");
    cudaStream_t st = (cudaStream_t)stream;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
    if (!pz->gside && cs == cudaStreamCaptureStatusNone && !getenv("CFD_XYZ_SERIAL")) {
        if (cudaStreamCreateWithFlags(&pz->gside, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&pz->gev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&pz->gev_join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            if (pz->gside) { cudaStreamDestroy(pz->gside); pz->gside = nullptr; }
        }
    }
    // Measured with the variants interleaved (scripts/time_gradient.py, profiles/r2k_time_gradient_two_streams.txt):
    // 128^3 0.0402 -> 0.0367 ms, 256^3 0.1403 -> 0.1318, 512^3 0.8773 -> 0.8630, 1024^3 7.55 -> 7.35-7.55 (never slower).
    if (!pz->gside || !pz->gev_fork || !pz->gev_join || getenv("CFD_XYZ_SERIAL")) {     // one stream: x/y, then z
        int rc = cfd_apply_xy(px, py, f, dfdx, dfdy, stream);
        return rc ? rc : cfd_apply(pz, f, dfdz, nullptr, nullptr, stream);
    }
    CUDA_TRY(cudaEventRecord(pz->gev_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(pz->gside, pz->gev_fork, 0));
    int rc;
    if (getenv("CFD_XYZ_XY_FIRST")) {
        rc = cfd_apply_xy(px, py, f, dfdx, dfdy, stream);
        if (!rc) rc = cfd_apply(pz, f, dfdz, nullptr, nullptr, pz->gside);
    } else {
        rc = cfd_apply(pz, f, dfdz, nullptr, nullptr, pz->gside);
        if (!rc) rc = cfd_apply_xy(px, py, f, dfdx, dfdy, stream);
    }
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(pz->gev_join, pz->gside));
    CUDA_TRY(cudaStreamWaitEvent(st, pz->gev_join, 0));
    return CFD_OK;
}

extern "C" int cfd_edge_faces(cfd_plan *p, const double *f, const double *halo_lo, const double *halo_hi,
                              double *faces, void *stream)
{
    return edge_impl(p, f, halo_lo, halo_hi, faces, nullptr, nullptr, nullptr, nullptr, 0, false, stream);
}

extern "C" int cfd_edge_faces_p2p(cfd_plan *p, const double *f, const double *halo_lo, const double *halo_hi,
                                  double *faces, double *peer_lo, double *peer_hi, unsigned long long *flag_lo,
                                  unsigned long long *flag_hi, unsigned long long seq, void *stream)
{
    return edge_impl(p, f, halo_lo, halo_hi, faces, peer_lo, peer_hi, flag_lo, flag_hi, seq, true, stream);
}

extern "C" int cfd_push_planes(const double *src0, double *dst0, const double *src1, double *dst1, long n,
                               unsigned long long *flag0, unsigned long long *flag1, unsigned long long seq,
                               void *stream)
{
    if (n < 2 || (n & 1)) return fail(CFD_EINVAL, "plane size %ld must be even", n);
    if ((dst0 && !src0) || (dst1 && !src1)) return fail(CFD_EINVAL, "destination without source");
    unsigned long long *done = nullptr;
    int rc = counter_pair(&done, (cudaStream_t)stream);
    if (rc) return rc;
    const int bs = 256;
    long blocks = (n / 2 + bs - 1) / bs;
    if (blocks > 592) blocks = 592;
    push_planes_kernel<<<(unsigned)blocks, bs, 0, (cudaStream_t)stream>>>(
        (const double2 *)src0, (double2 *)dst0, (const double2 *)src1, (double2 *)dst1, n / 2, flag0, flag1, done, seq);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CFD_OK;
}

extern "C" int cfd_wait_flags(const unsigned long long *flag0, const unsigned long long *flag1,
                              unsigned long long seq, void *stream)
{
    if (!flag0 && !flag1) return CFD_OK;
    wait_flags_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag0, flag1, seq, wait_params());
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CFD_OK;
}

extern "C" int cfd_edge_faces_push(cfd_plan *p, const double *f, double *faces, double *peer_face_lo,
                                   double *peer_face_hi, double *peer_halo_lo, double *peer_halo_hi,
                                   unsigned long long *flag_lo, unsigned long long *flag_hi, unsigned long long seq,
                                   void *stream)
{
    return edge_impl(p, f, nullptr, nullptr, faces, peer_face_lo, peer_face_hi, flag_lo, flag_hi, seq, true, stream,
                     true, peer_halo_lo, peer_halo_hi);
}

extern "C" int cfd_reduced_unknowns_deferred(cfd_plan *p, const double *faces_nb, const double *halo_lo,
                                             const double *halo_hi, const double *f, double *ab,
                                             const unsigned long long *flag0, const unsigned long long *flag1,
                                             unsigned long long seq, void *stream)
{
    if (!p || !faces_nb || !f || !ab) return fail(CFD_EINVAL, "NULL argument");
    if (p->size < 2) return fail(CFD_EINVAL, "plan has part_size 1: no interfaces");
    if (p->g.n < 2 * CH + 2) return fail(CFD_EUNSUPPORTED, "neighbour-only coupling needs >= %d rows per block", 2 * CH + 2);
    if ((p->rank > 0) != (halo_lo != nullptr) || (p->rank < p->size - 1) != (halo_hi != nullptr))
        return fail(CFD_EINVAL, "rank %d of %d: halo_lo / halo_hi must be given exactly where a neighbour exists", p->rank, p->size);
    const int bs = 256;
    CUDA_TRY(launch_k(reduced_planes_deferred_kernel, (unsigned)((p->g.nlines + bs - 1) / bs), bs, 0, (cudaStream_t)stream,
                      faces_nb, (const double *)p->d_lu_nb, p->g.nlines, p->nb_pv, p->nb_own, ab, halo_lo, halo_hi, f,
                      p->g.inner, p->g.n, p->w_lo, p->w_hi, flag0, flag1, seq, wait_params()));
    g_launches++;
    return CFD_OK;
}

static int edge_impl(cfd_plan *p, const double *f, const double *halo_lo, const double *halo_hi, double *faces,
                     double *peer_lo, double *peer_hi, unsigned long long *flag_lo, unsigned long long *flag_hi,
                     unsigned long long seq, bool p2p, void *stream, bool defer, double *push_lo, double *push_hi)
{
    if (!p || !f || !faces) return fail(CFD_EINVAL, "NULL argument");
    if (p->size < 2) return fail(CFD_EINVAL, "plan has part_size 1: no interfaces");
    if (p->g.n < 2 * CH + 2)
        return fail(CFD_EUNSUPPORTED, "cfd_edge_faces needs >= %d rows per block (have %d): use cfd_apply + "
                    "cfd_interface_pack + cfd_reduced_correct", 2 * CH + 2, p->g.n);
    if (!defer && !p->kp.lo_closure && !halo_lo) return fail(CFD_EINVAL, "rank %d of %d needs halo_lo", p->rank, p->size);
    if (!defer && !p->kp.hi_closure && !halo_hi) return fail(CFD_EINVAL, "rank %d of %d needs halo_hi", p->rank, p->size);
    EdgeP ep;
    memset(&ep, 0, sizeof ep);
    ep.nlines = p->g.nlines; ep.inner = p->g.inner; ep.n = p->g.n; ep.jl = p->g.jl;
    ep.lo_closure = p->kp.lo_closure; ep.hi_closure = p->kp.hi_closure;
    ep.sk_mid = p->kp.sk_mid; ep.l_mid = p->kp.l_mid; ep.s0c = p->kp.s0c; ep.snc = p->kp.snc;
    ep.sk_last = p->kp.tail.sk[p->g.jl]; ep.l_last = p->kp.tail.l[p->g.jl];
    ep.halo_lo = halo_lo; ep.halo_hi = halo_hi;
    ep.head = p->kp.head;
    ep.defer = defer ? 1 : 0;
    ep.push_lo = p->kp.lo_closure ? nullptr : push_lo;
    ep.push_hi = p->kp.hi_closure ? nullptr : push_hi;
    if (p2p) {
        ep.peer_lo = peer_lo; ep.peer_hi = peer_hi; ep.flag_lo = flag_lo; ep.flag_hi = flag_hi; ep.seq = seq;
        int rc = counter_pair(&ep.done, (cudaStream_t)stream, &p->pool);
        if (rc) return rc;
    }
    int bs = 128;
    if (const char *e = getenv("CFD_EDGE_BS")) { bs = atoi(e); if (bs < 32 || bs > 128 || bs % 32) bs = 128; }
    edge_faces_kernel<<<(unsigned)((ep.nlines + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(f, faces, ep);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CFD_OK;
}

extern "C" int cfd_compute_rhs(cfd_plan *p, const double *f, double *rhs, const double *halo_lo, const double *halo_hi,
                               void *stream)
{
    if (!p || !f || !rhs) return fail(CFD_EINVAL, "NULL argument");
    if (p->scheme != CFD_SCHEME_PADE4) return fail(CFD_EUNSUPPORTED, "the stage entry points serve the reference's Pade-4 scheme");
    if (f == rhs) return fail(CFD_EINVAL, "the right-hand side is computed out of place");
    if (!p->kp.lo_closure && !halo_lo) return fail(CFD_EINVAL, "rank %d of %d needs halo_lo", p->rank, p->size);
    if (!p->kp.hi_closure && !halo_hi) return fail(CFD_EINVAL, "rank %d of %d needs halo_hi", p->rank, p->size);
    const long total = p->g.nlines * p->g.n;
    const int bs = 256;
    rhs_kernel<<<(unsigned)((total + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(
        f, rhs, total, p->g.n, p->g.inner, p->h, p->kp.lo_closure, p->kp.hi_closure, halo_lo, halo_hi);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CFD_OK;
}

extern "C" int cfd_sum_solutions(cfd_plan *p, double *x, const double *alpha, const double *beta, void *stream)
{
    if (!p || !x || !alpha || !beta) return fail(CFD_EINVAL, "NULL argument");
    if (p->size < 2) return fail(CFD_EINVAL, "plan has part_size 1: no secondary solutions");
    const long total = p->g.nlines * p->g.n;
    const int bs = 256;
    sum_solutions_kernel<<<(unsigned)((total + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(
        x, alpha, beta, p->d_x_uh, p->d_x_lh, total, p->g.n, p->g.inner);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CFD_OK;
}

// Block-local coefficient list [b1, c1, ai, bi, ci, an, bn] of the plan (code/cuda/compact.py:159-166), so that a
// caller can build the matching NearToeplitzSolver for the reference's solve_primary_system stage.
extern "C" int cfd_plan_coeffs(const cfd_plan *p, double coeffs[7])
{
    if (!p || !coeffs) return fail(CFD_EINVAL, "NULL argument");
    if (p->scheme != CFD_SCHEME_PADE4) return fail(CFD_EUNSUPPORTED, "the stage entry points serve the reference's Pade-4 scheme");
    const LineCoeffs m = pade_block(p->rank, p->size);
    coeffs[0] = m.b1; coeffs[1] = m.c1; coeffs[2] = m.ai; coeffs[3] = m.bi; coeffs[4] = m.ci; coeffs[5] = m.an;
    coeffs[6] = m.bn;
    return CFD_OK;
}

extern "C" int cfd_interface_pack(cfd_plan *p, const double *df, double *faces, void *stream)
{
    if (!p || !df || !faces) return fail(CFD_EINVAL, "NULL argument");
    GeomP gp = {p->g.nlines, p->g.n, p->g.inner};
    const int bs = 256;
    interface_pack_kernel<<<(unsigned)((gp.nlines + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(
        df, faces, gp, p->kp.lo_closure, p->kp.hi_closure);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CFD_OK;
}

extern "C" int cfd_reduced_correct(cfd_plan *p, double *df, const double *faces_all, void *stream)
{
    if (!p || !df || !faces_all) return fail(CFD_EINVAL, "NULL argument");
    if (p->size < 2) return fail(CFD_EINVAL, "plan has part_size 1: nothing to correct");
    GeomP gp = {p->g.nlines, p->g.n, p->g.inner};
    const int bs = 128;
    if (p->g.contig) {
        const long threads = gp.nlines * 32;
        reduced_correct_kernel<true><<<(unsigned)((threads + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(
            df, faces_all, p->d_lu, p->d_x_uh, p->d_x_lh, gp, p->size, p->rank, p->wc);
    } else {
        reduced_correct_kernel<false><<<(unsigned)((gp.nlines + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(
            df, faces_all, p->d_lu, p->d_x_uh, p->d_x_lh, gp, p->size, p->rank, p->wc);
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CFD_OK;
}

static bool is_page_locked(const void *ptr)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

extern "C" int cfd_apply_host(cfd_plan *p, const double *f_host, double *df_host, int pinned)
{
    if (!p || !f_host || !df_host) return fail(CFD_EINVAL, "NULL argument");
    if (p->size != 1) return fail(CFD_EINVAL, "cfd_apply_host serves part_size == 1 only");
    // pinned != 0 is a promise the asynchronous copies rely on (they would silently fall back to the driver's staged
    // path otherwise): hold the caller to it
    if (pinned && !(is_page_locked(f_host) && is_page_locked(df_host)))
        return fail(CFD_EINVAL, "pinned = %d but the host buffers are not page-locked (cudaHostAlloc / cudaHostRegister)", pinned);
    const size_t bytes = (size_t)p->g.nlines * p->g.n * sizeof(double);
    if (!p->d_f) {
        CUDA_TRY(cudaMalloc(&p->d_f, bytes));
        CUDA_TRY(cudaMalloc(&p->d_df, bytes));
        CUDA_TRY(cudaStreamCreateWithFlags(&p->hstream, cudaStreamNonBlocking));
    }
    // Lines along x or y never leave a z-slab: with page-locked buffers the field moves in slabs, so that slab s+1
    // travels host -> device while slab s is differentiated and slab s-1 travels back (PCIe is full duplex).  Pageable
    // buffers (and z lines, which need the whole field) take the plain copy - kernel - copy sequence.
    const int slabs = (pinned && p->g.axis != 2 && p->g.nz >= 8 && p->scheme == CFD_SCHEME_PADE4) ? 4 : 1;
    if (slabs == 1) {
        CUDA_TRY(cudaMemcpyAsync(p->d_f, f_host, bytes, cudaMemcpyHostToDevice, p->hstream));
        int rc = cfd_apply(p, p->d_f, p->d_df, nullptr, nullptr, p->hstream);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(df_host, p->d_df, bytes, cudaMemcpyDeviceToHost, p->hstream));
        CUDA_TRY(cudaStreamSynchronize(p->hstream));
        return CFD_OK;
    }
    if (!p->hcopy) {
        CUDA_TRY(cudaStreamCreateWithFlags(&p->hcopy, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&p->hback, cudaStreamNonBlocking));
        for (int i = 0; i < 8; i++) CUDA_TRY(cudaEventCreateWithFlags(&p->hev[i], cudaEventDisableTiming));
    }
    const size_t plane = (size_t)p->g.ny * p->g.nx;
    for (int s = 0; s < slabs; s++) {
        const int z0 = (int)((long)p->g.nz * s / slabs), z1 = (int)((long)p->g.nz * (s + 1) / slabs);
        const size_t off = z0 * plane, cnt = (size_t)(z1 - z0) * plane * sizeof(double);
        CUDA_TRY(cudaMemcpyAsync(p->d_f + off, f_host + off, cnt, cudaMemcpyHostToDevice, p->hcopy));
        CUDA_TRY(cudaEventRecord(p->hev[s], p->hcopy));
        CUDA_TRY(cudaStreamWaitEvent(p->hstream, p->hev[s], 0));
        // a slab of a field is a field of fewer planes: same line tables, its own geometry
        Geometry gs;
        int rc = make_geometry(gs, z1 - z0, p->g.ny, p->g.nx, p->g.axis);
        if (rc) return rc;
        CUtensorMap tin, tout;
        rc = encode_maps(gs, p->d_f + off, p->d_df + off, &tin, &tout);
        if (rc) return rc;
        KParams kp = p->kp;
        kp.nb = gs.nb; kp.rows = gs.nlines; kp.outer = (int)gs.outer;
        kp.halo_lo = kp.halo_hi = kp.ab = nullptr; kp.nlines = gs.nlines;
        rc = gs.contig ? launch_stream<true, true>(gs, kp, tin, tout, p->hstream, &p->pool)
                       : launch_stream<false, true>(gs, kp, tin, tout, p->hstream, &p->pool);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(p->hev[4 + s], p->hstream));
        CUDA_TRY(cudaStreamWaitEvent(p->hback, p->hev[4 + s], 0));
        CUDA_TRY(cudaMemcpyAsync(df_host + off, p->d_df + off, cnt, cudaMemcpyDeviceToHost, p->hback));
    }
    CUDA_TRY(cudaStreamSynchronize(p->hback));
    return CFD_OK;
}

// ------------------------------------------------------------------------------------------------
// near-Toeplitz solver
// ------------------------------------------------------------------------------------------------
template <bool CONTIG, bool REVERSE>
static int launch_recurrence(const Geometry &g, const double *pc, const double *qc, const CUtensorMap &tm_in,
                             const CUtensorMap &tm_out, cudaStream_t stream, PairPool *pool)
{
    static DeviceInfo dinfo;
    if (!dinfo.ok) { int rc = device_info(dinfo); if (rc) return rc; }
    constexpr int per_warp = 4 * SLOT_BYTES + 3 * 16;
    int warps = 4;
    const long per_sm = (g.nb + dinfo.sms - 1) / dinfo.sms;
    if (per_sm < warps) warps = (int)(per_sm < 1 ? 1 : per_sm);
    const size_t smem = (size_t)warps * per_warp + 1024;
    auto kern = recurrence_kernel<CONTIG, REVERSE>;
    static size_t configured[MAX_DEVICES] = {0};
    int dev = 0;
    { int rc = current_device(dev); if (rc) return rc; }
    if (configured[dev] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * per_warp + 1024));
        configured[dev] = 4 * per_warp + 1024;
    }
    RParams rp;
    rp.K = g.K; rp.inner_tiles = g.inner_tiles; rp.nb = g.nb; rp.pc = pc; rp.qc = qc;
    int rc = counter_pair(&rp.counter, stream, pool);
    if (rc) return rc;
    long blocks = (g.nb + warps - 1) / warps;
    if (blocks > dinfo.sms) blocks = dinfo.sms;
    CUDA_TRY(launch_k(kern, (unsigned)blocks, warps * 32, smem, stream, tm_in, tm_out, rp));
    g_launches++;
    return CFD_OK;
}

extern "C" int nt_create(nt_plan **out, int nz, int ny, int nx, int axis, const double coeffs[7])
{
    if (!out || !coeffs) return fail(CFD_EINVAL, "NULL argument");
    *out = nullptr;
    for (int i = 0; i < 7; i++)
        if (!std::isfinite(coeffs[i])) return fail(CFD_EINVAL, "coefficient %d is not finite", i);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(CFD_ECUDA, "no CUDA device: libcfd_b200 has no CPU path");
    }
    nt_plan *p = new nt_plan();
    int rc = make_geometry(p->g, nz, ny, nx, axis);
    if (!rc) rc = ensure_counters();
    if (rc) { delete p; return rc; }
    const LineCoeffs m = {coeffs[0], coeffs[1], coeffs[2], coeffs[3], coeffs[4], coeffs[5], coeffs[6]};
    rc = fill_tables(p->kp, p->g, m, 1.0, true);
    if (rc == CFD_ESLOWPATH) {
        // Coupling too weak for one chunk of look-ahead.  Two chunks (|g|^64 below round-off: e.g. the 6th-order
        // scheme, alpha = 1/3) keep the solve one-pass in the general kernel; anything weaker takes the exact two-pass LU.
        const Pivots pg = build_pivots(p->g.n, as_matrix(m));
        const int la = pg.finite ? lookahead_chunks(pg, p->g.K) : 0;
        if (la == 2 && !getenv("CFD_NO_LA2")) {
            fill_general(p->gp, p->g, pg, 1.0);
            p->la = 2;
            g_err.clear();
            p->kp.lo_closure = 1; p->kp.hi_closure = 1;
            *out = p;
            return CFD_OK;
        }
    }
    if (rc == CFD_ESLOWPATH) {
        // exact two-pass LU: per-row tables on the device
        Pivots pv = build_pivots(p->g.n, m);
        if (!pv.finite) { delete p; return fail(CFD_EINVAL, "zero pivot: LU without pivoting breaks down"); }
        const int L = p->g.K * CH;
        std::vector<double> tab(4 * (size_t)L, 0.0);
        for (int i = 0; i < p->g.n; i++) {
            tab[i] = pv.beta[i]; tab[L + i] = pv.l[i];            // forward
            tab[2 * L + i] = 1.0; tab[3 * L + i] = pv.g[i];       // backward
        }
        if (cudaMalloc(&p->d_tab, tab.size() * sizeof(double)) != cudaSuccess) {
            delete p;
            return fail(CFD_ECUDA, "cudaMalloc of solver tables failed");
        }
        if (cudaMemcpy(p->d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            const cudaError_t e = cudaGetLastError();
            nt_destroy(p);
            return fail(CFD_ECUDA, "upload of solver tables failed: %s", cudaGetErrorString(e));
        }
        p->exact = true;
        g_err.clear();
        rc = CFD_OK;
    }
    if (rc) { delete p; return rc; }
    p->kp.lo_closure = 1; p->kp.hi_closure = 1;
    // (measured, scripts/sweep_solver.py: cutting the lines pays below ~256 bundles; at 512 whole lines win)
    if (!p->exact && p->g.K >= 16 && p->g.nb <= 256) {
        DeviceInfo di;
        if (device_info(di) == CFD_OK) {
            long want = (4L * di.sms * 4 + p->g.nb - 1) / p->g.nb;            // ~4 work items per warp
            if (want > p->g.K / 8) want = p->g.K / 8;                         // segments of >= 8 chunks
            if (want > 1) {
                p->aux_kseg = (int)((p->g.K + want - 1) / want);
                p->aux_nseg = (p->g.K + p->aux_kseg - 1) / p->aux_kseg;
                // the side buffer is a field of the same plane geometry with 2 * nseg chunks per line
                const int n_aux = 2 * p->aux_nseg * CH;
                int rc2;
                if (axis == 0)      rc2 = make_geometry(p->g_aux, nz, ny, n_aux, 0);
                else if (axis == 1) rc2 = make_geometry(p->g_aux, nz, n_aux, nx, 1);
                else                rc2 = make_geometry(p->g_aux, n_aux, ny, nx, 2);
                const size_t bytes = (size_t)p->g.nlines * n_aux * sizeof(double);
                if (rc2 != CFD_OK || cudaMalloc(&p->d_aux, bytes) != cudaSuccess) {
                    cudaGetLastError();
                    g_err.clear();
                    p->d_aux = nullptr;                                        // whole lines instead: slower, still right
                    p->aux_kseg = p->aux_nseg = 0;
                }
            }
        }
    }
    *out = p;
    return CFD_OK;
}

extern "C" int nt_is_exact_two_pass(const nt_plan *p) { return p ? (p->exact ? 1 : 0) : 0; }
extern "C" int nt_lookahead(const nt_plan *p) { return p ? (p->exact ? 0 : (p->la > 0 ? p->la : 1)) : 0; }

extern "C" int nt_solve(nt_plan *p, double *d, void *stream)
{
    if (!p || !d) return fail(CFD_EINVAL, "NULL argument");
    MapPair mp;
    int rc = get_maps(p->cache, p->g, d, d, mp);
    if (rc) return rc;
    if (p->la > 0) return launch_general<0>(p->g, p->gp, p->la, mp.tm_in, mp.tm_out, (cudaStream_t)stream, &p->pool);
    if (p->exact) {
        const long L = (long)p->g.K * CH;
        const double *t = p->d_tab;
        cudaStream_t st = (cudaStream_t)stream;
        if (p->g.contig) {
            rc = launch_recurrence<true, false>(p->g, t, t + L, mp.tm_in, mp.tm_out, st, &p->pool);
            if (rc) return rc;
            return launch_recurrence<true, true>(p->g, t + 2 * L, t + 3 * L, mp.tm_in, mp.tm_out, st, &p->pool);
        }
        rc = launch_recurrence<false, false>(p->g, t, t + L, mp.tm_in, mp.tm_out, st, &p->pool);
        if (rc) return rc;
        return launch_recurrence<false, true>(p->g, t + 2 * L, t + 3 * L, mp.tm_in, mp.tm_out, st, &p->pool);
    }
    KParams kp = p->kp;
    if (p->d_aux) {
        MapPair ma;                                   // only tm_out of this pair is used: the side buffer as a store target
        rc = get_maps(p->aux_cache, p->g_aux, p->d_aux, p->d_aux, ma);
        if (rc) return rc;
        rc = p->g.contig ? launch_stream<true, false>(p->g, kp, mp.tm_in, mp.tm_out, (cudaStream_t)stream, &p->pool, true,
                                                      p->aux_kseg, &ma.tm_out)
                         : launch_stream<false, false>(p->g, kp, mp.tm_in, mp.tm_out, (cudaStream_t)stream, &p->pool, true,
                                                       p->aux_kseg, &ma.tm_out);
        if (rc) return rc;
        const int n_aux = 2 * p->aux_nseg * CH;
        const long total = p->g.nlines * n_aux;
        const int bs = 256;
        scatter_aux_kernel<<<(unsigned)((total + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(
            d, p->d_aux, total, p->g.n, n_aux, p->g.inner, p->aux_kseg, p->g.K);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        return CFD_OK;
    }
    if (p->g.contig) return launch_stream<true, false>(p->g, kp, mp.tm_in, mp.tm_out, (cudaStream_t)stream, &p->pool, true);
    return launch_stream<false, false>(p->g, kp, mp.tm_in, mp.tm_out, (cudaStream_t)stream, &p->pool, true);
}

extern "C" void nt_destroy(nt_plan *p)
{
    if (!p) return;
    cudaFree(p->d_tab);
    cudaFree(p->d_aux);
    delete p;
}

// ------------------------------------------------------------------------------------------------
// pThomas
// ------------------------------------------------------------------------------------------------
struct pt_plan { int n = 0; double *d_lu = nullptr; };

extern "C" int cfd_pthomas_create(pt_plan **out, const double *a, const double *b, const double *c, int n)
{
    if (!out || !a || !b || !c) return fail(CFD_EINVAL, "NULL argument");
    *out = nullptr;
    if (n < 1 || n > 256) return fail(CFD_EINVAL, "n = %d (1..256)", n);
    std::vector<double> lu(3 * n);
    double piv = b[0];
    if (piv == 0.0 || !std::isfinite(piv)) return fail(CFD_EINVAL, "zero pivot at row 0");
    lu[0] = 0.0; lu[n] = 1.0 / piv; lu[2 * n] = c[0] / piv;
    for (int i = 1; i < n; i++) {
        piv = b[i] - a[i] * lu[2 * n + i - 1];
        if (piv == 0.0 || !std::isfinite(piv)) return fail(CFD_EINVAL, "zero pivot at row %d", i);
        lu[i] = a[i]; lu[n + i] = 1.0 / piv; lu[2 * n + i] = c[i] / piv;
    }
    pt_plan *p = new pt_plan();
    p->n = n;
    if (cudaMalloc(&p->d_lu, lu.size() * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(p->d_lu, lu.data(), lu.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        cudaFree(p->d_lu);
        delete p;
        return fail(CFD_ECUDA, "pThomas tables: %s", cudaGetErrorString(e));
    }
    *out = p;
    return CFD_OK;
}

extern "C" int cfd_pthomas_solve(pt_plan *p, double *d, long nsys, void *stream)
{
    if (!p || !d) return fail(CFD_EINVAL, "NULL argument");
    if (nsys < 1) return fail(CFD_EINVAL, "nsys = %ld", nsys);
    const int bs = 128;
    pthomas_kernel<<<(unsigned)((nsys + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(d, p->d_lu, p->n, nsys);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CFD_OK;
}

extern "C" void cfd_pthomas_destroy(pt_plan *p)
{
    if (!p) return;
    cudaFree(p->d_lu);
    delete p;
}

// One-shot convenience form (create + solve + destroy): allocates, and synchronises the stream before it frees.
extern "C" int cfd_pthomas(const double *a, const double *b, const double *c, double *d, int n, long nsys, void *stream)
{
    pt_plan *p = nullptr;
    int rc = cfd_pthomas_create(&p, a, b, c, n);
    if (rc) return rc;
    rc = cfd_pthomas_solve(p, d, nsys, stream);
    if (!rc && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
        rc = fail(CFD_ECUDA, "cfd_pthomas: %s", cudaGetErrorString(cudaGetLastError()));
    cfd_pthomas_destroy(p);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Partitioned line over NVLink peer memory, host side in C (SURVEY 8b cfd_mg_*): one cfd_zpart per rank owns the
// rank's receive buffer, maps the neighbours' (cudaIpc handles between processes, plain pointers inside one), and
// issues the launches of the fused multi-rank derivative.  Buffer layout, in doubles (plane = lines of the block):
//     halo  [2 parities][2][plane]   slot 0: last plane of rank-1, slot 1: first plane of rank+1
//     faces [2 parities][6][plane]   the neighbour-only interface planes (cfd_reduced_unknowns layout)
//     flags [16] (uint64)            2 / 3: the left / right neighbour's faces + halo of call `seq` have landed
//     ll    [2 parities][4][plane]   (16-byte words) receive arrays of the one-kernel path (kernels_zx.cuh): 0 / 1 = the left
//                                    neighbour's hi face / last row, 2 / 3 = the right neighbour's lo face / first row
// The parity is seq & 1 and flags only grow, so nothing is ever reset and no global barrier exists: a neighbour can
// be at most one call ahead (its call s+1 needs our producer launch of call s+1, which is stream-ordered after our
// consumer of call s), and then it writes the other parity.
// ------------------------------------------------------------------------------------------------
struct cfd_zpart {
    cfd_plan *plan = nullptr;
    long plane = 0;
    double *buf = nullptr;
    double *peer[2] = {nullptr, nullptr};
    bool opened[2] = {false, false};
    double *ab = nullptr;
    unsigned long long seq = 0;
    cudaEvent_t ev_begin = nullptr, ev_apply = nullptr;
    bool applied = false, pending = false;
    const double *pending_f = nullptr;
    cudaStream_t pending_stream = nullptr;
    int max_ctas = 0;                 // 0 = one CTA per SM; tests that keep several ranks on ONE device lower it
    double *halo(double *base, int par, int slot) const { return base + (long)(par * 2 + slot) * plane; }
    ulonglong2 *ll(double *base, int par, int i) const
    { return reinterpret_cast<ulonglong2 *>(base + 16 * plane + 16) + (long)(par * 4 + i) * plane; }
    double *faces(double *base, int par, int i) const { return base + 4 * plane + (long)(par * 6 + i) * plane; }
    unsigned long long *flag(double *base, int k) const { return reinterpret_cast<unsigned long long *>(base + 16 * plane) + k; }
};

// With lazy module loading (the CUDA default) the FIRST launch of a kernel loads its code, and that load can need the
// device to go idle.  A process that drives several ranks of a line on one device (tests, smoke) would then hang until
// the time-out: rank 0's consumer kernel spins for data that rank 1's producer -- whose first launch is stuck loading --
// never gets to store.  Every kernel the partitioned paths launch is therefore loaded when the first zpart is created.
static void preload_zpart_kernels()
{
    static bool done = false;
    if (done) return;
    cudaFuncAttributes at;
    cudaFuncGetAttributes(&at, edge_faces_kernel);
    cudaFuncGetAttributes(&at, reduced_planes_kernel);
    cudaFuncGetAttributes(&at, reduced_planes_deferred_kernel);
    cudaFuncGetAttributes(&at, push_planes_kernel);
    cudaFuncGetAttributes(&at, wait_flags_kernel);
    cudaFuncGetAttributes(&at, npts_edge_kernel<0>);
    cudaFuncGetAttributes(&at, npts_edge_kernel<1>);
    cudaFuncGetAttributes(&at, stream_kernel<false, true, 3>);
    cudaFuncGetAttributes(&at, stream_kernel<false, true, 4>);
    cudaFuncGetAttributes(&at, stream_kernel<false, true, 5>);
    cudaFuncGetAttributes(&at, stream_kernel<true, true, 3>);
    cudaFuncGetAttributes(&at, stream_kernel_zx<4>);
    cudaFuncGetAttributes(&at, stream_kernel_xy<4, false, false>);
    cudaFuncGetAttributes(&at, stream_kernel_xy<4, true, false>);
    cudaFuncGetAttributes(&at, stream_kernel_xy<4, false, true>);
    cudaFuncGetAttributes(&at, stream_kernel_xy<4, true, true>);
    cudaFuncGetAttributes(&at, stream_kernel_xy<3, false, false>);
    cudaFuncGetAttributes(&at, stream_kernel_xy<3, true, false>);
    cudaGetLastError();
    done = true;
}

extern "C" int cfd_zpart_create(cfd_zpart **out, cfd_plan *plan)
{
    if (!out || !plan) return fail(CFD_EINVAL, "NULL argument");
    *out = nullptr;
    preload_zpart_kernels();
    if (plan->size < 2) return fail(CFD_EINVAL, "plan has part_size 1: nothing to exchange");
    if (plan->scheme != CFD_SCHEME_PADE4) return fail(CFD_EUNSUPPORTED, "cfd_zpart serves the Pade-4 scheme");
    if (plan->g.axis != 2) return fail(CFD_EUNSUPPORTED, "cfd_zpart serves lines along z (contiguous boundary planes)");
    if (plan->g.n < 2 * CH + 2)
        return fail(CFD_EUNSUPPORTED, "cfd_zpart needs >= %d planes per slab (have %d)", 2 * CH + 2, plan->g.n);
    if (plan->g.nlines % 2) return fail(CFD_EINVAL, "plane size must be even");
    cfd_zpart *z = new cfd_zpart();
    z->plan = plan;
    z->plane = plan->g.nlines;
    const size_t bytes = (size_t)(32 * z->plane + 16) * sizeof(double);
    if (cudaMalloc(&z->buf, bytes) != cudaSuccess || cudaMemset(z->buf, 0, bytes) != cudaSuccess ||
        cudaMalloc(&z->ab, (size_t)2 * z->plane * sizeof(double)) != cudaSuccess ||
        cudaEventCreateWithFlags(&z->ev_begin, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&z->ev_apply, cudaEventDisableTiming) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        cfd_zpart_destroy(z);
        return fail(CFD_ECUDA, "cfd_zpart_create: %s", cudaGetErrorString(e));
    }
    *out = z;
    return CFD_OK;
}

extern "C" int cfd_zpart_export(cfd_zpart *z, void *handle)
{
    if (!z || !handle) return fail(CFD_EINVAL, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == CFD_IPC_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, z->buf));
    memcpy(handle, &h, sizeof h);
    return CFD_OK;
}

extern "C" void *cfd_zpart_buffer(cfd_zpart *z) { return z ? z->buf : nullptr; }

static int zpart_check_sides(cfd_zpart *z, const void *lo, const void *hi)
{
    const bool has_lo = z->plan->rank > 0, has_hi = z->plan->rank < z->plan->size - 1;
    if (has_lo != (lo != nullptr) || has_hi != (hi != nullptr))
        return fail(CFD_EINVAL, "rank %d of %d: neighbour buffers must be given exactly where a neighbour exists",
                    z->plan->rank, z->plan->size);
    return CFD_OK;
}

static void zpart_disconnect(cfd_zpart *z)
{
    for (int i = 0; i < 2; i++) {
        if (z->opened[i] && z->peer[i]) cudaIpcCloseMemHandle(z->peer[i]);
        z->peer[i] = nullptr;
        z->opened[i] = false;
    }
}

extern "C" int cfd_zpart_connect(cfd_zpart *z, const void *handle_lo, const void *handle_hi)
{
    if (!z) return fail(CFD_EINVAL, "NULL argument");
    int rc = zpart_check_sides(z, handle_lo, handle_hi);
    if (rc) return rc;
    zpart_disconnect(z);
    const void *hs[2] = {handle_lo, handle_hi};
    for (int i = 0; i < 2; i++) {
        if (!hs[i]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs[i], sizeof h);
        void *ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            zpart_disconnect(z);
            return fail(CFD_ECUDA, "cudaIpcOpenMemHandle (%s neighbour): %s", i ? "right" : "left", cudaGetErrorString(e));
        }
        z->peer[i] = (double *)ptr;
        z->opened[i] = true;
    }
    return CFD_OK;
}

extern "C" int cfd_zpart_connect_ptr(cfd_zpart *z, void *buffer_lo, void *buffer_hi)
{
    if (!z) return fail(CFD_EINVAL, "NULL argument");
    int rc = zpart_check_sides(z, buffer_lo, buffer_hi);
    if (rc) return rc;
    zpart_disconnect(z);
    z->peer[0] = (double *)buffer_lo;
    z->peer[1] = (double *)buffer_hi;
    return CFD_OK;
}

extern "C" void cfd_zpart_destroy(cfd_zpart *z)
{
    if (!z) return;
    zpart_disconnect(z);
    cudaFree(z->buf);
    cudaFree(z->ab);
    if (z->ev_begin) cudaEventDestroy(z->ev_begin);
    if (z->ev_apply) cudaEventDestroy(z->ev_apply);
    delete z;
}

// pointers of call `seq` on this rank
struct ZPtrs {
    double *faces_nb, *own_faces, *peer_face_lo, *peer_face_hi, *push_lo, *push_hi, *halo_lo, *halo_hi;
    unsigned long long *flag_lo, *flag_hi, *wait_lo, *wait_hi;
};

static int zpart_ptrs(cfd_zpart *z, unsigned long long seq, ZPtrs &q)
{
    const cfd_plan *p = z->plan;
    const bool has_lo = p->rank > 0, has_hi = p->rank < p->size - 1;
    if ((has_lo && !z->peer[0]) || (has_hi && !z->peer[1]))
        return fail(CFD_EINVAL, "cfd_zpart: neighbours are not connected (cfd_zpart_connect)");
    const int par = (int)(seq & 1), own = p->nb_own;
    const int own_left = (has_lo && p->rank - 1 > 0) ? 1 : 0;       // index of the left neighbour among ITS virtual ranks
    q.faces_nb = z->faces(z->buf, par, 0);
    q.own_faces = z->faces(z->buf, par, 2 * own);
    q.peer_face_lo = has_lo ? z->faces(z->peer[0], par, 2 * own_left + 2) : nullptr;   // its "right neighbour's faces[0]"
    q.peer_face_hi = has_hi ? z->faces(z->peer[1], par, 1) : nullptr;                  // its "left neighbour's faces[1]"
    q.push_lo = has_lo ? z->halo(z->peer[0], par, 1) : nullptr;
    q.push_hi = has_hi ? z->halo(z->peer[1], par, 0) : nullptr;
    q.flag_lo = has_lo ? z->flag(z->peer[0], 3) : nullptr;
    q.flag_hi = has_hi ? z->flag(z->peer[1], 2) : nullptr;
    q.halo_lo = has_lo ? z->halo(z->buf, par, 0) : nullptr;
    q.halo_hi = has_hi ? z->halo(z->buf, par, 1) : nullptr;
    q.wait_lo = has_lo ? z->flag(z->buf, 2) : nullptr;
    q.wait_hi = has_hi ? z->flag(z->buf, 3) : nullptr;
    return CFD_OK;
}

extern "C" int cfd_zpart_set_ctas(cfd_zpart *z, int max_ctas)
{
    if (!z || max_ctas < 0) return fail(CFD_EINVAL, "bad argument");
    z->max_ctas = max_ctas;
    return CFD_OK;
}

// The partitioned d/dz in ONE launch (kernels_zx.cuh): edge faces, exchange, reduced system and coupled solve per bundle.
static int launch_zx(cfd_zpart *z, const double *f, double *df, unsigned long long seq, cudaStream_t stream)
{
    static DeviceInfo dinfo;
    if (!dinfo.ok) { int rc = device_info(dinfo); if (rc) return rc; }
    cfd_plan *p = z->plan;
    const bool has_lo = p->rank > 0, has_hi = p->rank < p->size - 1;
    if ((has_lo && !z->peer[0]) || (has_hi && !z->peer[1]))
        return fail(CFD_EINVAL, "cfd_zpart: neighbours are not connected (cfd_zpart_connect)");
    MapPair mp;
    int rc = get_maps(p->cache, p->g, f, df, mp);
    if (rc) return rc;
    KParams kp = p->kp;
    kp.halo_lo = kp.halo_hi = kp.ab = nullptr;
    kp.nlines = p->g.nlines;
    kp.kseg = p->g.K; kp.nseg = 1;
    ZXParams q;
    memset(&q, 0, sizeof q);
    const int par = (int)(seq & 1);
    q.in_face_lo = z->ll(z->buf, par, 0); q.in_halo_lo = z->ll(z->buf, par, 1);
    q.in_face_hi = z->ll(z->buf, par, 2); q.in_halo_hi = z->ll(z->buf, par, 3);
    if (has_lo) { q.out_face_lo = z->ll(z->peer[0], par, 2); q.out_halo_lo = z->ll(z->peer[0], par, 3); }
    if (has_hi) { q.out_face_hi = z->ll(z->peer[1], par, 0); q.out_halo_hi = z->ll(z->peer[1], par, 1); }
    q.tag = (unsigned int)(seq & 0xffffffffULL);
    if (q.tag == 0) q.tag = 0x80000000u;                  // 0 is what a fresh buffer holds
    q.pv = p->nb_pv; q.own = p->nb_own;
    q.w_lo = p->w_lo; q.w_hi = p->w_hi;
    q.sk_last = p->kp.tail.sk[p->g.jl]; q.l_last = p->kp.tail.l[p->g.jl];
    if (p->lu_nb.size() > 36) return fail(CFD_EINVAL, "internal: neighbour table too large");
    for (size_t i = 0; i < p->lu_nb.size(); i++) q.lu[i] = p->lu_nb[i];
    q.wait = wait_params();
    q.hints = 7;
    if (const char *e = getenv("CFD_ZX_HINTS")) q.hints = atoi(e);
    constexpr int NS = 4;
    constexpr int per_warp = NS * SLOT_BYTES + NS * 16;
    int warps = g_warps ? g_warps : 7;
    if (const char *e = getenv("CFD_ZX_WARPS")) warps = atoi(e) > 0 ? atoi(e) : warps;
    if (warps > 7) warps = 7;
    const long per_sm = (p->g.nb + dinfo.sms - 1) / dinfo.sms;
    if (!g_warps && per_sm < warps) warps = (int)(per_sm < 1 ? 1 : per_sm);
    const size_t smem = (size_t)warps * per_warp + 1024;
    auto kern = stream_kernel_zx<NS>;
    static size_t configured[MAX_DEVICES] = {0};
    int dev = 0;
    rc = current_device(dev);
    if (rc) return rc;
    if (configured[dev] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * per_warp + 1024));
        configured[dev] = 7 * per_warp + 1024;
    }
    rc = counter_pair(&kp.counter, stream, &p->pool);
    if (rc) return rc;
    long blocks = (p->g.nb + warps - 1) / warps;
    if (blocks > dinfo.sms) blocks = dinfo.sms;
    if (z->max_ctas > 0 && blocks > z->max_ctas) blocks = z->max_ctas;
    CUDA_TRY(launch_k(kern, (unsigned)blocks, warps * 32, smem, stream, mp.tm_in, mp.tm_out, kp, q));
    g_launches++;
    return CFD_OK;
}

static bool zx_enabled() { return getenv("CFD_NO_ZX") == nullptr; }

extern "C" int cfd_zpart_begin(cfd_zpart *z, const double *f, void *stream)
{
    if (!z || !f) return fail(CFD_EINVAL, "NULL argument");
    int rc = cfd_async_status();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // the producer of call s+1 overwrites what the consumer of call s-1 read (same parity): order it after the last
    // coupled launch even when the caller runs begin() on a side stream -- and after an exchange that was begun on
    // another stream and never picked up, so that the producers of consecutive calls finish in order
    if (z->applied) CUDA_TRY(cudaStreamWaitEvent(st, z->ev_apply, 0));
    if (z->pending && z->pending_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, z->ev_begin, 0));
    ZPtrs q;
    rc = zpart_ptrs(z, z->seq + 1, q);
    if (rc) return rc;
    const unsigned long long seq = ++z->seq;
    rc = cfd_edge_faces_push(z->plan, f, q.own_faces, q.peer_face_lo, q.peer_face_hi, q.push_lo, q.push_hi, q.flag_lo,
                             q.flag_hi, seq, stream);
    if (rc) return rc;
    rc = cfd_reduced_unknowns_deferred(z->plan, q.faces_nb, q.halo_lo, q.halo_hi, f, z->ab, q.wait_lo, q.wait_hi, seq, stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(z->ev_begin, st));
    z->pending = true;
    z->pending_f = f;
    z->pending_stream = st;
    return CFD_OK;
}

extern "C" int cfd_zpart_apply(cfd_zpart *z, const double *f, double *df, void *stream)
{
    if (!z || !f || !df) return fail(CFD_EINVAL, "NULL argument");
    if (f == df) return fail(CFD_EINVAL, "the derivative is out of place: f and df must differ");
    cudaStream_t st = (cudaStream_t)stream;
    if (!(z->pending && z->pending_f == f) && zx_enabled()) {
        // nothing was begun early: the whole partitioned derivative is one launch
        int rc = cfd_async_status();
        if (rc) return rc;
        if (z->applied) CUDA_TRY(cudaStreamWaitEvent(st, z->ev_apply, 0));
        if (z->pending && z->pending_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, z->ev_begin, 0));
        z->pending = false;
        rc = launch_zx(z, f, df, z->seq + 1, st);
        if (rc) return rc;
        ++z->seq;
        CUDA_TRY(cudaEventRecord(z->ev_apply, st));
        z->applied = true;
        return CFD_OK;
    }
    if (!(z->pending && z->pending_f == f)) {
        int rc = cfd_zpart_begin(z, f, stream);
        if (rc) return rc;
    }
    if (z->pending_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, z->ev_begin, 0));
    z->pending = false;
    ZPtrs q;
    int rc = zpart_ptrs(z, z->seq, q);
    if (rc) return rc;
    rc = cfd_apply_coupled(z->plan, f, df, q.halo_lo, q.halo_hi, z->ab, stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(z->ev_apply, st));
    z->applied = true;
    return CFD_OK;
}

// The partitioned d/dz by the distributed npts method (kernels.cuh npts_edge_kernel): halo planes, u~ to the right,
// x~ to the left, coupled one-pass solve with this rank's slice of the global LU.  An alternative to the reduced-system
// method of cfd_zpart_apply (the measured path), kept for the reference's second algorithm; `zp` must have been
// created from a cfd_create_npts plan.
extern "C" int cfd_zpart_apply_npts(cfd_zpart *z, const double *f, double *df, void *stream)
{
    if (!z || !f || !df) return fail(CFD_EINVAL, "NULL argument");
    if (f == df) return fail(CFD_EINVAL, "the derivative is out of place: f and df must differ");
    cfd_plan *p = z->plan;
    if (!p->npts) return fail(CFD_EINVAL, "cfd_zpart_apply_npts needs a zpart of a cfd_create_npts plan");
    int rc = cfd_async_status();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (z->applied) CUDA_TRY(cudaStreamWaitEvent(st, z->ev_apply, 0));
    const bool has_lo = p->rank > 0, has_hi = p->rank < p->size - 1;
    if ((has_lo && !z->peer[0]) || (has_hi && !z->peer[1]))
        return fail(CFD_EINVAL, "cfd_zpart: neighbours are not connected (cfd_zpart_connect)");
    const unsigned long long seq = ++z->seq;
    const int par = (int)(seq & 1);
    const long plane = z->plane, n = p->g.n;
    // (1) halo planes of f: first plane -> left neighbour's halo slot 1, last plane -> right neighbour's slot 0
    rc = cfd_push_planes(has_lo ? f : nullptr, has_lo ? z->halo(z->peer[0], par, 1) : nullptr,
                         has_hi ? f + (n - 1) * plane : nullptr, has_hi ? z->halo(z->peer[1], par, 0) : nullptr, plane,
                         has_lo ? z->flag(z->peer[0], 1) : nullptr, has_hi ? z->flag(z->peer[1], 0) : nullptr, seq, stream);
    if (rc) return rc;
    rc = cfd_wait_flags(has_lo ? z->flag(z->buf, 0) : nullptr, has_hi ? z->flag(z->buf, 1) : nullptr, seq, stream);
    if (rc) return rc;
    double *halo_lo = has_lo ? z->halo(z->buf, par, 0) : nullptr, *halo_hi = has_hi ? z->halo(z->buf, par, 1) : nullptr;
    EdgeP ep;
    memset(&ep, 0, sizeof ep);
    ep.nlines = p->g.nlines; ep.inner = p->g.inner; ep.n = p->g.n; ep.jl = p->g.jl;
    ep.lo_closure = p->kp.lo_closure; ep.hi_closure = p->kp.hi_closure;
    ep.sk_mid = p->kp.sk_mid; ep.l_mid = p->kp.l_mid;
    ep.sk_last = p->kp.tail.sk[p->g.jl]; ep.l_last = p->kp.tail.l[p->g.jl];
    ep.halo_lo = halo_lo; ep.halo_hi = halo_hi;
    ep.head = p->kp.head;
    ep.seq = seq;
    const int bs = 128;
    const unsigned grid = (unsigned)((ep.nlines + bs - 1) / bs);
    // u~ / x~ live in the interface planes 0 / 1 of this call's parity, adjacent: exactly the `ab` planes the coupled
    // kernel reads
    double *ab = z->faces(z->buf, par, 0);
    if (has_hi) {                                            // (2) u at our last row -> the right neighbour's plane 0
        rc = counter_pair(&ep.done, st, &p->pool);
        if (rc) return rc;
        npts_edge_kernel<0><<<grid, bs, 0, st>>>(f, ep, nullptr, z->faces(z->peer[1], par, 0), z->flag(z->peer[1], 4));
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    if (has_lo) {                                            // (3) x at our first row -> the left neighbour's plane 1
        rc = cfd_wait_flags(z->flag(z->buf, 4), nullptr, seq, stream);
        if (rc) return rc;
        rc = counter_pair(&ep.done, st, &p->pool);
        if (rc) return rc;
        npts_edge_kernel<1><<<grid, bs, 0, st>>>(f, ep, ab, z->faces(z->peer[0], par, 1), z->flag(z->peer[0], 5));
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    } else {
        CUDA_TRY(cudaMemsetAsync(ab, 0, plane * sizeof(double), st));            // rank 0: no incoming u~
    }
    if (has_hi) {
        rc = cfd_wait_flags(z->flag(z->buf, 5), nullptr, seq, stream);
        if (rc) return rc;
    } else {
        CUDA_TRY(cudaMemsetAsync(ab + plane, 0, plane * sizeof(double), st));    // last rank: no incoming x~
    }
    // (4) both sweeps in one pass
    rc = apply_impl(p, f, df, halo_lo, halo_hi, ab, stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(z->ev_apply, st));
    z->applied = true;
    return CFD_OK;
}

// All three derivatives of a z-slab in three launches: the fused d/dx + d/dy kernel with the edge-face items of the
// partitioned d/dz drawn first (faces and halo rows travel over NVLink while x / y are computed), the reduced
// 2x2 / 6-unknown solve per line, the coupled d/dz kernel.
extern "C" int cfd_zpart_apply_xyz(cfd_zpart *z, cfd_plan *px, cfd_plan *py, const double *f, double *dfdx, double *dfdy,
                                   double *dfdz, void *stream)
{
    if (!z || !px || !py || !f || !dfdx || !dfdy || !dfdz) return fail(CFD_EINVAL, "NULL argument");
    cfd_plan *pz = z->plan;
    if (px->g.axis != 0 || py->g.axis != 1) return fail(CFD_EINVAL, "plans must be for axis 0 (x) and axis 1 (y)");
    if (px->size != 1 || py->size != 1) return fail(CFD_EINVAL, "x and y lines of a z-slab are unpartitioned plans");
    for (const cfd_plan *p : {px, py})
        if (p->g.nz != pz->g.nz || p->g.ny != pz->g.ny || p->g.nx != pz->g.nx) return fail(CFD_EINVAL, "plans are for different shapes");
    if (f == dfdx || f == dfdy || f == dfdz || dfdx == dfdy || dfdx == dfdz || dfdy == dfdz)
        return fail(CFD_EINVAL, "f and the three derivatives must be four different fields");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (zx_enabled()) {                                // two launches: the plain x/y launch, the one-kernel d/dz
        // The two are independent (same input, different outputs).  CFD_ZX_TWO_STREAMS=1: d/dz on plan_z's side stream,
        // forked and joined with events like cfd_apply_xyz, so that one kernel's CTAs take over as the other's retire.
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
        const bool two = getenv("CFD_ZX_TWO_STREAMS") != nullptr && !(z->pending && z->pending_f == f);
        if (two && !pz->gside && cs == cudaStreamCaptureStatusNone) {
            if (cudaStreamCreateWithFlags(&pz->gside, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&pz->gev_fork, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&pz->gev_join, cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                if (pz->gside) { cudaStreamDestroy(pz->gside); pz->gside = nullptr; }
            }
        }
        if (two && pz->gside && pz->gev_fork && pz->gev_join) {
            CUDA_TRY(cudaEventRecord(pz->gev_fork, st));
            CUDA_TRY(cudaStreamWaitEvent(pz->gside, pz->gev_fork, 0));
            rc = cfd_zpart_apply(z, f, dfdz, pz->gside);
            if (!rc) rc = cfd_apply_xy(px, py, f, dfdx, dfdy, stream);
            if (rc) return rc;
            CUDA_TRY(cudaEventRecord(pz->gev_join, pz->gside));
            CUDA_TRY(cudaStreamWaitEvent(st, pz->gev_join, 0));
            return CFD_OK;
        }
        rc = cfd_apply_xy(px, py, f, dfdx, dfdy, stream);
        return rc ? rc : cfd_zpart_apply(z, f, dfdz, stream);
    }
    if (!xy_eligible(px->g) || getenv("CFD_NO_XY") || getenv("CFD_NO_XY_EDGE") || (g_slots && g_slots != 4)) {
        rc = cfd_zpart_begin(z, f, stream);            // separate edge launch, then the x / y launch(es)
        if (!rc) rc = cfd_apply_xy(px, py, f, dfdx, dfdy, stream);
        if (!rc) rc = cfd_zpart_apply(z, f, dfdz, stream);
        return rc;
    }
    rc = cfd_async_status();
    if (rc) return rc;
    if (z->applied) CUDA_TRY(cudaStreamWaitEvent(st, z->ev_apply, 0));
    if (z->pending && z->pending_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, z->ev_begin, 0));   // see cfd_zpart_begin
    ZPtrs q;
    rc = zpart_ptrs(z, z->seq + 1, q);
    if (rc) return rc;
    MapPair mx, my, mz;
    rc = get_maps(px->cache, px->g, f, dfdx, mx);
    if (!rc) rc = get_maps(py->cache, py->g, f, dfdy, my);
    if (!rc) rc = get_maps(pz->cache, pz->g, f, dfdz, mz);
    if (rc) return rc;
    const unsigned long long seq = ++z->seq;
    EdgeX ex;
    memset(&ex, 0, sizeof ex);
    ex.nedge = (pz->g.nlines + CH - 1) / CH;
    ex.nlines = pz->g.nlines;
    ex.n = pz->g.n;
    ex.has_lo = pz->kp.lo_closure ? 0 : 1;
    ex.has_hi = pz->kp.hi_closure ? 0 : 1;
    ex.sk_mid = pz->kp.sk_mid; ex.l_mid = pz->kp.l_mid;
    ex.sk_last = pz->kp.tail.sk[pz->g.jl]; ex.l_last = pz->kp.tail.l[pz->g.jl];
    ex.f = f;
    ex.faces = q.own_faces;
    ex.peer_face_lo = q.peer_face_lo; ex.peer_face_hi = q.peer_face_hi;
    ex.push_lo = q.push_lo; ex.push_hi = q.push_hi;
    ex.flag_lo = q.flag_lo; ex.flag_hi = q.flag_hi;
    ex.seq = seq;
    ex.head = pz->kp.head;
    rc = counter_pair(&ex.done, st, &pz->pool);
    if (rc) return rc;
    const long nitems = (long)px->g.nz * (px->g.ny / CH + py->g.inner_tiles);
    rc = launch_xy<4>(px, py, mx, my, nitems, st, &ex, &mz.tm_in);
    if (rc) return rc;
    rc = cfd_reduced_unknowns_deferred(pz, q.faces_nb, q.halo_lo, q.halo_hi, f, z->ab, q.wait_lo, q.wait_hi, seq, stream);
    if (rc) return rc;
    z->pending = false;
    rc = cfd_apply_coupled(pz, f, dfdz, q.halo_lo, q.halo_hi, z->ab, stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(z->ev_apply, st));
    z->applied = true;
    return CFD_OK;
}
