/*
 * cfd_b200.h -- C ABI of libcfd_b200.so: the B200-native compact finite-difference derivative path.
 *
 * The reference (ashwinsrnth/compact-finite-differences) has no C ABI: its boundary is a set of Python
 * classes that launch runtime-compiled `extern "C"` kernels through PyCUDA `prepared_call`
 * (code/cuda/kernels.py:14-23).  This header is the boundary a maintainer of the reference would bind
 * instead (ctypes stub in INTEGRATION.md).  Each entry point names the reference interface it replaces.
 *
 * Conventions (kept from the reference):
 *   - fields are fp64, C order f[nz][ny][nx] (x fastest); axis numbering 0 = x, 1 = y, 2 = z
 *     (code/cuda/gpuDA.py:162);
 *   - every field pointer is a DEVICE pointer owned by the caller (code/cuda/compact.py:29); plans own only
 *     small coefficient tables (code/cuda/solvers/templated/near_toeplitz.py:62-69);
 *   - the tridiagonal solve is in place (near_toeplitz.py:78); the derivative is out of place (f -> df);
 *   - calls are asynchronous on the given CUDA stream (a cudaStream_t passed as void*; NULL = default
 *     stream), never allocate, never synchronise -- except the *_host convenience calls and the one-shot
 *     cfd_pthomas, which are synchronous and own their staging buffers -- and may be captured into CUDA graphs
 *     (a captured launch gets work counters of its own, so replays never collide with eager launches);
 *   - every function returns 0 on success or a negative CFD_E* code; cfd_last_error() gives the text.
 *     Nothing throws, nothing exits.  There is no CPU fallback: without a CUDA device create() fails
 *     with CFD_ECUDA.
 *
 * Constraints (narrower in the reference: nx a power of two <= 2048, extents multiples of 8/16):
 *   n (extent along the axis) >= 3; nx even (16-byte row pitch for TMA); pointers 16-byte aligned.
 */
#ifndef CFD_B200_H
#define CFD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CFD_B200_VERSION 200

#define CFD_OK            0
#define CFD_EINVAL       (-1)   /* bad shape / axis / pointer / coefficient                    */
#define CFD_ECUDA        (-2)   /* CUDA runtime or driver error (text in cfd_last_error)        */
#define CFD_EUNSUPPORTED (-3)   /* valid request this build cannot serve                        */
#define CFD_ETIMEOUT     (-4)   /* a kernel gave up waiting for a neighbour rank (cfd_async_status) */

typedef struct cfd_plan cfd_plan;   /* derivative operator for one (shape, axis, spacing, rank position) */
typedef struct nt_plan nt_plan;     /* batched near-Toeplitz tridiagonal solver                          */
typedef struct pt_plan pt_plan;     /* thread-parallel Thomas for one general tridiagonal matrix         */
typedef struct cfd_zpart cfd_zpart; /* one rank of a z-partitioned line: peer buffers + launch sequence  */

int cfd_version(void);
const char *cfd_last_error(void);   /* thread-local text of the last failure */

/* ---------------------------------------------------------------------------------------------------
 * Derivative operator.
 * Replaces CompactFiniteDifferenceSolver.__init__ (code/cuda/compact.py:18-27, coefficient choice
 * :159-173) for one direction.  (part_rank, part_size) is the position of this block along the
 * derivative line, exactly the reference's (line_da.rank, line_da.size): rank 0 carries the left
 * closure row [1 2], rank size-1 the right closure row [2 1], other block ends are the cut Toeplitz
 * rows.  part_size == 1 is the single-GPU operator.
 * ------------------------------------------------------------------------------------------------- */
int cfd_create(cfd_plan **plan, int nz, int ny, int nx, int axis, double h, int part_rank, int part_size);
void cfd_destroy(cfd_plan *plan);

/* Other compact schemes through the same one-pass streaming solve (the reference's solver takes any
 * [b1,c1,ai,bi,ci,an,bn], code/cuda/solvers/templated/near_toeplitz.py:49-50, its derivative operator only Pade-4):
 *   CFD_SCHEME_PADE4     the reference's scheme (= cfd_create with part (0, 1));
 *   CFD_SCHEME_COMPACT6  6th-order tridiagonal first derivative (Lele 1992 eq. 2.1, alpha = 1/3), 4th-order Pade rows
 *                        next to the reference's 3rd-order closure rows; interior coupling 0.382 per row, so the
 *                        backward sweep looks TWO chunks ahead (0.382^64 = 2e-27) -- still one pass, 16 B / point;
 *   CFD_SCHEME_PADE4_D2  4th-order SECOND derivative (eq. 2.2, alpha = 1/10) with the 3rd-order closure
 *                        f''_0 + 11 f''_1 = (13 f_0 - 27 f_1 + 15 f_2 - f_3) / h^2.
 * Plans of these schemes serve cfd_apply (and cfd_apply_host) on unpartitioned lines; cfd_plan_lookahead() = chunks of
 * look-ahead the plan's matrix needs (1 or 2), derived from its coefficients. */
#define CFD_SCHEME_PADE4    0
#define CFD_SCHEME_COMPACT6 1
#define CFD_SCHEME_PADE4_D2 2
int cfd_create_scheme(cfd_plan **plan, int nz, int ny, int nx, int axis, double h, int scheme);
int cfd_plan_lookahead(const cfd_plan *plan);

/* df = local solution x_R of the block: Pade RHS (code/cuda/kernels.cu:4-47 computeRHS) fused with the
 * block's tridiagonal solve (NearToeplitzSolver.solve, templated/near_toeplitz.py:78-107) -- one kernel,
 * f read once, df written once.  halo_lo / halo_hi: the neighbour's boundary plane of f (what
 * DA.global_to_local, code/cuda/gpuDA.py:61-132, puts in the ghost layer), one value per line, laid out
 * like a plane of the field with the axis removed; NULL where the block owns the physical end.
 * With part_size == 1 this IS the derivative (replaces dfdx, compact.py:29-44; dfdy/dfdz,
 * code/ocl/compact.py:41-61). */
int cfd_apply(cfd_plan *plan, const double *f, double *df, const double *halo_lo, const double *halo_hi,
              void *stream);

/* d/dx and d/dy of the same field in ONE launch (plans for axis 0 and axis 1 of the same shape, part_size 1).
 * Each derivative is the same one-pass solve as cfd_apply; the work of the two is interleaved plane by plane so
 * that every tile of f comes from HBM once and from L2 the second time.  Replaces two dfdx / dfdy calls of the
 * reference (code/ocl/compact.py:26-50).  Falls back to two launches when ny is not a multiple of 32. */
int cfd_apply_xy(cfd_plan *plan_x, cfd_plan *plan_y, const double *f, double *dfdx, double *dfdy, void *stream);
/* d/dx, d/dy and d/dz of the same unpartitioned field (plans for axes 0, 1, 2 of the same shape): cfd_apply_xy on
 * `stream` and the d/dz launch on a side stream owned by plan_z, forked and joined with events, so that the second
 * kernel's CTAs take over the SMs as the first one's retire instead of each launch paying its own ramp-up and tail.
 * On return everything is ordered on `stream` as if the three launches had been issued there. */
int cfd_apply_xyz(cfd_plan *plan_x, cfd_plan *plan_y, cfd_plan *plan_z, const double *f, double *dfdx, double *dfdy,
                  double *dfdz, void *stream);
/* Warps per SM of that launch for an axis-0 plan.  0 = the library's choice: 6, or 7 for short launches (at most 72
 * work items per SM and lines of fewer than 32 tiles, e.g. 256^3).  5 leaves room on the SMs for kernels that run
 * beside it on another stream, e.g. the exchange chain of a partitioned d/dz started before it. */
int cfd_plan_set_xy_warps(cfd_plan *plan_x, int warps_per_sm);

/* The reference's stages one by one, for callers that drive the path the way dfdx does (compact.py:40-44); the
 * fused entry points above never call them.
 *   cfd_compute_rhs   rhs = Pade right-hand side of f (computeRHS, code/cuda/kernels.cu:4-47, without the ghosted
 *                     copy of gpuDA.global_to_local); out of place.
 *   cfd_plan_coeffs   [b1,c1,ai,bi,ci,an,bn] of the block (compact.py:159-166) for a matching nt_create:
 *                     solve_primary_system (compact.py:62-64) = nt_solve on rhs.
 *   cfd_sum_solutions x += alpha*x_UH + beta*x_LH over the whole block (sumSolutions, kernels.cu:49-74), alpha and
 *                     beta planes as produced by cfd_reduced_unknowns. */
int cfd_compute_rhs(cfd_plan *plan, const double *f, double *rhs, const double *halo_lo, const double *halo_hi,
                    void *stream);
int cfd_plan_coeffs(const cfd_plan *plan, double coeffs[7]);
int cfd_sum_solutions(cfd_plan *plan, double *x, const double *alpha, const double *beta, void *stream);

/* Interface right-hand side of the reduced system: faces[0] = -x_R[first], faces[1] = -x_R[last], zero
 * at physical ends.  faces is [2][plane].  Replaces negateAndCopyFaces (code/cuda/kernels.cu:76-113). */
int cfd_interface_pack(cfd_plan *plan, const double *df, double *faces, void *stream);

/* faces_all = the all-gathered [2*part_size][plane] interface planes of every rank along the line.
 * Solves the 2P-unknown reduced system per line (rows of code/cuda/compact.py:96-111; Thomas as in
 * reducedSolverKernel, kernels.cu:115-145) redundantly on this rank and adds
 * alpha * x_UH + beta * x_LH to df (sumSolutions, kernels.cu:49-74; secondary systems
 * compact.py:128-154 are precomputed at cfd_create).  Only planes whose correction exceeds fp64
 * round-off are touched. */
int cfd_reduced_correct(cfd_plan *plan, double *df, const double *faces_all, void *stream);

/* Fused multi-rank path (no correction pass).
 *   cfd_edge_faces       the same interface planes as cfd_apply + cfd_interface_pack, but directly from f and from
 *                        the first 32 and the last 32 rows of the block only (x_R[first] / x_R[last] do not depend on rows
 *                        further away, to 0.268^31 = 2e-18).
 *   cfd_reduced_unknowns solves the reduced system of every line for THIS rank's two unknowns (rows of
 *                        code/cuda/compact.py:96-111, elimination of reducedSolverKernel kernels.cu:115-145):
 *                        ab[0] = alpha plane (the left neighbour's last point), ab[1] = beta plane (the right
 *                        neighbour's first point).  neighbours_only = 0: faces = all-gathered [2P][plane];
 *                        neighbours_only = 1: faces = [2V][plane] over the V = cfd_nb_layout() "virtual ranks"
 *                        (rank-1 if any, rank, rank+1 if any) -- planes 2*own, 2*own+1 are this rank's faces,
 *                        plane 2*own-1 the left neighbour's faces[1], plane 2*own+2 the right neighbour's
 *                        faces[0], the two outermost planes unused (zero).  For blocks of >= 64 rows the reduced
 *                        matrix is block diagonal in fp64 (couplings ~0.268^64 = 1e-37), so this is exact and a
 *                        rank needs ONE plane from each neighbour instead of 2P.
 *                        flag0/flag1/seq: optional arrival flags to wait on first (peer-memory exchange below).
 *   cfd_apply_coupled    the fused kernel with alpha, beta folded into rows 0 and n-1 of the block system, so df
 *                        leaves it as the FINAL derivative: computeRHS + solve + sumSolutions
 *                        (code/cuda/compact.py:40-44) in one pass.  Needs >= 66 rows per block; thinner blocks
 *                        use the three-call path above. */
int cfd_edge_faces(cfd_plan *plan, const double *f, const double *halo_lo, const double *halo_hi, double *faces,
                   void *stream);
int cfd_reduced_unknowns(cfd_plan *plan, const double *faces, int neighbours_only, double *ab,
                         const unsigned long long *flag0, const unsigned long long *flag1, unsigned long long seq,
                         void *stream);
int cfd_apply_coupled(cfd_plan *plan, const double *f, double *df, const double *halo_lo, const double *halo_hi,
                      const double *ab, void *stream);
int cfd_nb_layout(const cfd_plan *plan, int *virtual_ranks, int *own_index);

/* NVLink peer-memory exchange (no NCCL on the data path).  The caller maps each neighbour's receive buffers and
 * 64-bit arrival flags into this process (e.g. torch.distributed._symmetric_memory, cudaIpc) and passes the
 * PEER addresses here; `seq` is a call counter that only grows, so flags never need resetting.
 *   cfd_push_planes     copies two planes of n doubles to peer buffers (halo exchange) and raises the peers' flags;
 *   cfd_edge_faces_p2p  = cfd_edge_faces that also stores faces[0] / faces[1] into the left / right neighbour's
 *                         interface buffer and raises their flags when the whole grid has finished;
 *   cfd_wait_flags      stream-ordered wait (one thread, bounded spin) until this rank's local flags reach seq.
 * Null peer pointers / flags mean "no neighbour on that side". */
int cfd_push_planes(const double *src0, double *dst0, const double *src1, double *dst1, long n,
                    unsigned long long *flag0, unsigned long long *flag1, unsigned long long seq, void *stream);
int cfd_edge_faces_p2p(cfd_plan *plan, const double *f, const double *halo_lo, const double *halo_hi, double *faces,
                       double *peer_lo, double *peer_hi, unsigned long long *flag_lo, unsigned long long *flag_hi,
                       unsigned long long seq, void *stream);
int cfd_wait_flags(const unsigned long long *flag0, const unsigned long long *flag1, unsigned long long seq,
                   void *stream);

/* The same exchange in ONE producer launch, with no wait before it: the interface faces depend linearly on the
 * neighbour points of f, so
 *   cfd_edge_faces_push            computes them with f[-1] := f[0], f[n] := f[n-1], stores them locally (faces[0], faces[1]) and into
 *                                  the neighbours' interface buffers, stores this block's first / last row into the
 *                                  neighbours' HALO buffers, and raises the neighbours' flags when all of it has landed;
 *   cfd_reduced_unknowns_deferred  waits for this rank's flags, adds the missing terms w * (halo - guess) (plan-time weights) to the
 *                                  four faces next to this block, and solves the neighbour-only reduced system.
 * Replaces halo exchange + negateAndCopyFaces + Gather / Scatter of code/cuda/compact.py:46-51,65-126 by two
 * launches and one point-to-point synchronisation; cfd_apply_coupled then reads the halo buffers as before. */
int cfd_edge_faces_push(cfd_plan *plan, const double *f, double *faces, double *peer_face_lo, double *peer_face_hi,
                        double *peer_halo_lo, double *peer_halo_hi, unsigned long long *flag_lo,
                        unsigned long long *flag_hi, unsigned long long seq, void *stream);
int cfd_reduced_unknowns_deferred(cfd_plan *plan, const double *faces_nb, const double *halo_lo, const double *halo_hi,
                                  const double *f, double *ab, const unsigned long long *flag0,
                                  const unsigned long long *flag1, unsigned long long seq, void *stream);

/* Synchronous host-buffer form of cfd_apply for part_size == 1 (what the reference's OpenCL flavour
 * offers: ndarray in, ndarray out, code/ocl/compact.py:26-61).  Copies f to the device, runs the kernel,
 * copies df back; staging buffers belong to the plan.  pinned != 0 promises page-locked host memory: the
 * promise is checked (CFD_EINVAL otherwise) and used -- x / y derivatives then move in z-slabs, copies in both
 * directions overlapping the kernels; pinned == 0 takes the plain copy - kernel - copy sequence. */
int cfd_apply_host(cfd_plan *plan, const double *f_host, double *df_host, int pinned);

/* plane = number of lines of the block (elements of one halo / interface plane). */
long cfd_plane_elems(const cfd_plan *plan);

/* Debug/inspection: copies the plan's host tables (test infrastructure reads them to emulate the
 * kernel's chunked recurrences on the CPU).  out must hold cfd_tables_size() doubles. */
int cfd_tables_size(void);
int cfd_plan_tables(const cfd_plan *plan, double *out);
/* Host-only (no device): tables for a line of n rows of the matrix coeffs = [b1,c1,ai,bi,ci,an,bn];
 * same layout as cfd_plan_tables except out[.. + 7] = 1 if the streaming fast path accepts the matrix. */
int cfd_debug_tables(int n, const double coeffs[7], double scale, double *out);
int cfd_debug_secondary(int n, int part_rank, int part_size, double *x_uh, double *x_lh, double *ra, double *rb,
                        double *rc);
int cfd_debug_neighbour(int n, int part_rank, int part_size, int *virtual_ranks, int *own_index, double *va,
                        double *vb, double *vc);
/* Host-only: the draw order of cfd_apply_xy for nz planes of nxp x-bundles + nyp y-bundles with `active` planes (or
 * sub x sub-tile squares when sub > 0) in flight; active = 0: plane by plane.  Entries are (item << 3) | segment with
 * item = plane * (nxp + nyp) + index in plane and segment 0 = whole line, s >= 1 = output chunks [(s-1) sub, s sub).
 * Returns the number of entries (<0 on error); out may be NULL to query it.  And the two weights d(lo face)/d f[-1],
 * d(hi face)/d f[n] that cfd_reduced_unknowns_deferred applies (blocks of n rows). */
long cfd_debug_xy_order(int nz, int nxp, int nyp, double active, int sub, int *out, long capacity);
/* Host-only: the launch shape cfd_apply_xy picks for an [nz, ny, nx] field on a device of `sms` SMs -- warps per SM
 * (6, or 7 for short launches), planes in flight of the draw order, and the line cut (0 = whole lines, else x-segment
 * length | y-segment length << 8 in tiles). */
int cfd_debug_xy_shape(int nz, int ny, int nx, int sms, int *warps, double *active, int *sub);
/* Measurement yardstick, no part of any derivative path: a plain grid-stride streaming kernel over n doubles (n even,
 * 16-byte aligned device arrays) with the read : write mix of the library's launches -- c == NULL: b = a (a single
 * derivative's 8 B read + 8 B written per point); otherwise b and c written from a (the fused d/dx + d/dy launch's 8 +
 * 16 B).  bench.py times it next to the derivative launches (`roofline.yardstick`). */
int cfd_debug_stream(const double *a, double *b, double *c, long n, void *stream);
int cfd_debug_halo_weights(int n, double h, double *w_lo, double *w_hi);
/* Host-only: chunks of look-ahead (1, 2; 0 = two-pass) the one-pass solve needs for n rows of coeffs[7], and the
 * definition of a scheme -- out[37]: b1,c1,ai,bi,ci,an,bn | two special rows per end | a2,b2,c2 (row 1), am,bm,cm
 * (row n-2) | c0,c1,c2,sgn (interior r_i = c0 f_i + c1 (f_{i+1} + sgn f_{i-1}) + c2 (f_{i+2} + sgn f_{i-2})) | closure
 * rows per end | q[2][4] (rows 0, 1: sum_k q f[k]) | p[2][4] (rows n-1, n-2: sum_k p f[n-1-k]) | look-ahead | coupling. */
int cfd_debug_lookahead(int n, const double coeffs[7]);
int cfd_debug_scheme(int scheme, int n, double h, double *out);
/* secondary solutions x_UH, x_LH (each n doubles) and the reduced matrix a,b,c (each 2*part_size). */
int cfd_plan_secondary(const cfd_plan *plan, double *x_uh, double *x_lh, double *ra, double *rb, double *rc);

/* ---------------------------------------------------------------------------------------------------
 * Batched near-Toeplitz solver.
 * Replaces NearToeplitzSolver(shape, coeffs) / .solve(x_d)
 * (code/cuda/solvers/templated/near_toeplitz.py:36-107; globalmem/near_toeplitz.py).
 * coeffs = [b1, c1, ai, bi, ci, an, bn]; solves, in place, every line of d[nz][ny][nx] along `axis`
 * (the reference solves along x only).  Any matrix with non-zero LU pivots is accepted: diagonally dominant
 * ones (every compact scheme) run the one-pass streaming kernel, the rest an exact two-pass LU.
 * ------------------------------------------------------------------------------------------------- */
int nt_create(nt_plan **plan, int nz, int ny, int nx, int axis, const double coeffs[7]);
int nt_solve(nt_plan *plan, double *d, void *stream);
/* 1 if the plan uses the exact two-pass LU (one launch per sweep, 32 B/unknown) because the matrix is not
 * diagonally dominant enough for a one-pass kernel (pivots not converged by row 32, or |g|^64 > 1.2e-16).
 * nt_lookahead: chunks of backward look-ahead of the one-pass solve, derived from the coefficients
 * (ceil(log(1.2e-16) / log|g| / 32): 1 for |g| <= 0.316, 2 up to 0.563, e.g. alpha = 1/3), 0 for the two-pass LU. */
int nt_is_exact_two_pass(const nt_plan *plan);
int nt_lookahead(const nt_plan *plan);
void nt_destroy(nt_plan *plan);

/* ---------------------------------------------------------------------------------------------------
 * Thread-parallel Thomas for `nsys` interleaved systems sharing one general tridiagonal matrix.
 * Replaces ReducedSolver / reducedSolverKernel (code/cuda/reduced.py:5-18, code/cuda/kernels.cu:115-145).
 * a, b, c: HOST arrays of length n (a[0], c[n-1] ignored); d: DEVICE array [n][nsys], solved in place; n <= 256.
 * cfd_pthomas_create eliminates the matrix once (the reference redoes it in every thread of every call,
 * kernels.cu:127-133) and keeps the pivots on the device; cfd_pthomas_solve is then one asynchronous launch that
 * allocates nothing.  cfd_pthomas = create + solve + stream synchronise + destroy, for one-off calls.
 * ------------------------------------------------------------------------------------------------- */
int cfd_pthomas_create(pt_plan **plan, const double *a, const double *b, const double *c, int n);
int cfd_pthomas_solve(pt_plan *plan, double *d, long nsys, void *stream);
void cfd_pthomas_destroy(pt_plan *plan);
int cfd_pthomas(const double *a, const double *b, const double *c, double *d, int n, long nsys, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * The partitioned d/dz with the host side in C: what a caller of the reference's multi-rank dfdx
 * (code/cuda/compact.py:29-44 -- halo exchange gpuDA.py:86-113, Gather / Scatter compact.py:93-94,121-122) binds
 * when it has no torch: one cfd_zpart per rank (= per process and GPU) owns that rank's receive buffer, maps its
 * two z-neighbours' buffers and issues the launches.  Set-up, once:
 *     cfd_zpart_create(&zp, plan_z)              plan_z: axis 2, part_size > 1, >= 66 planes per slab
 *     cfd_zpart_export(zp, handle)               CFD_IPC_HANDLE_BYTES bytes to hand to both neighbours (MPI / any channel)
 *     cfd_zpart_connect(zp, handle_lo, handle_hi)  the neighbours' handles (NULL at a physical end); cudaIpc underneath
 *   (ranks living in ONE process -- tests, one process driving several devices with peer access enabled -- pass
 *    cfd_zpart_buffer() of the neighbours to cfd_zpart_connect_ptr instead).
 * Per call:
 *     cfd_zpart_apply(zp, f, dfdz, stream)       the partitioned d/dz in ONE launch (stream_kernel_zx): per 32-line
 *                                                  bundle the two interface faces, their exchange with the z-neighbours
 *                                                  (self-validating 16-byte words stored straight into the neighbours'
 *                                                  receive arrays over NVLink: no flag, no fence, no collective), the
 *                                                  reduced system in registers and the coupled one-pass solve; no host
 *                                                  synchronisation.  DRAM traffic = the unpartitioned d/dz's;
 *     cfd_zpart_apply_xyz(zp, px, py, f, ...)    the whole gradient of the slab in two launches: cfd_apply_xy, then the
 *                                                  above;
 *     cfd_zpart_begin(zp, f, side_stream)        optional, the three-launch form: cfd_edge_faces_push and
 *                                                  cfd_reduced_unknowns_deferred early, on another stream; the next
 *                                                  cfd_zpart_apply on the same f then only runs cfd_apply_coupled
 *                                                  (events order the streams both ways).  CFD_NO_ZX=1 in the environment
 *                                                  makes every call take this form (cfd_zpart_apply_xyz then carries the
 *                                                  edge-face work as the first work items of the x/y kernel).
 *   cfd_zpart_set_ctas(zp, n): the one-kernel d/dz is persistent (one CTA per SM) and its warps wait for data that the
 *   NEIGHBOURS' kernels produce: with one rank per GPU that is always there; a process that keeps several ranks of a
 *   line on ONE device (tests) must cap the CTA count so that all of their kernels are resident together.
 * A rank that never arrives makes the neighbours' reduced kernel give up after cfd_set_wait_timeout_ms (default
 * 120 s; it polls with back-off and never traps): the next cfd_zpart_* call, or cfd_async_status(), returns
 * CFD_ETIMEOUT and the results of that call are invalid.
 * ------------------------------------------------------------------------------------------------- */
#define CFD_IPC_HANDLE_BYTES 64
int cfd_zpart_create(cfd_zpart **zp, cfd_plan *plan_z);
int cfd_zpart_export(cfd_zpart *zp, void *handle);
int cfd_zpart_connect(cfd_zpart *zp, const void *handle_lo, const void *handle_hi);
void *cfd_zpart_buffer(cfd_zpart *zp);
int cfd_zpart_connect_ptr(cfd_zpart *zp, void *buffer_lo, void *buffer_hi);
int cfd_zpart_set_ctas(cfd_zpart *zp, int max_ctas);   /* 0 = one CTA per SM (default); see below */
int cfd_zpart_begin(cfd_zpart *zp, const double *f, void *stream);
int cfd_zpart_apply(cfd_zpart *zp, const double *f, double *dfdz, void *stream);
int cfd_zpart_apply_xyz(cfd_zpart *zp, cfd_plan *plan_x, cfd_plan *plan_y, const double *f, double *dfdx, double *dfdy,
                        double *dfdz, void *stream);
/* The LANL distributed npts method (lanl-implementation/npts.c:275-655, python/npts.py:172-382) as an alternative to
 * the reduced-system method: cfd_create_npts builds this rank's plan from the LU of the WHOLE line (pivots handed from
 * rank to rank, precompute_beta_gam), cfd_zpart_apply_npts runs the two sweeps u = phi + u~ psi, x = phi' + x~ psi' as
 * ONE coupled pass once the incoming values u~ (forward-eliminated value at the left neighbour's last row) and x~
 * (solution at the right neighbour's first row) have crossed NVLink -- each from the 32 rows next to the interface
 * (psi decays by 0.268 per row).  Launches: halo push, wait, u~ kernel, wait, x~ kernel, wait, coupled kernel.  The
 * zpart is created from the npts plan with cfd_zpart_create and connected like any other. */
int cfd_create_npts(cfd_plan **plan, int nz, int ny, int nx, int axis, double h, int part_rank, int part_size);
int cfd_zpart_apply_npts(cfd_zpart *zp, const double *f, double *dfdz, void *stream);
void cfd_zpart_destroy(cfd_zpart *zp);
int cfd_set_wait_timeout_ms(long milliseconds);
int cfd_async_status(void);

/* Tuning knobs for experiments (0 = built-in default).  Not part of the reference surface. */
int cfd_set_launch(int warps_per_cta, int ctas_per_sm, int ring_slots);
/* Line segmentation of the streaming kernel: 32-row chunks per work item (0 = automatic: whole lines unless there
 * are fewer bundles than warps and lines of >= 512 rows).  Any value is exact to fp64 (32-row warm-ups). */
int cfd_set_segments(int chunks_per_segment);

/* Number of kernels this library has launched since load (bench.py reports it as gpu_launches). */
long cfd_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
