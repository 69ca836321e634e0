/*
 * oracle/cfd_oracle.c -- CPU restatement of the reference's compact-derivative hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (compact_finite_differences_b200/, include/)
 * may link, import or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, and there only as the checker / the CPU arm.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function below against
 *   - the reference's own npts.c compiled unmodified (oracle/_ref/libnpts_ref.so, oracle/Makefile),
 *   - the golden "Average absolute error" values of lanl-implementation/test_npts.c
 *     (0.0000293338 / 0.0000010363 / 0.0000000420 / 0.0000000019 for NX = 32/64/128/256),
 *   - scipy.linalg.solve_banded in the banded form every reference test uses
 *     (code/cuda/compact.py:189-203), and the committed fixtures under tests/golden/.
 *
 * Each function cites the reference file:line it restates.  Written from the algorithm, in our
 * own loop structure (line-at-a-time, any axis), not copied from the reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

/* ------------------------------------------------------------------------------------------
 * Minimal parallel-for over lines (this image has no libgomp): the line loop of each routine
 * is a range worker handed to T pthreads; T = oracle_set_num_threads (default 1).
 * ------------------------------------------------------------------------------------------ */
static int g_threads = 1;
typedef void (*range_fn)(long lo, long hi, void *ctx);
typedef struct { range_fn fn; void *ctx; long lo, hi; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->lo, j->hi, j->ctx); return 0; }

static void parallel_for(long n, range_fn fn, void *ctx)
{
    int T = g_threads;
    if (T > n) T = (int)(n > 0 ? n : 1);
    if (T <= 1) { fn(0, n, ctx); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * T);
    job_t *jb = (job_t *)malloc(sizeof(job_t) * T);
    for (int t = 0; t < T; t++) {
        jb[t].fn = fn; jb[t].ctx = ctx;
        jb[t].lo = n * t / T; jb[t].hi = n * (t + 1) / T;
        pthread_create(&th[t], 0, job_main, &jb[t]);
    }
    for (int t = 0; t < T; t++) pthread_join(th[t], 0);
    free(th);
    free(jb);
}

/* ------------------------------------------------------------------------------------------
 * Geometry helper: a field is f[nz][ny][nx] (C order, x fastest).  A "line" along `axis`
 * (0 = x, 1 = y, 2 = z; numbering of code/cuda/gpuDA.py:162) has n points `stride` apart;
 * line number l starts at line_base(l).
 * ------------------------------------------------------------------------------------------ */
typedef struct { long n, stride, nlines, inner, outer_stride; } geom_t;

static geom_t make_geom(int nz, int ny, int nx, int axis)
{
    geom_t g;
    if (axis == 0)      { g.n = nx; g.stride = 1;             g.inner = 1;             g.outer_stride = nx; }
    else if (axis == 1) { g.n = ny; g.stride = nx;            g.inner = nx;            g.outer_stride = (long)nx * ny; }
    else                { g.n = nz; g.stride = (long)nx * ny; g.inner = (long)nx * ny; g.outer_stride = 0; }
    g.nlines = (long)nz * ny * nx / g.n;
    return g;
}

static long line_base(const geom_t *g, long l)
{
    /* line l = (outer o, inner c): base = o*outer_stride + c   (for x: inner == 1 so c == 0) */
    long o = l / g->inner, c = l % g->inner;
    return o * g->outer_stride + c;
}

/* ------------------------------------------------------------------------------------------
 * Pade right-hand side.
 *   interior : (3/(4h)) * (f[i+1] - f[i-1])                  code/cuda/kernels.cu:34
 *   i = 0    : (1/(2h)) * (-5 f[0] + 4 f[1] + f[2])          code/cuda/kernels.cu:38   (only on the first rank)
 *   i = n-1  : -(1/(2h)) * (-5 f[n-1] + 4 f[n-2] + f[n-3])   code/cuda/kernels.cu:44   (only on the last rank)
 * (same three formulas: lanl-implementation/test_npts.c:90,94,95.)
 * On a rank that does not own a physical end, the neighbour's point comes from the one-plane
 * halo (ghost layer of code/cuda/gpuDA.py:61-132): halo_lo / halo_hi are arrays of `nlines`
 * values, one per line, or NULL when the closure applies.
 * ------------------------------------------------------------------------------------------ */
typedef struct { geom_t g; const double *f; double *rhs; double h; const double *halo_lo, *halo_hi; } rhs_ctx;

static void rhs_range(long lo, long hi, void *vp)
{
    const rhs_ctx *c = (const rhs_ctx *)vp;
    const geom_t g = c->g;
    const long n = g.n, s = g.stride;
    const double h = c->h, *halo_lo = c->halo_lo, *halo_hi = c->halo_hi;
    for (long l = lo; l < hi; l++) {
        const double *fl = c->f + line_base(&g, l);
        double *rl = c->rhs + line_base(&g, l);
        for (long i = 1; i < n - 1; i++)
            rl[i * s] = (3. / (4 * h)) * (fl[(i + 1) * s] - fl[(i - 1) * s]);
        if (halo_lo) rl[0] = (3. / (4 * h)) * (fl[s] - halo_lo[l]);
        else         rl[0] = (1. / (2 * h)) * (-5 * fl[0] + 4 * fl[s] + fl[2 * s]);
        if (halo_hi) rl[(n - 1) * s] = (3. / (4 * h)) * (halo_hi[l] - fl[(n - 2) * s]);
        else         rl[(n - 1) * s] = -(1. / (2 * h)) * (-5 * fl[(n - 1) * s] + 4 * fl[(n - 2) * s] + fl[(n - 3) * s]);
    }
}

void oracle_rhs(const double *f, double *rhs, int nz, int ny, int nx, int axis, double h,
                const double *halo_lo, const double *halo_hi)
{
    rhs_ctx c = { make_geom(nz, ny, nx, axis), f, rhs, h, halo_lo, halo_hi };
    parallel_for(c.g.nlines, rhs_range, &c);
}

/* ------------------------------------------------------------------------------------------
 * npts pivots at one rank (lanl-implementation/npts.c:580-655 with npx = 1):
 *   beta_0 = 1, gam_0 = 0;  gam_i = beta_{i-1} * c_{i-1},  beta_i = 1 / (1 - a_i * gam_i)
 * for the matrix  [1 2; 1/4 1 1/4; ... ; 2 1]  (c_0 = 2, a_{n-1} = 2, otherwise 1/4).
 * ------------------------------------------------------------------------------------------ */
void oracle_npts_beta_gam(int n, double *beta, double *gam)
{
    beta[0] = 1.0;
    gam[0] = 0.0;
    for (int i = 1; i < n; i++) {
        double c_prev = (i == 1) ? 2.0 : 0.25;
        double a_i = (i == n - 1) ? 2.0 : 0.25;
        gam[i] = beta[i - 1] * c_prev;
        beta[i] = 1. / (1. - a_i * gam[i]);
    }
}

/* ------------------------------------------------------------------------------------------
 * npts solve at one rank, any axis, out of place (r -> u).
 *   L->R : phi_k = beta_k (r_k - a_k phi_{k-1})      lanl-implementation/npts.c:329-353 (+431-439: u = phi + u0*psi,
 *                                                     which at npx = 1 is exactly this recurrence started at beta_0 r_0)
 *   R->L : x_k = u_k - gam_{k+1} x_{k+1}             lanl-implementation/npts.c:456-478 (+554-562)
 * ------------------------------------------------------------------------------------------ */
typedef struct { geom_t g; const double *r; double *u; const double *beta, *gam; } npts_ctx;

static void npts_range(long lo, long hi, void *vp)
{
    const npts_ctx *c = (const npts_ctx *)vp;
    const geom_t g = c->g;
    const long n = g.n, s = g.stride;
    const double *beta = c->beta, *gam = c->gam;
    for (long l = lo; l < hi; l++) {
        const double *rl = c->r + line_base(&g, l);
        double *ul = c->u + line_base(&g, l);
        ul[0] = beta[0] * rl[0];
        for (long k = 1; k < n - 1; k++)
            ul[k * s] = beta[k] * (rl[k * s] - (1. / 4) * ul[(k - 1) * s]);
        ul[(n - 1) * s] = beta[n - 1] * (rl[(n - 1) * s] - 2 * ul[(n - 2) * s]);
        for (long k = n - 2; k >= 0; k--)
            ul[k * s] = ul[k * s] - gam[k + 1] * ul[(k + 1) * s];
    }
}

void oracle_npts_solve(const double *r, double *u, int nz, int ny, int nx, int axis,
                       const double *beta, const double *gam)
{
    npts_ctx c = { make_geom(nz, ny, nx, axis), r, u, beta, gam };
    parallel_for(c.g.nlines, npts_range, &c);
}

/* ------------------------------------------------------------------------------------------
 * General near-Toeplitz tridiagonal solve, in place, any axis.
 * Matrix semantics of NearToeplitzSolver (code/cuda/solvers/templated/near_toeplitz.py:36-50):
 *   a = (_, ai, ..., ai, an)   b = (b1, bi, ..., bi, bn)   c = (c1, ci, ..., ci, _)
 * coeffs = [b1, c1, ai, bi, ci, an, bn].  Plain LU without pivoting (Thomas), which is what both
 * the reference's cyclic reduction and npts amount to; reference tests compare against LAPACK's
 * banded solve (code/ocl/test/test_near_toeplitz.py:31-48).
 * ------------------------------------------------------------------------------------------ */
typedef struct { geom_t g; double *d; const double *cp, *pv; double ai, an; } nt_ctx;

static void nt_range(long lo, long hi, void *vp)
{
    const nt_ctx *c = (const nt_ctx *)vp;
    const geom_t g = c->g;
    const long n = g.n, s = g.stride;
    const double *cp = c->cp, *pv = c->pv, ai = c->ai, an = c->an;
    for (long l = lo; l < hi; l++) {
        double *x = c->d + line_base(&g, l);
        x[0] = x[0] / pv[0];
        for (long i = 1; i < n; i++) {
            double a_i = (i == n - 1) ? an : ai;
            x[i * s] = (x[i * s] - a_i * x[(i - 1) * s]) / pv[i];
        }
        for (long i = n - 2; i >= 0; i--)
            x[i * s] -= cp[i] * x[(i + 1) * s];
    }
}

void oracle_near_toeplitz_solve(double *d, int nz, int ny, int nx, int axis, const double *coeffs)
{
    geom_t g = make_geom(nz, ny, nx, axis);
    const long n = g.n;
    const double b1 = coeffs[0], c1 = coeffs[1], ai = coeffs[2], bi = coeffs[3], ci = coeffs[4],
                 an = coeffs[5], bn = coeffs[6];
    double *cp = (double *)malloc(sizeof(double) * n);   /* c'_i = c_i / pivot_i  */
    double *pv = (double *)malloc(sizeof(double) * n);   /* pivot_i               */
    pv[0] = b1;
    cp[0] = c1 / b1;
    for (long i = 1; i < n; i++) {
        double a_i = (i == n - 1) ? an : ai, b_i = (i == n - 1) ? bn : bi;
        pv[i] = b_i - a_i * cp[i - 1];
        cp[i] = ci / pv[i];
    }
    nt_ctx c = { g, d, cp, pv, ai, an };
    parallel_for(g.nlines, nt_range, &c);
    free(cp);
    free(pv);
}

/* ------------------------------------------------------------------------------------------
 * The reference GPU solver's own algorithm: cyclic reduction with precomputed per-level scalars.
 *   coefficient recurrences : code/cuda/solvers/templated/near_toeplitz.py:109-184
 *   forward reduction       : code/cuda/solvers/globalmem/kernels.cu:45-63
 *   2x2 solve               : code/cuda/solvers/globalmem/kernels.cu:29-44
 *   back substitution       : code/cuda/solvers/globalmem/kernels.cu:92-115
 * n must be a power of two >= 4 (reference assert, templated/near_toeplitz.py:56).  Contiguous
 * lines only (the reference solver is x-only).  Tables are L = log2(n) long.
 * ------------------------------------------------------------------------------------------ */
typedef struct { int L; double *a, *b, *c, *k1, *k2, *b_first, *k1_first, *k1_last; } cr_tab_t;

static void cr_precompute(int n, const double *coeffs, cr_tab_t *t)
{
    int L = 0;
    while ((1 << L) < n) L++;
    t->L = L;
    double *blk = (double *)calloc((size_t)8 * L, sizeof(double));
    t->a = blk; t->b = blk + L; t->c = blk + 2 * L; t->k1 = blk + 3 * L; t->k2 = blk + 4 * L;
    t->b_first = blk + 5 * L; t->k1_first = blk + 6 * L; t->k1_last = blk + 7 * L;
    const double b1 = coeffs[0], c1 = coeffs[1], ai = coeffs[2], bi = coeffs[3], ci = coeffs[4],
                 an = coeffs[5], bn = coeffs[6];
    double a_last = 0, b_last = 0;
    for (int i = 0; i < L - 1; i++) {
        /* level-i matrix is Toeplitz (pa, pb, pc) except its first diagonal entry pbf and last row */
        double pa = i ? t->a[i - 1] : ai, pb = i ? t->b[i - 1] : bi, pc = i ? t->c[i - 1] : ci;
        double pbf = i ? t->b_first[i - 1] : b1, pcf = i ? pc : c1;
        double pal = i ? a_last : an, pbl = i ? b_last : bn;
        t->k1[i] = pa / pb;
        t->k2[i] = pc / pb;
        t->a[i] = -pa * t->k1[i];
        t->b[i] = pb - pc * t->k1[i] - pa * t->k2[i];
        t->c[i] = -pc * t->k2[i];
        t->k1_first[i] = pa / pbf;
        t->b_first[i] = pb - pcf * t->k1_first[i] - pa * t->k2[i];
        t->k1_last[i] = pal / pb;
        a_last = -pa * t->k1_last[i];
        b_last = pbl - pc * t->k1_last[i];
    }
    t->a[L - 1] = a_last;
    t->b[L - 1] = b_last;
}

typedef struct { double *d; int n; const double *coeffs; const cr_tab_t *t; } cr_ctx;

static void cr_range(long lo, long hi, void *vp)
{
    const cr_ctx *cc = (const cr_ctx *)vp;
    const cr_tab_t t = *cc->t;
    const int L = t.L, n = cc->n;
    const double *coeffs = cc->coeffs;
    const double b1 = coeffs[0], c1 = coeffs[1], ai = coeffs[2], bi = coeffs[3], ci = coeffs[4];
    for (long l = lo; l < hi; l++) {
        double *x = cc->d + l * (long)n;
        /* forward reduction, levels 0 .. L-2 */
        for (int lev = 0; lev < L - 1; lev++) {
            int stride = 2 << lev, half = stride >> 1;
            for (int i = stride - 1; i < n; i += stride) {
                if (i == n - 1)          x[i] -= x[i - half] * t.k1_last[lev];
                else if (i == stride - 1) x[i] -= x[i - half] * t.k1_first[lev] + x[i + half] * t.k2[lev];
                else                      x[i] -= x[i - half] * t.k1[lev] + x[i + half] * t.k2[lev];
            }
        }
        /* 2x2 solve on (n/2-1, n-1) */
        {
            int m = n / 2 - 1, e = n - 1;
            double m00 = t.b_first[L - 2], m01 = t.c[L - 2], m10 = t.a[L - 1], m11 = t.b[L - 1];
            double det = m00 * m11 - m01 * m10;
            double xm = (x[m] * m11 - m01 * x[e]) / det;
            double xe = (m00 * x[e] - x[m] * m10) / det;
            x[m] = xm; x[e] = xe;
        }
        /* back substitution */
        for (int stride = n / 2; stride >= 4; stride >>= 1) {
            int half = stride >> 1, idx = 0;
            while ((4 << idx) < stride) idx++;           /* idx = log2(stride) - 2 */
            for (int i = half - 1; i < n; i += stride) {
                if (i == half - 1) x[i] = (x[i] - t.c[idx] * x[i + half]) / t.b_first[idx];
                else               x[i] = (x[i] - t.a[idx] * x[i - half] - t.c[idx] * x[i + half]) / t.b[idx];
            }
        }
        x[0] = (x[0] - c1 * x[1]) / b1;
        for (int i = 2; i < n; i += 2)
            x[i] = (x[i] - ai * x[i - 1] - ci * x[i + 1]) / bi;
    }
}

void oracle_cr_solve(double *d, long nlines, int n, const double *coeffs)
{
    cr_tab_t t;
    cr_precompute(n, coeffs, &t);
    cr_ctx c = { d, n, coeffs, &t };
    parallel_for(nlines, cr_range, &c);
    free(t.a);
}

/* ------------------------------------------------------------------------------------------
 * Thread-parallel Thomas ("pThomas", code/cuda/kernels.cu:115-145): one general tridiagonal
 * matrix (a, b, c of length n) shared by `nsys` systems whose elements are `nsys` apart.
 * ------------------------------------------------------------------------------------------ */
typedef struct { const double *a, *c2, *piv; double *d; int n; long nsys; } pt_ctx;

static void pt_range(long lo, long hi, void *vp)
{
    const pt_ctx *c = (const pt_ctx *)vp;
    const double *a = c->a, *c2 = c->c2, *piv = c->piv;
    double *d = c->d;
    const int n = c->n;
    const long nsys = c->nsys;
    for (long s = lo; s < hi; s++) {
        d[s] = d[s] / piv[0];
        for (int i = 1; i < n; i++) d[s + i * nsys] = (d[s + i * nsys] - a[i] * d[s + (i - 1) * nsys]) / piv[i];
        for (int i = n - 2; i >= 0; i--) d[s + i * nsys] -= c2[i] * d[s + (i + 1) * nsys];
    }
}

void oracle_pthomas(const double *a, const double *b, const double *c, double *d, int n, long nsys)
{
    double *c2 = (double *)malloc(sizeof(double) * n);
    double *piv = (double *)malloc(sizeof(double) * n);
    piv[0] = b[0];
    c2[0] = c[0] / b[0];
    for (int i = 1; i < n; i++) { piv[i] = b[i] - a[i] * c2[i - 1]; c2[i] = c[i] / piv[i]; }
    pt_ctx pc = { a, c2, piv, d, n, nsys };
    parallel_for(nsys, pt_range, &pc);
    free(c2);
    free(piv);
}

/* ------------------------------------------------------------------------------------------
 * Whole derivative on one rank = RHS + npts solve (the reference's CPU path of BASELINE.json
 * configs[0]: lanl-implementation/test_npts.c:86-97 then :126).  `df` doubles as the RHS buffer.
 * ------------------------------------------------------------------------------------------ */
void oracle_derivative(const double *f, double *df, int nz, int ny, int nx, int axis, double h)
{
    geom_t g = make_geom(nz, ny, nx, axis);
    double *beta = (double *)malloc(sizeof(double) * g.n), *gam = (double *)malloc(sizeof(double) * g.n);
    oracle_npts_beta_gam((int)g.n, beta, gam);
    oracle_rhs(f, df, nz, ny, nx, axis, h, 0, 0);
    oracle_npts_solve(df, df, nz, ny, nx, axis, beta, gam);
    free(beta);
    free(gam);
}

int oracle_num_threads(void) { return g_threads; }

void oracle_set_num_threads(int n) { g_threads = n > 0 ? n : 1; }
