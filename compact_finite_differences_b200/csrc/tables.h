// tables.h -- host-side coefficient tables for the streaming near-Toeplitz kernels.
//
// A line of the near-Toeplitz system  [b1 c1; ai bi ci; ...; an bn]
// (semantics of code/cuda/solvers/templated/near_toeplitz.py:36-50) is solved by LU with
// precomputed pivots, the same elimination the reference CPU path uses
// (lanl-implementation/npts.c:580-655: beta_i = 1/(b_i - a_i*gam_i), gam_i = beta_{i-1}*c_{i-1}):
//
//   forward : e_i = s_i * r_i - l_i * e_{i-1}      s_i = beta_i, l_i = a_i * beta_i
//   backward: x_i = e_i - g_i * x_{i+1}            g_i = c_i * beta_i  (= gam_{i+1} of npts)
//
// The kernels walk a line in chunks of CH = 32 rows.  Away from the two ends the pivots have
// converged to a constant (for the Pade matrix |beta_i - beta_inf| < 1e-37 by row 32), so only
// three coefficient sets are needed: a 32-row HEAD table (rows 0..31), a 32-row TAIL table (the
// last chunk, rows 32(K-1)..), and MID constants.  Rows >= n in the tail table are zero, which
// makes padded rows inert (e = x = 0) and the backward sweep exact from the true end.
#pragma once
#include <cmath>
#include <vector>

namespace cfd {

constexpr int CH = 32;   // rows per chunk == lanes per warp

struct RowTab {          // 768 B
    double sk[CH];       // multiplies the RHS stencil difference (derivative) or the RHS itself (solve)
    double l[CH];
    double g[CH];
};

struct LineCoeffs { double b1, c1, ai, bi, ci, an, bn; };

// Near-Toeplitz matrix whose first TWO and last TWO rows may differ from the interior (higher-order compact schemes
// drop to a lower-order scheme next to the closure rows):
//   row 0: (b1, c1)   row 1: (a2, b2, c2)   interior: (ai, bi, ci)   row n-2: (am, bm, cm)   row n-1: (an, bn)
struct LineMatrix {
    double b1, c1, ai, bi, ci, an, bn;
    bool two;
    double a2, b2, c2, am, bm, cm;
    double a(int i, int n) const { return i == 0 ? 0.0 : i == n - 1 ? an : (two && i == 1) ? a2 : (two && i == n - 2) ? am : ai; }
    double b(int i, int n) const { return i == 0 ? b1 : i == n - 1 ? bn : (two && i == 1) ? b2 : (two && i == n - 2) ? bm : bi; }
    double c(int i, int n) const { return i == 0 ? c1 : i == n - 1 ? 0.0 : (two && i == 1) ? c2 : (two && i == n - 2) ? cm : ci; }
};

inline LineMatrix as_matrix(const LineCoeffs &m)
{
    LineMatrix g = {m.b1, m.c1, m.ai, m.bi, m.ci, m.an, m.bn, false, 0, 0, 0, 0, 0, 0};
    return g;
}

struct Pivots {
    std::vector<double> beta, l, g;
    double decay = 0;        // |g| in the converged region: backward coupling per row
    bool converged = false;  // pivots constant (to ~1 ulp) over rows [CH, n-1)
    bool finite = true;
};

inline Pivots build_pivots(int n, const LineCoeffs &m)
{
    Pivots p;
    p.beta.resize(n); p.l.resize(n); p.g.resize(n);
    p.beta[0] = 1.0 / m.b1;
    p.l[0] = 0.0;
    for (int i = 1; i < n; i++) {
        const double a_i = (i == n - 1) ? m.an : m.ai;
        const double b_i = (i == n - 1) ? m.bn : m.bi;
        const double c_prev = (i == 1) ? m.c1 : m.ci;
        const double gam = p.beta[i - 1] * c_prev;          // npts.c:629-635
        p.beta[i] = 1.0 / (b_i - a_i * gam);                // npts.c:637-645
        p.l[i] = a_i * p.beta[i];
        p.g[i - 1] = gam;
    }
    p.g[n - 1] = 0.0;
    for (int i = 0; i < n; i++)
        if (!std::isfinite(p.beta[i])) p.finite = false;
    // convergence of the interior pivots and decay of the backward coupling
    p.converged = true;
    if (n > 2 * CH) {
        const double ref = p.beta[CH];
        for (int i = CH; i < n - 1; i++)
            if (std::fabs(p.beta[i] - ref) > 4e-16 * std::fabs(ref)) { p.converged = false; break; }
        p.decay = std::fabs(p.g[CH]);
    }
    return p;
}

// The same elimination for a LineMatrix.  `decay` / `converged` look at rows [CH, n-2): the rows the kernels serve with
// constants (rows n-2 and n-1 always come from a per-row table).
inline Pivots build_pivots(int n, const LineMatrix &m)
{
    Pivots p;
    p.beta.resize(n); p.l.resize(n); p.g.resize(n);
    p.beta[0] = 1.0 / m.b(0, n);
    p.l[0] = 0.0;
    for (int i = 1; i < n; i++) {
        const double gam = p.beta[i - 1] * m.c(i - 1, n);
        p.beta[i] = 1.0 / (m.b(i, n) - m.a(i, n) * gam);
        p.l[i] = m.a(i, n) * p.beta[i];
        p.g[i - 1] = gam;
    }
    p.g[n - 1] = 0.0;
    for (int i = 0; i < n; i++)
        if (!std::isfinite(p.beta[i])) p.finite = false;
    p.converged = true;
    if (n > 2 * CH) {
        const double ref = p.beta[CH];
        for (int i = CH; i < n - 2; i++)
            if (std::fabs(p.beta[i] - ref) > 4e-16 * std::fabs(ref)) { p.converged = false; break; }
        p.decay = std::fabs(p.g[CH]);
        if (std::fabs(p.l[CH]) > p.decay) p.decay = std::fabs(p.l[CH]);
    }
    return p;
}

// Table of chunk `chunk` (rows CH*chunk .. CH*chunk+31); `scale` folds the stencil factor 3/(4h).
inline RowTab chunk_table(const Pivots &p, int n, int chunk, double scale)
{
    RowTab t;
    for (int j = 0; j < CH; j++) {
        const int i = chunk * CH + j;
        if (i < n) { t.sk[j] = p.beta[i] * scale; t.l[j] = p.l[i]; t.g[j] = p.g[i]; }
        else       { t.sk[j] = 0.0; t.l[j] = 0.0; t.g[j] = 0.0; }
    }
    return t;
}

// Plain Thomas on one line with explicit diagonals (host; used for the secondary systems).
inline void thomas_host(int n, const std::vector<double> &a, const std::vector<double> &b,
                        const std::vector<double> &c, std::vector<double> &x)
{
    std::vector<double> cp(n);
    double piv = b[0];
    cp[0] = c[0] / piv;
    x[0] = x[0] / piv;
    for (int i = 1; i < n; i++) {
        piv = b[i] - a[i] * cp[i - 1];
        cp[i] = c[i] / piv;
        x[i] = (x[i] - a[i] * x[i - 1]) / piv;
    }
    for (int i = n - 2; i >= 0; i--) x[i] -= cp[i] * x[i + 1];
}

}  // namespace cfd
