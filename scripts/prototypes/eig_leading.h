// Prototype of the next Tucker eigen-solve, written the way the device code will be (plain loops over
// arrays, one "lane" per eigenvalue index, no library calls) so that it can be moved into tucker.cu with
// VT_HD = __device__ once a GPU is available to validate it.  NOT part of the product yet: nothing under
// vlasovtucker_b200/ includes this file.  Checked on the CPU by eig_leading_check.py against
// numpy.linalg.eigh (same ranks, same projector) including zero, rank-1 and fully degenerate matrices.
//
// Input: a symmetric tridiagonal matrix (d[0..n), e[1..n) sub-diagonal, e[0] unused) with unit trace,
// as the warp-level tred2 of tucker.cu leaves it.  Output: the eigenvalues above `thr` (at most kmax,
// at least one), largest first, and their eigenvectors in the tridiagonal basis.
#pragma once
#include <math.h>

#ifndef VT_HD
#define VT_HD
#endif

#define VT_EIG_NMAX 64

// Number of eigenvalues smaller than x: sign changes of the leading principal minors
// p_i = (d_i - x) p_{i-1} - e_i^2 p_{i-2}, kept in range by powers of two.  No division.
VT_HD inline int vt_sturm_count(const double* d, const double* e2, int n, double x)
{
    int cnt = 0;
    double pm2 = 0.0, pm1 = 1.0;
    for (int i = 0; i < n; i++) {
        double p = (d[i] - x) * pm1 - (i ? e2[i] * pm2 : 0.0);
        if (p == 0.0) p = pm1 > 0 ? -1e-300 : 1e-300;
        cnt += (p < 0) != (pm1 < 0);
        const double m = fmax(fabs(p), fabs(pm1));
        if (m < 1e-100 || m > 1e100) {
            int ex;
            frexp(m, &ex);
            p = ldexp(p, -ex);
            pm1 = ldexp(pm1, -ex);
        }
        pm2 = pm1;
        pm1 = p;
    }
    return cnt;
}

// k-th largest eigenvalue (k = 0 is the largest) inside [lo, hi] by bisection to the last bit.
VT_HD inline double vt_bisect_kth_largest(const double* d, const double* e2, int n, int k, double lo, double hi)
{
    for (int it = 0; it < 1100; it++) {
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;
        if (vt_sturm_count(d, e2, n, mid) <= n - 1 - k) lo = mid;
        else hi = mid;
    }
    return 0.5 * (lo + hi);
}

// One inverse-iteration solve (T - lam I) y = x in place, Gaussian elimination with partial pivoting on
// the tridiagonal (two super-diagonals of fill).  work: 4n doubles + n ints worth of room (5n doubles).
VT_HD inline void vt_tridiag_solve_shifted(const double* d, const double* e, int n, double lam, double tiny, double* x, double* work)
{
    double* a = work;           // pivots
    double* b = work + n;       // first super-diagonal of U
    double* c = work + 2 * n;   // multipliers
    double* f = work + 3 * n;   // second super-diagonal of U
    double* sw = work + 4 * n;  // 1.0 where rows k, k+1 were exchanged
    for (int i = 0; i < n; i++) {
        a[i] = d[i] - lam;
        b[i] = i + 1 < n ? e[i + 1] : 0.0;
        f[i] = 0.0;
    }
    for (int k = 0; k + 1 < n; k++) {
        const double sub = e[k + 1];
        if (fabs(a[k]) >= fabs(sub)) {
            if (fabs(a[k]) < tiny) a[k] = a[k] < 0 ? -tiny : tiny;
            c[k] = sub / a[k];
            a[k + 1] -= c[k] * b[k];
            sw[k] = 0.0;
        } else {
            const double mult = a[k] / sub;
            a[k] = sub;
            const double t = a[k + 1];
            a[k + 1] = b[k] - mult * t;
            if (k + 2 < n) {
                f[k] = b[k + 1];
                b[k + 1] = -mult * f[k];
            }
            b[k] = t;
            c[k] = mult;
            sw[k] = 1.0;
        }
    }
    if (fabs(a[n - 1]) < tiny) a[n - 1] = a[n - 1] < 0 ? -tiny : tiny;
    for (int k = 0; k + 1 < n; k++) {
        if (sw[k] == 0.0) x[k + 1] -= c[k] * x[k];
        else {
            const double t = x[k];
            x[k] = x[k + 1];
            x[k + 1] = t - c[k] * x[k];
        }
    }
    x[n - 1] /= a[n - 1];
    if (n > 1) x[n - 2] = (x[n - 2] - b[n - 2] * x[n - 1]) / a[n - 2];
    for (int k = n - 3; k >= 0; k--) x[k] = (x[k] - b[k] * x[k + 1] - f[k] * x[k + 2]) / a[k];
}

// Rank rule + leading eigenpairs of the unit-trace tridiagonal matrix.
//   thr   eigenvalue threshold (eps^2 / 3 for the relative rule sigma_j > eps |sigma| / sqrt 3)
//   vals  [kmax]      out, largest first
//   vecs  [kmax * n]  out, row j = eigenvector j in the tridiagonal basis, orthonormal
// Returns the rank r (1 <= r <= kmax).  On the device: loop "lane" j runs on lane j; the Gram-Schmidt
// pass is the only cross-lane step.
VT_HD inline int vt_leading_eigenpairs(const double* d, const double* e, int n, double thr, int kmax, double* vals, double* vecs)
{
    double e2[VT_EIG_NMAX], work[5 * VT_EIG_NMAX];
    double lo = d[0], hi = d[0], norm = 0.0;
    for (int i = 0; i < n; i++) {
        e2[i] = i ? e[i] * e[i] : 0.0;
        const double rad = (i ? fabs(e[i]) : 0.0) + (i + 1 < n ? fabs(e[i + 1]) : 0.0);
        lo = fmin(lo, d[i] - rad);
        hi = fmax(hi, d[i] + rad);
        norm = fmax(norm, fabs(d[i]) + rad);
    }
    int r = n - vt_sturm_count(d, e2, n, thr);
    r = r < 1 ? 1 : (r > kmax ? kmax : r);
    const double tiny = 2.3e-16 * (norm > 0 ? norm : 1.0);
    for (int j = 0; j < r; j++) vals[j] = vt_bisect_kth_largest(d, e2, n, j, lo, hi);
    for (int j = 0; j < r; j++) {
        double* x = vecs + (long)j * n;
        // lane-dependent start so that equal eigenvalues still span different directions
        unsigned s = 0x9E3779B9u * (unsigned)(j + 1);
        for (int i = 0; i < n; i++) {
            s = s * 1664525u + 1013904223u;
            x[i] = 0.5 + (double)(s >> 8) * (1.0 / 16777216.0);
        }
        for (int it = 0; it < 4; it++) {
            vt_tridiag_solve_shifted(d, e, n, vals[j], tiny, x, work);
            double big = 0.0;
            for (int i = 0; i < n; i++) big = fmax(big, fabs(x[i]));
            if (!(big > 0.0) || !isfinite(big)) {               // singular beyond repair: restart on a unit vector
                for (int i = 0; i < n; i++) x[i] = i == (j % n) ? 1.0 : 0.0;
                big = 1.0;
            }
            for (int i = 0; i < n; i++) x[i] /= big;
            for (int pass = 0; pass < 2; pass++)
                for (int q = 0; q < j; q++) {
                    const double* y = vecs + (long)q * n;
                    double dot = 0.0;
                    for (int i = 0; i < n; i++) dot += x[i] * y[i];
                    for (int i = 0; i < n; i++) x[i] -= dot * y[i];
                }
            double nn = 0.0;
            for (int i = 0; i < n; i++) nn += x[i] * x[i];
            if (nn <= 1e-300) {                                  // fell into the span of the earlier ones
                for (int i = 0; i < n; i++) x[i] = i == ((j + it + 1) % n) ? 1.0 : 0.0;
                continue;
            }
            nn = 1.0 / sqrt(nn);
            for (int i = 0; i < n; i++) x[i] *= nn;
        }
    }
    return r;
}
