// CPU harness for eig_leading.h: Householder tridiagonalisation of a symmetric matrix (the device has its
// own warp-level tred2), the prototype's leading eigenpairs, back-transformation.  Built and driven by
// eig_leading_check.py.
#include <cmath>
#include <vector>

#include "eig_leading.h"

extern "C" int leading_eig(const double* G, int n, double eps, int kmax, double* vals, double* U /* n x r, column major */)
{
    double tr = 0;
    for (int i = 0; i < n; i++) tr += G[i * n + i];
    std::vector<double> A(G, G + n * n), d(n), e(n, 0.0);
    if (tr > 0)
        for (auto& a : A) a /= tr;
    std::vector<std::vector<double>> vs;
    std::vector<double> betas;
    for (int k = 0; k + 2 < n; k++) {
        const int m = n - k - 1;
        std::vector<double> v(m);
        double nx = 0;
        for (int i = 0; i < m; i++) {
            v[i] = A[(k + 1 + i) * n + k];
            nx += v[i] * v[i];
        }
        nx = std::sqrt(nx);
        const double alpha = v[0] > 0 ? -nx : nx;
        v[0] -= alpha;
        double vv = 0;
        for (double t : v) vv += t * t;
        const double beta = vv > 0 ? 2.0 / vv : 0.0;
        vs.push_back(v);
        betas.push_back(beta);
        if (beta == 0.0) continue;
        std::vector<double> p(m, 0.0), w(m);
        for (int i = 0; i < m; i++)
            for (int j = 0; j < m; j++) p[i] += A[(k + 1 + i) * n + (k + 1 + j)] * v[j];
        double vp = 0;
        for (int i = 0; i < m; i++) {
            p[i] *= beta;
            vp += v[i] * p[i];
        }
        for (int i = 0; i < m; i++) w[i] = p[i] - 0.5 * beta * vp * v[i];
        for (int i = 0; i < m; i++)
            for (int j = 0; j < m; j++) A[(k + 1 + i) * n + (k + 1 + j)] -= v[i] * w[j] + w[i] * v[j];
        A[(k + 1) * n + k] = A[k * n + k + 1] = alpha;
        for (int i = 1; i < m; i++) A[(k + 1 + i) * n + k] = A[k * n + k + 1 + i] = 0.0;
    }
    for (int i = 0; i < n; i++) {
        d[i] = A[i * n + i];
        if (i) e[i] = A[i * n + i - 1];
    }
    std::vector<double> Y((size_t)kmax * n);
    const int r = vt_leading_eigenpairs(d.data(), e.data(), n, eps * eps / 3.0, kmax, vals, Y.data());
    for (int j = 0; j < r; j++) {
        vals[j] *= tr > 0 ? tr : 1.0;
        double* y = Y.data() + (size_t)j * n;
        for (int k = (int)vs.size() - 1; k >= 0; k--) {
            const int m = n - k - 1;
            double dot = 0;
            for (int i = 0; i < m; i++) dot += vs[k][i] * y[k + 1 + i];
            dot *= betas[k];
            for (int i = 0; i < m; i++) y[k + 1 + i] -= dot * vs[k][i];
        }
        for (int i = 0; i < n; i++) U[(size_t)j * n + i] = y[i];
    }
    return r;
}
