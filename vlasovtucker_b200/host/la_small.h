// Small dense linear algebra for the host-side Tucker class: Householder QR and a one-sided
// Jacobi (Hestenes) SVD.  Both are backward stable, which matters because the reference's
// truncation rule keeps every singular value above eps*|sigma|/sqrt(3) with eps as small as
// 1e-10 (particle_data.h:49) — a Gram-matrix eigen-decomposition would lose those.
#pragma once
#include <Eigen/Dense>
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

namespace VlasovTucker {
namespace la {

using Eigen::MatrixXd;

// Thin QR of A (n x c): Q is n x m with m = min(n, c), orthonormal columns, range(A) in range(Q).
inline MatrixXd ThinQ(const MatrixXd& A)
{
    const long n = A.rows(), c = A.cols(), m = std::min(n, c);
    MatrixXd R = A;
    std::vector<std::vector<double>> vs;
    for (long k = 0; k < m; k++) {
        std::vector<double> v((size_t)(n - k));
        double nrm = 0;
        for (long i = k; i < n; i++) {
            v[(size_t)(i - k)] = R(i, k);
            nrm += R(i, k) * R(i, k);
        }
        nrm = std::sqrt(nrm);
        if (nrm > 0) {
            v[0] += (v[0] >= 0 ? nrm : -nrm);
            double vn = 0;
            for (double x : v) vn += x * x;
            if (vn > 0) {
                for (long j = k; j < c; j++) {
                    double s = 0;
                    for (long i = k; i < n; i++) s += v[(size_t)(i - k)] * R(i, j);
                    s = 2 * s / vn;
                    for (long i = k; i < n; i++) R(i, j) -= s * v[(size_t)(i - k)];
                }
            } else {
                v.assign(v.size(), 0.0);
            }
        } else {
            v.assign(v.size(), 0.0);
        }
        vs.push_back(std::move(v));
    }
    MatrixXd Q = MatrixXd::Identity(n, m);
    for (long k = m - 1; k >= 0; k--) {
        const auto& v = vs[(size_t)k];
        double vn = 0;
        for (double x : v) vn += x * x;
        if (vn == 0) continue;
        for (long j = 0; j < m; j++) {
            double s = 0;
            for (long i = k; i < n; i++) s += v[(size_t)(i - k)] * Q(i, j);
            s = 2 * s / vn;
            for (long i = k; i < n; i++) Q(i, j) -= s * v[(size_t)(i - k)];
        }
    }
    return Q;
}

// Left singular vectors and singular values of A (r x c), sorted descending; U is r x min(r, c).
inline void LeftSVD(const MatrixXd& A, MatrixXd& U, std::vector<double>& sigma)
{
    const long r = A.rows(), c = A.cols();
    // Hestenes on W = A^T (c x r): rotate column pairs until orthogonal; A = V diag(|w_j|) (W V / |w_j|)^T
    MatrixXd W = A.transpose();
    MatrixXd V = MatrixXd::Identity(r, r);
    const double eps = 1e-15;
    for (int sweep = 0; sweep < 60; sweep++) {
        bool rotated = false;
        for (long p = 0; p < r - 1; p++)
            for (long q = p + 1; q < r; q++) {
                double app = 0, aqq = 0, apq = 0;
                for (long i = 0; i < c; i++) {
                    app += W(i, p) * W(i, p);
                    aqq += W(i, q) * W(i, q);
                    apq += W(i, p) * W(i, q);
                }
                if (std::fabs(apq) <= eps * std::sqrt(app * aqq) || apq == 0) continue;
                rotated = true;
                const double zeta = (aqq - app) / (2 * apq);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                const double cs = 1 / std::sqrt(1 + t * t), sn = cs * t;
                for (long i = 0; i < c; i++) {
                    const double wp = W(i, p), wq = W(i, q);
                    W(i, p) = cs * wp - sn * wq;
                    W(i, q) = sn * wp + cs * wq;
                }
                for (long i = 0; i < r; i++) {
                    const double vp = V(i, p), vq = V(i, q);
                    V(i, p) = cs * vp - sn * vq;
                    V(i, q) = sn * vp + cs * vq;
                }
            }
        if (!rotated) break;
    }
    std::vector<double> s((size_t)r);
    for (long j = 0; j < r; j++) {
        double n2 = 0;
        for (long i = 0; i < c; i++) n2 += W(i, j) * W(i, j);
        s[(size_t)j] = std::sqrt(n2);
    }
    std::vector<long> order((size_t)r);
    std::iota(order.begin(), order.end(), 0L);
    std::stable_sort(order.begin(), order.end(), [&](long a, long b) { return s[(size_t)a] > s[(size_t)b]; });
    const long k = std::min(r, c);
    U = MatrixXd(r, k);
    sigma.assign((size_t)k, 0.0);
    for (long j = 0; j < k; j++) {
        sigma[(size_t)j] = s[(size_t)order[(size_t)j]];
        for (long i = 0; i < r; i++) U(i, j) = V(i, order[(size_t)j]);
    }
}

}  // namespace la
}  // namespace VlasovTucker
