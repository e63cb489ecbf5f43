// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).
// Restates src/tucker.cpp:34-98, 100-145, 190-316, 442-465 (Tucker algebra and rounding) and the
// TensorType = Tucker instantiation of src/solver.cpp:141-212, 258-293, 314-361 and
// src/particle_data.cpp:23-125.  Eigen's dense kernels (BDCSVD / JacobiSVD, ColPivHouseholderQR,
// GEMM, kroneckerProduct) live in the missing Eigen/src/Core + SVD/QR modules; they are restated
// by their published algorithms: Householder QR and a one-sided Jacobi SVD (singular vectors are
// unique up to sign/rotation inside degenerate subspaces, which is why Tucker parity is defined
// on reconstructed tensors and moments).  The dense R0*G*kron(R1,R2)^T of tucker.cpp:83 is
// evaluated as three sequential mode products (mathematically identical).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <numeric>
#include <stdexcept>

#include "oracle.h"

namespace oracle {

namespace {

// column-major matrix
struct Mat {
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)i + (size_t)r * j]; }
    double operator()(int i, int j) const { return a[(size_t)i + (size_t)r * j]; }
};

Mat MatMul(const Mat& A, const Mat& B)
{
    Mat C(A.r, B.c);
    for (int j = 0; j < B.c; j++)
        for (int k = 0; k < A.c; k++) {
            const double b = B(k, j);
            for (int i = 0; i < A.r; i++) C(i, j) += A(i, k) * b;
        }
    return C;
}
Mat Transposed(const Mat& A)
{
    Mat T(A.c, A.r);
    for (int i = 0; i < A.r; i++)
        for (int j = 0; j < A.c; j++) T(j, i) = A(i, j);
    return T;
}

// Householder QR: orthonormal Q (n x min(n,c)) with range(A) inside range(Q)   (tucker.cpp:72-79)
Mat ThinQ(const Mat& A)
{
    const int n = A.r, c = A.c, m = std::min(n, c);
    Mat R = A;
    std::vector<std::vector<double>> refl;
    for (int k = 0; k < m; k++) {
        std::vector<double> v(n - k);
        double s = 0;
        for (int i = k; i < n; i++) {
            v[i - k] = R(i, k);
            s += v[i - k] * v[i - k];
        }
        const double alpha = std::sqrt(s);
        if (alpha > 0) {
            v[0] += v[0] >= 0 ? alpha : -alpha;
            double vv = 0;
            for (double x : v) vv += x * x;
            for (int j = k; j < c; j++) {
                double d = 0;
                for (int i = k; i < n; i++) d += v[i - k] * R(i, j);
                d *= 2 / vv;
                for (int i = k; i < n; i++) R(i, j) -= d * v[i - k];
            }
        } else {
            std::fill(v.begin(), v.end(), 0.0);
        }
        refl.push_back(v);
    }
    Mat Q(n, m);
    for (int i = 0; i < m; i++) Q(i, i) = 1;
    for (int k = m - 1; k >= 0; k--) {
        const auto& v = refl[k];
        double vv = 0;
        for (double x : v) vv += x * x;
        if (vv == 0) continue;
        for (int j = 0; j < m; j++) {
            double d = 0;
            for (int i = k; i < n; i++) d += v[i - k] * Q(i, j);
            d *= 2 / vv;
            for (int i = k; i < n; i++) Q(i, j) -= d * v[i - k];
        }
    }
    return Q;
}

// thin left singular vectors (r x min(r,c)) and singular values, descending   (tucker.cpp:446-448)
void LeftSVD(const Mat& A, Mat& U, std::vector<double>& sv)
{
    const int r = A.r, c = A.c;
    Mat W = Transposed(A);   // c x r; orthogonalise its columns (Hestenes)
    Mat V(r, r);
    for (int i = 0; i < r; i++) V(i, i) = 1;
    for (int sweep = 0; sweep < 80; sweep++) {
        bool any = false;
        for (int p = 0; p + 1 < r; p++)
            for (int q = p + 1; q < r; q++) {
                double a = 0, b = 0, g = 0;
                for (int i = 0; i < c; i++) {
                    a += W(i, p) * W(i, p);
                    b += W(i, q) * W(i, q);
                    g += W(i, p) * W(i, q);
                }
                if (g == 0 || std::fabs(g) <= 1e-15 * std::sqrt(a * b)) continue;
                any = true;
                const double z = (b - a) / (2 * g);
                const double t = (z >= 0 ? 1.0 : -1.0) / (std::fabs(z) + std::sqrt(1 + z * z));
                const double cs = 1 / std::sqrt(1 + t * t), sn = cs * t;
                for (int i = 0; i < c; i++) {
                    const double x = W(i, p), y = W(i, q);
                    W(i, p) = cs * x - sn * y;
                    W(i, q) = sn * x + cs * y;
                }
                for (int i = 0; i < r; i++) {
                    const double x = V(i, p), y = V(i, q);
                    V(i, p) = cs * x - sn * y;
                    V(i, q) = sn * x + cs * y;
                }
            }
        if (!any) break;
    }
    std::vector<double> s(r);
    for (int j = 0; j < r; j++) {
        double n2 = 0;
        for (int i = 0; i < c; i++) n2 += W(i, j) * W(i, j);
        s[j] = std::sqrt(n2);
    }
    std::vector<int> ord(r);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return s[x] > s[y]; });
    const int k = std::min(r, c);
    U = Mat(r, k);
    sv.assign(k, 0.0);
    for (int j = 0; j < k; j++) {
        sv[j] = s[ord[j]];
        for (int i = 0; i < r; i++) U(i, j) = V(i, ord[j]);
    }
}

// dense 3-D tensor, i0 fastest
struct Ten {
    int d[3] = {0, 0, 0};
    std::vector<double> a;
    Ten() {}
    Ten(int d0, int d1, int d2) : a((size_t)d0 * d1 * d2, 0.0)
    {
        d[0] = d0;
        d[1] = d1;
        d[2] = d2;
    }
    double& operator()(int i0, int i1, int i2) { return a[(size_t)i0 + (size_t)d[0] * (i1 + (size_t)d[1] * i2)]; }
    double operator()(int i0, int i1, int i2) const { return a[(size_t)i0 + (size_t)d[0] * (i1 + (size_t)d[1] * i2)]; }
};

// Y = X x_mode M, M is (q x d[mode])
Ten ModeProduct(const Ten& X, const Mat& M, int mode)
{
    int e[3] = {X.d[0], X.d[1], X.d[2]};
    e[mode] = M.r;
    Ten Y(e[0], e[1], e[2]);
    const size_t stride = mode == 0 ? 1 : (mode == 1 ? (size_t)X.d[0] : (size_t)X.d[0] * X.d[1]);
    const size_t ostride = mode == 0 ? 1 : (mode == 1 ? (size_t)e[0] : (size_t)e[0] * e[1]);
    for (int i2 = 0; i2 < (mode == 2 ? 1 : X.d[2]); i2++)
        for (int i1 = 0; i1 < (mode == 1 ? 1 : X.d[1]); i1++)
            for (int i0 = 0; i0 < (mode == 0 ? 1 : X.d[0]); i0++) {
                const size_t in0 = (size_t)i0 + (size_t)X.d[0] * (i1 + (size_t)X.d[1] * i2);
                const size_t out0 = (size_t)i0 + (size_t)e[0] * (i1 + (size_t)e[1] * i2);
                for (int q = 0; q < M.r; q++) {
                    double s = 0;
                    for (int k = 0; k < X.d[mode]; k++) s += M(q, k) * X.a[in0 + k * stride];
                    Y.a[out0 + q * ostride] = s;
                }
            }
    return Y;
}

// Unfolding with the reference's column order (tucker.cpp:337-392)
Mat Unfold(const Ten& t, int mode)
{
    const int a = (mode + 1) % 3, b = (mode + 2) % 3;
    Mat m(t.d[mode], t.d[a] * t.d[b]);
    int i[3];
    for (i[2] = 0; i[2] < t.d[2]; i[2]++)
        for (i[1] = 0; i[1] < t.d[1]; i[1]++)
            for (i[0] = 0; i[0] < t.d[0]; i[0]++) m(i[mode], i[b] + i[a] * t.d[b]) = t(i[0], i[1], i[2]);
    return m;
}

}  // namespace

struct TuckerT {
    int n[3] = {0, 0, 0}, r[3] = {0, 0, 0};
    Mat U[3];
    Ten core;

    TuckerT() {}
    // Tucker(tensor, precision, maxRank), tucker.cpp:34-50 with _ComputeU :442-465
    TuckerT(const Ten& x, double precision, int rmax)
    {
        for (int i = 0; i < 3; i++) n[i] = x.d[i];
        for (int i = 0; i < 3; i++) {
            Mat Ui;
            std::vector<double> sv;
            LeftSVD(Unfold(x, i), Ui, sv);
            double n2 = 0;
            for (double s : sv) n2 += s * s;
            const double threshold = precision * std::sqrt(n2) / std::sqrt(3);
            std::vector<int> keep;
            for (int j = 0; j < Ui.c; j++)
                if (keep.empty() || (sv[j] > threshold && (int)keep.size() < rmax)) keep.push_back(j);
            U[i] = Mat(Ui.r, (int)keep.size());
            for (size_t cidx = 0; cidx < keep.size(); cidx++)
                for (int row = 0; row < Ui.r; row++) U[i](row, (int)cidx) = Ui(row, keep[cidx]);
            r[i] = (int)keep.size();
        }
        Ten c = ModeProduct(x, Transposed(U[0]), 0);
        c = ModeProduct(c, Transposed(U[1]), 1);
        core = ModeProduct(c, Transposed(U[2]), 2);
    }

    // tucker.cpp:66-98
    void Compress(double precision, int rmax)
    {
        Mat Q[3], R[3];
        for (int i = 0; i < 3; i++) {
            Q[i] = ThinQ(U[i]);
            R[i] = MatMul(Transposed(Q[i]), U[i]);
        }
        Ten aux = ModeProduct(core, R[0], 0);
        aux = ModeProduct(aux, R[1], 1);
        aux = ModeProduct(aux, R[2], 2);
        TuckerT small(aux, precision, rmax);
        core = small.core;
        for (int i = 0; i < 3; i++) {
            U[i] = MatMul(Q[i], small.U[i]);
            r[i] = U[i].c;
        }
    }

    // tucker.cpp:100-104
    Ten Reconstructed() const
    {
        Ten t = ModeProduct(core, U[0], 0);
        t = ModeProduct(t, U[1], 1);
        return ModeProduct(t, U[2], 2);
    }

    // tucker.cpp:131-145 sums every reconstructed entry; same value by linearity
    double Sum() const
    {
        std::vector<double> s[3];
        for (int i = 0; i < 3; i++) {
            s[i].assign(r[i], 0.0);
            for (int j = 0; j < r[i]; j++)
                for (int k = 0; k < n[i]; k++) s[i][j] += U[i](k, j);
        }
        double total = 0;
        for (int j0 = 0; j0 < r[0]; j0++)
            for (int j1 = 0; j1 < r[1]; j1++)
                for (int j2 = 0; j2 < r[2]; j2++) total += core(j0, j1, j2) * s[0][j0] * s[1][j1] * s[2][j2];
        return total;
    }
};

namespace {

// tucker.cpp:190-228
TuckerT Add(const TuckerT& a, const TuckerT& b)
{
    TuckerT r;
    for (int i = 0; i < 3; i++) {
        r.n[i] = a.n[i];
        r.r[i] = a.r[i] + b.r[i];
        r.U[i] = Mat(a.n[i], r.r[i]);
        for (int row = 0; row < a.n[i]; row++) {
            for (int j = 0; j < a.r[i]; j++) r.U[i](row, j) = a.U[i](row, j);
            for (int j = 0; j < b.r[i]; j++) r.U[i](row, a.r[i] + j) = b.U[i](row, j);
        }
    }
    r.core = Ten(r.r[0], r.r[1], r.r[2]);
    for (int k2 = 0; k2 < a.r[2]; k2++)
        for (int k1 = 0; k1 < a.r[1]; k1++)
            for (int k0 = 0; k0 < a.r[0]; k0++) r.core(k0, k1, k2) = a.core(k0, k1, k2);
    for (int k2 = 0; k2 < b.r[2]; k2++)
        for (int k1 = 0; k1 < b.r[1]; k1++)
            for (int k0 = 0; k0 < b.r[0]; k0++) r.core(a.r[0] + k0, a.r[1] + k1, a.r[2] + k2) = b.core(k0, k1, k2);
    return r;
}
// tucker.cpp:302-316
TuckerT Scale(double d, const TuckerT& t)
{
    TuckerT r = t;
    for (auto& x : r.core.a) x *= d;
    return r;
}
TuckerT Sub(const TuckerT& a, const TuckerT& b) { return Add(a, Scale(-1.0, b)); }   // tucker.cpp:254-257
// tucker.cpp:259-300
TuckerT Hadamard(const TuckerT& a, const TuckerT& b)
{
    TuckerT r;
    for (int i = 0; i < 3; i++) {
        r.n[i] = a.n[i];
        r.r[i] = a.r[i] * b.r[i];
        r.U[i] = Mat(a.n[i], r.r[i]);
        for (int row = 0; row < a.n[i]; row++)
            for (int ja = 0; ja < a.r[i]; ja++)
                for (int jb = 0; jb < b.r[i]; jb++) r.U[i](row, ja * b.r[i] + jb) = a.U[i](row, ja) * b.U[i](row, jb);
    }
    r.core = Ten(r.r[0], r.r[1], r.r[2]);
    for (int k2 = 0; k2 < r.r[2]; k2++)
        for (int k1 = 0; k1 < r.r[1]; k1++)
            for (int k0 = 0; k0 < r.r[0]; k0++)
                r.core(k0, k1, k2) = a.core(k0 / b.r[0], k1 / b.r[1], k2 / b.r[2]) * b.core(k0 % b.r[0], k1 % b.r[1], k2 % b.r[2]);
    return r;
}

Ten TenFrom(const std::vector<double>& v, const int n[3])
{
    Ten t(n[0], n[1], n[2]);
    t.a = v;
    return t;
}

}  // namespace

// Solver<Tucker> + ParticleData<Tucker> for one species on a periodic / walled mesh
struct TuckerSim {
    const Mesh* mesh;
    VGrid vg;
    double mass = 1, charge = 1, timeStep = 0, comprErr = 1e-10;
    int maxRank;
    Vec3 externalField = {0, 0, 0};
    std::vector<TuckerT> pdf;
    std::vector<int> faceBCType;
    std::vector<int> faceSource;                // index into `sources` for Source faces
    std::vector<char> faceCollect;              // ParticleBC::collectCharge
    std::vector<TuckerT> sources;
    std::map<int, double> wallCharge;           // entity -> absorbed charge (solver.cpp:171-178)
    std::vector<TuckerT> vNormal, vNormalAbs;   // per face (solver.cpp:258-293)

    // particle_data.cpp:23-90: the initial tensors are built with precision 0 (uncompressed ranks)
    void SetDense(const double* f)
    {
        const int nT = mesh->nTets(), N = vg.nTotal;
        pdf.clear();
        for (int t = 0; t < nT; t++) {
            Ten x(vg.n[0], vg.n[1], vg.n[2]);
            std::memcpy(x.a.data(), f + (size_t)t * N, (size_t)N * 8);
            pdf.push_back(TuckerT(x, 0.0, 1000000));
        }
    }
    void Precompute()
    {
        const size_t nf = mesh->facePoints.size();
        vNormal.resize(nf);
        vNormalAbs.resize(nf);
        for (size_t fi = 0; fi < nf; fi++) {
            const Vec3& nrm = mesh->faceNormal[fi];
            Ten vn(vg.n[0], vg.n[1], vg.n[2]), va(vg.n[0], vg.n[1], vg.n[2]);
            for (int e = 0; e < vg.nTotal; e++) {
                vn.a[e] = nrm[0] * vg.v[0][e] + nrm[1] * vg.v[1][e] + nrm[2] * vg.v[2][e];
                va.a[e] = std::fabs(vn.a[e]);
            }
            vNormal[fi] = TuckerT(vn, 0.0, 1000000);
            vNormalAbs[fi] = TuckerT(va, 0.0, 1000000);
            vNormal[fi].Compress(comprErr, 1000000);
            vNormalAbs[fi].Compress(comprErr, 6);
        }
    }
    // solver.cpp:314-346
    TuckerT Flux(int t, int f) const
    {
        const int fi = 4 * t + f;
        const int bc = faceBCType[fi];
        const TuckerT& A = pdf[t];
        if (bc == PBC_NonBoundary || bc == PBC_Periodic) {
            const TuckerT& B = pdf[mesh->adj[fi]];
            return Scale(0.5, Sub(Hadamard(vNormal[fi], Add(B, A)), Hadamard(vNormalAbs[fi], Sub(B, A))));
        } else if (bc == PBC_Absorbing) {
            return Scale(0.5, Add(Hadamard(vNormal[fi], A), Hadamard(vNormalAbs[fi], A)));
        }
        if (bc == PBC_Source) {
            const TuckerT& B = sources[faceSource[fi]];
            return Scale(0.5, Sub(Hadamard(vNormal[fi], Add(B, A)), Hadamard(vNormalAbs[fi], Sub(B, A))));
        }
        return Hadamard(vNormal[fi], A);   // Free
    }
    // solver.cpp:348-361: U_k <- D_k U_k
    TuckerT Derivative(int t, int k) const
    {
        TuckerT d = pdf[t];
        const int nk = vg.n[k];
        Mat D(nk, nk);
        for (int i = 0; i < nk; i++)
            for (int j = 0; j < nk; j++) D(i, j) = vg.d[k][(size_t)i * nk + j];
        d.U[k] = MatMul(D, d.U[k]);
        return d;
    }
    // solver.cpp:141-212
    void UpdatePDF(const double* E)
    {
        const int nT = mesh->nTets();
        std::vector<TuckerT> rhs(nT);
        const Ten zero(vg.n[0], vg.n[1], vg.n[2]);
        for (int t = 0; t < nT; t++) rhs[t] = TuckerT(zero, 0.0, 1000000);
#pragma omp parallel for schedule(dynamic)
        for (int t = 0; t < nT; t++)
            for (int f = 0; f < 4; f++) {
                const int fi = 4 * t + f;
                const TuckerT flux = Flux(t, f);
                rhs[t] = Sub(rhs[t], Scale(mesh->faceArea[fi] / mesh->tetVolume[t], flux));
                if (faceBCType[fi] == PBC_Absorbing && faceCollect[fi]) {
                    const double dq = charge * (timeStep * mesh->faceArea[fi] * flux.Sum() * vg.cellVolume);
#pragma omp critical
                    wallCharge[mesh->faceEntity[fi]] += dq;
                }
                rhs[t].Compress(comprErr, maxRank);
            }
#pragma omp parallel for schedule(dynamic)
        for (int t = 0; t < nT; t++) {
            for (int k = 0; k < 3; k++) {
                const double force = (charge / mass) * (E[3 * t + k] + externalField[k]);
                rhs[t] = Sub(rhs[t], Scale(force, Derivative(t, k)));
            }
            rhs[t].Compress(comprErr, maxRank);
        }
#pragma omp parallel for schedule(dynamic)
        for (int t = 0; t < nT; t++) {
            pdf[t] = Add(pdf[t], Scale(timeStep, rhs[t]));
            pdf[t].Compress(comprErr, maxRank);
        }
    }
};

}  // namespace oracle

using namespace oracle;

extern "C" {

// stand-alone Tucker objects (tucker_test.cpp usage)
void* orc_tucker_from_full(const double* x, const int* n, double precision, int rmax)
{
    return new TuckerT(TenFrom(std::vector<double>(x, x + (size_t)n[0] * n[1] * n[2]), n), precision, rmax);
}
void* orc_tucker_clone(void* h) { return new TuckerT(*(TuckerT*)h); }
void orc_tucker_free(void* h) { delete (TuckerT*)h; }
void orc_tucker_ranks(void* h, int* r)
{
    for (int i = 0; i < 3; i++) r[i] = ((TuckerT*)h)->r[i];
}
void orc_tucker_reconstruct(void* h, double* out)
{
    Ten t = ((TuckerT*)h)->Reconstructed();
    std::memcpy(out, t.a.data(), t.a.size() * 8);
}
void orc_tucker_compress(void* h, double precision, int rmax) { ((TuckerT*)h)->Compress(precision, rmax); }
double orc_tucker_sum(void* h) { return ((TuckerT*)h)->Sum(); }
// a <- a + s*b   (operator+=, operator-= with scalar operator*)
void orc_tucker_axpy(void* a, double s, void* b) { *(TuckerT*)a = Add(*(TuckerT*)a, Scale(s, *(TuckerT*)b)); }
void orc_tucker_hadamard(void* a, void* b) { *(TuckerT*)a = Hadamard(*(TuckerT*)a, *(TuckerT*)b); }

// Solver<Tucker>
void* orc_tsim_create(void* mesh, const int* n, const double* minV, const double* maxV, double mass, double charge,
                      double comprErr, int maxRank)
{
    TuckerSim* s = new TuckerSim();
    s->mesh = (Mesh*)mesh;
    s->vg = MakeVGrid({n[0], n[1], n[2]}, {minV[0], minV[1], minV[2]}, {maxV[0], maxV[1], maxV[2]});
    s->mass = mass;
    s->charge = charge;
    s->comprErr = comprErr;
    s->maxRank = maxRank > 0 ? maxRank : std::max({n[0], n[1], n[2]});   // particle_data.cpp:18
    const size_t nf = s->mesh->facePoints.size();
    s->faceBCType.assign(nf, PBC_NonBoundary);
    s->faceSource.assign(nf, -1);
    s->faceCollect.assign(nf, 0);
    for (auto& pr : s->mesh->periodicPairs)
        for (int mark : pr)
            for (size_t i = 0; i < nf; i++)
                if (s->mesh->faceEntity[i] == mark) s->faceBCType[i] = PBC_Periodic;
    return s;
}
void orc_tsim_free(void* h) { delete (TuckerSim*)h; }
void orc_tsim_set_particle_bc(void* h, int entity, int type)
{
    TuckerSim& s = *(TuckerSim*)h;
    for (size_t i = 0; i < s.mesh->facePoints.size(); i++)
        if (s.mesh->faceEntity[i] == entity) s.faceBCType[i] = type;
}
// SetParticleBC with collectCharge / sourcePDF (dense, precision 0); source may be NULL
void orc_tsim_set_particle_bc_ex(void* h, int entity, int type, int collect, const double* source)
{
    TuckerSim& s = *(TuckerSim*)h;
    int id = -1;
    if (source) {
        Ten x(s.vg.n[0], s.vg.n[1], s.vg.n[2]);
        std::memcpy(x.a.data(), source, (size_t)s.vg.nTotal * 8);
        s.sources.push_back(TuckerT(x, 0.0, 1000000));
        id = (int)s.sources.size() - 1;
    }
    for (size_t i = 0; i < s.mesh->facePoints.size(); i++)
        if (s.mesh->faceEntity[i] == entity) {
            s.faceBCType[i] = type;
            s.faceCollect[i] = (char)collect;
            s.faceSource[i] = id;
        }
}
double orc_tsim_wall_charge(void* h, int entity)
{
    TuckerSim& s = *(TuckerSim*)h;
    auto it = s.wallCharge.find(entity);
    return it == s.wallCharge.end() ? 0.0 : it->second;
}
void orc_tsim_set_pdf(void* h, const double* f)
{
    TuckerSim& s = *(TuckerSim*)h;
    s.SetDense(f);
    if (s.vNormal.empty()) s.Precompute();
}
void orc_tsim_get_pdf(void* h, double* f)
{
    TuckerSim& s = *(TuckerSim*)h;
    for (int t = 0; t < s.mesh->nTets(); t++) {
        Ten x = s.pdf[t].Reconstructed();
        std::memcpy(f + (size_t)t * s.vg.nTotal, x.a.data(), (size_t)s.vg.nTotal * 8);
    }
}
void orc_tsim_ranks(void* h, int* r)
{
    TuckerSim& s = *(TuckerSim*)h;
    for (int t = 0; t < s.mesh->nTets(); t++)
        for (int i = 0; i < 3; i++) r[3 * t + i] = s.pdf[t].r[i];
}
void orc_tsim_vnabs(void* h, int face, double* out)
{
    TuckerSim& s = *(TuckerSim*)h;
    Ten x = s.vNormalAbs[face].Reconstructed();
    std::memcpy(out, x.a.data(), x.a.size() * 8);
}
void orc_tsim_update_pdf(void* h, double dt, const double* E, const double* ext)
{
    TuckerSim& s = *(TuckerSim*)h;
    s.timeStep = dt;
    if (ext) s.externalField = {ext[0], ext[1], ext[2]};
    s.UpdatePDF(E);
}
void orc_tsim_density(void* h, double* out)
{
    TuckerSim& s = *(TuckerSim*)h;
    for (int t = 0; t < s.mesh->nTets(); t++) out[t] = s.pdf[t].Sum() * s.vg.cellVolume;   // particle_data.cpp:99
}
}
