// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).
// Restates src/poisson.cpp:14-455.  The linear solve restates Eigen 3.4.0's
// ConjugateGradient<SparseMatrix<double>, Upper, DiagonalPreconditioner>
// (include/eigen-3.4.0/Eigen/src/IterativeLinearSolvers/ConjugateGradient.h:28-91,
// IterativeSolverBase.h: tolerance = epsilon, maxIterations = 2*cols;
// BasicPreconditioners.h: invdiag = 1/diag, or 1 where diag == 0).  The reference's
// default SparseLU (poisson.h:20) is a direct solve; the oracle substitutes the same CG
// run to its 2.2e-16 tolerance and says so — tests cross-check it against scipy's splu.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>

namespace oracle {

namespace {
const double kEps0 = 8.85e-12;  // constants.h:10
Vec3 sub(const Vec3& a, const Vec3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
Vec3 add(const Vec3& a, const Vec3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
Vec3 divs(const Vec3& a, double d) { return {a[0] / d, a[1] / d, a[2] / d}; }
Vec3 muls(const Vec3& a, double d) { return {a[0] * d, a[1] * d, a[2] * d}; }
double dot(const Vec3& a, const Vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
double norm(const Vec3& a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// Eigen::Matrix3d::fullPivLu().solve(rhs) (poisson.cpp:440): Gaussian elimination with
// complete pivoting.
Vec3 FullPivSolve3(double m[3][3], const Vec3& rhs)
{
    double a[3][4];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) a[i][j] = m[i][j];
        a[i][3] = rhs[i];
    }
    int colPerm[3] = {0, 1, 2};
    for (int k = 0; k < 3; k++) {
        int pr = k, pc = k;
        double best = -1;
        for (int i = k; i < 3; i++)
            for (int j = k; j < 3; j++)
                if (std::fabs(a[i][j]) > best) {
                    best = std::fabs(a[i][j]);
                    pr = i;
                    pc = j;
                }
        if (pr != k)
            for (int j = 0; j < 4; j++) std::swap(a[k][j], a[pr][j]);
        if (pc != k) {
            for (int i = 0; i < 3; i++) std::swap(a[i][k], a[i][pc]);
            std::swap(colPerm[k], colPerm[pc]);
        }
        for (int i = k + 1; i < 3; i++) {
            double l = a[i][k] / a[k][k];
            for (int j = k; j < 4; j++) a[i][j] -= l * a[k][j];
        }
    }
    double y[3];
    for (int i = 2; i >= 0; i--) {
        double s = a[i][3];
        for (int j = i + 1; j < 3; j++) s -= a[i][j] * y[j];
        y[i] = s / a[i][i];
    }
    Vec3 x;
    for (int i = 0; i < 3; i++) x[colPerm[i]] = y[i];
    return x;
}
}  // namespace

Poisson::Poisson(const Mesh* m) : mesh(m)
{
    solutionIsUnique = false;
    faceBC.assign(m->facePoints.size(), PoissonBC());
    PoissonBC periodic;
    periodic.type = QBC_Periodic;
    for (auto& pr : m->periodicPairs)
        for (int mark : pr) SetBC(mark, periodic);
}

void Poisson::SetBC(int boundaryInd, const PoissonBC& bc)
{
    if (bc.type == QBC_Dirichlet) solutionIsUnique = true;
    auto it = mesh->entityToFaces.find(boundaryInd);
    if (it == mesh->entityToFaces.end()) throw std::out_of_range("EntityToFaces().at");  // poisson.cpp:90
    for (int fi : it->second) faceBC[fi] = bc;
}

// d = adjCentroid - centroid + faceCentroid - partnerFaceCentroid (poisson.cpp:152-162)
Vec3 Poisson::PeriodicShiftedDistance(int t, int j) const
{
    int a = mesh->adj[4 * t + j];
    Vec3 d = sub(mesh->tetCentroid[a], mesh->tetCentroid[t]);
    int k = 0;
    while (mesh->adj[4 * a + k] != t) k++;
    return sub(add(d, mesh->faceCentroid[4 * t + j]), mesh->faceCentroid[4 * a + k]);
}

void Poisson::Initialize()
{
    int n = mesh->nTets();
    std::vector<std::map<int, double>> rows(n);
    for (int i = 0; i < n; i++) {
        // _FillLineCoeffs, poisson.cpp:126-177
        if (!solutionIsUnique && i == 0) {
            rows[0][0] += 1;
            continue;
        }
        for (int j = 0; j < 4; j++) {
            int fi = 4 * i + j;
            int bc = faceBC[fi].type;
            double A = mesh->faceArea[fi];
            const Vec3& nrm = mesh->faceNormal[fi];
            if (bc == QBC_NonBoundary) {
                int a = mesh->adj[fi];
                Vec3 d = sub(mesh->tetCentroid[a], mesh->tetCentroid[i]);
                rows[i][a] += A / dot(d, nrm);
                rows[i][i] += -A / dot(d, nrm);
            } else if (bc == QBC_Periodic) {
                int a = mesh->adj[fi];
                Vec3 d = PeriodicShiftedDistance(i, j);
                rows[i][a] += A / dot(d, nrm);
                rows[i][i] += -A / dot(d, nrm);
            } else if (bc == QBC_Dirichlet) {
                Vec3 d = sub(mesh->faceCentroid[fi], mesh->tetCentroid[i]);
                rows[i][i] += -A / dot(d, nrm);
            }
        }
    }
    rowPtr.assign(n + 1, 0);
    colInd.clear();
    val.clear();
    for (int i = 0; i < n; i++) {
        for (auto& kv : rows[i]) {
            colInd.push_back(kv.first);
            val.push_back(kv.second);
        }
        rowPtr[i + 1] = (int)colInd.size();
    }
}

// Jacobi-PCG on the Upper self-adjoint view (poisson.h:41-44): entry (i,j), j>=i, is used
// for both (i,j) and (j,i); strictly-lower entries are ignored.
std::vector<double> Poisson::SolveSystem(const std::vector<double>& rhs, bool useGuess)
{
    int n = (int)rhs.size();
    auto spmv = [&](const std::vector<double>& x, std::vector<double>& y) {
        std::fill(y.begin(), y.end(), 0.0);
        for (int i = 0; i < n; i++)
            for (int k = rowPtr[i]; k < rowPtr[i + 1]; k++) {
                int j = colInd[k];
                if (j < i) continue;
                y[i] += val[k] * x[j];
                if (j != i) y[j] += val[k] * x[i];
            }
    };
    std::vector<double> invDiag(n, 1.0);
    for (int i = 0; i < n; i++)
        for (int k = rowPtr[i]; k < rowPtr[i + 1]; k++)
            if (colInd[k] == i && val[k] != 0) invDiag[i] = 1.0 / val[k];

    std::vector<double> x(n, 0.0);
    if (useGuess && (int)guess.size() == n) x = guess;  // poisson.cpp:47-48

    double tol = std::numeric_limits<double>::epsilon();
    int maxIters = 2 * n;
    std::vector<double> residual(n), p(n), z(n), tmp(n);
    spmv(x, tmp);
    for (int i = 0; i < n; i++) residual[i] = rhs[i] - tmp[i];
    double rhsNorm2 = 0;
    for (double v : rhs) rhsNorm2 += v * v;
    if (rhsNorm2 == 0) {
        lastIterations = 0;
        lastError = 0;
        return std::vector<double>(n, 0.0);
    }
    double threshold = std::max(tol * tol * rhsNorm2, std::numeric_limits<double>::min());
    double residualNorm2 = 0;
    for (double v : residual) residualNorm2 += v * v;
    if (residualNorm2 < threshold) {
        lastIterations = 0;
        lastError = std::sqrt(residualNorm2 / rhsNorm2);
        return x;
    }
    for (int i = 0; i < n; i++) p[i] = invDiag[i] * residual[i];
    double absNew = 0;
    for (int i = 0; i < n; i++) absNew += residual[i] * p[i];
    int it = 0;
    while (it < maxIters) {
        spmv(p, tmp);
        double pAp = 0;
        for (int i = 0; i < n; i++) pAp += p[i] * tmp[i];
        double alpha = absNew / pAp;
        for (int i = 0; i < n; i++) x[i] += alpha * p[i];
        for (int i = 0; i < n; i++) residual[i] -= alpha * tmp[i];
        residualNorm2 = 0;
        for (double v : residual) residualNorm2 += v * v;
        if (residualNorm2 < threshold) break;
        for (int i = 0; i < n; i++) z[i] = invDiag[i] * residual[i];
        double absOld = absNew;
        absNew = 0;
        for (int i = 0; i < n; i++) absNew += residual[i] * z[i];
        double beta = absNew / absOld;
        for (int i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
        it++;
    }
    lastIterations = it;
    totalIterations += it;
    lastError = std::sqrt(residualNorm2 / rhsNorm2);
    return x;
}

Vec3 Poisson::WeightedGradient(int t, int f) const
{
    int fi = 4 * t + f;
    int bc = faceBC[fi].type;
    Vec3 d = sub(mesh->faceCentroid[fi], mesh->tetCentroid[t]);
    if (bc == QBC_NonBoundary || bc == QBC_Periodic) {
        int a = mesh->adj[fi];
        Vec3 adjD = sub(mesh->faceCentroid[fi], mesh->tetCentroid[a]);
        double g = norm(adjD) / (norm(adjD) + norm(d));
        double adjG = norm(d) / (norm(adjD) + norm(d));
        Vec3 w;
        for (int i = 0; i < 3; i++) w[i] = gradient[t][i] * g + gradient[a][i] * adjG;
        return w;
    }
    // Dirichlet (poisson.cpp:300-303); Neumann is never asked for (poisson.cpp:306-359)
    return gradient[t];
}

Vec3 Poisson::TetLSG(int t) const
{
    double valT = solution[t];
    double adjVal[4];
    Vec3 dist[4];
    for (int i = 0; i < 4; i++) {
        int fi = 4 * t + i;
        int bc = faceBC[fi].type;
        if (bc == QBC_NonBoundary) {
            int a = mesh->adj[fi];
            adjVal[i] = solution[a];
            dist[i] = sub(mesh->tetCentroid[a], mesh->tetCentroid[t]);
        } else if (bc == QBC_Dirichlet) {
            adjVal[i] = faceBC[fi].value;
            dist[i] = sub(mesh->faceCentroid[fi], mesh->tetCentroid[t]);
        } else if (bc == QBC_Neumann) {
            Vec3 d = sub(mesh->faceCentroid[fi], mesh->tetCentroid[t]);
            Vec3 x = muls(mesh->faceNormal[fi], dot(mesh->faceNormal[fi], d));
            dist[i] = x;
            adjVal[i] = valT + norm(x) * faceBC[fi].normalGrad;
        } else {  // Periodic
            adjVal[i] = solution[mesh->adj[fi]];
            dist[i] = PeriodicShiftedDistance(t, i);
        }
    }
    double w[4];
    for (int i = 0; i < 4; i++) w[i] = 1 / norm(dist[i]);
    double m[3][3];
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) {
            m[k][i] = 0;
            for (int j = 0; j < 4; j++) m[k][i] += 2 * w[j] * dist[j][k] * dist[j][i];
        }
    double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) -
                 m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                 m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    if (det == 0) throw std::runtime_error("Degenerate matrix in LSG.");
    Vec3 rhs;
    for (int k = 0; k < 3; k++) {
        rhs[k] = 0;
        for (int j = 0; j < 4; j++) rhs[k] -= 2 * w[j] * dist[j][k] * (valT - adjVal[j]);
    }
    return FullPivSolve3(m, rhs);
}

void Poisson::Solve(const std::vector<double>& rho)
{
    int n = mesh->nTets();
    std::vector<double> rhs(n);
    for (int i = 0; i < n; i++) {
        rhs[i] = (-rho[i] / kEps0) * mesh->tetVolume[i];
        // _FillLineRHS, poisson.cpp:246-274
        if (!solutionIsUnique && i == 0) {
            rhs[0] = 0;
            continue;
        }
        for (int j = 0; j < 4; j++) {
            int fi = 4 * i + j;
            if (faceBC[fi].type == QBC_Dirichlet) {
                Vec3 d = sub(mesh->faceCentroid[fi], mesh->tetCentroid[i]);
                rhs[i] -= mesh->faceArea[fi] / dot(d, mesh->faceNormal[fi]) * faceBC[fi].value;
            } else if (faceBC[fi].type == QBC_Neumann) {
                rhs[i] -= faceBC[fi].normalGrad * mesh->faceArea[fi];
            }
        }
    }
    auto computeGradient = [&]() {
        std::vector<Vec3> g(n);
        for (int i = 0; i < n; i++) g[i] = TetLSG(i);
        return g;
    };
    if (gradient.empty()) {
        // poisson.cpp:192-199: first call solves without correction (guess not yet set)
        solution = SolveSystem(rhs, false);
        gradient = computeGradient();
    }
    // _CorrectRHS, poisson.cpp:306-359
    for (int i = 0; i < n; i++) {
        if (!solutionIsUnique && i == 0) continue;
        for (int j = 0; j < 4; j++) {
            int fi = 4 * i + j;
            int bc = faceBC[fi].type;
            double cross = 0;
            Vec3 e;
            bool has = true;
            if (bc == QBC_NonBoundary) e = sub(mesh->tetCentroid[mesh->adj[fi]], mesh->tetCentroid[i]);
            else if (bc == QBC_Periodic) e = PeriodicShiftedDistance(i, j);
            else if (bc == QBC_Dirichlet) e = sub(mesh->faceCentroid[fi], mesh->tetCentroid[i]);
            else has = false;
            if (has) {
                e = divs(e, norm(e));
                Vec3 wg = WeightedGradient(i, j);
                const Vec3& nrm = mesh->faceNormal[fi];
                Vec3 q = sub(nrm, divs(e, dot(e, nrm)));
                cross = mesh->faceArea[fi] * dot(wg, q);
            }
            rhs[i] -= cross;
        }
    }
    guess = solution;                     // poisson.cpp:208
    solution = SolveSystem(rhs, true);    // poisson.cpp:211
    gradient = computeGradient();         // poisson.cpp:212
}

std::vector<Vec3> Poisson::ElectricField() const
{
    std::vector<Vec3> f(gradient.size());
    for (size_t i = 0; i < gradient.size(); i++)
        for (int k = 0; k < 3; k++) f[i][k] = -gradient[i][k];
    return f;
}

}  // namespace oracle
