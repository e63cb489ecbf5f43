// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).
// Restates src/full.cpp, src/velocity_grid.cpp, src/particle_data.cpp, src/solver.cpp and
// src/multicomponent_solver.cpp for TensorType = Full.  Compile with -ffp-contract=off:
// the reference materialises one tensor per operator (full.cpp:69-101), so no multiply
// and add of different operators can ever be contracted into an FMA there.
#include "oracle.h"

#include <cmath>
#include <stdexcept>

namespace oracle {

// ---------------------------------------------------------------- Full (full.cpp:24-101)
double Full::Sum() const
{
    double s = 0;
    for (double x : a) s += x;
    return s;
}
Full operator+(const Full& x, const Full& y)
{
    Full r(x.a.size());
    for (size_t i = 0; i < x.a.size(); i++) r.a[i] = x.a[i] + y.a[i];
    return r;
}
Full operator-(const Full& x, const Full& y)
{
    Full r(x.a.size());
    for (size_t i = 0; i < x.a.size(); i++) r.a[i] = x.a[i] - y.a[i];
    return r;
}
Full operator*(const Full& x, const Full& y)
{
    Full r(x.a.size());
    for (size_t i = 0; i < x.a.size(); i++) r.a[i] = x.a[i] * y.a[i];
    return r;
}
Full operator*(double d, const Full& x)
{
    Full r(x.a.size());
    for (size_t i = 0; i < x.a.size(); i++) r.a[i] = d * x.a[i];
    return r;
}

// ---------------------------------------------------------------- VelocityGrid (velocity_grid.cpp:9-53)
VGrid MakeVGrid(const std::array<int, 3>& n, const Vec3& minV, const Vec3& maxV)
{
    VGrid g;
    g.n = n;
    g.minV = minV;
    g.maxV = maxV;
    g.nTotal = n[0] * n[1] * n[2];
    for (int j = 0; j < 3; j++) g.step[j] = (maxV[j] - minV[j]) / (n[j] - 1);
    g.cellVolume = g.step[0] * g.step[1] * g.step[2];
    for (int j = 0; j < 3; j++) {
        g.v[j].assign(g.nTotal, 0.0);
        int ind[3];
        for (ind[0] = 0; ind[0] < n[0]; ind[0]++)
            for (ind[1] = 0; ind[1] < n[1]; ind[1]++)
                for (ind[2] = 0; ind[2] < n[2]; ind[2]++)
                    g.v[j][g.idx(ind[0], ind[1], ind[2])] = minV[j] + ind[j] * g.step[j];
    }
    for (int j = 0; j < 3; j++) {
        int nj = n[j];
        g.d[j].assign((size_t)nj * nj, 0.0);
        auto D = [&](int r, int c) -> double& { return g.d[j][(size_t)r * nj + c]; };
        D(0, 1) = 1;
        D(0, nj - 1) = 0;
        for (int i = 1; i < nj - 1; i++) {
            D(i, i + 1) = 1;
            D(i, i - 1) = -1;
        }
        D(nj - 1, 0) = 0;
        D(nj - 1, nj - 2) = -1;
        for (auto& x : g.d[j]) x /= 2 * g.step[j];
    }
    return g;
}

// ---------------------------------------------------------------- ParticleData (particle_data.cpp)
static const double kBoltz = 1.38e-23;   // constants.h:12
static const double kEps0 = 8.85e-12;    // constants.h:10

void Species::SetMaxwell(const std::vector<double>& physDensity, double temperature,
                         const Vec3& mpv)
{
    const VGrid& g = *vg;
    int n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
    std::vector<double> t3d(g.nTotal, 0.0);
    pdf.clear();
    for (int t = 0; t < mesh->nTets(); t++) {
        if (temperature != 0.0) {
            double normConst = 0;
            for (int i0 = 0; i0 < n0; i0++)
                for (int i1 = 0; i1 < n1; i1++)
                    for (int i2 = 0; i2 < n2; i2++) {
                        double at[3] = {g.minV[0] + i0 * g.step[0], g.minV[1] + i1 * g.step[1],
                                        g.minV[2] + i2 * g.step[2]};  // velocity_grid.cpp:55-60
                        double velSquared = 0;
                        for (int j = 0; j < 3; j++) {
                            double velJ = at[j] - mpv[j];
                            velSquared += velJ * velJ;
                        }
                        double e = std::exp(-mass * velSquared / (2 * kBoltz * temperature));
                        t3d[g.idx(i0, i1, i2)] = e;
                        normConst += e;
                    }
            for (int i0 = 0; i0 < n0; i0++)
                for (int i1 = 0; i1 < n1; i1++)
                    for (int i2 = 0; i2 < n2; i2++)
                        t3d[g.idx(i0, i1, i2)] *= physDensity[t] / (g.cellVolume * normConst);
        }
        if (temperature == 0.0) {
            std::fill(t3d.begin(), t3d.end(), 0.0);
            int i0 = (int)((mpv[0] - g.minV[0]) / g.step[0]);
            int i1 = (int)((mpv[1] - g.minV[1]) / g.step[1]);
            int i2 = (int)((mpv[2] - g.minV[2]) / g.step[2]);
            t3d[g.idx(i0, i1, i2)] = physDensity[t] / g.cellVolume;
        }
        pdf.push_back(Full(t3d));
    }
}

std::vector<double> Species::Density() const
{
    std::vector<double> r(mesh->nTets());
    for (int i = 0; i < mesh->nTets(); i++) r[i] = pdf[i].Sum() * vg->cellVolume;
    return r;
}

std::vector<Vec3> Species::Velocity() const
{
    std::vector<Vec3> r(mesh->nTets());
    std::vector<double> density = Density();
    for (int i = 0; i < mesh->nTets(); i++)
        for (int k = 0; k < 3; k++) {
            Full vPDF = Full(vg->v[k]) * pdf[i];
            r[i][k] = density[i] != 0 ? vPDF.Sum() * vg->cellVolume / density[i] : 0.0;
        }
    return r;
}

// ---------------------------------------------------------------- Solver<Full> (solver.cpp)
FullSolver::FullSolver(const Mesh* m, const VGrid* g, Species* s)
    : mesh(m), vg(g), sp(s), poisson(m)
{
    size_t nf = m->facePoints.size();
    faceBCType.assign(nf, PBC_NonBoundary);
    faceCollect.assign(nf, 0);
    faceSource.assign(nf, -1);
    for (auto& pr : m->periodicPairs)
        for (int mark : pr) SetParticleBC(mark, PBC_Periodic, false, -1);
}

void FullSolver::SetParticleBC(int boundaryInd, int type, bool collect, int sourceId)
{
    for (size_t i = 0; i < mesh->facePoints.size(); i++)
        if (mesh->faceEntity[i] == boundaryInd) {
            faceBCType[i] = type;
            faceCollect[i] = collect;
            faceSource[i] = sourceId;
        }
}

void FullSolver::SetFieldBCPotential(int boundaryInd, double potential)
{
    PoissonBC bc;
    bc.type = QBC_Dirichlet;
    bc.value = potential;
    poisson.SetBC(boundaryInd, bc);
}

void FullSolver::SetFieldBCCharge(int boundaryInd, double chargeDensity)
{
    PoissonBC bc;
    bc.type = QBC_Neumann;
    bc.normalGrad = chargeDensity / (2 * kEps0);
    poisson.SetBC(boundaryInd, bc);
}

void FullSolver::InitializeWallCharge()
{
    for (size_t fi = 0; fi < mesh->facePoints.size(); fi++)
        if (faceBCType[fi] == PBC_Absorbing && faceCollect[fi]) {
            int e = mesh->faceEntity[fi];
            if (!wallCharge.count(e)) {
                wallCharge[e] = 0;
                wallArea[e] = 0;
            }
            wallArea[e] += mesh->faceArea[fi];
        }
}

// _PrecomputeNormalTensors (solver.cpp:258-293); Full::Compress is a no-op (full.cpp:33-37).
void FullSolver::PrecomputeNormalTensors()
{
    const VGrid& g = *vg;
    size_t nf = mesh->facePoints.size();
    vNormal.assign(nf, Full());
    vNormalAbs.assign(nf, Full());
    for (size_t fi = 0; fi < nf; fi++) {
        const Vec3& nrm = mesh->faceNormal[fi];
        vNormal[fi] = Full(g.nTotal);
        vNormalAbs[fi] = Full(g.nTotal);
        for (int e = 0; e < g.nTotal; e++) {
            double x = nrm[0] * g.v[0][e] + nrm[1] * g.v[1][e] + nrm[2] * g.v[2][e];
            vNormal[fi].a[e] = x;
            vNormalAbs[fi].a[e] = std::fabs(x);
        }
    }
}

Full FullSolver::Flux(int t, int f, int bcType) const
{
    int fi = 4 * t + f;
    const Full& vN = vNormal[fi];
    const Full& vNabs = vNormalAbs[fi];
    const Full& A = sp->pdf[t];
    if (bcType == PBC_NonBoundary || bcType == PBC_Periodic) {
        const Full& B = sp->pdf[mesh->adj[fi]];
        return 0.5 * (vN * (B + A) - vNabs * (B - A));
    } else if (bcType == PBC_Absorbing) {
        return 0.5 * (vN * A + vNabs * A);
    } else if (bcType == PBC_Source) {
        const Full& B = sourcePDFs[faceSource[fi]];
        return 0.5 * (vN * (B + A) - vNabs * (B - A));
    } else if (bcType == PBC_Free) {
        return vN * A;
    }
    return Full(vg->nTotal);
}

Full FullSolver::PDFDerivative(int t, int ind) const
{
    const VGrid& g = *vg;
    Full der(g.nTotal);
    const std::vector<double>& f = sp->pdf[t].a;
    int i[3];
    for (i[0] = 0; i[0] < g.n[0]; i[0]++)
        for (i[1] = 0; i[1] < g.n[1]; i[1]++)
            for (i[2] = 0; i[2] < g.n[2]; i[2]++) {
                int ip[3] = {i[0], i[1], i[2]}, im[3] = {i[0], i[1], i[2]};
                if (i[ind] == 0) {
                    ip[ind] = 1;
                    im[ind] = g.n[ind] - 1;
                } else if (i[ind] == g.n[ind] - 1) {
                    ip[ind] = 0;
                    im[ind] = g.n[ind] - 2;
                } else {
                    ip[ind] += 1;
                    im[ind] -= 1;
                }
                der.a[g.idx(i[0], i[1], i[2])] =
                    (f[g.idx(ip[0], ip[1], ip[2])] - f[g.idx(im[0], im[1], im[2])]) / (2 * g.step[ind]);
            }
    return der;
}

void FullSolver::UpdatePDF()
{
    if (fused) {
        UpdatePDFFused();
        return;
    }
    if (vNormal.empty()) PrecomputeNormalTensors();
    int nT = mesh->nTets();
    std::vector<Full> rhs(nT, Full(vg->nTotal));

#pragma omp parallel for
    for (int t = 0; t < nT; t++) {
        for (int f = 0; f < 4; f++) {
            int fi = 4 * t + f;
            int bc = faceBCType[fi];
            Full flux = Flux(t, f, bc);
            rhs[t] = rhs[t] - (mesh->faceArea[fi] / mesh->tetVolume[t]) * flux;
            if (bc == PBC_Absorbing && faceCollect[fi]) {
                double particlesAbsorbed = timeStep * mesh->faceArea[fi] * flux.Sum() * vg->cellVolume;
#pragma omp critical
                wallCharge[mesh->faceEntity[fi]] += sp->charge * particlesAbsorbed;
            }
        }
    }

#pragma omp parallel for
    for (int t = 0; t < nT; t++) {
        for (int k = 0; k < 3; k++) {
            double electricField = field[t][k] + externalField[k];
            double forceComponent = (sp->charge / sp->mass) * electricField;
            rhs[t] = rhs[t] - forceComponent * PDFDerivative(t, k);
        }
    }

#pragma omp parallel for
    for (int t = 0; t < nT; t++) sp->pdf[t] = sp->pdf[t] + timeStep * rhs[t];
}

// Same arithmetic (same operation order per element, no FMA), one pass, no temporaries.
// Used for the "fused CPU" baseline figure and checked bit-identical to
// UpdatePDF() in tests.
void FullSolver::UpdatePDFFused()
{
    const VGrid& g = *vg;
    int nT = mesh->nTets();
    int N = g.nTotal, n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
    std::vector<std::vector<double>> next(nT);
    std::vector<double> wallAdd(mesh->facePoints.size(), 0.0);

#pragma omp parallel for
    for (int t = 0; t < nT; t++) {
        const double* A = sp->pdf[t].a.data();
        std::vector<double> rhs(N, 0.0);
        for (int f = 0; f < 4; f++) {
            int fi = 4 * t + f;
            int bc = faceBCType[fi];
            const Vec3& nrm = mesh->faceNormal[fi];
            double c = mesh->faceArea[fi] / mesh->tetVolume[t];
            const double* B = nullptr;
            if (bc == PBC_NonBoundary || bc == PBC_Periodic) B = sp->pdf[mesh->adj[fi]].a.data();
            if (bc == PBC_Source) B = sourcePDFs[faceSource[fi]].a.data();
            double fsum = 0;
            for (int e = 0; e < N; e++) {
                double vn = nrm[0] * g.v[0][e] + nrm[1] * g.v[1][e] + nrm[2] * g.v[2][e];
                double va = std::fabs(vn);
                double flux;
                if (B) flux = 0.5 * (vn * (B[e] + A[e]) - va * (B[e] - A[e]));
                else if (bc == PBC_Absorbing) flux = 0.5 * (vn * A[e] + va * A[e]);
                else flux = vn * A[e];
                fsum += flux;
                rhs[e] = rhs[e] - c * flux;
            }
            if (bc == PBC_Absorbing && faceCollect[fi])
                wallAdd[fi] = sp->charge * (timeStep * mesh->faceArea[fi] * fsum * g.cellVolume);
        }
        for (int k = 0; k < 3; k++) {
            double force = (sp->charge / sp->mass) * (field[t][k] + externalField[k]);
            int stride = k == 0 ? 1 : (k == 1 ? n0 : n0 * n1);
            int nk = g.n[k];
            double den = 2 * g.step[k];
            for (int i2 = 0; i2 < n2; i2++)
                for (int i1 = 0; i1 < n1; i1++)
                    for (int i0 = 0; i0 < n0; i0++) {
                        int e = i0 + n0 * (i1 + n1 * i2);
                        int ik = k == 0 ? i0 : (k == 1 ? i1 : i2);
                        int ep = ik == nk - 1 ? e - (nk - 1) * stride : e + stride;
                        int em = ik == 0 ? e + (nk - 1) * stride : e - stride;
                        rhs[e] = rhs[e] - force * ((A[ep] - A[em]) / den);
                    }
        }
        next[t].resize(N);
        for (int e = 0; e < N; e++) next[t][e] = A[e] + timeStep * rhs[e];
    }
    for (int t = 0; t < nT; t++) sp->pdf[t].a.swap(next[t]);
    for (size_t fi = 0; fi < wallAdd.size(); fi++)
        if (faceBCType[fi] == PBC_Absorbing && faceCollect[fi]) wallCharge[mesh->faceEntity[fi]] += wallAdd[fi];
}

void FullSolver::StepOnce()
{
    // solver.cpp:98-105
    rho = sp->Density();
    for (size_t i = 0; i < rho.size(); i++) {
        rho[i] *= sp->charge;
        if (!backgroundChargeDensity.empty()) rho[i] += backgroundChargeDensity[i];
    }
    // solver.cpp:108-110
    poisson.Solve(rho);
    phi = poisson.solution;
    field = poisson.ElectricField();
    // solver.cpp:115
    UpdatePDF();
    // solver.cpp:120-132
    for (auto& kv : wallCharge) SetFieldBCCharge(kv.first, kv.second / wallArea[kv.first]);
}

void MultiStepOnce(std::vector<FullSolver*>& solvers, const std::vector<int>& mult, int iteration)
{
    FullSolver* base = solvers[0];
    int nT = base->mesh->nTets();
    // multicomponent_solver.cpp:61-74
    std::vector<double> rho(nT, 0.0);
    for (auto* s : solvers) {
        std::vector<double> density = s->sp->Density();
        for (int i = 0; i < nT; i++) rho[i] += s->sp->charge * density[i];
    }
    if (!base->backgroundChargeDensity.empty())
        for (int i = 0; i < nT; i++) rho[i] += base->backgroundChargeDensity[i];
    // :77-83
    base->poisson.Solve(rho);
    for (auto* s : solvers) {
        s->rho = rho;
        s->phi = base->poisson.solution;
        s->field = base->poisson.ElectricField();
    }
    // :86-94
    for (size_t k = 0; k < solvers.size(); k++) {
        if (iteration % mult[k]) continue;
        solvers[k]->UpdatePDF();
    }
    // :99-126
    std::map<int, double> wall = base->wallCharge;
    for (auto* s : solvers) {
        if (s == base) continue;
        for (auto& kv : s->wallCharge) wall[kv.first] += kv.second;
    }
    for (auto& kv : wall) base->SetFieldBCCharge(kv.first, kv.second / base->wallArea[kv.first]);
}

}  // namespace oracle
