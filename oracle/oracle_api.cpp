// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).  Flat C interface for ctypes.
#include "oracle.h"

#include <cstring>
#include <memory>
#include <string>

using namespace oracle;

namespace {
thread_local std::string g_err;
template <class F>
int guard(F f)
{
    try {
        f();
        return 0;
    } catch (std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
std::vector<std::array<int, 2>> pairsOf(const int* pairs, int n)
{
    std::vector<std::array<int, 2>> p;
    for (int i = 0; i < n; i++) p.push_back({pairs[2 * i], pairs[2 * i + 1]});
    return p;
}
struct SimState {
    Mesh* mesh;
    std::vector<std::unique_ptr<VGrid>> grids;
    std::vector<std::unique_ptr<Species>> species;
    std::vector<std::unique_ptr<FullSolver>> solvers;
    std::vector<FullSolver*> order;
    std::vector<int> mult;
};
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// ---- mesh
void* orc_mesh_load(const char* file, const int* pairs, int nPairs, double scale)
{
    Mesh* m = nullptr;
    if (guard([&] { m = new Mesh(LoadMesh(file, pairsOf(pairs, nPairs), scale)); })) return nullptr;
    return m;
}
void* orc_mesh_from_arrays(const double* nodes, int nNodes, const int* tets, int nTets,
                           const int* tris, const int* triEntity, int nTris, const int* pairs,
                           int nPairs, double scale)
{
    Mesh* m = nullptr;
    if (guard([&] {
            m = new Mesh(MeshFromArrays(nodes, nNodes, tets, nTets, tris, triEntity, nTris,
                                        pairsOf(pairs, nPairs), scale));
        }))
        return nullptr;
    return m;
}
void orc_mesh_free(void* h) { delete (Mesh*)h; }
int orc_mesh_ntets(void* h) { return ((Mesh*)h)->nTets(); }
int orc_mesh_npoints(void* h) { return (int)((Mesh*)h)->points.size(); }
double orc_mesh_average_cell_size(void* h) { return AverageCellSize(*(Mesh*)h); }
void orc_mesh_get(void* h, double* points, int* tets, double* tetCentroid, double* tetVolume,
                  int* facePoints, double* faceNormal, double* faceCentroid, double* faceArea,
                  int* faceEntity, unsigned char* faceBoundary, int* adj)
{
    Mesh& m = *(Mesh*)h;
    size_t nt = m.tets.size(), nf = 4 * nt;
    if (points) memcpy(points, m.points.data(), m.points.size() * 24);
    if (tets) memcpy(tets, m.tets.data(), nt * 16);
    if (tetCentroid) memcpy(tetCentroid, m.tetCentroid.data(), nt * 24);
    if (tetVolume) memcpy(tetVolume, m.tetVolume.data(), nt * 8);
    if (facePoints) memcpy(facePoints, m.facePoints.data(), nf * 12);
    if (faceNormal) memcpy(faceNormal, m.faceNormal.data(), nf * 24);
    if (faceCentroid) memcpy(faceCentroid, m.faceCentroid.data(), nf * 24);
    if (faceArea) memcpy(faceArea, m.faceArea.data(), nf * 8);
    if (faceEntity) memcpy(faceEntity, m.faceEntity.data(), nf * 4);
    if (faceBoundary) memcpy(faceBoundary, m.faceBoundary.data(), nf);
    if (adj) memcpy(adj, m.adj.data(), nf * 4);
}
int orc_mesh_entity_faces(void* h, int entity, int* out, int cap)
{
    Mesh& m = *(Mesh*)h;
    auto it = m.entityToFaces.find(entity);
    if (it == m.entityToFaces.end()) return -1;
    int n = (int)it->second.size();
    for (int i = 0; i < n && i < cap; i++) out[i] = it->second[i];
    return n;
}
int orc_mesh_labels(void* h, char* buf, int cap)
{
    Mesh& m = *(Mesh*)h;
    std::string s;
    for (auto& kv : m.entityToPhysGroups) {
        s += std::to_string(kv.first) + ":";
        for (size_t i = 0; i < kv.second.size(); i++) s += (i ? "|" : "") + kv.second[i];
        s += "\n";
    }
    strncpy(buf, s.c_str(), cap - 1);
    buf[cap - 1] = 0;
    return (int)s.size();
}

// ---- stand-alone Poisson (test/poisson_test.cpp usage)
void* orc_poisson_create(void* mesh)
{
    return new Poisson((Mesh*)mesh);
}
void orc_poisson_free(void* p) { delete (Poisson*)p; }
int orc_poisson_set_bc(void* p, int entity, int type, double value, double normalGrad)
{
    return guard([&] {
        PoissonBC bc;
        bc.type = type;
        bc.value = value;
        bc.normalGrad = normalGrad;
        ((Poisson*)p)->SetBC(entity, bc);
    });
}
int orc_poisson_initialize(void* p) { return guard([&] { ((Poisson*)p)->Initialize(); }); }
int orc_poisson_nnz(void* p) { return (int)((Poisson*)p)->val.size(); }
void orc_poisson_csr(void* p, int* rowPtr, int* colInd, double* val)
{
    Poisson& q = *(Poisson*)p;
    memcpy(rowPtr, q.rowPtr.data(), q.rowPtr.size() * 4);
    memcpy(colInd, q.colInd.data(), q.colInd.size() * 4);
    memcpy(val, q.val.data(), q.val.size() * 8);
}
int orc_poisson_solve(void* p, const double* rho, double* phi, double* E)
{
    Poisson& q = *(Poisson*)p;
    int n = q.mesh->nTets();
    return guard([&] {
        q.Solve(std::vector<double>(rho, rho + n));
        if (phi) memcpy(phi, q.solution.data(), n * 8);
        if (E) {
            auto f = q.ElectricField();
            memcpy(E, f.data(), n * 24);
        }
    });
}
// the linear solve alone (SparseSolver::Solve, poisson.cpp:39-53) on a caller-built RHS
int orc_poisson_solve_system(void* p, const double* rhs, const double* guess, double* x)
{
    Poisson& q = *(Poisson*)p;
    int n = q.mesh->nTets();
    return guard([&] {
        if (guess) q.guess.assign(guess, guess + n);
        auto r = q.SolveSystem(std::vector<double>(rhs, rhs + n), guess != nullptr);
        memcpy(x, r.data(), n * 8);
    });
}
int orc_poisson_last_iterations(void* p) { return ((Poisson*)p)->lastIterations; }
double orc_poisson_last_error(void* p) { return ((Poisson*)p)->lastError; }

// ---- simulation (Solver<Full> / MulticomponentSolver<Full>)
void* orc_sim_create(void* mesh)
{
    SimState* s = new SimState();
    s->mesh = (Mesh*)mesh;
    return s;
}
void orc_sim_free(void* h) { delete (SimState*)h; }

// Adds VelocityGrid + ParticleData + Solver for one species; first added is the base solver.
int orc_sim_add_species(void* h, const int* n, const double* minV, const double* maxV, double mass,
                        double charge, int stepMultiplier)
{
    SimState& s = *(SimState*)h;
    s.grids.emplace_back(new VGrid(MakeVGrid({n[0], n[1], n[2]}, {minV[0], minV[1], minV[2]},
                                             {maxV[0], maxV[1], maxV[2]})));
    Species* sp = new Species();
    sp->mesh = s.mesh;
    sp->vg = s.grids.back().get();
    sp->mass = mass;
    sp->charge = charge;
    s.species.emplace_back(sp);
    s.solvers.emplace_back(new FullSolver(s.mesh, sp->vg, sp));
    s.order.push_back(s.solvers.back().get());
    s.mult.push_back(stepMultiplier);
    return (int)s.solvers.size() - 1;
}
int orc_sim_set_maxwell(void* h, int sp, const double* physDensity, double temperature,
                        const double* mpv)
{
    SimState& s = *(SimState*)h;
    return guard([&] {
        int n = s.mesh->nTets();
        s.species[sp]->SetMaxwell(std::vector<double>(physDensity, physDensity + n), temperature,
                                  {mpv[0], mpv[1], mpv[2]});
    });
}
void orc_sim_set_pdf(void* h, int sp, const double* f)
{
    SimState& s = *(SimState*)h;
    int n = s.mesh->nTets(), N = s.grids[sp]->nTotal;
    s.species[sp]->pdf.assign(n, Full((size_t)N));
    for (int t = 0; t < n; t++) memcpy(s.species[sp]->pdf[t].a.data(), f + (size_t)t * N, (size_t)N * 8);
}
void orc_sim_get_pdf(void* h, int sp, double* f)
{
    SimState& s = *(SimState*)h;
    int n = s.mesh->nTets(), N = s.grids[sp]->nTotal;
    for (int t = 0; t < n; t++) memcpy(f + (size_t)t * N, s.species[sp]->pdf[t].a.data(), (size_t)N * 8);
}
void orc_sim_set_params(void* h, int sp, double timeStep, const double* ext, const double* background,
                        int fused)
{
    SimState& s = *(SimState*)h;
    FullSolver& so = *s.solvers[sp];
    so.timeStep = timeStep;
    if (ext) so.externalField = {ext[0], ext[1], ext[2]};
    if (background) so.backgroundChargeDensity.assign(background, background + s.mesh->nTets());
    so.fused = fused != 0;
}
int orc_sim_set_particle_bc(void* h, int sp, int entity, int type, int collect, const double* sourcePDF)
{
    SimState& s = *(SimState*)h;
    FullSolver& so = *s.solvers[sp];
    int src = -1;
    if (sourcePDF) {
        int N = s.grids[sp]->nTotal;
        so.sourcePDFs.push_back(Full(std::vector<double>(sourcePDF, sourcePDF + N)));
        src = (int)so.sourcePDFs.size() - 1;
    }
    so.SetParticleBC(entity, type, collect != 0, src);
    return 0;
}
int orc_sim_set_field_bc(void* h, int sp, int entity, int isPotential, double value)
{
    SimState& s = *(SimState*)h;
    return guard([&] {
        if (isPotential) s.solvers[sp]->SetFieldBCPotential(entity, value);
        else s.solvers[sp]->SetFieldBCCharge(entity, value);
    });
}
// Solver::Solve prologue (solver.cpp:82-88) / MulticomponentSolver::Solve prologue (:38-51)
int orc_sim_begin(void* h)
{
    SimState& s = *(SimState*)h;
    return guard([&] {
        for (auto& so : s.solvers) so->InitializeWallCharge();
        s.solvers[0]->poisson.Initialize();
    });
}
// _InitializeWallCharge alone (solver.cpp:296-311), for kernel-level tests without field BCs
void orc_sim_init_wall(void* h)
{
    SimState& s = *(SimState*)h;
    for (auto& so : s.solvers) so->InitializeWallCharge();
}
// One loop iteration; single-species uses Solver::Solve's body, otherwise the multicomponent body.
int orc_sim_step(void* h, int iteration)
{
    SimState& s = *(SimState*)h;
    return guard([&] {
        if (s.solvers.size() == 1) s.solvers[0]->StepOnce();
        else MultiStepOnce(s.order, s.mult, iteration);
    });
}
// _UpdatePDF alone with a caller-supplied field (3 doubles per tet)
int orc_sim_update_pdf(void* h, int sp, const double* E)
{
    SimState& s = *(SimState*)h;
    return guard([&] {
        FullSolver& so = *s.solvers[sp];
        int n = s.mesh->nTets();
        so.field.resize(n);
        memcpy(so.field.data(), E, (size_t)n * 24);
        so.UpdatePDF();
    });
}
void orc_sim_get_fields(void* h, int sp, double* rho, double* phi, double* E)
{
    SimState& s = *(SimState*)h;
    FullSolver& so = *s.solvers[sp];
    int n = s.mesh->nTets();
    if (rho && (int)so.rho.size() == n) memcpy(rho, so.rho.data(), n * 8);
    if (phi && (int)so.phi.size() == n) memcpy(phi, so.phi.data(), n * 8);
    if (E && (int)so.field.size() == n) memcpy(E, so.field.data(), n * 24);
}
void orc_sim_density(void* h, int sp, double* out)
{
    SimState& s = *(SimState*)h;
    auto d = s.species[sp]->Density();
    memcpy(out, d.data(), d.size() * 8);
}
void orc_sim_velocity(void* h, int sp, double* out)
{
    SimState& s = *(SimState*)h;
    auto d = s.species[sp]->Velocity();
    memcpy(out, d.data(), d.size() * 24);
}
double orc_sim_wall_charge(void* h, int sp, int entity)
{
    SimState& s = *(SimState*)h;
    auto& w = s.solvers[sp]->wallCharge;
    auto it = w.find(entity);
    return it == w.end() ? 0.0 : it->second;
}
double orc_sim_wall_area(void* h, int sp, int entity)
{
    SimState& s = *(SimState*)h;
    auto& w = s.solvers[sp]->wallArea;
    auto it = w.find(entity);
    return it == w.end() ? 0.0 : it->second;
}
int orc_sim_poisson_iterations(void* h) { return ((SimState*)h)->solvers[0]->poisson.lastIterations; }

}  // extern "C"
