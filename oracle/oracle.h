// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Eigen-free CPU restatement of the reference's per-time-step kinetic update
// (DmitriiGurev/VlasovTucker, /root/reference).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.  The
// product (vlasovtucker_b200/) never includes or links anything from oracle/.
//
// PARITY STATUS: "parity unpinned" for _UpdatePDF/_Flux/_PDFDerivative/Density/wall
// charge/MulticomponentSolver — the reference ships no golden vectors for them and
// cannot be compiled here (its vendored Eigen lacks Eigen/Core).  The
// oracle is pinned only where the reference's own tests give known answers: Full
// sums (test/test_tensors.cpp:10-18), Tucker add/scale/round invariance
// (test/tucker_test.cpp:173-203), Poisson analytic solutions
// (test/poisson_test.cpp:28-297) and fixture-derived mesh counts.
//
// Every function cites the reference file:line it restates.
#pragma once

#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace oracle {

typedef std::array<double, 3> Vec3;

// ---------------------------------------------------------------- mesh (src/mesh.cpp, src/primitives.cpp)
struct Mesh {
    // points[i] = i-th $Nodes line, already multiplied by the scale factor (mesh.cpp:113-127)
    std::vector<Vec3> points;
    // tets[t] = t-th type-4 element in file order, 0-based node ids (mesh.cpp:129-149)
    std::vector<std::array<int, 4>> tets;
    std::vector<Vec3> tetCentroid;
    std::vector<double> tetVolume;
    // faces[4t+j], vertex order of mesh.cpp:157-160
    std::vector<std::array<int, 3>> facePoints;
    std::vector<Vec3> faceNormal;
    std::vector<Vec3> faceCentroid;
    std::vector<double> faceArea;
    std::vector<int> faceEntity;    // -1 = internal (primitives.h:76)
    std::vector<uint8_t> faceBoundary;
    // adj[4t+j] = tet index across face j, -1 where the reference holds nullptr
    std::vector<int> adj;
    // entity tag -> face indices in triangle file order, duplicates kept (mesh.cpp:183-206)
    std::map<int, std::vector<int>> entityToFaces;
    std::map<int, std::vector<std::string>> entityToPhysGroups;
    std::vector<std::array<int, 2>> periodicPairs;

    int nTets() const { return (int)tets.size(); }
};

// Load an MSH 2.2 ASCII file and run Mesh::Reconstruct (mesh.cpp:13-29, 94-303).
Mesh LoadMesh(const std::string& file, const std::vector<std::array<int, 2>>& periodicPairs,
              double scale);
Mesh MeshFromArrays(const double* nodes, int nNodes, const int* tets, int nTets, const int* tris,
                    const int* triEntity, int nTris,
                    const std::vector<std::array<int, 2>>& periodicPairs, double scale);
double AverageCellSize(const Mesh& m);  // mesh.cpp:86-92

// ---------------------------------------------------------------- velocity grid (src/velocity_grid.cpp)
struct VGrid {
    std::array<int, 3> n;
    int nTotal;
    Vec3 minV, maxV;
    std::array<double, 3> step;
    double cellVolume;
    // v[j] as dense tensors, col-major (i0 fastest), velocity_grid.cpp:19-34
    std::array<std::vector<double>, 3> v;
    // d[j]: n_j x n_j central difference, zero outside the grid, row-major here (velocity_grid.cpp:36-52)
    std::array<std::vector<double>, 3> d;
    inline int idx(int i0, int i1, int i2) const { return i0 + n[0] * (i1 + n[1] * i2); }
};
VGrid MakeVGrid(const std::array<int, 3>& n, const Vec3& minV, const Vec3& maxV);

// ---------------------------------------------------------------- Full tensor (src/full.cpp)
// One materialised temporary per operator, exactly as full.cpp:69-101.
struct Full {
    std::vector<double> a;
    Full() {}
    explicit Full(size_t n) : a(n, 0.0) {}
    explicit Full(const std::vector<double>& v) : a(v) {}
    double Sum() const;  // full.cpp:29-31
};
Full operator+(const Full& x, const Full& y);
Full operator-(const Full& x, const Full& y);
Full operator*(const Full& x, const Full& y);
Full operator*(double d, const Full& x);

// ---------------------------------------------------------------- boundary-condition enums (src/solver.h:25, src/poisson.h:47)
enum ParticleBCType : int { PBC_NonBoundary = 0, PBC_Periodic = 1, PBC_Source = 2, PBC_Absorbing = 3, PBC_Free = 4 };
enum PoissonBCType : int { QBC_NonBoundary = 0, QBC_Neumann = 1, QBC_Dirichlet = 2, QBC_Periodic = 3 };

// ---------------------------------------------------------------- Poisson (src/poisson.cpp)
struct PoissonBC {
    int type = QBC_NonBoundary;
    double value = 0;
    double normalGrad = 0;
};

struct Poisson {
    const Mesh* mesh = nullptr;
    std::vector<PoissonBC> faceBC;
    bool solutionIsUnique = false;
    // CSR of _system (poisson.cpp:107-124); duplicates from setFromTriplets summed
    std::vector<int> rowPtr, colInd;
    std::vector<double> val;
    std::vector<double> solution;
    std::vector<Vec3> gradient;
    std::vector<double> guess;
    int lastIterations = 0;
    double lastError = 0;
    long totalIterations = 0;

    explicit Poisson(const Mesh* m);                  // poisson.cpp:67-83
    void SetBC(int boundaryInd, const PoissonBC& bc); // poisson.cpp:85-92
    void Initialize();                                // poisson.cpp:107-124
    void Solve(const std::vector<double>& rho);       // poisson.cpp:179-213
    std::vector<Vec3> ElectricField() const;          // poisson.cpp:220-229

    // helpers
    Vec3 PeriodicShiftedDistance(int t, int j) const; // poisson.cpp:152-162
    std::vector<double> SolveSystem(const std::vector<double>& rhs, bool useGuess);
    Vec3 TetLSG(int t) const;                         // poisson.cpp:361-442
    Vec3 WeightedGradient(int t, int f) const;        // poisson.cpp:276-304
};

// ---------------------------------------------------------------- species + solver (src/particle_data.cpp, src/solver.cpp)
struct Species {
    const Mesh* mesh;
    const VGrid* vg;
    double mass = 1, charge = 1;
    std::vector<Full> pdf;

    void SetMaxwell(const std::vector<double>& physDensity, double temperature,
                    const Vec3& mostProbableV);        // particle_data.cpp:23-90
    std::vector<double> Density() const;               // particle_data.cpp:93-102
    std::vector<Vec3> Velocity() const;                // particle_data.cpp:105-125
};

struct FullSolver {
    const Mesh* mesh;
    const VGrid* vg;
    Species* sp;
    Poisson poisson;
    double timeStep = 0;
    std::vector<double> backgroundChargeDensity;
    Vec3 externalField = {0, 0, 0};
    std::vector<double> rho, phi;
    std::vector<Vec3> field;
    std::vector<int> faceBCType;        // per face, ParticleBCType
    std::vector<uint8_t> faceCollect;   // per face
    std::vector<int> faceSource;        // per face, index into sourcePDFs or -1
    std::vector<Full> sourcePDFs;
    std::map<int, double> wallCharge, wallArea;
    bool fused = false;                 // false = reference-faithful temporaries
    std::vector<Full> vNormal, vNormalAbs;   // per face (solver.h:90-91), built lazily

    FullSolver(const Mesh* m, const VGrid* g, Species* s);          // solver.cpp:16-41
    void SetParticleBC(int boundaryInd, int type, bool collect, int sourceId);  // solver.cpp:62-71
    void SetFieldBCPotential(int boundaryInd, double potential);    // solver.cpp:45-59
    void SetFieldBCCharge(int boundaryInd, double chargeDensity);   // solver.cpp:45-59
    void InitializeWallCharge();                                    // solver.cpp:296-311
    void PrecomputeNormalTensors();                                 // solver.cpp:258-293
    Full Flux(int t, int f, int bcType) const;                      // solver.cpp:314-346
    Full PDFDerivative(int t, int k) const;                         // solver.cpp:363-404
    void UpdatePDF();                                               // solver.cpp:141-212
    void UpdatePDFFused();                                          // same arithmetic, no temporaries
    void StepOnce();                                                // solver.cpp:91-133 loop body
};

// multicomponent_solver.cpp:27-135, one iteration of the loop body
void MultiStepOnce(std::vector<FullSolver*>& solvers, const std::vector<int>& multipliers,
                   int iteration);

}  // namespace oracle
