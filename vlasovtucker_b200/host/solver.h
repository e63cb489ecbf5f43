// Single-species Vlasov-Poisson time loop of the public API (reference: src/solver.h:16-96).
// The loop body — Density -> rho -> Poisson -> _UpdatePDF -> wall charge — runs on the GPU
// through include/vt_b200.h; this class keeps the reference's public names and call sequence.
#pragma once
#include <limits.h>

#include <map>
#include <unordered_map>
#include <vector>

#include "log.h"
#include "mesh.h"
#include "particle_data.h"
#include "poisson.h"
#include "typedefs.h"

namespace VlasovTucker {
// ---- boundary conditions as the drivers set them
enum class ParticleBCType { NonBoundary, Periodic, Source, Absorbing, Free };   // numeric values = VT_PBC_*
enum class FieldBCType { ConstantPotential, ChargedPlane };

template <typename TensorType>
struct ParticleBC {
    ParticleBCType type = ParticleBCType::NonBoundary;
    bool collectCharge = false;   // Absorbing: accumulate the absorbed charge of the entity
    TensorType sourcePDF;         // Source: takes the neighbour's place in the flux
};

struct FieldBC {
    FieldBCType type;
    double potential;       // ConstantPotential -> Dirichlet value
    double chargeDensity;   // ChargedPlane      -> Neumann value sigma / (2 eps0)
};

template <typename TensorType>
class Solver {
    using Data = ParticleData<TensorType>;

public:
    // run parameters (public data, assigned by the driver before Solve)
    int nIterations = 0;
    double timeStep = 0;
    int writeStep = INT_MAX;
    Vector3d externalField = {0, 0, 0};
    std::vector<double> backgroundChargeDensity;

    Solver(const Mesh* mesh, const VelocityGrid* velocityGrid, Data* particleData);

    void SetParticleBC(int boundaryInd, const ParticleBC<TensorType>& bc);
    void SetFieldBC(int boundaryInd, const FieldBC& bc);
    void SetSparseSolverType(SparseSolverType type);
    void Solve();

private:
    // device side of one iteration
    void _PushParticleBC();     // upload the per-face BC tables when they changed
    void _UpdatePDF();          // vt_step_full / vt_step_tucker
    void _PullWallCharge();
    // host side
    void _InitializeWallCharge();
    void _WriteResults(int iteration);

    const Mesh* _mesh;
    const VelocityGrid* _vGrid;
    Data* _pData;
    PoissonSolver _poissonSolver;
    Log _log;

    bool _bcDirty = true;
    std::vector<ParticleBC<TensorType>> _faceParticleBC;   // one per face of the mesh
    std::unordered_map<int, double> _wallCharge;           // entity -> absorbed charge so far
    std::unordered_map<int, double> _wallArea;             // entity -> area of its collecting faces

    // last downloaded fields (for the VTK dumps)
    std::vector<double> _rho, _phi;
    std::vector<Vector3d> _field;

    template <typename TensorTypeM>
    friend class MulticomponentSolver;
};
}  // namespace VlasovTucker
