// Single-species Vlasov-Poisson time loop of the public API (reference: src/solver.h:16-96).
// The loop body — Density -> rho -> Poisson -> _UpdatePDF -> wall charge — runs on the GPU
// through include/vt_b200.h; this class keeps the reference's members and call sequence.
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
enum class FieldBCType { ConstantPotential, ChargedPlane };

struct FieldBC {
    FieldBCType type;
    double potential;
    double chargeDensity;
};

enum class ParticleBCType { NonBoundary, Periodic, Source, Absorbing, Free };

template <typename TensorType>
struct ParticleBC {
    ParticleBCType type = ParticleBCType::NonBoundary;
    TensorType sourcePDF;
    bool collectCharge = false;
};

template <typename TensorType>
class Solver {
public:
    Solver(const Mesh* mesh, const VelocityGrid* velocityGrid, ParticleData<TensorType>* particleData);

    void SetFieldBC(int boundaryInd, const FieldBC& bc);
    void SetParticleBC(int boundaryInd, const ParticleBC<TensorType>& bc);
    void SetSparseSolverType(SparseSolverType type);
    void Solve();

private:
    void _InitializeWallCharge();
    void _UpdatePDF();
    void _WriteResults(int iteration);
    void _PushParticleBC();     // upload the per-face BC tables when they changed
    void _PullWallCharge();

public:
    double timeStep = 0;
    int nIterations = 0;
    int writeStep = INT_MAX;
    std::vector<double> backgroundChargeDensity;
    Vector3d externalField = {0, 0, 0};

private:
    const Mesh* _mesh;
    const VelocityGrid* _vGrid;
    ParticleData<TensorType>* _pData;
    PoissonSolver _poissonSolver;

    std::vector<double> _rho;
    std::vector<double> _phi;
    std::vector<Vector3d> _field;

    std::vector<ParticleBC<TensorType>> _faceParticleBC;
    bool _bcDirty = true;
    std::unordered_map<int, double> _wallCharge;
    std::unordered_map<int, double> _wallArea;
    Log _log;

public:
    template <typename TensorTypeM>
    friend class MulticomponentSolver;
};
}  // namespace VlasovTucker
