// Poisson solver of the public API (reference: src/poisson.h:15-103).  Assembly data, the
// Jacobi-PCG solve, the deferred non-orthogonal correction and the least-squares gradient all
// run on the device (csrc/poisson.cu); this class keeps the per-face boundary conditions and
// host copies of the last solution.  Both SparseSolverType values map to the device PCG run to
// Eigen's default tolerance (2.2e-16 relative residual): the reference's SparseLU default is a
// direct solve of the same system.
#pragma once
#include <array>
#include <memory>
#include <vector>

#include "constants.h"
#include "mesh.h"

namespace VlasovTucker {
namespace device {
struct MeshContext;
}

enum class SparseSolverType { SparseLU, ConjugateGradient };
enum class PoissonBCType { NonBoundary, Neumann, Dirichlet, Periodic };

struct PoissonBC {
    PoissonBCType type = PoissonBCType::NonBoundary;
    double value = 0;
    double normalGrad = 0;
};

class PoissonSolver {
public:
    PoissonSolver();
    PoissonSolver(const Mesh* mesh);

    void SetBC(int boundaryInd, const PoissonBC& bc);
    void SetSparseSolverType(SparseSolverType type);
    void Initialize();
    void Solve(std::vector<double> rho);

    const std::vector<double>& Potential() const;
    std::vector<Vector3d> ElectricField() const;

    // device-side entry used by the solvers: rho already assembled on the device
    void SolveOnDevice(bool download);
    int LastIterations() const;

private:
    void PushBCValues();

    const Mesh* _mesh;
    std::vector<PoissonBC> _faceBC;
    bool _initialized = false;
    bool _valuesDirty = false;
    SparseSolverType _type = SparseSolverType::SparseLU;
    std::shared_ptr<device::MeshContext> _dev;
    std::vector<double> _solution;
    std::vector<Vector3d> _field;
};
}  // namespace VlasovTucker
