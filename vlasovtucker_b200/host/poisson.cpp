#include "poisson.h"

#include <iostream>

#include "device.h"
#include "log.h"

namespace VlasovTucker {

PoissonSolver::PoissonSolver() : _mesh(nullptr) {}

PoissonSolver::PoissonSolver(const Mesh* mesh) : _mesh(mesh)
{
    _faceBC.assign(mesh->faces.size(), PoissonBC());
    PoissonBC periodic;
    periodic.type = PoissonBCType::Periodic;
    for (const auto& pair : mesh->PeriodicBoundaries())
        for (int mark : pair) SetBC(mark, periodic);
}

void PoissonSolver::SetBC(int boundaryInd, const PoissonBC& bc)
{
    // .at(): an unknown entity is std::out_of_range, as in the reference (poisson.cpp:90)
    for (Face* f : _mesh->EntityToFaces().at(boundaryInd)) {
        if (_initialized && _faceBC[f->index].type != bc.type)
            throw std::runtime_error("PoissonSolver: a boundary-condition type cannot change after Initialize()");
        _faceBC[f->index] = bc;
    }
    _valuesDirty = true;
}

void PoissonSolver::SetSparseSolverType(SparseSolverType type) { _type = type; }

void PoissonSolver::Initialize()
{
    _dev = device::ContextOf(_mesh);
    const size_t nf = _faceBC.size();
    std::vector<uint8_t> type(nf);
    std::vector<double> value(nf), grad(nf);
    for (size_t i = 0; i < nf; i++) {
        type[i] = (uint8_t)_faceBC[i].type;   // enum order matches VT_QBC_*
        value[i] = _faceBC[i].value;
        grad[i] = _faceBC[i].normalGrad;
    }
    const FlatMesh& fm = _mesh->Flat();
    device::Check(vt_poisson_setup(_dev->ctx, fm.tetCentroid.data(), fm.faceCentroid.data(), type.data(), value.data(),
                                   grad.data()));
    _dev->poissonOwner = this;
    _initialized = true;
    _valuesDirty = false;
}

void PoissonSolver::PushBCValues()
{
    if (!_valuesDirty) return;
    const size_t nf = _faceBC.size();
    std::vector<double> value(nf), grad(nf);
    for (size_t i = 0; i < nf; i++) {
        value[i] = _faceBC[i].value;
        grad[i] = _faceBC[i].normalGrad;
    }
    device::Check(vt_poisson_update_bc_values(_dev->ctx, value.data(), grad.data()));
    _valuesDirty = false;
}

void PoissonSolver::Solve(std::vector<double> rho)
{
    if (!_initialized) throw std::runtime_error("PoissonSolver::Initialize must be called before Solve");
    if (_dev->poissonOwner != this) Initialize();   // another solver on the same mesh re-assembled
    PushBCValues();
    const size_t n = _mesh->tets.size();
    _solution.resize(n);
    _field.resize(n);
    device::Check(vt_poisson_solve(_dev->ctx, rho.data(), _solution.data(), &_field[0][0]));
}

void PoissonSolver::SolveOnDevice(bool download)
{
    if (!_initialized) throw std::runtime_error("PoissonSolver::Initialize must be called before Solve");
    if (_dev->poissonOwner != this) Initialize();
    PushBCValues();
    const size_t n = _mesh->tets.size();
    if (download) {
        _solution.resize(n);
        _field.resize(n);
        device::Check(vt_poisson_solve(_dev->ctx, nullptr, _solution.data(), &_field[0][0]));
    } else {
        device::Check(vt_poisson_solve(_dev->ctx, nullptr, nullptr, nullptr));
    }
}

int PoissonSolver::LastIterations() const
{
    int it = 0;
    double res = 0;
    if (_dev) vt_poisson_stats(_dev->ctx, &it, &res);
    return it;
}

const std::vector<double>& PoissonSolver::Potential() const { return _solution; }
std::vector<Vector3d> PoissonSolver::ElectricField() const { return _field; }

}  // namespace VlasovTucker
