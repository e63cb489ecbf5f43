#include "solver.h"

#include <stdexcept>

#include "device.h"
#include "full.h"
#include "timer.h"
#include "tucker.h"
#include "vtk.h"

namespace VlasovTucker {

template <typename T>
Solver<T>::Solver(const Mesh* mesh, const VelocityGrid* vGrid, ParticleData<T>* pData)
    : _mesh(mesh), _vGrid(vGrid), _pData(pData), _poissonSolver(mesh)
{
    _log = Log(LogLevel::Console);
    // v.n and |v.n| are recomputed inside the step kernel; the reference stores 8 tensors per tet here
    // (_PrecomputeNormalTensors, solver.cpp:258-293) and logs their compression ratio — there is no
    // such tensor to report on, so nothing is logged
    _faceParticleBC.resize(mesh->faces.size());
    ParticleBC<T> periodic;
    periodic.type = ParticleBCType::Periodic;
    for (const auto& pair : mesh->PeriodicBoundaries())
        for (int mark : pair) SetParticleBC(mark, periodic);
}

template <typename T>
void Solver<T>::SetFieldBC(int boundaryInd, const FieldBC& bc)
{
    PoissonBC p;
    if (bc.type == FieldBCType::ConstantPotential) {
        p.type = PoissonBCType::Dirichlet;
        p.value = bc.potential;
    } else {
        p.type = PoissonBCType::Neumann;
        p.normalGrad = bc.chargeDensity / (2 * epsilon0);
    }
    _poissonSolver.SetBC(boundaryInd, p);
}

template <typename T>
void Solver<T>::SetParticleBC(int boundaryInd, const ParticleBC<T>& bc)
{
    for (size_t i = 0; i < _mesh->faces.size(); i++)
        if (_mesh->faces[i]->entity == boundaryInd) _faceParticleBC[i] = bc;
    _bcDirty = true;
}

template <typename T>
void Solver<T>::SetSparseSolverType(SparseSolverType type)
{
    _poissonSolver.SetSparseSolverType(type);
}

template <typename T>
void Solver<T>::_InitializeWallCharge()
{
    for (Face* f : _mesh->faces) {
        const auto& bc = _faceParticleBC[f->index];
        if (bc.type == ParticleBCType::Absorbing && bc.collectCharge) {
            if (!_wallCharge.count(f->entity)) {
                _wallCharge[f->entity] = 0;
                _wallArea[f->entity] = 0;
            }
            _wallArea[f->entity] += f->area;
        }
    }
}

template <typename T>
void Solver<T>::_PushParticleBC()
{
    if (!_bcDirty) return;
    auto dev = _pData->DeviceContext();
    const size_t nf = _faceParticleBC.size();
    std::vector<uint8_t> type(nf), collect(nf);
    std::vector<int32_t> source(nf, -1);
    std::vector<double> sources;
    const int N = _vGrid->nCellsTotal;
    int nSrc = 0;
    for (size_t i = 0; i < nf; i++) {
        const auto& bc = _faceParticleBC[i];
        type[i] = (uint8_t)bc.type;   // enum order matches VT_PBC_*
        collect[i] = bc.collectCharge ? 1 : 0;
        if (bc.type == ParticleBCType::Source) {
            const Tensor3d t = bc.sourcePDF.Reconstructed();
            if (t.size() != N) throw std::invalid_argument("Different shapes in source PDF");
            sources.insert(sources.end(), t.data(), t.data() + N);
            source[i] = nSrc++;
        }
    }
    device::Check(vt_species_set_source_pdfs(dev->ctx, _pData->DeviceSpecies(), nSrc, sources.data()));
    device::Check(vt_species_set_face_bc(dev->ctx, _pData->DeviceSpecies(), type.data(), collect.data(), source.data()));
    _bcDirty = false;
}

template <typename T>
void Solver<T>::_PullWallCharge()
{
    auto dev = _pData->DeviceContext();
    for (auto& kv : _wallCharge)
        device::Check(vt_wall_charge_get(dev->ctx, _pData->DeviceSpecies(), kv.first, &kv.second));
}

template <>
void Solver<Full>::_UpdatePDF()
{
    _log << Indent(2) << "Compute the right-hand side\n";
    _log << Indent(3) << "Boltzmann part\n";
    _log << Indent(3) << "Vlasov part\n";
    _log << Indent(2) << "Time integration\n";
    _pData->PushParams();
    _PushParticleBC();
    auto dev = _pData->DeviceContext();
    const double ext[3] = {externalField[0], externalField[1], externalField[2]};
    device::Check(vt_step_full(dev->ctx, _pData->DeviceSpecies(), timeStep, ext));
    if (!_wallCharge.empty()) _PullWallCharge();
}

template <>
void Solver<Tucker>::_UpdatePDF()
{
    _log << Indent(2) << "Compute the right-hand side\n";
    _log << Indent(3) << "Boltzmann part\n";
    _log << Indent(3) << "Vlasov part\n";
    _log << Indent(2) << "Time integration\n";
    _pData->PushParams();
    _PushParticleBC();
    auto dev = _pData->DeviceContext();
    const double ext[3] = {externalField[0], externalField[1], externalField[2]};
    // flux per face with rounding after each face, acceleration term, rounding, Euler update,
    // rounding — all on the device (csrc/tucker.cu)
    device::Check(vt_step_tucker(dev->ctx, _pData->DeviceSpecies(), timeStep, ext));
    if (!_wallCharge.empty()) _PullWallCharge();
}

template <typename T>
void Solver<T>::Solve()
{
    _pData->PushParams();
    _PushParticleBC();   // also restarts the device wall-charge accumulators
    _InitializeWallCharge();
    _log << "Initialize the Poisson solver\n";
    Timer timer;
    _poissonSolver.Initialize();
    timer.PrintSectionTime("Poisson solver initialization");
    auto dev = _pData->DeviceContext();
    const int sp = _pData->DeviceSpecies();

    _log << "Start the main loop\n";
    for (int iteration = 0; iteration < nIterations; iteration++) {
        _log << "\n" << "Iteration #" << iteration << "\n";
        _log << "Time: " << iteration * timeStep << "\n";
        _log << Indent(1) << "Compute the electric field\n";
        // rho = charge * Density() + background, assembled on the device
        _pData->PushParams();
        device::Check(vt_charge_density(dev->ctx, &sp, 1, backgroundChargeDensity.empty() ? nullptr : backgroundChargeDensity.data()));
        const bool write = iteration % writeStep == 0;
        _poissonSolver.SolveOnDevice(write);
        timer.PrintSectionTime(Indent(1) + "Done");

        _log << Indent(1) << "Update the PDF\n";
        _UpdatePDF();
        timer.PrintSectionTime(Indent(1) + "Done");

        _log << Indent(1) << "Update the boundary conditions\n";
        for (const auto& kv : _wallCharge) {
            FieldBC bc;
            bc.type = FieldBCType::ChargedPlane;
            bc.chargeDensity = kv.second / _wallArea[kv.first];
            SetFieldBC(kv.first, bc);
        }
        if (write) {
            const size_t n = _mesh->tets.size();
            _rho.resize(n);
            _phi.resize(n);
            _field.resize(n);
            device::Check(vt_field_get(dev->ctx, _rho.data(), nullptr, nullptr));
            _phi = _poissonSolver.Potential();
            _field = _poissonSolver.ElectricField();
            _pData->SyncFromDevice();   // Tucker: the host mirror feeds the distribution dump
            _WriteResults(iteration);
        }
    }
    _pData->SyncFromDevice();
}

template <typename T>
void Solver<T>::_WriteResults(int iteration)
{
    _log << "Particle species: " << _pData->species << "\n";
    double averageTensorSize = 0;
    for (size_t i = 0; i < _mesh->tets.size(); i++) averageTensorSize += _pData->pdf[i].Size();
    averageTensorSize /= (double)_mesh->tets.size();
    _log << Indent(1) << "Average PDF size = " << averageTensorSize << "\n";
    _log << Indent(1) << "(Uncompressed: " << _vGrid->nCellsTotal << ")\n";

    const std::string prefix = "solution/";
    const std::string postfix = "_" + std::to_string(iteration / writeStep);
    WriteCellScalarDataVTK(prefix + "density/density_" + _pData->species + postfix, *_mesh, _pData->Density());
    WriteCellVectorDataVTK(prefix + "velocity/velocity_" + _pData->species + postfix, *_mesh, _pData->Velocity());
    const int tetInd = (int)_mesh->tets.size() / 2;
    WriteDistributionVTK(prefix + "distribution/distribution_" + _pData->species + postfix, *_vGrid,
                         _pData->pdf[tetInd].Reconstructed());
    // file names keep the reference's double underscore (solver.cpp:245-253)
    WriteCellScalarDataVTK(prefix + "charge_density/charge_density_" + postfix, *_mesh, _rho);
    WriteCellScalarDataVTK(prefix + "phi/phi_" + postfix, *_mesh, _phi);
    WriteCellVectorDataVTK(prefix + "field/e_" + postfix, *_mesh, _field);
}

template class Solver<Full>;
template class Solver<Tucker>;

}  // namespace VlasovTucker
