#include "multicomponent_solver.h"

#include <stdexcept>
#include <unordered_map>

#include "device.h"
#include "full.h"
#include "timer.h"
#include "tucker.h"

namespace VlasovTucker {

template <typename T>
MulticomponentSolver<T>::MulticomponentSolver(Solver<T>* base) : _log(LogLevel::Console), _solvers({base})
{
    _log = Log(LogLevel::Console);
}

template <typename T>
void MulticomponentSolver<T>::AddSolver(Solver<T>* solver)
{
    _solvers.push_back(solver);
}

// One body for both tensor formats: everything it calls (vt_charge_density, the device Poisson solve,
// Solver<T>::_UpdatePDF, the wall-charge pull) is format-agnostic (multicomponent_solver.cpp:27-135).
template <typename T>
void MulticomponentSolver<T>::Solve()
{
    Solver<T>* base = _solvers[0];
    for (auto* s : _solvers)
        if (!stepMultipliers.count(s)) stepMultipliers[s] = 1;
    for (auto* s : _solvers) {
        s->timeStep = timeStep * stepMultipliers[s];
        s->writeStep = writeStep;
    }
    for (auto* s : _solvers) {
        if (s->_mesh != base->_mesh) throw std::invalid_argument("All species must share one mesh");
        s->_pData->PushParams();
        s->_PushParticleBC();
        s->_InitializeWallCharge();
    }
    _log << "Initialize the Poisson solver\n";
    Timer timer;
    base->_poissonSolver.Initialize();
    timer.PrintSectionTime("Poisson solver initialization");
    auto dev = base->_pData->DeviceContext();
    std::vector<int> species;
    for (auto* s : _solvers) species.push_back(s->_pData->DeviceSpecies());

    _log << "Start the main loop\n";
    for (int iteration = 0; iteration < nIterations; iteration++) {
        _log << "\n" << "Iteration #" << iteration << "\n";
        _log << "Time: " << iteration * timeStep << "\n";
        _log << Indent(1) << "Compute the electric field\n";
        for (auto* s : _solvers) s->_pData->PushParams();
        // total charge density of all species (+ the base solver's background), on the device
        device::Check(vt_charge_density(dev->ctx, species.data(), (int)species.size(),
                                        base->backgroundChargeDensity.empty() ? nullptr : base->backgroundChargeDensity.data()));
        const bool write = iteration % writeStep == 0;
        base->_poissonSolver.SolveOnDevice(write);
        timer.PrintSectionTime(Indent(1) + "Done");

        for (auto* s : _solvers) {
            if (iteration % stepMultipliers[s]) continue;   // sub-cycling
            _log << Indent(1) << "Update the PDF: " << s->_pData->species << "\n";
            s->_UpdatePDF();
        }
        timer.PrintSectionTime(Indent(1) + "Done");

        _log << Indent(1) << "Update the boundary conditions\n";
        std::unordered_map<int, double> wallCharge = base->_wallCharge;
        for (auto* s : _solvers) {
            if (s == base) continue;
            for (const auto& kv : s->_wallCharge) wallCharge[kv.first] += kv.second;
        }
        for (const auto& kv : wallCharge) {
            FieldBC bc;
            bc.type = FieldBCType::ChargedPlane;
            bc.chargeDensity = kv.second / base->_wallArea[kv.first];
            base->SetFieldBC(kv.first, bc);
        }
        if (write) {
            const size_t n = base->_mesh->tets.size();
            std::vector<double> rho(n);
            device::Check(vt_field_get(dev->ctx, rho.data(), nullptr, nullptr));
            for (auto* s : _solvers) {
                s->_rho = rho;
                s->_phi = base->_poissonSolver.Potential();
                s->_field = base->_poissonSolver.ElectricField();
                s->_pData->SyncFromDevice();   // Tucker: the host mirror feeds the distribution dump
                s->_WriteResults(iteration);
            }
        }
    }
    for (auto* s : _solvers) s->_pData->SyncFromDevice();
}

template class MulticomponentSolver<Full>;
template class MulticomponentSolver<Tucker>;

}  // namespace VlasovTucker
