// Several species advanced together: one shared Poisson solve per iteration, each species stepped
// by its own Solver with its own sub-cycling multiplier (the public type of
// src/multicomponent_solver.h:7-30; the loop itself is multicomponent_solver.cpp:27-135).
#pragma once
#include <limits.h>

#include <map>
#include <vector>

#include "solver.h"

namespace VlasovTucker {
template <typename TensorType>
class MulticomponentSolver {
    using SolverPtr = Solver<TensorType>*;

public:
    // run parameters (public data, assigned by the driver before Solve)
    int nIterations = 0;
    double timeStep = 0;
    int writeStep = INT_MAX;
    // species s updates when iteration % stepMultipliers[s] == 0, with dt * multiplier;
    // keyed by solver address as the drivers do (examples/sheath.cpp:131-132)
    std::map<SolverPtr, int> stepMultipliers;

    // baseSolver owns the field solve (its PoissonSolver and field BCs are the ones used)
    MulticomponentSolver(SolverPtr baseSolver);
    void AddSolver(SolverPtr solver);
    void Solve();

private:
    Log _log;
    std::vector<SolverPtr> _solvers;   // baseSolver first
};
}  // namespace VlasovTucker
