// Multi-species loop with one shared Poisson solve and per-species sub-cycling
// (reference: src/multicomponent_solver.h:7-30).
#pragma once
#include <limits.h>

#include <map>
#include <vector>

#include "solver.h"

namespace VlasovTucker {
template <typename TensorType>
class MulticomponentSolver {
public:
    MulticomponentSolver(Solver<TensorType>* baseSolver);
    void AddSolver(Solver<TensorType>* solver);
    void Solve();

public:
    double timeStep = 0;
    std::map<Solver<TensorType>*, int> stepMultipliers;   // update a species once in several steps
    int nIterations = 0;
    int writeStep = INT_MAX;

private:
    std::vector<Solver<TensorType>*> _solvers;
    Log _log;
};
}  // namespace VlasovTucker
