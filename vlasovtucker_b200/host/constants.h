// Physical constants; the values are those of the reference (src/constants.h:5-12) because they
// enter the results (epsilon0 in the Poisson RHS, boltzConst in the Maxwellian).
#pragma once

namespace VlasovTucker {
constexpr double pi = 3.14159265358979323846;
constexpr double atomicMass = 1.66e-27;      // kg
constexpr double elMass = 9.1e-31;           // kg
constexpr double elCharge = 1.6e-19;         // C
constexpr double epsilon0 = 8.85e-12;        // F/m
constexpr double electronvolt = 11604.518;   // K
constexpr double boltzConst = 1.38e-23;      // J/K
}  // namespace VlasovTucker
