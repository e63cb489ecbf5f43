// Everything a driver needs (reference: src/header.h): the drivers in examples/ and test/ include
// only this file.
#pragma once
#include "log.h"
#include "timer.h"
#include "typedefs.h"
#include "vtk.h"

#include "full.h"
#include "tucker.h"

#include "mesh.h"
#include "velocity_grid.h"
#include "particle_data.h"

#include "solver.h"
#include "multicomponent_solver.h"
