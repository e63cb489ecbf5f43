// Umbrella include of the public API (reference: src/header.h:1-14).
#pragma once
#include "typedefs.h"

#include "mesh.h"
#include "particle_data.h"
#include "velocity_grid.h"

#include "full.h"
#include "tucker.h"

#include "multicomponent_solver.h"
#include "solver.h"

#include "log.h"
#include "timer.h"
#include "vtk.h"
