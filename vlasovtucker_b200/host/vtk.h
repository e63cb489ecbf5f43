// Legacy-VTK (ASCII) output of the public API.  Host-side I/O only: nothing here touches the device;
// the solvers hand over host copies of the moments (Solver::SyncFromDevice).
// Reference interface: src/vtk.h:15-28 — same four entry points, same argument order and defaults.
#pragma once
// user code written against the reference gets these through this header (test/poisson_test.cpp opens
// std::ofstream without including <fstream> itself), so they stay part of the interface
#include <array>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "mesh.h"
#include "typedefs.h"
#include "velocity_grid.h"

namespace VlasovTucker {

// unstructured grid (points + tets) alone
void WriteMeshVTK(std::string fileName, const Mesh& mesh);

// the grid plus one value per tet; an empty vector writes the grid only
void WriteCellScalarDataVTK(std::string fileName, const Mesh& mesh, const std::vector<double>& data = {});

// the grid plus one 3-vector per tet
void WriteCellVectorDataVTK(std::string fileName, const Mesh& mesh, const std::vector<Vector3d>& data = {});

// one tet's distribution function as a structured-points volume over the velocity grid
void WriteDistributionVTK(std::string fileName, const VelocityGrid& velocityGrid, const Tensor3d& distribution);

}  // namespace VlasovTucker
