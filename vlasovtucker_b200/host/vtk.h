// ASCII legacy-VTK writers of the public API (reference: src/vtk.h:15-28).  Host I/O only.
#pragma once
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "mesh.h"
#include "typedefs.h"
#include "velocity_grid.h"

namespace VlasovTucker {
void WriteCellScalarDataVTK(std::string fileName, const Mesh& mesh, const std::vector<double>& data = {});
void WriteCellVectorDataVTK(std::string fileName, const Mesh& mesh, const std::vector<Vector3d>& data = {});
void WriteMeshVTK(std::string fileName, const Mesh& mesh);
void WriteDistributionVTK(std::string fileName, const VelocityGrid& velocityGrid, const Tensor3d& distribution);
}  // namespace VlasovTucker
