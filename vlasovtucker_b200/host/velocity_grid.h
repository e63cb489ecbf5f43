// Uniform 3-D velocity grid (reference: src/velocity_grid.h:11-30).
#pragma once
#include <Eigen/Dense>
#include <array>

#include "typedefs.h"

namespace VlasovTucker {
struct VelocityGrid {
    VelocityGrid(std::array<int, 3> nCells, Vector3d minV, Vector3d maxV);

    Vector3d At(int i0, int i1, int i2) const;

    std::array<int, 3> nCells;
    int nCellsTotal;
    std::array<double, 3> step;
    double cellVolume;
    Vector3d maxV;
    Vector3d minV;
    std::array<Tensor3d, 3> v;           // coordinate tensors v_j(i0,i1,i2)
    std::array<Eigen::MatrixXd, 3> d;    // central-difference matrices, zero outside the grid
};
}  // namespace VlasovTucker
