// Uniform 3-D velocity grid (the public type of src/velocity_grid.h:11-30; same member names).
// The device keeps only n, min and step (vt_species_create); the coordinate tensors and the
// difference matrices below exist for host-side users of the API (drivers, VTK output, Tucker tests).
#pragma once
#include <Eigen/Dense>
#include <array>

#include "typedefs.h"

namespace VlasovTucker {
struct VelocityGrid {
    // extents and resolution, as given
    std::array<int, 3> nCells;
    Vector3d minV;
    Vector3d maxV;
    // derived: node spacing (max - min)/(n - 1), number of nodes, quadrature weight prod(step)
    std::array<double, 3> step;
    int nCellsTotal;
    double cellVolume;
    // v[k](i0,i1,i2) = k-th velocity component at node (i0,i1,i2); d[k] = central difference along
    // axis k with zeros outside the grid
    std::array<Tensor3d, 3> v;
    std::array<Eigen::MatrixXd, 3> d;

    VelocityGrid(std::array<int, 3> nCells, Vector3d minV, Vector3d maxV);

    // velocity vector of node (i0, i1, i2)
    Vector3d At(int i0, int i1, int i2) const;
};
}  // namespace VlasovTucker
