// Uniform velocity-space lattice (reference: src/velocity_grid.cpp:9-59): node coordinates as three
// rank-3 tensors for the Tucker algebra, and the dense central-difference operators whose first and
// last rows see zeros beyond the lattice (the Tucker derivative; the full-format kernels wrap instead).
#include "velocity_grid.h"

namespace VlasovTucker {

namespace {
// (D f)_i = (f_{i+1} - f_{i-1}) / 2h with f = 0 outside [0, m)
Eigen::MatrixXd CentralDifference(int m, double h)
{
    Eigen::MatrixXd D = Eigen::MatrixXd::Zero(m, m);
    const double w = 1.0 / (2 * h);
    for (int i = 0; i + 1 < m; i++) {
        D(i, i + 1) = w;
        D(i + 1, i) = -w;
    }
    return D;
}
}  // namespace

VelocityGrid::VelocityGrid(std::array<int, 3> n, Vector3d lo, Vector3d hi) : nCells(n), minV(lo), maxV(hi)
{
    nCellsTotal = n[0] * n[1] * n[2];
    cellVolume = 1.0;
    for (int axis = 0; axis < 3; axis++) {
        step[axis] = (hi[axis] - lo[axis]) / (n[axis] - 1);        // n counts nodes
        cellVolume *= step[axis];
    }
    for (int axis = 0; axis < 3; axis++) {
        Tensor3d& coord = v[axis];
        coord = Tensor3d(n[0], n[1], n[2]);
        for (int i2 = 0; i2 < n[2]; i2++)
            for (int i1 = 0; i1 < n[1]; i1++)
                for (int i0 = 0; i0 < n[0]; i0++) {
                    const int node = axis == 0 ? i0 : axis == 1 ? i1 : i2;
                    coord(i0, i1, i2) = lo[axis] + node * step[axis];
                }
        d[axis] = CentralDifference(n[axis], step[axis]);
    }
}

Vector3d VelocityGrid::At(int i0, int i1, int i2) const
{
    Vector3d p;
    const int idx[3] = {i0, i1, i2};
    for (int axis = 0; axis < 3; axis++) p[axis] = minV[axis] + idx[axis] * step[axis];
    return p;
}

}  // namespace VlasovTucker
