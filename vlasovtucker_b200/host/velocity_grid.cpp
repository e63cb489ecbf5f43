#include "velocity_grid.h"

namespace VlasovTucker {

VelocityGrid::VelocityGrid(std::array<int, 3> n, Vector3d lo, Vector3d hi) : nCells(n), minV(lo), maxV(hi)
{
    nCellsTotal = n[0] * n[1] * n[2];
    for (int j = 0; j < 3; j++) step[j] = (maxV[j] - minV[j]) / (nCells[j] - 1);   // nodes, not cells
    cellVolume = step[0] * step[1] * step[2];
    for (int j = 0; j < 3; j++) {
        v[j] = Tensor3d(n[0], n[1], n[2]);
        for (int i2 = 0; i2 < n[2]; i2++)
            for (int i1 = 0; i1 < n[1]; i1++)
                for (int i0 = 0; i0 < n[0]; i0++) {
                    const int idx[3] = {i0, i1, i2};
                    v[j](i0, i1, i2) = minV[j] + idx[j] * step[j];
                }
        // tridiagonal central difference; the PDF is taken as zero outside the grid
        d[j] = Eigen::MatrixXd::Zero(n[j], n[j]);
        for (int i = 0; i < n[j]; i++) {
            if (i + 1 < n[j]) d[j](i, i + 1) = 1;
            if (i > 0) d[j](i, i - 1) = -1;
        }
        d[j] /= 2 * step[j];
    }
}

Vector3d VelocityGrid::At(int i0, int i1, int i2) const
{
    return {minV[0] + i0 * step[0], minV[1] + i1 * step[1], minV[2] + i2 * step[2]};
}

}  // namespace VlasovTucker
