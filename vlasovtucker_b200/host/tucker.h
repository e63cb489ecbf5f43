// Tucker-format 3-D tensor of the public API (reference: src/tucker.h:9-87): core r0 x r1 x r2
// plus three factor matrices n_i x r_i.  This host class is the user-facing value type (tests
// construct it from a Tensor3d, add, scale and round it); the per-step Tucker update of the
// solver runs on the device.
#pragma once
#include <Eigen/Dense>
#include <array>
#include <ostream>

#include "typedefs.h"

namespace VlasovTucker {
class Tucker {
    using Factors = std::array<Eigen::MatrixXd, 3>;

public:
    // ---- construction
    Tucker();
    Tucker(const Tensor3d& core, const Factors& u);                             // from parts
    Tucker(const Tensor3d& tensor, double precision = 0, int maxRank = 1e+6);   // truncated HOSVD
    Tucker(int n0, int n1, int n2, int r0, int r1, int r2);                      // zero tensor of given ranks

    // ---- element-wise algebra; ranks add under +/-, multiply under * (Hadamard), scalars touch the core
    friend Tucker operator-(const Tucker& t);
    friend Tucker operator*(double d, const Tucker& t);
    friend Tucker operator*(const Tucker& t, double d);
    friend Tucker operator*(const Tucker& t1, const Tucker& t2);
    friend Tucker operator-(const Tucker& t1, const Tucker& t2);
    friend Tucker operator+(const Tucker& t1, const Tucker& t2);
    Tucker& operator*=(double d);
    Tucker& operator*=(const Tucker& t);
    Tucker& operator-=(const Tucker& t);
    Tucker& operator+=(const Tucker& t);

    // ---- rounding: QR of the factors, HOSVD of the transformed core, keep sigma_j > precision*|sigma|/sqrt(3)
    Tucker& Compress(double precision = 0, int maxRank = 1e+6);

    // ---- queries
    double operator()(int i0, int i1, int i2) const;
    Tensor3d Reconstructed() const;
    double Sum() const;
    double Norm() const;
    Tensor3d Core() const;
    Factors U() const;
    std::array<int, 3> Ranks() const;
    std::array<int, 3> Dimensions() const;
    int Size() const;   // doubles stored: core + factors

    friend std::ostream& operator<<(std::ostream& out, const Tucker& t);

private:
    void _ComputeU(const Tensor3d& tensor, double precision, int maxRank);

    std::array<int, 3> _n;
    std::array<int, 3> _r;
    Factors _u;
    Tensor3d _core;
};

// Tensor-matrix conversion (tucker.h:76-80): the first rows*cols entries of the tensor's column-major
// storage seen as a rows x cols matrix.  The reference returns an Eigen::Map onto the tensor; with
// the stand-in dense types this is a copy.
template <typename Scalar, int rank, typename sizeType>
Eigen::MatrixXd TensorToMatrix(const Eigen::Tensor<Scalar, rank>& tensor, const sizeType rows, const sizeType cols)
{
    Eigen::MatrixXd m((int)rows, (int)cols);
    const Scalar* src = tensor.data();
    for (long i = 0; i < (long)rows * (long)cols; i++) m.data()[i] = (double)src[i];
    return m;
}

// Mode-`index` unfolding with the reference's column order (tucker.cpp:337-392): mode 0 columns
// run i2 fastest then i1; mode 1: i0 fastest then i2; mode 2: i1 fastest then i0.
Eigen::MatrixXd Unfolding(const Tensor3d& tensor, int index);
Tensor3d Folding(int I0, int I1, int I2, const Eigen::MatrixXd& unfolding, int index);
}  // namespace VlasovTucker
