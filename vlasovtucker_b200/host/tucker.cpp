#include "tucker.h"

#include <cmath>
#include <stdexcept>
#include <vector>

#include "la_small.h"

using Eigen::MatrixXd;

namespace VlasovTucker {

namespace {
// Y = X x_mode M  (M is q x dim_mode): contracts dimension `mode` of X with the columns of M.
Tensor3d ModeProduct(const Tensor3d& X, const MatrixXd& M, int mode)
{
    const long d[3] = {X.dimension(0), X.dimension(1), X.dimension(2)};
    long e[3] = {d[0], d[1], d[2]};
    e[mode] = M.rows();
    Tensor3d Y(e[0], e[1], e[2]);
    for (long i2 = 0; i2 < e[2]; i2++)
        for (long i1 = 0; i1 < e[1]; i1++)
            for (long i0 = 0; i0 < e[0]; i0++) {
                const long out[3] = {i0, i1, i2};
                long in[3] = {i0, i1, i2};
                double s = 0;
                for (long k = 0; k < d[mode]; k++) {
                    in[mode] = k;
                    s += M(out[mode], k) * X(in[0], in[1], in[2]);
                }
                Y(i0, i1, i2) = s;
            }
    return Y;
}
}  // namespace

Tucker::Tucker() : _n({0, 0, 0}), _r({0, 0, 0}) {}

Tucker::Tucker(int n0, int n1, int n2, int r0, int r1, int r2) : _n({n0, n1, n2}), _r({r0, r1, r2})
{
    _u[0] = MatrixXd::Zero(n0, r0);
    _u[1] = MatrixXd::Zero(n1, r1);
    _u[2] = MatrixXd::Zero(n2, r2);
    _core = Tensor3d(r0, r1, r2);
}

Tucker::Tucker(const Tensor3d& tensor, double precision, int maxRank)
{
    _n = {(int)tensor.dimension(0), (int)tensor.dimension(1), (int)tensor.dimension(2)};
    _ComputeU(tensor, precision, maxRank);
    _r = {(int)_u[0].cols(), (int)_u[1].cols(), (int)_u[2].cols()};
    // core = X x1 U0^T x2 U1^T x3 U2^T   (the reference forms U0^T X_(0) kron(U1,U2), tucker.cpp:48)
    Tensor3d c = ModeProduct(tensor, _u[0].transpose(), 0);
    c = ModeProduct(c, _u[1].transpose(), 1);
    _core = ModeProduct(c, _u[2].transpose(), 2);
}

Tucker::Tucker(const Tensor3d& core, const std::array<MatrixXd, 3>& u) : _u(u), _core(core)
{
    _n = {(int)_u[0].rows(), (int)_u[1].rows(), (int)_u[2].rows()};
    _r = {(int)_core.dimension(0), (int)_core.dimension(1), (int)_core.dimension(2)};
}

// Rounding (tucker.cpp:66-98): orthogonalise the factors, pull the triangular parts into the
// core, HOSVD-truncate the small core, rotate the factors back.  Economy sizes are used: the
// reference pads Q with zero columns when a factor has more columns than rows, which only adds
// zero rows/singular values to the auxiliary core.
Tucker& Tucker::Compress(double precision, int maxRank)
{
    std::array<MatrixXd, 3> Q, R;
    for (int i = 0; i < 3; i++) {
        Q[i] = la::ThinQ(_u[i]);
        R[i] = Q[i].transpose() * _u[i];
    }
    Tensor3d aux = ModeProduct(_core, R[0], 0);
    aux = ModeProduct(aux, R[1], 1);
    aux = ModeProduct(aux, R[2], 2);
    Tucker small(aux, precision, maxRank);
    _core = small._core;
    for (int i = 0; i < 3; i++) _u[i] = Q[i] * small._u[i];
    _r = {(int)_u[0].cols(), (int)_u[1].cols(), (int)_u[2].cols()};
    return *this;
}

Tensor3d Tucker::Reconstructed() const
{
    Tensor3d t = ModeProduct(_core, _u[0], 0);
    t = ModeProduct(t, _u[1], 1);
    return ModeProduct(t, _u[2], 2);
}

int Tucker::Size() const { return (int)(_core.size() + _u[0].size() + _u[1].size() + _u[2].size()); }
std::array<int, 3> Tucker::Dimensions() const { return _n; }
std::array<int, 3> Tucker::Ranks() const { return _r; }
std::array<MatrixXd, 3> Tucker::U() const { return _u; }
Tensor3d Tucker::Core() const { return _core; }

double Tucker::Sum() const
{
    // sum over all entries = core x1 (1^T U0) x2 (1^T U1) x3 (1^T U2)
    std::array<std::vector<double>, 3> s;
    for (int i = 0; i < 3; i++) {
        s[i].assign((size_t)_r[i], 0.0);
        for (int j = 0; j < _r[i]; j++)
            for (int k = 0; k < _n[i]; k++) s[i][(size_t)j] += _u[i](k, j);
    }
    double total = 0;
    for (int j0 = 0; j0 < _r[0]; j0++)
        for (int j1 = 0; j1 < _r[1]; j1++)
            for (int j2 = 0; j2 < _r[2]; j2++) total += _core(j0, j1, j2) * s[0][(size_t)j0] * s[1][(size_t)j1] * s[2][(size_t)j2];
    return total;
}

double Tucker::Norm() const
{
    const Tensor3d t = Reconstructed();
    double s = 0;
    for (long i = 0; i < t.size(); i++) s += t.data()[i] * t.data()[i];
    return std::sqrt(s);
}

double Tucker::operator()(int i0, int i1, int i2) const
{
    double el = 0;
    for (int j0 = 0; j0 < _r[0]; j0++)
        for (int j1 = 0; j1 < _r[1]; j1++)
            for (int j2 = 0; j2 < _r[2]; j2++) el += _core(j0, j1, j2) * _u[0](i0, j0) * _u[1](i1, j1) * _u[2](i2, j2);
    return el;
}

std::ostream& operator<<(std::ostream& out, const Tucker& t)
{
    out << "This is a 3D tensor in the Tucker format with \n";
    out << "r0 = " << t._r[0] << ", n0 = " << t._n[0] << "\n";
    out << "r1 = " << t._r[1] << ", n1 = " << t._n[1] << "\n";
    out << "r2 = " << t._r[2] << ", n2 = " << t._n[2];
    return out;
}

// Sum: block-diagonal core, concatenated factors (ranks add, tucker.cpp:190-228).
Tucker operator+(const Tucker& a, const Tucker& b)
{
    if (a.Dimensions() != b.Dimensions()) throw std::invalid_argument("Different shapes in sum");
    Tucker r(a._n[0], a._n[1], a._n[2], a._r[0] + b._r[0], a._r[1] + b._r[1], a._r[2] + b._r[2]);
    for (int k2 = 0; k2 < a._r[2]; k2++)
        for (int k1 = 0; k1 < a._r[1]; k1++)
            for (int k0 = 0; k0 < a._r[0]; k0++) r._core(k0, k1, k2) = a._core(k0, k1, k2);
    for (int k2 = 0; k2 < b._r[2]; k2++)
        for (int k1 = 0; k1 < b._r[1]; k1++)
            for (int k0 = 0; k0 < b._r[0]; k0++) r._core(a._r[0] + k0, a._r[1] + k1, a._r[2] + k2) = b._core(k0, k1, k2);
    for (int i = 0; i < 3; i++)
        for (int row = 0; row < a._n[i]; row++) {
            for (int j = 0; j < a._r[i]; j++) r._u[i](row, j) = a._u[i](row, j);
            for (int j = 0; j < b._r[i]; j++) r._u[i](row, a._r[i] + j) = b._u[i](row, j);
        }
    return r;
}

Tucker& Tucker::operator+=(const Tucker& t) { return *this = *this + t; }
Tucker& Tucker::operator-=(const Tucker& t) { return *this = *this - t; }
Tucker& Tucker::operator*=(const Tucker& t) { return *this = *this * t; }
Tucker& Tucker::operator*=(double d) { return *this = *this * d; }

Tucker operator-(const Tucker& a, const Tucker& b) { return a + (-1.0) * b; }

// Hadamard product: Kronecker core (index kA*rB + kB), row-wise Kronecker factors (ranks multiply,
// tucker.cpp:259-300).
Tucker operator*(const Tucker& a, const Tucker& b)
{
    if (a.Dimensions() != b.Dimensions()) throw std::invalid_argument("Different shapes in mult");
    Tucker r(a._n[0], a._n[1], a._n[2], a._r[0] * b._r[0], a._r[1] * b._r[1], a._r[2] * b._r[2]);
    for (int k2 = 0; k2 < r._r[2]; k2++)
        for (int k1 = 0; k1 < r._r[1]; k1++)
            for (int k0 = 0; k0 < r._r[0]; k0++)
                r._core(k0, k1, k2) = a._core(k0 / b._r[0], k1 / b._r[1], k2 / b._r[2]) *
                                      b._core(k0 % b._r[0], k1 % b._r[1], k2 % b._r[2]);
    for (int i = 0; i < 3; i++)
        for (int row = 0; row < a._n[i]; row++)
            for (int ja = 0; ja < a._r[i]; ja++)
                for (int jb = 0; jb < b._r[i]; jb++) r._u[i](row, ja * b._r[i] + jb) = a._u[i](row, ja) * b._u[i](row, jb);
    return r;
}

Tucker operator*(double d, const Tucker& t)
{
    Tucker r = t;   // scaling touches the core only (tucker.cpp:302-316)
    for (long i = 0; i < r._core.size(); i++) r._core.data()[i] *= d;
    return r;
}
Tucker operator*(const Tucker& t, double d) { return d * t; }
Tucker operator-(const Tucker& t) { return (-1.0) * t; }

MatrixXd Unfolding(const Tensor3d& t, int index)
{
    const int I[3] = {(int)t.dimension(0), (int)t.dimension(1), (int)t.dimension(2)};
    const int a = (index + 1) % 3, b = (index + 2) % 3;   // column = i_b + i_a * I_b
    MatrixXd m(I[index], I[a] * I[b]);
    int i[3];
    for (i[2] = 0; i[2] < I[2]; i[2]++)
        for (i[1] = 0; i[1] < I[1]; i[1]++)
            for (i[0] = 0; i[0] < I[0]; i[0]++) m(i[index], i[b] + i[a] * I[b]) = t(i[0], i[1], i[2]);
    return m;
}

Tensor3d Folding(int I0, int I1, int I2, const MatrixXd& m, int index)
{
    const int I[3] = {I0, I1, I2};
    const int a = (index + 1) % 3, b = (index + 2) % 3;
    Tensor3d t(I0, I1, I2);
    int i[3];
    for (i[2] = 0; i[2] < I2; i[2]++)
        for (i[1] = 0; i[1] < I1; i[1]++)
            for (i[0] = 0; i[0] < I0; i[0]++) t(i[0], i[1], i[2]) = m(i[index], i[b] + i[a] * I[b]);
    return t;
}

// Factor i = leading left singular vectors of the mode-i unfolding.  Truncation rule of the
// reference (tucker.cpp:450-461): keep sigma_j while sigma_j > precision*|sigma|_2/sqrt(3) and
// fewer than maxRank are kept; always keep the first.
void Tucker::_ComputeU(const Tensor3d& tensor, double precision, int rmax)
{
    for (int i = 0; i < 3; i++) {
        MatrixXd U;
        std::vector<double> sv;
        la::LeftSVD(Unfolding(tensor, i), U, sv);
        double n2 = 0;
        for (double s : sv) n2 += s * s;
        const double threshold = precision * std::sqrt(n2) / std::sqrt(3);
        std::vector<int> keep;
        for (int j = 0; j < (int)U.cols(); j++)
            if (keep.empty() || (sv[(size_t)j] > threshold && (int)keep.size() < rmax)) keep.push_back(j);
        _u[i] = MatrixXd(U.rows(), (long)keep.size());
        for (size_t c = 0; c < keep.size(); c++)
            for (long r = 0; r < U.rows(); r++) _u[i](r, (long)c) = U(r, keep[c]);
    }
}

}  // namespace VlasovTucker
