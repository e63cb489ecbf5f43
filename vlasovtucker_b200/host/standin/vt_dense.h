// Minimal stand-ins for the Eigen types that leak through VlasovTucker's public API
// (Tensor3d = Eigen::Tensor<double,3>, typedefs.h:9; Eigen::MatrixXd in Tucker::U() and
// VelocityGrid::d).  The reference vendors Eigen 3.4.0 without Eigen/Core, so its own headers do
// not compile anywhere; these classes implement just the members the drivers, tests and host
// classes use, with Eigen's storage order (column-major, first index fastest).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <initializer_list>
#include <ostream>
#include <stdexcept>
#include <vector>

namespace Eigen {

typedef long Index;
constexpr int Dynamic = -1;

template <typename Scalar, int Rank>
class Tensor {
public:
    Tensor() { dims_.fill(0); }
    template <typename... Ix>
    explicit Tensor(Index d0, Ix... rest) : dims_{{d0, static_cast<Index>(rest)...}}
    {
        static_assert(sizeof...(Ix) + 1 == Rank, "one extent per dimension");
        Index n = 1;
        for (Index d : dims_) n *= d;
        data_.assign((size_t)n, Scalar());
    }
    Index dimension(int i) const { return dims_[i]; }
    const std::array<Index, Rank>& dimensions() const { return dims_; }
    Index size() const { return (Index)data_.size(); }
    Scalar* data() { return data_.data(); }
    const Scalar* data() const { return data_.data(); }

    template <typename... Ix>
    Scalar& operator()(Index i0, Ix... rest)
    {
        return data_[offset({{i0, static_cast<Index>(rest)...}})];
    }
    template <typename... Ix>
    const Scalar& operator()(Index i0, Ix... rest) const
    {
        return data_[offset({{i0, static_cast<Index>(rest)...}})];
    }

    Tensor& setZero() { return setConstant(Scalar(0)); }
    Tensor& setConstant(Scalar v)
    {
        std::fill(data_.begin(), data_.end(), v);
        return *this;
    }
    // Eigen's setRandom() draws uniformly from [0, 1) for floating-point tensors
    Tensor& setRandom()
    {
        for (auto& x : data_) x = Scalar(std::rand()) / (Scalar(RAND_MAX) + Scalar(1));
        return *this;
    }
    Tensor abs() const
    {
        Tensor r(*this);
        for (auto& x : r.data_) x = std::abs(x);
        return r;
    }
    Tensor& operator+=(const Tensor& o) { return zip(o, [](Scalar a, Scalar b) { return a + b; }); }
    Tensor& operator-=(const Tensor& o) { return zip(o, [](Scalar a, Scalar b) { return a - b; }); }
    Tensor& operator*=(const Tensor& o) { return zip(o, [](Scalar a, Scalar b) { return a * b; }); }
    friend Tensor operator+(Tensor a, const Tensor& b) { return a += b; }
    friend Tensor operator-(Tensor a, const Tensor& b) { return a -= b; }
    friend Tensor operator*(Tensor a, const Tensor& b) { return a *= b; }
    friend Tensor operator*(Scalar s, Tensor a)
    {
        for (auto& x : a.data_) x = s * x;
        return a;
    }
    friend Tensor operator*(Tensor a, Scalar s) { return s * a; }
    Scalar sumAll() const
    {
        Scalar s = 0;
        for (auto x : data_) s += x;
        return s;
    }

    friend std::ostream& operator<<(std::ostream& os, const Tensor& t)
    {
        // rank-3 tensors print as the matrix dim0 x (dim1*dim2), as Eigen does
        const Index rows = Rank > 0 ? t.dims_[0] : 1;
        const Index cols = rows ? t.size() / rows : 0;
        for (Index r = 0; r < rows; r++) {
            for (Index c = 0; c < cols; c++) os << (c ? " " : "") << t.data_[(size_t)(r + rows * c)];
            if (r + 1 < rows) os << "\n";
        }
        return os;
    }

private:
    size_t offset(const std::array<Index, Rank>& ix) const
    {
        size_t o = 0, stride = 1;
        for (int d = 0; d < Rank; d++) {
            o += (size_t)ix[d] * stride;
            stride *= (size_t)dims_[d];
        }
        return o;
    }
    template <class F>
    Tensor& zip(const Tensor& o, F f)
    {
        if (o.dims_ != dims_) throw std::invalid_argument("tensor shapes differ");
        for (size_t i = 0; i < data_.size(); i++) data_[i] = f(data_[i], o.data_[i]);
        return *this;
    }
    std::array<Index, Rank> dims_;
    std::vector<Scalar> data_;
};

// Column-major dense matrix / vector of doubles.
class MatrixXd {
public:
    MatrixXd() : r_(0), c_(0) {}
    MatrixXd(Index r, Index c) : r_(r), c_(c), a_((size_t)(r * c), 0.0) {}
    static MatrixXd Zero(Index r, Index c) { return MatrixXd(r, c); }
    static MatrixXd Identity(Index r, Index c)
    {
        MatrixXd m(r, c);
        for (Index i = 0; i < std::min(r, c); i++) m(i, i) = 1.0;
        return m;
    }
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    void resize(Index r, Index c)
    {
        r_ = r;
        c_ = c;
        a_.assign((size_t)(r * c), 0.0);
    }
    double& operator()(Index i, Index j) { return a_[(size_t)(i + r_ * j)]; }
    const double& operator()(Index i, Index j) const { return a_[(size_t)(i + r_ * j)]; }
    double* data() { return a_.data(); }
    const double* data() const { return a_.data(); }
    MatrixXd transpose() const
    {
        MatrixXd t(c_, r_);
        for (Index i = 0; i < r_; i++)
            for (Index j = 0; j < c_; j++) t(j, i) = (*this)(i, j);
        return t;
    }
    MatrixXd& operator/=(double d)
    {
        for (auto& x : a_) x /= d;
        return *this;
    }
    MatrixXd& operator*=(double d)
    {
        for (auto& x : a_) x *= d;
        return *this;
    }
    friend MatrixXd operator*(const MatrixXd& a, const MatrixXd& b)
    {
        if (a.c_ != b.r_) throw std::invalid_argument("matrix shapes differ in product");
        MatrixXd m(a.r_, b.c_);
        for (Index j = 0; j < b.c_; j++)
            for (Index k = 0; k < a.c_; k++) {
                const double bkj = b(k, j);
                for (Index i = 0; i < a.r_; i++) m(i, j) += a(i, k) * bkj;
            }
        return m;
    }
    MatrixXd eval() const { return *this; }
    double norm() const
    {
        double s = 0;
        for (double x : a_) s += x * x;
        return std::sqrt(s);
    }
    friend std::ostream& operator<<(std::ostream& os, const MatrixXd& m)
    {
        for (Index i = 0; i < m.r_; i++) {
            for (Index j = 0; j < m.c_; j++) os << (j ? " " : "") << m(i, j);
            if (i + 1 < m.r_) os << "\n";
        }
        return os;
    }

private:
    Index r_, c_;
    std::vector<double> a_;
};

class VectorXd {
public:
    VectorXd() {}
    explicit VectorXd(Index n) : a_((size_t)n, 0.0) {}
    Index size() const { return (Index)a_.size(); }
    double& operator()(Index i) { return a_[(size_t)i]; }
    const double& operator()(Index i) const { return a_[(size_t)i]; }
    double& operator[](Index i) { return a_[(size_t)i]; }
    const double& operator[](Index i) const { return a_[(size_t)i]; }
    double norm() const
    {
        double s = 0;
        for (double x : a_) s += x * x;
        return std::sqrt(s);
    }

private:
    std::vector<double> a_;
};

}  // namespace Eigen
