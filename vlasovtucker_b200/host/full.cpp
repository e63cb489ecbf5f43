#include "full.h"

#include "device.h"

namespace VlasovTucker {

Full::Full(const Tensor3d& tensor) : _tensor(tensor)
{
    _dims = {(int)tensor.dimension(0), (int)tensor.dimension(1), (int)tensor.dimension(2)};
}

Full::Full(const Full& o) : _tensor(o.Value()), _dims(o._dims) {}

Full::Full(Full&& o) noexcept
    : _tensor(std::move(o._tensor)), _dev(std::move(o._dev)), _species(o._species), _tet(o._tet), _dims(o._dims)
{
    o._tet = -1;
}

Full& Full::operator=(const Full& o)
{
    if (this == &o) return *this;
    if (OnDevice()) {
        const Tensor3d v = o.Value();
        if ((int)v.dimension(0) != _dims[0] || (int)v.dimension(1) != _dims[1] || (int)v.dimension(2) != _dims[2])
            throw std::invalid_argument("Different shapes in assignment to a device-resident PDF");
        device::Check(vt_species_set_pdf(_dev->ctx, _species, _tet, 1, v.data()));
    } else {
        _tensor = o.Value();
        _dims = o._dims;
    }
    return *this;
}

Full& Full::operator=(Full&& o)
{
    if (this == &o) return *this;
    if (OnDevice()) {
        // assignment to a device-resident row uploads; a shape mismatch or a CUDA failure propagates
        // exactly as from the copy assignment (nothing is swallowed)
        *this = static_cast<const Full&>(o);
        return *this;
    }
    _tensor = std::move(o._tensor);
    _dev = std::move(o._dev);
    _species = o._species;
    _tet = o._tet;
    _dims = o._dims;
    o._tet = -1;
    return *this;
}

Full Full::DeviceRow(std::shared_ptr<device::MeshContext> ctx, int species, int tet, std::array<int, 3> dims)
{
    Full f;
    f._dev = std::move(ctx);
    f._species = species;
    f._tet = tet;
    f._dims = dims;
    return f;
}

Tensor3d Full::Value() const
{
    if (!OnDevice()) return _tensor;
    Tensor3d t(_dims[0], _dims[1], _dims[2]);
    device::Check(vt_species_get_pdf(_dev->ctx, _species, _tet, 1, t.data()));
    return t;
}

int Full::Size() const { return _dims[0] * _dims[1] * _dims[2]; }
std::array<int, 3> Full::Dimensions() const { return _dims; }
double Full::operator()(int i0, int i1, int i2) const { return Value()(i0, i1, i2); }
Tensor3d Full::Reconstructed() const { return Value(); }
double Full::Sum() const { return Value().sumAll(); }
Full& Full::Compress(double, int) { return *this; }

std::ostream& operator<<(std::ostream& out, const Full& t) { return out << t.Value(); }

Full& Full::operator+=(const Full& t) { return *this = *this + t; }
Full& Full::operator-=(const Full& t) { return *this = *this - t; }
Full& Full::operator*=(const Full& t) { return *this = *this * t; }
Full& Full::operator*=(double d) { return *this = *this * d; }

Full operator+(const Full& a, const Full& b) { return Full(a.Value() + b.Value()); }
Full operator-(const Full& a, const Full& b) { return Full(a.Value() - b.Value()); }
Full operator*(const Full& a, const Full& b) { return Full(a.Value() * b.Value()); }
Full operator*(double d, const Full& t) { return Full(d * t.Value()); }
Full operator*(const Full& t, double d) { return d * t; }
Full operator-(const Full& t) { return (-1) * t; }

}  // namespace VlasovTucker
