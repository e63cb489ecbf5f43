// Dense 3-D tensor of the public API (reference: src/full.h:9-44).
//
// A Full is either a host value (what drivers and tests build from a Tensor3d) or a handle to
// one tet's row of a species' state on the GPU (what ParticleData::pdf holds).  A handle
// materialises on demand (Reconstructed, operator(), Sum, arithmetic); assigning to a handle
// writes the row back to the device, so `pdf[i] = Full(tensor)` keeps working.
#pragma once
#include <array>
#include <memory>
#include <ostream>

#include "typedefs.h"

namespace VlasovTucker {
namespace device {
struct MeshContext;
}

class Full {
public:
    // ---- device binding (not part of the reference API)
    static Full DeviceRow(std::shared_ptr<device::MeshContext> ctx, int species, int tet, std::array<int, 3> dims);
    bool OnDevice() const { return _tet >= 0; }

    // ---- construction / assignment
    Full() {}
    Full(const Tensor3d& tensor);
    Full(Full&& other) noexcept;             // moves keep the binding (std::vector growth)
    Full(const Full& other);                 // copies materialise: the copy is a host value
    Full& operator=(Full&& other);   // uploads when the target is a device-resident row: may throw
    Full& operator=(const Full& other);      // into a handle: uploads the row

    // ---- element-wise algebra (same operator set as Tucker, so Solver<T> code is format agnostic)
    friend Full operator-(const Full& t);
    friend Full operator*(double d, const Full& t);
    friend Full operator*(const Full& t, double d);
    friend Full operator*(const Full& t1, const Full& t2);
    friend Full operator-(const Full& t1, const Full& t2);
    friend Full operator+(const Full& t1, const Full& t2);
    Full& operator*=(double d);
    Full& operator*=(const Full& t);
    Full& operator-=(const Full& t);
    Full& operator+=(const Full& t);
    Full& Compress(double precision = 0, int maxRank = 1e+6);   // no-op, as the reference

    // ---- queries
    double operator()(int i0, int i1, int i2) const;
    Tensor3d Reconstructed() const;
    double Sum() const;
    std::array<int, 3> Dimensions() const;
    int Size() const;

    friend std::ostream& operator<<(std::ostream& out, const Full& t);

private:
    Tensor3d Value() const;   // host value, fetched from the device for handles

    Tensor3d _tensor;
    std::shared_ptr<device::MeshContext> _dev;
    int _species = -1, _tet = -1;
    std::array<int, 3> _dims{{0, 0, 0}};
};
}  // namespace VlasovTucker
