// The two value types every public signature of the VlasovTucker API is written in
// (reference: src/typedefs.h:9-10).
//   Tensor3d  a dense rank-3 array of doubles, first index fastest: one tet's f(v0, v1, v2) in full
//             format, a Tucker core, a velocity-coordinate table.  On the device the same memory layout
//             is one row of the species state (include/vt_b200.h, vt_species_set_pdf).
//   Vector3d  a plain triple (position, velocity, field), no alignment requirement, trivially copyable
//             into the double[3] arguments of the C ABI.
#pragma once
#include <array>
#include <unsupported/Eigen/CXX11/Tensor>

namespace VlasovTucker {
using Vector3d = std::array<double, 3>;
using Tensor3d = Eigen::Tensor<double, 3>;
}  // namespace VlasovTucker
