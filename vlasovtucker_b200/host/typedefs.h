// Public typedefs of the VlasovTucker API (reference: src/typedefs.h:9-10).
#pragma once
#include <array>
#include <unsupported/Eigen/CXX11/Tensor>

namespace VlasovTucker {
using Tensor3d = Eigen::Tensor<double, 3>;
using Vector3d = std::array<double, 3>;
}  // namespace VlasovTucker
