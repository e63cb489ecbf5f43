// Device binding of the host classes: one vt_ctx per Mesh, created on first use after
// Mesh::Reconstruct, shared by every ParticleData / Solver / PoissonSolver built on that mesh.
// Everything below include/vt_b200.h runs on the GPU; there is no CPU fallback.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>

#include "../../include/vt_b200.h"
#include "mesh.h"

namespace VlasovTucker {
namespace device {

struct MeshContext {
    vt_ctx* ctx = nullptr;
    const Mesh* mesh = nullptr;
    const void* poissonOwner = nullptr;   // which PoissonSolver last called vt_poisson_setup
    ~MeshContext();
};

// Throws std::runtime_error carrying vt_last_error() when rc != 0.
void Check(int rc);
// The context of `mesh` (uploads the mesh tables on first call).
std::shared_ptr<MeshContext> ContextOf(const Mesh* mesh);

}  // namespace device
}  // namespace VlasovTucker
