#include "device.h"

#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

namespace VlasovTucker {
namespace device {

MeshContext::~MeshContext()
{
    if (ctx) vt_ctx_destroy(ctx);
}

void Check(int rc)
{
    if (rc) throw std::runtime_error(vt_last_error());
}

std::shared_ptr<MeshContext> ContextOf(const Mesh* mesh)
{
    static std::mutex mu;
    static std::map<const Mesh*, std::weak_ptr<MeshContext>> table;
    std::lock_guard<std::mutex> lock(mu);
    auto it = table.find(mesh);
    if (it != table.end())
        if (auto alive = it->second.lock()) return alive;
    if (mesh->tets.empty()) throw std::runtime_error("Mesh::Reconstruct must be called before the mesh is used");
    auto mc = std::make_shared<MeshContext>();
    mc->mesh = mesh;
    // VT_DEVICES=0,1,2,...: a device group — the mesh is partitioned over the listed GPUs (an index may
    // repeat: several partitions on one GPU) and every Solver / PoissonSolver call runs on all of them.
    // VT_DEVICE=k (or nothing): one GPU.
    const char* devs = std::getenv("VT_DEVICES");
    std::vector<int> list;
    if (devs && *devs) {
        const char* p = devs;
        while (*p) {
            char* end = nullptr;
            const long v = std::strtol(p, &end, 10);
            if (end == p) break;
            list.push_back((int)v);
            p = (*end == ',') ? end + 1 : end;
        }
    }
    if (list.size() > 1) {
        Check(vt_ctx_create_group(list.data(), (int)list.size(), &mc->ctx));
    } else {
        const char* dev = std::getenv("VT_DEVICE");
        Check(vt_ctx_create(!list.empty() ? list[0] : (dev ? std::atoi(dev) : 0), &mc->ctx));
    }
    const FlatMesh& fm = mesh->Flat();
    Check(vt_mesh_upload(mc->ctx, (int)mesh->tets.size(), 0, fm.nbr.data(), fm.area.data(), fm.volume.data(),
                         fm.normal.data(), fm.entity.data(), fm.order.data()));
    table[mesh] = mc;
    return mc;
}

}  // namespace device
}  // namespace VlasovTucker
