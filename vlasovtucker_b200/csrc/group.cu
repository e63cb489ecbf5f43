// Device groups: one vt_ctx that drives several GPUs (or several "virtual ranks" on one GPU) from one
// host thread.  This is what lets the C++ host classes — Solver<T>::Solve, MulticomponentSolver<T>::
// Solve, PoissonSolver — and therefore the unchanged examples/oscillations.cpp and examples/sheath.cpp
// use every GPU of the box (VT_DEVICES=0,1,...): the group partitions the mesh, creates one ordinary
// context per device, wires their ghost rows directly to each other (vt_halo_attach_local,
// vt_poisson_comm_attach_local) and translates every call of the C ABI that carries per-tet data into
// calls on the members.  Everything that runs per step is what the one-process-per-GPU path runs: the
// step kernels with the fused halo push, the device-side barriers, the partitioned Poisson solve.
//
// Partition: the mesh arrives with a locality permutation (vt_mesh_upload's `order`: Morton order of
// the centroids for .msh meshes, bricks for the synthetic box); member r owns the r-th contiguous
// chunk of that sequence — compact parts without needing coordinates at upload time, deterministic,
// and the chunk order doubles as the member's own locality order.
#include "vt_internal.h"

#include <algorithm>
#include <cstring>
#include <numeric>

namespace vt {

struct Group {
    std::vector<vt_ctx*> m;
    int nTets = 0;
    std::vector<int32_t> owner, local;              // per global tet: member, caller-local index there
    std::vector<std::vector<int32_t>> owned, ghost; // per member: global ids of its owned / ghost rows
    std::vector<std::vector<int32_t>> peers;        // per member: ranks it exchanges with (sorted)
    std::vector<std::vector<int32_t>> pushPeer, pushRank, pushRow;   // per member, 4 per owned tet
    std::vector<int32_t> nbr;                       // global tables kept for the ghost geometry
    std::vector<double> area, normal;
    std::vector<int> speciesN;                      // velocity nodes per species
    std::vector<char> ghostsDirty;                  // per species: peers' ghost rows are stale
};

namespace {

void check(int rc)
{
    if (rc) throw std::runtime_error(vt_last_error());
}

template <class F>
int guard(F f)
{
    try {
        f();
        return 0;
    } catch (std::exception& e) {
        vt_set_error(e.what());
        return 1;
    }
}

// per-tet array (k values per tet, global caller order) -> the slice of member r
template <class T>
std::vector<T> slice(const Group& g, int r, const T* a, int k)
{
    const auto& ow = g.owned[r];
    std::vector<T> out(std::max<size_t>(1, ow.size() * k));
    for (size_t i = 0; i < ow.size(); i++)
        for (int c = 0; c < k; c++) out[i * k + c] = a[(size_t)ow[i] * k + c];
    return out;
}
template <class T>
void scatter_back(const Group& g, int r, const std::vector<T>& part, T* a, int k)
{
    const auto& ow = g.owned[r];
    for (size_t i = 0; i < ow.size(); i++)
        for (int c = 0; c < k; c++) a[(size_t)ow[i] * k + c] = part[i * k + c];
}

void refresh_ghosts(Group& g, int sp)
{
    if (sp < 0 || sp >= (int)g.ghostsDirty.size() || !g.ghostsDirty[sp]) return;
    for (auto* c : g.m) check(vt_halo_push_current(c, sp));
    for (auto* c : g.m) check(vt_halo_barrier(c));
    for (auto* c : g.m) check(vt_sync(c));
    g.ghostsDirty[sp] = 0;
}

}  // namespace

void group_destroy(Group* g)
{
    if (!g) return;
    for (auto* c : g->m) vt_ctx_destroy(c);
    delete g;
}

int group_mesh_upload(vt_ctx* ctx, int nTets, int nGhost, const int32_t* nbr, const double* area, const double* volume,
                      const double* normal, const int32_t* entity, const int32_t* order)
{
    return guard([&] {
        Group& g = *ctx->group;
        if (nGhost != 0) throw std::invalid_argument("a device group takes the whole mesh (nGhost = 0)");
        const int R = (int)g.m.size();
        g.nTets = nTets;
        ctx->nOwned = nTets;
        g.nbr.assign(nbr, nbr + 4 * (size_t)nTets);
        g.area.assign(area, area + 4 * (size_t)nTets);
        g.normal.assign(normal, normal + 12 * (size_t)nTets);
        g.owner.assign(nTets, 0);
        g.local.assign(nTets, 0);
        g.owned.assign(R, {});
        g.ghost.assign(R, {});
        for (int r = 0; r < R; r++) {
            const long long lo = (long long)nTets * r / R, hi = (long long)nTets * (r + 1) / R;
            for (long long p = lo; p < hi; p++) {
                const int t = order ? order[p] : (int)p;
                g.owner[t] = r;
                g.local[t] = (int)g.owned[r].size();
                g.owned[r].push_back(t);
            }
        }
        // ghost rows of member r: neighbours owned elsewhere, grouped by owner, sorted by global id
        for (int r = 0; r < R; r++) {
            std::vector<int32_t> gh;
            for (int t : g.owned[r])
                for (int j = 0; j < 4; j++) {
                    const int a = nbr[4 * (size_t)t + j];
                    if (a >= 0 && g.owner[a] != r) gh.push_back(a);
                }
            std::sort(gh.begin(), gh.end(), [&](int a, int b) { return g.owner[a] != g.owner[b] ? g.owner[a] < g.owner[b] : a < b; });
            gh.erase(std::unique(gh.begin(), gh.end()), gh.end());
            g.ghost[r] = gh;
        }
        g.peers.assign(R, {});
        g.pushPeer.assign(R, {});
        g.pushRank.assign(R, {});
        g.pushRow.assign(R, {});
        for (int r = 0; r < R; r++) {
            for (int a : g.ghost[r])
                if (g.peers[r].empty() || g.peers[r].back() != g.owner[a]) g.peers[r].push_back(g.owner[a]);
            g.pushPeer[r].assign(4 * std::max<size_t>(1, g.owned[r].size()), -1);
            g.pushRank[r].assign(4 * std::max<size_t>(1, g.owned[r].size()), -1);
            g.pushRow[r].assign(4 * std::max<size_t>(1, g.owned[r].size()), -1);
        }
        // push lists: q's ghost row `idx` that r owns is written by r's step kernel
        for (int q = 0; q < R; q++)
            for (size_t idx = 0; idx < g.ghost[q].size(); idx++) {
                const int a = g.ghost[q][idx], r = g.owner[a], li = g.local[a];
                int slot = 0;
                while (slot < 4 && g.pushRank[r][4 * (size_t)li + slot] >= 0) slot++;
                if (slot == 4) throw std::runtime_error("a tet is a ghost on more than 4 members");
                const auto it = std::find(g.peers[r].begin(), g.peers[r].end(), q);
                if (it == g.peers[r].end()) throw std::runtime_error("device group: asymmetric adjacency");
                g.pushPeer[r][4 * (size_t)li + slot] = (int32_t)(it - g.peers[r].begin());
                g.pushRank[r][4 * (size_t)li + slot] = q;
                g.pushRow[r][4 * (size_t)li + slot] = (int32_t)(g.owned[q].size() + idx);
            }
        for (int r = 0; r < R; r++) {
            const auto& ow = g.owned[r];
            const size_t nO = ow.size();
            std::vector<int32_t> g2l(nTets, -1);
            for (size_t i = 0; i < nO; i++) g2l[ow[i]] = (int32_t)i;
            for (size_t i = 0; i < g.ghost[r].size(); i++) g2l[g.ghost[r][i]] = (int32_t)(nO + i);
            std::vector<int32_t> lnbr(4 * std::max<size_t>(1, nO), -1);
            for (size_t i = 0; i < nO; i++)
                for (int j = 0; j < 4; j++) {
                    const int a = nbr[4 * (size_t)ow[i] + j];
                    lnbr[4 * i + j] = a < 0 ? -1 : g2l[a];
                }
            auto ar = slice(g, r, area, 4);
            auto vo = slice(g, r, volume, 1);
            auto no = slice(g, r, normal, 12);
            auto en = slice(g, r, entity, 4);
            check(vt_mesh_upload(g.m[r], (int)nO, (int)g.ghost[r].size(), lnbr.data(), ar.data(), vo.data(), no.data(),
                                 en.data(), nullptr));
        }
    });
}

int group_species_create(vt_ctx* ctx, const int32_t n[3], const double vmin[3], const double vmax[3], double mass,
                         double charge, int* species)
{
    return guard([&] {
        Group& g = *ctx->group;
        const int R = (int)g.m.size();
        std::vector<int> ids(R);
        for (int r = 0; r < R; r++) check(vt_species_create(g.m[r], n, vmin, vmax, mass, charge, &ids[r]));
        for (int r = 1; r < R; r++)
            if (ids[r] != ids[0]) throw std::runtime_error("device group: species ids diverged");
        for (int r = 0; r < R; r++) {
            std::vector<vt_ctx*> pc;
            std::vector<int32_t> ps(g.peers[r].size(), ids[0]);
            for (int q : g.peers[r]) pc.push_back(g.m[q]);
            if (!pc.empty()) {
                check(vt_halo_attach_local(g.m[r], ids[0], r, (int)pc.size(), g.peers[r].data(), pc.data(), ps.data()));
                check(vt_halo_set_push(g.m[r], ids[0], g.pushPeer[r].data(), g.pushRow[r].data()));
            }
        }
        g.speciesN.resize(ids[0] + 1);
        g.speciesN[ids[0]] = n[0] * n[1] * n[2];
        g.ghostsDirty.resize(ids[0] + 1, 1);
        g.ghostsDirty[ids[0]] = 1;
        *species = ids[0];
    });
}

int group_species_set_params(vt_ctx* ctx, int sp, double mass, double charge)
{
    return guard([&] {
        for (auto* c : ctx->group->m) check(vt_species_set_params(c, sp, mass, charge));
    });
}

int group_species_set_face_bc(vt_ctx* ctx, int sp, const uint8_t* bcType, const uint8_t* collect, const int32_t* sourceId)
{
    return guard([&] {
        Group& g = *ctx->group;
        for (int r = 0; r < (int)g.m.size(); r++) {
            auto bc = slice(g, r, bcType, 4);
            std::vector<uint8_t> co;
            std::vector<int32_t> so;
            if (collect) co = slice(g, r, collect, 4);
            if (sourceId) so = slice(g, r, sourceId, 4);
            check(vt_species_set_face_bc(g.m[r], sp, bc.data(), collect ? co.data() : nullptr, sourceId ? so.data() : nullptr));
            // the member rebuilt its tet records: the halo push lists go back in
            if (!g.peers[r].empty()) check(vt_halo_set_push(g.m[r], sp, g.pushPeer[r].data(), g.pushRow[r].data()));
        }
    });
}

int group_species_set_source_pdfs(vt_ctx* ctx, int sp, int nSource, const double* pdf)
{
    return guard([&] {
        for (auto* c : ctx->group->m) check(vt_species_set_source_pdfs(c, sp, nSource, pdf));
    });
}

// rows [first, first+count) of the caller order, in runs that are contiguous on one member
template <class F>
void for_runs(const Group& g, int first, int count, F f)
{
    int i = 0;
    while (i < count) {
        const int t = first + i, r = g.owner[t], l0 = g.local[t];
        int len = 1;
        while (i + len < count && g.owner[t + len] == r && g.local[t + len] == l0 + len) len++;
        f(r, l0, len, i);
        i += len;
    }
}

int group_species_set_pdf(vt_ctx* ctx, int sp, int first, int count, const double* pdf)
{
    return guard([&] {
        Group& g = *ctx->group;
        if (first < 0 || count < 0 || first + count > g.nTets) throw std::out_of_range("vt_species_set_pdf: tet range");
        const size_t N = g.speciesN.at(sp);
        for_runs(g, first, count, [&](int r, int l0, int len, int i) { check(vt_species_set_pdf(g.m[r], sp, l0, len, pdf + (size_t)i * N)); });
        g.ghostsDirty[sp] = 1;
    });
}

int group_species_get_pdf(vt_ctx* ctx, int sp, int first, int count, double* pdf)
{
    return guard([&] {
        Group& g = *ctx->group;
        if (first < 0 || count < 0 || first + count > g.nTets) throw std::out_of_range("vt_species_get_pdf: tet range");
        const size_t N = g.speciesN.at(sp);
        for_runs(g, first, count, [&](int r, int l0, int len, int i) { check(vt_species_get_pdf(g.m[r], sp, l0, len, pdf + (size_t)i * N)); });
    });
}

int group_species_set_maxwell(vt_ctx* ctx, int sp, const double* physDensity, double temperature, const double mpv[3])
{
    return guard([&] {
        Group& g = *ctx->group;
        for (int r = 0; r < (int)g.m.size(); r++) {
            auto d = slice(g, r, physDensity, 1);
            check(vt_species_set_maxwell(g.m[r], sp, d.data(), temperature, mpv));
        }
        g.ghostsDirty[sp] = 1;
    });
}

int group_species_density(vt_ctx* ctx, int sp, double* density)
{
    return guard([&] {
        Group& g = *ctx->group;
        for (int r = 0; r < (int)g.m.size(); r++) {
            std::vector<double> d(std::max<size_t>(1, g.owned[r].size()));
            check(vt_species_density(g.m[r], sp, density ? d.data() : nullptr));
            if (density) scatter_back(g, r, d, density, 1);
        }
    });
}

int group_species_velocity(vt_ctx* ctx, int sp, double* velocity)
{
    return guard([&] {
        Group& g = *ctx->group;
        for (int r = 0; r < (int)g.m.size(); r++) {
            std::vector<double> v(3 * std::max<size_t>(1, g.owned[r].size()));
            check(vt_species_velocity(g.m[r], sp, v.data()));
            scatter_back(g, r, v, velocity, 3);
        }
    });
}

int group_field_set(vt_ctx* ctx, const double* E)
{
    return guard([&] {
        Group& g = *ctx->group;
        for (int r = 0; r < (int)g.m.size(); r++) {
            auto e = slice(g, r, E, 3);
            check(vt_field_set(g.m[r], e.data()));
        }
    });
}

int group_field_get(vt_ctx* ctx, double* rho, double* phi, double* E)
{
    return guard([&] {
        Group& g = *ctx->group;
        for (int r = 0; r < (int)g.m.size(); r++) {
            const size_t nO = std::max<size_t>(1, g.owned[r].size());
            std::vector<double> a(nO), b(nO), c(3 * nO);
            check(vt_field_get(g.m[r], rho ? a.data() : nullptr, phi ? b.data() : nullptr, E ? c.data() : nullptr));
            if (rho) scatter_back(g, r, a, rho, 1);
            if (phi) scatter_back(g, r, b, phi, 1);
            if (E) scatter_back(g, r, c, E, 3);
        }
    });
}

int group_step(vt_ctx* ctx, int sp, double dt, const double ext[3], bool tucker)
{
    return guard([&] {
        Group& g = *ctx->group;
        refresh_ghosts(g, sp);
        // every member's step is in flight before any barrier is queued: nothing a member waits for on
        // the device may depend on a later host call that could block (allocation, module load)
        for (auto* c : g.m) check(tucker ? vt_step_tucker(c, sp, dt, ext) : vt_step_full(c, sp, dt, ext));
        for (auto* c : g.m) check(vt_halo_barrier(c));
    });
}

int group_wall_charge_get(vt_ctx* ctx, int sp, int entity, double* charge)
{
    return guard([&] {
        double tot = 0.0;
        for (auto* c : ctx->group->m) {   // fixed member order: deterministic sum
            double q = 0.0;
            check(vt_wall_charge_get(c, sp, entity, &q));
            tot += q;
        }
        *charge = tot;
    });
}

int group_wall_charge_reset(vt_ctx* ctx, int sp)
{
    return guard([&] {
        for (auto* c : ctx->group->m) check(vt_wall_charge_reset(c, sp));
    });
}

int group_charge_density(vt_ctx* ctx, const int* species, int nSpecies, const double* background)
{
    return guard([&] {
        Group& g = *ctx->group;
        for (int r = 0; r < (int)g.m.size(); r++) {
            std::vector<double> bg;
            if (background) bg = slice(g, r, background, 1);
            check(vt_charge_density(g.m[r], species, nSpecies, background ? bg.data() : nullptr));
        }
    });
}

int group_poisson_setup(vt_ctx* ctx, const double* tetCentroid, const double* faceCentroid, const uint8_t* bcType,
                        const double* bcValue, const double* bcNormalGrad)
{
    return guard([&] {
        Group& g = *ctx->group;
        const int R = (int)g.m.size();
        bool anyDirichlet = false;
        for (size_t i = 0; i < 4 * (size_t)g.nTets; i++)
            if (bcType[i] == VT_QBC_DIRICHLET) anyDirichlet = true;
        for (int r = 0; r < R; r++) {
            const auto& ow = g.owned[r];
            const auto& gh = g.ghost[r];
            const size_t nO = ow.size(), nG = gh.size();
            std::vector<int32_t> g2l(g.nTets, -1), gid(std::max<size_t>(1, nO + nG));
            for (size_t i = 0; i < nO; i++) {
                g2l[ow[i]] = (int32_t)i;
                gid[i] = ow[i];
            }
            for (size_t i = 0; i < nG; i++) {
                g2l[gh[i]] = (int32_t)(nO + i);
                gid[nO + i] = gh[i];
            }
            std::vector<int32_t> gn(4 * std::max<size_t>(1, nG), -1);
            std::vector<double> ga(4 * std::max<size_t>(1, nG)), gnr(12 * std::max<size_t>(1, nG)), gc(3 * std::max<size_t>(1, nG)),
                gfc(12 * std::max<size_t>(1, nG));
            for (size_t i = 0; i < nG; i++) {
                const size_t a = gh[i];
                for (int j = 0; j < 4; j++) {
                    const int b = g.nbr[4 * a + j];
                    gn[4 * i + j] = b < 0 ? -1 : g2l[b];
                    ga[4 * i + j] = g.area[4 * a + j];
                }
                for (int c = 0; c < 12; c++) {
                    gnr[12 * i + c] = g.normal[12 * a + c];
                    gfc[12 * i + c] = faceCentroid[12 * a + c];
                }
                for (int c = 0; c < 3; c++) gc[3 * i + c] = tetCentroid[3 * a + c];
            }
            check(vt_mesh_set_ghost_geometry(g.m[r], g.nTets, gid.data(), gn.data(), ga.data(), gnr.data(), gc.data(), gfc.data()));
            check(vt_poisson_set_global_dirichlet(g.m[r], anyDirichlet ? 1 : 0));
            auto tc = slice(g, r, tetCentroid, 3);
            auto fc = slice(g, r, faceCentroid, 12);
            auto bt = slice(g, r, bcType, 4);
            auto bv = slice(g, r, bcValue, 4);
            auto bg = slice(g, r, bcNormalGrad, 4);
            check(vt_poisson_setup(g.m[r], tc.data(), fc.data(), bt.data(), bv.data(), bg.data()));
        }
        for (int r = 0; r < R; r++) {
            check(vt_poisson_comm_attach_local(g.m[r], r, R, g.m.data()));
            check(vt_poisson_set_push(g.m[r], g.pushRank[r].data(), g.pushRow[r].data()));
        }
    });
}

int group_poisson_update_bc_values(vt_ctx* ctx, const double* bcValue, const double* bcNormalGrad)
{
    return guard([&] {
        Group& g = *ctx->group;
        for (int r = 0; r < (int)g.m.size(); r++) {
            auto bv = slice(g, r, bcValue, 4);
            auto bg = slice(g, r, bcNormalGrad, 4);
            check(vt_poisson_update_bc_values(g.m[r], bv.data(), bg.data()));
        }
    });
}

int group_poisson_solve(vt_ctx* ctx, const double* rho, double* phi, double* E)
{
    return guard([&] {
        Group& g = *ctx->group;
        // all members' solves are launched before anything is read back (they wait for each other on the device)
        for (int r = 0; r < (int)g.m.size(); r++) {
            std::vector<double> rs;
            if (rho) rs = slice(g, r, rho, 1);
            check(vt_poisson_solve(g.m[r], rho ? rs.data() : nullptr, nullptr, nullptr));
        }
        if (phi || E) check(group_field_get(ctx, nullptr, phi, E));
    });
}

int group_poisson_stats(vt_ctx* ctx, int* its, double* res) { return vt_poisson_stats(ctx->group->m[0], its, res); }

int group_tucker_enable(vt_ctx* ctx, int sp, double comprErr, int maxRank)
{
    return guard([&] {
        Group& g = *ctx->group;
        const int R = (int)g.m.size();
        for (auto* c : g.m) check(vt_tucker_enable(c, sp, comprErr, maxRank));
        for (int r = 0; r < R; r++) {
            std::vector<vt_ctx*> pc;
            std::vector<int32_t> ps(g.peers[r].size(), sp);
            for (int q : g.peers[r]) pc.push_back(g.m[q]);
            if (!pc.empty()) check(vt_tucker_halo_attach_local(g.m[r], sp, (int)pc.size(), pc.data(), ps.data()));
        }
        g.ghostsDirty[sp] = 1;
    });
}

int group_tucker_get_factors(vt_ctx* ctx, int sp, int tet, int32_t ranks[3], double* core, double* u0, double* u1, double* u2)
{
    Group& g = *ctx->group;
    if (tet < 0 || tet >= g.nTets) {
        vt_set_error("tet index");
        return 1;
    }
    return vt_tucker_get_factors(g.m[g.owner[tet]], sp, g.local[tet], ranks, core, u0, u1, u2);
}

int group_tucker_get_ranks(vt_ctx* ctx, int sp, int32_t* ranks)
{
    return guard([&] {
        Group& g = *ctx->group;
        for (int r = 0; r < (int)g.m.size(); r++) {
            std::vector<int32_t> p(3 * std::max<size_t>(1, g.owned[r].size()));
            check(vt_tucker_get_ranks(g.m[r], sp, p.data()));
            scatter_back(g, r, p, ranks, 3);
        }
    });
}

int group_sync(vt_ctx* ctx)
{
    return guard([&] {
        for (auto* c : ctx->group->m) check(vt_sync(c));
    });
}

long group_launch_count(vt_ctx* ctx)
{
    long n = 0;
    for (auto* c : ctx->group->m) n += vt_launch_count(c);
    return n;
}

}  // namespace vt

extern "C" int vt_ctx_create_group(const int* devices, int nDevices, vt_ctx** out)
{
    using namespace vt;
    return guard([&] {
        if (nDevices < 1 || nDevices > 64) throw std::invalid_argument("vt_ctx_create_group: 1..64 devices");
        Group* g = new Group();
        try {
            for (int i = 0; i < nDevices; i++) {
                vt_ctx* c = nullptr;
                check(vt_ctx_create(devices[i], &c));
                g->m.push_back(c);
            }
        } catch (...) {
            group_destroy(g);
            throw;
        }
        vt_ctx* ctx = new vt_ctx();
        ctx->device = devices[0];
        ctx->prop = g->m[0]->prop;
        ctx->group = g;
        *out = ctx;
    });
}

extern "C" int vt_group_size(vt_ctx* ctx) { return ctx->group ? (int)ctx->group->m.size() : 1; }
