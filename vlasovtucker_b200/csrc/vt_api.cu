// C ABI of libvt_b200.so (include/vt_b200.h): context, mesh tables, species state.
#include "vt_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace {
thread_local std::string g_err;

template <class F>
int guard(F f)
{
    try {
        f();
        return 0;
    } catch (std::exception& e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown error";
        return 1;
    }
}

// rows of N doubles: dst[rowMap[i]] = src[i]  (scatter)  or dst[i] = src[rowMap[i]] (gather)
__global__ void k_rows_scatter(const double* __restrict__ src, double* __restrict__ dst,
                               const int32_t* __restrict__ rowMap, int first, int N)
{
    const int i = blockIdx.x;
    const double* s = src + (size_t)i * N;
    double* d = dst + (size_t)rowMap[first + i] * N;
    for (int e = threadIdx.x; e < N; e += blockDim.x) d[e] = s[e];
}
__global__ void k_rows_gather(const double* __restrict__ src, double* __restrict__ dst,
                              const int32_t* __restrict__ rowMap, int first, int N)
{
    const int i = blockIdx.x;
    const double* s = src + (size_t)rowMap[first + i] * N;
    double* d = dst + (size_t)i * N;
    for (int e = threadIdx.x; e < N; e += blockDim.x) d[e] = s[e];
}
// f[row][e] = table[e] * scale[row]   (particle_data.cpp:57-66)
__global__ void k_maxwell_fill(double* __restrict__ f, const double* __restrict__ table,
                               const double* __restrict__ scale, int N)
{
    const int row = blockIdx.x;
    const double s = scale[row];
    double* d = f + (size_t)row * N;
    for (int e = threadIdx.x; e < N; e += blockDim.x) d[e] = table[e] * s;
}
// f[row][e] = sum_k amp[row][k] * a0[k][i0] * a1[k][i1] * a2[k][i2]: a sum of nTerms separable terms
// (shifted / anisotropic Maxwellians of the C5 inputs: multilinear rank nTerms exactly)
__global__ void k_separable_fill(double* __restrict__ f, const double* __restrict__ amp, const double* __restrict__ a0,
                                 const double* __restrict__ a1, const double* __restrict__ a2, int nTerms, int n0, int n1,
                                 int n2)
{
    const int row = blockIdx.x;
    const int N = n0 * n1 * n2;
    double* d = f + (size_t)row * N;
    for (int e = threadIdx.x; e < N; e += blockDim.x) {
        const int i0 = e % n0, i1 = (e / n0) % n1, i2 = e / (n0 * n1);
        double s = 0.0;
        for (int k = 0; k < nTerms; k++)
            s += amp[(size_t)row * nTerms + k] * ((a0[k * n0 + i0] * a1[k * n1 + i1]) * a2[k * n2 + i2]);
        d[e] = s;
    }
}
// FP64 FMA peak probe: 8 independent chains per thread
__global__ void k_dfma_probe(double* out, int iters)
{
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// u_k = sum_v v_k f * cellVolume / density   (particle_data.cpp:105-125)
__global__ void k_velocity(const double* __restrict__ f, const double* __restrict__ density,
                           double* __restrict__ vel, int n0, int n1, int n2, double m0, double m1,
                           double m2, double s0, double s1, double s2, double cellVolume)
{
    __shared__ double red[3][32];
    const int N = n0 * n1 * n2;
    const double* row = f + (size_t)blockIdx.x * N;
    double a0 = 0, a1 = 0, a2 = 0;
    for (int e = threadIdx.x; e < N; e += blockDim.x) {
        int i0 = e % n0, i1 = (e / n0) % n1, i2 = e / (n0 * n1);
        double x = row[e];
        a0 += __dadd_rn(m0, __dmul_rn((double)i0, s0)) * x;
        a1 += __dadd_rn(m1, __dmul_rn((double)i1, s1)) * x;
        a2 += __dadd_rn(m2, __dmul_rn((double)i2, s2)) * x;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = a0;
        red[1][threadIdx.x >> 5] = a1;
        red[2][threadIdx.x >> 5] = a2;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[threadIdx.x][w];
        double d = density[blockIdx.x];
        vel[3 * (size_t)blockIdx.x + threadIdx.x] = d != 0 ? s * cellVolume / d : 0.0;
    }
}
// rho[t] (+)= charge * density[t] (+ background[t])
__global__ void k_charge_accum(double* __restrict__ rho, const double* __restrict__ density, double charge,
                               int n, int first)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double v = charge * density[t];
    rho[t] = first ? v : rho[t] + v;
}
__global__ void k_add(double* __restrict__ a, const double* __restrict__ b, int n)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) a[t] += b[t];
}

vt::Species& species_of(vt_ctx* ctx, int s)
{
    if (s < 0 || s >= (int)ctx->species.size()) throw std::invalid_argument("bad species id");
    return *ctx->species[s];
}

// caller-order host vector (k doubles per tet) -> device-order device array
void upload_tet_array(vt_ctx* ctx, const double* host, double* dev, int k)
{
    const int n = ctx->nOwned;
    double* pin = vt::ctx_pinned(ctx, (size_t)n * k * sizeof(double));
    for (int p = 0; p < n; p++) {
        const int t = ctx->order[p];
        for (int j = 0; j < k; j++) pin[(size_t)p * k + j] = host[(size_t)t * k + j];
    }
    VT_CUDA(cudaMemcpyAsync(dev, pin, (size_t)n * k * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    VT_CUDA(cudaStreamSynchronize(ctx->stream));
}
void download_tet_array(vt_ctx* ctx, const double* dev, double* host, int k)
{
    const int n = ctx->nOwned;
    double* pin = vt::ctx_pinned(ctx, (size_t)n * k * sizeof(double));
    VT_CUDA(cudaMemcpyAsync(pin, dev, (size_t)n * k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VT_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < n; p++) {
        const int t = ctx->order[p];
        for (int j = 0; j < k; j++) host[(size_t)t * k + j] = pin[(size_t)p * k + j];
    }
}

void rebuild_tet_records(vt_ctx* ctx, vt::Species& sp)
{
    const int n = ctx->nOwned;
    // absorbed charge collected so far, by entity: a change of the boundary conditions (or of the halo
    // push lists) renumbers the accumulator slots but must not lose the charge — the reference keeps
    // _wallCharge across SetParticleBC and across Solve() calls (solver.cpp:296-311 only adds entities)
    std::map<int, double> kept;
    if (sp.wall && !sp.wallEntities.empty()) {
        std::vector<double> old(sp.wallEntities.size());
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
        VT_CUDA(cudaMemcpy(old.data(), sp.wall, old.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < old.size(); i++) kept[sp.wallEntities[i]] = old[i];
    }
    sp.recHost.assign(n, vt::TetRec());
    sp.wallEntities.clear();
    sp.danglingFaces = 0;
    sp.fastOnly = -1;
    std::map<int, int> slotOf;
    for (int p = 0; p < n; p++) {
        const int t = ctx->order[p];
        vt::TetRec& r = sp.recHost[p];
        for (int j = 0; j < 4; j++) {
            const size_t fi = 4 * (size_t)t + j;
            r.coef[j] = ctx->area[fi] / ctx->volume[t];  // solver.cpp:168
            r.area[j] = ctx->area[fi];
            for (int k = 0; k < 3; k++) r.nrm[j][k] = ctx->normal[3 * fi + k];
            const int bc = sp.bcType.empty() ? VT_PBC_NONBOUNDARY : sp.bcType[fi];
            r.bc[j] = (uint8_t)bc;
            r.wallSlot[j] = -1;
            r.pushPeer[j] = sp.pushPeer.empty() ? -1 : sp.pushPeer[fi];
            r.pushRow[j] = sp.pushRow.empty() ? -1 : sp.pushRow[fi];
            r.nbr[j] = ctx->nbrHost[4 * (size_t)p + j];
            if (r.nbr[j] >= ctx->nOwned) r.ghostFaces |= 1 << j;
            if (bc == VT_PBC_SOURCE) {
                const int sid = sp.sourceId.empty() ? -1 : sp.sourceId[fi];
                if (sid < 0 || sid >= sp.nSrc) throw std::invalid_argument("Source face without a source PDF");
                r.nbr[j] = -2 - sid;
            } else if (bc == VT_PBC_ABSORBING || bc == VT_PBC_FREE) {
                r.nbr[j] = -1;
                if (bc == VT_PBC_ABSORBING && !sp.collect.empty() && sp.collect[fi]) {
                    const int ent = ctx->entity[fi];
                    auto it = slotOf.find(ent);
                    if (it == slotOf.end()) {
                        it = slotOf.emplace(ent, (int)sp.wallEntities.size()).first;
                        sp.wallEntities.push_back(ent);
                    }
                    if (it->second > 127) throw std::runtime_error("too many charge-collecting entities");
                    r.wallSlot[j] = (int8_t)it->second;
                }
            } else if (r.nbr[j] < 0) {
                // solver.cpp:319 dereferences adjTets[f] unconditionally: a NonBoundary/Periodic
                // face without a neighbour is a null dereference in the reference
                // here; checked when a step is requested so BCs can be set after creation
                sp.danglingFaces++;
            }
        }
        // halo push slots are not tied to faces: pack the used ones to the front
        int used = 0;
        for (int j = 0; j < 4; j++)
            if (r.pushPeer[j] >= 0) {
                const int32_t pp = r.pushPeer[j], pr = r.pushRow[j];
                r.pushPeer[j] = -1;
                r.pushRow[j] = -1;
                r.pushPeer[used] = pp;
                r.pushRow[used] = pr;
                used++;
            }
    }
    if (!sp.rec) VT_CUDA(cudaMalloc(&sp.rec, std::max<size_t>(1, n) * sizeof(vt::TetRec)));
    VT_CUDA(cudaMemcpy(sp.rec, sp.recHost.data(), (size_t)n * sizeof(vt::TetRec), cudaMemcpyHostToDevice));
    // new accumulator slots; entities that keep a slot keep their charge (vt_wall_charge_reset zeroes them)
    if (sp.wall) VT_CUDA(cudaFree(sp.wall));
    sp.wall = nullptr;
    std::vector<double> init(std::max<size_t>(1, sp.wallEntities.size()), 0.0);
    for (size_t i = 0; i < sp.wallEntities.size(); i++) {
        auto it = kept.find(sp.wallEntities[i]);
        if (it != kept.end()) init[i] = it->second;
    }
    VT_CUDA(cudaMalloc(&sp.wall, init.size() * sizeof(double)));
    VT_CUDA(cudaMemcpy(sp.wall, init.data(), init.size() * sizeof(double), cudaMemcpyHostToDevice));
}
}  // namespace

namespace vt {
void rebuild_tet_records_public(vt_ctx* ctx, Species& sp) { rebuild_tet_records(ctx, sp); }

double* ctx_stage(vt_ctx* ctx, size_t bytes)
{
    if (ctx->stageBytes < bytes) {
        if (ctx->stage) VT_CUDA(cudaFree(ctx->stage));
        ctx->stage = nullptr;
        VT_CUDA(cudaMalloc(&ctx->stage, bytes));
        ctx->stageBytes = bytes;
    }
    return ctx->stage;
}
double* ctx_pinned(vt_ctx* ctx, size_t bytes)
{
    if (ctx->pinnedBytes < bytes) {
        if (ctx->pinned) VT_CUDA(cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr;
        VT_CUDA(cudaMallocHost(&ctx->pinned, bytes));
        ctx->pinnedBytes = bytes;
    }
    return ctx->pinned;
}
}  // namespace vt

extern "C" {

const char* vt_last_error(void) { return g_err.c_str(); }
void vt_set_error(const char* msg) { g_err = msg; }
int vt_version(void) { return 100; }

int vt_ctx_create(int device, vt_ctx** out)
{
    return guard([&] {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw std::runtime_error(std::string("no CUDA device: libvt_b200 has no CPU fallback (") +
                                     cudaGetErrorString(e) + ")");
        if (device < 0 || device >= count) throw std::invalid_argument("bad device index");
        VT_CUDA(cudaSetDevice(device));
        vt_ctx* c = new vt_ctx();
        c->device = device;
        VT_CUDA(cudaGetDeviceProperties(&c->prop, device));
        VT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        VT_CUDA(cudaEventCreate(&c->ev0));
        VT_CUDA(cudaEventCreate(&c->ev1));
        *out = c;
    });
}

void vt_ctx_destroy(vt_ctx* ctx)
{
    if (!ctx) return;
    if (ctx->group) {
        vt::group_destroy(ctx->group);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto* sp : ctx->species) {
        cudaFree(sp->f[0]);
        cudaFree(sp->f[1]);
        cudaFree(sp->rec);
        cudaFree(sp->src);
        cudaFree(sp->density);
        cudaFree(sp->densPartial);
        cudaFree(sp->wall);
        cudaFree(sp->tetLists);
        vt::tucker_destroy(sp->tucker);
        delete sp;
    }
    for (void* p : ctx->ipcOpened) cudaIpcCloseMemHandle(p);
    cudaFree(ctx->flags);
    if (ctx->haloStatusHost) cudaFreeHost(ctx->haloStatusHost);
    cudaFree(ctx->haloTable);
    cudaFree(ctx->workCounter);
    if (ctx->poisson) vt::poisson_destroy(ctx->poisson);
    cudaFree(ctx->E);
    cudaFree(ctx->rho);
    cudaFree(ctx->phi);
    cudaFree(ctx->stage);
    cudaFree(ctx->orderDev);
    cudaFree(ctx->invDev);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int vt_sync(vt_ctx* ctx)
{
    if (ctx->group) return vt::group_sync(ctx);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
        vt::check_halo_status(ctx);
    });
}

int vt_device_info(vt_ctx* ctx, int* sm_count, size_t* l2_bytes, size_t* hbm_bytes)
{
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (l2_bytes) *l2_bytes = (size_t)ctx->prop.l2CacheSize;
    if (hbm_bytes) *hbm_bytes = ctx->prop.totalGlobalMem;
    return 0;
}

int vt_mesh_upload(vt_ctx* ctx, int nOwned, int nGhost, const int32_t* nbr, const double* area,
                   const double* volume, const double* normal, const int32_t* entity, const int32_t* order)
{
    if (ctx->group) return vt::group_mesh_upload(ctx, nOwned, nGhost, nbr, area, volume, normal, entity, order);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (nOwned < 0 || nGhost < 0) throw std::invalid_argument("negative tet count");
        if (!ctx->species.empty()) throw std::runtime_error("vt_mesh_upload must precede vt_species_create");
        if (ctx->orderDev) throw std::runtime_error("vt_mesh_upload: this context already holds a mesh (create a new context)");
        ctx->nOwned = nOwned;
        ctx->nGhost = nGhost;
        ctx->order.resize(nOwned);
        ctx->inv.assign(nOwned, -1);
        ctx->identityOrder = true;
        for (int p = 0; p < nOwned; p++) {
            const int t = order ? order[p] : p;
            if (t < 0 || t >= nOwned || ctx->inv[t] != -1) throw std::invalid_argument("order is not a permutation");
            ctx->order[p] = t;
            ctx->inv[t] = p;
            if (t != p) ctx->identityOrder = false;
        }
        ctx->area.assign(area, area + 4 * (size_t)nOwned);
        ctx->volume.assign(volume, volume + nOwned);
        ctx->normal.assign(normal, normal + 12 * (size_t)nOwned);
        ctx->entity.assign(entity, entity + 4 * (size_t)nOwned);
        ctx->nbrHost.resize(4 * (size_t)nOwned);
        for (int p = 0; p < nOwned; p++)
            for (int j = 0; j < 4; j++) {
                int a = nbr[4 * (size_t)ctx->order[p] + j];
                if (a >= nOwned + nGhost) throw std::invalid_argument("neighbour index out of range");
                ctx->nbrHost[4 * (size_t)p + j] = a < 0 ? -1 : (a < nOwned ? ctx->inv[a] : a);
            }
        const size_t nAlloc = std::max(1, nOwned);
        VT_CUDA(cudaMalloc(&ctx->orderDev, nAlloc * sizeof(int32_t)));
        VT_CUDA(cudaMalloc(&ctx->invDev, nAlloc * sizeof(int32_t)));
        VT_CUDA(cudaMemcpy(ctx->orderDev, ctx->order.data(), nOwned * sizeof(int32_t), cudaMemcpyHostToDevice));
        VT_CUDA(cudaMemcpy(ctx->invDev, ctx->inv.data(), nOwned * sizeof(int32_t), cudaMemcpyHostToDevice));
        VT_CUDA(cudaMalloc(&ctx->E, 3 * nAlloc * sizeof(double)));
        VT_CUDA(cudaMemset(ctx->E, 0, 3 * nAlloc * sizeof(double)));
        VT_CUDA(cudaMalloc(&ctx->rho, nAlloc * sizeof(double)));
        VT_CUDA(cudaMalloc(&ctx->phi, nAlloc * sizeof(double)));
        VT_CUDA(cudaMemset(ctx->rho, 0, nAlloc * sizeof(double)));
        VT_CUDA(cudaMemset(ctx->phi, 0, nAlloc * sizeof(double)));
    });
}

int vt_mesh_set_ghost_geometry(vt_ctx* ctx, int globalTets, const int32_t* globalId, const int32_t* ghostNbr,
                               const double* ghostArea, const double* ghostNormal, const double* ghostTetCentroid,
                               const double* ghostFaceCentroid)
{
    if (ctx->group) { vt_set_error("vt_mesh_set_ghost_geometry: not available on a device group (call it on a single-device context)"); return 1; };
    return guard([&] {
        const size_t nO = ctx->nOwned, nG = ctx->nGhost;
        if (globalTets < (int)nO) throw std::invalid_argument("vt_mesh_set_ghost_geometry: globalTets < nOwned");
        ctx->globalRows = globalTets;
        ctx->globalId.assign(globalId, globalId + nO + nG);
        ctx->ghostNbr.assign(ghostNbr, ghostNbr + 4 * nG);
        for (size_t i = 0; i < 4 * nG; i++)
            if (ctx->ghostNbr[i] >= (int)(nO + nG)) throw std::invalid_argument("ghost neighbour index out of range");
        ctx->ghostArea.assign(ghostArea, ghostArea + 4 * nG);
        ctx->ghostNormal.assign(ghostNormal, ghostNormal + 12 * nG);
        ctx->ghostCentroid.assign(ghostTetCentroid, ghostTetCentroid + 3 * nG);
        ctx->ghostFaceCentroid.assign(ghostFaceCentroid, ghostFaceCentroid + 12 * nG);
    });
}

int vt_species_create(vt_ctx* ctx, const int32_t n[3], const double vmin[3], const double vmax[3], double mass,
                      double charge, int* species)
{
    if (ctx->group) return vt::group_species_create(ctx, n, vmin, vmax, mass, charge, species);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (n[0] < 2 || n[1] < 2 || n[2] < 2) throw std::invalid_argument("velocity grid needs >= 2 nodes per axis");
        vt::Species* sp = new vt::Species();
        for (int k = 0; k < 3; k++) {
            sp->n[k] = n[k];
            sp->vmin[k] = vmin[k];
            sp->vmax[k] = vmax[k];
            sp->step[k] = (vmax[k] - vmin[k]) / (n[k] - 1);  // velocity_grid.cpp:15
        }
        sp->N = n[0] * n[1] * n[2];
        sp->cellVolume = sp->step[0] * sp->step[1] * sp->step[2];  // velocity_grid.cpp:17
        sp->mass = mass;
        sp->charge = charge;
        const size_t rows = std::max(1, ctx->nOwned + ctx->nGhost);
        for (int b = 0; b < 2; b++) {
            VT_CUDA(cudaMalloc(&sp->f[b], rows * sp->N * sizeof(double)));
            VT_CUDA(cudaMemsetAsync(sp->f[b], 0, rows * sp->N * sizeof(double), ctx->stream));
        }
        VT_CUDA(cudaMalloc(&sp->density, std::max(1, ctx->nOwned) * sizeof(double)));
        // the buffers may be exported to peer GPUs right away: nothing of ours may still be pending
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->species.push_back(sp);
        rebuild_tet_records(ctx, *sp);
        *species = (int)ctx->species.size() - 1;
    });
}

int vt_species_set_params(vt_ctx* ctx, int species, double mass, double charge)
{
    if (ctx->group) return vt::group_species_set_params(ctx, species, mass, charge);
    return guard([&] {
        vt::Species& sp = species_of(ctx, species);
        sp.mass = mass;
        sp.charge = charge;
    });
}

int vt_species_set_face_bc(vt_ctx* ctx, int species, const uint8_t* bcType, const uint8_t* collect,
                           const int32_t* sourceId)
{
    if (ctx->group) return vt::group_species_set_face_bc(ctx, species, bcType, collect, sourceId);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        const size_t nf = 4 * (size_t)ctx->nOwned;
        sp.bcType.assign(bcType, bcType + nf);
        if (collect) sp.collect.assign(collect, collect + nf);
        else sp.collect.assign(nf, 0);
        if (sourceId) sp.sourceId.assign(sourceId, sourceId + nf);
        else sp.sourceId.assign(nf, -1);
        rebuild_tet_records(ctx, sp);
    });
}

int vt_species_set_source_pdfs(vt_ctx* ctx, int species, int nSource, const double* pdf)
{
    if (ctx->group) return vt::group_species_set_source_pdfs(ctx, species, nSource, pdf);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        if (sp.src) VT_CUDA(cudaFree(sp.src));
        sp.src = nullptr;
        sp.nSrc = nSource;
        if (nSource > 0) {
            VT_CUDA(cudaMalloc(&sp.src, (size_t)nSource * sp.N * sizeof(double)));
            VT_CUDA(cudaMemcpy(sp.src, pdf, (size_t)nSource * sp.N * sizeof(double), cudaMemcpyHostToDevice));
        }
    });
}

int vt_species_set_pdf(vt_ctx* ctx, int species, int first, int count, const double* pdf)
{
    if (ctx->group) return vt::group_species_set_pdf(ctx, species, first, count, pdf);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        if (first < 0 || count < 0 || first + count > ctx->nOwned + ctx->nGhost)
            throw std::out_of_range("vt_species_set_pdf: tet range");
        if (sp.tucker) vt::tucker_begin_dense_write(ctx, sp);   // partial ranges keep the other rows
        double* dst = sp.f[sp.cur];
        const size_t rowB = (size_t)sp.N * sizeof(double);
        // ghost rows and identity order: straight copies
        if (ctx->identityOrder || first >= ctx->nOwned) {
            VT_CUDA(cudaMemcpyAsync(dst + (size_t)first * sp.N, pdf, count * rowB, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            if (first + count > ctx->nOwned) throw std::out_of_range("range straddles owned and ghost rows");
            const int batch = (int)std::max<size_t>(1, std::min<size_t>(count, (256u << 20) / rowB));
            double* stage = vt::ctx_stage(ctx, batch * rowB);
            for (int done = 0; done < count; done += batch) {
                const int nb = std::min(batch, count - done);
                VT_CUDA(cudaMemcpyAsync(stage, pdf + (size_t)done * sp.N, nb * rowB, cudaMemcpyHostToDevice, ctx->stream));
                k_rows_scatter<<<nb, 256, 0, ctx->stream>>>(stage, dst, ctx->invDev, first + done, sp.N);
                ctx->launches++;
                VT_CUDA(cudaGetLastError());
            }
        }
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
        sp.densityValid = false;
        if (sp.tucker) vt::tucker_end_dense_write(ctx, sp);
    });
}

int vt_species_get_pdf(vt_ctx* ctx, int species, int first, int count, double* pdf)
{
    if (ctx->group) return vt::group_species_get_pdf(ctx, species, first, count, pdf);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        if (first < 0 || count < 0 || first + count > ctx->nOwned + ctx->nGhost)
            throw std::out_of_range("vt_species_get_pdf: tet range");
        if (sp.tucker) vt::tucker_materialize(ctx, sp);
        const double* src = sp.f[sp.cur];
        const size_t rowB = (size_t)sp.N * sizeof(double);
        if (ctx->identityOrder || first >= ctx->nOwned) {
            VT_CUDA(cudaMemcpyAsync(pdf, src + (size_t)first * sp.N, count * rowB, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            if (first + count > ctx->nOwned) throw std::out_of_range("range straddles owned and ghost rows");
            const int batch = (int)std::max<size_t>(1, std::min<size_t>(count, (256u << 20) / rowB));
            double* stage = vt::ctx_stage(ctx, batch * rowB);
            for (int done = 0; done < count; done += batch) {
                const int nb = std::min(batch, count - done);
                k_rows_gather<<<nb, 256, 0, ctx->stream>>>(src, stage, ctx->invDev, first + done, sp.N);
                ctx->launches++;
                VT_CUDA(cudaGetLastError());
                VT_CUDA(cudaMemcpyAsync(pdf + (size_t)done * sp.N, stage, nb * rowB, cudaMemcpyDeviceToHost, ctx->stream));
            }
        }
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int vt_species_set_maxwell(vt_ctx* ctx, int species, const double* physDensity, double temperature,
                           const double mpv[3])
{
    if (ctx->group) return vt::group_species_set_maxwell(ctx, species, physDensity, temperature, mpv);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        const double boltzConst = 1.38e-23;  // constants.h:12
        const int n0 = sp.n[0], n1 = sp.n[1], n2 = sp.n[2];
        std::vector<double> table(sp.N, 0.0);
        double normConst = 0;
        if (temperature != 0.0) {
            // particle_data.cpp:38-56, same loop nest so normConst sums in the same order
            for (int i0 = 0; i0 < n0; i0++)
                for (int i1 = 0; i1 < n1; i1++)
                    for (int i2 = 0; i2 < n2; i2++) {
                        const double at[3] = {sp.vmin[0] + i0 * sp.step[0], sp.vmin[1] + i1 * sp.step[1],
                                              sp.vmin[2] + i2 * sp.step[2]};
                        double velSquared = 0;
                        for (int j = 0; j < 3; j++) {
                            const double velJ = at[j] - mpv[j];
                            velSquared += velJ * velJ;
                        }
                        const double e = std::exp(-sp.mass * velSquared / (2 * boltzConst * temperature));
                        table[i0 + n0 * (i1 + n1 * i2)] = e;
                        normConst += e;
                    }
        } else {
            // particle_data.cpp:70-77
            const int i0 = (int)((mpv[0] - sp.vmin[0]) / sp.step[0]);
            const int i1 = (int)((mpv[1] - sp.vmin[1]) / sp.step[1]);
            const int i2 = (int)((mpv[2] - sp.vmin[2]) / sp.step[2]);
            if (i0 < 0 || i0 >= n0 || i1 < 0 || i1 >= n1 || i2 < 0 || i2 >= n2)
                throw std::out_of_range("mostProbableV outside the velocity grid");
            table[i0 + n0 * (i1 + n1 * i2)] = 1.0;
        }
        const int n = ctx->nOwned;
        std::vector<double> scale(n);
        for (int p = 0; p < n; p++) {
            const double d = physDensity[ctx->order[p]];
            scale[p] = temperature != 0.0 ? d / (sp.cellVolume * normConst) : d / sp.cellVolume;
        }
        double* tableDev = vt::ctx_stage(ctx, ((size_t)sp.N + n) * sizeof(double));
        double* scaleDev = tableDev + sp.N;
        VT_CUDA(cudaMemcpyAsync(tableDev, table.data(), (size_t)sp.N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        VT_CUDA(cudaMemcpyAsync(scaleDev, scale.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (n > 0) {
            k_maxwell_fill<<<n, 256, 0, ctx->stream>>>(sp.f[sp.cur], tableDev, scaleDev, sp.N);
            ctx->launches++;
            VT_CUDA(cudaGetLastError());
        }
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
        sp.densityValid = false;
        if (sp.tucker) vt::tucker_from_dense(ctx, sp);
    });
}

int vt_species_set_separable(vt_ctx* ctx, int species, int nTerms, const double* amp, const double* a0,
                             const double* a1, const double* a2)
{
    if (ctx->group) { vt_set_error("vt_species_set_separable: not available on a device group (call it on a single-device context)"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        if (nTerms < 1 || nTerms > 64) throw std::invalid_argument("vt_species_set_separable: 1..64 terms");
        const int n = ctx->nOwned;
        if (n == 0) return;
        const size_t nAmp = (size_t)n * nTerms, nAx = (size_t)nTerms * (sp.n[0] + sp.n[1] + sp.n[2]);
        double* dev = vt::ctx_stage(ctx, (nAmp + nAx) * sizeof(double));
        std::vector<double> host(nAmp + nAx);
        for (int p = 0; p < n; p++)
            for (int k = 0; k < nTerms; k++) host[(size_t)p * nTerms + k] = amp[(size_t)ctx->order[p] * nTerms + k];
        double* h = host.data() + nAmp;
        std::memcpy(h, a0, (size_t)nTerms * sp.n[0] * sizeof(double));
        std::memcpy(h + (size_t)nTerms * sp.n[0], a1, (size_t)nTerms * sp.n[1] * sizeof(double));
        std::memcpy(h + (size_t)nTerms * (sp.n[0] + sp.n[1]), a2, (size_t)nTerms * sp.n[2] * sizeof(double));
        VT_CUDA(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        const double* d0 = dev + nAmp;
        k_separable_fill<<<n, 256, 0, ctx->stream>>>(sp.f[sp.cur], dev, d0, d0 + (size_t)nTerms * sp.n[0],
                                                     d0 + (size_t)nTerms * (sp.n[0] + sp.n[1]), nTerms, sp.n[0], sp.n[1],
                                                     sp.n[2]);
        ctx->launches++;
        VT_CUDA(cudaGetLastError());
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
        sp.densityValid = false;
        if (sp.tucker) vt::tucker_from_dense(ctx, sp);
    });
}

int vt_measure_dfma_peak(vt_ctx* ctx, double* tflops)
{
    if (ctx->group) { vt_set_error("vt_measure_dfma_peak: not available on a device group (call it on a single-device context)"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        const int blocks = 8 * ctx->prop.multiProcessorCount, threads = 256, iters = 20000;
        double* out = vt::ctx_stage(ctx, (size_t)blocks * threads * sizeof(double));
        cudaEvent_t e0, e1;
        VT_CUDA(cudaEventCreate(&e0));
        VT_CUDA(cudaEventCreate(&e1));
        k_dfma_probe<<<blocks, threads, 0, ctx->stream>>>(out, iters);   // warm-up
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            VT_CUDA(cudaEventRecord(e0, ctx->stream));
            k_dfma_probe<<<blocks, threads, 0, ctx->stream>>>(out, iters);
            VT_CUDA(cudaEventRecord(e1, ctx->stream));
            VT_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            VT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
        }
        ctx->launches += 4;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *tflops = 2.0 * 8 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    });
}

int vt_species_density(vt_ctx* ctx, int species, double* density)
{
    if (ctx->group) return vt::group_species_density(ctx, species, density);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        if (!sp.densityValid) vt::launch_density(ctx, sp);
        if (density) download_tet_array(ctx, sp.density, density, 1);
    });
}

int vt_species_velocity(vt_ctx* ctx, int species, double* velocity)
{
    if (ctx->group) return vt::group_species_velocity(ctx, species, velocity);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        if (!sp.densityValid) vt::launch_density(ctx, sp);
        if (sp.tucker) vt::tucker_materialize(ctx, sp);
        const int n = ctx->nOwned;
        double* vel = vt::ctx_stage(ctx, 3 * (size_t)std::max(1, n) * sizeof(double));
        if (n > 0) {
            k_velocity<<<n, 256, 0, ctx->stream>>>(sp.f[sp.cur], sp.density, vel, sp.n[0], sp.n[1], sp.n[2], sp.vmin[0],
                                                   sp.vmin[1], sp.vmin[2], sp.step[0], sp.step[1], sp.step[2],
                                                   sp.cellVolume);
            ctx->launches++;
            VT_CUDA(cudaGetLastError());
        }
        download_tet_array(ctx, vel, velocity, 3);
    });
}

int vt_field_set(vt_ctx* ctx, const double* E)
{
    if (ctx->group) return vt::group_field_set(ctx, E);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        upload_tet_array(ctx, E, ctx->E, 3);
    });
}

int vt_field_get(vt_ctx* ctx, double* rho, double* phi, double* E)
{
    if (ctx->group) return vt::group_field_get(ctx, rho, phi, E);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (rho) download_tet_array(ctx, ctx->rho, rho, 1);
        if (phi) download_tet_array(ctx, ctx->phi, phi, 1);
        if (E) download_tet_array(ctx, ctx->E, E, 3);
    });
}

int vt_step_full(vt_ctx* ctx, int species, double dt, const double ext[3])
{
    if (ctx->group) return vt::group_step(ctx, species, dt, ext, false);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::check_halo_status(ctx);
        vt::launch_full_step(ctx, species_of(ctx, species), dt, ext);
    });
}

int vt_step_full_host(vt_ctx* ctx, int species, double dt, const double ext[3], const double* E, double* density)
{
    if (ctx->group) { vt_set_error("vt_step_full_host: not available on a device group (call it on a single-device context)"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        const int n = ctx->nOwned;
        // one pinned staging area: [E in (3n) | density out (n)]
        double* pin = vt::ctx_pinned(ctx, 4 * (size_t)n * sizeof(double));
        for (int p = 0; p < n; p++) {
            const int t = ctx->order[p];
            pin[3 * (size_t)p] = E[3 * (size_t)t];
            pin[3 * (size_t)p + 1] = E[3 * (size_t)t + 1];
            pin[3 * (size_t)p + 2] = E[3 * (size_t)t + 2];
        }
        VT_CUDA(cudaMemcpyAsync(ctx->E, pin, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        vt::launch_full_step(ctx, sp, dt, ext);
        double* dpin = pin + 3 * (size_t)n;
        VT_CUDA(cudaMemcpyAsync(dpin, sp.density, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int p = 0; p < n; p++) density[ctx->order[p]] = dpin[p];
    });
}

int vt_step_config(vt_ctx* ctx, int chunkPlanes, int brickTets, int variant)
{
    ctx->chunkPlanes = chunkPlanes;
    ctx->brickTets = brickTets;
    ctx->variant = variant;
    return 0;
}

int vt_step_last_ms(vt_ctx* ctx, float* ms)
{
    if (ctx->group) { vt_set_error("vt_step_last_ms: not available on a device group (call it on a single-device context)"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        VT_CUDA(cudaEventSynchronize(ctx->ev1));
        VT_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    });
}

long vt_launch_count(vt_ctx* ctx) { return ctx->group ? vt::group_launch_count(ctx) : ctx->launches; }

int vt_profile_begin(vt_ctx* ctx)
{
    if (ctx->group) { vt_set_error("vt_profile_begin: not available on a device group (call it on a single-device context)"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (!ctx->regionStart) {
            VT_CUDA(cudaEventCreate(&ctx->regionStart));
            VT_CUDA(cudaEventCreate(&ctx->regionStop));
        }
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->profiling = true;
        ctx->kernelEventsUsed = 0;
        VT_CUDA(cudaEventRecord(ctx->regionStart, ctx->stream));
    });
}

int vt_profile_end(vt_ctx* ctx, float* region_ms, float* step_kernel_ms, int* step_kernels)
{
    if (ctx->group) { vt_set_error("vt_profile_end: not available on a device group (call it on a single-device context)"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (!ctx->profiling) throw std::runtime_error("vt_profile_end without vt_profile_begin");
        VT_CUDA(cudaEventRecord(ctx->regionStop, ctx->stream));
        VT_CUDA(cudaEventSynchronize(ctx->regionStop));
        ctx->profiling = false;
        float ms = 0;
        VT_CUDA(cudaEventElapsedTime(&ms, ctx->regionStart, ctx->regionStop));
        if (region_ms) *region_ms = ms;
        float sum = 0;
        for (size_t i = 0; i + 1 < ctx->kernelEventsUsed; i += 2) {
            float k = 0;
            VT_CUDA(cudaEventElapsedTime(&k, ctx->kernelEvents[i], ctx->kernelEvents[i + 1]));
            sum += k;
        }
        if (step_kernel_ms) *step_kernel_ms = sum;
        if (step_kernels) *step_kernels = (int)(ctx->kernelEventsUsed / 2);
    });
}

int vt_wall_charge_get(vt_ctx* ctx, int species, int entity, double* charge)
{
    if (ctx->group) return vt::group_wall_charge_get(ctx, species, entity, charge);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        *charge = 0.0;
        for (size_t s = 0; s < sp.wallEntities.size(); s++)
            if (sp.wallEntities[s] == entity) {
                VT_CUDA(cudaStreamSynchronize(ctx->stream));
                VT_CUDA(cudaMemcpy(charge, sp.wall + s, sizeof(double), cudaMemcpyDeviceToHost));
            }
    });
}

int vt_wall_charge_reset(vt_ctx* ctx, int species)
{
    if (ctx->group) return vt::group_wall_charge_reset(ctx, species);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        VT_CUDA(cudaMemsetAsync(sp.wall, 0, std::max<size_t>(1, sp.wallEntities.size()) * sizeof(double), ctx->stream));
    });
}

int vt_charge_density(vt_ctx* ctx, const int* species, int nSpecies, const double* background)
{
    if (ctx->group) return vt::group_charge_density(ctx, species, nSpecies, background);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        const int n = ctx->nOwned;
        if (n == 0) return;
        for (int i = 0; i < nSpecies; i++) {
            vt::Species& sp = species_of(ctx, species[i]);
            if (!sp.densityValid) vt::launch_density(ctx, sp);
            k_charge_accum<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->rho, sp.density, sp.charge, n, i == 0);
            ctx->launches++;
        }
        if (background) {
            double* bg = vt::ctx_stage(ctx, (size_t)n * sizeof(double));
            upload_tet_array(ctx, background, bg, 1);
            k_add<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->rho, bg, n);
            ctx->launches++;
        }
        VT_CUDA(cudaGetLastError());
    });
}

}  // extern "C"
