// Tucker path on the device — Solver<Tucker>::_UpdatePDF (src/solver.cpp:141-212 with
// _Flux :314-346, _PDFDerivative<Tucker> :348-361, _PrecomputeNormalTensors :258-293) and the
// Tucker algebra it calls (src/tucker.cpp: operator+ :190-228, operator* :259-300, scalar
// :302-316, Compress :66-98 with _ComputeU :442-465).
//
// Formulation.  The reference adds and multiplies in Tucker format (ranks grow to r_rhs + 16 r per
// face) and then rounds: QR of the stacked factors, core transform, truncated HOSVD of the small
// core.  Because Q is an orthonormal basis of the stacked factors' range, that rounding is exactly
// the truncated HOSVD of the tensor the sum represents, with the reference's per-singular-value
// rule (keep sigma_j > eps*|sigma|_2/sqrt(3), at least one, at most maxRank).  For the stacked ranks
// of this update (R >= n almost immediately) the intermediate is a full n^3 tensor
// anyway, so the device evaluates each sum densely and applies the same truncated HOSVD at the same six
// points per tet-step; the state between steps stays compressed (core r^3 + three n x r factors per tet).
//
// Files.  This one is the host side (state, dispatch, the ABI entry points).  Two kernels run a step:
//   k_tucker_slab  tucker_slab.cu     slab-streaming, for grids of 33..48 nodes per axis, rank caps <= 16 and
//                                     compression errors >= 5.5e-7 (see its header)
//   k_tucker       tucker_kernel.inl  everything else, and the compress / reconstruct / |v.n| table modes of every
//                                     species; compiled once per instantiation by tucker_inst.cu.  One CTA owns a tet
//                                     from reconstruction to the re-compressed result, dense work arrays in a per-CTA
//                                     global scratch:
//     reconstruct      core x1 U0 x2 U1 x3 U2                                       (tucker.cpp:100-104)
//     hosvd_truncate   Gram matrices of the three unfoldings (FP64 tensor cores up to 32 nodes per axis,
//                      register-blocked DFMA above), three symmetric eigen-problems by three warps (Householder
//                      tridiagonalisation + implicit QL), for eps < 5.5e-7 a two-level refinement of the trailing
//                      eigen-directions inside their own subspace (Gram eigenvalues carry ~1e-16 of the trace as
//                      noise), rank rule, projection X x_k U_k^T                    (tucker.cpp:34-50, 442-465)
#include "tucker_internal.h"

#include <algorithm>
#include <cstring>

namespace vt {



static void tucker_block_pointers(void* block, size_t rows, size_t slot, double* buf[2], int* ranks[2])
{
    buf[0] = static_cast<double*>(block);
    buf[1] = buf[0] + rows * slot;
    ranks[0] = reinterpret_cast<int*>(buf[1] + rows * slot);
    ranks[1] = ranks[0] + rows * 3;
}

void tucker_destroy(TuckerState* ts)
{
    if (!ts) return;
    cudaFree(ts->block);
    cudaFree(ts->vnabs);
    cudaFree(ts->vnabsRanks);
    cudaFree(ts->scratch);
    delete ts;
}

namespace {

void fill_params(vt_ctx* ctx, Species& sp, TuckerState& ts, TuckerParams& P)
{
    std::memset(&P, 0, sizeof(P));
    P.nOwned = ctx->nOwned;
    for (int k = 0; k < 3; k++) {
        P.n[k] = sp.n[k];
        P.rcap[k] = ts.rcap[k];
        P.vmin[k] = sp.vmin[k];
        P.step[k] = sp.step[k];
        P.inv2h[k] = 1.0 / (2 * sp.step[k]);
    }
    P.N = sp.N;
    P.coreCap = ts.coreCap;
    P.slot = ts.slot;
    P.rec = sp.rec;
    P.E = ctx->E;
    P.vnabs = ts.vnabs;
    P.vnabsRanks = ts.vnabsRanks;
    P.vslot = ts.vslot;
    P.src = sp.src;
    P.density = sp.density;
    P.wall = sp.wall;
    P.scratch = ts.scratch;
    P.scratchPerCTA = (size_t)6 * sp.N + 5 * (size_t)kMaxN * kMaxN;
    P.qm = sp.charge / sp.mass;
    P.cellVolume = sp.cellVolume;
    P.eps = ts.comprErr;
    P.maxRank = ts.maxRank;
    // Gram matrices: FP64 tensor cores (mma.sync m8n8k4) or DFMA.  Measured on B200 (profiles/
    // r2_tucker_gram_dmma_vs_dfma.md): the two pipes have the same peak (37.0 vs 35.0 TFLOP/s), and the
    // tensor-core version wins where the operand traffic from shared memory dominates — 11^3: 4.8 vs 5.3
    // ms, 32^3: 36.9 vs 44.8 ms per step — and loses at 48^3 (32.5 vs 29.4 ms), where the DFMA version's
    // register blocking already amortises the loads.  Default: tensor cores up to 32 nodes per axis;
    // VT_TUCKER_GRAM=dmma|dfma forces either.
    static const char* mode = std::getenv("VT_TUCKER_GRAM");
    const int nmax = std::max({sp.n[0], sp.n[1], sp.n[2]});
    P.gramDmma = mode ? (std::string(mode) == "dmma" ? 1 : 0) : (nmax <= 32 ? 1 : 0);
}

void launch(vt_ctx* ctx, TuckerState& ts, const TuckerParams& P)
{
    if (ctx->nOwned == 0) return;
    const int grid = std::min(ctx->nOwned, ts.scratchCTAs);
    const int nmax = std::max({P.n[0], P.n[1], P.n[2]});
    const size_t smem = (3 * (size_t)(nmax | 1) * nmax + 2 * (size_t)tile_cap(nmax)) * sizeof(double);
    if (nmax <= 16) launch_k_tucker_16(grid, smem, ctx->stream, P);
    else if (nmax <= 32) launch_k_tucker_32(grid, smem, ctx->stream, P);
    else launch_k_tucker_64(grid, smem, ctx->stream, P);
    ctx->launches++;
    VT_CUDA(cudaGetLastError());
}

TuckerState& state_of(Species& sp)
{
    if (!sp.tucker) throw std::runtime_error("species is not in Tucker format (vt_tucker_enable)");
    return *sp.tucker;
}

Species& species_of(vt_ctx* ctx, int s)
{
    if (s < 0 || s >= (int)ctx->species.size()) throw std::invalid_argument("bad species id");
    return *ctx->species[s];
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

}  // namespace

// Dense copy of the current Tucker tensors in sp.f[sp.cur] (the buffers the Full-format moment
// and download kernels read), refreshed lazily after a step or an upload.
void tucker_materialize(vt_ctx* ctx, Species& sp)
{
    TuckerState& ts = state_of(sp);
    if (ts.densePending) tucker_from_dense(ctx, sp);   // rows uploaded since the last compression
    if (ts.denseValid) return;
    TuckerParams P;
    fill_params(ctx, sp, ts, P);
    P.mode = 2;
    P.in = ts.buf[ts.cur];
    P.rin = ts.ranks[ts.cur];
    P.denseOut = sp.f[sp.cur];
    launch(ctx, ts, P);
    ts.denseValid = true;
}

// Compress the dense rows in sp.f[sp.cur] into the Tucker state, exactly (precision 0) as
// ParticleData<Tucker>::SetMaxwellPDF does (particle_data.cpp:64-69), ranks capped by the slot size.
void tucker_begin_dense_write(vt_ctx* ctx, Species& sp)
{
    TuckerState& ts = state_of(sp);
    if (!ts.densePending) tucker_materialize(ctx, sp);   // while an upload is pending the dense copy IS the state
}

void tucker_end_dense_write(vt_ctx*, Species& sp)
{
    TuckerState& ts = state_of(sp);
    ts.densePending = true;
    ts.denseValid = true;
    sp.densityValid = false;
}

static void ensure_compressed(vt_ctx* ctx, Species& sp)
{
    if (state_of(sp).densePending) tucker_from_dense(ctx, sp);
}

void tucker_from_dense(vt_ctx* ctx, Species& sp)
{
    TuckerState& ts = state_of(sp);
    ts.densePending = false;
    TuckerParams P;
    fill_params(ctx, sp, ts, P);
    P.mode = 1;
    P.denseIn = sp.f[sp.cur];
    P.out = ts.buf[ts.cur];
    P.rout = ts.ranks[ts.cur];
    launch(ctx, ts, P);
    sp.densityValid = false;
    // sp.f keeps the caller's tensors; they equal the reconstruction unless the rank cap binds
    ts.denseValid = ts.maxRank >= std::max({sp.n[0], sp.n[1], sp.n[2]});
}

// Initial ghost fill of a partitioned Tucker species (vt_halo_push_current dispatches here).
void tucker_push_current(vt_ctx* ctx, Species& sp)
{
    TuckerState& ts = state_of(sp);
    ensure_compressed(ctx, sp);
    if (ctx->nOwned == 0 || sp.nPeers == 0) return;
    if (ts.nPeers != sp.nPeers) throw std::runtime_error("partitioned Tucker species: vt_tucker_halo_attach has not been called");
    TuckerParams P;
    fill_params(ctx, sp, ts, P);
    P.mode = 4;
    P.in = ts.buf[ts.cur];
    P.rin = ts.ranks[ts.cur];
    for (int i = 0; i < kMaxPeers; i++) {
        P.peerOut[i] = i < ts.nPeers ? ts.peerBuf[i][ts.cur] : nullptr;
        P.peerRout[i] = i < ts.nPeers ? ts.peerRanks[i][ts.cur] : nullptr;
    }
    launch(ctx, ts, P);
    VT_CUDA(cudaStreamSynchronize(ctx->stream));
}

namespace {

void ensure_vnabs(vt_ctx* ctx, Species& sp, TuckerState& ts)
{
    if (ts.vnabsValid) return;
    TuckerParams P;
    fill_params(ctx, sp, ts, P);
    P.mode = 3;
    P.epsAbs = ts.comprErr;
    launch(ctx, ts, P);
    ts.vnabsValid = true;
}

}  // namespace
}  // namespace vt

using namespace vt;

extern "C" {

int vt_tucker_enable(vt_ctx* ctx, int species, double comprErr, int maxRank)
{
    if (ctx->group) return vt::group_tucker_enable(ctx, species, comprErr, maxRank);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        Species& sp = species_of(ctx, species);
        for (int k = 0; k < 3; k++)
            if (sp.n[k] > kMaxN) throw std::runtime_error("Tucker path: velocity grid larger than 64 nodes per axis");
        if (sp.tucker) {
            // re-configuration (SetCompressionError / SetMaxRank after the PDF was set): keep the state
            tucker_materialize(ctx, sp);
            VT_CUDA(cudaStreamSynchronize(ctx->stream));
            tucker_destroy(sp.tucker);
            sp.tucker = nullptr;
        }
        TuckerState* ts = new TuckerState();
        sp.tucker = ts;
        ts->comprErr = comprErr;
        const int nmax = std::max({sp.n[0], sp.n[1], sp.n[2]});
        ts->maxRank = maxRank > 0 ? maxRank : nmax;   // particle_data.cpp:18
        for (int k = 0; k < 3; k++) ts->rcap[k] = std::min(ts->maxRank, sp.n[k]);
        ts->coreCap = (size_t)ts->rcap[0] * ts->rcap[1] * ts->rcap[2];
        ts->slot = ts->coreCap + (size_t)sp.n[0] * ts->rcap[0] + (size_t)sp.n[1] * ts->rcap[1] + (size_t)sp.n[2] * ts->rcap[2];
        const size_t nA = std::max(1, ctx->nOwned);
        ts->rows = (size_t)std::max(1, ctx->nOwned + ctx->nGhost);
        const size_t blockBytes = 2 * ts->rows * ts->slot * sizeof(double) + 2 * ts->rows * 3 * sizeof(int);
        VT_CUDA(cudaMalloc(&ts->block, blockBytes));
        VT_CUDA(cudaMemset(ts->block, 0, blockBytes));
        tucker_block_pointers(ts->block, ts->rows, ts->slot, ts->buf, ts->ranks);
        ts->vslot = 216 + 6 * (size_t)(sp.n[0] + sp.n[1] + sp.n[2]);
        VT_CUDA(cudaMalloc(&ts->vnabs, nA * 4 * ts->vslot * sizeof(double)));
        VT_CUDA(cudaMalloc(&ts->vnabsRanks, nA * 12 * sizeof(int)));
        ts->scratchCTAs = (nmax <= 32 ? 4 : 2) * ctx->prop.multiProcessorCount;
        const size_t per = (size_t)6 * sp.N + 5 * (size_t)kMaxN * kMaxN;
        VT_CUDA(cudaMalloc(&ts->scratch, (size_t)ts->scratchCTAs * per * sizeof(double)));
        tucker_from_dense(ctx, sp);   // whatever the species holds (zeros after vt_species_create)
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

// ---- multi-GPU: the Tucker state of boundary tets is mirrored into the peers' ghost rows
struct TuckerIpc {
    cudaIpcMemHandle_t block;
    unsigned long long rows, slot;
    unsigned char pad[48];
};
static_assert(sizeof(TuckerIpc) == 128, "Tucker halo handle is 128 bytes");

int vt_tucker_halo_export(vt_ctx* ctx, int species, void* handle)
{
    if (ctx->group) { vt_set_error("vt_tucker_halo_export: not available on a device group"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        Species& sp = species_of(ctx, species);
        TuckerState& ts = state_of(sp);
        TuckerIpc pk;
        std::memset(&pk, 0, sizeof(pk));
        VT_CUDA(cudaIpcGetMemHandle(&pk.block, ts.block));
        pk.rows = ts.rows;
        pk.slot = ts.slot;
        std::memcpy(handle, &pk, sizeof(pk));
    });
}

int vt_tucker_halo_attach(vt_ctx* ctx, int species, int nPeers, const void* peerHandles)
{
    if (ctx->group) { vt_set_error("vt_tucker_halo_attach: not available on a device group"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        Species& sp = species_of(ctx, species);
        TuckerState& ts = state_of(sp);
        if (nPeers != sp.nPeers) throw std::runtime_error("vt_tucker_halo_attach: call vt_halo_attach with the same peers first");
        const TuckerIpc* pk = static_cast<const TuckerIpc*>(peerHandles);
        for (int i = 0; i < nPeers; i++) {
            if (pk[i].slot != ts.slot) throw std::runtime_error("peer Tucker state has a different slot size (maxRank / grid mismatch)");
            void* base = nullptr;
            VT_CUDA(cudaIpcOpenMemHandle(&base, pk[i].block, cudaIpcMemLazyEnablePeerAccess));
            ctx->ipcOpened.push_back(base);
            tucker_block_pointers(base, (size_t)pk[i].rows, (size_t)pk[i].slot, ts.peerBuf[i], ts.peerRanks[i]);
        }
        ts.nPeers = nPeers;
    });
}

int vt_tucker_halo_attach_local(vt_ctx* ctx, int species, int nPeers, vt_ctx* const* peerCtx, const int32_t* peerSpecies)
{
    if (ctx->group) { vt_set_error("vt_tucker_halo_attach_local: not available on a device group"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        Species& sp = species_of(ctx, species);
        TuckerState& ts = state_of(sp);
        if (nPeers != sp.nPeers) throw std::runtime_error("vt_tucker_halo_attach_local: call vt_halo_attach_local with the same peers first");
        for (int i = 0; i < nPeers; i++) {
            TuckerState& ps = state_of(species_of(peerCtx[i], peerSpecies[i]));
            if (ps.slot != ts.slot) throw std::runtime_error("peer Tucker state has a different slot size (maxRank / grid mismatch)");
            tucker_block_pointers(ps.block, ps.rows, ps.slot, ts.peerBuf[i], ts.peerRanks[i]);
        }
        ts.nPeers = nPeers;
    });
}

int vt_tucker_set_pdf(vt_ctx* ctx, int species, const double* dense)
{
    if (ctx->group) return vt::group_species_set_pdf(ctx, species, 0, ctx->nOwned, dense);
    // vt_species_set_pdf re-compresses a Tucker species after the upload
    if (!ctx || species < 0 || species >= (int)ctx->species.size() || !ctx->species[species]->tucker) {
        vt_set_error("species is not in Tucker format (vt_tucker_enable)");
        return 1;
    }
    return vt_species_set_pdf(ctx, species, 0, ctx->nOwned, dense);
}

int vt_tucker_get_pdf(vt_ctx* ctx, int species, double* dense)
{
    if (ctx->group) return vt::group_species_get_pdf(ctx, species, 0, ctx->nOwned, dense);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        Species& sp = species_of(ctx, species);
        TuckerState& ts = state_of(sp);
        (void)ts;
        tucker_materialize(ctx, sp);
        const double* d = sp.f[sp.cur];
        for (int p = 0; p < ctx->nOwned; p++)
            VT_CUDA(cudaMemcpyAsync(dense + (size_t)ctx->order[p] * sp.N, d + (size_t)p * sp.N, (size_t)sp.N * sizeof(double),
                                    cudaMemcpyDeviceToHost, ctx->stream));
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int vt_tucker_get_ranks(vt_ctx* ctx, int species, int32_t* ranks)
{
    if (ctx->group) return vt::group_tucker_get_ranks(ctx, species, ranks);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        Species& sp = species_of(ctx, species);
        TuckerState& ts = state_of(sp);
        ensure_compressed(ctx, sp);
        std::vector<int> r(3 * (size_t)ctx->nOwned);
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
        VT_CUDA(cudaMemcpy(r.data(), ts.ranks[ts.cur], r.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (int p = 0; p < ctx->nOwned; p++)
            for (int k = 0; k < 3; k++) ranks[3 * (size_t)ctx->order[p] + k] = r[3 * (size_t)p + k];
    });
}

int vt_tucker_get_factors(vt_ctx* ctx, int species, int tet, int32_t ranks[3], double* core, double* u0, double* u1,
                          double* u2)
{
    if (ctx->group) return vt::group_tucker_get_factors(ctx, species, tet, ranks, core, u0, u1, u2);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        Species& sp = species_of(ctx, species);
        TuckerState& ts = state_of(sp);
        ensure_compressed(ctx, sp);
        if (tet < 0 || tet >= ctx->nOwned) throw std::out_of_range("tet index");
        const int p = ctx->inv[tet];
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
        int r[3];
        VT_CUDA(cudaMemcpy(r, ts.ranks[ts.cur] + 3 * (size_t)p, sizeof(r), cudaMemcpyDeviceToHost));
        std::vector<double> slot(ts.slot);
        VT_CUDA(cudaMemcpy(slot.data(), ts.buf[ts.cur] + (size_t)p * ts.slot, ts.slot * sizeof(double), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 3; k++) ranks[k] = r[k];
        std::memcpy(core, slot.data(), (size_t)r[0] * r[1] * r[2] * sizeof(double));
        const double* u = slot.data() + ts.coreCap;
        double* outs[3] = {u0, u1, u2};
        for (int k = 0; k < 3; k++) {
            std::memcpy(outs[k], u, (size_t)sp.n[k] * r[k] * sizeof(double));
            u += (size_t)sp.n[k] * ts.rcap[k];
        }
    });
}

int vt_tucker_density(vt_ctx* ctx, int species, double* density)
{
    if (ctx->group) return vt::group_species_density(ctx, species, density);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        Species& sp = species_of(ctx, species);
        TuckerState& ts = state_of(sp);
        (void)ts;
        // Sum() of the current tensors (particle_data.cpp:93-102); launch_density materialises them
        if (!sp.densityValid) launch_density(ctx, sp);
        if (density) {
            std::vector<double> tmp(ctx->nOwned);
            VT_CUDA(cudaStreamSynchronize(ctx->stream));
            VT_CUDA(cudaMemcpy(tmp.data(), sp.density, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (int p = 0; p < ctx->nOwned; p++) density[ctx->order[p]] = tmp[p];
        }
    });
}

int vt_tucker_last_kernel(vt_ctx* ctx, int species, int* kernel)
{
    if (ctx->group) {
        vt_set_error("vt_tucker_last_kernel: not available on a device group");
        return 1;
    }
    return guard([&] {
        Species& sp = species_of(ctx, species);
        *kernel = state_of(sp).lastKernel;
    });
}

int vt_step_tucker(vt_ctx* ctx, int species, double dt, const double ext[3])
{
    if (ctx->group) return vt::group_step(ctx, species, dt, ext, true);
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        Species& sp = species_of(ctx, species);
        TuckerState& ts = state_of(sp);
        if (sp.danglingFaces > 0)
            throw std::runtime_error(std::to_string(sp.danglingFaces) + " boundary faces have no neighbour and no particle BC");
        if (sp.nPeers > 0 && ts.nPeers != sp.nPeers)
            throw std::runtime_error("partitioned Tucker species: vt_tucker_halo_attach has not been called");
        ensure_compressed(ctx, sp);
        check_halo_status(ctx);
        ensure_vnabs(ctx, sp, ts);
        TuckerParams P;
        fill_params(ctx, sp, ts, P);
        P.mode = 0;
        P.in = ts.buf[ts.cur];
        P.rin = ts.ranks[ts.cur];
        P.out = ts.buf[ts.cur ^ 1];
        P.rout = ts.ranks[ts.cur ^ 1];
        P.dt = dt;
        for (int k = 0; k < 3; k++) P.ext[k] = ext ? ext[k] : 0.0;
        P.wallScale = sp.charge * dt * sp.cellVolume;
        for (int i = 0; i < kMaxPeers; i++) {
            P.peerOut[i] = i < ts.nPeers ? ts.peerBuf[i][ts.cur ^ 1] : nullptr;
            P.peerRout[i] = i < ts.nPeers ? ts.peerRanks[i][ts.cur ^ 1] : nullptr;
        }
        // VT_TUCKER_PROFILE=1: per-phase clock cycles of thread 0 of CTA 0, printed to stderr
        static const bool profile = getenv("VT_TUCKER_PROFILE") != nullptr;
        long long* profDev = nullptr;
        if (profile) {
            VT_CUDA(cudaMalloc(&profDev, 24 * sizeof(long long)));
            VT_CUDA(cudaMemsetAsync(profDev, 0, 24 * sizeof(long long), ctx->stream));
            P.prof = profDev;
        }
        cudaEvent_t e0 = ctx->ev0, e1 = ctx->ev1;
        if (ctx->profiling) {   // vt_profile_begin/end: one event pair per step-kernel launch
            if (ctx->kernelEventsUsed + 2 > ctx->kernelEvents.size()) {
                cudaEvent_t a, b;
                VT_CUDA(cudaEventCreate(&a));
                VT_CUDA(cudaEventCreate(&b));
                ctx->kernelEvents.push_back(a);
                ctx->kernelEvents.push_back(b);
            }
            e0 = ctx->kernelEvents[ctx->kernelEventsUsed++];
            e1 = ctx->kernelEvents[ctx->kernelEventsUsed++];
        }
        VT_CUDA(cudaEventRecord(e0, ctx->stream));
        const bool slab = slab_eligible(ctx, P);   // tucker_slab.cu: slab-streaming kernel where it applies
        if (slab) launch_tucker_slab(ctx, ts, P);
        else launch(ctx, ts, P);
        ts.lastKernel = slab ? 2 : 1;
        VT_CUDA(cudaEventRecord(e1, ctx->stream));
        if (profile) {
            long long h[24];
            VT_CUDA(cudaStreamSynchronize(ctx->stream));
            VT_CUDA(cudaMemcpy(h, profDev, sizeof(h), cudaMemcpyDeviceToHost));
            cudaFree(profDev);
            if (slab)
                fprintf(stderr,
                        "[vt_step_tucker, slab kernel] cycles of CTA 0: pass 1 (expand + X + G0, G1) %lld  pass 2 (G2) %lld  "
                        "Cholesky + L^T L %lld  Jacobi %lld  select %lld  pass 3 (core) %lld  total %lld | pass 1 phases: M2 %lld  T %lld  "
                        "expand %lld  X %lld  Gram %lld | Jacobi rounds %lld sweeps %lld: scan %lld  phase 1 %lld  phase 2 %lld | active sets of the last problem, by sweep (modes 0|1|2): %llx %llx %llx %llx %llx %llx\n",
                        h[0], h[1], h[5], h[2], h[3], h[4], h[7], h[8], h[9], h[10], h[11], h[12], h[13], h[14], h[15], h[16], h[17], h[18], h[19], h[20], h[21], h[22], h[23]);
            else
                fprintf(stderr,
                        "[vt_step_tucker] cycles of CTA 0: gram %lld  eig %lld  select %lld  project %lld  reconstruct %lld  "
                        "flux %lld  derivative %lld  total %lld\n",
                        h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
        }
        ts.cur ^= 1;
        ts.denseValid = false;
        sp.densityValid = true;
    });
}

}  // extern "C"
