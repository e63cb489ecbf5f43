// Multi-GPU halo plumbing: CUDA-IPC mapped peer buffers, initial ghost fill, device-side barrier.
// The per-step exchange itself is fused into the step kernel (full_step.cu, HALO path).
#include "vt_internal.h"

#include <cstdlib>
#include <cstring>

namespace {

struct IpcPack {
    cudaIpcMemHandle_t f0, f1, flags;
};
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");

// copy the current state of every pushed tet into the ghost rows of the peers' current buffer
__global__ void k_push_rows(const double* __restrict__ f, const vt::TetRec* __restrict__ rec, int N,
                            double* const* __restrict__ peerBase)
{
    const int t = blockIdx.x;
    const vt::TetRec& r = rec[t];
    const double* src = f + (size_t)t * N;
    for (int q = 0; q < 4; q++) {
        if (r.pushPeer[q] < 0) continue;
        double* dst = peerBase[r.pushPeer[q]] + (size_t)r.pushRow[q] * N;
        for (int e = threadIdx.x; e < N; e += blockDim.x) dst[e] = src[e];
    }
}

// All ranks announce `epoch` in every peer's flag array, then wait for every peer's announcement.
// Stream order guarantees the step kernel (and its peer stores) finished before this runs.
__global__ void k_halo_barrier(uint32_t* myFlags, uint32_t* const* peerFlags, const int* peerRank, int nPeers,
                               int myRank, uint32_t epoch, volatile int* status, unsigned long long timeoutNs)
{
    const int i = threadIdx.x;
    if (i >= nPeers) return;
    __threadfence_system();
    volatile uint32_t* out = peerFlags[i] + myRank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(out), "r"(epoch) : "memory");
    const uint32_t* in = myFlags + peerRank[i];
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    uint32_t v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(in) : "memory");
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if ((int32_t)(v - epoch) < 0 && (*status != 0 || t - t0 > timeoutNs)) {
            // a peer did not arrive (it died, or it is far behind): do not hang the GPU.  The flag lives
            // in mapped host memory and is sticky: every later vt_step_* / vt_halo_barrier / vt_sync of
            // this context fails instead of computing with ghost rows that may be stale.
            *status = 1;
            __threadfence_system();
            break;
        }
    } while ((int32_t)(v - epoch) < 0);
    __threadfence_system();
}

vt::Species& species_of(vt_ctx* ctx, int s)
{
    if (s < 0 || s >= (int)ctx->species.size()) throw std::invalid_argument("bad species id");
    return *ctx->species[s];
}

void ensure_flags(vt_ctx* ctx)
{
    if (ctx->flags) return;
    VT_CUDA(cudaMalloc(&ctx->flags, 64 * sizeof(uint32_t)));
    VT_CUDA(cudaMemset(ctx->flags, 0, 64 * sizeof(uint32_t)));
    // the barrier's time-out flag: mapped host memory, so the host sees it without synchronising
    VT_CUDA(cudaHostAlloc(&ctx->haloStatusHost, sizeof(int), cudaHostAllocMapped));
    *ctx->haloStatusHost = 0;
    VT_CUDA(cudaHostGetDevicePointer(&ctx->haloStatus, ctx->haloStatusHost, 0));
    if (const char* t = std::getenv("VT_COMM_TIMEOUT_MS")) ctx->haloTimeoutNs = (unsigned long long)std::atoll(t) * 1000000ULL;
}

// the barrier kernel reads the peers' flag pointers and ranks from a small device table
void upload_halo_table(vt_ctx* ctx)
{
    struct Table {
        uint32_t* pf[vt::kMaxPeers];
        int pr[vt::kMaxPeers];
    } tb;
    for (int i = 0; i < vt::kMaxPeers; i++) {
        tb.pf[i] = ctx->peerFlags[i];
        tb.pr[i] = ctx->peerRank[i];
    }
    if (!ctx->haloTable) VT_CUDA(cudaMalloc(&ctx->haloTable, sizeof(Table)));
    VT_CUDA(cudaMemcpy(ctx->haloTable, &tb, sizeof(tb), cudaMemcpyHostToDevice));
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

namespace vt {
void rebuild_tet_records_public(vt_ctx* ctx, Species& sp);
void check_halo_status(vt_ctx* ctx)
{
    if (ctx->haloStatusHost && *static_cast<volatile int*>(ctx->haloStatusHost) != 0)
        throw std::runtime_error("halo barrier timed out earlier: a peer rank did not reach the step barrier "
                                 "(VT_COMM_TIMEOUT_MS); the ghost rows of this context are not trustworthy");
}
}

extern "C" {

int vt_halo_export(vt_ctx* ctx, int species, void* handles)
{
    if (ctx->group) { vt_set_error("vt_halo_export: not available on a device group"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        ensure_flags(ctx);
        IpcPack pk;
        VT_CUDA(cudaIpcGetMemHandle(&pk.f0, sp.f[0]));
        VT_CUDA(cudaIpcGetMemHandle(&pk.f1, sp.f[1]));
        VT_CUDA(cudaIpcGetMemHandle(&pk.flags, ctx->flags));
        std::memcpy(handles, &pk, sizeof(pk));
    });
}

int vt_halo_attach(vt_ctx* ctx, int species, int myRank, int nPeers, const int32_t* peerRanks, const void* peerHandles)
{
    if (ctx->group) { vt_set_error("vt_halo_attach: not available on a device group"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        if (nPeers > vt::kMaxPeers) throw std::runtime_error("too many halo peers");
        if (myRank < 0 || myRank >= 64) throw std::runtime_error("rank out of range for the barrier flags");
        ctx->rank = myRank;
        ctx->nPeers = nPeers;
        sp.nPeers = nPeers;
        const IpcPack* pk = static_cast<const IpcPack*>(peerHandles);
        for (int i = 0; i < nPeers; i++) {
            if (peerRanks[i] < 0 || peerRanks[i] >= 64) throw std::runtime_error("peer rank out of range");
            ctx->peerRank[i] = peerRanks[i];
            void *p0 = nullptr, *p1 = nullptr, *pf = nullptr;
            VT_CUDA(cudaIpcOpenMemHandle(&p0, pk[i].f0, cudaIpcMemLazyEnablePeerAccess));
            VT_CUDA(cudaIpcOpenMemHandle(&p1, pk[i].f1, cudaIpcMemLazyEnablePeerAccess));
            sp.peerF[i][0] = static_cast<double*>(p0);
            sp.peerF[i][1] = static_cast<double*>(p1);
            ctx->ipcOpened.push_back(p0);
            ctx->ipcOpened.push_back(p1);
            if (!ctx->peerFlags[i]) {
                VT_CUDA(cudaIpcOpenMemHandle(&pf, pk[i].flags, cudaIpcMemLazyEnablePeerAccess));
                ctx->peerFlags[i] = static_cast<uint32_t*>(pf);
                ctx->ipcOpened.push_back(pf);
            }
        }
        upload_halo_table(ctx);
    });
}

int vt_halo_attach_local(vt_ctx* ctx, int species, int myRank, int nPeers, const int32_t* peerRanks,
                         vt_ctx* const* peerCtx, const int32_t* peerSpecies)
{
    if (ctx->group) { vt_set_error("vt_halo_attach_local: not available on a device group"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        if (nPeers > vt::kMaxPeers) throw std::runtime_error("too many halo peers");
        if (myRank < 0 || myRank >= 64) throw std::runtime_error("rank out of range for the barrier flags");
        ensure_flags(ctx);
        ctx->rank = myRank;
        ctx->nPeers = nPeers;
        sp.nPeers = nPeers;
        for (int i = 0; i < nPeers; i++) {
            if (peerRanks[i] < 0 || peerRanks[i] >= 64) throw std::runtime_error("peer rank out of range");
            vt_ctx* pc = peerCtx[i];
            if (!pc || pc == ctx) throw std::invalid_argument("vt_halo_attach_local: bad peer context");
            vt::Species& psp = species_of(pc, peerSpecies[i]);
            if (psp.N != sp.N) throw std::invalid_argument("vt_halo_attach_local: peer species has another velocity grid");
            if (pc->device != ctx->device) {
                int can = 0;
                VT_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, pc->device));
                if (!can) throw std::runtime_error("vt_halo_attach_local: no peer access between the devices");
                cudaError_t e = cudaDeviceEnablePeerAccess(pc->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) VT_CUDA(e);
                (void)cudaGetLastError();
            }
            VT_CUDA(cudaSetDevice(pc->device));
            ensure_flags(pc);
            VT_CUDA(cudaSetDevice(ctx->device));
            ctx->peerRank[i] = peerRanks[i];
            sp.peerF[i][0] = psp.f[0];
            sp.peerF[i][1] = psp.f[1];
            ctx->peerFlags[i] = pc->flags;
        }
        upload_halo_table(ctx);
    });
}

int vt_halo_set_push(vt_ctx* ctx, int species, const int32_t* pushPeer, const int32_t* pushRow)
{
    if (ctx->group) { vt_set_error("vt_halo_set_push: not available on a device group"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        const size_t n4 = 4 * (size_t)ctx->nOwned;
        sp.pushPeer.assign(pushPeer, pushPeer + n4);
        sp.pushRow.assign(pushRow, pushRow + n4);
        for (size_t i = 0; i < n4; i++)
            if (sp.pushPeer[i] >= sp.nPeers) throw std::invalid_argument("push peer index out of range");
        vt::rebuild_tet_records_public(ctx, sp);
    });
}

int vt_halo_push_current(vt_ctx* ctx, int species)
{
    if (ctx->group) { vt_set_error("vt_halo_push_current: not available on a device group"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        vt::Species& sp = species_of(ctx, species);
        if (sp.tucker) {
            vt::tucker_push_current(ctx, sp);
            return;
        }
        if (ctx->nOwned == 0 || sp.nPeers == 0) return;
        double* base[vt::kMaxPeers];
        for (int i = 0; i < vt::kMaxPeers; i++) base[i] = i < sp.nPeers ? sp.peerF[i][sp.cur] : nullptr;
        double** baseDev = reinterpret_cast<double**>(vt::ctx_stage(ctx, sizeof(base)));
        VT_CUDA(cudaMemcpyAsync(baseDev, base, sizeof(base), cudaMemcpyHostToDevice, ctx->stream));
        k_push_rows<<<ctx->nOwned, 256, 0, ctx->stream>>>(sp.f[sp.cur], sp.rec, sp.N, baseDev);
        ctx->launches++;
        VT_CUDA(cudaGetLastError());
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int vt_halo_barrier(vt_ctx* ctx)
{
    if (ctx->group) { vt_set_error("vt_halo_barrier: not available on a device group"); return 1; };
    return guard([&] {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (ctx->nPeers == 0) return;
        vt::check_halo_status(ctx);
        ctx->epoch++;
        // peer flag pointers and ranks: the device table written by vt_halo_attach(_local)
        if (!ctx->haloTable) upload_halo_table(ctx);
        uint32_t* const* pfDev = reinterpret_cast<uint32_t* const*>(ctx->haloTable);
        const int* prDev = reinterpret_cast<const int*>(reinterpret_cast<const char*>(ctx->haloTable) +
                                                        vt::kMaxPeers * sizeof(uint32_t*));
        k_halo_barrier<<<1, 32, 0, ctx->stream>>>(ctx->flags, pfDev, prDev, ctx->nPeers, ctx->rank, ctx->epoch,
                                                  ctx->haloStatus, ctx->haloTimeoutNs);
        ctx->launches++;
        VT_CUDA(cudaGetLastError());
    });
}

}  // extern "C"
