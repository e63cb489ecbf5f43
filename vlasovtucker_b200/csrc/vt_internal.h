// Internal structures of libvt_b200 (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vt_b200.h"

#define VT_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

namespace vt {

// One record per owned tet, device order.  Read once per CTA into shared memory.
struct TetRec {
    double coef[4];      // face->area / tet->volume            (solver.cpp:168)
    double nrm[4][3];    // face->normal                        (solver.cpp:271)
    int32_t nbr[4];      // device row of tet->adjTets[f]; -1 none; <= -2: source PDF -2-id
    uint8_t bc[4];       // ParticleBCType                      (solver.h:25)
    int8_t wallSlot[4];  // >= 0: Absorbing && collectCharge -> accumulator slot (solver.cpp:171)
    double area[4];      // face->area (wall charge)            (solver.cpp:173)
    // multi-GPU halo: this tet is a ghost row on up to 4 peer GPUs; the step kernel stores its
    // new state there as well (peer index into StepParams::peerFn, row in the peer's buffer)
    int32_t pushPeer[4]; // -1 = unused
    int32_t pushRow[4];
    int32_t ghostFaces;  // bit f: the neighbour across face f is a ghost row (owned by a peer): where v.n_f >= 0 that peer reads this tet
    int32_t pad;         // sizeof == 224: records are copied with 16-byte cp.async
};
static_assert(sizeof(TetRec) == 224, "TetRec must stay a multiple of 16 bytes");

constexpr int kMaxPeers = 16;

struct StepParams {
    const double* f;      // state at step n, rows of N doubles (owned then ghost)
    double* fn;           // state at step n+1
    const TetRec* rec;
    const double* E;      // 3 per owned tet, device order
    const double* src;    // source PDFs, rows of N doubles
    double* densPartial;  // [nOwned * nChunks] sum over the chunk of f^{n+1}
    double* wall;         // [nSlots] absorbed charge accumulators
    int nOwned;
    int n0, n1, n2, N;
    int nvec0;            // n0 / VEC
    int nLG;              // line groups per CTA = blockDim / nvec0
    int chunkPlanes, nChunks, brickTets;
    int densSplit;        // partial sums written per (tet, chunk): 1, or one per consumer warp
    double vmin[3], step[3], inv2h[3];
    double qm;            // charge / mass
    double ext[3];
    double dt;
    double wallScale;     // charge * dt * cellVolume
    double* peerFn[kMaxPeers];   // peers' step-n+1 buffers (CUDA-IPC mapped), for the halo push
};

struct Species {
    int n[3];
    int N;
    double vmin[3], vmax[3], step[3], cellVolume;
    double mass, charge;
    double* f[2] = {nullptr, nullptr};
    int cur = 0;
    TetRec* rec = nullptr;            // device, per owned tet (BCs folded in)
    std::vector<TetRec> recHost;
    double* src = nullptr;
    int nSrc = 0;
    double* density = nullptr;        // device, owned tets, device order
    double* densPartial = nullptr;
    int densPartialCap = 0;
    bool densityValid = false;
    double* wall = nullptr;           // device accumulators
    std::vector<int> wallEntities;    // slot -> entity
    std::vector<uint8_t> bcType, collect;
    std::vector<int32_t> sourceId;
    int danglingFaces = 0;            // boundary faces with neither neighbour nor particle BC
    int fastOnly = -1;                // 1: no tet needs the boundary/halo branches; -1: not evaluated
    int32_t* tetLists = nullptr;      // device: boundary/halo tets then interior tets (bulk-copy step kernel)
    int nGeneric = 0, nFast = 0;
    // halo (multi-GPU): peers' ping-pong buffers opened through CUDA IPC
    int nPeers = 0;
    double* peerF[kMaxPeers][2] = {};
    std::vector<int32_t> pushPeer, pushRow;   // 4 per owned tet, caller order, -1 unused
    struct TuckerState* tucker = nullptr;     // compressed state when the species is in Tucker format
};

struct PoissonData;
struct Group;
void tucker_destroy(TuckerState* ts);

}  // namespace vt

struct vt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaDeviceProp prop;
    long launches = 0;
    float lastStepMs = 0;
    // profiling region (vt_profile_begin/end): one event pair per step-kernel launch
    bool profiling = false;
    cudaEvent_t regionStart = nullptr, regionStop = nullptr;
    std::vector<cudaEvent_t> kernelEvents;   // pairs, grown on demand and reused
    size_t kernelEventsUsed = 0;

    int nOwned = 0, nGhost = 0;
    std::vector<int32_t> order, inv;   // order[p] = caller tet at device row p; inv = inverse
    bool identityOrder = true;
    std::vector<int32_t> nbrHost;      // device-order, device-row neighbour indices
    std::vector<double> area, volume, normal;   // caller order
    std::vector<int32_t> entity;                // caller order
    int32_t* orderDev = nullptr;
    int32_t* invDev = nullptr;
    // partitioned mesh (vt_mesh_set_ghost_geometry): what the Poisson assembly needs to know about
    // the ghost tets, and the reference's (global) tet index of every local row
    std::vector<int32_t> globalId;              // [nOwned + nGhost], caller order; empty = identity
    std::vector<int32_t> ghostNbr;              // [4 * nGhost] local caller index across each ghost face, -1 = not local
    std::vector<double> ghostArea, ghostNormal, ghostCentroid, ghostFaceCentroid;
    int globalRows = 0;                         // tets of the whole mesh (0 = nOwned)
    int globalDirichlet = -1;                   // partition: does ANY rank hold a Dirichlet field BC (-1 = decide locally)

    std::vector<vt::Species*> species;

    double* E = nullptr;       // device, 3 per owned tet, device order
    double* rho = nullptr;     // device order
    double* phi = nullptr;
    double* stage = nullptr;   // staging buffer for permuted host<->device copies
    size_t stageBytes = 0;
    double* pinned = nullptr;  // pinned host staging
    size_t pinnedBytes = 0;

    int chunkPlanes = 0, brickTets = 0;
    int variant = 64;   // vt_step_config bits; 64 = choose the step kernel from the velocity grid
    unsigned long long* workCounter = nullptr;   // device: heads of the persistent kernels' work queues (2)

    vt::PoissonData* poisson = nullptr;
    vt::Group* group = nullptr;   // device group (vt_ctx_create_group): this context only dispatches to its members

    // multi-GPU: flag words for the device-side barrier between steps (CUDA-IPC shared)
    int rank = 0, nPeers = 0;
    int peerRank[vt::kMaxPeers] = {};
    uint32_t* flags = nullptr;                       // [64] one word per source rank, local
    uint32_t* peerFlags[vt::kMaxPeers] = {};         // peers' flag arrays
    uint32_t epoch = 0;
    int* haloStatus = nullptr;                       // device view of haloStatusHost: != 0 when a barrier timed out (sticky)
    int* haloStatusHost = nullptr;                   // mapped host memory
    unsigned long long haloTimeoutNs = 20000000000ULL;   // VT_COMM_TIMEOUT_MS
    std::vector<void*> ipcOpened;
    void* haloTable = nullptr;                       // device copy of the peer flag pointers/ranks
};

extern "C" void vt_set_error(const char* msg);   // internal: sets vt_last_error()

namespace vt {
// device groups (group.cu): the C ABI entry points hand a group context over to these
void group_destroy(Group* g);
int group_mesh_upload(vt_ctx* ctx, int nTets, int nGhost, const int32_t* nbr, const double* area, const double* volume,
                      const double* normal, const int32_t* entity, const int32_t* order);
int group_species_create(vt_ctx* ctx, const int32_t n[3], const double vmin[3], const double vmax[3], double mass,
                         double charge, int* species);
int group_species_set_params(vt_ctx* ctx, int sp, double mass, double charge);
int group_species_set_face_bc(vt_ctx* ctx, int sp, const uint8_t* bcType, const uint8_t* collect, const int32_t* sourceId);
int group_species_set_source_pdfs(vt_ctx* ctx, int sp, int nSource, const double* pdf);
int group_species_set_pdf(vt_ctx* ctx, int sp, int first, int count, const double* pdf);
int group_species_get_pdf(vt_ctx* ctx, int sp, int first, int count, double* pdf);
int group_species_set_maxwell(vt_ctx* ctx, int sp, const double* physDensity, double temperature, const double mpv[3]);
int group_species_density(vt_ctx* ctx, int sp, double* density);
int group_species_velocity(vt_ctx* ctx, int sp, double* velocity);
int group_field_set(vt_ctx* ctx, const double* E);
int group_field_get(vt_ctx* ctx, double* rho, double* phi, double* E);
int group_step(vt_ctx* ctx, int sp, double dt, const double ext[3], bool tucker);
int group_wall_charge_get(vt_ctx* ctx, int sp, int entity, double* charge);
int group_wall_charge_reset(vt_ctx* ctx, int sp);
int group_charge_density(vt_ctx* ctx, const int* species, int nSpecies, const double* background);
int group_poisson_setup(vt_ctx* ctx, const double* tetCentroid, const double* faceCentroid, const uint8_t* bcType,
                        const double* bcValue, const double* bcNormalGrad);
int group_poisson_update_bc_values(vt_ctx* ctx, const double* bcValue, const double* bcNormalGrad);
int group_poisson_solve(vt_ctx* ctx, const double* rho, double* phi, double* E);
int group_poisson_stats(vt_ctx* ctx, int* its, double* res);
int group_tucker_enable(vt_ctx* ctx, int sp, double comprErr, int maxRank);
int group_tucker_get_factors(vt_ctx* ctx, int sp, int tet, int32_t ranks[3], double* core, double* u0, double* u1, double* u2);
int group_tucker_get_ranks(vt_ctx* ctx, int sp, int32_t* ranks);
int group_sync(vt_ctx* ctx);
long group_launch_count(vt_ctx* ctx);

void launch_full_step(vt_ctx* ctx, Species& sp, double dt, const double ext[3]);
void launch_density(vt_ctx* ctx, Species& sp);
// Tucker species: refresh the dense copy of the state in sp.f[sp.cur] (no-op when current)
void tucker_materialize(vt_ctx* ctx, Species& sp);
// Tucker species: re-compress the dense rows in sp.f[sp.cur] into the Tucker state (precision 0)
void tucker_from_dense(vt_ctx* ctx, Species& sp);
// Tucker species, partial uploads (vt_species_set_pdf streams a snapshot in batches): make the dense copy current
// before rows are written into it, and defer the re-compression until the state is used — one HOSVD pass over
// the tets for the whole upload instead of one per batch
void tucker_begin_dense_write(vt_ctx* ctx, Species& sp);
void tucker_end_dense_write(vt_ctx* ctx, Species& sp);
// Tucker species, multi-GPU: copy the current slots of the boundary tets into the peers' ghost rows
void tucker_push_current(vt_ctx* ctx, Species& sp);
void poisson_destroy(PoissonData* p);
void check_halo_status(vt_ctx* ctx);   // throws when a halo barrier of this context has timed out
double* ctx_stage(vt_ctx* ctx, size_t bytes);
double* ctx_pinned(vt_ctx* ctx, size_t bytes);
}  // namespace vt
