// Internal structures of the Tucker path shared by tucker.cu (general kernel, host side) and
// tucker_slab.cu (slab-streaming step kernel).  Not part of the ABI.
#pragma once

#include "vt_internal.h"

namespace vt {

constexpr int kMaxN = 64;   // nodes per axis served by the Tucker path

struct TuckerState {
    int rcap[3];                 // stored rank capacity per mode = min(maxRank, n)
    size_t coreCap, slot;        // doubles per tet: core, whole slot (core + 3 factors)
    void* block = nullptr;       // one allocation (one CUDA-IPC handle): buf[0] | buf[1] | ranks[0] | ranks[1]
    size_t rows = 0;             // owned + ghost rows
    double* buf[2] = {nullptr, nullptr};   // compressed state, ping-pong
    int* ranks[2] = {nullptr, nullptr};    // 3 per tet
    // multi-GPU: the peers' blocks, for the ghost copies of boundary tets
    double* peerBuf[kMaxPeers][2] = {};
    int* peerRanks[kMaxPeers][2] = {};
    int nPeers = 0;
    double* vnabs = nullptr;     // |v.n| per face as rank-<=6 Tucker tensors (solver.cpp:282): 4 slots per owned tet
    int* vnabsRanks = nullptr;   // 3 per (tet, face)
    size_t vslot = 0;            // doubles per slot: 6^3 core + 6 (n0 + n1 + n2)
    double* scratch = nullptr;   // per-CTA dense work space
    int scratchCTAs = 0;
    double comprErr = 1e-10;
    int maxRank = 0;
    int cur = 0;
    bool vnabsValid = false;
    bool densePending = false;   // rows were written into the dense copy (vt_species_set_pdf): re-compress before the next use
    int lastKernel = 0;          // 0 none yet, 1 k_tucker, 2 k_tucker_slab
    bool denseValid = false;     // sp.f[sp.cur] holds the reconstruction of buf[cur]
};

struct TuckerParams {
    int nOwned;
    int n[3], N;
    int rcap[3];
    size_t coreCap, slot;
    const double* in;      // compressed state at step n
    const int* rin;
    double* out;           // compressed state at step n+1
    int* rout;
    const TetRec* rec;
    const double* E;
    double* vnabs;         // [nOwned][4][vslot]: core 6^3, then U0 (n0 x 6), U1, U2
    int* vnabsRanks;       // [nOwned][4][3]
    size_t vslot;
    const double* src;     // dense source PDFs (Source BC), rows of N
    double* density;
    double* wall;
    double* scratch;       // per CTA: 5 N + 3 kMaxN*kMaxN(U work) doubles
    size_t scratchPerCTA;
    double vmin[3], step[3], inv2h[3];
    double qm, ext[3], dt, wallScale, cellVolume;
    double eps;
    int maxRank;
    const double* denseIn;   // set_pdf path: dense rows to compress (mode 1)
    double* denseOut;        // get_pdf path: dense rows reconstructed (mode 2)
    int first;
    int mode;                // 0 step, 1 compress dense input, 2 reconstruct, 3 |v.n| tables
    double epsAbs;           // mode 3: compression error for |v.n| (rank cap 6)
    long long* prof;         // optional: 8 phase timers in clock cycles (VT_TUCKER_PROFILE)
    int rK;                  // slab kernel: columns of the staged factor matrices (rank cap rounded up to 8)
    int gramDmma;            // 1: Gram matrices by mma.sync f64 (the default), 0: DFMA (VT_TUCKER_GRAM=dfma)
    double* peerOut[kMaxPeers];   // multi-GPU: peers' state buffers receiving the ghost copies
    int* peerRout[kMaxPeers];
};


// ---- general kernel k_tucker: launch geometry shared by tucker.cu (host) and tucker_inst.cu (one instantiation each)
constexpr int kThreads = 256;        // CTA size for velocity grids above 32 nodes per axis
constexpr int kThreadsSmall = 128;   // ... and up to 32: more tets in flight per SM hide the eigen-solver's latency
constexpr int kQB = 8;               // output rows per thread of a mode product
__host__ __device__ constexpr int pad_ld(int n)   // smallest ld >= roundup(n, 8) with ld % 16 == 4
{
    const int p = (n + 7) / 8 * 8;
    return p + ((4 - p % 16) + 16) % 16;
}
// doubles per staging tile buffer: the DFMA Gram / mode-product layouts or the zero-padded DMMA tiles
__host__ __device__ constexpr int tile_cap(int nmax)
{
    const int a = ((nmax | 1) > (nmax + kQB - 1) / kQB * kQB ? (nmax | 1) : (nmax + kQB - 1) / kQB * kQB) * nmax;
    const int b = pad_ld(nmax) * ((nmax + 7) / 8 * 8);
    return a > b ? a : b;
}
void launch_k_tucker_16(int grid, size_t smem, cudaStream_t stream, const TuckerParams& P);
void launch_k_tucker_32(int grid, size_t smem, cudaStream_t stream, const TuckerParams& P);
void launch_k_tucker_64(int grid, size_t smem, cudaStream_t stream, const TuckerParams& P);

// tucker_slab.cu: the slab-streaming step kernel (mode 0 only).  slab_eligible says whether it serves
// this grid / rank cap / compression error; launch_tucker_slab runs one step of all owned tets.
bool slab_eligible(const vt_ctx* ctx, const TuckerParams& P);
void launch_tucker_slab(vt_ctx* ctx, TuckerState& ts, TuckerParams& P);

}  // namespace vt
