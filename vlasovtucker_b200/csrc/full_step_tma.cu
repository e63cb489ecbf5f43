// K1 (TMA variant) — the fused full-format kinetic update as a persistent, bulk-async-copy
// pipelined kernel for sm_100a.  Same arithmetic and same results as full_step.cu (which stays
// as the general path for velocity grids this layout does not cover); see that file for the
// reference citations of each term.
//
// Why: the register-staged kernel is latency bound — ~120 registers per thread cap it at 16
// warps per SM, each with nine 512-byte loads in flight, which covers only about a third of
// HBM's bandwidth-latency product (profiles/round1_notes.md).  Here the loads are issued by one
// thread per CTA as bulk async copies (cp.async.bulk, the 1-D TMA path; SASS UBLKCP) into a
// shared-memory ring, completion is tracked by mbarriers, and no register is tied up by data
// in flight: two to three complete plane sets (own plane + 4 neighbour planes, 40 KiB each at
// 32x32) are always on their way while the 512 threads compute from shared memory.
//
// Work: one CTA per SM, persistent; work items (tet, chunk of i2-planes) are dealt round-robin
// in the same brick-major order as the general kernel, so the CTAs in flight still share an L2
// working set.  A thread owns KPT fixed (i0-pair, i1) columns of the plane and marches along
// i2 keeping the own-row values prev/cur/next in registers; the i0 and i1 stencil neighbours and
// the four neighbour tets' values come from shared memory.
#include "vt_internal.h"

namespace vt {

namespace {

constexpr int kOwnRing = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ bool is_pair(int bc)
{
    return bc == VT_PBC_NONBOUNDARY || bc == VT_PBC_PERIODIC || bc == VT_PBC_SOURCE;
}

struct Item {
    int tet, chunk, pl0, npl;
};

__device__ __forceinline__ bool decode_item(const StepParams& p, long long w, long long total, Item& it)
{
    if (w >= total) return false;
    const int perBrick = p.brickTets * p.nChunks;
    const int brick = (int)(w / perBrick);
    const int base = brick * p.brickTets;
    const int nb = min(p.brickTets, p.nOwned - base);
    const int r = (int)(w - (long long)brick * perBrick);
    it.chunk = r / nb;
    it.tet = base + (r - it.chunk * nb);
    it.pl0 = it.chunk * p.chunkPlanes;
    it.npl = min(p.chunkPlanes, p.n2 - it.pl0);
    return true;
}

struct TmaParams {
    StepParams s;
    int S;            // neighbour ring depth
    int planeElems;   // n0*n1
    int PV;           // double2 per plane
};

struct Smem {
    double* ownRing;     // [kOwnRing][PE]
    double* nbrRing;     // [S][4][PE]
    uint64_t* barOwn;    // [kOwnRing]
    uint64_t* barNbr;    // [S]
    TetRec* rec;         // [2]
    double* tz;          // [2][n2][4]
    double* red;         // [16][5]
};

// Producer cursor: which stage of which item is issued next, and how far both rings are filled.
struct Producer {
    long long pw;        // work index of the item being issued
    int pk;              // its ordinal in this CTA's sequence
    int ps;              // next stage of that item (0 .. npl+1)
    Item it;
    unsigned ownIssued, nbrIssued, ownFreed, nbrFreed;
};

// Issue stages in stream order while both rings have room; never run past item kConsumer+1
// (only two tet records are resident).  Stage s of an item carries own plane pl0-1+s and, for
// 1 <= s <= npl, the four neighbour planes pl0+s-1.
__device__ __noinline__ void produce(const TmaParams& P, const Smem& sm, Producer& pr, int kConsumer, long long total)
{
    const StepParams& p = P.s;
    const int PE = P.planeElems;
    const uint32_t PB = (uint32_t)PE * 8u;
    while (pr.pw < total && pr.pk <= kConsumer + 1) {
        const bool needNbr = pr.ps >= 1 && pr.ps <= pr.it.npl;
        if (pr.ownIssued - pr.ownFreed >= (unsigned)kOwnRing) break;
        if (needNbr && pr.nbrIssued - pr.nbrFreed >= (unsigned)P.S) break;
        const TetRec& r = sm.rec[pr.pk & 1];
        int ip = pr.it.pl0 - 1 + pr.ps;                  // periodic in v2 (solver.cpp:380-389)
        if (ip < 0) ip += p.n2;
        if (ip >= p.n2) ip -= p.n2;
        {
            const unsigned slot = pr.ownIssued % kOwnRing;
            mbar_expect_tx(sm.barOwn + slot, PB);
            bulk_g2s(sm.ownRing + (size_t)slot * PE, p.f + (size_t)pr.it.tet * p.N + (size_t)ip * PE, PB, sm.barOwn + slot);
            pr.ownIssued++;
        }
        if (needNbr) {
            const unsigned slot = pr.nbrIssued % P.S;
            uint64_t* bar = sm.barNbr + slot;
            int cnt = 0;
            for (int f = 0; f < 4; f++) cnt += is_pair(r.bc[f]) ? 1 : 0;
            mbar_expect_tx(bar, PB * cnt);
            const int plane = pr.it.pl0 + pr.ps - 1;
            for (int f = 0; f < 4; f++) {
                if (!is_pair(r.bc[f])) continue;
                const int n = r.nbr[f];
                const double* row = n >= 0 ? p.f + (size_t)n * p.N : p.src + (size_t)(-2 - n) * p.N;
                bulk_g2s(sm.nbrRing + ((size_t)slot * 4 + f) * PE, row + (size_t)plane * PE, PB, bar);
            }
            pr.nbrIssued++;
        }
        pr.ps++;
        if (pr.ps > pr.it.npl + 1) {
            pr.pw += gridDim.x;
            pr.pk++;
            pr.ps = 0;
            if (pr.pw < total) decode_item(p, pr.pw, total, pr.it);
        }
    }
}

// All planes of one work item.  GENERIC: per-face boundary conditions and halo push (branches
// are uniform across the CTA); otherwise four pair faces and no push.
template <int KPT, bool UPWIND, bool GENERIC>
__device__ __forceinline__ void item_compute(const TmaParams& P, const Smem& sm, Producer& pr, const Item& cur,
                                             const int k, const unsigned ownBase, const unsigned nbrBase,
                                             const long long total, const int (&colV)[KPT], const int (&colI0)[KPT],
                                             const int (&colI1)[KPT], const bool (&colOn)[KPT], double& accDens,
                                             double (&accWall)[4])
{
    const StepParams& p = P.s;
    const int PE = P.planeElems;
    const int tid = threadIdx.x;
    const TetRec& rec = sm.rec[k & 1];
    const double* tzk = sm.tz + (size_t)(k & 1) * 4 * p.n2;

    double cxy[KPT][4][2], hc[4];
    bool pairF[4], absF[4], colF[4];
#pragma unroll
    for (int f = 0; f < 4; f++) {
        const int bc = GENERIC ? rec.bc[f] : VT_PBC_NONBOUNDARY;
        pairF[f] = is_pair(bc);
        absF[f] = GENERIC && bc == VT_PBC_ABSORBING;
        colF[f] = GENERIC && rec.wallSlot[f] >= 0;
        hc[f] = pairF[f] ? 0.5 * rec.coef[f] : rec.coef[f];
        const double pre = (UPWIND && pairF[f]) ? rec.coef[f] : 1.0;
#pragma unroll
        for (int kk = 0; kk < KPT; kk++) {
            const double v1 = __dadd_rn(p.vmin[1], __dmul_rn((double)colI1[kk], p.step[1]));
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const double v0 = __dadd_rn(p.vmin[0], __dmul_rn((double)(colI0[kk] + u), p.step[0]));
                cxy[kk][f][u] = pre * (rec.nrm[f][0] * v0 + rec.nrm[f][1] * v1);
            }
        }
    }
    double g[3];
#pragma unroll
    for (int q = 0; q < 3; q++) g[q] = (p.qm * (p.E[3 * (size_t)cur.tet + q] + p.ext[q])) * p.inv2h[q];
    double* nrow = p.fn + (size_t)cur.tet * p.N;
    double* push[4] = {nullptr, nullptr, nullptr, nullptr};
    if (GENERIC) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (rec.pushPeer[q] >= 0) push[q] = p.peerFn[rec.pushPeer[q]] + (size_t)rec.pushRow[q] * p.N;
    }

    // first two own stages: prev, cur
    double2 prv[KPT], cr[KPT];
    {
        const unsigned g0 = ownBase, g1 = ownBase + 1;
        mbar_wait(sm.barOwn + (g0 % kOwnRing), (g0 / kOwnRing) & 1);
        mbar_wait(sm.barOwn + (g1 % kOwnRing), (g1 / kOwnRing) & 1);
        const double* s0 = sm.ownRing + (size_t)(g0 % kOwnRing) * PE;
        const double* s1 = sm.ownRing + (size_t)(g1 % kOwnRing) * PE;
#pragma unroll
        for (int kk = 0; kk < KPT; kk++) {
            prv[kk] = *reinterpret_cast<const double2*>(s0 + 2 * colV[kk]);
            cr[kk] = *reinterpret_cast<const double2*>(s1 + 2 * colV[kk]);
        }
    }

    for (int j = 0; j < cur.npl; j++) {
        const unsigned gc = ownBase + j + 1, gn = ownBase + j + 2, cn = nbrBase + j;
        mbar_wait(sm.barOwn + (gn % kOwnRing), (gn / kOwnRing) & 1);
        mbar_wait(sm.barNbr + (cn % P.S), (cn / P.S) & 1);
        const double* sc = sm.ownRing + (size_t)(gc % kOwnRing) * PE;       // plane j: i0/i1 neighbours
        const double* sn = sm.ownRing + (size_t)(gn % kOwnRing) * PE;       // plane j+1
        const double* sb = sm.nbrRing + (size_t)(cn % P.S) * 4 * PE;
        const double tzf[4] = {tzk[4 * j], tzk[4 * j + 1], tzk[4 * j + 2], tzk[4 * j + 3]};
        const size_t gplane = (size_t)(cur.pl0 + j) * PE;
#pragma unroll
        for (int kk = 0; kk < KPT; kk++) {
            if (!colOn[kk]) continue;
            const int ev = 2 * colV[kk];
            const int i0 = colI0[kk], i1 = colI1[kk];
            const double2 nx = *reinterpret_cast<const double2*>(sn + ev);
            const double2 um = *reinterpret_cast<const double2*>(sc + ev + ((i1 == 0) ? (p.n1 - 1) : -1) * p.n0);
            const double2 up = *reinterpret_cast<const double2*>(sc + ev + ((i1 == p.n1 - 1) ? -(p.n1 - 1) : 1) * p.n0);
            const double fl = sc[ev + ((i0 == 0) ? (p.n0 - 1) : -1)];
            const double fr = sc[ev + 1 + ((i0 + 2 == p.n0) ? -(p.n0 - 1) : 1)];
            double2 fa[4];
#pragma unroll
            for (int f = 0; f < 4; f++)
                fa[f] = (!GENERIC || pairF[f]) ? *reinterpret_cast<const double2*>(sb + (size_t)f * PE + ev)
                                               : make_double2(0.0, 0.0);
            const double fcv[2] = {cr[kk].x, cr[kk].y};
            const double xm[2] = {fl, cr[kk].x};
            const double xp[2] = {cr[kk].y, fr};
            const double y1m[2] = {um.x, um.y}, y1p[2] = {up.x, up.y};
            const double z2m[2] = {prv[kk].x, prv[kk].y}, z2p[2] = {nx.x, nx.y};
            double out[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const double fv = fcv[u];
                double rhs = 0.0;
#pragma unroll
                for (int f = 0; f < 4; f++) {
                    const double vn = cxy[kk][f][u] + tzf[f];
                    const double fau = u == 0 ? fa[f].x : fa[f].y;
                    if (!GENERIC || pairF[f]) {
                        if (UPWIND) {
                            rhs = fma(-vn, vn > 0.0 ? fv : fau, rhs);
                        } else {
                            const double s = fau + fv, d = fau - fv;
                            rhs = fma(-hc[f], fma(vn, s, -(fabs(vn) * d)), rhs);
                        }
                    } else if (absF[f]) {
                        const double flux = 0.5 * (vn * fv + fabs(vn) * fv);
                        if (colF[f]) accWall[f] += flux;
                        rhs = fma(-hc[f], flux, rhs);
                    } else {
                        rhs = fma(-hc[f], vn * fv, rhs);
                    }
                }
                rhs = fma(-g[0], xp[u] - xm[u], rhs);
                rhs = fma(-g[1], y1p[u] - y1m[u], rhs);
                rhs = fma(-g[2], z2p[u] - z2m[u], rhs);
                out[u] = fma(p.dt, rhs, fv);
                accDens += out[u];
            }
            const double2 o = make_double2(out[0], out[1]);
            *reinterpret_cast<double2*>(nrow + gplane + ev) = o;
            if (GENERIC) {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (push[q]) *reinterpret_cast<double2*>(push[q] + gplane + ev) = o;
            }
            prv[kk] = cr[kk];
            cr[kk] = nx;
        }
        __syncthreads();   // every thread is done with own stage gc and neighbour slot cn
        if (tid == 0) {
            pr.ownFreed = ownBase + ((j == cur.npl - 1) ? cur.npl + 2 : j + 2);
            pr.nbrFreed = nbrBase + j + 1;
            produce(P, sm, pr, k, total);
        }
    }
}

template <int KPT, bool UPWIND>
__global__ void __launch_bounds__(512, 1) k_full_step_tma(const TmaParams P)
{
    const StepParams& p = P.s;
    extern __shared__ __align__(128) unsigned char smraw[];
    const int PE = P.planeElems;
    Smem sm;
    sm.ownRing = reinterpret_cast<double*>(smraw);
    sm.nbrRing = sm.ownRing + (size_t)kOwnRing * PE;
    sm.barOwn = reinterpret_cast<uint64_t*>(sm.nbrRing + (size_t)P.S * 4 * PE);
    sm.barNbr = sm.barOwn + kOwnRing;
    sm.rec = reinterpret_cast<TetRec*>(sm.barNbr + 8);
    sm.tz = reinterpret_cast<double*>(sm.rec + 2);
    sm.red = sm.tz + 2 * 4 * p.n2;

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const long long total = (long long)p.nOwned * p.nChunks;
    const int G = gridDim.x;

    // fixed columns of this thread
    int colV[KPT], colI0[KPT], colI1[KPT];
    bool colOn[KPT];
#pragma unroll
    for (int kk = 0; kk < KPT; kk++) {
        const int v = tid + kk * nthr;
        colOn[kk] = v < P.PV;
        colV[kk] = colOn[kk] ? v : 0;
        colI0[kk] = (colV[kk] % p.nvec0) * 2;
        colI1[kk] = colV[kk] / p.nvec0;
    }

    if (tid == 0) {
        for (int i = 0; i < kOwnRing; i++) mbar_init(sm.barOwn + i, 1);
        for (int i = 0; i < P.S; i++) mbar_init(sm.barNbr + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    Item cur, nxt;
    const bool haveCur = decode_item(p, blockIdx.x, total, cur);
    bool haveNxt = decode_item(p, (long long)blockIdx.x + G, total, nxt);
    auto load_rec = [&](int slot, int tet) {
        const int* gsrc = reinterpret_cast<const int*>(p.rec + tet);
        int* s = reinterpret_cast<int*>(sm.rec + slot);
        for (int i = tid; i < (int)(sizeof(TetRec) / 4); i += nthr) s[i] = gsrc[i];
    };
    if (haveCur) load_rec(0, cur.tet);
    if (haveNxt) load_rec(1, nxt.tet);
    __syncthreads();
    if (!haveCur) return;

    Producer pr;
    pr.pw = blockIdx.x;
    pr.pk = 0;
    pr.ps = 0;
    pr.it = cur;
    pr.ownIssued = pr.nbrIssued = pr.ownFreed = pr.nbrFreed = 0;
    if (tid == 0) produce(P, sm, pr, 0, total);

    unsigned ownBase = 0, nbrBase = 0;   // global stage / compute counters at the start of the item
    int k = 0;
    long long w = blockIdx.x;
    while (true) {
        const TetRec& rec = sm.rec[k & 1];
        // per-plane part of v.n for this item: (A/V) n_z v2(i2)
        double* tzk = sm.tz + (size_t)(k & 1) * 4 * p.n2;
        for (int i = tid; i < 4 * cur.npl; i += nthr) {
            const int f = i & 3, pl = i >> 2;
            const double v2 = __dadd_rn(p.vmin[2], __dmul_rn((double)(cur.pl0 + pl), p.step[2]));
            const double pre = (UPWIND && is_pair(rec.bc[f])) ? rec.coef[f] : 1.0;
            tzk[4 * pl + f] = pre * (rec.nrm[f][2] * v2);
        }
        __syncthreads();
        const bool fast = is_pair(rec.bc[0]) && is_pair(rec.bc[1]) && is_pair(rec.bc[2]) && is_pair(rec.bc[3]) &&
                          rec.pushPeer[0] < 0;
        double accDens = 0.0;
        double accWall[4] = {0.0, 0.0, 0.0, 0.0};
        if (fast) item_compute<KPT, UPWIND, false>(P, sm, pr, cur, k, ownBase, nbrBase, total, colV, colI0, colI1, colOn, accDens, accWall);
        else item_compute<KPT, UPWIND, true>(P, sm, pr, cur, k, ownBase, nbrBase, total, colV, colI0, colI1, colOn, accDens, accWall);

        // ---- item epilogue: sum_v f' (Density) and absorbed flux (wall charge)
        const bool anyWall = (rec.wallSlot[0] >= 0) | (rec.wallSlot[1] >= 0) | (rec.wallSlot[2] >= 0) | (rec.wallSlot[3] >= 0);
        const int warp = tid >> 5, lane = tid & 31;
        const double sd = warp_sum(accDens);
        if (lane == 0) sm.red[warp * 5] = sd;
        if (anyWall) {
#pragma unroll
            for (int f = 0; f < 4; f++) {
                const double wv = warp_sum(accWall[f]);
                if (lane == 0) sm.red[warp * 5 + 1 + f] = wv;
            }
        }
        ownBase += cur.npl + 2;
        nbrBase += cur.npl;
        const Item done = cur;
        w += G;
        k++;
        const bool more = haveNxt;
        if (more) {
            cur = nxt;
            haveNxt = decode_item(p, w + G, total, nxt);
        }
        __syncthreads();
        if (tid == 0) {
            const int nw = nthr >> 5;
            double d = 0.0;
            for (int q = 0; q < nw; q++) d += sm.red[q * 5];
            p.densPartial[(size_t)done.tet * p.nChunks + done.chunk] = d;
            if (anyWall) {
                const TetRec& rd = sm.rec[(k - 1) & 1];
                for (int f = 0; f < 4; f++) {
                    if (rd.wallSlot[f] < 0) continue;
                    double q2 = 0.0;
                    for (int q = 0; q < nw; q++) q2 += sm.red[q * 5 + 1 + f];
                    // charge * (timeStep * area * flux.Sum() * cellVolume), solver.cpp:173-177
                    atomicAdd(p.wall + rd.wallSlot[f], p.wallScale * rd.area[f] * q2);
                }
            }
        }
        if (!more) break;
        __syncthreads();                       // red[] and rec[(k-1)&1] are free again
        if (haveNxt) load_rec((k + 1) & 1, nxt.tet);
        __syncthreads();
        if (tid == 0) produce(P, sm, pr, k, total);
    }
}

}  // namespace

// Returns false when the velocity grid does not fit this layout (caller falls back).
bool launch_full_step_tma(vt_ctx* ctx, Species& sp, StepParams& p, bool upwind, cudaEvent_t e0, cudaEvent_t e1)
{
    const int n0 = sp.n[0], n1 = sp.n[1], n2 = sp.n[2];
    if (n0 % 2) return false;
    const int PE = n0 * n1;
    if ((PE * 8) % 16) return false;
    const int PV = PE / 2;
    int nthr, kpt;
    if (PV <= 512) {
        nthr = ((PV + 31) / 32) * 32;
        if (nthr < 128) nthr = 128;
        kpt = 1;
    } else if (PV <= 1024) {
        nthr = ((PV / 2 + 31) / 32) * 32;
        kpt = 2;
    } else if (PV <= 1536) {
        nthr = ((PV / 3 + 31) / 32) * 32;
        kpt = 3;
    } else {
        return false;
    }
    if (nthr * kpt < PV) nthr += 32;
    if (nthr > 512) return false;
    const size_t PB = (size_t)PE * 8;
    const size_t fixed = (kOwnRing + 8) * 8 + 2 * sizeof(TetRec) + (size_t)2 * 4 * n2 * 8 + 16 * 5 * 8 + 256;
    const size_t maxSmem = 227 * 1024;
    int S = 4;
    while (S >= 2 && kOwnRing * PB + (size_t)S * 4 * PB + fixed > maxSmem) S--;
    if (S < 2) return false;
    const size_t smem = kOwnRing * PB + (size_t)S * 4 * PB + fixed;

    TmaParams P;
    P.s = p;
    P.S = S;
    P.planeElems = PE;
    P.PV = PV;
    const long long total = (long long)ctx->nOwned * p.nChunks;
    const int grid = (int)std::min<long long>(total, ctx->prop.multiProcessorCount);

    auto launch = [&](auto kern) {
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VT_CUDA(cudaEventRecord(e0, ctx->stream));
        kern<<<grid, nthr, smem, ctx->stream>>>(P);
        VT_CUDA(cudaEventRecord(e1, ctx->stream));
    };
    if (kpt == 1) upwind ? launch(k_full_step_tma<1, true>) : launch(k_full_step_tma<1, false>);
    else if (kpt == 2) upwind ? launch(k_full_step_tma<2, true>) : launch(k_full_step_tma<2, false>);
    else upwind ? launch(k_full_step_tma<3, true>) : launch(k_full_step_tma<3, false>);
    VT_CUDA(cudaGetLastError());
    return true;
}

}  // namespace vt
