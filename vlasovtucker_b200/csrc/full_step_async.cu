// K1 (async-pipeline variant) — the fused full-format kinetic update as a persistent kernel whose
// loads are cp.async copies into a shared-memory ring.  Same arithmetic and results as
// full_step.cu (which remains the general path and carries the reference citations).
//
// Why (profiles/round1_notes.md): the register-staged kernel is latency bound.  ~120 registers
// per thread cap it at 16 warps/SM with nine 512-byte loads each, about 74 KB in flight per SM,
// which sustains ~6 TB/s of L2->SM traffic for the 48 B/update this stencil moves through L2 —
// a third of what the 21-28 TB/s L2 can deliver and too little to saturate HBM.  Registers are
// the limiter, so here no register holds data in flight: every thread issues 16-byte cp.async
// (LDGSTS) copies for planes D tickets ahead and computes from shared memory.  This kernel is
// superseded as the default by the warp-specialised cp.async.bulk pipeline of full_step_tma.cu
// (a dedicated producer thread, mbarrier rings, no block-wide barrier: 18.6 ms against 28.8-36 ms
// here on the C4 share) and stays as an opt-in variant (vt_step_config bit 3) and a parity case.
//
// Structure: one CTA per SM, persistent.  Work items (tet, chunk of i2-planes) are handed out in
// brick-major order through a global atomic counter, so all SMs stay inside a narrow window of
// the locality order (static round-robin lets CTAs drift apart and destroys L2 reuse).  An item
// of npl planes is npl+2 "tickets": ticket t loads own plane pl0-1+t (periodic) and, for
// 1 <= t <= npl, the four neighbour planes pl0+t-1; at ticket t >= 2 plane t-2 is computed.  A
// thread owns KPT fixed (i0-pair, i1) columns and marches along i2 with prev/cur/next in
// registers; i0/i1 stencil neighbours and the neighbour tets' values come from shared memory.
// One __syncthreads per ticket both publishes the landed copies and retires the ring slot that
// the next copies overwrite (ring depth D+2).
#include "vt_internal.h"

#include <cstdio>
#include <cstdlib>

namespace vt {

namespace {

constexpr int kItemRing = 8;    // fetched-ahead work items (descriptor + tet record + field)
constexpr int kRecBytes = 224;  // sizeof(TetRec)
constexpr int kLead = 4;        // descriptors fetched beyond the loader's item

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
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

struct ItemDesc {
    int tet, chunk, pl0, npl;   // npl == 0: no more work
};

struct AsyncParams {
    StepParams s;
    int planeElems;   // n0*n1
    int PV;           // double2 per plane
    unsigned long long* counter;   // work queue head (zeroed before the launch)
};

__device__ __forceinline__ void decode_item(const StepParams& p, long long w, long long total, ItemDesc& it)
{
    if (w >= total) {
        it.tet = 0;
        it.chunk = 0;
        it.pl0 = 0;
        it.npl = 0;
        return;
    }
    const int perBrick = p.brickTets * p.nChunks;
    const int brick = (int)(w / perBrick);
    const int base = brick * p.brickTets;
    const int nb = min(p.brickTets, p.nOwned - base);
    const int r = (int)(w - (long long)brick * perBrick);
    it.chunk = r / nb;
    it.tet = base + (r - it.chunk * nb);
    it.pl0 = it.chunk * p.chunkPlanes;
    it.npl = min(p.chunkPlanes, p.n2 - it.pl0);
}

// Per-item constants of the consumer, held in registers for the npl planes of the item.
template <int KPT>
struct ItemRegs {
    double cxy[KPT][4][2];   // (A/V)(n_x v0 + n_y v1) per column/face/element (A/V only in UPWIND pair form)
    double cz[4];            // (A/V) n_z
    double hc[4];            // 0.5 A/V (pair) or A/V (wall)
    double g[3];             // (q/m)(E+ext)/(2 step)
    double* nrow;
    double* push[4];
    unsigned pairMask, absMask, colMask;
};

// One plane of one item.  GENERIC: per-face boundary conditions and halo push (uniform branches).
template <int KPT, bool UPWIND, bool GENERIC>
__device__ __forceinline__ void compute_plane(const StepParams& p, const ItemRegs<KPT>& it, const double* sc,
                                              const double* sn, const double* sb, const int PE, const int plane,
                                              const int (&colE)[KPT], const int (&offU)[KPT], const int (&offD)[KPT],
                                              const int (&offL)[KPT], const int (&offR)[KPT], const bool (&colOn)[KPT],
                                              double2 (&prv)[KPT], double2 (&cr)[KPT], double& accDens,
                                              double (&accWall)[4])
{
    const double v2 = __dadd_rn(p.vmin[2], __dmul_rn((double)plane, p.step[2]));
    double tz[4];
#pragma unroll
    for (int f = 0; f < 4; f++) tz[f] = it.cz[f] * v2;
    const size_t gplane = (size_t)plane * PE;
#pragma unroll
    for (int kk = 0; kk < KPT; kk++) {
        if (!colOn[kk]) continue;
        const int ev = colE[kk];
        const double2 nx = *reinterpret_cast<const double2*>(sn + ev);
        const double2 um = *reinterpret_cast<const double2*>(sc + ev + offD[kk]);
        const double2 up = *reinterpret_cast<const double2*>(sc + ev + offU[kk]);
        const double fl = sc[ev + offL[kk]];
        const double fr = sc[ev + offR[kk]];
        double2 fa[4];
#pragma unroll
        for (int f = 0; f < 4; f++)
            fa[f] = (!GENERIC || ((it.pairMask >> f) & 1)) ? *reinterpret_cast<const double2*>(sb + f * PE + ev)
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
                const double vn = it.cxy[kk][f][u] + tz[f];
                const double fau = u == 0 ? fa[f].x : fa[f].y;
                if (!GENERIC || ((it.pairMask >> f) & 1)) {
                    if (UPWIND) {
                        rhs = fma(-vn, vn > 0.0 ? fv : fau, rhs);
                    } else {
                        const double s = fau + fv, d = fau - fv;
                        rhs = fma(-it.hc[f], fma(vn, s, -(fabs(vn) * d)), rhs);
                    }
                } else if ((it.absMask >> f) & 1) {
                    const double flux = 0.5 * (vn * fv + fabs(vn) * fv);
                    if ((it.colMask >> f) & 1) accWall[f] += flux;
                    rhs = fma(-it.hc[f], flux, rhs);
                } else {
                    rhs = fma(-it.hc[f], vn * fv, rhs);
                }
            }
            rhs = fma(-it.g[0], xp[u] - xm[u], rhs);
            rhs = fma(-it.g[1], y1p[u] - y1m[u], rhs);
            rhs = fma(-it.g[2], z2p[u] - z2m[u], rhs);
            out[u] = fma(p.dt, rhs, fv);
            accDens += out[u];
        }
        const double2 o = make_double2(out[0], out[1]);
        *reinterpret_cast<double2*>(it.nrow + gplane + ev) = o;
        if (GENERIC) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (it.push[q]) *reinterpret_cast<double2*>(it.push[q] + gplane + ev) = o;
        }
        prv[kk] = cr[kk];
        cr[kk] = nx;
    }
}

template <int KPT, int D, bool UPWIND>
__global__ void __launch_bounds__(KPT == 1 ? 512 : 256, 1) k_full_step_async(const AsyncParams P)
{
    const StepParams& p = P.s;
    extern __shared__ __align__(128) unsigned char smraw[];
    const int PE = P.planeElems;
    constexpr int R = D + 2;
    double* ownRing = reinterpret_cast<double*>(smraw);                         // [R][PE]
    double* nbrRing = ownRing + (size_t)R * PE;                                  // [R][4][PE]
    unsigned char* recRaw = reinterpret_cast<unsigned char*>(nbrRing + (size_t)R * 4 * PE);   // [kItemRing][kRecBytes]
    double* fieldS = reinterpret_cast<double*>(recRaw + kItemRing * kRecBytes);  // [kItemRing][4]
    ItemDesc* itemS = reinterpret_cast<ItemDesc*>(fieldS + kItemRing * 4);       // [kItemRing]
    double* red = reinterpret_cast<double*>(itemS + kItemRing);                  // [2][16][5]
    const uint32_t ownRing32 = smem_u32(ownRing), nbrRing32 = smem_u32(nbrRing);
    const uint32_t PB = (uint32_t)PE * 8u;

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const long long total = (long long)p.nOwned * p.nChunks;

    // fixed columns of this thread and their stencil offsets (periodic in velocity space)
    int colE[KPT], offU[KPT], offD[KPT], offL[KPT], offR[KPT], colI0[KPT], colI1[KPT];
    bool colOn[KPT];
#pragma unroll
    for (int kk = 0; kk < KPT; kk++) {
        const int v = tid + kk * nthr;
        colOn[kk] = v < P.PV;
        const int vv = colOn[kk] ? v : 0;
        colE[kk] = 2 * vv;
        colI0[kk] = (vv % p.nvec0) * 2;
        colI1[kk] = vv / p.nvec0;
        offD[kk] = ((colI1[kk] == 0) ? (p.n1 - 1) : -1) * p.n0;
        offU[kk] = ((colI1[kk] == p.n1 - 1) ? -(p.n1 - 1) : 1) * p.n0;
        offL[kk] = (colI0[kk] == 0) ? (p.n0 - 1) : -1;
        offR[kk] = 1 + ((colI0[kk] + 2 == p.n0) ? -(p.n0 - 1) : 1);
    }

    // ---- work-item ring: descriptors are fetched kLead items ahead of the loader; the tet record
    // and the field of a fetched item follow by cp.async in the next iteration's group
    int fetched = 0, recIssued = 0;
    auto fetch_one = [&]() {   // thread 0 only
        const unsigned long long w = atomicAdd(P.counter, 1ULL);
        ItemDesc d;
        decode_item(p, (long long)w, total, d);
        itemS[fetched % kItemRing] = d;
    };
    auto issue_rec = [&](int ord) {   // all threads; the copies join the caller's commit group
        const ItemDesc d = itemS[ord % kItemRing];
        if (d.npl == 0) return;
        if (tid < kRecBytes / 16)
            cp_async16(smem_u32(recRaw + (size_t)(ord % kItemRing) * kRecBytes + 16 * tid),
                       reinterpret_cast<const unsigned char*>(p.rec + d.tet) + 16 * tid);
        else if (tid >= 32 && tid < 35)
            cp_async8(smem_u32(fieldS + (ord % kItemRing) * 4 + (tid - 32)), p.E + 3 * (size_t)d.tet + (tid - 32));
    };
    if (tid == 0)
        for (int i = 0; i < kLead; i++) {
            fetch_one();
            fetched++;
        }
    fetched = kLead;
    __syncthreads();
    for (; recIssued < kLead; recIssued++) issue_rec(recIssued);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (itemS[0].npl == 0) return;

    // ---- loader cursor (uniform across the CTA); per-item values cached in registers
    int lOrd = 0, lT = 0, lPl0 = 0, lNpl = 0;
    int lOwnSlot = 0, lNbrSlot = 0;
    bool lDone = false;
    const double* lOwn = nullptr;
    auto loader_enter = [&]() {
        const ItemDesc d = itemS[lOrd % kItemRing];
        if (d.npl == 0) {
            lDone = true;
            return;
        }
        lPl0 = d.pl0;
        lNpl = d.npl;
        lT = 0;
        lOwn = p.f + (size_t)d.tet * p.N;
    };
    auto load_ticket = [&]() {
        if (lDone) return;
        int ip = lPl0 - 1 + lT;                           // periodic in v2 (solver.cpp:380-389)
        if (ip < 0) ip += p.n2;
        if (ip >= p.n2) ip -= p.n2;
        const double* osrc = lOwn + (size_t)ip * PE;
        const uint32_t odst = ownRing32 + (uint32_t)lOwnSlot * PB;
#pragma unroll
        for (int kk = 0; kk < KPT; kk++)
            if (colOn[kk]) cp_async16(odst + 8u * colE[kk], osrc + colE[kk]);
        if (lT >= 1 && lT <= lNpl) {
            const size_t poff = (size_t)(lPl0 + lT - 1) * PE;
            const uint32_t ndst = nbrRing32 + (uint32_t)lNbrSlot * 4u * PB;
            // neighbour rows are looked up in the (shared-memory) tet record at every ticket: once
            // per plane, and it keeps no per-item pointer state alive across the item switch
            const TetRec& r = *reinterpret_cast<const TetRec*>(recRaw + (size_t)(lOrd % kItemRing) * kRecBytes);
#pragma unroll
            for (int f = 0; f < 4; f++) {
                if (!is_pair(r.bc[f])) continue;
                const int n = r.nbr[f];
                const double* row = (n >= 0 ? p.f + (size_t)n * p.N : p.src + (size_t)(-2 - n) * p.N) + poff;
#pragma unroll
                for (int kk = 0; kk < KPT; kk++)
                    if (colOn[kk]) cp_async16(ndst + (uint32_t)f * PB + 8u * colE[kk], row + colE[kk]);
            }
            lNbrSlot = (lNbrSlot + 1 == R) ? 0 : lNbrSlot + 1;
        }
        lOwnSlot = (lOwnSlot + 1 == R) ? 0 : lOwnSlot + 1;
        lT++;
        if (lT > lNpl + 1) {
            lOrd++;
            loader_enter();
        }
    };
    loader_enter();
    for (int i = 0; i < D; i++) {   // prologue: D tickets in flight before the first compute
        load_ticket();
        cp_async_commit();
    }

    // ---- consumer cursor
    int cOrd = 0, cT = 0;
    int cSlot = 0, cPrevSlot = R - 1;        // own-ring slots of this ticket and the previous one
    int cNbrNext = 0, cNbrLast = R - 1;      // neighbour-ring slot of the next / most recent stage
    ItemDesc cur = itemS[0];
    ItemRegs<KPT> it;
    bool generic = false;
    double2 prv[KPT], cr[KPT];
    double accDens = 0.0, accWall[4] = {0.0, 0.0, 0.0, 0.0};
    int pendingEpilogue = -1;    // item ordinal whose partial sums wait in red[] for thread 0
    ItemDesc pendingItem = cur;

    auto finish_pending = [&]() {   // thread 0, after a sync that published red[]
        const double* rd = red + (size_t)(pendingEpilogue & 1) * 16 * 5;
        const int nw = nthr >> 5;
        double dsum = 0.0;
        for (int q = 0; q < nw; q++) dsum += rd[q * 5];
        p.densPartial[(size_t)pendingItem.tet * p.nChunks + pendingItem.chunk] = dsum;
        const TetRec& rr = *reinterpret_cast<const TetRec*>(recRaw + (size_t)(pendingEpilogue % kItemRing) * kRecBytes);
        for (int f = 0; f < 4; f++) {
            if (rr.wallSlot[f] < 0) continue;
            double q2 = 0.0;
            for (int q = 0; q < nw; q++) q2 += rd[q * 5 + 1 + f];
            atomicAdd(p.wall + rr.wallSlot[f], p.wallScale * rr.area[f] * q2);   // solver.cpp:173-177
        }
    };

    while (true) {
        cp_async_wait<D - 1>();
        __syncthreads();
        // this iteration's copy group: records of items published before the sync, then ticket +D
        for (; recIssued < fetched; recIssued++) issue_rec(recIssued);
        load_ticket();
        cp_async_commit();
        if (fetched < lOrd + kLead) {   // visible to the other threads after the next sync
            if (tid == 0) fetch_one();
            fetched++;
        }
        if (pendingEpilogue >= 0) {
            if (tid == 0) finish_pending();
            pendingEpilogue = -1;
        }

        if (cT == 1) {
            // item constants; prev/cur from own stages 0 and 1
            const TetRec& rec = *reinterpret_cast<const TetRec*>(recRaw + (size_t)(cOrd % kItemRing) * kRecBytes);
            it.pairMask = it.absMask = it.colMask = 0;
#pragma unroll
            for (int f = 0; f < 4; f++) {
                const int bc = rec.bc[f];
                const bool pr = is_pair(bc);
                it.pairMask |= (pr ? 1u : 0u) << f;
                it.absMask |= (bc == VT_PBC_ABSORBING ? 1u : 0u) << f;
                it.colMask |= (rec.wallSlot[f] >= 0 ? 1u : 0u) << f;
                it.hc[f] = pr ? 0.5 * rec.coef[f] : rec.coef[f];
                const double pre = (UPWIND && pr) ? rec.coef[f] : 1.0;
                it.cz[f] = pre * rec.nrm[f][2];
#pragma unroll
                for (int kk = 0; kk < KPT; kk++) {
                    const double v1 = __dadd_rn(p.vmin[1], __dmul_rn((double)colI1[kk], p.step[1]));
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const double v0 = __dadd_rn(p.vmin[0], __dmul_rn((double)(colI0[kk] + u), p.step[0]));
                        it.cxy[kk][f][u] = pre * (rec.nrm[f][0] * v0 + rec.nrm[f][1] * v1);
                    }
                }
            }
            const double* Es = fieldS + (cOrd % kItemRing) * 4;
#pragma unroll
            for (int q = 0; q < 3; q++) it.g[q] = (p.qm * (Es[q] + p.ext[q])) * p.inv2h[q];
            it.nrow = p.fn + (size_t)cur.tet * p.N;
#pragma unroll
            for (int q = 0; q < 4; q++)
                it.push[q] = rec.pushPeer[q] >= 0 ? p.peerFn[rec.pushPeer[q]] + (size_t)rec.pushRow[q] * p.N : nullptr;
            generic = it.pairMask != 0xFu || it.push[0] != nullptr;
            const double* s0 = ownRing + (size_t)cPrevSlot * PE;
            const double* s1 = ownRing + (size_t)cSlot * PE;
#pragma unroll
            for (int kk = 0; kk < KPT; kk++) {
                prv[kk] = *reinterpret_cast<const double2*>(s0 + colE[kk]);
                cr[kk] = *reinterpret_cast<const double2*>(s1 + colE[kk]);
            }
            accDens = 0.0;
#pragma unroll
            for (int f = 0; f < 4; f++) accWall[f] = 0.0;
        } else if (cT >= 2) {
            const double* sc = ownRing + (size_t)cPrevSlot * PE;       // plane j: i0/i1 neighbours
            const double* sn = ownRing + (size_t)cSlot * PE;           // plane j+1
            const double* sb = nbrRing + (size_t)cNbrLast * 4 * PE;    // neighbour planes j
            const int plane = cur.pl0 + cT - 2;
            if (generic)
                compute_plane<KPT, UPWIND, true>(p, it, sc, sn, sb, PE, plane, colE, offU, offD, offL, offR, colOn, prv, cr, accDens, accWall);
            else
                compute_plane<KPT, UPWIND, false>(p, it, sc, sn, sb, PE, plane, colE, offU, offD, offL, offR, colOn, prv, cr, accDens, accWall);
        }
        // advance the consumer
        if (cT >= 1 && cT <= cur.npl) {   // this ticket carried a neighbour stage
            cNbrLast = cNbrNext;
            cNbrNext = (cNbrNext + 1 == R) ? 0 : cNbrNext + 1;
        }
        cPrevSlot = cSlot;
        cSlot = (cSlot + 1 == R) ? 0 : cSlot + 1;
        cT++;
        if (cT > cur.npl + 1) {
            // item finished: leave the partial sums for thread 0 (next iteration, after the sync)
            double* rd = red + (size_t)(cOrd & 1) * 16 * 5;
            const int warp = tid >> 5, lane = tid & 31;
            const double sd = warp_sum(accDens);
            if (lane == 0) rd[warp * 5] = sd;
            if (it.colMask) {
#pragma unroll
                for (int f = 0; f < 4; f++) {
                    const double wv = warp_sum(accWall[f]);
                    if (lane == 0) rd[warp * 5 + 1 + f] = wv;
                }
            }
            pendingEpilogue = cOrd;
            pendingItem = cur;
            cOrd++;
            cT = 0;
            cur = itemS[cOrd % kItemRing];
            if (cur.npl == 0) break;
        }
    }
    __syncthreads();
    if (pendingEpilogue >= 0 && tid == 0) finish_pending();
    cp_async_wait<0>();
}

}  // namespace

// Returns false when the velocity grid does not fit this layout (caller falls back).
bool launch_full_step_async(vt_ctx* ctx, Species& sp, StepParams& p, bool upwind, cudaEvent_t e0, cudaEvent_t e1)
{
    const int n0 = sp.n[0], n1 = sp.n[1];
    if (n0 % 2) return false;
    const int PE = n0 * n1;
    if ((PE * 8) % 16) return false;
    const int PV = PE / 2;
    int nthr, kpt;
    const bool wide = std::getenv("VT_ASYNC_WIDE") != nullptr;   // experiment: 1 column per thread, up to 512 threads
    if (PV <= 128 || (wide && PV <= 512)) {
        nthr = std::max(128, ((PV + 31) / 32) * 32);
        kpt = 1;
    } else if (PV <= 512) {
        kpt = 2;
        nthr = (((PV + 1) / 2 + 31) / 32) * 32;
    } else if (PV <= 1024) {
        kpt = 4;
        nthr = (((PV + 3) / 4 + 31) / 32) * 32;
    } else {
        return false;
    }
    if (nthr > (kpt == 1 ? 512 : 256)) return false;
    const size_t PB = (size_t)PE * 8;
    const size_t fixed = (size_t)kItemRing * kRecBytes + kItemRing * 4 * 8 + kItemRing * sizeof(ItemDesc) + 2 * 16 * 5 * 8 + 256;
    const size_t maxSmem = 227 * 1024;
    int D = 3;
    while (D >= 2 && (size_t)(D + 2) * 5 * PB + fixed > maxSmem) D--;
    if (D < 2) return false;
    const size_t smem = (size_t)(D + 2) * 5 * PB + fixed;

    if (!ctx->workCounter) VT_CUDA(cudaMalloc(&ctx->workCounter, 2 * sizeof(unsigned long long)));
    AsyncParams P;
    P.s = p;
    P.planeElems = PE;
    P.PV = PV;
    P.counter = ctx->workCounter;
    const long long total = (long long)ctx->nOwned * p.nChunks;
    const int grid = (int)std::min<long long>(total, ctx->prop.multiProcessorCount);

    auto launch = [&](auto kern) {
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VT_CUDA(cudaMemsetAsync(ctx->workCounter, 0, sizeof(unsigned long long), ctx->stream));
        VT_CUDA(cudaEventRecord(e0, ctx->stream));
        kern<<<grid, nthr, smem, ctx->stream>>>(P);
        VT_CUDA(cudaEventRecord(e1, ctx->stream));
    };
#define VT_ASYNC_CASE(K, DD)                                                                        \
    if (kpt == K && D == DD) {                                                                      \
        upwind ? launch(k_full_step_async<K, DD, true>) : launch(k_full_step_async<K, DD, false>);  \
        VT_CUDA(cudaGetLastError());                                                                \
        return true;                                                                                \
    }
    VT_ASYNC_CASE(1, 3)
    VT_ASYNC_CASE(1, 2)
    VT_ASYNC_CASE(2, 3)
    VT_ASYNC_CASE(2, 2)
    VT_ASYNC_CASE(4, 3)
    VT_ASYNC_CASE(4, 2)
#undef VT_ASYNC_CASE
    return false;
}

}  // namespace vt
