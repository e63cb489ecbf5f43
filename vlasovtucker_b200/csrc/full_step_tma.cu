// K1 (bulk-copy pipeline) — the fused full-format kinetic update as a persistent, warp-specialised
// kernel for sm_100a.  Same arithmetic and same results as full_step.cu (which stays as the
// general path for velocity grids this layout does not cover); see that file for the reference
// citations of each term.
//
// Why: the register-staged kernel is latency bound — 128 registers per thread cap it at 16 warps
// per SM and every byte in flight is held by a register (profiles/r1_k_full_step_summary.json:
// 73 % of stall samples are long-scoreboard).  Here one producer warp per CTA issues the loads as
// bulk async copies (cp.async.bulk, the 1-D TMA path; SASS UBLKCP) of whole velocity planes
// (n0*n1 doubles, 8 KiB at 32x32) into shared-memory rings; completion is tracked by mbarriers
// ("full"), consumer warps hand slots back through a second set ("empty"), and no register is
// tied up by data in flight: ~150 KiB per SM are permanently on their way while eight consumer
// warps compute from shared memory.  There is no CTA-wide barrier in the steady state.
//
// Work: one CTA per SM, persistent.  Work items (tet, chunk of i2-planes) are handed out through a
// global atomic queue in brick-major, chunk, tet order, so that at any moment all SMs work on the
// same few hundred tets and the same i2 window — the neighbour rows another CTA needs are then in
// L2 (the static round-robin of the first version of this file measured 3.1x the algorithmic DRAM
// traffic).  A consumer thread owns KPT fixed (i0-pair, i1) columns of the plane and marches along
// i2 keeping the own-row values prev/cur/next in registers; the i0 and i1 stencil neighbours and
// the four neighbour tets' values come from shared memory.
//
// Instances: 16 consumer warps x KPT columns per thread (planes of any even n0 up to 2048 double2), or
// eight warps x two i1-adjacent columns when the plane is exactly 512 double2, with the 32x32 shape of
// the bench grid known at compile time.  Tets that need boundary-condition or halo-push code are
// launched first over their own list, the interior ones after them with those branches compiled
// out; all instances use explicit roundings so that they produce the same bits.
//
// Stage stream of one item with npl planes starting at pl0 (periodic in i2, solver.cpp:380-389):
//   own stage s = 0 .. npl+1  : plane pl0-1+s of the tet's own row          (ring of OD planes)
//   nbr stage j = 0 .. npl-1  : plane pl0+j of the (up to) four neighbours  (ring of S x 4 planes)
// The producer issues own(0), own(1), nbr(0), own(2), nbr(1), ...; computing plane j needs
// own(j+1) [centre, i0/i1 neighbours], own(j+2) [i2+1], nbr(j), and prev/cur from registers.
#include "vt_internal.h"

#include <algorithm>
#include <cmath>

namespace vt {

namespace {

constexpr int kItemRing = 4;
constexpr int kMaxRing = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// 16-byte async copy global -> shared through the LSU (SASS LDGSTS.BYPASS), L2 only
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
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

struct ItemHdr {
    int tet, chunk, pl0, npl;   // tet < 0: no more work
    double E[3];                // field of the tet, fetched by the producer one item ahead
    double pad;
};

struct BulkParams {
    StepParams s;
    int OD, S;        // own ring depth (planes), neighbour ring depth (4-plane sets)
    int planeElems;   // n0*n1
    int PV;           // double2 per plane
    unsigned long long* queue;
    long long total;
    const int* tetList;   // nullptr: all owned tets in device order
    int nTets;
    int orderMode;    // 0: brick, chunk, tet (tet fastest); 1: tet, chunk (chunk fastest)
    // single-precision copies for the neighbour-load predicate (consumer-side loads only)
    float vmin0f, vmin1f, vmin2f, step0f, step1f, step2f, guard;
};

// Work item w of a launch over nTets tets (all owned tets, or the sub-list tetList of them).
// orderMode 0: brick-major, then chunk, then tet — all SMs sweep one velocity window of one brick;
// orderMode 1: tet-major, chunk fastest — the chunks of one tet are in flight on neighbouring CTAs at
// the same time, so the rows in flight are total/nChunks tets and the i2 halo planes two chunks share
// are fetched from DRAM once.
__device__ __forceinline__ bool decode_item(const BulkParams& P, long long w64, ItemHdr& it)
{
    if (w64 >= P.total) return false;
    const StepParams& p = P.s;
    const unsigned w = (unsigned)w64;   // total < 2^31 (checked by the launcher)
    int pos;
    if (P.orderMode == 1) {
        pos = (int)(w / (unsigned)p.nChunks);
        it.chunk = (int)(w - (unsigned)pos * (unsigned)p.nChunks);
    } else {
        const unsigned perBrick = (unsigned)p.brickTets * (unsigned)p.nChunks;
        const unsigned brick = w / perBrick;
        const int base = (int)(brick * (unsigned)p.brickTets);
        const int nb = min(p.brickTets, P.nTets - base);
        const unsigned r = w - brick * perBrick;
        it.chunk = (int)(r / (unsigned)nb);
        pos = base + (int)(r - (unsigned)it.chunk * (unsigned)nb);
    }
    it.tet = P.tetList ? P.tetList[pos] : pos;
    it.pl0 = it.chunk * p.chunkPlanes;
    it.npl = min(p.chunkPlanes, p.n2 - it.pl0);
    return true;
}

struct Smem {
    double* ownRing;     // [OD][PE]
    double* nbrRing;     // [S][4][PE]
    TetRec* rec;         // [kItemRing]
    ItemHdr* hdr;        // [kItemRing]
    uint64_t *ownFull, *ownEmpty, *nbrFull, *nbrEmpty, *itemFull, *itemEmpty;
};

// position in a ring of D slots; `phase` flips at every wrap (mbarrier parity of the slot's use)
struct Cursor {
    int slot = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int D)
    {
        if (++slot == D) {
            slot = 0;
            phase ^= 1u;
        }
    }
};

// ---- producer: one thread; claims items from the queue and keeps the rings full.
// NBRCP: the consumers fetch the neighbour values themselves (see NbrFetch), the producer streams
// the own planes only and never has to wait for a tet record.
template <bool NBRCP>
__device__ void producer_loop(const BulkParams& P, const Smem& sm)
{
    const StepParams& p = P.s;
    const int PE = P.planeElems;
    const uint32_t PB = (uint32_t)PE * 8u;
    Cursor cItem, cPost, cOwn, cNbr;

    // An item slot completes on two arrivals: the record's bulk copy (posted here) and the field
    // values, which the producer loads into registers here and stores one iteration later
    // (finish_item) so that their latency is never waited for.
    double ePend[3] = {0.0, 0.0, 0.0};
    int pendSlot = -1;
    auto post_item = [&](long long w, ItemHdr& h) {   // header + tet record of the item into slot cPost
        mbar_wait(sm.itemEmpty + cPost.slot, cPost.phase ^ 1u);
        if (decode_item(P, w, h)) {
            sm.hdr[cPost.slot].tet = h.tet;
            sm.hdr[cPost.slot].chunk = h.chunk;
            sm.hdr[cPost.slot].pl0 = h.pl0;
            sm.hdr[cPost.slot].npl = h.npl;
#pragma unroll
            for (int q = 0; q < 3; q++) ePend[q] = __ldg(p.E + 3 * (size_t)h.tet + q);
            mbar_expect_tx(sm.itemFull + cPost.slot, (uint32_t)sizeof(TetRec));
            bulk_g2s(sm.rec + cPost.slot, p.rec + h.tet, (uint32_t)sizeof(TetRec), sm.itemFull + cPost.slot);
        } else {
            h.tet = -1;
            sm.hdr[cPost.slot].tet = -1;
            mbar_arrive(sm.itemFull + cPost.slot);
        }
        pendSlot = cPost.slot;
        cPost.advance(kItemRing);
    };
    auto finish_item = [&]() {
#pragma unroll
        for (int q = 0; q < 3; q++) sm.hdr[pendSlot].E[q] = ePend[q];
        mbar_arrive(sm.itemFull + pendSlot);
    };

    ItemHdr h, hNext;
    long long wNext = (long long)atomicAdd(P.queue, 1ULL);
    post_item(wNext, h);
    wNext = (long long)atomicAdd(P.queue, 1ULL);
    while (true) {
        finish_item();                                      // item k
        post_item(wNext, hNext);                            // item k+1: its record lands while item k streams
        wNext = (long long)atomicAdd(P.queue, 1ULL);        // item k+2: the ticket is used one iteration later
        if (h.tet < 0) break;
        const double* rows[4] = {nullptr, nullptr, nullptr, nullptr};
        int cnt = 0;
        if (!NBRCP) {
            mbar_wait(sm.itemFull + cItem.slot, cItem.phase);
            const TetRec& r = sm.rec[cItem.slot];
#pragma unroll
            for (int f = 0; f < 4; f++) {
                if (is_pair(r.bc[f])) {
                    const int n = r.nbr[f];
                    rows[f] = n >= 0 ? p.f + (size_t)n * p.N : p.src + (size_t)(-2 - n) * p.N;
                    cnt++;
                }
            }
        }
        const double* own = p.f + (size_t)h.tet * p.N;
        for (int s = 0; s <= h.npl + 1; s++) {
            int ip = h.pl0 - 1 + s;
            if (ip < 0) ip += p.n2;
            if (ip >= p.n2) ip -= p.n2;
            mbar_wait(sm.ownEmpty + cOwn.slot, cOwn.phase ^ 1u);
            mbar_expect_tx(sm.ownFull + cOwn.slot, PB);
            bulk_g2s(sm.ownRing + (size_t)cOwn.slot * PE, own + (size_t)ip * PE, PB, sm.ownFull + cOwn.slot);
            cOwn.advance(P.OD);
            if (!NBRCP && s >= 1 && s <= h.npl) {
                const size_t plane = (size_t)(h.pl0 + s - 1) * PE;
                uint64_t* bar = sm.nbrFull + cNbr.slot;
                mbar_wait(sm.nbrEmpty + cNbr.slot, cNbr.phase ^ 1u);
                if (cnt > 0) mbar_expect_tx(bar, PB * cnt);
                else mbar_arrive(bar);
#pragma unroll
                for (int f = 0; f < 4; f++)
                    if (rows[f]) bulk_g2s(sm.nbrRing + ((size_t)cNbr.slot * 4 + f) * PE, rows[f] + plane, PB, bar);
                cNbr.advance(P.S);
            }
        }
        cItem.advance(kItemRing);
        h = hNext;
    }
}

// ---- consumers: all planes of one work item.  GENERIC: per-face boundary conditions and halo
// push (branches are uniform across the CTA); otherwise four pair faces and no push.
// Per-thread constants of one column (fixed for the whole kernel): byte offset of its double2 inside
// a plane and of its i1-1 / i1+1 / i0-1 / i0+2 neighbours (periodic wrap), all relative to the
// plane start.
struct Column {
    uint32_t evB, dUmB, dUpB, dFlB, dFrB;
    int i0, i1;
    bool on;
};

__device__ __forceinline__ double2 lds128(uint32_t a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void mbar_wait32(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive32(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Consumer-side view of the rings: 32-bit shared addresses, ring depths are powers of two so a
// stage counter gives slot = c & mask and parity = (c >> shift) & 1.
struct ConsRings {
    uint32_t own, nbr;                       // ring data
    uint32_t ownFull, ownEmpty, nbrFull, nbrEmpty;
    uint32_t PB;                             // bytes per plane
    uint32_t odMask, odShift, sMask, sShift;
};

// ---- consumer-side neighbour loads (NBD > 0).
// With the upwind-select flux a tet reads its neighbour across face f only where v.n_f <= 0 (inflow);
// where v.n_f > 0 the flux uses the tet's own value.  On average that is half of every neighbour
// row, and a bulk copy of whole planes moves the other half through L2 and the crossbar for
// nothing — the round-1 kernel ran at 5,940 B/cycle of L2->SM traffic, the chip's limit.  Here every
// consumer thread requests exactly the neighbour double2s its own columns will use, NBD-1 planes
// ahead, as predicated 16-byte cp.async (LDGSTS, L2 only) into its private positions of the
// neighbour ring; it is the only reader of those positions, so completion is the thread's own
// cp.async group count and the ring needs no barrier at all.  The predicate is evaluated in single
// precision with a guard band (BulkParams::guard) wide enough that every value the FP64 select can
// pick has been requested; values in the guard band are loaded and not used.
template <int KPT>
struct NbrFetch {
    int itemSlot = 0;
    uint32_t itemPhase = 0;   // next item slot to open
    int left = 0;             // planes of the open item not requested yet
    bool done = false;        // the end-of-queue item has been seen
    uint32_t cnt = 0;         // planes requested so far (ring slot = cnt & (NBD-1))
    float v2 = 0.f;           // velocity of the next plane to request
    const char* gp[4];        // neighbour row f at that plane, at the thread's first column
    float cmin[KPT][4];       // min over the double2 of (n_x v0 + n_y v1), minus the guard; +inf: never load
    float cz[4];              // n_z
};

template <int KPT, int NBD>
__device__ __forceinline__ void nbr_fetch_plane(const BulkParams& P, const Smem& sm, const ConsRings& R, const uint32_t PB,
                                                NbrFetch<KPT>& F, const uint32_t (&oEv)[KPT], const float (&vf)[KPT][3])
{
    const StepParams& p = P.s;
    if (F.left == 0 && !F.done) {
        mbar_wait(sm.itemFull + F.itemSlot, F.itemPhase);
        const int tet = sm.hdr[F.itemSlot].tet;
        if (tet < 0) {
            F.done = true;
        } else {
            const TetRec& r = sm.rec[F.itemSlot];
            const int pl0 = sm.hdr[F.itemSlot].pl0;
            F.left = sm.hdr[F.itemSlot].npl;
            F.v2 = fmaf((float)pl0, P.step2f, P.vmin2f);
#pragma unroll
            for (int f = 0; f < 4; f++) {
                const int n = r.nbr[f];
                const bool have = is_pair(r.bc[f]) && n != -1;
                const double* row = n >= 0 ? p.f + (size_t)n * p.N : p.src + (size_t)(n <= -2 ? -2 - n : 0) * p.N;
                F.gp[f] = reinterpret_cast<const char*>(row + (size_t)pl0 * P.planeElems) + oEv[0];
                const float nx = (float)r.nrm[f][0], ny = (float)r.nrm[f][1];
                F.cz[f] = (float)r.nrm[f][2];
#pragma unroll
                for (int kk = 0; kk < KPT; kk++)
                    F.cmin[kk][f] = have ? fminf(nx * vf[kk][0], nx * vf[kk][1]) + ny * vf[kk][2] - P.guard : INFINITY;
            }
        }
        if (++F.itemSlot == kItemRing) {
            F.itemSlot = 0;
            F.itemPhase ^= 1u;
        }
    }
    if (F.left > 0) {
        const uint32_t dst0 = R.nbr + (F.cnt & (uint32_t)(NBD - 1)) * 4u * PB + oEv[0];
#pragma unroll
        for (int f = 0; f < 4; f++) {
#pragma unroll
            for (int kk = 0; kk < KPT; kk++) {
                const float t = fmaf(F.cz[f], F.v2, F.cmin[kk][f]);
                const uint32_t d = oEv[kk] - oEv[0];
                if (t <= 0.f) cp_async16(dst0 + (uint32_t)f * PB + d, F.gp[f] + d);
            }
            F.gp[f] += PB;
        }
        F.v2 += P.step2f;
        F.left--;
    }
    F.cnt++;
    cp_async_commit();
}

template <int KPT, bool UPWIND, bool GENERIC, int NCW, int SN, int NBD>
__device__ __forceinline__ void item_compute(const BulkParams& P, const Smem& sm, const ConsRings& R, const ItemHdr& cur,
                                             const TetRec& rec, uint32_t& cOwn, uint32_t& cNbr, const Column (&col)[KPT],
                                             const uint32_t (&oEv)[KPT], NbrFetch<KPT>& F, const float (&vf)[KPT][3])
{
    const StepParams& p = P.s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr bool PAIRED = KPT == 2 && NCW == 8;   // see the column mapping in k_full_step_bulk

    double cxy[KPT][4][2], hc[4], cz[4];
    bool pairF[4], absF[4], colF[4];
#pragma unroll
    for (int f = 0; f < 4; f++) {
        const int bc = GENERIC ? rec.bc[f] : VT_PBC_NONBOUNDARY;
        pairF[f] = is_pair(bc);
        absF[f] = GENERIC && bc == VT_PBC_ABSORBING;
        colF[f] = GENERIC && rec.wallSlot[f] >= 0;
        hc[f] = pairF[f] ? 0.5 * rec.coef[f] : rec.coef[f];
        const double pre = (UPWIND && pairF[f]) ? rec.coef[f] : 1.0;
        cz[f] = pre * rec.nrm[f][2];
#pragma unroll
        for (int kk = 0; kk < KPT; kk++) {
            const double v1 = __dadd_rn(p.vmin[1], __dmul_rn((double)col[kk].i1, p.step[1]));
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const double v0 = __dadd_rn(p.vmin[0], __dmul_rn((double)(col[kk].i0 + u), p.step[0]));
                // explicit roundings here and below: every template instance of this function has to
                // produce the same bits (a partitioned run mixes them, and must equal the single-GPU run)
                cxy[kk][f][u] = __dmul_rn(pre, __fma_rn(rec.nrm[f][0], v0, __dmul_rn(rec.nrm[f][1], v1)));
            }
        }
    }
    double g[3];
#pragma unroll
    for (int q = 0; q < 3; q++) g[q] = (p.qm * (cur.E[q] + p.ext[q])) * p.inv2h[q];
    // running output pointers (advance one plane per iteration)
    char* outp[KPT];
#pragma unroll
    for (int kk = 0; kk < KPT; kk++)
        outp[kk] = reinterpret_cast<char*>(p.fn + (size_t)cur.tet * p.N + (size_t)cur.pl0 * P.planeElems) + oEv[kk];
    long long pushOff[4] = {0, 0, 0, 0};   // byte distance from this tet's row to its ghost copies
    bool pushOn[4] = {false, false, false, false};
    if (GENERIC) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (rec.pushPeer[q] >= 0) {
                pushOn[q] = true;
                pushOff[q] = reinterpret_cast<char*>(p.peerFn[rec.pushPeer[q]] + (size_t)rec.pushRow[q] * p.N) -
                             reinterpret_cast<char*>(p.fn + (size_t)cur.tet * p.N);
            }
    }
    // Halo push, upwind-select arithmetic: the peer that holds this tet as a ghost row reads it only where ITS
    // v.n <= 0, i.e. where v.n_f >= 0 for the face f shared with that peer (the consumer-side inflow-half loads
    // above, or the select `vn > 0 ? f : fa`).  So only that half of the row goes over NVLink — the union over
    // the faces whose neighbour is a ghost row, with the guard band of the receiver's single-precision predicate.
    // The expression-shape arithmetic reads the whole neighbour row: everything is pushed then.
    bool ghostF[4] = {false, false, false, false};
    if (GENERIC && UPWIND) {
#pragma unroll
        for (int f = 0; f < 4; f++) ghostF[f] = ((rec.ghostFaces >> f) & 1) != 0;
    }
    const double pushGuard = -(double)P.guard;
    double accDens = 0.0;
    double accWall[4] = {0.0, 0.0, 0.0, 0.0};

    // own stages 0 and 1: prev (registers only) and cur (registers + i0/i1 neighbours from smem)
    double2 prv[KPT], cr[KPT];
    {
        const uint32_t slot = cOwn & R.odMask;
        mbar_wait32(R.ownFull + slot * 8, (cOwn >> R.odShift) & 1u);
        const uint32_t a = R.own + slot * R.PB;
#pragma unroll
        for (int kk = 0; kk < KPT; kk++) prv[kk] = lds128(a + oEv[kk]);
        __syncwarp();
        if (lane == 0) mbar_arrive32(R.ownEmpty + slot * 8);
        cOwn++;
    }
    uint32_t scSlot = cOwn & R.odMask;
    mbar_wait32(R.ownFull + scSlot * 8, (cOwn >> R.odShift) & 1u);
    uint32_t scA = R.own + scSlot * R.PB;
    cOwn++;
#pragma unroll
    for (int kk = 0; kk < KPT; kk++) cr[kk] = lds128(scA + oEv[kk]);

    // Per-thread byte offsets of the stencil operands.  With a compile-time plane shape (SN x SN) and
    // paired columns the second column is the first plus one row, so most addresses become
    // register + immediate.
    constexpr bool SPEC = SN > 0 && PAIRED;
    const uint32_t PB = SPEC ? (uint32_t)(SN * SN * 8) : R.PB;
    uint32_t oUm[KPT], oUp[KPT], oFl[KPT], oFr[KPT];
#pragma unroll
    for (int kk = 0; kk < KPT; kk++) {
        oUm[kk] = col[kk].dUmB;
        oUp[kk] = col[kk].dUpB;
        oFl[kk] = (SPEC && kk == 1) ? col[0].dFlB + SN * 8 : col[kk].dFlB;
        oFr[kk] = (SPEC && kk == 1) ? col[0].dFrB + SN * 8 : col[kk].dFrB;
    }

    double2 nxt[KPT];
    int j = 0;
    // One plane: pv/cv hold planes j-1 and j of the thread's columns, nv receives plane j+1.  The
    // plane loop below calls it with the three register sets rotating, so no values are moved.
    auto plane = [&](const double2 (&pv)[KPT], const double2 (&cv)[KPT], double2 (&nv)[KPT]) {
        // request the neighbour values of plane j+NBD-1 (of this item or the next ones) before waiting
        if (NBD > 0) nbr_fetch_plane<KPT, (NBD > 0 ? NBD : 1)>(P, sm, R, PB, F, oEv, vf);
        const uint32_t snSlot = cOwn & R.odMask;
        mbar_wait32(R.ownFull + snSlot * 8, (cOwn >> R.odShift) & 1u);
        const uint32_t snA = R.own + snSlot * PB;       // plane j+1
        cOwn++;
        const uint32_t nbSlot = NBD > 0 ? (cNbr & (uint32_t)(NBD - 1)) : (cNbr & R.sMask);
        if (NBD > 0) cp_async_wait<(NBD > 0 ? NBD - 1 : 0)>();
        else mbar_wait32(R.nbrFull + nbSlot * 8, (cNbr >> R.sShift) & 1u);
        const uint32_t sbA = R.nbr + nbSlot * 4 * PB;
        cNbr++;

        const double v2 = __dadd_rn(p.vmin[2], __dmul_rn((double)(cur.pl0 + j), p.step[2]));
        double tzf[4];
#pragma unroll
        for (int f = 0; f < 4; f++) tzf[f] = __dmul_rn(cz[f], v2);
#pragma unroll
        for (int kk = 0; kk < KPT; kk++) {
            if (!PAIRED && !col[kk].on) continue;   // the paired layout covers the plane exactly
            const double2 nx = lds128(snA + oEv[kk]);
            nv[kk] = nx;
            // PAIRED: the thread's two columns are neighbours in i1, so each one's i1 neighbour on
            // that side is the other one's centre value, already in registers
            const double2 um = (PAIRED && kk == 1) ? cv[0] : lds128(scA + oUm[kk]);
            const double2 up = (PAIRED && kk == 0) ? cv[KPT - 1] : lds128(scA + oUp[kk]);
            const double fl = lds64(scA + oFl[kk]);
            const double fr = lds64(scA + oFr[kk]);
            double2 fa[4];
            {
                uint32_t a = sbA + oEv[kk];
#pragma unroll
                for (int f = 0; f < 4; f++) {
                    fa[f] = (!GENERIC || pairF[f]) ? lds128(a) : make_double2(0.0, 0.0);
                    a += PB;
                }
            }
            const double fcv[2] = {cv[kk].x, cv[kk].y};
            const double xm[2] = {fl, cv[kk].x};
            const double xp[2] = {cv[kk].y, fr};
            const double y1m[2] = {um.x, um.y}, y1p[2] = {up.x, up.y};
            const double z2m[2] = {pv[kk].x, pv[kk].y}, z2p[2] = {nx.x, nx.y};
            double out[2];
            bool pushEl = !UPWIND;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const double fv = fcv[u];
                double rhs = 0.0;
#pragma unroll
                for (int f = 0; f < 4; f++) {
                    const double vn = __dadd_rn(cxy[kk][f][u], tzf[f]);
                    if (GENERIC && UPWIND && ghostF[f] && vn >= pushGuard) pushEl = true;
                    const double fau = u == 0 ? fa[f].x : fa[f].y;
                    if (!GENERIC || pairF[f]) {
                        if (UPWIND) {
                            // with NBD > 0 fau is defined only where vn <= 0 (plus the guard band)
                            rhs = fma(-vn, vn > 0.0 ? fv : fau, rhs);
                        } else {
                            const double s = __dadd_rn(fau, fv), d = __dsub_rn(fau, fv);
                            rhs = __fma_rn(-hc[f], __fma_rn(vn, s, -__dmul_rn(fabs(vn), d)), rhs);
                        }
                    } else if (absF[f]) {
                        const double flux = 0.5 * (vn * fv + fabs(vn) * fv);
                        if (colF[f]) accWall[f] += flux;
                        rhs = fma(-hc[f], flux, rhs);
                    } else {
                        rhs = fma(-hc[f], vn * fv, rhs);
                    }
                }
                rhs = __fma_rn(-g[0], __dsub_rn(xp[u], xm[u]), rhs);
                rhs = __fma_rn(-g[1], __dsub_rn(y1p[u], y1m[u]), rhs);
                rhs = __fma_rn(-g[2], __dsub_rn(z2p[u], z2m[u]), rhs);
                out[u] = fma(p.dt, rhs, fv);
                accDens = __dadd_rn(accDens, out[u]);
            }
            const double2 o = make_double2(out[0], out[1]);
            // streaming store (evict-first): the new state is not read again in this launch and should
            // not push the step-n rows, which four neighbours still want, out of L2
            __stcs(reinterpret_cast<double2*>(outp[kk]), o);
            if (GENERIC) {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (pushOn[q] && pushEl) *reinterpret_cast<double2*>(outp[kk] + pushOff[q]) = o;
            }
            outp[kk] += PB;
        }
        __syncwarp();   // every lane is done with own stage sc and neighbour slot nbSlot
        if (lane == 0) {
            mbar_arrive32(R.ownEmpty + scSlot * 8);
            if (NBD == 0) mbar_arrive32(R.nbrEmpty + nbSlot * 8);
        }
        scA = snA;
        scSlot = snSlot;
        j++;
    };
    while (true) {
        plane(prv, cr, nxt);
        if (j == cur.npl) break;
        plane(cr, nxt, prv);
        if (j == cur.npl) break;
        plane(nxt, prv, cr);
        if (j == cur.npl) break;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive32(R.ownEmpty + scSlot * 8);   // the last own stage (plane pl0+npl)

    // ---- item epilogue: sum_v f' (Density) per warp, reduced in fixed order by k_density_reduce;
    // absorbed flux (wall charge)
    const double sd = warp_sum(accDens);
    if (lane == 0) p.densPartial[((size_t)cur.tet * p.nChunks + cur.chunk) * NCW + warp] = sd;
    if (GENERIC) {
#pragma unroll
        for (int f = 0; f < 4; f++) {
            if (!colF[f]) continue;
            const double wv = warp_sum(accWall[f]);
            // charge * (timeStep * area * flux.Sum() * cellVolume), solver.cpp:173-177
            if (lane == 0) atomicAdd(p.wall + rec.wallSlot[f], p.wallScale * rec.area[f] * wv);
        }
    }
}

// NBD = 0: the producer bulk-copies whole neighbour planes; NBD = 2 or 4: the consumers request the
// neighbour values they use themselves, NBD-1 planes ahead (upwind-select arithmetic only).
// Register budget: the CTA is NCW consumer warps plus one auxiliary warp GROUP (four warps, of which
// one lane runs the producer).  A separate producer WARP would make 9 (17) warps, i.e. three (five)
// on one SM sub-partition, and cap every thread at 168 (96) registers; with whole warp groups the
// auxiliary group hands its registers back (setmaxnreg.dec) and the consumer groups take them
// (setmaxnreg.inc): 232 registers per consumer thread with eight consumer warps, 104 with sixteen (RegSplit).
constexpr int kAuxThreads = 128;
// registers after the reallocation; 4 * aux + NCW * consumer <= (NCW + 4) * launch count (168 / 96)
template <int NCW, int NBD>
struct RegSplit {
    static constexpr int aux = (NCW == 8 && NBD > 0) ? 40 : 56;   // the producer that also copies neighbour planes needs more
    static constexpr int consumer = NCW == 8 ? (NBD > 0 ? 232 : 224) : 104;
};

template <int KPT, bool UPWIND, int NCW, bool ALLFAST, int SN, int NBD>
__global__ void __launch_bounds__(NCW * 32 + kAuxThreads, 1) k_full_step_bulk(const BulkParams P)
{
    static_assert(NBD == 0 || UPWIND, "consumer-side neighbour loads rely on the upwind select");
    const StepParams& p = P.s;
    extern __shared__ __align__(128) unsigned char smraw[];
    const int PE = P.planeElems;
    Smem sm;
    sm.ownRing = reinterpret_cast<double*>(smraw);
    sm.nbrRing = sm.ownRing + (size_t)P.OD * PE;
    sm.rec = reinterpret_cast<TetRec*>(sm.nbrRing + (size_t)P.S * 4 * PE);
    sm.hdr = reinterpret_cast<ItemHdr*>(sm.rec + kItemRing);
    sm.ownFull = reinterpret_cast<uint64_t*>(sm.hdr + kItemRing);
    sm.ownEmpty = sm.ownFull + kMaxRing;
    sm.nbrFull = sm.ownEmpty + kMaxRing;
    sm.nbrEmpty = sm.nbrFull + kMaxRing;
    sm.itemFull = sm.nbrEmpty + kMaxRing;
    sm.itemEmpty = sm.itemFull + kItemRing;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < P.OD; i++) {
            mbar_init(sm.ownFull + i, 1);
            mbar_init(sm.ownEmpty + i, NCW);
        }
        for (int i = 0; i < P.S; i++) {
            mbar_init(sm.nbrFull + i, 1);
            mbar_init(sm.nbrEmpty + i, NCW);
        }
        for (int i = 0; i < kItemRing; i++) {
            mbar_init(sm.itemFull + i, 2);
            mbar_init(sm.itemEmpty + i, NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= NCW * 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RegSplit<NCW, NBD>::aux));
        if (tid == NCW * 32) producer_loop<(NBD > 0)>(P, sm);
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RegSplit<NCW, NBD>::consumer));

    // fixed columns of this consumer thread
    constexpr bool SPEC = SN > 0 && KPT == 2 && NCW == 8;
    Column col[KPT];
    uint32_t oEv[KPT];
    float vf[KPT][3];   // v0(i0), v0(i0+1), v1(i1) of the column in single precision (neighbour-load predicate)
#pragma unroll
    for (int kk = 0; kk < KPT; kk++) {
        // eight warps x two columns (launched only when the plane is exactly 2 x 256 double2 and n1 is
        // even): the two columns of a thread are the rows 2m and 2m+1 of the same i0 pair
        const int v = (KPT == 2 && NCW == 8) ? (tid % p.nvec0) + p.nvec0 * (2 * (tid / p.nvec0) + kk) : tid + kk * NCW * 32;
        col[kk].on = v < P.PV;
        const int cv = col[kk].on ? v : 0;
        const int i0 = (cv % p.nvec0) * 2, i1 = cv / p.nvec0;
        col[kk].i0 = i0;
        col[kk].i1 = i1;
        col[kk].evB = 16u * cv;
        col[kk].dUmB = 8u * (uint32_t)(2 * cv + ((i1 == 0) ? (p.n1 - 1) : -1) * p.n0);
        col[kk].dUpB = 8u * (uint32_t)(2 * cv + ((i1 == p.n1 - 1) ? -(p.n1 - 1) : 1) * p.n0);
        col[kk].dFlB = 8u * (uint32_t)(2 * cv + ((i0 == 0) ? (p.n0 - 1) : -1));
        col[kk].dFrB = 8u * (uint32_t)(2 * cv + 1 + ((i0 + 2 == p.n0) ? -(p.n0 - 1) : 1));
        oEv[kk] = (SPEC && kk == 1) ? col[0].evB + SN * 8 : col[kk].evB;
        // a column beyond the plane (the last columns of a thread when the plane is not a multiple of the thread
        // count) must never request neighbour values: NaN makes the load predicate false whatever the normal is
        vf[kk][0] = col[kk].on ? fmaf((float)i0, P.step0f, P.vmin0f) : __int_as_float(0x7fc00000);
        vf[kk][1] = col[kk].on ? fmaf((float)(i0 + 1), P.step0f, P.vmin0f) : __int_as_float(0x7fc00000);
        vf[kk][2] = col[kk].on ? fmaf((float)i1, P.step1f, P.vmin1f) : __int_as_float(0x7fc00000);
    }
    ConsRings R;
    R.own = smem_u32(sm.ownRing);
    R.nbr = smem_u32(sm.nbrRing);
    R.ownFull = smem_u32(sm.ownFull);
    R.ownEmpty = smem_u32(sm.ownEmpty);
    R.nbrFull = smem_u32(sm.nbrFull);
    R.nbrEmpty = smem_u32(sm.nbrEmpty);
    R.PB = (uint32_t)PE * 8u;
    R.odMask = (uint32_t)P.OD - 1u;
    R.odShift = 31u - (uint32_t)__clz(P.OD);
    R.sMask = (uint32_t)P.S - 1u;
    R.sShift = 31u - (uint32_t)__clz(P.S);
    Cursor cItem;
    uint32_t cOwn = 0, cNbr = 0;
    const int lane = tid & 31;
    NbrFetch<KPT> F;
    if (NBD > 0) {
        // prime the neighbour pipeline: planes 0 .. NBD-2 of the CTA's plane stream
        const uint32_t PBk = SPEC ? (uint32_t)(SN * SN * 8) : R.PB;
#pragma unroll 1
        for (int d = 0; d < NBD - 1; d++) nbr_fetch_plane<KPT, (NBD > 0 ? NBD : 1)>(P, sm, R, PBk, F, oEv, vf);
    }
    while (true) {
        mbar_wait(sm.itemFull + cItem.slot, cItem.phase);
        const ItemHdr cur = sm.hdr[cItem.slot];
        if (cur.tet < 0) break;
        const TetRec& rec = sm.rec[cItem.slot];
        if (ALLFAST) {
            item_compute<KPT, UPWIND, false, NCW, SN, NBD>(P, sm, R, cur, rec, cOwn, cNbr, col, oEv, F, vf);
        } else {
            const bool fast = is_pair(rec.bc[0]) && is_pair(rec.bc[1]) && is_pair(rec.bc[2]) && is_pair(rec.bc[3]) &&
                              rec.pushPeer[0] < 0;
            if (fast) item_compute<KPT, UPWIND, false, NCW, SN, NBD>(P, sm, R, cur, rec, cOwn, cNbr, col, oEv, F, vf);
            else item_compute<KPT, UPWIND, true, NCW, SN, NBD>(P, sm, R, cur, rec, cOwn, cNbr, col, oEv, F, vf);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.itemEmpty + cItem.slot);
        cItem.advance(kItemRing);
    }
    if (NBD > 0) cp_async_wait<0>();
}

template <int KPT, int NCW, int SN, int NBD>
void launch_cfg(vt_ctx* ctx, const BulkParams& P, bool upwind, bool allFast, size_t smem, cudaStream_t stream, int maxCTAs)
{
    if (P.total == 0) return;
    const int grid = (int)std::min<long long>(P.total, maxCTAs);
    auto launch = [&](auto kern) {
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VT_CUDA(cudaMemsetAsync(P.queue, 0, sizeof(unsigned long long), stream));
        kern<<<grid, NCW * 32 + kAuxThreads, smem, stream>>>(P);
        ctx->launches++;
    };
    if constexpr (NBD > 0) {
        // consumer-side neighbour loads exist for the upwind-select arithmetic only
        allFast ? launch(k_full_step_bulk<KPT, true, NCW, true, SN, NBD>) : launch(k_full_step_bulk<KPT, true, NCW, false, SN, NBD>);
    } else {
        if (allFast) upwind ? launch(k_full_step_bulk<KPT, true, NCW, true, SN, 0>) : launch(k_full_step_bulk<KPT, false, NCW, true, SN, 0>);
        else upwind ? launch(k_full_step_bulk<KPT, true, NCW, false, SN, 0>) : launch(k_full_step_bulk<KPT, false, NCW, false, SN, 0>);
    }
}

}  // namespace

// Returns false when the velocity grid does not fit this layout (caller falls back).  On success
// p.densSplit tells the caller how many partial sums per (tet, chunk) were written.
bool launch_full_step_tma(vt_ctx* ctx, Species& sp, StepParams& p, bool upwind, cudaEvent_t e0, cudaEvent_t e1)
{
    const int n0 = sp.n[0], n1 = sp.n[1];
    if (n0 % 2) return false;
    const int PE = n0 * n1;
    if ((PE * 8) % 16) return false;
    if (((size_t)sp.N * 8) % 16) return false;
    const int PV = PE / 2;
    // variant bit 5: eight consumer warps with two columns per thread instead of sixteen with one
    const bool wide = (ctx->variant & 32) != 0 && PV == 512 && n1 % 2 == 0 && 256 % (n0 / 2) == 0;
    // planes above 32x32 (up to 50x50): also eight consumer warps, three to five columns per thread in the generic
    // column mapping — the sixteen-warp instances are capped at 96 registers per thread and spill from three
    // columns on (cuobjdump -res-usage: 96-528 bytes of stack), eight warps get 232 after the register re-split.
    // Opt-in (VT_STEP_WIDE8=1) until it has seen the whole parity suite.
    const bool wide8 = (ctx->variant & 32) != 0 && !wide && PV > 2 * 256 && PV <= 5 * 256 && std::getenv("VT_STEP_WIDE8") != nullptr;
    const int ncw = (wide || wide8) ? 8 : 16;
    const int kpt = (PV + ncw * 32 - 1) / (ncw * 32);
    if (kpt > (wide8 ? 5 : 4)) return false;
    // variant bit 7 switches the consumer-side neighbour loads off (whole neighbour planes by bulk copy)
    const bool nbrSelf = upwind && (wide || wide8) && !(ctx->variant & 128);
    const size_t PB = (size_t)PE * 8;
    const size_t fixed = kItemRing * (sizeof(TetRec) + sizeof(ItemHdr)) + (4 * kMaxRing + 2 * kItemRing) * 8 + 128;
    const size_t maxSmem = 227 * 1024;
    // ring depths are powers of two (the consumers derive slot and parity from a stage counter)
    int S = nbrSelf ? 4 : kMaxRing, OD = kMaxRing;
    while (S > 2 && (size_t)S * 4 * PB + (size_t)4 * PB + fixed > maxSmem) S /= 2;
    while (OD > 4 && (size_t)OD * PB + (size_t)S * 4 * PB + fixed > maxSmem) OD /= 2;
    if ((size_t)OD * PB + (size_t)S * 4 * PB + fixed > maxSmem) return false;
    const size_t smem = (size_t)OD * PB + (size_t)S * 4 * PB + fixed;
    if ((long long)ctx->nOwned * p.nChunks > 2147483647LL) return false;

    if (sp.fastOnly < 0) {
        // Tets whose four faces are all paired with a neighbour (or source) row and that push no halo
        // copy take the kernel without the boundary-condition branches.  When both kinds exist they
        // are launched separately — boundary/halo tets first, so that their peer stores overlap the
        // interior work — each over its own list (device order kept, so bricks still mean locality).
        std::vector<int32_t> gen, fast;
        for (int t = 0; t < ctx->nOwned; t++) {
            const TetRec& r = sp.recHost[t];
            bool g = r.pushPeer[0] >= 0;
            for (int f = 0; f < 4; f++)
                if (!(r.bc[f] == VT_PBC_NONBOUNDARY || r.bc[f] == VT_PBC_PERIODIC || r.bc[f] == VT_PBC_SOURCE)) g = true;
            (g ? gen : fast).push_back(t);
        }
        sp.fastOnly = gen.empty() ? 1 : 0;
        if (sp.tetLists) VT_CUDA(cudaFree(sp.tetLists));
        sp.tetLists = nullptr;
        sp.nGeneric = (int)gen.size();
        sp.nFast = (int)fast.size();
        if (!gen.empty()) {
            gen.insert(gen.end(), fast.begin(), fast.end());
            VT_CUDA(cudaMalloc(&sp.tetLists, gen.size() * sizeof(int32_t)));
            VT_CUDA(cudaMemcpy(sp.tetLists, gen.data(), gen.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        }
    }
    if (!ctx->workCounter) VT_CUDA(cudaMalloc(&ctx->workCounter, 2 * sizeof(unsigned long long)));
    BulkParams P;
    P.s = p;
    P.OD = OD;
    P.S = S;
    P.planeElems = PE;
    P.PV = PV;
    P.queue = ctx->workCounter;
    // variant bit 8: work items in tet-major order, the chunks of one tet adjacent in the queue
    P.orderMode = (ctx->variant & 256) ? 1 : 0;
    P.vmin0f = (float)sp.vmin[0];
    P.vmin1f = (float)sp.vmin[1];
    P.vmin2f = (float)sp.vmin[2];
    P.step0f = (float)sp.step[0];
    P.step1f = (float)sp.step[1];
    P.step2f = (float)sp.step[2];
    {
        double vs = 0;
        for (int k = 0; k < 3; k++) vs += std::max(std::fabs(sp.vmin[k]), std::fabs(sp.vmax[k]));
        // single-precision evaluation of n.v (unit normal) is good to ~1e-6 of this sum, including the
        // running sum of v2 over a row; the band makes the predicate err on the side of loading
        P.guard = (float)(1e-4 * vs);
    }

    const int sms = ctx->prop.multiProcessorCount;
    auto run = [&](const int* list, int nTets, bool allFast, cudaStream_t stream, int maxCTAs, int counter) {
        P.tetList = list;
        P.nTets = nTets;
        P.total = (long long)nTets * p.nChunks;
        P.queue = ctx->workCounter + counter;
        if (wide && n0 == 32 && n1 == 32) {
            if (nbrSelf && S == 4) launch_cfg<2, 8, 32, 4>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
            else launch_cfg<2, 8, 32, 0>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
        } else if (wide) {
            if (nbrSelf && S == 4) launch_cfg<2, 8, 0, 4>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
            else if (nbrSelf && S == 2) launch_cfg<2, 8, 0, 2>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
            else launch_cfg<2, 8, 0, 0>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
        } else if (wide8) {
            const int nbd = nbrSelf ? S : 0;   // S is 4 or 2 here
            if (kpt == 3) {
                if (nbd == 4) launch_cfg<3, 8, 0, 4>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
                else if (nbd == 2) launch_cfg<3, 8, 0, 2>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
                else launch_cfg<3, 8, 0, 0>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
            } else if (kpt == 4) {
                if (nbd == 4) launch_cfg<4, 8, 0, 4>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
                else if (nbd == 2) launch_cfg<4, 8, 0, 2>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
                else launch_cfg<4, 8, 0, 0>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
            } else {
                if (nbd == 4) launch_cfg<5, 8, 0, 4>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
                else if (nbd == 2) launch_cfg<5, 8, 0, 2>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
                else launch_cfg<5, 8, 0, 0>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
            }
        } else if (kpt == 1) launch_cfg<1, 16, 0, 0>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
        else if (kpt == 2) launch_cfg<2, 16, 0, 0>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
        else if (kpt == 3) launch_cfg<3, 16, 0, 0>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
        else launch_cfg<4, 16, 0, 0>(ctx, P, upwind, allFast, smem, stream, maxCTAs);
    };
    VT_CUDA(cudaEventRecord(e0, ctx->stream));
    if (sp.fastOnly == 1) {
        run(nullptr, ctx->nOwned, true, ctx->stream, sms, 0);
    } else {
        // boundary/halo tets first, then the interior ones, on the same stream.  (Running the first
        // launch concurrently on a few SMs of a side stream was measured: 32 ms instead of 20 ms per
        // 2-GPU step — a handful of SMs cannot feed the NVLink stores of all ghost rows.)
        // VT_STEP_MIXED=1 (experiment): one launch over the whole list — boundary tets at the head of
        // the queue, interior tets behind them in the same kernel, so that the NVLink stores drain
        // behind the interior work.  Measured slower than the two launches at every N (8 GPUs: 23.7
        // against 22.8 ms per step): the interior path of the kernel that carries the boundary
        // branches costs more than the overlap gains.
        static const bool mixed = std::getenv("VT_STEP_MIXED") && std::atoi(std::getenv("VT_STEP_MIXED")) != 0;
        if (mixed) {
            run(sp.tetLists, sp.nGeneric + sp.nFast, false, ctx->stream, sms, 0);
        } else {
            run(sp.tetLists, sp.nGeneric, false, ctx->stream, sms, 0);
            run(sp.tetLists + sp.nGeneric, sp.nFast, true, ctx->stream, sms, 1);
        }
    }
    VT_CUDA(cudaEventRecord(e1, ctx->stream));
    VT_CUDA(cudaGetLastError());
    p.densSplit = ncw;
    return true;
}

}  // namespace vt
