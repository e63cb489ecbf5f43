// Tucker step, slab-streaming kernel — Solver<Tucker>::_UpdatePDF (src/solver.cpp:141-212) with the
// six Compress calls (src/tucker.cpp:66-98, rank rule :442-465) evaluated as truncated HOSVDs of the
// dense sums (the formulation of tucker.cu, see its header), restructured for the machine:
//
//   * a 512-thread CTA (one per SM) owns a tet for the whole step; dense data only ever exist as ONE velocity
//     slab in shared memory; the rounding's input X and the tet's own f live in a per-CTA global scratch
//     (L2-sized), every Tucker operand (the neighbour, |v.n| of the face, the previous rounded right-hand
//     side) is re-expanded slab by slab from its factors
//   * every contraction runs on the FP64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA): the slab expansions
//     (U0 M2) U1^T, the Gram matrices S S^T / S^T S / C^T C (upper triangle of 8x8 blocks, accumulators in
//     registers for a whole pass), the projection U0^T S
//   * the three symmetric eigen-problems of a rounding are first reduced by a pivoted Cholesky factorisation
//     (the Gram matrix of a smooth tensor has numerical rank 12-36 of 48) and then solved by a parallel
//     two-sided Jacobi method that only rotates where it matters for a rounding of relative error eps
//
// Passes of one rounding (X = what the reference rounds at this point):
//   1  i2 slabs:  T = U0 M2 of every operand, then per 8x8 block: expand all operands into accumulator fragments and
//                 combine them in registers (flux / acceleration / Euler update) -> X to the scratch; one barrier per slab
//   G  i2 slabs of X: G0 += S S^T, G1 += S^T S;  i1 slabs of X (n0 x n2): G2 += C^T C
//   E  pivoted Cholesky G = L L^T, Jacobi on L^T L, U = L W Lambda^(-1/2), rank rule (tucker.cpp:450-461)
//   3  i2 slabs of X:  W += U2(i2, :) (x) (U0^T S) in registers, then core = W x_2 U1^T
//
// Shared memory: ~190 KB dynamic (+9 KB static), so the L1 is almost gone and anything on the stack costs an L2
// round trip: kernel parameters, layout and operand descriptors live in shared memory and every pass is its
// own __noinline__ function.  profiles/r2_tucker_slab_notes.md has the measurement trail.
//
// Served here: compression errors that do not need the small-eps refinement of tucker.cu (eps >= 5.5e-7),
// grids of 33..48 nodes per axis, rank caps <= 16.  Everything else stays on k_tucker.
#include "tucker_internal.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace vt {
namespace {

constexpr int kMaxP = 48;       // padded nodes per axis
constexpr int kMaxSlots = 32;   // rotation slots per Jacobi round (>= kMaxP / 2, one lane each)
constexpr int kMaxSweeps = 30;

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}
__host__ __device__ constexpr int pad_ld(int n)   // smallest ld >= roundup(n, 8) with ld % 16 == 4: both fragment walks are conflict free
{
    const int p = (n + 7) / 8 * 8;
    return p + ((4 - p % 16) + 16) % 16;
}
__device__ __forceinline__ void cp_async16(double* dst, const double* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(double* dst, const double* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One 8x8 output block: acc += sum_k A(row, k) B(k, col), element A(row, k) at A[row*sAr + k*sAk], B(k, col) at
// B[k*sBk + col*sBc], K4 steps of four.  Lane l holds A(l/4, l%4), B(l%4, l/4) and C(l/4, 2(l%4) + {0,1}).
__device__ __forceinline__ void warp_mma(double (&acc)[2], const double* A, int sAr, int sAk, const double* B, int sBk, int sBc, int K4)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const double* a = A + g * sAr + c * sAk;
    const double* b = B + c * sBk + g * sBc;
    for (int kk = 0; kk < K4; kk++) dmma884(acc, a[4 * kk * sAk], b[4 * kk * sBk]);
}
__device__ __forceinline__ void warp_store(const double (&acc)[2], double* C, int sCr, int sCc)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    C[g * sCr + (2 * c) * sCc] = acc[0];
    C[g * sCr + (2 * c + 1) * sCc] = acc[1];
}

// Shared-memory map (doubles).  Region E is used twice: by the pass buffers (slabs, operand factors,
// expansion staging) and, between pass 2 and pass 3, by the Gram / eigenvector matrices.
struct SlabLay {
    int p[3], LU[3], LD, slab, rK;
    int oSlab[5];        // A ring (2), B, |v.n|, R / X
    int oFac[3][3];      // operand s (0 neighbour, 1 |v.n|, 2 previous rhs / own f), factor k < 2: LU[k] x rK
    int oT[3];           // U0 M2 of operand s: LD x rK, double buffered
    int oM2[3];          // core x_3 U2(i2, :) of operand s: rK x rK, double buffered
    int oCore[3], coreCap;
    int oG[3], oV[3], oW[3];   // Gram matrix (then L^T L), its eigenvectors, the Cholesky factor L: p[k] columns, leading dimension LU[k]
    int oUnew[3];        // factors chosen by the rounding: LU[k] x rK, columns >= rank are zero
    int total;
};
__host__ __device__ inline SlabLay slab_layout(const int n[3], int rK)
{
    SlabLay L;
    for (int k = 0; k < 3; k++) {
        L.p[k] = (n[k] + 7) / 8 * 8;
        L.LU[k] = pad_ld(L.p[k]);
    }
    L.LD = L.LU[0];
    L.slab = L.LD * (L.p[1] > L.p[2] ? L.p[1] : L.p[2]);
    L.rK = rK;
    int o = 0;
    for (int b = 0; b < 5; b++) {
        L.oSlab[b] = o;
        o += L.slab;
    }
    for (int s = 0; s < 3; s++)
        for (int k = 0; k < 3; k++) {   // the third factor is read where it lies (one row per slab): nothing staged
            L.oFac[s][k] = o;
            if (k < 2) o += L.LU[k] * rK;
        }
    for (int s = 0; s < 3; s++) {   // two sets (slab parity), the second one 3 LD rK further on
        L.oT[s] = o;
        o += L.LD * rK;
    }
    o += 3 * L.LD * rK;
    for (int s = 0; s < 3; s++) {   // likewise, 3 rK^2 further on
        L.oM2[s] = o;
        o += rK * rK;
    }
    o += 3 * rK * rK;
    L.coreCap = rK <= 8 ? rK * rK * rK : 0;   // larger cores are read from global memory (L1)
    for (int s = 0; s < 3; s++) {
        L.oCore[s] = o;
        o += L.coreCap;
    }
    const int passEnd = o;
    o = 0;
    for (int k = 0; k < 3; k++) {
        L.oG[k] = o;
        o += L.LU[k] * L.p[k];
    }
    for (int k = 0; k < 3; k++) {
        L.oV[k] = o;
        o += L.LU[k] * L.p[k];
    }
    for (int k = 0; k < 3; k++) {
        L.oW[k] = o;
        o += L.LU[k] * L.p[k];
    }
    o = o > passEnd ? o : passEnd;
    if (o < L.oV[0] + 4 * L.slab) o = L.oV[0] + 4 * L.slab;   // pass 2 keeps its four slab buffers behind G0, G1, G2
    if (o < rK * rK * L.p[1]) o = rK * rK * L.p[1];           // pass 3 collects W (rK x p1 x rK) at the bottom of the region
    for (int k = 0; k < 3; k++) {
        L.oUnew[k] = o;
        o += L.LU[k] * rK;
    }
    L.total = o;
    return L;
}

// a Tucker tensor staged for slab-wise expansion
struct Operand {
    const double* core;   // r0 x r1 x r2 packed (shared or global)
    const double* U[3];   // where the factors come from (global state, |v.n| tables, or the previous rounding's in shared memory)
    int ldu[3];
    int r[3];
    int on;
};

__device__ __forceinline__ int ceil8(int x) { return (x + 7) & ~7; }
__device__ __forceinline__ int ceil4(int x) { return (x + 3) & ~3; }

// idx -> block pair (I, J), I <= J, idx = I + J (J + 1) / 2
__device__ __forceinline__ void pair_of(int idx, int& I, int& J)
{
    J = 0;
    while ((J + 1) * (J + 2) / 2 <= idx) J++;
    I = idx - J * (J + 1) / 2;
}

struct JacobiWork {
    double* G[3];
    double* V[3];
    double* W[3];
    int n[3], p[3], ld[3];
    double* rotC;        // [3][kMaxSlots] rotation of a slot: cosine, sine
    double* rotS;
    int* rotPQ;          // p | q << 8 | rotated << 16
    int* act;            // [3][64] indices that still have an off-diagonal entry above the threshold
    int* na;             // [3] their number (made even with an idle index)
    unsigned* mask;      // [3][2] the same set as bits
    double* crit;        // [3][3] see needs_rotation
    double* invTr;       // [3] 1 / trace
    long long* prof;
};

// Does the pair (p, q) still need a rotation?  Tolerance-aware: the eigenvectors only have to be good enough
// for a rounding with relative error eps.  Leaving g_pq alone leaves the two directions mixed by the angle
// g_pq / (g_max - g_min); if one of them is kept and the other dropped, the rounded tensor changes by
// ~ |g_pq| / sigma_max relative to |X| = sqrt(tr).  So a pair is rotated only while
//     g_pq^2 > (eta eps)^2 tr max(g_pp, g_qq)            (eta = 1e-3: a thousandth of the compression error)
// and never when both directions are certain to be dropped (g_pp + g_qq below a tenth of the rank
// threshold eps^2 tr / 3: the 2x2 block is positive semi-definite, its eigenvalues are below its trace),
// nor below the rounding noise of the Gram matrix itself.  This keeps the numerically-zero trailing block —
// most of the matrix — out of the sweeps.  crit = {floor, (eta eps)^2 tr, 0.03 eps^2 tr}; the floor 0.02 eps^2 tr
// is the same bound for a direction at the rank threshold (sigma = eps sqrt(tr / 3)): 3 % of eps, and never
// below the rounding noise of the Gram matrix 1e-16 tr.
__device__ __forceinline__ bool needs_rotation(double gpq, double gpp, double gqq, const double* crit)
{
    const double a2 = gpq * gpq, gp = fabs(gpp), gq = fabs(gqq);
    return fabs(gpq) > crit[0] && a2 > crit[1] * fmax(gp, gq) && gp + gq > crit[2] && a2 > 1e-32 * gp * gq;
}

// The rotation (c, s) for the pair with off-diagonal gpq and diagonal difference d = gqq - gpp.  The tangent is
// computed in single precision (no double-precision division or square root on the critical path of a round):
// an inexact angle only means that g_pq is not annihilated completely, which the next sweep sees.  What has to
// be exact is c^2 + s^2 = 1 — the transformation must stay orthogonal — so the pair is renormalised in double
// precision with the series of 1 / sqrt(1 + delta).  invTr scales the entries into single-precision range.
__device__ __forceinline__ void rotation_of(double gpq, double d, double invTr, double& c, double& sn)
{
    const float af = (float)(2.0 * gpq * invTr), df = (float)(d * invTr);
    const float den = fabsf(df) + sqrtf(fmaf(df, df, af * af));
    float tf = den > 0.0f ? af / den : 0.0f;
    if (df < 0.0f) tf = -tf;
    const float cf = rsqrtf(fmaf(tf, tf, 1.0f));
    double cc = (double)cf, ss = (double)(tf * cf);
    const double delta = fma(cc, cc, fma(ss, ss, -1.0));
    const double corr = fma(delta, fma(delta, 0.375, -0.5), 1.0);
    c = cc * corr;
    sn = ss * corr;
}

// Parallel two-sided Jacobi on three symmetric matrices at once (scripts/prototypes/jacobi_round_robin.py
// states the method on the CPU).  Before every sweep the ACTIVE indices are collected — rows that still
// hold an off-diagonal entry that needs a rotation — and only they are paired (circle method: na - 1
// rounds of na / 2 disjoint pairs).  A round is two phases and two barriers.  Phase 1: one lane per pair
// computes the rotation that annihilates g_pq.  Phase 2: the rotations of a round commute; G <- J^T G J
// is applied to the 2x2 blocks (rows of pair i, columns of pair j, i <= j, mirrored write — each block
// needs only itself and the two rotations), to the rows of the idle indices and to V <- V J.
// Ends when no index is active.  On exit the columns of V are the eigenvectors and the diagonal of G the
// eigenvalues, both to the tolerance above.
// Work split: matrix k belongs to the warps GW k .. GW k + GW - 1, GW = 5 of sixteen (the last warp only keeps the barriers).  What a
// round costs is the length of the instruction stream of its slowest warp — about five cycles per instruction,
// nothing else to overlap with — so each warp runs the code of ONE matrix (a version that staged the loads of
// all three matrices in every warp for more instruction-level parallelism took 2.4x as long per round).
template <int T>
__device__ void jacobi3(const JacobiWork& wShared)
{
    constexpr int GW = (T / 32) / 3;   // warps per matrix: five of sixteen (two of eight in the 256-thread instance)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = warp / GW < 3 ? warp / GW : 2, wg = warp % GW;
    const bool member = warp < 3 * GW;
    const JacobiWork& w = wShared;
    const int n = w.n[k], ld = w.ld[k], pk = w.p[k];
    double* const G = w.G[k];
    double* const V = w.V[k];
    int* const act = w.act + 64 * k;
    double* const rotC = w.rotC + kMaxSlots * k;
    double* const rotS = w.rotS + kMaxSlots * k;
    int* const rotPQ = w.rotPQ + kMaxSlots * k;
    const double crit[3] = {w.crit[3 * k], w.crit[3 * k + 1], w.crit[3 * k + 2]};
    const double invTr = w.invTr[k];
    long long* const wprof = tid == 0 ? w.prof : nullptr;
    for (int sweep = 0; sweep < kMaxSweeps; sweep++) {
        long long tj = wprof ? clock64() : 0;
        if (wprof && sweep == 0)
            for (int q = 18; q < 24; q++) wprof[q] = 0;
        // ---- active sets: rows wg, wg + 5, ... of matrix k, the lanes share the columns
        if (tid < 6) w.mask[tid] = 0;
        __syncthreads();
        if (member) {
            for (int i = wg; i < n; i += GW) {
                const double gii = G[i + ld * i];
                bool hit = false;
                for (int q = lane; q < n; q += 32)
                    if (q != i) hit |= needs_rotation(G[i + ld * q], gii, G[q + ld * q], crit);
                if (__any_sync(0xffffffffu, hit) && lane == 0) atomicOr(w.mask + 2 * k + (i >> 5), 1u << (i & 31));
            }
        }
        __syncthreads();
        if (member && wg == 0) {
            unsigned m0 = w.mask[2 * k], m1 = w.mask[2 * k + 1];
            int na = __popc(m0) + __popc(m1);
            if (na & 1) {   // an idle index completes the last pair (it exists: the padded dimension is even)
                int d = 0;
                while (d < pk && ((d < 32 ? m0 >> d : m1 >> (d - 32)) & 1u)) d++;
                if (d < 32) m0 |= 1u << d;
                else m1 |= 1u << (d - 32);
                na++;
            }
            const unsigned below = (1u << lane) - 1u;
            if ((m0 >> lane) & 1u) act[__popc(m0 & below)] = lane;
            if ((m1 >> lane) & 1u) act[__popc(m0) + __popc(m1 & below)] = 32 + lane;
            if (lane == 0) {
                w.na[k] = na;
                w.mask[2 * k] = m0;
                w.mask[2 * k + 1] = m1;
            }
        }
        __syncthreads();
        const int na = w.na[k];
        const int maxRounds = max(w.na[0], max(w.na[1], w.na[2])) - 1;
        if (maxRounds <= 0) break;
        const unsigned m0 = w.mask[2 * k], m1 = w.mask[2 * k + 1];
        const int h = na >> 1;
        if (wprof) {
            if (sweep < 6) wprof[18 + sweep] = w.na[0] | (w.na[1] << 8) | (w.na[2] << 16);   // the last eigen-problem's active sets
            wprof[13] += maxRounds;
            wprof[14] += 1;
            const long long now = clock64();
            wprof[15] += now - tj;
            tj = now;
        }
        for (int s = 0; s < maxRounds; s++) {
            const bool live = member && s < na - 1;
            // ---- phase 1
            if (live && wg == 0 && lane < h) {
                int a, b;
                if (lane == 0) {
                    a = na - 1;
                    b = s;
                } else {
                    a = s + lane;
                    if (a >= na - 1) a -= na - 1;
                    b = s - lane;
                    if (b < 0) b += na - 1;
                }
                int p = act[a], q = act[b];
                if (p > q) {
                    const int xx = p;
                    p = q;
                    q = xx;
                }
                const double gpq = G[p + ld * q], gpp = G[p + ld * p], gqq = G[q + ld * q];
                double c = 1.0, sn = 0.0;
                int rotated = 0;
                if (needs_rotation(gpq, gpp, gqq, crit)) {
                    rotation_of(gpq, gqq - gpp, invTr, c, sn);
                    rotated = 1;
                }
                rotC[lane] = c;
                rotS[lane] = sn;
                rotPQ[lane] = p | (q << 8) | (rotated << 16);
            }
            __syncthreads();
            if (wprof) {
                const long long now = clock64();
                wprof[16] += now - tj;
                tj = now;
            }
            // ---- phase 2
            if (live) {
                // 2x2 blocks: lane = pair i, warp of the group (+ 5) = pair j
                for (int j = wg; j < h; j += GW) {
                    const int i = lane;
                    if (i > j) continue;
                    const int pqi = rotPQ[i], pqj = rotPQ[j];
                    if (((pqi | pqj) >> 16) == 0) continue;
                    const int P = pqi & 255, Q = (pqi >> 8) & 255, R = pqj & 255, S = (pqj >> 8) & 255;
                    const double ci = rotC[i], si = rotS[i], cj = rotC[j], sj = rotS[j];
                    const double mPR = G[P + ld * R], mPS = G[P + ld * S], mQR = G[Q + ld * R], mQS = G[Q + ld * S];
                    const double nPR = cj * mPR - sj * mPS, nPS = sj * mPR + cj * mPS;
                    const double nQR = cj * mQR - sj * mQS, nQS = sj * mQR + cj * mQS;
                    const double oPR = ci * nPR - si * nQR, oQR = si * nPR + ci * nQR;
                    const double oPS = ci * nPS - si * nQS, oQS = si * nPS + ci * nQS;
                    if (i == j) {   // rows and columns are the same pair; what is left of g_PQ is kept: the similarity stays exact
                        G[P + ld * P] = oPR;
                        G[Q + ld * Q] = oQS;
                        G[P + ld * Q] = oPS;
                        G[Q + ld * P] = oPS;
                    } else {
                        G[P + ld * R] = oPR;
                        G[R + ld * P] = oPR;
                        G[P + ld * S] = oPS;
                        G[S + ld * P] = oPS;
                        G[Q + ld * R] = oQR;
                        G[R + ld * Q] = oQR;
                        G[Q + ld * S] = oQS;
                        G[S + ld * Q] = oQS;
                    }
                }
                // rows: V <- V J, and G(x, .) of the idle indices x
                for (int j = wg; j < h; j += GW) {
                    const int pq = rotPQ[j];
                    if ((pq >> 16) == 0) continue;
                    const int P = pq & 255, Q = (pq >> 8) & 255;
                    const double c = rotC[j], sn = rotS[j];
                    for (int x = lane; x < n; x += 32) {
                        const double vp = V[x + ld * P], vq = V[x + ld * Q];
                        V[x + ld * P] = c * vp - sn * vq;
                        V[x + ld * Q] = sn * vp + c * vq;
                        if (((x < 32 ? m0 >> x : m1 >> (x - 32)) & 1u) == 0) {
                            const double gp = G[x + ld * P], gq = G[x + ld * Q];
                            const double np = c * gp - sn * gq, nq = sn * gp + c * gq;
                            G[x + ld * P] = np;
                            G[P + ld * x] = np;
                            G[x + ld * Q] = nq;
                            G[Q + ld * x] = nq;
                        }
                    }
                }
            }
            __syncthreads();
            if (wprof) {
                const long long now = clock64();
                wprof[17] += now - tj;
                tj = now;
            }
        }
    }
}

// The same method for ONE matrix on ONE warp (no block barrier): after the Cholesky reduction the problems are
// small (12-20 rows for smooth tensors), a round then is a handful of rotations and the three matrices run side
// by side on three warps.  Used when every reduced matrix has at most 24 rows; larger ones take jacobi3.
__device__ void jacobi_warp(const JacobiWork& ws, int k)
{
    const int lane = threadIdx.x & 31;
    const int n = ws.n[k], ld = ws.ld[k], pk = ws.p[k];
    double* const G = ws.G[k];
    double* const V = ws.V[k];
    int* const act = ws.act + 64 * k;
    double* const rotC = ws.rotC + kMaxSlots * k;
    double* const rotS = ws.rotS + kMaxSlots * k;
    int* const rotPQ = ws.rotPQ + kMaxSlots * k;
    const double crit[3] = {ws.crit[3 * k], ws.crit[3 * k + 1], ws.crit[3 * k + 2]};
    const double invTr = ws.invTr[k];
    long long* const wprof = (lane == 0 && k == 0) ? ws.prof : nullptr;
    for (int sweep = 0; sweep < kMaxSweeps; sweep++) {
        bool hit0 = false, hit1 = false;
        if (lane < n) {
            const double gii = G[lane + ld * lane];
            for (int q = 0; q < n; q++)
                if (q != lane) hit0 |= needs_rotation(G[lane + ld * q], gii, G[q + ld * q], crit);
        }
        if (lane + 32 < n) {
            const int r = lane + 32;
            const double gii = G[r + ld * r];
            for (int q = 0; q < n; q++)
                if (q != r) hit1 |= needs_rotation(G[r + ld * q], gii, G[q + ld * q], crit);
        }
        unsigned m0 = __ballot_sync(0xffffffffu, hit0), m1 = __ballot_sync(0xffffffffu, hit1);
        int na = __popc(m0) + __popc(m1);
        if (na == 0) break;
        if (na & 1) {   // an idle index completes the last pair (it exists: the padded dimension is even)
            int d = 0;
            while (d < pk && ((d < 32 ? m0 >> d : m1 >> (d - 32)) & 1u)) d++;
            if (d < 32) m0 |= 1u << d;
            else m1 |= 1u << (d - 32);
            na++;
        }
        const unsigned below = (1u << lane) - 1u;
        if ((m0 >> lane) & 1u) act[__popc(m0 & below)] = lane;
        if ((m1 >> lane) & 1u) act[__popc(m0) + __popc(m1 & below)] = 32 + lane;
        __syncwarp();
        const int h = na >> 1;
        if (wprof) {
            wprof[13] += na - 1;
            wprof[14] += 1;
        }
        for (int s = 0; s < na - 1; s++) {
            if (lane < h) {
                int a, b;
                if (lane == 0) {
                    a = na - 1;
                    b = s;
                } else {
                    a = s + lane;
                    if (a >= na - 1) a -= na - 1;
                    b = s - lane;
                    if (b < 0) b += na - 1;
                }
                int p = act[a], q = act[b];
                if (p > q) {
                    const int xx = p;
                    p = q;
                    q = xx;
                }
                const double gpq = G[p + ld * q], gpp = G[p + ld * p], gqq = G[q + ld * q];
                double c = 1.0, sn = 0.0;
                int rotated = 0;
                if (needs_rotation(gpq, gpp, gqq, crit)) {
                    rotation_of(gpq, gqq - gpp, invTr, c, sn);
                    rotated = 1;
                }
                rotC[lane] = c;
                rotS[lane] = sn;
                rotPQ[lane] = p | (q << 8) | (rotated << 16);
            }
            __syncwarp();
            // 2x2 blocks of G over the pairs i <= j
            for (int e = lane; e < h * h; e += 32) {
                const int i = e % h, j = e / h;
                if (i > j) continue;
                const int pqi = rotPQ[i], pqj = rotPQ[j];
                if (((pqi | pqj) >> 16) == 0) continue;
                const int P = pqi & 255, Q = (pqi >> 8) & 255, R = pqj & 255, S = (pqj >> 8) & 255;
                const double ci = rotC[i], si = rotS[i], cj = rotC[j], sj = rotS[j];
                const double mPR = G[P + ld * R], mPS = G[P + ld * S], mQR = G[Q + ld * R], mQS = G[Q + ld * S];
                const double nPR = cj * mPR - sj * mPS, nPS = sj * mPR + cj * mPS;
                const double nQR = cj * mQR - sj * mQS, nQS = sj * mQR + cj * mQS;
                const double oPR = ci * nPR - si * nQR, oQR = si * nPR + ci * nQR;
                const double oPS = ci * nPS - si * nQS, oQS = si * nPS + ci * nQS;
                if (i == j) {   // what is left of g_PQ is kept: the similarity transformation stays exact
                    G[P + ld * P] = oPR;
                    G[Q + ld * Q] = oQS;
                    G[P + ld * Q] = oPS;
                    G[Q + ld * P] = oPS;
                } else {
                    G[P + ld * R] = oPR;
                    G[R + ld * P] = oPR;
                    G[P + ld * S] = oPS;
                    G[S + ld * P] = oPS;
                    G[Q + ld * R] = oQR;
                    G[R + ld * Q] = oQR;
                    G[Q + ld * S] = oQS;
                    G[S + ld * Q] = oQS;
                }
            }
            // rows: V <- V J, and G(x, .) of the idle indices x
            for (int e = lane; e < n * h; e += 32) {
                const int x = e % n, j = e / n;
                const int pq = rotPQ[j];
                if ((pq >> 16) == 0) continue;
                const int P = pq & 255, Q = (pq >> 8) & 255;
                const double c = rotC[j], sn = rotS[j];
                const double vp = V[x + ld * P], vq = V[x + ld * Q];
                V[x + ld * P] = c * vp - sn * vq;
                V[x + ld * Q] = sn * vp + c * vq;
                if (((x < 32 ? m0 >> x : m1 >> (x - 32)) & 1u) == 0) {
                    const double gp = G[x + ld * P], gq = G[x + ld * Q];
                    const double np = c * gp - sn * gq, nq = sn * gp + c * gq;
                    G[x + ld * P] = np;
                    G[P + ld * x] = np;
                    G[x + ld * Q] = nq;
                    G[Q + ld * x] = nq;
                }
            }
            __syncwarp();
        }
    }
}

// Per-thread plan of one slab copy global -> shared through cp.async: which 16-byte (8-byte for an odd n0) pieces
// this thread moves.  Built once per pass, so that the index arithmetic — a division and a modulo per piece —
// stays out of the slab loops.  Slab element (r, c), r < rows contiguous in both spaces, sits at shared r + LD c
// and at global r + gStride c.
template <int T>
struct CopyPlan {
    static constexpr int kMax = (kMaxP * kMaxP + T - 1) / T;
    int n;
    int soff[kMax], goff[kMax];
    bool vec2;
    __device__ __forceinline__ void build(int tid, int rows, int cols, int LD, int gStride)
    {
        vec2 = (rows % 2) == 0 && (gStride % 2) == 0;
        const int w = vec2 ? rows / 2 : rows, total = w * cols, step = vec2 ? 2 : 1;
        n = 0;
#pragma unroll
        for (int m = 0; m < kMax; m++) {
            const int e = tid + T * m;
            soff[m] = goff[m] = 0;
            if (e < total) {
                const int r = step * (e % w), c = e / w;
                soff[m] = r + LD * c;
                goff[m] = r + gStride * c;
                n = m + 1;
            }
        }
    }
    __device__ __forceinline__ void issue(double* buf, const double* src) const
    {
#pragma unroll
        for (int m = 0; m < kMax; m++) {
            if (m < n) {
                if (vec2) cp_async16(buf + soff[m], src + goff[m]);
                else cp_async8(buf + soff[m], src + goff[m]);
            }
        }
    }
};

enum { KIND_FACE = 0, KIND_ACCEL = 1, KIND_FINAL = 2 };

// Everything the passes share, in SHARED memory.  With ~190 KB of shared memory in use the L1 cache is almost
// gone: a value on the stack (a spilled accumulator, a struct indexed at run time, kernel parameters whose
// address is taken) costs an L2 round trip.  So the kernel parameters are copied here once, the state of the
// rounding in progress is written here by thread 0, and each pass is its own __noinline__ function that
// takes only this record: register allocation starts afresh in every pass and nothing long-lived crowds the
// accumulators out of the register file.
template <int NW, int PPW>
struct SlabShared {
    TuckerParams P;
    SlabLay L;
    Operand op[3];
    JacobiWork jw;
    TetRec rec;
    // the rounding in progress
    int t, rd, kind, f, pairBC, srcBC, absBC, genA, useR, wallOn;
    double coef, nrm[3], gacc[3];
    double *Aglob, *Xglob, *coreW;
    int profOn;
    int kdim[3];     // numerical rank of each Gram matrix = size of the eigen-problem actually solved
    double trace[3];
    // eigen-solve and rank selection
    double lam[3][kMaxN];
    int ord[3][kMaxN];
    int rsel[3];
    double rotC[3 * kMaxSlots], rotS[3 * kMaxSlots];
    int rotPQ[3 * kMaxSlots];
    int act[3 * 64], na[3];
    unsigned mask[6];
    double crit[9], invTr[3];
    double red[NW][5];
    double colSum[3][16];
    // Gram tasks of every warp, one packed word each: mode (pass 1) or half of the sum (pass 2) | I << 4 | J << 8; 15 = none
    unsigned short task1[NW * PPW], task2[NW * PPW];
    long long prof[24];
};

// slab buffers 0..4 of the pass area: the A ring (0, 1), the neighbour (2), |v.n| (3), previous rhs / X (4)
__device__ __forceinline__ double* slab_buf(double* sm, const SlabLay& L, int b) { return sm + b * L.slab; }

// ---- operands of the rounding: factors zero padded to LU[k] x rK, small cores into shared memory
template <int T, int NW, int PPW>
__device__ __noinline__ void stage_operands(SlabShared<NW, PPW>& S, double* sm)
{
    const int tid = threadIdx.x;
    const SlabLay& L = S.L;
    const int rK = L.rK;
    if (tid == 0) {
        const TuckerParams& P = S.P;
        const int f = S.f, t = S.t, n0 = P.n[0], n1 = P.n[1];
        for (int s = 0; s < 3; s++) {
            Operand& o = S.op[s];
            o.on = s == 0 ? S.pairBC : (s == 1 ? (S.pairBC || S.srcBC || S.absBC) : 1);
            if (!o.on) continue;
            if (s == 0 || (s == 2 && S.genA)) {
                const int row = s == 0 ? S.rec.nbr[f] : t;
                const double* base = P.in + (size_t)row * P.slot;
                o.core = base;
                o.U[0] = base + P.coreCap;
                o.U[1] = o.U[0] + (size_t)n0 * P.rcap[0];
                o.U[2] = o.U[1] + (size_t)n1 * P.rcap[1];
                for (int k = 0; k < 3; k++) {
                    o.r[k] = P.rin[3 * row + k];
                    o.ldu[k] = P.n[k];
                }
            } else if (s == 1) {
                const double* vs = P.vnabs + ((size_t)t * 4 + f) * P.vslot;
                o.core = vs;
                o.U[0] = vs + 216;
                o.U[1] = vs + 216 + 6 * n0;
                o.U[2] = vs + 216 + 6 * (n0 + n1);
                const int* vr = P.vnabsRanks + ((size_t)t * 4 + f) * 3;
                for (int k = 0; k < 3; k++) {
                    o.r[k] = vr[k];
                    o.ldu[k] = P.n[k];
                }
            } else {
                o.core = S.coreW;
                for (int k = 0; k < 3; k++) {
                    o.U[k] = sm + L.oUnew[k];
                    o.r[k] = S.rsel[k];   // ranks chosen by the previous rounding
                    o.ldu[k] = L.LU[k];
                }
            }
        }
    }
    for (int e = tid; e < 5 * L.slab; e += T) sm[e] = 0.0;
    __syncthreads();
    for (int s = 0; s < 3; s++) {
        if (!S.op[s].on) continue;
        for (int k = 0; k < 2; k++) {
            double* dst = sm + L.oFac[s][k];
            const double* src = S.op[s].U[k];
            const int lu = L.LU[k], nk = S.P.n[k], rk = S.op[s].r[k], ldu = S.op[s].ldu[k];
            for (int e = tid; e < lu * rK; e += T) {
                const int i = e % lu, a = e / lu;
                dst[e] = (i < nk && a < rk) ? src[i + (size_t)ldu * a] : 0.0;
            }
        }
        if (L.coreCap > 0) {
            const int cn = S.op[s].r[0] * S.op[s].r[1] * S.op[s].r[2];
            double* dst = sm + L.oCore[s];
            const double* src = S.op[s].core;
            for (int e = tid; e < cn; e += T) dst[e] = src[e];
        }
    }
    __syncthreads();
    if (tid == 0 && L.coreCap > 0)
        for (int s = 0; s < 3; s++)
            if (S.op[s].on) S.op[s].core = sm + L.oCore[s];
    __syncthreads();
}

// ---- pass 1: the rounding's input X, slab by slab along i2, into the global scratch (L2).
// Every Tucker operand (neighbour, |v.n|, previous rounded right-hand side — or the tet's own tensor in the first
// rounding, which is also kept dense in Aglob) is expanded as  slab = (U0 M2) U1^T,  M2 = core x_3 U2(i2, :).
// One barrier per slab:  phase A  T(i2) = U0 M2(i2) and M2(i2 + 1) for every operand (both double buffered),
// phase B  a warp takes an 8x8 block of the slab, expands ALL operands on the tensor cores into its accumulator
// fragments and combines them right there in registers (flux / acceleration / Euler update): the expanded
// operands never touch shared memory.  The tet's own dense slab comes from Aglob through a cp.async ring.
// Returns this thread's part of the sum of the face flux (wall charge).
template <int T, int NW, int PPW>
__device__ __noinline__ double pass1(SlabShared<NW, PPW>& S, double* sm)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
    const SlabLay& L = S.L;
    const int n0 = S.P.n[0], n1 = S.P.n[1], n2 = S.P.n[2], M = n0 * n1;
    const int p0 = L.p[0], LD = L.LD, rK = L.rK;
    const int nb0 = p0 / 8, nblk = nb0 * (L.p[1] / 8);
    const int kind = S.kind;
    const bool genA = S.genA, useR = S.useR && !S.genA, pairBC = S.pairBC, srcBC = S.srcBC, absBC = S.absBC, wallOn = S.wallOn;
    const bool vec2 = (n0 % 2) == 0;
    const bool prof = S.profOn && tid == 0;
    double* const Aglob = S.Aglob;
    double* const Xglob = S.Xglob;
    const int tSet = 3 * LD * rK, mSet = 3 * rK * rK, lu1 = L.LU[1];
    const bool on0 = S.op[0].on, on1 = S.op[1].on;
    const int K40 = on0 ? ceil4(S.op[0].r[1]) / 4 : 0, K41 = on1 ? ceil4(S.op[1].r[1]) / 4 : 0, K42 = ceil4(S.op[2].r[1]) / 4;
    const double vmin0 = S.P.vmin[0], vmin1 = S.P.vmin[1], st0 = S.P.step[0], st1 = S.P.step[1];
    const double coef = S.coef, nrm0 = S.nrm[0], nrm1 = S.nrm[1];
    const double g0 = S.gacc[0], g1 = S.gacc[1], g2 = S.gacc[2], dt = S.P.dt;
    double wallSum = 0.0;
    // (a CopyPlan here costs more in spilled registers than the index arithmetic it saves: measured)
    auto issueA = [&](int j, double* buf) {
        const double* src = Aglob + (size_t)j * M;
        if (vec2) {
            const int h0 = n0 / 2;
            for (int e = tid; e < M / 2; e += T) cp_async16(buf + 2 * (e % h0) + LD * (e / h0), src + 2 * e);
        } else {
            for (int e = tid; e < M; e += T) cp_async8(buf + (e % n0) + LD * (e / n0), src + e);
        }
    };
    // the dense slab j of the tet's own tensor lives in buffer ringA(j): with one barrier per slab a buffer may be
    // refilled only two slabs after its last use — three buffers (0, 1, 3; 2 holds a source PDF's slab) for the
    // face and Euler roundings, five for the acceleration term (which also reads the slabs i2 - 1 and i2 + 1)
    auto ringA = [&](int j) {
        if (kind == KIND_ACCEL) return slab_buf(sm, L, j % 5);
        const int q = j % 3;
        return slab_buf(sm, L, q == 2 ? 3 : q);
    };
    auto make_M2 = [&](int i2, int set) {
        for (int s = 0; s < 3; s++) {
            if (!S.op[s].on) continue;
            const int r0 = S.op[s].r[0], r1 = S.op[s].r[1], r2 = S.op[s].r[2];
            const int r0K = ceil4(r0), r1K = ceil8(r1);
            const double* u2 = S.op[s].U[2] + i2;   // global state / |v.n| table, or the previous rounding's factor in shared memory
            const int lu2 = S.op[s].ldu[2];
            const double* core = S.op[s].core;
            double* M2 = sm + L.oM2[s] + set * mSet;
            for (int e = tid; e < r0K * r1K; e += T) {
                const int a = e % r0K, b = e / r0K;
                double v = 0.0;
                if (a < r0 && b < r1)
                    for (int q = 0; q < r2; q++) v = fma(core[a + r0 * (b + r1 * q)], u2[lu2 * q], v);
                M2[a + rK * b] = v;
            }
        }
    };
    if (!genA) {
        issueA(0, ringA(0));
        cp_async_commit();
        if (kind == KIND_ACCEL) {
            if (n2 > 1) issueA(1, ringA(1));
            cp_async_commit();
        }
    }
    make_M2(0, 0);
    __syncthreads();
    for (int i2 = 0; i2 < n2; i2++) {
        long long tq = prof ? clock64() : 0;
        const int set = i2 & 1;
        // ---- phase A: T = U0 M2 on the tensor cores (p0 x r1, the columns a phase-B step reads beyond the rank are
        // zero because M2's are), M2 of the next slab, prefetches
        {
            int task = warp;
            for (int s = 0; s < 3; s++) {
                if (!S.op[s].on) continue;
                const int K4 = ceil4(S.op[s].r[0]) / 4, nJ = ceil8(S.op[s].r[1]) / 8;
                const double* U0 = sm + L.oFac[s][0];
                const double* M2 = sm + L.oM2[s] + set * mSet;
                double* Ts = sm + L.oT[s] + set * tSet;
                for (; task < nb0 * nJ; task += NW) {
                    const int I = task % nb0, J = task / nb0;
                    double acc[2] = {0.0, 0.0};
                    warp_mma(acc, U0 + 8 * I, 1, LD, M2 + rK * 8 * J, 1, rK, K4);
                    warp_store(acc, Ts + 8 * I + LD * 8 * J, 1, LD);
                }
                task -= nb0 * nJ;
            }
        }
        if (i2 + 1 < n2) make_M2(i2 + 1, set ^ 1);
        if (!genA) {
            if (kind == KIND_ACCEL) {
                if (i2 + 2 < n2) issueA(i2 + 2, ringA(i2 + 2));
            } else {
                if (i2 + 1 < n2) issueA(i2 + 1, ringA(i2 + 1));
            }
        }
        if (srcBC) {   // the source PDF takes the neighbour's place (solver.cpp:335-338); kept dense on the device
            const double* srow = S.P.src + (size_t)(-2 - S.rec.nbr[S.f]) * S.P.N + (size_t)i2 * M;
            double* slabB = slab_buf(sm, L, 2 + 2 * set);   // buffers 2 and 4 in turn: no barrier separates phase B from the next phase A
            for (int e = tid; e < M; e += T) cp_async8(slabB + (e % n0) + LD * (e / n0), srow + e);
        }
        cp_async_commit();
        if (srcBC) cp_async_wait<0>();
        else cp_async_wait<1>();
        __syncthreads();
        if (prof) {
            const long long now = clock64();
            S.prof[9] += now - tq;
            tq = now;
        }
        // ---- phase B: 8x8 blocks of the slab, all operands expanded and combined in registers
        {
            const double* T0 = sm + L.oT[0] + set * tSet + g + LD * c;
            const double* T1 = sm + L.oT[1] + set * tSet + g + LD * c;
            const double* T2 = sm + L.oT[2] + set * tSet + g + LD * c;
            const double* U10 = sm + L.oFac[0][1] + g + lu1 * c;
            const double* U11 = sm + L.oFac[1][1] + g + lu1 * c;
            const double* U12 = sm + L.oFac[2][1] + g + lu1 * c;
            const double* slabB = slab_buf(sm, L, 2 + 2 * set);
            const double* Ac = ringA(i2);
            const double* Ap = (kind == KIND_ACCEL && i2 > 0) ? ringA(i2 - 1) : nullptr;
            const double* An = (kind == KIND_ACCEL && i2 + 1 < n2) ? ringA(i2 + 1) : nullptr;
            const double v2 = __dadd_rn(S.P.vmin[2], __dmul_rn((double)i2, S.P.step[2]));
            const double nv2 = S.nrm[2] * v2;
            double* const Xs = Xglob + (size_t)i2 * M;
            double* const As = Aglob + (size_t)i2 * M;
            for (int task = warp; task < nblk; task += NW) {
                const int I = task % nb0, J = task / nb0;
                double aB[2] = {0.0, 0.0}, aV[2] = {0.0, 0.0}, aR[2] = {0.0, 0.0};
                const int Kmax = max(K40, max(K41, K42));
                for (int kk = 0; kk < Kmax; kk++) {   // three independent accumulation chains
                    if (kk < K40) dmma884(aB, T0[8 * I + 4 * kk * LD], U10[8 * J + 4 * kk * lu1]);
                    if (kk < K41) dmma884(aV, T1[8 * I + 4 * kk * LD], U11[8 * J + 4 * kk * lu1]);
                    if (kk < K42) dmma884(aR, T2[8 * I + 4 * kk * LD], U12[8 * J + 4 * kk * lu1]);
                }
                const int i0 = 8 * I + g;
                const double v0 = __dadd_rn(vmin0, __dmul_rn((double)i0, st0));
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int i1 = 8 * J + 2 * c + u;
                    if (i0 >= n0 || i1 >= n1) continue;
                    const int idx = i0 + LD * i1;
                    const double a = genA ? aR[u] : Ac[idx];
                    const double rp = useR ? aR[u] : 0.0;
                    double x;
                    if (kind == KIND_FACE) {
                        const double v1 = __dadd_rn(vmin1, __dmul_rn((double)i1, st1));
                        const double vn = nrm0 * v0 + nrm1 * v1 + nv2;
                        double flux;
                        if (pairBC || srcBC) {
                            const double b = pairBC ? aB[u] : slabB[idx], va = aV[u];
                            flux = 0.5 * (vn * (b + a) - va * (b - a));   // solver.cpp:325-327
                        } else if (absBC) {
                            const double va = aV[u];
                            flux = 0.5 * (vn * a + va * a);               // solver.cpp:331-332
                            if (wallOn) wallSum += flux;
                        } else {
                            flux = vn * a;                                // Free, solver.cpp:342
                        }
                        x = rp - coef * flux;                             // solver.cpp:168
                        if (genA) As[i0 + n0 * i1] = a;
                    } else if (kind == KIND_ACCEL) {
                        // rhs -= (q/m)(E_k + ext_k) D_k f, D = zero-outside central difference   solver.cpp:187-200, 348-361
                        x = rp;
                        x = x - g0 * ((i0 + 1 < n0 ? Ac[idx + 1] : 0.0) - (i0 > 0 ? Ac[idx - 1] : 0.0));
                        x = x - g1 * ((i1 + 1 < n1 ? Ac[idx + LD] : 0.0) - (i1 > 0 ? Ac[idx - LD] : 0.0));
                        x = x - g2 * ((An ? An[idx] : 0.0) - (Ap ? Ap[idx] : 0.0));
                    } else {
                        x = a + dt * rp;                                  // solver.cpp:207
                    }
                    Xs[i0 + n0 * i1] = x;
                }
            }
        }
        if (prof) S.prof[10] += clock64() - tq;
    }
    cp_async_wait<0>();
    __syncthreads();
    return wallSum;
}

// ---- Gram passes over X in the global scratch (L2): G0 += S S^T and G1 += S^T S over the slabs S = X(:, :, i2),
// then G2 += C^T C over the slabs C = X(:, i1, :).  Nothing but loads, barriers and DMMA in the loops, the
// accumulators of this warp's block pairs stay in registers for a whole pass (kept apart from make_slab, whose
// register needs pushed them onto the stack).  Four slab buffers behind the Gram matrices, prefetch distance 2,
// one barrier per slab.
template <int T, int NW, int PPW>
__device__ __noinline__ void gram_pass01(SlabShared<NW, PPW>& S, double* sm)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const SlabLay& L = S.L;
    const int n0 = S.P.n[0], n1 = S.P.n[1], n2 = S.P.n[2], M = n0 * n1;
    const int p0 = L.p[0], p1 = L.p[1], LD = L.LD;
    double* const ring0 = sm + L.oV[0];
    const int slabSz = L.slab;
    const double* const Xglob = S.Xglob;
    const int g = lane >> 2, c = lane & 3;
    const bool prof = S.profOn && tid == 0;
    long long tp = prof ? clock64() : 0;
    for (int e = tid; e < L.oV[0] + 4 * slabSz; e += T) sm[e] = 0.0;   // G0, G1, G2 and the buffers' padding
    __syncthreads();
    // ---- G0, G1
    {
        double acc[PPW][2];
        int offA[PPW], offB[PPW], kstep[PPW], K[PPW];
#pragma unroll
        for (int j = 0; j < PPW; j++) {
            acc[j][0] = acc[j][1] = 0.0;
            const int tk = S.task1[warp * PPW + j], m = tk & 15, I = (tk >> 4) & 15, J = (tk >> 8) & 15;
            if (m == 0) {          // G0: rows of the slab against rows, the sum runs over i1
                offA[j] = g + 8 * I + LD * c;
                offB[j] = g + 8 * J + LD * c;
                kstep[j] = 4 * LD;
                K[j] = p1 / 4;
            } else if (m == 1) {   // G1: columns against columns, the sum runs over i0
                offA[j] = c + LD * (8 * I + g);
                offB[j] = c + LD * (8 * J + g);
                kstep[j] = 4;
                K[j] = p0 / 4;
            } else {
                offA[j] = offB[j] = kstep[j] = K[j] = 0;
            }
        }
        CopyPlan<T> plan;
        plan.build(tid, n0, n1, LD, n0);
        auto issueS = [&](int j, double* buf) { plan.issue(buf, Xglob + (size_t)j * M); };
        issueS(0, ring0);
        cp_async_commit();
        if (n2 > 1) issueS(1, ring0 + slabSz);
        cp_async_commit();
        const int Kmax = max(p0, p1) / 4;
        for (int i2 = 0; i2 < n2; i2++) {
            if (i2 + 2 < n2) issueS(i2 + 2, ring0 + ((i2 + 2) & 3) * slabSz);
            cp_async_commit();
            cp_async_wait<2>();
            __syncthreads();
            const double* Sl = ring0 + (i2 & 3) * slabSz;
#pragma unroll 1
            for (int kk = 0; kk < Kmax; kk++) {   // the tasks of a warp advance together: independent accumulation chains
#pragma unroll
                for (int j = 0; j < PPW; j++)
                    if (kk < K[j]) dmma884(acc[j], Sl[offA[j] + kk * kstep[j]], Sl[offB[j] + kk * kstep[j]]);
            }
        }
        cp_async_wait<0>();
#pragma unroll
        for (int j = 0; j < PPW; j++) {
            const int tk = S.task1[warp * PPW + j], k = tk & 15;
            if (k == 15) continue;
            const int nk = S.P.n[k], ld = L.LU[k];
            double* G = sm + L.oG[k];
            const int i = 8 * ((tk >> 4) & 15) + g;
            for (int u = 0; u < 2; u++) {
                const int jj = 8 * ((tk >> 8) & 15) + 2 * c + u;
                if (i < nk && jj < nk && i <= jj) {   // diagonal blocks: one triangle, mirrored
                    G[i + ld * jj] = acc[j][u];
                    G[jj + ld * i] = acc[j][u];
                }
            }
        }
        __syncthreads();
    }
    if (prof) S.prof[12] += clock64() - tp;
}

template <int T, int NW, int PPW>
__device__ __noinline__ void gram_pass2(SlabShared<NW, PPW>& S, double* sm)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const SlabLay& L = S.L;
    const int n0 = S.P.n[0], n1 = S.P.n[1], n2 = S.P.n[2], M = n0 * n1;
    const int p0 = L.p[0], LD = L.LD;
    double* const ring0 = sm + L.oV[0];
    const int slabSz = L.slab;
    const double* const Xglob = S.Xglob;
    const int g = lane >> 2, c = lane & 3;
    const bool prof = S.profOn && tid == 0;
    long long tp = prof ? clock64() : 0;
    // ---- G2 (the slabs are n0 x n2 now: clear what the n0 x n1 ones left in the padding)
    for (int e = tid; e < 4 * slabSz; e += T) ring0[e] = 0.0;
    __syncthreads();
    {
        double acc[PPW][2];
        int offA[PPW], offB[PPW], K[PPW];
        const int K4 = p0 / 4, Kh = (K4 + 1) / 2;
#pragma unroll
        for (int j = 0; j < PPW; j++) {
            acc[j][0] = acc[j][1] = 0.0;
            const int tk = S.task2[warp * PPW + j], m = tk & 15, I = (tk >> 4) & 15, J = (tk >> 8) & 15;
            const int k0 = m == 0 ? 0 : Kh;
            offA[j] = 4 * k0 + c + LD * (8 * I + g);
            offB[j] = 4 * k0 + c + LD * (8 * J + g);
            K[j] = m == 15 ? 0 : (m == 0 ? Kh : K4 - Kh);
        }
        CopyPlan<T> plan;
        plan.build(tid, n0, n2, LD, M);
        auto issueC = [&](int i1, double* buf) { plan.issue(buf, Xglob + (size_t)n0 * i1); };
        issueC(0, ring0);
        cp_async_commit();
        if (n1 > 1) issueC(1, ring0 + slabSz);
        cp_async_commit();
        for (int i1 = 0; i1 < n1; i1++) {
            if (i1 + 2 < n1) issueC(i1 + 2, ring0 + ((i1 + 2) & 3) * slabSz);
            cp_async_commit();
            cp_async_wait<2>();
            __syncthreads();
            const double* C = ring0 + (i1 & 3) * slabSz;
#pragma unroll 1
            for (int kk = 0; kk < Kh; kk++) {
#pragma unroll
                for (int j = 0; j < PPW; j++)
                    if (kk < K[j]) dmma884(acc[j], C[offA[j] + 4 * kk], C[offB[j] + 4 * kk]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        for (int half = 0; half < 2; half++) {
#pragma unroll
            for (int j = 0; j < PPW; j++) {
                const int tk = S.task2[warp * PPW + j];
                if ((tk & 15) != half) continue;
                const int ld = L.LU[2];
                double* G = sm + L.oG[2];
                const int i = 8 * ((tk >> 4) & 15) + g;
                for (int u = 0; u < 2; u++) {
                    const int jj = 8 * ((tk >> 8) & 15) + 2 * c + u;
                    if (i < n2 && jj < n2 && i <= jj) {
                        const double v = (half == 0 ? 0.0 : G[i + ld * jj]) + acc[j][u];
                        G[i + ld * jj] = v;
                        G[jj + ld * i] = v;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (prof) S.prof[1] += clock64() - tp;
}

// ---- eigen-decomposition of the three Gram matrices, rank rule (tucker.cpp:450-461), factors.
// The Gram matrix of a smooth tensor is numerically of low rank (20-25 of 48 at the 1e-17 level), so the n x n
// problem is first reduced EXACTLY (to the rounding noise of G) by a diagonally pivoted Cholesky factorisation
// G = L L^T + R, L n x k, stopped when every remaining diagonal entry is below delta tr (R is positive semi-
// definite with trace <= (n - k) delta tr).  One warp per matrix, rows on the lanes, no block barrier.  The
// eigenvalues of G are those of the k x k matrix H = L^T L, its eigenvectors U = L W Lambda^(-1/2): the Jacobi
// sweeps run on H — a third of the rounds, a quarter of the work per round.
template <int T, int NW, int PPW>
__device__ __noinline__ void eigen_and_select(SlabShared<NW, PPW>& S, double* sm)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const SlabLay& L = S.L;
    const int rK = L.rK;
    const bool prof = S.profOn && tid == 0;
    long long tp = prof ? clock64() : 0;
    // ---- pivoted Cholesky: warp k factorises G_k into W_k (= L), rows lane and lane + 32 in this lane's registers
    if (warp < 3) {
        const int k = warp, n = S.P.n[k], ld = L.LU[k];
        const double* G = sm + L.oG[k];
        double* Lm = sm + L.oW[k];
        const int i0 = lane, i1 = lane + 32;
        double d0 = i0 < n ? G[i0 + ld * i0] : -1.0, d1 = i1 < n ? G[i1 + ld * i1] : -1.0;
        double tr = fmax(d0, 0.0) + fmax(d1, 0.0);
        for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
        const double eps = S.P.eps;
        const double stop = fmax(1e-17, 1e-5 * eps * eps) * tr;
        int kd = 0;
        for (int j = 0; j < n; j++) {
            // pivot = the largest remaining diagonal entry (ties: the lower index)
            double best = d0 >= d1 ? d0 : d1;
            int bi = d0 >= d1 ? i0 : i1;
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) {
                    best = ob;
                    bi = oi;
                }
            }
            if (!(best > stop)) break;
            const double inv = rsqrt(best);
            // column j: (G(:, piv) - L(:, :j) L(piv, :j)^T) / sqrt(d_piv); rows already used as pivots are exactly zero
            double c0 = (i0 < n && d0 >= 0.0) ? G[i0 + ld * bi] : 0.0, c1 = (i1 < n && d1 >= 0.0) ? G[i1 + ld * bi] : 0.0;
            for (int q = 0; q < j; q++) {
                const double lp = Lm[bi + ld * q];
                if (i0 < n) c0 = fma(-Lm[i0 + ld * q], lp, c0);
                if (i1 < n) c1 = fma(-Lm[i1 + ld * q], lp, c1);
            }
            if (i0 == bi) c0 = best;
            if (i1 == bi) c1 = best;
            const double l0 = (i0 < n && d0 >= 0.0) ? c0 * inv : 0.0, l1 = (i1 < n && d1 >= 0.0) ? c1 * inv : 0.0;
            if (i0 < n) Lm[i0 + ld * j] = l0;   // rows >= n are never read (the leading dimension may be below 32)
            if (i1 < n) Lm[i1 + ld * j] = l1;
            if (d0 >= 0.0) d0 = i0 == bi ? -1.0 : fmax(d0 - l0 * l0, 0.0);
            if (d1 >= 0.0) d1 = i1 == bi ? -1.0 : fmax(d1 - l1 * l1, 0.0);
            kd = j + 1;
            __syncwarp();
        }
        if (kd == 0) {   // the zero tensor: one arbitrary direction, as an SVD would return
            if (i0 < n) Lm[i0] = i0 == 0 ? 1.0 : 0.0;
            if (i1 < n) Lm[i1] = 0.0;
            kd = 1;
        }
        if (lane == 0) {
            S.kdim[k] = kd;
            S.trace[k] = tr;
            S.crit[3 * k] = fmax(1e-16, 0.02 * eps * eps) * tr;
            S.crit[3 * k + 1] = 1e-6 * eps * eps * tr;
            S.crit[3 * k + 2] = 0.03 * eps * eps * tr;
            S.invTr[k] = tr > 0.0 ? 1.0 / tr : 0.0;
        }
    }
    __syncthreads();
    // ---- H = L^T L into the place of G (zero beyond k x k), V = I
    for (int e = tid; e < L.oW[0]; e += T) sm[e] = 0.0;
    __syncthreads();
    for (int k = 0; k < 3; k++) {
        const int kd = S.kdim[k], n = S.P.n[k], ld = L.LU[k];
        const double* Lm = sm + L.oW[k];
        double* H = sm + L.oG[k];
        double* V = sm + L.oV[k];
        for (int e = tid; e < kd * kd; e += T) {
            const int a = e % kd, b = e / kd;
            if (a > b) continue;
            double v = 0.0;
            for (int i = 0; i < n; i++) v = fma(Lm[i + ld * a], Lm[i + ld * b], v);
            H[a + ld * b] = v;
            H[b + ld * a] = v;
        }
        for (int i = tid; i < L.p[k]; i += T) V[i + ld * i] = 1.0;
    }
    if (tid == 0)
        for (int k = 0; k < 3; k++) S.jw.n[k] = S.kdim[k];
    __syncthreads();
    if (prof) {
        const long long now = clock64();
        S.prof[5] += now - tp;
        tp = now;
    }
    if (max(S.kdim[0], max(S.kdim[1], S.kdim[2])) <= 24) {
        if (warp < 3) jacobi_warp(S.jw, warp);
        __syncthreads();
    } else {
        jacobi3<T>(S.jw);
    }
    if (prof) {
        const long long now = clock64();
        S.prof[2] += now - tp;
        tp = now;
    }
    for (int q = tid; q < 3 * kMaxN; q += T) {
        const int k = q / kMaxN, j = q % kMaxN;
        if (j < S.kdim[k]) S.lam[k][j] = sm[L.oG[k] + j + L.LU[k] * j];
    }
    __syncthreads();
    // descending order by rank counting (stable)
    for (int q = tid; q < 3 * kMaxN; q += T) {
        const int k = q / kMaxN, j = q % kMaxN, n = S.kdim[k];
        if (j >= n) continue;
        const double* lam = S.lam[k];
        int rank = 0;
        for (int i = 0; i < n; i++)
            if (lam[i] > lam[j] || (lam[i] == lam[j] && i < j)) rank++;
        S.ord[k][rank] = j;
    }
    __syncthreads();
    if (tid < 3) {
        const int k = tid, n = S.kdim[k];
        const double* lam = S.lam[k];
        // sigma_j = sqrt(lambda_j); |sigma|^2 = sum of all lambda_j = trace of G        (tucker.cpp:450)
        const double thr = S.P.eps * sqrt(S.trace[k]) / sqrt(3.0);
        const int cap = min(S.P.maxRank, S.P.rcap[k]);
        int r = 0;
        for (int j = 0; j < n; j++) {
            const double sig = sqrt(fmax(lam[S.ord[k][j]], 0.0));
            if (r == 0 || (sig > thr && r < cap)) r++;   // sorted: a prefix is kept   (tucker.cpp:454-460)
            else break;
        }
        S.rsel[k] = r;
    }
    __syncthreads();
    // ---- factors: U(:, a) = L W(:, ord[a]) / sqrt(lambda_ord[a])
    for (int k = 0; k < 3; k++) {
        double* U = sm + L.oUnew[k];
        const double* V = sm + L.oV[k];
        const double* Lm = sm + L.oW[k];
        const int lu = L.LU[k], nk = S.P.n[k], rk = S.rsel[k], kd = S.kdim[k];
        const int* ord = S.ord[k];
        const double* lam = S.lam[k];
        for (int e = tid; e < lu * rK; e += T) {
            const int i = e % lu, a = e / lu;
            double v = 0.0;
            if (i < nk && a < rk) {
                const int col = ord[a];
                for (int j = 0; j < kd; j++) v = fma(Lm[i + lu * j], V[j + lu * col], v);
                const double la = lam[col];
                v *= la > 0.0 ? rsqrt(la) : 0.0;
            }
            U[e] = v;
        }
    }
    __syncthreads();
    // one Gram-Schmidt pass (twice is enough) per factor, one warp each: the weakest kept columns come out of
    // L W / sqrt(lambda) orthonormal only to ~1e-16 tr / lambda
    if (warp < 3) {
        const int k = warp, nk = S.P.n[k], rk = S.rsel[k], lu = L.LU[k];
        double* U = sm + L.oUnew[k];
        const int i0 = lane, i1 = lane + 32;
        for (int a = 0; a < rk; a++) {
            double u0 = i0 < nk ? U[i0 + lu * a] : 0.0, u1 = i1 < nk ? U[i1 + lu * a] : 0.0;
            for (int rep = 0; rep < 2; rep++)
                for (int b = 0; b < a; b++) {
                    const double w0 = i0 < nk ? U[i0 + lu * b] : 0.0, w1 = i1 < nk ? U[i1 + lu * b] : 0.0;
                    double dot = u0 * w0 + u1 * w1;
                    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                    u0 -= dot * w0;
                    u1 -= dot * w1;
                }
            double nn = u0 * u0 + u1 * u1;
            for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
            const double inv = nn > 0.0 ? rsqrt(nn) : 0.0;
            if (i0 < nk) U[i0 + lu * a] = u0 * inv;
            if (i1 < nk) U[i1 + lu * a] = u1 * inv;
            __syncwarp();
        }
    }
    __syncthreads();
    if (prof) S.prof[3] += clock64() - tp;
}

// ---- pass 3: core = X x_1 U0^T x_2 U1^T x_3 U2^T, slab by slab; then the results of the rounding
template <int T, int NW, int PPW, int RK>
__device__ __noinline__ void pass3(SlabShared<NW, PPW>& S, double* sm, double wall0, double wall1, double wall2, double wall3)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const SlabLay& L = S.L;
    const int n0 = S.P.n[0], n1 = S.P.n[1], n2 = S.P.n[2], M = n0 * n1;
    const int p0 = L.p[0], p1 = L.p[1], LD = L.LD, rK = L.rK, nb1 = p1 / 8;
    constexpr int CPT = (RK * RK * RK + T - 1) / T;
    const int r0 = S.rsel[0], r1 = S.rsel[1], r2 = S.rsel[2], cn = r0 * r1 * r2;
    const int r0B = ceil8(r0) / 8;
    const double* const Xglob = S.Xglob;
    const int slabSz = L.slab;
    const int lu0 = L.LU[0], lu1 = L.LU[1], lu2 = L.LU[2];
    for (int e = tid; e < 4 * slabSz; e += T) sm[e] = 0.0;
    __syncthreads();
    CopyPlan<T> planX;
    planX.build(tid, n0, n1, LD, n0);
    auto issueX = [&](int j, double* buf) { planX.issue(buf, Xglob + (size_t)j * M); };
    issueX(0, sm);
    cp_async_commit();
    if (n2 > 1) issueX(1, sm + slabSz);
    cp_async_commit();
    // W(a, i1, c) = sum over i2 of U2(i2, c) (U0^T S_i2)(a, i1), accumulated in REGISTERS over the slabs: a warp owns an
    // 8 x 8 block (a-block I, i1-block J) of P = U0^T S — with one a-block the sum over i0 is split between two warps —
    // computes it on the tensor cores and multiplies its two entries per lane into the r2 accumulators right away.
    // No shared-memory round trip and no second barrier per slab; the partial sums meet in shared memory at the end.
    const int KS = r0B == 1 ? 2 : 1;                 // warps per block along the contraction
    const int nTask = r0B * nb1 * KS;                // <= 14 for grids up to 48 nodes and 16 warps
    const bool mine = warp < nTask;
    const int tI = mine ? warp % r0B : 0, tJ = mine ? (warp / r0B) % nb1 : 0, tK = mine ? warp / (r0B * nb1) : 0;
    const int K4 = p0 / 4, kBeg = tK * ((K4 + KS - 1) / KS), kEnd = min(K4, kBeg + (K4 + KS - 1) / KS);
    const int g = lane >> 2, c = lane & 3;
    double wacc[2][RK];
#pragma unroll
    for (int q = 0; q < RK; q++) wacc[0][q] = wacc[1][q] = 0.0;
    {
        const double* U0 = sm + L.oUnew[0] + lu0 * (8 * tI + g) + c;
        const double* U2 = sm + L.oUnew[2];
        for (int i2 = 0; i2 < n2; i2++) {
            if (i2 + 2 < n2) issueX(i2 + 2, sm + ((i2 + 2) & 3) * slabSz);
            cp_async_commit();
            cp_async_wait<2>();
            __syncthreads();
            if (!mine) continue;
            const double* Sl = sm + (i2 & 3) * slabSz + c + LD * (8 * tJ + g);
            double pacc[2] = {0.0, 0.0};
            for (int kk = kBeg; kk < kEnd; kk++) dmma884(pacc, U0[4 * kk], Sl[4 * kk]);
#pragma unroll
            for (int q = 0; q < RK; q++) {
                if (q < r2) {
                    const double u = U2[i2 + lu2 * q];
                    wacc[0][q] = fma(pacc[0], u, wacc[0][q]);
                    wacc[1][q] = fma(pacc[1], u, wacc[1][q]);
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // W to shared memory: Ws(a, i1, c) at a + rK (i1 + p1 c); the second half of a split sum is added to the first
    double* const Ws = sm;
    for (int half = 0; half < KS; half++) {
        if (mine && tK == half) {
            const int a = 8 * tI + g;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int i1 = 8 * tJ + 2 * c + u;
#pragma unroll
                for (int q = 0; q < RK; q++) {
                    if (q < r2) {
                        double* dst = Ws + a + rK * (i1 + p1 * q);
                        *dst = (half == 0 ? 0.0 : *dst) + wacc[u][q];
                    }
                }
            }
        }
        __syncthreads();
    }
    // core(a, b, c) = sum over i1 of W(a, i1, c) U1(i1, b)
    double cacc[CPT];
    {
        const double* U1 = sm + L.oUnew[1];
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const int e = tid + T * j;
            if (e < cn) {
                const int a = e % r0, b = (e / r0) % r1, cc = e / (r0 * r1);
                const double* w = Ws + a + rK * p1 * cc;
                const double* u1 = U1 + lu1 * b;
                double v = 0.0;
                for (int i1 = 0; i1 < n1; i1++) v = fma(w[rK * i1], u1[i1], v);
                cacc[j] = v;
            } else {
                cacc[j] = 0.0;
            }
        }
    }
    cp_async_wait<0>();
    // ---- results of the rounding
    if (S.rd < 5) {
        double* coreW = S.coreW;
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const int e = tid + T * j;
            if (e < cn) coreW[e] = cacc[j];
        }
    } else {
        const TuckerParams& P = S.P;
        const int t = S.t;
        double* base = P.out + (size_t)t * P.slot;
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const int e = tid + T * j;
            if (e < cn) base[e] = cacc[j];
        }
        {
            double* UO = base + P.coreCap;
            for (int k = 0; k < 3; k++) {
                const double* U = sm + L.oUnew[k];
                const int nk = P.n[k], lu = L.LU[k], rk = S.rsel[k];
                for (int e = tid; e < nk * rk; e += T) UO[e] = U[(e % nk) + lu * (e / nk)];
                UO += (size_t)nk * P.rcap[k];
            }
        }
        if (tid < 3) P.rout[3 * t + tid] = S.rsel[tid];
        // Density() sums the rounded tensor (particle_data.cpp:93-102): sum = core x_1 (1^T U0) x_2 (1^T U1) x_3 (1^T U2)
        if (tid < 3 * 16) {
            const int k = tid / 16, a = tid % 16;
            double sum = 0.0;
            if (a < S.rsel[k]) {
                const double* U = sm + L.oUnew[k] + L.LU[k] * a;
                for (int i = 0; i < P.n[k]; i++) sum += U[i];
            }
            S.colSum[k][a] = sum;
        }
        __syncthreads();
        double dens = 0.0;
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const int e = tid + T * j;
            if (e < cn) {
                const int a = e % r0, b = (e / r0) % r1, c = e / (r0 * r1);
                dens += cacc[j] * S.colSum[0][a] * S.colSum[1][b] * S.colSum[2][c];
            }
        }
        double vals[5] = {dens, wall0, wall1, wall2, wall3};
#pragma unroll
        for (int q = 0; q < 5; q++) {
            double v = vals[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) S.red[warp][q] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double tot[5] = {0, 0, 0, 0, 0};
            for (int wv = 0; wv < NW; wv++)
#pragma unroll
                for (int q = 0; q < 5; q++) tot[q] += S.red[wv][q];
            P.density[t] = tot[0] * P.cellVolume;
#pragma unroll
            for (int ff = 0; ff < 4; ff++)
                if (S.rec.wallSlot[ff] >= 0) atomicAdd(P.wall + S.rec.wallSlot[ff], P.wallScale * S.rec.area[ff] * tot[1 + ff]);
        }
        // multi-GPU: the new slot of a boundary tet also goes into the ghost rows of the peers
        __syncthreads();
        for (int q = 0; q < 4; q++) {
            const int peer = S.rec.pushPeer[q];
            if (peer < 0) continue;
            const double* src = P.out + (size_t)t * P.slot;
            double* dst = P.peerOut[peer] + (size_t)S.rec.pushRow[q] * P.slot;
            for (size_t e = tid; e < P.slot; e += T) dst[e] = src[e];
            if (tid < 3) P.peerRout[peer][3 * (size_t)S.rec.pushRow[q] + tid] = S.rsel[tid];
        }
    }
    __syncthreads();
}

template <int T, int PPW>
__global__ void __launch_bounds__(T, T >= 512 ? 1 : 2) k_tucker_slab(const TuckerParams Pin)
{
    extern __shared__ __align__(16) double sm[];
    constexpr int NW = T / 32;
    __shared__ SlabShared<NW, PPW> S;
    const int tid = threadIdx.x;
    {   // kernel parameters -> shared memory
        const int* src = reinterpret_cast<const int*>(&Pin);
        int* dst = reinterpret_cast<int*>(&S.P);
        for (int i = tid; i < (int)(sizeof(TuckerParams) / 4); i += T) dst[i] = src[i];
    }
    if (tid < 24) S.prof[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        S.L = slab_layout(S.P.n, S.P.rK);
        const SlabLay& L = S.L;
        JacobiWork& jw = S.jw;
        for (int k = 0; k < 3; k++) {
            jw.G[k] = sm + L.oG[k];
            jw.V[k] = sm + L.oV[k];
            jw.W[k] = sm + L.oW[k];
            jw.n[k] = S.P.n[k];
            jw.p[k] = L.p[k];
            jw.ld[k] = L.LU[k];
        }
        jw.rotC = S.rotC;
        jw.rotS = S.rotS;
        jw.rotPQ = S.rotPQ;
        jw.act = S.act;
        jw.na = S.na;
        jw.mask = S.mask;
        jw.crit = S.crit;
        jw.invTr = S.invTr;
        S.profOn = (S.P.prof && blockIdx.x == 0) ? 1 : 0;
        jw.prof = S.profOn ? S.prof : nullptr;
        double* scr = S.P.scratch + (size_t)blockIdx.x * S.P.scratchPerCTA;
        S.Aglob = scr;
        S.Xglob = scr + S.P.N;
        S.coreW = scr + 2 * (size_t)S.P.N;   // core of the previous rounded right-hand side
    }
    __syncthreads();
    {   // Gram tasks: pass 1 = block pairs of modes 0 and 1, pass 2 = block pairs of mode 2 x two halves of the sum
        const SlabLay& L = S.L;
        const int nb0 = L.p[0] / 8, nb1 = L.p[1] / 8, nb2 = L.p[2] / 8;
        const int np0 = nb0 * (nb0 + 1) / 2, np1 = nb1 * (nb1 + 1) / 2, np2 = nb2 * (nb2 + 1) / 2;
        for (int e = tid; e < NW * PPW; e += T) {
            const int wv = e / PPW, j = e % PPW, q = wv + NW * j;
            int I = 0, J = 0;
            const int m1 = q < np0 ? 0 : (q < np0 + np1 ? 1 : 15);
            if (m1 != 15) pair_of(m1 == 0 ? q : q - np0, I, J);
            S.task1[e] = (unsigned short)(m1 | (I << 4) | (J << 8));
            I = J = 0;
            const int m2 = q < 2 * np2 ? q / np2 : 15;
            if (m2 != 15) pair_of(q % np2, I, J);
            S.task2[e] = (unsigned short)(m2 | (I << 4) | (J << 8));
        }
    }
    __syncthreads();
    const long long tKernel = clock64();
    for (int t = blockIdx.x; t < S.P.nOwned; t += gridDim.x) {
        __syncthreads();
        {
            const int* gsrc = reinterpret_cast<const int*>(S.P.rec + t);
            int* d = reinterpret_cast<int*>(&S.rec);
            for (int i = tid; i < (int)(sizeof(TetRec) / 4); i += T) d[i] = gsrc[i];
        }
        __syncthreads();
        double wall0 = 0.0, wall1 = 0.0, wall2 = 0.0, wall3 = 0.0;
        for (int rd = 0; rd < 6; rd++) {
            if (tid == 0) {
                const TuckerParams& P = S.P;
                const int kind = rd < 4 ? KIND_FACE : (rd == 4 ? KIND_ACCEL : KIND_FINAL);
                const int f = rd < 4 ? rd : 0;
                const int bc = S.rec.bc[f];
                S.t = t;
                S.rd = rd;
                S.kind = kind;
                S.f = f;
                S.pairBC = kind == KIND_FACE && (bc == VT_PBC_NONBOUNDARY || bc == VT_PBC_PERIODIC);
                S.srcBC = kind == KIND_FACE && bc == VT_PBC_SOURCE;
                S.absBC = kind == KIND_FACE && bc == VT_PBC_ABSORBING;
                S.genA = rd == 0;       // the own tensor is expanded here and kept dense in Aglob
                S.useR = rd > 0;
                S.wallOn = S.absBC && S.rec.wallSlot[f] >= 0;
                S.coef = S.rec.coef[f];
                for (int k = 0; k < 3; k++) {
                    S.nrm[k] = S.rec.nrm[f][k];
                    S.gacc[k] = (P.qm * (P.E[3 * (size_t)t + k] + P.ext[k])) * P.inv2h[k];
                }
            }
            __syncthreads();
            const bool prof = S.profOn && tid == 0;
            long long tp = prof ? clock64() : 0;
            stage_operands<T, NW, PPW>(S, sm);
            const double wallSum = pass1<T, NW, PPW>(S, sm);
            if (rd == 0) wall0 += wallSum;
            else if (rd == 1) wall1 += wallSum;
            else if (rd == 2) wall2 += wallSum;
            else if (rd == 3) wall3 += wallSum;
            if (prof) {
                const long long now = clock64();
                S.prof[0] += now - tp;
                tp = now;
            }
            gram_pass01<T, NW, PPW>(S, sm);
            gram_pass2<T, NW, PPW>(S, sm);
            eigen_and_select<T, NW, PPW>(S, sm);
            if (prof) tp = clock64();
            if (S.L.rK <= 8) pass3<T, NW, PPW, 8>(S, sm, wall0, wall1, wall2, wall3);
            else pass3<T, NW, PPW, 16>(S, sm, wall0, wall1, wall2, wall3);
            if (prof) S.prof[4] += clock64() - tp;
        }
    }
    if (S.profOn && tid == 0) {
        S.prof[7] = clock64() - tKernel;
        for (int i = 0; i < 24; i++) S.P.prof[i] = S.prof[i];
    }
}

int rank_cols(const TuckerParams& P)
{
    int cap = 6;   // |v.n| tables
    for (int k = 0; k < 3; k++) cap = std::max(cap, std::min(P.maxRank, P.rcap[k]));
    return (cap + 7) / 8 * 8;
}

}  // namespace

bool slab_eligible(const vt_ctx* ctx, const TuckerParams& P)
{
    const char* force = std::getenv("VT_TUCKER_KERNEL");   // "general": always k_tucker (read per call: tests switch it)
    if (force && std::strcmp(force, "general") == 0) return false;
    const int nmax = std::max({P.n[0], P.n[1], P.n[2]});
    const int nmin = std::min({P.n[0], P.n[1], P.n[2]});
    // up to 32 nodes per axis the general kernel (four 128-thread CTAs per SM) is the faster one: measured 36.9 ms
    // against 52 ms for 3072 tets x 32^3 (profiles/r2_tucker_slab_notes.md)
    static const int minN = std::getenv("VT_TUCKER_SLAB_MIN_N") ? std::atoi(std::getenv("VT_TUCKER_SLAB_MIN_N")) : 33;
    if (nmax > kMaxP || nmax < minN || nmax <= 16 || nmin < 2) return false;
    if (!(P.eps > 0.0) || P.eps * P.eps / 3.0 < 1e-13) return false;   // small-eps refinement lives in k_tucker
    const int rK = rank_cols(P);
    if (rK > 16) return false;
    const SlabLay L = slab_layout(P.n, rK);
    return (size_t)L.total * sizeof(double) + 8 * 1024 <= (size_t)ctx->prop.sharedMemPerBlockOptin;
}

void launch_tucker_slab(vt_ctx* ctx, TuckerState& ts, TuckerParams& P)
{
    if (ctx->nOwned == 0) return;
    P.rK = rank_cols(P);
    const SlabLay L = slab_layout(P.n, P.rK);
    const size_t smem = (size_t)L.total * sizeof(double);
    const int nmax = std::max({P.n[0], P.n[1], P.n[2]});
    const int sms = ctx->prop.multiProcessorCount;
    if (nmax <= 32) {   // experiment (VT_TUCKER_SLAB_MIN_N <= 32): 256 threads, two CTAs per SM
        const int grid = std::min({ctx->nOwned, ts.scratchCTAs, 2 * sms});
        VT_CUDA(cudaFuncSetAttribute(k_tucker_slab<256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_tucker_slab<256, 3><<<grid, 256, smem, ctx->stream>>>(P);
    } else {
        const int grid = std::min({ctx->nOwned, ts.scratchCTAs, sms});
        VT_CUDA(cudaFuncSetAttribute(k_tucker_slab<512, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_tucker_slab<512, 4><<<grid, 512, smem, ctx->stream>>>(P);
    }
    ctx->launches++;
    VT_CUDA(cudaGetLastError());
}

}  // namespace vt
