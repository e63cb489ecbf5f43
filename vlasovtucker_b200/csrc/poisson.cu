// Poisson path on the device — PoissonSolver of the reference (src/poisson.cpp).
//
//   host, once   vt_poisson_setup: per-face coefficients of _FillLineCoeffs (poisson.cpp:126-177),
//                the least-squares distances of _TetLSG (:361-405) and the cross-diffusion
//                vectors of _CorrectRHS (:306-359), flattened to SoA in device order.
//   per solve    k_rhs        -rho/eps0*V + Dirichlet/Neumann terms          (:184-190, :246-274)
//                k_pcg        Jacobi-preconditioned CG, one persistent cooperative kernel
//                             (Eigen ConjugateGradient<Upper, DiagonalPreconditioner>, :39-53)
//                k_gradient   weighted least-squares gradient, E = -grad     (:361-455, :220-229)
//                k_correct    rhs -= cross-diffusion of the previous gradient (:201-205, :306-359)
//
// The matrix is stored ELL-like (diagonal + one coefficient per face): a tet has at most four
// neighbours.  The reference hands Eigen's CG the Upper triangle only (poisson.h:41-44), so the
// coefficient used for the pair (i,j) is the one assembled in row min(i,j) — in the reference's
// tet numbering — and the pinned row 0 decouples; the same choice is made here on the host.
#include "vt_internal.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace cg = cooperative_groups;

namespace vt {

struct PoissonData {
    int n = 0;
    bool solutionIsUnique = false;
    int pinnedRow = -1;            // device row of the reference's tet 0 when no Dirichlet BC
    // per face (4 per row), device order
    int32_t* nbr = nullptr;
    uint8_t* bc = nullptr;
    double* offU = nullptr;        // symmetric (Upper-view) off-diagonal coefficient, 0 where none
    double* coefD = nullptr;       // A/(d.n) of Dirichlet faces (RHS), 0 elsewhere
    double* area = nullptr;
    double* dist = nullptr;        // 3 per face: LSG distance vector
    double* q = nullptr;           // 3 per face: A*(n - e/(e.n)), 0 for Neumann
    double* wOwn = nullptr;        // weight of own gradient in _WeightedGradient
    double* wAdj = nullptr;
    double* bcValue = nullptr;
    double* bcGrad = nullptr;
    // per row
    double* diag = nullptr;
    double* invDiag = nullptr;
    double* volume = nullptr;
    double* rhs = nullptr;
    double* grad = nullptr;        // 3 per row
    double* x = nullptr;
    double* r = nullptr;
    double* z = nullptr;
    double* tmp = nullptr;
    double* p[2] = {nullptr, nullptr};
    double* partial = nullptr;     // per-CTA partial sums (3 per CTA)
    int* status = nullptr;         // [iterations, flag]
    double* statusD = nullptr;     // [residualNorm2, rhsNorm2]
    bool haveGradient = false;
    bool statsPending = false;     // the last solve's status has not been read back yet
    int lastIterations = 0;
    double lastRelResidual = 0;
    // ---- partitioned solve (multi-GPU): rows = owned tets, ghost values of x / z / grad arrive by peer
    // stores, the two dot products of an iteration are summed over the ranks inside the kernel.
    // One allocation holds everything a peer writes into: [x | z | grad | red | flagR | flagB].
    int nTot = 0;                  // owned + ghost rows
    void* xchg = nullptr;
    size_t xchgBytes = 0;
    unsigned long long* ll = nullptr;   // [2][kMaxRanks][4]: the ranks' partial sums as {epoch, payload} words, double buffered
    uint32_t* flagB = nullptr;     // [kMaxRanks]: barrier epoch announced by rank s
    int commRank = 0, commWorld = 1;
    void* peerXchg[64] = {};       // base of every rank's exchange block (own entry = own block)
    int peerNTot[64] = {};
    void* commTable = nullptr;     // device copy of CommTable
    int32_t* pushRank = nullptr;   // [4 * n] device order: rank that holds this row as a ghost, -1 unused
    int32_t* pushRow = nullptr;    // ... and the ghost row there
    uint32_t* epochDev = nullptr;  // reduction epoch counter, lives on the device across solves
    uint32_t barEpoch = 0;
    unsigned long long timeoutNs = 20000000000ULL;   // 20 s; VT_COMM_TIMEOUT_MS
    std::vector<void*> ipcOpened;
    double tol = std::numeric_limits<double>::epsilon();
    int gridBlocks = 0;
    int globalRows = 0;            // rows of the whole system (= n on one GPU)
    std::vector<uint8_t> bcHost;       // caller order, 4 per tet
};

namespace {

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

const double kEps0 = 8.85e-12;  // constants.h:10
constexpr int kMaxRanks = 64;

// what the kernels need to reach the other ranks (device memory, written once at attach time)
struct CommTable {
    double* x[kMaxRanks];
    double* z[kMaxRanks];
    double* grad[kMaxRanks];
    unsigned long long* ll[kMaxRanks];   // [2][kMaxRanks][4] words {epoch : 32, payload : 32}
    uint32_t* flagB[kMaxRanks];
};

// layout of a rank's exchange block for nTot rows
struct XchgLayout {
    size_t x, z, grad, ll, flagB, bytes;
    explicit XchgLayout(size_t nTot)
    {
        x = 0;
        z = x + nTot * sizeof(double);
        grad = z + nTot * sizeof(double);
        ll = grad + 3 * nTot * sizeof(double);
        flagB = ll + 2 * kMaxRanks * 4 * sizeof(unsigned long long);
        bytes = flagB + kMaxRanks * sizeof(uint32_t);
    }
};

struct PcgParams {
    int n;
    const int32_t* nbr;
    const double* offU;
    const double* diag;
    const double* invDiag;
    const double* rhs;
    double* x;
    double* r;
    double* z;
    double* tmp;
    double* p0;
    double* p1;
    double* partial;
    int* status;
    double* statusD;
    double tol;
    int maxIters;
    int useGuess;
    // partitioned solve
    int nTot;                     // owned + ghost rows (x, z, p0, p1 have nTot entries)
    int rank, world;
    const CommTable* comm;
    const int32_t* pushRank;
    const int32_t* pushRow;
    unsigned long long* ll;       // own reduction slots
    uint32_t* epoch;
    unsigned long long timeoutNs; // a rank that does not arrive within this time is given up on (status[1] = 1)
};

__device__ __forceinline__ double block_sum(double v, double* sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; w++) s += sh[w];
    return s;
}

// deterministic grid-wide sum of two values: per-CTA partials, one grid sync, then every CTA
// re-sums all partials in the same fixed order.  Two partial buffers alternate so a buffer is
// rewritten only after a later sync has retired all its readers.
// COMM: the sum then goes over the ranks of a partitioned solve as well — CTA 0 stores this rank's
// sums and the reduction epoch into every other rank's slots (peer stores), every CTA waits for the
// other ranks' epochs in its own slots and adds the values in rank order, so that all CTAs of all
// ranks hold bit-identical results.  Slots are double buffered by epoch parity: a rank can only be
// one reduction ahead of the slowest one.  Peer stores issued before this call by any thread of the
// grid (the z halo) are visible to a peer once it has seen the epoch (fence + grid sync + release).
template <bool COMM>
__device__ __forceinline__ void grid_sum2(cg::grid_group& grid, const PcgParams& P, uint32_t& epoch, double a, double b,
                                          double* partial, int& buf, double* sh, double& outA, double& outB,
                                          bool pushed = false)
{
    const double sa = block_sum(a, sh);
    const double sb = block_sum(b, sh);
    double* pb = partial + (size_t)buf * 2 * gridDim.x;
    buf ^= 1;
    if (threadIdx.x == 0) {
        pb[2 * blockIdx.x] = sa;
        pb[2 * blockIdx.x + 1] = sb;
    }
    if (COMM && pushed) __threadfence_system();   // this thread's z halo stores are ordered before the sums below
    grid.sync();
    double ta = 0.0, tb = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        ta += pb[2 * i];
        tb += pb[2 * i + 1];
    }
    outA = block_sum(ta, sh);
    outB = block_sum(tb, sh);
    if (COMM) {
        // Sum over the ranks, LL style: every 8-byte word a rank stores into a peer carries 32 bits of
        // payload and the 32-bit reduction epoch, so the word itself says when it is valid — no separate
        // flag, no fence between data and flag on the critical path.  The two sums travel as four such
        // words (thread 4q+h of CTA 0 stores word h into rank q).  Every CTA polls its own rank's slots and
        // adds the ranks' values in rank order: all CTAs of all ranks end up with the same bits.  Slots are
        // double buffered by epoch parity (a rank is at most one reduction ahead of the slowest one).
        epoch++;
        const int par = epoch & 1u;
        const int t = threadIdx.x, q = t >> 2, h = t & 3;
        __shared__ uint32_t halves[kMaxRanks][4];
        if (q < P.world && q != P.rank) {
            if (blockIdx.x == 0) {
                const unsigned long long bits = (unsigned long long)__double_as_longlong(h < 2 ? outA : outB);
                const unsigned long long word = ((unsigned long long)epoch << 32) | ((h & 1) ? (bits >> 32) : (bits & 0xffffffffULL));
                unsigned long long* dst = P.comm->ll[q] + ((size_t)par * kMaxRanks + P.rank) * 4 + h;
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
            }
            unsigned long long* in = P.ll + ((size_t)par * kMaxRanks + q) * 4 + h;
            unsigned long long v;
            const unsigned long long t0 = global_ns();
            while (true) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(in) : "memory");
                if ((uint32_t)(v >> 32) == epoch) break;
                // CTA 0 gives up on a rank that does not arrive (a dead peer, or virtual ranks whose kernels
                // were not scheduled side by side): it fills the slot with NaN itself, which releases the
                // other CTAs with the same (poisoned) sum, ends the iteration and is reported by the host
                if (blockIdx.x == 0 && (P.status[1] != 0 || global_ns() - t0 > P.timeoutNs)) {
                    v = ((unsigned long long)epoch << 32) | ((h & 1) ? 0x7ff80000ULL : 0ULL);
                    P.status[1] = 1;
                    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(in), "l"(v) : "memory");
                    break;
                }
            }
            halves[q][h] = (uint32_t)v;
        }
        __syncthreads();
        double ga = 0.0, gb = 0.0;
        for (int s = 0; s < P.world; s++) {
            if (s == P.rank) {
                ga += outA;
                gb += outB;
            } else {
                ga += __longlong_as_double(((long long)halves[s][1] << 32) | halves[s][0]);
                gb += __longlong_as_double(((long long)halves[s][3] << 32) | halves[s][2]);
            }
        }
        outA = ga;
        outB = gb;
        __syncthreads();
    }
}

// Eigen 3.4 conjugate_gradient (ConjugateGradient.h:28-91) with the Jacobi preconditioner
// (BasicPreconditioners.h:66-78), as ONE persistent cooperative kernel: two grid syncs per
// iteration.  The search direction p = z + beta*p is recomputed on the fly for neighbour rows
// inside the SpMV (from z and the previous p, both stable across the phase), which removes the
// third sync a separate p-update would need.
// COMM (partitioned solve): rows [n, nTot) are ghost rows.  x of the ghost rows was pushed by the
// owners after the previous solve; z of the boundary rows is stored into the peers' ghost rows as it
// is computed, right before the reduction that publishes it; p of a ghost row is advanced locally.
template <bool COMM>
__global__ void __launch_bounds__(256) k_pcg(PcgParams P)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[8];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    int buf = 0;
    uint32_t epoch = COMM ? *P.epoch : 0u;
    const int nAll = COMM ? P.nTot : P.n;

    auto push_z = [&](int i, double zi) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int rk = P.pushRank[4 * i + q];
            if (rk >= 0) P.comm->z[rk][P.pushRow[4 * i + q]] = zi;
        }
    };

    if (!P.useGuess) {
        for (int i = tid; i < nAll; i += nth) P.x[i] = 0.0;
        grid.sync();
    }
    // residual = rhs - A x ; rhsNorm2
    double l0 = 0.0, l1 = 0.0;
    for (int i = tid; i < P.n; i += nth) {
        double y = P.diag[i] * P.x[i];
#pragma unroll
        for (int f = 0; f < 4; f++) {
            const int a = P.nbr[4 * i + f];
            const double w = P.offU[4 * i + f];
            if (a >= 0 && w != 0.0) y += w * P.x[a];
        }
        const double ri = P.rhs[i] - y;
        P.r[i] = ri;
        l0 += P.rhs[i] * P.rhs[i];
        l1 += ri * ri;
    }
    double rhsNorm2, residualNorm2;
    grid_sum2<COMM>(grid, P, epoch, l0, l1, P.partial, buf, sh, rhsNorm2, residualNorm2);
    auto finish = [&](int it) {
        if (tid == 0) {
            P.status[0] = it;
            P.statusD[0] = residualNorm2;
            P.statusD[1] = rhsNorm2;
            if (COMM) *P.epoch = epoch;
        }
    };
    if (!(rhsNorm2 > 0.0)) {   // zero right-hand side (Eigen returns x = 0), or a poisoned sum
        for (int i = tid; i < nAll; i += nth) P.x[i] = 0.0;
        residualNorm2 = 0.0;
        finish(0);
        return;
    }
    const double threshold = fmax(P.tol * P.tol * rhsNorm2, 2.2250738585072014e-308);
    if (!(residualNorm2 >= threshold)) {
        finish(0);
        return;
    }
    // z = M^-1 r ; absNew = r.z ; previous direction = 0 so that p = z + 0*p on the first pass
    double* pc = P.p0;
    double* pn = P.p1;
    l0 = 0.0;
    for (int i = tid; i < nAll; i += nth) {
        pc[i] = 0.0;
        if (i >= P.n) continue;
        const double zi = P.invDiag[i] * P.r[i];
        P.z[i] = zi;
        if (COMM) push_z(i, zi);
        l0 += P.r[i] * zi;
    }
    double absNew, dummy;
    grid_sum2<COMM>(grid, P, epoch, l0, 0.0, P.partial, buf, sh, absNew, dummy);
    double beta = 0.0;

    int it = 0;
    while (it < P.maxIters) {
        // p = z + beta p ; tmp = A p ; p.tmp
        l0 = 0.0;
        for (int i = tid; i < nAll; i += nth) {
            const double pi = P.z[i] + beta * pc[i];
            pn[i] = pi;
            if (i >= P.n) continue;   // ghost row: only its search direction is advanced
            double t = P.diag[i] * pi;
#pragma unroll
            for (int f = 0; f < 4; f++) {
                const int a = P.nbr[4 * i + f];
                const double w = P.offU[4 * i + f];
                if (a >= 0 && w != 0.0) t += w * (P.z[a] + beta * pc[a]);
            }
            P.tmp[i] = t;
            l0 += pi * t;
        }
        double pAp;
        grid_sum2<COMM>(grid, P, epoch, l0, 0.0, P.partial, buf, sh, pAp, dummy);
        const double alpha = absNew / pAp;
        // x += alpha p ; r -= alpha tmp ; z = M^-1 r ; |r|^2 ; r.z
        l0 = 0.0;
        l1 = 0.0;
        for (int i = tid; i < P.n; i += nth) {
            P.x[i] += alpha * pn[i];
            const double ri = P.r[i] - alpha * P.tmp[i];
            P.r[i] = ri;
            const double zi = P.invDiag[i] * ri;
            P.z[i] = zi;
            if (COMM) push_z(i, zi);
            l0 += ri * ri;
            l1 += ri * zi;
        }
        double rz;
        grid_sum2<COMM>(grid, P, epoch, l0, l1, P.partial, buf, sh, residualNorm2, rz);
        if (!(residualNorm2 >= threshold)) break;   // converged — or poisoned by a rank that never arrived
        beta = rz / absNew;
        absNew = rz;
        double* t = pc;
        pc = pn;
        pn = t;
        it++;
    }
    finish(it);
}

// partitioned solve: copy k doubles per boundary row into the ghost rows of the ranks that hold it
__global__ void k_push_vals(int n, int k, const double* __restrict__ src, const int32_t* __restrict__ pushRank,
                            const int32_t* __restrict__ pushRow, double* const* __restrict__ peerBase)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int q = 0; q < 4; q++) {
        const int rk = pushRank[4 * i + q];
        if (rk < 0) continue;
        double* dst = peerBase[rk] + (size_t)pushRow[4 * i + q] * k;
        for (int c = 0; c < k; c++) dst[c] = src[(size_t)i * k + c];
    }
}

// partitioned solve: barrier over all ranks on the context stream (after the pushes above)
__global__ void k_comm_barrier(const CommTable* comm, uint32_t* myFlags, int rank, int world, uint32_t epoch,
                               int* status, unsigned long long timeoutNs)
{
    const int q = threadIdx.x;
    if (q >= world || q == rank) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(comm->flagB[q] + rank), "r"(epoch) : "memory");
    const uint32_t* in = myFlags + q;
    uint32_t v;
    const unsigned long long t0 = global_ns();
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(in) : "memory");
        if ((int32_t)(v - epoch) < 0 && (status[1] != 0 || global_ns() - t0 > timeoutNs)) {
            status[1] = 1;   // reported by vt_poisson_stats / vt_sync
            break;
        }
    } while ((int32_t)(v - epoch) < 0);
    __threadfence_system();
}

// rhs_i = (-rho_i/eps0) V_i - sum_Dirichlet A/(d.n) value - sum_Neumann g A ; pinned row -> 0
__global__ void k_rhs(int n, const double* __restrict__ rho, const double* __restrict__ volume,
                      const uint8_t* __restrict__ bc, const double* __restrict__ coefD,
                      const double* __restrict__ area, const double* __restrict__ bcValue,
                      const double* __restrict__ bcGrad, int pinnedRow, double* __restrict__ rhs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == pinnedRow) {
        rhs[i] = 0.0;
        return;
    }
    double r = (-rho[i] / kEps0) * volume[i];
    for (int f = 0; f < 4; f++) {
        const int b = bc[4 * i + f];
        if (b == VT_QBC_DIRICHLET) r -= coefD[4 * i + f] * bcValue[4 * i + f];
        else if (b == VT_QBC_NEUMANN) r -= bcGrad[4 * i + f] * area[4 * i + f];
    }
    rhs[i] = r;
}

// rhs_i -= sum_f A (gbar . (n - e/(e.n)))  with gbar the distance-weighted face gradient
__global__ void k_correct(int n, const int32_t* __restrict__ nbr, const uint8_t* __restrict__ bc,
                          const double* __restrict__ q, const double* __restrict__ wOwn,
                          const double* __restrict__ wAdj, const double* __restrict__ grad, int pinnedRow,
                          double* __restrict__ rhs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == pinnedRow) return;
    double r = rhs[i];
    const double g0 = grad[3 * i], g1 = grad[3 * i + 1], g2 = grad[3 * i + 2];
    for (int f = 0; f < 4; f++) {
        const int b = bc[4 * i + f];
        double cross = 0.0;
        if (b == VT_QBC_NONBOUNDARY || b == VT_QBC_PERIODIC) {
            const int a = nbr[4 * i + f];
            const double wo = wOwn[4 * i + f], wa = wAdj[4 * i + f];
            const double w0 = g0 * wo + grad[3 * a] * wa;
            const double w1 = g1 * wo + grad[3 * a + 1] * wa;
            const double w2 = g2 * wo + grad[3 * a + 2] * wa;
            cross = w0 * q[12 * i + 3 * f] + w1 * q[12 * i + 3 * f + 1] + w2 * q[12 * i + 3 * f + 2];
        } else if (b == VT_QBC_DIRICHLET) {
            cross = g0 * q[12 * i + 3 * f] + g1 * q[12 * i + 3 * f + 1] + g2 * q[12 * i + 3 * f + 2];
        }
        r -= cross;
    }
    rhs[i] = r;
}

// _TetLSG (poisson.cpp:361-442): 3x3 weighted least squares solved with full pivoting
__global__ void k_gradient(int n, const int32_t* __restrict__ nbr, const uint8_t* __restrict__ bc,
                           const double* __restrict__ dist, const double* __restrict__ bcValue,
                           const double* __restrict__ bcGrad, const double* __restrict__ phi,
                           double* __restrict__ grad, double* __restrict__ E)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double val = phi[i];
    double d[4][3], adjVal[4], w[4];
    for (int f = 0; f < 4; f++) {
        for (int k = 0; k < 3; k++) d[f][k] = dist[12 * i + 3 * f + k];
        const double len = sqrt(d[f][0] * d[f][0] + d[f][1] * d[f][1] + d[f][2] * d[f][2]);
        w[f] = 1 / len;
        const int b = bc[4 * i + f];
        if (b == VT_QBC_DIRICHLET) adjVal[f] = bcValue[4 * i + f];
        else if (b == VT_QBC_NEUMANN) adjVal[f] = val + len * bcGrad[4 * i + f];
        else adjVal[f] = phi[nbr[4 * i + f]];
    }
    double a[3][4];
    for (int k = 0; k < 3; k++) {
        for (int c = 0; c < 3; c++) {
            double m = 0;
            for (int j = 0; j < 4; j++) m += 2 * w[j] * d[j][k] * d[j][c];
            a[k][c] = m;
        }
        double r = 0;
        for (int j = 0; j < 4; j++) r -= 2 * w[j] * d[j][k] * (val - adjVal[j]);
        a[k][3] = r;
    }
    int perm[3] = {0, 1, 2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int pr = k, pc = k;
        double best = -1;
        for (int r = k; r < 3; r++)
            for (int c = k; c < 3; c++)
                if (fabs(a[r][c]) > best) {
                    best = fabs(a[r][c]);
                    pr = r;
                    pc = c;
                }
        if (pr != k)
            for (int c = 0; c < 4; c++) {
                double t = a[k][c];
                a[k][c] = a[pr][c];
                a[pr][c] = t;
            }
        if (pc != k) {
            for (int r = 0; r < 3; r++) {
                double t = a[r][k];
                a[r][k] = a[r][pc];
                a[r][pc] = t;
            }
            int t = perm[k];
            perm[k] = perm[pc];
            perm[pc] = t;
        }
        for (int r = k + 1; r < 3; r++) {
            const double l = a[r][k] / a[k][k];
            for (int c = k; c < 4; c++) a[r][c] -= l * a[k][c];
        }
    }
    double y[3];
    for (int r = 2; r >= 0; r--) {
        double s = a[r][3];
        for (int c = r + 1; c < 3; c++) s -= a[r][c] * y[c];
        y[r] = s / a[r][r];
    }
    double g[3];
    for (int k = 0; k < 3; k++) g[perm[k]] = y[k];
    for (int k = 0; k < 3; k++) {
        grad[3 * i + k] = g[k];
        E[3 * i + k] = -g[k];   // poisson.cpp:220-229
    }
}

struct V3 {
    double x, y, z;
};
V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
V3 divs(V3 a, double d) { return {a.x / d, a.y / d, a.z / d}; }
V3 muls(V3 a, double d) { return {a.x * d, a.y * d, a.z * d}; }
double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
double norm(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

template <class T>
T* to_device(const std::vector<T>& v)
{
    T* d = nullptr;
    VT_CUDA(cudaMalloc(&d, std::max<size_t>(1, v.size()) * sizeof(T)));
    VT_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}
template <class T>
void free_dev(T*& p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

void upload_bc_values(vt_ctx* ctx, PoissonData& P, const double* bcValue, const double* bcGrad)
{
    const int n = P.n;
    std::vector<double> v(4 * (size_t)n), g(4 * (size_t)n);
    for (int p = 0; p < n; p++)
        for (int f = 0; f < 4; f++) {
            v[4 * (size_t)p + f] = bcValue[4 * (size_t)ctx->order[p] + f];
            g[4 * (size_t)p + f] = bcGrad[4 * (size_t)ctx->order[p] + f];
        }
    VT_CUDA(cudaMemcpy(P.bcValue, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
    VT_CUDA(cudaMemcpy(P.bcGrad, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice));
}

void read_stats(vt_ctx* ctx, PoissonData& P)
{
    if (!P.statsPending) return;
    int st[2];
    double sd[2];
    VT_CUDA(cudaMemcpyAsync(st, P.status, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
    VT_CUDA(cudaMemcpyAsync(sd, P.statusD, sizeof(sd), cudaMemcpyDeviceToHost, ctx->stream));
    VT_CUDA(cudaStreamSynchronize(ctx->stream));
    P.lastIterations = st[0];
    P.lastRelResidual = sd[1] > 0 ? std::sqrt(sd[0] / sd[1]) : 0.0;
    P.statsPending = false;
    if (st[1] != 0)
        throw std::runtime_error("partitioned Poisson solve: a rank did not reach a reduction or barrier in time "
                                 "(VT_COMM_TIMEOUT_MS); the fields of this step are invalid");
}

// Launches the solve and returns: nothing here waits for the device (a partitioned solve driven by
// one host thread — virtual ranks, several devices — must have every rank's kernel in flight before
// any of them can finish); vt_poisson_stats reads the iteration count back on demand.
void run_pcg(vt_ctx* ctx, PoissonData& P, bool useGuess)
{
    PcgParams pp;
    pp.n = P.n;
    pp.nbr = P.nbr;
    pp.offU = P.offU;
    pp.diag = P.diag;
    pp.invDiag = P.invDiag;
    pp.rhs = P.rhs;
    pp.x = P.x;
    pp.r = P.r;
    pp.z = P.z;
    pp.tmp = P.tmp;
    pp.p0 = P.p[0];
    pp.p1 = P.p[1];
    pp.partial = P.partial;
    pp.status = P.status;
    pp.statusD = P.statusD;
    pp.tol = P.tol;
    pp.maxIters = 2 * P.globalRows;   // IterativeSolverBase default: 2 * rows of the (global) system
    pp.useGuess = useGuess ? 1 : 0;
    pp.nTot = P.nTot;
    pp.rank = P.commRank;
    pp.world = P.commWorld;
    pp.comm = static_cast<const CommTable*>(P.commTable);
    pp.pushRank = P.pushRank;
    pp.pushRow = P.pushRow;
    pp.ll = P.ll;
    pp.epoch = P.epochDev;
    pp.timeoutNs = P.timeoutNs;
    void* args[] = {&pp};
    int blocks = std::min(P.gridBlocks, std::max(1, (P.nTot + 255) / 256));
    const bool comm = P.commWorld > 1;
    if (comm && (!P.commTable || !P.pushRank))
        throw std::runtime_error("partitioned Poisson: vt_poisson_comm_attach / vt_poisson_set_push have not been called");
    VT_CUDA(cudaLaunchCooperativeKernel(comm ? (void*)k_pcg<true> : (void*)k_pcg<false>, dim3(blocks), dim3(256), args, 0,
                                        ctx->stream));
    ctx->launches++;
    P.statsPending = true;
}

// partitioned solve: boundary rows of `src` (k doubles per row) into the peers' ghost rows, then a
// barrier over all ranks
void push_and_barrier(vt_ctx* ctx, PoissonData& P, const double* src, int k, bool grad)
{
    if (P.commWorld <= 1) return;
    const CommTable* ct = static_cast<const CommTable*>(P.commTable);
    double* const* base = grad ? ct->grad : ct->x;   // device addresses of the pointer arrays inside the table
    k_push_vals<<<(P.n + 127) / 128, 128, 0, ctx->stream>>>(P.n, k, src, P.pushRank, P.pushRow, base);
    P.barEpoch++;
    k_comm_barrier<<<1, kMaxRanks, 0, ctx->stream>>>(ct, P.flagB, P.commRank, P.commWorld, P.barEpoch, P.status, P.timeoutNs);
    ctx->launches += 2;
    VT_CUDA(cudaGetLastError());
}

void gradient(vt_ctx* ctx, PoissonData& P)
{
    const int n = P.n;
    push_and_barrier(ctx, P, P.x, 1, false);      // phi of the ghost rows
    k_gradient<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, P.nbr, P.bc, P.dist, P.bcValue, P.bcGrad, P.x, P.grad,
                                                         ctx->E);
    ctx->launches++;
    VT_CUDA(cudaGetLastError());
    push_and_barrier(ctx, P, P.grad, 3, true);    // gradient of the ghost rows, for the next correction
}

}  // namespace

void poisson_destroy(PoissonData* P)
{
    if (!P) return;
    free_dev(P->nbr); free_dev(P->bc); free_dev(P->offU); free_dev(P->coefD); free_dev(P->area);
    free_dev(P->dist); free_dev(P->q); free_dev(P->wOwn); free_dev(P->wAdj); free_dev(P->bcValue);
    free_dev(P->bcGrad); free_dev(P->diag); free_dev(P->invDiag); free_dev(P->volume); free_dev(P->rhs);
    free_dev(P->r); free_dev(P->tmp); free_dev(P->p[0]);
    free_dev(P->p[1]); free_dev(P->partial); free_dev(P->status); free_dev(P->statusD);
    for (void* q : P->ipcOpened) cudaIpcCloseMemHandle(q);
    if (P->xchg) cudaFree(P->xchg);   // x, z, grad, reduction slots and flags live in this block
    free_dev(P->commTable); free_dev(P->pushRank); free_dev(P->pushRow); free_dev(P->epochDev);
    delete P;
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_poisson_setup(vt_ctx* ctx, const double* tetCentroid, const double* faceCentroid, const uint8_t* bcType,
                     const double* bcValue, const double* bcNormalGrad)
{
    if (ctx->group) return vt::group_poisson_setup(ctx, tetCentroid, faceCentroid, bcType, bcValue, bcNormalGrad);
    try {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (ctx->nGhost > 0 && ctx->globalId.empty())
            throw std::runtime_error("vt_poisson_setup on a partition needs vt_mesh_set_ghost_geometry first");
        if (ctx->poisson) poisson_destroy(ctx->poisson);
        ctx->poisson = new PoissonData();
        PoissonData& P = *ctx->poisson;
        const int n = P.n = ctx->nOwned;
        const int nG = ctx->nGhost;
        const int nTot = P.nTot = n + nG;
        P.globalRows = ctx->globalRows > 0 ? ctx->globalRows : n;
        if (const char* t = std::getenv("VT_POISSON_TOL")) P.tol = std::atof(t);
        // geometry of local row t (caller order): owned rows from the arguments and vt_mesh_upload,
        // ghost rows (t >= n) from vt_mesh_set_ghost_geometry
        auto C = [&](int t) {
            const double* c = t < n ? tetCentroid + 3 * (size_t)t : ctx->ghostCentroid.data() + 3 * (size_t)(t - n);
            return V3{c[0], c[1], c[2]};
        };
        auto FC = [&](int t, int j) {
            const double* c = t < n ? faceCentroid + 12 * (size_t)t + 3 * j : ctx->ghostFaceCentroid.data() + 12 * (size_t)(t - n) + 3 * j;
            return V3{c[0], c[1], c[2]};
        };
        auto NR = [&](int t, int j) {
            const double* c = t < n ? ctx->normal.data() + 12 * (size_t)t + 3 * j : ctx->ghostNormal.data() + 12 * (size_t)(t - n) + 3 * j;
            return V3{c[0], c[1], c[2]};
        };
        auto AREA = [&](int t, int j) { return t < n ? ctx->area[4 * (size_t)t + j] : ctx->ghostArea[4 * (size_t)(t - n) + j]; };
        // the reference's tet index of a local row: decides which row of a pair holds the Upper entry
        auto GID = [&](int t) { return ctx->globalId.empty() ? t : ctx->globalId[t]; };
        P.bcHost.assign(bcType, bcType + 4 * (size_t)n);
        P.solutionIsUnique = false;
        for (size_t i = 0; i < 4 * (size_t)n; i++)
            if (bcType[i] == VT_QBC_DIRICHLET) P.solutionIsUnique = true;   // poisson.cpp:87-88
        // on a partition the Dirichlet faces may all belong to other ranks: vt_poisson_set_global_dirichlet
        if (ctx->globalDirichlet >= 0) P.solutionIsUnique = ctx->globalDirichlet != 0;
        // the pinned row is the reference's tet 0 (poisson.cpp:128-134): one rank owns it
        P.pinnedRow = -1;
        if (!P.solutionIsUnique)
            for (int t = 0; t < n; t++)
                if (GID(t) == 0) P.pinnedRow = ctx->inv[t];

        // caller-order neighbour table over owned and ghost rows
        std::vector<int32_t> adj(4 * (size_t)nTot, -1);
        for (int p = 0; p < n; p++)
            for (int j = 0; j < 4; j++) {
                const int a = ctx->nbrHost[4 * (size_t)p + j];
                adj[4 * (size_t)ctx->order[p] + j] = a < 0 ? -1 : (a < n ? ctx->order[a] : a);
            }
        for (int g = 0; g < nG; g++)
            for (int j = 0; j < 4; j++) adj[4 * (size_t)(n + g) + j] = ctx->ghostNbr[4 * (size_t)g + j];
        // d for interior/periodic faces (poisson.cpp:145-162): periodic adds the plane offset
        auto faceDistance = [&](int t, int j, int bc) {
            const int a = adj[4 * (size_t)t + j];
            V3 d = sub(C(a), C(t));
            if (bc == VT_QBC_PERIODIC) {
                int k = 0;
                while (k < 4 && adj[4 * (size_t)a + k] != t) k++;
                if (k == 4) throw std::runtime_error("periodic adjacency is not symmetric");
                d = sub(add(d, FC(t, j)), FC(a, k));
            }
            return d;
        };
        // A/(d.n) of face j of local row t (owned or ghost); the BC kind of a pair face is the same on both sides
        auto coefOf = [&](int t, int j, int bc) { return AREA(t, j) / dot(faceDistance(t, j, bc), NR(t, j)); };
        std::vector<double> coef(4 * (size_t)n, 0.0);    // caller order: this row's A/(d.n)
        for (int t = 0; t < n; t++)
            for (int j = 0; j < 4; j++) {
                const size_t fi = 4 * (size_t)t + j;
                const int bc = bcType[fi];
                if (bc == VT_QBC_NONBOUNDARY || bc == VT_QBC_PERIODIC) {
                    if (adj[fi] < 0) throw std::runtime_error("Poisson: boundary face without a field BC (null adjTets, poisson.cpp:142-147)");
                    coef[fi] = coefOf(t, j, bc);
                } else if (bc == VT_QBC_DIRICHLET) {
                    coef[fi] = ctx->area[fi] / dot(sub(FC(t, j), C(t)), NR(t, j));
                }
            }
        std::vector<int32_t> nbrD(4 * (size_t)n, -1);
        std::vector<uint8_t> bcD(4 * (size_t)n);
        std::vector<double> offU(4 * (size_t)n, 0.0), coefD(4 * (size_t)n, 0.0), areaD(4 * (size_t)n), dist(12 * (size_t)n, 0.0),
            q(12 * (size_t)n, 0.0), wOwn(4 * (size_t)n, 0.0), wAdj(4 * (size_t)n, 0.0), diag(n, 0.0), invDiag(n, 1.0), vol(n);
        const bool pinned = !P.solutionIsUnique;
        for (int p = 0; p < n; p++) {
            const int t = ctx->order[p];
            vol[p] = ctx->volume[t];
            double dg = 0.0;
            for (int j = 0; j < 4; j++) {
                const size_t fi = 4 * (size_t)t + j, fo = 4 * (size_t)p + j;
                const int bc = bcType[fi];
                bcD[fo] = (uint8_t)bc;
                areaD[fo] = ctx->area[fi];
                const V3 nrm = NR(t, j);
                V3 dv{0, 0, 0}, e{0, 0, 0};
                bool hasE = false;
                if (bc == VT_QBC_NONBOUNDARY || bc == VT_QBC_PERIODIC) {
                    const int a = adj[fi];
                    nbrD[fo] = a < n ? ctx->inv[a] : a;
                    dg += -coef[fi];
                    // Upper view: coefficient assembled in the row with the smaller reference index;
                    // the pinned row 0 holds only its diagonal (poisson.cpp:128-134)
                    double w;
                    if (a != t && GID(a) > GID(t)) w = (pinned && GID(t) == 0) ? 0.0 : coef[fi];
                    else if (a != t) {
                        if (pinned && GID(a) == 0) w = 0.0;
                        else {
                            // sum of the neighbour row's entries towards t is split per face: use the
                            // back face matching this one (first k with adj[a][k]==t, duplicates in order)
                            int seen = 0;
                            for (int jj = 0; jj < j; jj++)
                                if (adj[4 * (size_t)t + jj] == a) seen++;
                            int k = -1, cnt = 0;
                            for (int kk = 0; kk < 4; kk++)
                                if (adj[4 * (size_t)a + kk] == t) {
                                    if (cnt == seen) { k = kk; break; }
                                    cnt++;
                                }
                            if (k < 0) throw std::runtime_error("adjacency is not symmetric");
                            w = a < n ? coef[4 * (size_t)a + k] : coefOf(a, k, bc);
                        }
                    } else {
                        w = 0.0;   // self-neighbour: folded into the diagonal below
                        dg += coef[fi];
                    }
                    offU[fo] = w;
                    dv = faceDistance(t, j, bc);
                    e = dv;
                    hasE = true;
                    // _WeightedGradient (poisson.cpp:276-299): note adjD is NOT shifted for periodic faces
                    const V3 dOwn = sub(FC(t, j), C(t));
                    const V3 dAdj = sub(FC(t, j), C(a));
                    wOwn[fo] = norm(dAdj) / (norm(dAdj) + norm(dOwn));
                    wAdj[fo] = norm(dOwn) / (norm(dAdj) + norm(dOwn));
                } else if (bc == VT_QBC_DIRICHLET) {
                    dg += -coef[fi];
                    coefD[fo] = coef[fi];
                    dv = sub(FC(t, j), C(t));
                    e = dv;
                    hasE = true;
                } else {  // Neumann (poisson.cpp:389-398)
                    const V3 d = sub(FC(t, j), C(t));
                    dv = muls(nrm, dot(nrm, d));
                }
                dist[12 * (size_t)p + 3 * j] = dv.x;
                dist[12 * (size_t)p + 3 * j + 1] = dv.y;
                dist[12 * (size_t)p + 3 * j + 2] = dv.z;
                if (hasE) {
                    e = divs(e, norm(e));
                    const V3 qq = muls(sub(nrm, divs(e, dot(e, nrm))), ctx->area[fi]);   // poisson.cpp:328-329
                    q[12 * (size_t)p + 3 * j] = qq.x;
                    q[12 * (size_t)p + 3 * j + 1] = qq.y;
                    q[12 * (size_t)p + 3 * j + 2] = qq.z;
                }
            }
            if (pinned && GID(t) == 0) {
                dg = 1.0;
                for (int j = 0; j < 4; j++) offU[4 * (size_t)p + j] = 0.0;
            }
            diag[p] = dg;
            invDiag[p] = dg != 0.0 ? 1.0 / dg : 1.0;   // DiagonalPreconditioner
        }
        P.nbr = to_device(nbrD);
        P.bc = to_device(bcD);
        P.offU = to_device(offU);
        P.coefD = to_device(coefD);
        P.area = to_device(areaD);
        P.dist = to_device(dist);
        P.q = to_device(q);
        P.wOwn = to_device(wOwn);
        P.wAdj = to_device(wAdj);
        P.diag = to_device(diag);
        P.invDiag = to_device(invDiag);
        P.volume = to_device(vol);
        const size_t nA = std::max(1, n);
        VT_CUDA(cudaMalloc(&P.bcValue, 4 * nA * sizeof(double)));
        VT_CUDA(cudaMalloc(&P.bcGrad, 4 * nA * sizeof(double)));
        upload_bc_values(ctx, P, bcValue, bcNormalGrad);
        const size_t nT = std::max(1, nTot);
        for (double** v : {&P.rhs, &P.r, &P.tmp}) {
            VT_CUDA(cudaMalloc(v, nA * sizeof(double)));
            VT_CUDA(cudaMemset(*v, 0, nA * sizeof(double)));
        }
        for (double** v : {&P.p[0], &P.p[1]}) {
            VT_CUDA(cudaMalloc(v, nT * sizeof(double)));
            VT_CUDA(cudaMemset(*v, 0, nT * sizeof(double)));
        }
        // everything a peer rank writes into lives in one block (one CUDA-IPC handle)
        const XchgLayout lay(nT);
        P.xchgBytes = lay.bytes;
        VT_CUDA(cudaMalloc(&P.xchg, lay.bytes));
        VT_CUDA(cudaMemset(P.xchg, 0, lay.bytes));
        char* xb = static_cast<char*>(P.xchg);
        P.x = reinterpret_cast<double*>(xb + lay.x);
        P.z = reinterpret_cast<double*>(xb + lay.z);
        P.grad = reinterpret_cast<double*>(xb + lay.grad);
        P.ll = reinterpret_cast<unsigned long long*>(xb + lay.ll);
        P.flagB = reinterpret_cast<uint32_t*>(xb + lay.flagB);
        VT_CUDA(cudaMalloc(&P.epochDev, sizeof(uint32_t)));
        VT_CUDA(cudaMemset(P.epochDev, 0, sizeof(uint32_t)));
        int perSm = 0;
        VT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_pcg<true>, 256, 0));
        P.gridBlocks = std::max(1, std::min(perSm, 2) * ctx->prop.multiProcessorCount);
        VT_CUDA(cudaMalloc(&P.partial, 4 * (size_t)P.gridBlocks * sizeof(double)));
        VT_CUDA(cudaMalloc(&P.status, 2 * sizeof(int)));
        VT_CUDA(cudaMemset(P.status, 0, 2 * sizeof(int)));
        if (const char* t = std::getenv("VT_COMM_TIMEOUT_MS")) P.timeoutNs = (unsigned long long)std::atoll(t) * 1000000ULL;
        // ... and every kernel of the solve path has to be resident in the context before the first solve:
        // with lazy module loading the first launch of a kernel loads it, which can wait for running
        // kernels — such as another rank's solve that is itself waiting for this rank
        {
            cudaFuncAttributes fa;
            VT_CUDA(cudaFuncGetAttributes(&fa, k_pcg<true>));
            VT_CUDA(cudaFuncGetAttributes(&fa, k_pcg<false>));
            VT_CUDA(cudaFuncGetAttributes(&fa, k_rhs));
            VT_CUDA(cudaFuncGetAttributes(&fa, k_correct));
            VT_CUDA(cudaFuncGetAttributes(&fa, k_gradient));
            VT_CUDA(cudaFuncGetAttributes(&fa, k_push_vals));
            VT_CUDA(cudaFuncGetAttributes(&fa, k_comm_barrier));
        }
        // nothing may be allocated once solves are in flight: an allocation (device or page-locked host
        // memory) is a device-wide synchronisation point, and a rank whose kernels wait for another rank
        // of the same process would then wait forever
        (void)ctx_pinned(ctx, 4 * nA * sizeof(double));
        VT_CUDA(cudaMalloc(&P.statusD, 2 * sizeof(double)));
        P.haveGradient = false;
        return 0;
    } catch (std::exception& e) {
        vt_set_error(e.what());
        return 1;
    }
}

// ---- partitioned solve: wiring of the ranks ---------------------------------------------------------
struct PoissonIpc {
    cudaIpcMemHandle_t block;
    long long nTot;
    unsigned char pad[56];
};
static_assert(sizeof(PoissonIpc) == 128, "Poisson comm handle is 128 bytes");

static void fill_comm_table(vt_ctx* ctx, PoissonData& P)
{
    CommTable tb;
    std::memset(&tb, 0, sizeof(tb));
    for (int r = 0; r < P.commWorld; r++) {
        char* base = static_cast<char*>(P.peerXchg[r]);
        if (!base) continue;
        const XchgLayout lay((size_t)std::max(1, P.peerNTot[r]));
        tb.x[r] = reinterpret_cast<double*>(base + lay.x);
        tb.z[r] = reinterpret_cast<double*>(base + lay.z);
        tb.grad[r] = reinterpret_cast<double*>(base + lay.grad);
        tb.ll[r] = reinterpret_cast<unsigned long long*>(base + lay.ll);
        tb.flagB[r] = reinterpret_cast<uint32_t*>(base + lay.flagB);
    }
    if (!P.commTable) VT_CUDA(cudaMalloc(&P.commTable, sizeof(CommTable)));
    VT_CUDA(cudaMemcpy(P.commTable, &tb, sizeof(tb), cudaMemcpyHostToDevice));
}

int vt_poisson_set_global_dirichlet(vt_ctx* ctx, int anyDirichlet)
{
    ctx->globalDirichlet = anyDirichlet ? 1 : 0;
    return 0;
}

int vt_poisson_comm_export(vt_ctx* ctx, void* handle)
{
    if (ctx->group) { vt_set_error("vt_poisson_comm_export: not available on a device group"); return 1; };
    try {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (!ctx->poisson) throw std::runtime_error("vt_poisson_setup has not been called");
        PoissonIpc pk;
        std::memset(&pk, 0, sizeof(pk));
        VT_CUDA(cudaIpcGetMemHandle(&pk.block, ctx->poisson->xchg));
        pk.nTot = ctx->poisson->nTot;
        std::memcpy(handle, &pk, sizeof(pk));
        return 0;
    } catch (std::exception& e) {
        vt_set_error(e.what());
        return 1;
    }
}

int vt_poisson_comm_attach(vt_ctx* ctx, int myRank, int world, const void* handles)
{
    if (ctx->group) { vt_set_error("vt_poisson_comm_attach: not available on a device group"); return 1; };
    try {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (!ctx->poisson) throw std::runtime_error("vt_poisson_setup has not been called");
        if (world < 1 || world > kMaxRanks || myRank < 0 || myRank >= world) throw std::invalid_argument("bad rank / world size");
        PoissonData& P = *ctx->poisson;
        P.commRank = myRank;
        P.commWorld = world;
        const PoissonIpc* pk = static_cast<const PoissonIpc*>(handles);
        for (int r = 0; r < world; r++) {
            if (r == myRank) {
                P.peerXchg[r] = P.xchg;
                P.peerNTot[r] = P.nTot;
                continue;
            }
            void* base = nullptr;
            VT_CUDA(cudaIpcOpenMemHandle(&base, pk[r].block, cudaIpcMemLazyEnablePeerAccess));
            P.ipcOpened.push_back(base);
            P.peerXchg[r] = base;
            P.peerNTot[r] = (int)pk[r].nTot;
        }
        fill_comm_table(ctx, P);
        return 0;
    } catch (std::exception& e) {
        vt_set_error(e.what());
        return 1;
    }
}

int vt_poisson_comm_attach_local(vt_ctx* ctx, int myRank, int world, vt_ctx* const* ranks)
{
    if (ctx->group) { vt_set_error("vt_poisson_comm_attach_local: not available on a device group"); return 1; };
    try {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (!ctx->poisson) throw std::runtime_error("vt_poisson_setup has not been called");
        if (world < 1 || world > kMaxRanks || myRank < 0 || myRank >= world) throw std::invalid_argument("bad rank / world size");
        PoissonData& P = *ctx->poisson;
        P.commRank = myRank;
        P.commWorld = world;
        int sameDevice = 0;
        for (int r = 0; r < world; r++) {
            vt_ctx* pc = ranks[r];
            if (!pc || !pc->poisson) throw std::runtime_error("vt_poisson_comm_attach_local: a rank has no Poisson set-up yet");
            if (pc->device == ctx->device) sameDevice++;
            else {
                int can = 0;
                VT_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, pc->device));
                if (!can) throw std::runtime_error("vt_poisson_comm_attach_local: no peer access between the devices");
                cudaError_t e = cudaDeviceEnablePeerAccess(pc->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) VT_CUDA(e);
                (void)cudaGetLastError();
            }
            P.peerXchg[r] = pc->poisson->xchg;
            P.peerNTot[r] = pc->poisson->nTot;
        }
        // ranks that share this device ("virtual ranks") run their persistent solve kernels side by side:
        // each takes its share of the co-resident CTAs, or they would wait for each other forever
        P.gridBlocks = std::max(1, P.gridBlocks / std::max(1, sameDevice));
        fill_comm_table(ctx, P);
        return 0;
    } catch (std::exception& e) {
        vt_set_error(e.what());
        return 1;
    }
}

int vt_poisson_set_push(vt_ctx* ctx, const int32_t* pushRank, const int32_t* pushRow)
{
    if (ctx->group) { vt_set_error("vt_poisson_set_push: not available on a device group"); return 1; };
    try {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (!ctx->poisson) throw std::runtime_error("vt_poisson_setup has not been called");
        PoissonData& P = *ctx->poisson;
        const int n = P.n;
        std::vector<int32_t> pr(4 * (size_t)std::max(1, n), -1), prow(4 * (size_t)std::max(1, n), -1);
        for (int p = 0; p < n; p++) {
            const int t = ctx->order[p];
            int used = 0;
            for (int j = 0; j < 4; j++) {
                const int rk = pushRank[4 * (size_t)t + j];
                if (rk < 0) continue;
                if (rk >= kMaxRanks) throw std::invalid_argument("push rank out of range");
                pr[4 * (size_t)p + used] = rk;
                prow[4 * (size_t)p + used] = pushRow[4 * (size_t)t + j];
                used++;
            }
        }
        free_dev(P.pushRank);
        free_dev(P.pushRow);
        P.pushRank = to_device(pr);
        P.pushRow = to_device(prow);
        return 0;
    } catch (std::exception& e) {
        vt_set_error(e.what());
        return 1;
    }
}

int vt_poisson_update_bc_values(vt_ctx* ctx, const double* bcValue, const double* bcNormalGrad)
{
    if (ctx->group) return vt::group_poisson_update_bc_values(ctx, bcValue, bcNormalGrad);
    try {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (!ctx->poisson) throw std::runtime_error("vt_poisson_setup has not been called");
        VT_CUDA(cudaStreamSynchronize(ctx->stream));
        upload_bc_values(ctx, *ctx->poisson, bcValue, bcNormalGrad);
        return 0;
    } catch (std::exception& e) {
        vt_set_error(e.what());
        return 1;
    }
}

int vt_poisson_solve(vt_ctx* ctx, const double* rho, double* phi, double* E)
{
    if (ctx->group) return vt::group_poisson_solve(ctx, rho, phi, E);
    try {
        VT_CUDA(cudaSetDevice(ctx->device));
        if (!ctx->poisson) throw std::runtime_error("vt_poisson_setup has not been called");
        PoissonData& P = *ctx->poisson;
        const int n = P.n;
        if (n == 0) return 0;
        if (rho) {
            double* pin = ctx_pinned(ctx, (size_t)n * sizeof(double));
            for (int p = 0; p < n; p++) pin[p] = rho[ctx->order[p]];
            VT_CUDA(cudaMemcpyAsync(ctx->rho, pin, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        }
        const int blocks = (n + 255) / 256;
        k_rhs<<<blocks, 256, 0, ctx->stream>>>(n, ctx->rho, P.volume, P.bc, P.coefD, P.area, P.bcValue, P.bcGrad,
                                               P.pinnedRow, P.rhs);
        ctx->launches++;
        if (!P.haveGradient) {
            // first call: solve without correction to get an initial gradient (poisson.cpp:192-199)
            run_pcg(ctx, P, false);
            gradient(ctx, P);
            P.haveGradient = true;
        }
        k_correct<<<blocks, 256, 0, ctx->stream>>>(n, P.nbr, P.bc, P.q, P.wOwn, P.wAdj, P.grad, P.pinnedRow, P.rhs);
        ctx->launches++;
        VT_CUDA(cudaGetLastError());
        run_pcg(ctx, P, true);   // guess = previous solution (poisson.cpp:208), x still holds it
        gradient(ctx, P);
        VT_CUDA(cudaMemcpyAsync(ctx->phi, P.x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        if (phi || E) {
            double* pin = ctx_pinned(ctx, 4 * (size_t)n * sizeof(double));
            VT_CUDA(cudaMemcpyAsync(pin, P.x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            VT_CUDA(cudaMemcpyAsync(pin + n, ctx->E, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            VT_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int p = 0; p < n; p++) {
                const int t = ctx->order[p];
                if (phi) phi[t] = pin[p];
                if (E)
                    for (int k = 0; k < 3; k++) E[3 * (size_t)t + k] = pin[n + 3 * (size_t)p + k];
            }
        }
        return 0;
    } catch (std::exception& e) {
        vt_set_error(e.what());
        return 1;
    }
}

int vt_poisson_stats(vt_ctx* ctx, int* lastIterations, double* lastRelResidual)
{
    if (ctx->group) return vt::group_poisson_stats(ctx, lastIterations, lastRelResidual);
    if (!ctx->poisson) return 1;
    try {
        VT_CUDA(cudaSetDevice(ctx->device));
        read_stats(ctx, *ctx->poisson);
    } catch (std::exception& e) {
        vt_set_error(e.what());
        return 1;
    }
    if (lastIterations) *lastIterations = ctx->poisson->lastIterations;
    if (lastRelResidual) *lastRelResidual = ctx->poisson->lastRelResidual;
    return 0;
}
}
