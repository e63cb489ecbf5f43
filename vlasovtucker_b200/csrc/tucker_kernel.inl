// Device code of the general Tucker kernel k_tucker (see tucker.cu for the formulation): included by tucker_inst.cu,
// which is compiled once per instantiation — the three together take eight minutes of cicc in one translation unit.
namespace {


struct Dims {
    int n[3];
    int N;
};

// dst (dims with dim[mode] -> rows) = src x_mode M.  M is addressed M[out + ldm*in] (transpose=false,
// an "rows x n_mode" matrix stored column-major) or M[in + ldm*out] (transpose=true: apply U^T).
// Ms: shared staging for M (>= rowsOut * din[mode] doubles).  A thread takes one fibre of the
// contracted mode and a block of kQB output rows: every element read from src feeds kQB FMAs whose
// matrix operands are shared-memory broadcasts; consecutive threads take consecutive fibres, which
// are consecutive i0 (coalesced) for modes 1 and 2.
__device__ void mode_apply(const double* __restrict__ src, double* __restrict__ dst, const int din[3], int mode,
                           const double* __restrict__ M, int ldm, int rowsOut, bool transpose, double* __restrict__ Ms)
{
    const int K = din[mode];
    const int QS = (rowsOut + kQB - 1) / kQB * kQB;   // padded row count: Ms[k * QS + q]
    for (int e = threadIdx.x; e < QS * K; e += blockDim.x) {
        const int q = e % QS, k = e / QS;
        Ms[e] = q < rowsOut ? (transpose ? M[k + ldm * q] : M[q + ldm * k]) : 0.0;
    }
    __syncthreads();
    const int nFib = din[0] * din[1] * din[2] / K;
    const int nQB = QS / kQB;
    const int strideIn = mode == 0 ? 1 : (mode == 1 ? din[0] : din[0] * din[1]);
    const int strideOut = mode == 0 ? 1 : strideIn;
    for (int w = threadIdx.x; w < nFib * nQB; w += blockDim.x) {
        const int fib = w % nFib, q0 = (w / nFib) * kQB;
        int baseIn, baseOut;
        if (mode == 0) {
            baseIn = fib * K;
            baseOut = fib * rowsOut;
        } else if (mode == 1) {
            const int i0 = fib % din[0], i2 = fib / din[0];
            baseIn = i0 + din[0] * K * i2;
            baseOut = i0 + din[0] * rowsOut * i2;
        } else {
            baseIn = fib;
            baseOut = fib;
        }
        double acc[kQB];
#pragma unroll
        for (int b = 0; b < kQB; b++) acc[b] = 0.0;
        const double* mrow = Ms + q0;
        for (int k = 0; k < K; k++) {
            const double x = src[baseIn + k * strideIn];
#pragma unroll
            for (int b = 0; b < kQB; b++) acc[b] = fma(mrow[k * QS + b], x, acc[b]);
        }
#pragma unroll
        for (int b = 0; b < kQB; b++)
            if (q0 + b < rowsOut) dst[baseOut + (q0 + b) * strideOut] = acc[b];
    }
    __syncthreads();
}

// ---- Gram matrices of the three unfoldings, G_k = X_(k) X_(k)^T, accumulated from shared-memory
// tiles.  A group is one row i against a block of kGB consecutive columns j0..j0+kGB-1 (entries with
// j < i are computed and dropped); a thread owns a fixed set of groups (table in shared memory) and
// keeps their sums in registers across tiles: one operand a_i feeds kGB FMAs, 1.25 shared loads per
// FMA instead of 2.  Consecutive threads take consecutive rows of the same column block, so the a_i
// loads of a warp are consecutive words and the b_j loads are broadcasts.  G is written symmetric
// with leading dimension ld = n | 1 (odd, so that both row and column walks are bank-conflict free
// in the eigen-solver).
constexpr int kGB = 4;
// number of table entries that cover an n x n Gram matrix when the table was laid out for nmax:
// column block jb holds the rows 0 .. min(kGB*jb + kGB, nmax) - 1
__host__ __device__ constexpr int gram_groups(int n, int nmax)
{
    int g = 0;
    for (int jb = 0; jb * kGB < n; jb++) g += (kGB * jb + kGB < nmax) ? kGB * jb + kGB : nmax;
    return g;
}
// groups per thread for a kernel instance serving grids of up to NM nodes per axis with T threads
template <int T, int NM> struct GroupsPerThread { static constexpr int value = (gram_groups(NM, NM) + T - 1) / T; };

struct GramWork {
    double* tile;              // two buffers of tileCap doubles each
    int tileCap;
    const unsigned short* groups;   // group q -> i | (j0 << 8); enumerated block by block, rows
                                    // consecutive; laid out for nmax (rows >= n of a smaller matrix are skipped)
    int nmax;
};

__device__ void build_groups(unsigned short* groups, int nmax)
{
    for (int jb = threadIdx.x; jb * kGB < nmax; jb += blockDim.x) {
        const int base = gram_groups(jb * kGB, nmax);
        const int rows = min(kGB * jb + kGB, nmax);
        for (int i = 0; i < rows; i++) groups[base + i] = (unsigned short)(i | ((jb * kGB) << 8));
    }
    __syncthreads();
}

// 8-byte asynchronous copy global -> shared (the padded tile layouts are only 8-byte aligned)
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

// The tile is double buffered (two halves of tileCap doubles): while the groups of one slab/chunk
// are accumulated from shared memory, cp.async brings the next one in.
template <int T, int NM>
__device__ void gram_mode(const double* __restrict__ X, const int d[3], int mode, double* __restrict__ G, const GramWork& gw)
{
    constexpr int GPT = GroupsPerThread<T, NM>::value;
    const int n = d[mode];
    const int ng = gram_groups(n, gw.nmax);
    double acc[GPT][kGB];
    int gi[GPT], gj[GPT], gc[GPT];   // row, first column, columns in the group (0: no group)
#pragma unroll
    for (int k = 0; k < GPT; k++) {
#pragma unroll
        for (int u = 0; u < kGB; u++) acc[k][u] = 0.0;
        const int q = threadIdx.x + k * T;
        const unsigned short e = q < ng ? gw.groups[q] : 0;
        gi[k] = e & 255;
        gj[k] = e >> 8;
        gc[k] = (q < ng && gi[k] < n) ? min(kGB, n - gj[k]) : 0;
    }
    const int n0 = d[0], n1 = d[1], n2 = d[2], M = n0 * n1;
    // accumulate one tile: element (row r of operand i) sits at tile[aStep * r + aOff(i)]
    auto accumulate = [&](const double* tile, int len, int rowStride, int colStride) {
#pragma unroll
        for (int k = 0; k < GPT; k++) {
            if (gc[k] == 0) continue;
            const double* a = tile + colStride * gi[k];
            int off[kGB];   // columns past the end of a short block repeat its last column (their sums are dropped)
#pragma unroll
            for (int u = 0; u < kGB; u++) off[u] = colStride * (gj[k] + min(u, gc[k] - 1));
            for (int r = 0; r < len; r++) {
                const double av = a[rowStride * r];
#pragma unroll
                for (int u = 0; u < kGB; u++) acc[k][u] = fma(av, tile[rowStride * r + off[u]], acc[k][u]);
            }
        }
    };
    if (mode == 0 || mode == 1) {
        // slabs A = X(:, :, i2), n0 x n1; mode 0: G += A A^T, mode 1: G += A^T A
        const int ld = mode == 0 ? n0 : (n0 | 1);
        auto issue = [&](int i2) {
            const double* slab = X + (size_t)i2 * M;
            double* t = gw.tile + (size_t)(i2 & 1) * gw.tileCap;
            for (int e = threadIdx.x; e < M; e += T) cp_async8(t + (e % n0) + ld * (e / n0), slab + e);
            cp_async_commit();
        };
        issue(0);
        for (int i2 = 0; i2 < n2; i2++) {
            if (i2 + 1 < n2) {
                issue(i2 + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const double* tile = gw.tile + (size_t)(i2 & 1) * gw.tileCap;
            if (mode == 0) accumulate(tile, n1, ld, 1);   // operand i = row i of A: stride ld along the sum
            else accumulate(tile, n0, 1, ld);             // operand i = column i of A
            __syncthreads();
        }
    } else {
        // X as an M x n2 matrix B (column a = plane a); G = B^T B over row chunks
        const int ldmax = gw.tileCap / n2;                       // the chunk (ld x n2) has to fit half the tile
        const int rows = min(M, (ldmax & 1) ? ldmax : ldmax - 1);
        const int ld = rows | 1;
        const int nChunks = (M + rows - 1) / rows;
        auto issue = [&](int c) {
            const int r0 = c * rows, nr = min(rows, M - r0);
            double* t = gw.tile + (size_t)(c & 1) * gw.tileCap;
            for (int e = threadIdx.x; e < nr * n2; e += T) {
                const int r = e % nr, a = e / nr;
                cp_async8(t + r + ld * a, X + (size_t)r0 + r + (size_t)M * a);
            }
            cp_async_commit();
        };
        issue(0);
        for (int c = 0; c < nChunks; c++) {
            if (c + 1 < nChunks) {
                issue(c + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            accumulate(gw.tile + (size_t)(c & 1) * gw.tileCap, min(rows, M - c * rows), 1, ld);
            __syncthreads();
        }
    }
    const int ldg = n | 1;
#pragma unroll
    for (int k = 0; k < GPT; k++) {
#pragma unroll
        for (int u = 0; u < kGB; u++) {
            if (u >= gc[k] || gj[k] + u < gi[k]) continue;   // upper triangle (j >= i) only
            G[gi[k] + ldg * (gj[k] + u)] = acc[k][u];
            G[(gj[k] + u) + ldg * gi[k]] = acc[k][u];
        }
    }
    __syncthreads();
}

// ---- the same three Gram matrices on the FP64 tensor cores: mma.sync.aligned.m8n8k4.f64 (SASS DMMA),
// 256 FMA per warp instruction.  G = T T^T over shared-memory tiles T: for an 8-row block B and four
// consecutive contraction indices k0..k0+3 a lane l holds the "fragment" T[8B + l/4][k0 + l%4] — as the A
// operand for block row I and, unchanged, as the B operand for block column J — so one shared load per
// lane feeds a whole 8x8x4 product, against 1.25 loads per FMA in the DFMA version above.  Tiles are
// stored with a leading dimension == 4 (mod 16), which makes both fragment walks (along and across the
// leading dimension) bank-conflict free, and are zero padded to multiples of 8 / 4.  The upper
// triangle of 8x8 blocks is dealt out to the warps (a block pair stays with one warp for the whole
// accumulation, so nothing is reduced across warps): slabs X(:,:,i2) feed modes 0 and 1 at once (half
// of the warps each), row chunks of the M x n2 matrix feed mode 2.
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}
// 16-byte asynchronous copy global -> shared
__device__ __forceinline__ void cp_async16(double* dst, const double* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}


template <int T, int NM>
struct DmmaPairs {
    static constexpr int nb = (NM + 7) / 8;
    static constexpr int pairs = nb * (nb + 1) / 2;
    static constexpr int warps = T / 32;
    static constexpr int perHalf = (pairs + warps / 2 - 1) / (warps / 2);   // slab pass: half of the warps per mode
    static constexpr int perAll = (pairs + warps - 1) / warps;              // chunk pass: all warps on mode 2
};

// Block pair (I, J), I <= J, has the index p = I + J(J+1)/2 and belongs to warp p % NGW of its group,
// where it is accumulator p / NGW: all of this is resolved at compile time inside fully unrolled
// loops, so fragments and accumulators stay in registers (the only run-time test is the warp-uniform
// "is this pair mine").
// Accumulate `nk` contraction steps of 4 from one tile.  alongLd == false: the Gram index runs along
// the contiguous direction of the tile (fragment T[(8B+g) + ld*(k0+c)]); true: across it
// (fragment T[(k0+c) + ld*(8B+g)]).
template <int NB, int NGW, int PP>
__device__ __forceinline__ void dmma_tile(const double* tile, int ld, bool alongLd, int nk, int nbUsed, int wm,
                                          double (&acc)[PP][2])
{
    const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const int rowStride = alongLd ? ld : 1, kStride = alongLd ? 1 : ld;
    const double* base = tile + g * rowStride + c * kStride;
    for (int k = 0; k < nk; k++) {
        double frag[NB];
#pragma unroll
        for (int b = 0; b < NB; b++) frag[b] = b < nbUsed ? base[(8 * b) * rowStride + (4 * k) * kStride] : 0.0;
#pragma unroll
        for (int J = 0; J < NB; J++) {
#pragma unroll
            for (int I = 0; I <= J; I++) {
                constexpr int dummy = 0;
                (void)dummy;
                const int p = I + J * (J + 1) / 2;
                if (p % NGW == wm && J < nbUsed) dmma884(acc[p / NGW], frag[I], frag[J]);
            }
        }
    }
}

template <int NB, int NGW, int PP>
__device__ __forceinline__ void dmma_store(double* G, int n, int nbUsed, int wm, const double (&acc)[PP][2])
{
    const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const int ldg = n | 1;
#pragma unroll
    for (int J = 0; J < NB; J++) {
#pragma unroll
        for (int I = 0; I <= J; I++) {
            const int p = I + J * (J + 1) / 2;
            if (p % NGW != wm || J >= nbUsed) continue;
            const int i = 8 * I + g;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int j = 8 * J + 2 * c + u;
                if (i < n && j < n) {
                    G[i + ldg * j] = acc[p / NGW][u];
                    G[j + ldg * i] = acc[p / NGW][u];   // diagonal blocks write both triangles from the exact same sums
                }
            }
        }
    }
}

template <int T, int NM>
__device__ void gram_all_dmma(const double* __restrict__ X, const int d[3], double* const G[3], const GramWork& gw)
{
    using DP = DmmaPairs<T, NM>;
    constexpr int NB = DP::nb;
    const int warp = threadIdx.x >> 5;
    constexpr int W = T / 32, HW = W / 2;
    const int n0 = d[0], n1 = d[1], n2 = d[2], M = n0 * n1;
    double* tile0 = gw.tile;
    double* tile1 = gw.tile + gw.tileCap;
    auto zero_tiles = [&]() {
        for (int e = threadIdx.x; e < 2 * gw.tileCap; e += T) gw.tile[e] = 0.0;
        __syncthreads();
    };
    // ---- pass A: slabs, modes 0 and 1
    {
        const int mode = warp < HW ? 0 : 1;
        const int wm = warp < HW ? warp : warp - HW;
        const int n = d[mode], nb = (n + 7) / 8;
        double acc[DP::perHalf][2];
#pragma unroll
        for (int q = 0; q < DP::perHalf; q++) acc[q][0] = acc[q][1] = 0.0;
        const int ld = pad_ld(n0);
        zero_tiles();
        const bool vec2 = (n0 % 2 == 0) && ((reinterpret_cast<size_t>(X) & 15) == 0);
        auto issue = [&](int i2) {
            const double* slab = X + (size_t)i2 * M;
            double* t = (i2 & 1) ? tile1 : tile0;
            if (vec2) {
                const int h0 = n0 / 2;
                for (int e = threadIdx.x; e < M / 2; e += T) cp_async16(t + 2 * (e % h0) + ld * (e / h0), slab + 2 * e);
            } else {
                for (int e = threadIdx.x; e < M; e += T) cp_async8(t + (e % n0) + ld * (e / n0), slab + e);
            }
            cp_async_commit();
        };
        const int nk = mode == 0 ? (n1 + 3) / 4 : (n0 + 3) / 4;
        issue(0);
        for (int i2 = 0; i2 < n2; i2++) {
            if (i2 + 1 < n2) {
                issue(i2 + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            dmma_tile<NB, HW, DP::perHalf>((i2 & 1) ? tile1 : tile0, ld, mode == 1, nk, nb, wm, acc);
            __syncthreads();
        }
        dmma_store<NB, HW, DP::perHalf>(G[mode], n, nb, wm, acc);
    }
    // ---- pass B: X as an M x n2 matrix, row chunks, mode 2
    {
        const int n = n2, nb = (n + 7) / 8;
        double acc[DP::perAll][2];
#pragma unroll
        for (int q = 0; q < DP::perAll; q++) acc[q][0] = acc[q][1] = 0.0;
        // chunk of `rows` consecutive rows: tile[r + ld * a], ld == 4 (mod 16), 8*nb columns
        int ld = gw.tileCap / (8 * nb);
        ld -= ((ld % 16) - 4 + 16) % 16;
        const int rows = ld / 4 * 4;
        const int nChunks = (M + rows - 1) / rows;
        zero_tiles();
        const bool vec2 = (M % 2 == 0) && (rows % 2 == 0) && ((reinterpret_cast<size_t>(X) & 15) == 0);
        auto issue = [&](int cidx) {
            const int r0 = cidx * rows, nr = min(rows, M - r0);
            double* t = (cidx & 1) ? tile1 : tile0;
            if (nr < rows) {   // the last chunk is short: clear what the chunk before it left behind
                for (int e = threadIdx.x; e < (rows - nr) * n2; e += T) t[nr + e % (rows - nr) + ld * (e / (rows - nr))] = 0.0;
            }
            if (vec2 && nr % 2 == 0) {
                const int h = nr / 2;
                for (int e = threadIdx.x; e < h * n2; e += T) {
                    const int r = 2 * (e % h), a = e / h;
                    cp_async16(t + r + ld * a, X + (size_t)r0 + r + (size_t)M * a);
                }
            } else {
                for (int e = threadIdx.x; e < nr * n2; e += T) {
                    const int r = e % nr, a = e / nr;
                    cp_async8(t + r + ld * a, X + (size_t)r0 + r + (size_t)M * a);
                }
            }
            cp_async_commit();
        };
        issue(0);
        for (int cidx = 0; cidx < nChunks; cidx++) {
            if (cidx + 1 < nChunks) {
                issue(cidx + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const int nr = min(rows, M - cidx * rows);
            dmma_tile<NB, W, DP::perAll>((cidx & 1) ? tile1 : tile0, ld, true, (nr + 3) / 4, nb, warp, acc);
            __syncthreads();
        }
        dmma_store<NB, W, DP::perAll>(G[2], n, nb, warp, acc);
    }
    __syncthreads();
}

__device__ __forceinline__ double wsum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Symmetric eigen-decomposition by one warp, in place: Householder tridiagonalisation followed by
// the implicit QL iteration (the EISPACK tred2/tql2 pair in the form popularised by JAMA), inner
// loops spread over the 32 lanes.  V (n x n, leading dimension ld) holds the matrix on entry and
// the eigenvectors (columns) on exit; dv the eigenvalues (unordered), ev is work space.
#define VV(i, j) V[(i) + ld * (j)]
__device__ void eig_sym_warp(double* V, int n, int ld, double* dv, double* ev)
{
    const int lane = threadIdx.x & 31;
    for (int j = lane; j < n; j += 32) dv[j] = VV(n - 1, j);
    __syncwarp();
    for (int i = n - 1; i > 0; i--) {
        double part = 0.0;
        for (int k = lane; k < i; k += 32) part += fabs(dv[k]);
        const double scale = wsum(part);
        double h = 0.0;
        if (scale == 0.0) {
            if (lane == 0) ev[i] = dv[i - 1];
            __syncwarp();
            for (int j = lane; j < i; j += 32) {
                dv[j] = VV(i - 1, j);
                VV(i, j) = 0.0;
                VV(j, i) = 0.0;
            }
        } else {
            part = 0.0;
            for (int k = lane; k < i; k += 32) {
                const double x = dv[k] / scale;   // scale may be subnormal: no reciprocal here
                dv[k] = x;
                part += x * x;
            }
            h = wsum(part);
            __syncwarp();
            const double f = dv[i - 1];
            double g = sqrt(h);
            if (f > 0) g = -g;
            h -= f * g;
            __syncwarp();
            if (lane == 0) {
                ev[i] = scale * g;
                dv[i - 1] = f - g;
            }
            __syncwarp();
            // e = A d on the leading i x i block (lower triangle valid), and store the Householder vector
            for (int j = lane; j < i; j += 32) {
                double s = 0.0;
                for (int k = 0; k <= j; k++) s = fma(VV(j, k), dv[k], s);
                for (int k = j + 1; k < i; k++) s = fma(VV(k, j), dv[k], s);
                ev[j] = s;
            }
            __syncwarp();
            for (int j = lane; j < i; j += 32) VV(j, i) = dv[j];
            part = 0.0;
            const double hinv = 1.0 / h;
            for (int j = lane; j < i; j += 32) {
                const double x = ev[j] * hinv;
                ev[j] = x;
                part += x * dv[j];
            }
            const double hh = wsum(part) / (h + h);
            __syncwarp();
            for (int j = lane; j < i; j += 32) ev[j] -= hh * dv[j];
            __syncwarp();
            // rank-2 update of the lower triangle: row k, columns j <= k
            for (int k = lane; k < i; k += 32) {
                const double ek = ev[k], dk = dv[k];
                for (int j = 0; j <= k; j++) VV(k, j) -= dv[j] * ek + ev[j] * dk;
            }
            __syncwarp();
            for (int j = lane; j < i; j += 32) {
                dv[j] = VV(i - 1, j);
                VV(i, j) = 0.0;
            }
        }
        __syncwarp();
        if (lane == 0) dv[i] = h;
        __syncwarp();
    }
    // accumulate the transformations
    for (int i = 0; i < n - 1; i++) {
        if (lane == 0) {
            VV(n - 1, i) = VV(i, i);
            VV(i, i) = 1.0;
        }
        __syncwarp();
        const double h = dv[i + 1];
        if (h != 0.0) {
            const double hinv = 1.0 / h;
            for (int k = lane; k <= i; k += 32) dv[k] = VV(k, i + 1) * hinv;
            __syncwarp();
            for (int j = lane; j <= i; j += 32) {
                double g = 0.0;
                for (int k = 0; k <= i; k++) g = fma(VV(k, i + 1), VV(k, j), g);
                for (int k = 0; k <= i; k++) VV(k, j) -= g * dv[k];
            }
        }
        __syncwarp();
        for (int k = lane; k <= i; k += 32) VV(k, i + 1) = 0.0;
        __syncwarp();
    }
    for (int j = lane; j < n; j += 32) {
        dv[j] = VV(n - 1, j);
        VV(n - 1, j) = 0.0;
    }
    __syncwarp();
    if (lane == 0) {
        VV(n - 1, n - 1) = 1.0;
        ev[0] = 0.0;
    }
    __syncwarp();
    // implicit QL
    {
        const double t0 = (1 + lane < n) ? ev[1 + lane] : 0.0, t1 = (33 + lane < n) ? ev[33 + lane] : 0.0;
        __syncwarp();
        if (1 + lane < n) ev[lane] = t0;
        if (33 + lane < n) ev[32 + lane] = t1;
    }
    __syncwarp();
    if (lane == 0) ev[n - 1] = 0.0;
    __syncwarp();
    double f = 0.0, tst1 = 0.0;
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n; l++) {
        tst1 = fmax(tst1, fabs(dv[l]) + fabs(ev[l]));
        int m = l;
        while (m < n) {
            if (fabs(ev[m]) <= eps * tst1) break;
            m++;
        }
        if (m > l) {
            int iter = 0;
            do {
                iter++;
                double g = dv[l];
                double p = (dv[l + 1] - g) / (2.0 * ev[l]);
                double r = sqrt(fma(p, p, 1.0));
                if (p < 0) r = -r;
                const double el = ev[l];
                const double dl = el / (p + r), dl1 = el * (p + r);
                double h = g - dl;
                __syncwarp();
                if (lane == 0) {
                    dv[l] = dl;
                    dv[l + 1] = dl1;
                }
                for (int i = l + 2 + lane; i < n; i += 32) dv[i] -= h;
                __syncwarp();
                f += h;
                p = dv[m];
                double c = 1.0, c2 = 1.0, c3 = 1.0, s = 0.0, s2 = 0.0;
                const double el1 = ev[l + 1];
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2;
                    c2 = c;
                    s2 = s;
                    const double ei = ev[i], di = dv[i];
                    g = c * ei;
                    h = c * p;
                    // r = hypot(p, e_i), s = e_i / r, c = p / r.  Scaling by a power of two (exact) keeps the
                    // squares of the subnormal noise entries of a rank-deficient Gram matrix from
                    // underflowing; one reciprocal square root replaces the divisions.
                    const double q2plain = fma(p, p, ei * ei);
                    if (q2plain > 1e-280 && q2plain < 1e280) {   // the usual case: no scaling needed
                        const double qinv = rsqrt(q2plain);
                        r = q2plain * qinv;
                        s = ei * qinv;
                        c = p * qinv;
                    } else {
                        const double mx = fmax(fabs(p), fabs(ei));
                        if (mx == 0.0) {
                            r = 0.0;
                            s = 0.0;
                            c = 1.0;
                        } else {
                            int ex;
                            (void)frexp(mx, &ex);
                            const double a = scalbn(p, -ex), b = scalbn(ei, -ex);
                            const double q2 = fma(a, a, b * b);      // in [0.25, 2)
                            const double qinv = rsqrt(q2);
                            r = scalbn(q2 * qinv, ex);
                            s = b * qinv;
                            c = a * qinv;
                        }
                    }
                    p = c * di - s * g;
                    __syncwarp();
                    if (lane == 0) {
                        // e[i+1] = s_prev * r, d[i+1] = h + s (c g + s d[i])
                        ev[i + 1] = s2 * r;
                        dv[i + 1] = h + s * (c * g + s * di);
                    }
                    for (int k = lane; k < n; k += 32) {
                        const double vh = VV(k, i + 1), vl = VV(k, i);
                        VV(k, i + 1) = s * vl + c * vh;
                        VV(k, i) = c * vl - s * vh;
                    }
                    __syncwarp();
                }
                p = -s * s2 * c3 * el1 * ev[l] / dl1;
                __syncwarp();
                if (lane == 0) {
                    ev[l] = s * p;
                    dv[l] = c * p;
                }
                __syncwarp();
            } while (fabs(ev[l]) > eps * tst1 && iter < 60);
        }
        __syncwarp();
        if (lane == 0) {
            dv[l] = dv[l] + f;
            ev[l] = 0.0;
        }
        __syncwarp();
    }
}
#undef VV

struct TruncWork {
    double* G[3];     // shared: Gram matrices -> eigenvectors, leading dimension n_k | 1
    double* dv;       // [3][kMaxN] shared: eigenvalues
    double* ev;       // [3][kMaxN] shared: work
    int* order;       // [3][kMaxN] shared: descending order
    int* rsel;        // [3] shared: selected ranks
    GramWork gw;
    bool dmma;        // Gram matrices on the FP64 tensor cores (gram_all_dmma) instead of DFMA (gram_mode)
    double* refV;     // global, 2 x kMaxN^2: sorted eigenvectors / rotated trailing block (small-eps refinement)
    double* trInv;    // [3] shared: 1 / trace of each Gram matrix
    int* misc;        // [4] shared: scratch integers of the refinement
    double* sorted;   // [kMaxN] shared: eigenvalues in descending order (refinement)
    long long* prof;  // optional phase timers (thread 0 of CTA 0): gram, eig, select, project, reconstruct, flux, derivative
};

// Truncated HOSVD of the dense tensor X (dims d).  Writes the factors to Uout[k] (leading dimension
// d[k], rcap[k] columns available), the core to coreOut (r0 x r1 x r2 packed), the ranks to rsel.
// W1/W2 are dense work buffers (>= N doubles each).
template <int T, int NM>
__device__ void hosvd_truncate(const double* X, const int d[3], double eps, int rmax, const int rcap[3], double* const Uout[3],
                               double* coreOut, double* W1, double* W2, const TruncWork& w)
{
    long long t0 = clock64();
    if (w.dmma) gram_all_dmma<T, NM>(X, d, w.G, w.gw);
    else
        for (int k = 0; k < 3; k++) gram_mode<T, NM>(X, d, k, w.G[k], w.gw);
    long long t1 = clock64();
    if (w.prof) w.prof[0] += t1 - t0;
    const int warp = threadIdx.x >> 5;
    if (warp < 3) {
        // unit trace: the rank rule only uses ratios of eigenvalues, and the QL iteration can then
        // use plain square roots
        double* G = w.G[warp];
        const int n = d[warp], ld = n | 1, lane = threadIdx.x & 31;
        double tr = 0.0;
        for (int i = lane; i < n; i += 32) tr += G[i + ld * i];
        tr = wsum(tr);
        if (tr > 0.0) {
            const double inv = 1.0 / tr;
            for (int j = 0; j < n; j++)
                for (int i = lane; i < n; i += 32) G[i + ld * j] *= inv;
        }
        if (lane == 0) w.trInv[warp] = tr > 0.0 ? 1.0 / tr : 0.0;
        __syncwarp();
    }
    if (warp < 3) eig_sym_warp(w.G[warp], d[warp], d[warp] | 1, w.dv + warp * kMaxN, w.ev + warp * kMaxN);
    __syncthreads();
    t0 = clock64();
    if (w.prof) w.prof[1] += t0 - t1;
    // descending order by rank counting (stable); eigenvalues of a Gram matrix are >= 0 up to rounding
    for (int q = threadIdx.x; q < 3 * kMaxN; q += blockDim.x) {
        const int k = q / kMaxN, j = q % kMaxN;
        if (j >= d[k]) continue;
        const double* lam = w.dv + k * kMaxN;
        int rank = 0;
        for (int i = 0; i < d[k]; i++)
            if (lam[i] > lam[j] || (lam[i] == lam[j] && i < j)) rank++;
        w.order[k * kMaxN + rank] = j;
    }
    __syncthreads();
    // ---- small compression errors.  Eigenvalues of a Gram matrix carry an absolute error of ~1e-16 of
    // the trace, i.e. singular values below ~1e-8 |sigma| are noise, while the rank rule compares them
    // with eps |sigma| / sqrt(3) (tucker.cpp:450-461; the class default is eps = 1e-10,
    // particle_data.h:49).  When the threshold is that low, the trailing eigen-directions are resolved
    // a second time inside their own subspace: Y = V_s^T X_(k), Gram matrix of Y, eigen-decomposition
    // W, V_s <- V_s W.  The cut between "leading" and "trailing" sits at sigma = 1e-4 |sigma| (lambda =
    // 1e-8 of the trace): a leading direction leaks into the computed trailing vectors with an amplitude
    // of ~1e-16 / lambda, i.e. it pollutes the trailing singular values by 1e-16 / sigma <= 1e-12, and
    // inside the block (largest value 1e-4) the Gram eigenvalues are again good to 1e-8 of that, so
    // singular values come out with ~1e-12 |sigma| of noise — enough for eps >= ~1e-11.  Below that
    // (and for precision 0 with a binding rank cap) a second pass resolves the block sigma < 1e-8 the
    // same way, which brings the noise to ~1e-16.
    {
        bool refine = eps > 0.0 ? (eps * eps / 3.0 < 1e-13) : false;
        if (eps == 0.0)
            for (int k = 0; k < 3; k++)
                if (min(rmax, rcap[k]) < d[k]) refine = true;   // precision 0 with a binding rank cap: the order matters
        const int nPass = !refine ? 0 : ((eps == 0.0 || eps < 1e-11) ? 2 : 1);
        for (int pass = 0; pass < nPass; pass++) {
            const double cut = pass == 0 ? 1e-8 : 1e-16;
            for (int k = 0; k < 3; k++) {
                const int n = d[k], ld = n | 1;
                double* lam = w.dv + k * kMaxN;
                int* ord = w.order + k * kMaxN;
                if (threadIdx.x == 0) {
                    int m = 0;
                    while (m < n && lam[ord[m]] >= cut) m++;
                    w.misc[0] = m;
                }
                __syncthreads();
                const int m = w.misc[0], sN = n - m;
                if (sN == 0) continue;
                double* Vs = w.refV;                        // n x n, sorted columns
                double* Rt = w.refV + (size_t)kMaxN * kMaxN;   // n x sN rotated trailing block
                for (int p = threadIdx.x; p < n * n; p += blockDim.x) Vs[p] = w.G[k][(p % n) + ld * ord[p / n]];
                for (int j = threadIdx.x; j < n; j += blockDim.x) w.sorted[j] = lam[ord[j]];   // sorted eigenvalues
                __syncthreads();
                // Y = X x_k V_s^T  (mode-k dimension sN)
                int dd[3] = {d[0], d[1], d[2]};
                mode_apply(X, W1, dd, k, Vs + (size_t)n * m, n, sN, true, w.gw.tile);
                dd[k] = sN;
                gram_mode<T, NM>(W1, dd, k, w.G[k], w.gw);   // sN x sN, leading dimension sN | 1
                const int lds = sN | 1;
                if (threadIdx.x < 32) {
                    const int lane = threadIdx.x;
                    double tr = 0.0;
                    for (int i = lane; i < sN; i += 32) tr += w.G[k][i + lds * i];
                    tr = wsum(tr);
                    const double inv = tr > 0.0 ? 1.0 / tr : 0.0;
                    for (int j = 0; j < sN; j++)
                        for (int i = lane; i < sN; i += 32) w.G[k][i + lds * j] *= inv;
                    __syncwarp();
                    if (tr > 0.0) eig_sym_warp(w.G[k], sN, lds, lam, w.ev + k * kMaxN);
                    __syncwarp();
                    if (lane == 0) {
                        // descending order of the sN values (insertion sort), scaled back to the unit-trace units of lam
                        const double scale = tr * w.trInv[k];
                        for (int j = 0; j < sN; j++) ord[j] = j;
                        if (tr > 0.0)
                            for (int a = 1; a < sN; a++) {
                                const int key = ord[a];
                                int b = a - 1;
                                while (b >= 0 && lam[ord[b]] < lam[key]) {
                                    ord[b + 1] = ord[b];
                                    b--;
                                }
                                ord[b + 1] = key;
                            }
                        w.misc[1] = tr > 0.0 ? 1 : 0;
                        for (int j = 0; j < sN; j++) lam[j] = fmax(lam[j], 0.0) * scale;
                    }
                }
                __syncthreads();
                const bool rotated = w.misc[1] != 0;
                // rotated trailing vectors: Rt(:, j) = sum_q Vs(:, m+q) W(q, ord[j])
                for (int p = threadIdx.x; p < n * sN; p += blockDim.x) {
                    const int i = p % n, j = p / n;
                    double acc = 0.0;
                    if (rotated)
                        for (int q = 0; q < sN; q++) acc = fma(Vs[i + (size_t)n * (m + q)], w.G[k][q + lds * ord[j]], acc);
                    else acc = Vs[i + (size_t)n * (m + j)];
                    Rt[p] = acc;
                }
                __syncthreads();
                // reassemble: sorted leading part, refined trailing part; eigenvalues and order follow
                if (threadIdx.x == 0) {
                    double tmp[kMaxN];
                    for (int j = 0; j < sN; j++) tmp[j] = rotated ? lam[ord[j]] : w.sorted[m + j];
                    for (int j = 0; j < m; j++) lam[j] = w.sorted[j];
                    for (int j = 0; j < sN; j++) lam[m + j] = tmp[j];
                }
                __syncthreads();
                for (int p = threadIdx.x; p < n * n; p += blockDim.x) {
                    const int i = p % n, j = p / n;
                    w.G[k][i + ld * j] = j < m ? Vs[p] : Rt[i + (size_t)n * (j - m)];
                }
                for (int j = threadIdx.x; j < n; j += blockDim.x) ord[j] = j;
                __syncthreads();
            }
        }
    }
    if (threadIdx.x < 3) {
        const int k = threadIdx.x, n = d[k];
        const double* lam = w.dv + k * kMaxN;
        const int* ord = w.order + k * kMaxN;
        // sigma_j = sqrt(lambda_j); |sigma|^2 = sum lambda_j              (tucker.cpp:450)
        double s2 = 0;
        for (int j = 0; j < n; j++) s2 += fmax(lam[j], 0.0);
        const double thr = eps * sqrt(s2) / sqrt(3.0);
        int r = 0;
        const int cap = min(rmax, rcap[k]);
        // precision 0 ("exact", particle_data.cpp:64-69): the reference keeps every sigma_j > 0, and an SVD
        // returns tiny positive values for the numerically-zero ones, i.e. everything is kept; here a
        // numerically-zero Gram eigenvalue may come out <= 0, which must not cut the prefix short (a
        // component of 1e-9 |sigma| has lambda = 1e-18, far below the eigenvalue noise)
        const bool keepAll = eps == 0.0 && s2 > 0.0;
        for (int j = 0; j < n; j++) {
            const double sig = sqrt(fmax(lam[ord[j]], 0.0));
            if (r == 0 || ((keepAll || sig > thr) && r < cap)) r++;   // sorted: a prefix is kept       (tucker.cpp:454-460)
            else break;
        }
        w.rsel[k] = r;
    }
    __syncthreads();
    for (int k = 0; k < 3; k++) {
        const int n = d[k], r = w.rsel[k], ld = n | 1;
        const int* ord = w.order + k * kMaxN;
        for (int p = threadIdx.x; p < n * r; p += blockDim.x) {
            const int i = p % n, j = p / n;
            Uout[k][i + n * j] = w.G[k][i + ld * ord[j]];
        }
    }
    __syncthreads();
    // core = X x1 U0^T x2 U1^T x3 U2^T
    t1 = clock64();
    if (w.prof) w.prof[2] += t1 - t0;
    // (mode products commute: the big coalesced contraction goes first, the strided mode-0 one last
    // on the smallest tensor)
    int dd[3] = {d[0], d[1], d[2]};
    mode_apply(X, W1, dd, 2, Uout[2], d[2], w.rsel[2], true, w.gw.tile);
    dd[2] = w.rsel[2];
    mode_apply(W1, W2, dd, 1, Uout[1], d[1], w.rsel[1], true, w.gw.tile);
    dd[1] = w.rsel[1];
    mode_apply(W2, coreOut, dd, 0, Uout[0], d[0], w.rsel[0], true, w.gw.tile);
    if (w.prof) w.prof[3] += clock64() - t1;
}

// dense = core x1 U0 x2 U1 x3 U2
__device__ void reconstruct(const double* core, const int r[3], double* const U[3], const int d[3], double* out, double* W1,
                            double* W2, double* Ms, long long* prof = nullptr)
{
    const long long t0 = clock64();
    int dd[3] = {r[0], r[1], r[2]};
    mode_apply(core, W1, dd, 0, U[0], d[0], d[0], false, Ms);
    dd[0] = d[0];
    mode_apply(W1, W2, dd, 1, U[1], d[1], d[1], false, Ms);
    dd[1] = d[1];
    mode_apply(W2, out, dd, 2, U[2], d[2], d[2], false, Ms);
    if (prof) prof[4] += clock64() - t0;
}


__device__ void slot_ptrs(double* base, const TuckerParams& P, double*& core, double* U[3])
{
    core = base;
    U[0] = base + P.coreCap;
    U[1] = U[0] + (size_t)P.n[0] * P.rcap[0];
    U[2] = U[1] + (size_t)P.n[1] * P.rcap[1];
}

template <int T, int NM>
__global__ void __launch_bounds__(T, T == kThreads ? 2 : 4) k_tucker(const TuckerParams P)
{
    extern __shared__ double sDyn[];   // 3 Gram/eigenvector matrices + one staging tile, (nmax|1)*nmax doubles each
    const int nmaxS = max(P.n[0], max(P.n[1], P.n[2]));
    const int matElems = (nmaxS | 1) * nmaxS;
    __shared__ double sDv[3 * kMaxN], sEv[3 * kMaxN];
    __shared__ int sOrder[3 * kMaxN];
    __shared__ unsigned short sGroups[gram_groups(kMaxN, kMaxN)];
    __shared__ int sR[3];
    __shared__ double sRed[T / 32][5];
    __shared__ TetRec rec;
    TruncWork w;
    for (int k = 0; k < 3; k++) w.G[k] = sDyn + (size_t)k * matElems;
    w.dv = sDv;
    w.ev = sEv;
    w.order = sOrder;
    w.rsel = sR;
    __shared__ double sTrInv[3], sSorted[kMaxN];
    __shared__ int sMisc[4];
    w.trInv = sTrInv;
    w.misc = sMisc;
    w.sorted = sSorted;
    w.gw.tile = sDyn + (size_t)3 * matElems;
    w.gw.tileCap = tile_cap(nmaxS);   // also stages a padded factor (mode_apply) and the padded DMMA tiles
    w.dmma = P.gramDmma != 0;
    w.gw.groups = sGroups;
    w.gw.nmax = nmaxS;
    build_groups(sGroups, nmaxS);
    __shared__ long long sProf[8];
    if (threadIdx.x < 8) sProf[threadIdx.x] = 0;
    w.prof = (P.prof && blockIdx.x == 0 && threadIdx.x == 0) ? sProf : nullptr;
    const long long tKernel = clock64();

    const int d[3] = {P.n[0], P.n[1], P.n[2]};
    const int N = P.N;
    double* scr = P.scratch + (size_t)blockIdx.x * P.scratchPerCTA;
    double* A = scr;
    double* B = A + N;
    double* RHS = B + N;
    double* W1 = RHS + N;
    double* W2 = W1 + N;
    double* Uw[3] = {W2 + N, W2 + N + (size_t)kMaxN * kMaxN, W2 + N + 2 * (size_t)kMaxN * kMaxN};   // factors of intermediates
    double* coreW = Uw[2] + (size_t)kMaxN * kMaxN;                                                  // [N]
    w.refV = coreW + N;                                                                             // 2 x kMaxN^2: small-eps refinement
    const int fullcap[3] = {d[0], d[1], d[2]};

    for (int t = blockIdx.x; t < P.nOwned; t += gridDim.x) {
        if (P.mode == 2) {   // reconstruct tet t into denseOut
            double *core, *U[3];
            slot_ptrs(const_cast<double*>(P.in) + (size_t)t * P.slot, P, core, U);
            const int r[3] = {P.rin[3 * t], P.rin[3 * t + 1], P.rin[3 * t + 2]};
            reconstruct(core, r, U, d, P.denseOut + (size_t)t * N, W1, W2, w.gw.tile, w.prof);
            continue;
        }
        if (P.mode == 1) {   // compress dense input into the slot (initial condition: exact, precision 0)
            double *core, *U[3];
            slot_ptrs(P.out + (size_t)t * P.slot, P, core, U);
            for (int e = threadIdx.x; e < N; e += blockDim.x) A[e] = P.denseIn[(size_t)t * N + e];
            __syncthreads();
            hosvd_truncate<T, NM>(A, d, 0.0, P.maxRank, P.rcap, U, core, W1, W2, w);
            if (threadIdx.x < 3) P.rout[3 * t + threadIdx.x] = sR[threadIdx.x];
            __syncthreads();
            continue;
        }
        // stage the tet record
        {
            const int* g = reinterpret_cast<const int*>(P.rec + t);
            int* s = reinterpret_cast<int*>(&rec);
            for (int i = threadIdx.x; i < (int)(sizeof(TetRec) / 4); i += blockDim.x) s[i] = g[i];
        }
        __syncthreads();
        if (P.mode == 4) {   // initial ghost fill: copy the current slot of every pushing tet to its peers
            for (int q = 0; q < 4; q++) {
                if (rec.pushPeer[q] < 0) continue;
                const double* src = P.in + (size_t)t * P.slot;
                double* dst = P.peerOut[rec.pushPeer[q]] + (size_t)rec.pushRow[q] * P.slot;
                for (size_t e = threadIdx.x; e < P.slot; e += blockDim.x) dst[e] = src[e];
                if (threadIdx.x < 3) P.peerRout[rec.pushPeer[q]][3 * (size_t)rec.pushRow[q] + threadIdx.x] = P.rin[3 * t + threadIdx.x];
            }
            __syncthreads();
            continue;
        }
        if (P.mode == 3) {   // |v.n| per face, rounded to rank <= 6 (solver.cpp:276-282)
            for (int f = 0; f < 4; f++) {
                for (int e = threadIdx.x; e < N; e += blockDim.x) {
                    const int i0 = e % d[0], i1 = (e / d[0]) % d[1], i2 = e / (d[0] * d[1]);
                    const double v0 = __dadd_rn(P.vmin[0], __dmul_rn((double)i0, P.step[0]));
                    const double v1 = __dadd_rn(P.vmin[1], __dmul_rn((double)i1, P.step[1]));
                    const double v2 = __dadd_rn(P.vmin[2], __dmul_rn((double)i2, P.step[2]));
                    A[e] = fabs(rec.nrm[f][0] * v0 + rec.nrm[f][1] * v1 + rec.nrm[f][2] * v2);
                }
                __syncthreads();
                double* vs = P.vnabs + ((size_t)t * 4 + f) * P.vslot;
                double* Uv[3] = {vs + 216, vs + 216 + 6 * d[0], vs + 216 + 6 * (d[0] + d[1])};
                const int cap6[3] = {min(6, d[0]), min(6, d[1]), min(6, d[2])};
                hosvd_truncate<T, NM>(A, d, P.epsAbs, 6, cap6, Uv, vs, W1, W2, w);
                if (threadIdx.x < 3) P.vnabsRanks[((size_t)t * 4 + f) * 3 + threadIdx.x] = sR[threadIdx.x];
                __syncthreads();
            }
            continue;
        }

        // ---- mode 0: one explicit step of tet t
        {
            double *core, *U[3];
            slot_ptrs(const_cast<double*>(P.in) + (size_t)t * P.slot, P, core, U);
            const int r[3] = {P.rin[3 * t], P.rin[3 * t + 1], P.rin[3 * t + 2]};
            reconstruct(core, r, U, d, A, W1, W2, w.gw.tile, w.prof);
        }
        for (int e = threadIdx.x; e < N; e += blockDim.x) RHS[e] = 0.0;
        __syncthreads();
        double wallAcc[4] = {0, 0, 0, 0};
        for (int f = 0; f < 4; f++) {
            const int bc = rec.bc[f];
            const bool pair = bc == VT_PBC_NONBOUNDARY || bc == VT_PBC_PERIODIC || bc == VT_PBC_SOURCE;
            if (bc == VT_PBC_SOURCE) {
                // sourcePDF takes the neighbour's place (solver.cpp:335-338); kept dense on the device
                const double* srow = P.src + (size_t)(-2 - rec.nbr[f]) * N;
                for (int e = threadIdx.x; e < N; e += blockDim.x) B[e] = srow[e];
                __syncthreads();
            } else if (pair) {
                const int nb = rec.nbr[f];
                double *core, *U[3];
                slot_ptrs(const_cast<double*>(P.in) + (size_t)nb * P.slot, P, core, U);
                const int r[3] = {P.rin[3 * nb], P.rin[3 * nb + 1], P.rin[3 * nb + 2]};
                reconstruct(core, r, U, d, B, W1, W2, w.gw.tile, w.prof);
            }
            const double coef = rec.coef[f];
            // |v.n| of this face from its rank-<=6 factors (coreW is free until the rounding below)
            {
                double* vs = P.vnabs + ((size_t)t * 4 + f) * P.vslot;
                double* Uv[3] = {vs + 216, vs + 216 + 6 * d[0], vs + 216 + 6 * (d[0] + d[1])};
                const int* vr = P.vnabsRanks + ((size_t)t * 4 + f) * 3;
                const int r[3] = {vr[0], vr[1], vr[2]};
                reconstruct(vs, r, Uv, d, coreW, W1, W2, w.gw.tile, w.prof);
            }
            const double* va = coreW;
            const long long tFlux = clock64();
            for (int e = threadIdx.x; e < N; e += blockDim.x) {
                const int i0 = e % d[0], i1 = (e / d[0]) % d[1], i2 = e / (d[0] * d[1]);
                const double v0 = __dadd_rn(P.vmin[0], __dmul_rn((double)i0, P.step[0]));
                const double v1 = __dadd_rn(P.vmin[1], __dmul_rn((double)i1, P.step[1]));
                const double v2 = __dadd_rn(P.vmin[2], __dmul_rn((double)i2, P.step[2]));
                const double vn = rec.nrm[f][0] * v0 + rec.nrm[f][1] * v1 + rec.nrm[f][2] * v2;
                const double a = A[e];
                double flux;
                if (pair) flux = 0.5 * (vn * (B[e] + a) - va[e] * (B[e] - a));        // solver.cpp:325-327
                else if (bc == VT_PBC_ABSORBING) {
                    flux = 0.5 * (vn * a + va[e] * a);                                // solver.cpp:331-332
                    if (rec.wallSlot[f] >= 0) wallAcc[f] += flux;
                } else flux = vn * a;                                                 // Free, solver.cpp:342
                RHS[e] = RHS[e] - coef * flux;                                        // solver.cpp:168
            }
            __syncthreads();
            if (w.prof) w.prof[5] += clock64() - tFlux;
            // rhs.Compress(comprErr, maxRank)                                           solver.cpp:182
            hosvd_truncate<T, NM>(RHS, d, P.eps, P.maxRank, fullcap, Uw, coreW, W1, W2, w);
            const int r[3] = {sR[0], sR[1], sR[2]};
            reconstruct(coreW, r, Uw, d, RHS, W1, W2, w.gw.tile, w.prof);
        }
        // acceleration: rhs -= (q/m)(E_k+ext_k) D_k f, D = zero-outside central difference       solver.cpp:187-200, 348-361
        {
            double g[3];
            for (int k = 0; k < 3; k++) g[k] = (P.qm * (P.E[3 * (size_t)t + k] + P.ext[k])) * P.inv2h[k];
            const int stride[3] = {1, d[0], d[0] * d[1]};
            const long long tDer = clock64();
            for (int e = threadIdx.x; e < N; e += blockDim.x) {
                const int i[3] = {e % d[0], (e / d[0]) % d[1], e / (d[0] * d[1])};
                double r = RHS[e];
                for (int k = 0; k < 3; k++) {
                    const double up = i[k] + 1 < d[k] ? A[e + stride[k]] : 0.0;
                    const double dn = i[k] > 0 ? A[e - stride[k]] : 0.0;
                    r = r - g[k] * (up - dn);
                }
                W1[e] = r;
            }
            __syncthreads();
            for (int e = threadIdx.x; e < N; e += blockDim.x) RHS[e] = W1[e];
            __syncthreads();
            if (w.prof) w.prof[6] += clock64() - tDer;
            hosvd_truncate<T, NM>(RHS, d, P.eps, P.maxRank, fullcap, Uw, coreW, W1, W2, w);   // solver.cpp:199
            const int r[3] = {sR[0], sR[1], sR[2]};
            reconstruct(coreW, r, Uw, d, RHS, W1, W2, w.gw.tile, w.prof);
        }
        // pdf += dt*rhs ; pdf.Compress                                                         solver.cpp:207-210
        for (int e = threadIdx.x; e < N; e += blockDim.x) B[e] = A[e] + P.dt * RHS[e];
        __syncthreads();
        {
            double *core, *U[3];
            slot_ptrs(P.out + (size_t)t * P.slot, P, core, U);
            hosvd_truncate<T, NM>(B, d, P.eps, P.maxRank, P.rcap, U, core, W1, W2, w);
            if (threadIdx.x < 3) P.rout[3 * t + threadIdx.x] = sR[threadIdx.x];
            const int r[3] = {sR[0], sR[1], sR[2]};
            reconstruct(core, r, U, d, B, W1, W2, w.gw.tile, w.prof);   // Density() sums the rounded tensor (particle_data.cpp:99)
            // multi-GPU: the new slot of a boundary tet also goes into the ghost rows of the peers
            for (int q = 0; q < 4; q++) {
                if (rec.pushPeer[q] < 0) continue;
                const double* src = P.out + (size_t)t * P.slot;
                double* dst = P.peerOut[rec.pushPeer[q]] + (size_t)rec.pushRow[q] * P.slot;
                for (size_t e = threadIdx.x; e < P.slot; e += blockDim.x) dst[e] = src[e];
                if (threadIdx.x < 3) P.peerRout[rec.pushPeer[q]][3 * (size_t)rec.pushRow[q] + threadIdx.x] = r[threadIdx.x];
            }
        }
        double acc = 0.0;
        for (int e = threadIdx.x; e < N; e += blockDim.x) acc += B[e];
        double vals[5] = {acc, wallAcc[0], wallAcc[1], wallAcc[2], wallAcc[3]};
        for (int q = 0; q < 5; q++) {
            double v = vals[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5][q] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot[5] = {0, 0, 0, 0, 0};
            for (int wv = 0; wv < T / 32; wv++)
                for (int q = 0; q < 5; q++) tot[q] += sRed[wv][q];
            P.density[t] = tot[0] * P.cellVolume;
            for (int f = 0; f < 4; f++)
                if (rec.wallSlot[f] >= 0) atomicAdd(P.wall + rec.wallSlot[f], P.wallScale * rec.area[f] * tot[1 + f]);
        }
        __syncthreads();
    }
    if (w.prof) {
        sProf[7] = clock64() - tKernel;
        for (int i = 0; i < 8; i++) P.prof[i] = sProf[i];
    }
}

}  // namespace
