// K1 — fused full-format kinetic update for sm_100a.
//
// One launch performs, for every owned tet t and velocity node v, what the reference spreads
// over three OpenMP loops and ~45 tensor temporaries (src/solver.cpp:141-212):
//   rhs  = - sum_f (A_f/V_t) * flux_f(v)                      (solver.cpp:159-184, _Flux :314-346)
//          - sum_k (q/m)(E_k+ext_k) * d f/d v_k               (solver.cpp:187-200, :363-404)
//   f'   = f + dt * rhs                                       (solver.cpp:204-211)
// plus the two reductions the next step needs: sum_v f' (ParticleData::Density,
// particle_data.cpp:98-99) and sum_v flux on absorbing+collecting faces (solver.cpp:171-178).
//
// Work decomposition: CTA = (tet, chunk of consecutive i2-planes).  CTAs are numbered
// brick-major, chunk-next, tet-fastest so that the CTAs in flight touch one velocity chunk of
// one spatially compact brick of tets: the four neighbour reads of every tet then hit the
// 126 MB L2 instead of HBM (the state is laid out in a locality order chosen by the host).
// Inside a CTA a thread owns one i0-vector (VEC doubles) and walks lines (i1,i2); the
// face-normal velocity v.n is split into a per-thread part n_x*v0(i0) kept in registers and a
// per-line part n_y*v1(i1)+n_z*v2(i2) tabulated once per CTA in shared memory.
#include "vt_internal.h"

namespace vt {

namespace {

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
    double x;
};
template <>
struct Vec<2> {
    double x, y;
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> ldv(const double* p);
template <>
__device__ __forceinline__ Vec<1> ldv<1>(const double* p)
{
    Vec<1> r;
    r.x = __ldg(p);
    return r;
}
template <>
__device__ __forceinline__ Vec<2> ldv<2>(const double* p)
{
    double2 t = __ldg(reinterpret_cast<const double2*>(p));
    Vec<2> r;
    r.x = t.x;
    r.y = t.y;
    return r;
}
// neighbour rows are streamed: read-only path, do not pollute L1
template <int VEC>
__device__ __forceinline__ Vec<VEC> ldv_stream(const double* p);
template <>
__device__ __forceinline__ Vec<1> ldv_stream<1>(const double* p)
{
    Vec<1> r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r.x) : "l"(p));
    return r;
}
template <>
__device__ __forceinline__ Vec<2> ldv_stream<2>(const double* p)
{
    Vec<2> r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
template <int VEC>
__device__ __forceinline__ void stv(double* p, const Vec<VEC>& v);
template <>
__device__ __forceinline__ void stv<1>(double* p, const Vec<1>& v)
{
    *p = v.x;
}
template <>
__device__ __forceinline__ void stv<2>(double* p, const Vec<2>& v)
{
    *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One face's contribution for one element, keeping the reference's expression shape
// 0.5*(vn*(fa+f) - |vn|*(fa-f)) (solver.cpp:325-327) so differences stay at FMA level.
__device__ __forceinline__ double flux_pair(double vn, double fa, double f)
{
    return 0.5 * (vn * (fa + f) - fabs(vn) * (fa - f));
}
__device__ __forceinline__ double flux_absorb(double vn, double f)
{
    return 0.5 * (vn * f + fabs(vn) * f);  // solver.cpp:331-332
}

template <int VEC, bool SHFL>
__global__ void __launch_bounds__(256) k_full_step(const StepParams p)
{
    extern __shared__ double sm[];
    __shared__ TetRec rec;
    __shared__ double red[8][5];

    // ---- decode (brick, chunk, tet)
    const int perBrick = p.brickTets * p.nChunks;
    const int brick = blockIdx.x / perBrick;
    const int base = brick * p.brickTets;
    const int nb = min(p.brickTets, p.nOwned - base);
    const int r = blockIdx.x - brick * perBrick;
    const int chunk = r / nb;
    const int tet = base + (r - chunk * nb);
    const int pl0 = chunk * p.chunkPlanes;
    const int npl = min(p.chunkPlanes, p.n2 - pl0);
    const int nLines = npl * p.n1;

    const int tid = threadIdx.x;
    // ---- stage the tet record and the velocity tables
    {
        const int* g = reinterpret_cast<const int*>(p.rec + tet);
        int* s = reinterpret_cast<int*>(&rec);
        for (int i = tid; i < (int)(sizeof(TetRec) / 4); i += blockDim.x) s[i] = g[i];
    }
    double* v0 = sm;                   // [n0]
    double* bl = sm + p.n0;            // [nLines][4]
    // velocity_grid.cpp:30: v = min + i*step, two roundings as in the reference
    for (int i = tid; i < p.n0; i += blockDim.x) v0[i] = __dadd_rn(p.vmin[0], __dmul_rn((double)i, p.step[0]));
    __syncthreads();
    for (int l = tid; l < nLines; l += blockDim.x) {
        int i1 = l % p.n1;
        int i2 = pl0 + l / p.n1;
        double v1 = __dadd_rn(p.vmin[1], __dmul_rn((double)i1, p.step[1]));
        double v2 = __dadd_rn(p.vmin[2], __dmul_rn((double)i2, p.step[2]));
#pragma unroll
        for (int f = 0; f < 4; f++) bl[4 * l + f] = rec.nrm[f][1] * v1 + rec.nrm[f][2] * v2;
    }
    __syncthreads();

    const int c0 = tid % p.nvec0;
    const int lg = tid / p.nvec0;
    const bool active = lg < p.nLG;

    double accDens = 0.0;
    double accWall[4] = {0.0, 0.0, 0.0, 0.0};

    if (active) {
        const int i0 = c0 * VEC;
        // per-thread part of v.n
        double a[4][VEC];
#pragma unroll
        for (int f = 0; f < 4; f++)
#pragma unroll
            for (int u = 0; u < VEC; u++) a[f][u] = rec.nrm[f][0] * v0[i0 + u];

        const double* frow = p.f + (size_t)tet * p.N;
        double* nrow = p.fn + (size_t)tet * p.N;
        const double* nb_row[4];
#pragma unroll
        for (int f = 0; f < 4; f++) {
            int n = rec.nbr[f];
            nb_row[f] = n >= 0 ? p.f + (size_t)n * p.N : (n <= -2 ? p.src + (size_t)(-2 - n) * p.N : nullptr);
        }
        // force_k / (2 step_k): (q/m)*(E_k+ext_k) as solver.cpp:193-194
        double g[3];
#pragma unroll
        for (int k = 0; k < 3; k++) g[k] = (p.qm * (p.E[3 * (size_t)tet + k] + p.ext[k])) * p.inv2h[k];

        const int plane = p.n0 * p.n1;
        // x-neighbour offsets (periodic wrap, solver.cpp:380-389)
        const int offL = (i0 == 0) ? (p.n0 - 1) : -1;                          // left of the first element
        const int offR = (VEC - 1) + ((i0 + VEC == p.n0) ? -(p.n0 - 1) : 1);   // right of the last element
        int bcs[4];
        double coef[4];
        bool collect[4];
#pragma unroll
        for (int f = 0; f < 4; f++) {
            bcs[f] = rec.bc[f];
            coef[f] = rec.coef[f];
            collect[f] = rec.wallSlot[f] >= 0;
        }

        for (int l = lg; l < nLines; l += p.nLG) {
            const int i1 = l % p.n1;
            const int i2 = pl0 + l / p.n1;
            const int e = (i2 * p.n1 + i1) * p.n0 + i0;

            Vec<VEC> fc = ldv<VEC>(frow + e);
            // neighbour tets first: longest latency
            Vec<VEC> fa[4];
#pragma unroll
            for (int f = 0; f < 4; f++) {
                fa[f].x = 0.0;
                if (VEC == 2) ((double*)&fa[f])[VEC - 1] = 0.0;
                if (nb_row[f]) fa[f] = ldv_stream<VEC>(nb_row[f] + e);
            }

            // own-row stencil
            const int e1m = e + ((i1 == 0) ? (p.n1 - 1) : -1) * p.n0;
            const int e1p = e + ((i1 == p.n1 - 1) ? -(p.n1 - 1) : 1) * p.n0;
            const int e2m = e + ((i2 == 0) ? (p.n2 - 1) : -1) * plane;
            const int e2p = e + ((i2 == p.n2 - 1) ? -(p.n2 - 1) : 1) * plane;
            Vec<VEC> f1m = ldv<VEC>(frow + e1m), f1p = ldv<VEC>(frow + e1p);
            Vec<VEC> f2m = ldv<VEC>(frow + e2m), f2p = ldv<VEC>(frow + e2p);
            double fl, fr;  // left of first element, right of last element
            if (SHFL) {
                const double last = (VEC == 2) ? ((const double*)&fc)[VEC - 1] : fc.x;
                fl = __shfl_sync(0xffffffffu, last, (c0 + p.nvec0 - 1) & (p.nvec0 - 1), p.nvec0);
                fr = __shfl_sync(0xffffffffu, fc.x, (c0 + 1) & (p.nvec0 - 1), p.nvec0);
            } else {
                fl = __ldg(frow + e + offL);
                fr = __ldg(frow + e + offR);
            }

            const double* fcv = (const double*)&fc;
            double out[VEC];
#pragma unroll
            for (int u = 0; u < VEC; u++) {
                const double fv = fcv[u];
                double rhs = 0.0;
#pragma unroll
                for (int f = 0; f < 4; f++) {
                    const double vn = a[f][u] + bl[4 * l + f];
                    const int bc = bcs[f];
                    double flux;
                    if (bc == VT_PBC_NONBOUNDARY || bc == VT_PBC_PERIODIC || bc == VT_PBC_SOURCE) {
                        flux = flux_pair(vn, ((const double*)&fa[f])[u], fv);
                    } else if (bc == VT_PBC_ABSORBING) {
                        flux = flux_absorb(vn, fv);
                        if (collect[f]) accWall[f] += flux;
                    } else {
                        flux = vn * fv;  // Free, solver.cpp:342
                    }
                    rhs = rhs - coef[f] * flux;  // solver.cpp:168
                }
                // d/dv0
                const double xm = (u == 0) ? fl : fcv[u - 1];
                const double xp = (u == VEC - 1) ? fr : fcv[u + 1];
                rhs = rhs - g[0] * (xp - xm);
                rhs = rhs - g[1] * (((const double*)&f1p)[u] - ((const double*)&f1m)[u]);
                rhs = rhs - g[2] * (((const double*)&f2p)[u] - ((const double*)&f2m)[u]);
                out[u] = fv + p.dt * rhs;  // solver.cpp:207
                accDens += out[u];
            }
            Vec<VEC> o;
            o.x = out[0];
            if (VEC == 2) ((double*)&o)[1] = out[VEC - 1];
            stv<VEC>(nrow + e, o);
        }
    }

    // ---- reductions: sum_v f' for Density(), sum_v flux for the wall charge
    const bool anyWall = (rec.wallSlot[0] >= 0) | (rec.wallSlot[1] >= 0) | (rec.wallSlot[2] >= 0) | (rec.wallSlot[3] >= 0);
    const int warp = tid >> 5, lane = tid & 31;
    double s = warp_sum(accDens);
    if (lane == 0) red[warp][0] = s;
    if (anyWall) {
#pragma unroll
        for (int f = 0; f < 4; f++) {
            double w = warp_sum(accWall[f]);
            if (lane == 0) red[warp][1 + f] = w;
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int nw = blockDim.x >> 5;
        double d = 0.0;
        for (int w = 0; w < nw; w++) d += red[w][0];
        p.densPartial[(size_t)tet * p.nChunks + chunk] = d;
        if (anyWall) {
            for (int f = 0; f < 4; f++) {
                if (rec.wallSlot[f] < 0) continue;
                double q = 0.0;
                for (int w = 0; w < nw; w++) q += red[w][1 + f];
                // charge * (timeStep * area * flux.Sum() * cellVolume), solver.cpp:173-177
                atomicAdd(p.wall + rec.wallSlot[f], p.wallScale * rec.area[f] * q);
            }
        }
    }
}

__global__ void k_density_reduce(const double* __restrict__ partial, double* __restrict__ density,
                                 int nOwned, int nChunks, double cellVolume)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nOwned) return;
    double s = 0.0;
    for (int c = 0; c < nChunks; c++) s += partial[(size_t)t * nChunks + c];
    density[t] = s * cellVolume;  // particle_data.cpp:99
}

// Density() of the current state when no step produced it (initial condition).
__global__ void k_density_full(const double* __restrict__ f, double* __restrict__ density, int N,
                               double cellVolume)
{
    __shared__ double red[32];
    const double* row = f + (size_t)blockIdx.x * N;
    double s = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s += row[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double d = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) d += red[w];
        density[blockIdx.x] = d * cellVolume;
    }
}

}  // namespace

void launch_density(vt_ctx* ctx, Species& sp)
{
    if (ctx->nOwned == 0) return;
    k_density_full<<<ctx->nOwned, 256, 0, ctx->stream>>>(sp.f[sp.cur], sp.density, sp.N, sp.cellVolume);
    ctx->launches++;
    VT_CUDA(cudaGetLastError());
    sp.densityValid = true;
}

void launch_full_step(vt_ctx* ctx, Species& sp, double dt, const double ext[3])
{
    if (ctx->nOwned == 0) return;
    if (sp.danglingFaces > 0)
        throw std::runtime_error(std::to_string(sp.danglingFaces) +
                                 " boundary faces have no neighbour and no particle BC (Absorbing/Free/Source); "
                                 "the reference dereferences a null adjTets there (solver.cpp:319)");
    StepParams p;
    p.f = sp.f[sp.cur];
    p.fn = sp.f[sp.cur ^ 1];
    p.rec = sp.rec;
    p.E = ctx->E;
    p.src = sp.src;
    p.wall = sp.wall;
    p.nOwned = ctx->nOwned;
    p.n0 = sp.n[0];
    p.n1 = sp.n[1];
    p.n2 = sp.n[2];
    p.N = sp.N;
    const int VEC = (sp.n[0] % 2 == 0) ? 2 : 1;
    p.nvec0 = sp.n[0] / VEC;
    const int threads = 256;
    if (p.nvec0 > threads) throw std::runtime_error("vt_step_full: n0 too large for one CTA line");
    p.nLG = threads / p.nvec0;
    int cp = ctx->chunkPlanes > 0 ? ctx->chunkPlanes : sp.n[2];
    if (cp > sp.n[2]) cp = sp.n[2];
    // keep the per-line table within shared memory
    while ((size_t)(sp.n[0] + 4 * cp * sp.n[1]) * 8 > 160 * 1024 && cp > 1) cp = (cp + 1) / 2;
    p.chunkPlanes = cp;
    p.nChunks = (sp.n[2] + cp - 1) / cp;
    p.brickTets = ctx->brickTets > 0 ? ctx->brickTets : ctx->nOwned;
    if (p.brickTets > ctx->nOwned) p.brickTets = ctx->nOwned;
    for (int k = 0; k < 3; k++) {
        p.vmin[k] = sp.vmin[k];
        p.step[k] = sp.step[k];
        p.inv2h[k] = 1.0 / (2 * sp.step[k]);
        p.ext[k] = ext ? ext[k] : 0.0;
    }
    p.qm = sp.charge / sp.mass;
    p.dt = dt;
    p.wallScale = sp.charge * dt * sp.cellVolume;

    size_t need = (size_t)ctx->nOwned * p.nChunks;
    if ((size_t)sp.densPartialCap < need) {
        if (sp.densPartial) VT_CUDA(cudaFree(sp.densPartial));
        VT_CUDA(cudaMalloc(&sp.densPartial, need * sizeof(double)));
        sp.densPartialCap = (int)need;
    }
    p.densPartial = sp.densPartial;

    const size_t smem = (size_t)(sp.n[0] + 4 * cp * sp.n[1]) * sizeof(double);
    const long long grid = (long long)ctx->nOwned * p.nChunks;
    if (grid > 2147483647LL) throw std::runtime_error("vt_step_full: grid too large");
    const int nLines = cp * sp.n[1];
    const bool pow2 = (p.nvec0 & (p.nvec0 - 1)) == 0;
    const bool shfl = VEC == 2 && pow2 && p.nvec0 <= 32 && p.nLG * p.nvec0 == threads &&
                      (nLines % p.nLG == 0) && (sp.n[2] % cp == 0) && ctx->variant == 0;

    auto launch = [&](auto kern) {
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaEvent_t e0 = ctx->ev0, e1 = ctx->ev1;
        if (ctx->profiling) {
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
        kern<<<(unsigned)grid, threads, smem, ctx->stream>>>(p);
        VT_CUDA(cudaEventRecord(e1, ctx->stream));
    };
    if (VEC == 2 && shfl) launch(k_full_step<2, true>);
    else if (VEC == 2) launch(k_full_step<2, false>);
    else launch(k_full_step<1, false>);
    ctx->launches++;
    VT_CUDA(cudaGetLastError());

    k_density_reduce<<<(ctx->nOwned + 255) / 256, 256, 0, ctx->stream>>>(sp.densPartial, sp.density, ctx->nOwned,
                                                                        p.nChunks, sp.cellVolume);
    ctx->launches++;
    VT_CUDA(cudaGetLastError());
    sp.cur ^= 1;
    sp.densityValid = true;
}

}  // namespace vt
