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
//
// Two arithmetic variants of the interior face flux (vt_step_config's `variant`, bit 1):
//   reference shape   0.5*(vn*(fa+f) - |vn|*(fa-f))                       (solver.cpp:325-327)
//   upwind select     vn>0 ? vn*f : vn*fa, with A_f/V folded into vn: the same first-order
//                     upwind flux in 2 FP64 operations per face instead of 6; differs from the
//                     reference expression by rounding only (the parity tests run both).
#include "vt_internal.h"

namespace vt {

namespace {

template <int VEC>
struct Vec {
    double v[VEC];
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> ldv(const double* p)
{
    Vec<VEC> r;
    if (VEC == 2) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(p));
        r.v[0] = t.x;
        r.v[VEC - 1] = t.y;
    } else {
        r.v[0] = __ldg(p);
    }
    return r;
}
// neighbour rows are streamed once per CTA: read-only path, do not pollute L1
template <int VEC>
__device__ __forceinline__ Vec<VEC> ldv_stream(const double* p)
{
    Vec<VEC> r;
    if (VEC == 2) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                     : "=d"(r.v[0]), "=d"(r.v[VEC - 1])
                     : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r.v[0]) : "l"(p));
    }
    return r;
}
template <int VEC>
__device__ __forceinline__ void stv(double* p, const Vec<VEC>& x)
{
    if (VEC == 2) *reinterpret_cast<double2*>(p) = make_double2(x.v[0], x.v[VEC - 1]);
    else *p = x.v[0];
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

// The per-line loop of one CTA.  GENERIC: per-face boundary-condition dispatch (CTA-uniform
// branches); otherwise all four faces are interior/periodic/source pairs.
template <int VEC, bool SHFL, bool GENERIC, bool UPWIND, bool HALO>
__device__ __forceinline__ void line_loop(const StepParams& p, const TetRec& rec, const int tet, const int c0,
                                          const int lg, const int pl0, const int nLines,
                                          const double* __restrict__ v0, const double* __restrict__ bl,
                                          double& accDens, double (&accWall)[4])
{
    const int n0 = p.n0, n1 = p.n1, n2 = p.n2;
    const int i0 = c0 * VEC;
    double a[4][VEC], hc[4];
    bool pair[4], absorbing[4], collect[4];
#pragma unroll
    for (int f = 0; f < 4; f++) {
        const int bc = GENERIC ? rec.bc[f] : VT_PBC_NONBOUNDARY;
        pair[f] = is_pair(bc);
        absorbing[f] = GENERIC && bc == VT_PBC_ABSORBING;
        collect[f] = GENERIC && rec.wallSlot[f] >= 0;
        const bool pre = UPWIND && pair[f];
        // reference shape: rhs -= coef*(0.5*t) == rhs - (0.5*coef)*t (the halving is exact)
        hc[f] = pair[f] ? 0.5 * rec.coef[f] : rec.coef[f];
#pragma unroll
        for (int u = 0; u < VEC; u++) a[f][u] = (pre ? rec.coef[f] : 1.0) * (rec.nrm[f][0] * v0[i0 + u]);
    }
    const double* frow = p.f + (size_t)tet * p.N;
    double* nrow = p.fn + (size_t)tet * p.N;
    const double* nbp[4];
#pragma unroll
    for (int f = 0; f < 4; f++) {
        const int n = rec.nbr[f];
        nbp[f] = n >= 0 ? p.f + (size_t)n * p.N : (n <= -2 ? p.src + (size_t)(-2 - n) * p.N : frow);
    }
    double* push[4] = {nullptr, nullptr, nullptr, nullptr};
    if (HALO) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (rec.pushPeer[q] >= 0) push[q] = p.peerFn[rec.pushPeer[q]] + (size_t)rec.pushRow[q] * p.N;
    }
    // force_k / (2 step_k) with force = (q/m)*(E_k+ext_k) as solver.cpp:193-194
    double g[3];
#pragma unroll
    for (int k = 0; k < 3; k++) g[k] = (p.qm * (p.E[3 * (size_t)tet + k] + p.ext[k])) * p.inv2h[k];

    const int plane = n0 * n1;
    // periodic wrap in velocity space (solver.cpp:380-389)
    const int offL = (i0 == 0) ? (n0 - 1) : -1;                         // left of the first element
    const int offR = (VEC - 1) + ((i0 + VEC == n0) ? -(n0 - 1) : 1);    // right of the last element

    int l = lg;
    int i1 = l % n1, pl = l / n1;
    const int dI1 = p.nLG % n1, dPl = p.nLG / n1;
    int e = (pl0 * n1 + l) * n0 + i0;
    const int de = p.nLG * n0;

    for (; l < nLines; l += p.nLG) {
        const int i2 = pl0 + pl;
        const Vec<VEC> fc = ldv<VEC>(frow + e);
        Vec<VEC> fa[4];
#pragma unroll
        for (int f = 0; f < 4; f++) {
            if (!GENERIC || pair[f]) fa[f] = ldv_stream<VEC>(nbp[f] + e);
            else
#pragma unroll
                for (int u = 0; u < VEC; u++) fa[f].v[u] = 0.0;
        }
        const int e1m = e + ((i1 == 0) ? (n1 - 1) : -1) * n0;
        const int e1p = e + ((i1 == n1 - 1) ? -(n1 - 1) : 1) * n0;
        const int e2m = e + ((i2 == 0) ? (n2 - 1) : -1) * plane;
        const int e2p = e + ((i2 == n2 - 1) ? -(n2 - 1) : 1) * plane;
        const Vec<VEC> f1m = ldv<VEC>(frow + e1m), f1p = ldv<VEC>(frow + e1p);
        const Vec<VEC> f2m = ldv<VEC>(frow + e2m), f2p = ldv<VEC>(frow + e2p);
        double fl, fr;   // left of the first element, right of the last
        if (SHFL) {
            fl = __shfl_sync(0xffffffffu, fc.v[VEC - 1], (c0 + p.nvec0 - 1) & (p.nvec0 - 1), p.nvec0);
            fr = __shfl_sync(0xffffffffu, fc.v[0], (c0 + 1) & (p.nvec0 - 1), p.nvec0);
        } else {
            fl = __ldg(frow + e + offL);
            fr = __ldg(frow + e + offR);
        }
        double blf[4];
#pragma unroll
        for (int f = 0; f < 4; f++) blf[f] = bl[4 * l + f];

        Vec<VEC> out;
#pragma unroll
        for (int u = 0; u < VEC; u++) {
            const double fv = fc.v[u];
            double rhs = 0.0;
#pragma unroll
            for (int f = 0; f < 4; f++) {
                const double vn = a[f][u] + blf[f];
                if (!GENERIC || pair[f]) {
                    if (UPWIND) {
                        // vn here is (A/V)*v.n; upwind value: own cell for outflow, neighbour for inflow
                        rhs = fma(-vn, vn > 0.0 ? fv : fa[f].v[u], rhs);
                    } else {
                        const double s = fa[f].v[u] + fv, d = fa[f].v[u] - fv;
                        rhs = fma(-hc[f], fma(vn, s, -(fabs(vn) * d)), rhs);   // solver.cpp:325-327, :168
                    }
                } else if (absorbing[f]) {
                    const double flux = 0.5 * (vn * fv + fabs(vn) * fv);      // solver.cpp:331-332
                    if (collect[f]) accWall[f] += flux;
                    rhs = fma(-hc[f], flux, rhs);
                } else {
                    rhs = fma(-hc[f], vn * fv, rhs);                          // Free, solver.cpp:342
                }
            }
            const double xm = (u == 0) ? fl : fc.v[u - 1 >= 0 ? u - 1 : 0];
            const double xp = (u == VEC - 1) ? fr : fc.v[u + 1 < VEC ? u + 1 : VEC - 1];
            rhs = fma(-g[0], xp - xm, rhs);
            rhs = fma(-g[1], f1p.v[u] - f1m.v[u], rhs);
            rhs = fma(-g[2], f2p.v[u] - f2m.v[u], rhs);
            out.v[u] = fma(p.dt, rhs, fv);                                    // solver.cpp:207
            accDens += out.v[u];
        }
        stv<VEC>(nrow + e, out);
        if (HALO) {
            // fused halo exchange: the same registers go to the ghost rows on the peer GPUs
            // (NVLink peer stores), so the transfer overlaps the sweep tile by tile
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (push[q]) stv<VEC>(push[q] + e, out);
        }

        e += de;
        i1 += dI1;
        pl += dPl;
        if (i1 >= n1) {
            i1 -= n1;
            pl++;
        }
    }
}

template <int VEC, bool SHFL, bool UPWIND, int MINB>
__global__ void __launch_bounds__(256, MINB) k_full_step(const StepParams p)
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
    double* v0 = sm;                          // [n0] (padded to an even count)
    double* bl = sm + ((p.n0 + 1) & ~1);      // [nLines][4]
    // velocity_grid.cpp:30: v = min + i*step, two roundings as in the reference
    for (int i = tid; i < p.n0; i += blockDim.x) v0[i] = __dadd_rn(p.vmin[0], __dmul_rn((double)i, p.step[0]));
    __syncthreads();
    for (int l = tid; l < nLines; l += blockDim.x) {
        const int i1 = l % p.n1;
        const int i2 = pl0 + l / p.n1;
        const double v1 = __dadd_rn(p.vmin[1], __dmul_rn((double)i1, p.step[1]));
        const double v2 = __dadd_rn(p.vmin[2], __dmul_rn((double)i2, p.step[2]));
#pragma unroll
        for (int f = 0; f < 4; f++) {
            const double b = rec.nrm[f][1] * v1 + rec.nrm[f][2] * v2;
            bl[4 * l + f] = (UPWIND && is_pair(rec.bc[f])) ? rec.coef[f] * b : b;
        }
    }
    __syncthreads();

    const int c0 = tid % p.nvec0;
    const int lg = tid / p.nvec0;

    double accDens = 0.0;
    double accWall[4] = {0.0, 0.0, 0.0, 0.0};
    const bool allPair = is_pair(rec.bc[0]) && is_pair(rec.bc[1]) && is_pair(rec.bc[2]) && is_pair(rec.bc[3]);
    if (lg < p.nLG) {
        const bool halo = rec.pushPeer[0] >= 0;
        if (allPair && !halo) line_loop<VEC, SHFL, false, UPWIND, false>(p, rec, tet, c0, lg, pl0, nLines, v0, bl, accDens, accWall);
        else if (allPair) line_loop<VEC, SHFL, false, UPWIND, true>(p, rec, tet, c0, lg, pl0, nLines, v0, bl, accDens, accWall);
        else line_loop<VEC, false, true, UPWIND, true>(p, rec, tet, c0, lg, pl0, nLines, v0, bl, accDens, accWall);
    }

    // ---- reductions: sum_v f' for Density(), sum_v flux for the wall charge
    const bool anyWall = (rec.wallSlot[0] >= 0) | (rec.wallSlot[1] >= 0) | (rec.wallSlot[2] >= 0) | (rec.wallSlot[3] >= 0);
    const int warp = tid >> 5, lane = tid & 31;
    const double s = warp_sum(accDens);
    if (lane == 0) red[warp][0] = s;
    if (anyWall) {
#pragma unroll
        for (int f = 0; f < 4; f++) {
            const double w = warp_sum(accWall[f]);
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
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
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

bool launch_full_step_tma(vt_ctx* ctx, Species& sp, StepParams& p, bool upwind, cudaEvent_t e0, cudaEvent_t e1);

void launch_density(vt_ctx* ctx, Species& sp)
{
    if (ctx->nOwned == 0) return;
    if (sp.tucker) tucker_materialize(ctx, sp);
    k_density_full<<<ctx->nOwned, 256, 0, ctx->stream>>>(sp.f[sp.cur], sp.density, sp.N, sp.cellVolume);
    ctx->launches++;
    VT_CUDA(cudaGetLastError());
    sp.densityValid = true;
}

void launch_full_step(vt_ctx* ctx, Species& sp, double dt, const double ext[3])
{
    if (ctx->nOwned == 0) return;
    if (sp.tucker) throw std::runtime_error("vt_step_full: the species is in Tucker format (use vt_step_tucker)");
    if (sp.danglingFaces > 0)
        throw std::runtime_error(std::to_string(sp.danglingFaces) +
                                 " boundary faces have no neighbour and no particle BC (Absorbing/Free/Source); "
                                 "the reference dereferences a null adjTets there (solver.cpp:319)");
    // variant bit 6 (the default): pick the kernel from the grid — the bulk-copy pipeline with eight
    // consumer warps and the upwind-select arithmetic where whole planes of >= 4 KiB can be staged,
    // the register-staged kernel otherwise.  Explicit bits are honoured as given.
    struct VariantScope {
        vt_ctx* c;
        int saved;
        ~VariantScope() { c->variant = saved; }
    } variantScope{ctx, ctx->variant};
    if (ctx->variant & 64) {
        const int PE = sp.n[0] * sp.n[1];
        const bool planes = sp.n[0] % 2 == 0 && PE * 8 >= 4096 && PE / 2 <= 2048;
        // bits 7 and up are modifiers of the bulk-copy pipeline and stay as given
        ctx->variant = (ctx->variant & ~127) | (planes ? (2 | 16 | 32) : 2);
    }
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
    while ((size_t)(sp.n[0] + 2 + 4 * cp * sp.n[1]) * 8 > 160 * 1024 && cp > 1) cp = (cp + 1) / 2;
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
    for (int i = 0; i < kMaxPeers; i++) p.peerFn[i] = i < sp.nPeers ? sp.peerF[i][sp.cur ^ 1] : nullptr;

    p.densSplit = 1;
    const size_t need = (size_t)ctx->nOwned * p.nChunks * ((ctx->variant & 16) ? 16 : 1);
    if (need > 2147483647ULL) throw std::runtime_error("vt_step_full: too many density partial sums");
    if ((size_t)sp.densPartialCap < need) {
        if (sp.densPartial) VT_CUDA(cudaFree(sp.densPartial));
        sp.densPartial = nullptr;
        VT_CUDA(cudaMalloc(&sp.densPartial, need * sizeof(double)));
        sp.densPartialCap = (int)need;
    }
    p.densPartial = sp.densPartial;

    const size_t smem = (size_t)(((sp.n[0] + 1) & ~1) + 4 * cp * sp.n[1]) * sizeof(double);
    const long long grid = (long long)ctx->nOwned * p.nChunks;
    if (grid > 2147483647LL) throw std::runtime_error("vt_step_full: grid too large");
    const int nLines = cp * sp.n[1];
    const bool pow2 = (p.nvec0 & (p.nvec0 - 1)) == 0;
    // variant bit 0: disable the shuffle path; bit 1: upwind-select arithmetic
    const bool shfl = VEC == 2 && pow2 && p.nvec0 <= 32 && p.nLG * p.nvec0 == threads &&
                      (nLines % p.nLG == 0) && (sp.n[2] % cp == 0) && !(ctx->variant & 1);
    const bool upwind = (ctx->variant & 2) != 0;

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
    auto launch = [&](auto kern) {
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VT_CUDA(cudaEventRecord(e0, ctx->stream));
        kern<<<(unsigned)grid, threads, smem, ctx->stream>>>(p);
        VT_CUDA(cudaEventRecord(e1, ctx->stream));
    };
    // bit 4: the persistent bulk-copy (cp.async.bulk) producer/consumer pipeline — it counts its own
    // launches (one, or two when boundary and interior tets are launched separately).  Falls through
    // to the register-staged kernel where it does not apply.  (Bit 3 selected a cp.async ring in round
    // 1; it lost to the bulk-copy pipeline at every size and was removed.)
    bool useTma = false;
    if (ctx->variant & 16) {
        useTma = launch_full_step_tma(ctx, sp, p, upwind, e0, e1);
        if (useTma) ctx->launches--;   // compensates the increment below
    }
    // variant bit 2: ask the compiler for 3 resident CTAs per SM (<= 85 registers) instead of 2
    const bool dense = (ctx->variant & 4) != 0;
    if (useTma) {
        // launched above
    } else if (VEC == 2 && shfl) {
        if (upwind) dense ? launch(k_full_step<2, true, true, 3>) : launch(k_full_step<2, true, true, 2>);
        else dense ? launch(k_full_step<2, true, false, 3>) : launch(k_full_step<2, true, false, 2>);
    } else if (VEC == 2) {
        if (upwind) dense ? launch(k_full_step<2, false, true, 3>) : launch(k_full_step<2, false, true, 2>);
        else dense ? launch(k_full_step<2, false, false, 3>) : launch(k_full_step<2, false, false, 2>);
    } else {
        if (upwind) launch(k_full_step<1, false, true, 2>);
        else launch(k_full_step<1, false, false, 2>);
    }
    ctx->launches++;
    VT_CUDA(cudaGetLastError());

    k_density_reduce<<<(ctx->nOwned + 255) / 256, 256, 0, ctx->stream>>>(sp.densPartial, sp.density, ctx->nOwned,
                                                                        p.nChunks * p.densSplit, sp.cellVolume);
    ctx->launches++;
    VT_CUDA(cudaGetLastError());
    sp.cur ^= 1;
    sp.densityValid = true;
}

}  // namespace vt
