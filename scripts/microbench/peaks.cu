// Micro-benchmarks behind the roofline denominators that MEASURED_PEAKS.json does not carry:
// L2->SM read bandwidth (L2-resident working set), HBM read bandwidth, and the FP64 DFMA peak.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void k_read(const double2* __restrict__ p, size_t n, int reps, double* out)
{
    double acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n; i += 4 * stride) {
            double2 a, b, c, d;
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(a.x), "=d"(a.y) : "l"(p + i));
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(b.x), "=d"(b.y) : "l"(p + i + stride));
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(c.x), "=d"(c.y) : "l"(p + i + 2 * stride));
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(d.x), "=d"(d.y) : "l"(p + i + 3 * stride));
            acc += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
        }
    }
    if (acc == 123.456) *out = acc;
}

__global__ void k_dfma(double* out, int iters)
{
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// FP64 tensor-core peak: mma.sync.aligned.m8n8k4 (256 FMA per warp instruction), 8 independent accumulators
__global__ void k_dmma(double* out, int iters)
{
    double c[8][2];
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = threadIdx.x + i;
    const double a = 1.0000001, b = 0.9999999;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float timeit(void (*launch)(void*), void* arg, int n)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch(arg);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < n; i++) launch(arg);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / n;
}

struct RA { const double2* p; size_t n; int reps; double* out; int blocks; };
static void launch_read(void* a) { RA* r = (RA*)a; k_read<<<r->blocks, 512>>>(r->p, r->n, r->reps, r->out); }
struct FA { double* out; int iters; int blocks; };
static void launch_dfma(void* a) { FA* f = (FA*)a; k_dfma<<<f->blocks, 256>>>(f->out, f->iters); }
static void launch_dmma(void* a) { FA* f = (FA*)a; k_dmma<<<f->blocks, 256>>>(f->out, f->iters); }

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out;
    cudaMalloc(&out, 1 << 24);
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    const size_t sizesMB[] = {16, 32, 48, 64, 96, 4096};
    for (size_t s : sizesMB) {
        const size_t bytes = s << 20;
        double2* buf;
        cudaMalloc(&buf, bytes);
        cudaMemset(buf, 0, bytes);
        RA ra{buf, bytes / 16, s >= 1024 ? 1 : 20, out, sms * 4};
        const float ms = timeit(launch_read, &ra, 5);
        const double gbs = (double)bytes * ra.reps / (ms * 1e-3) / 1e9;
        printf(", \"read_%zuMB_GBs\": %.1f", s, gbs);
        cudaFree(buf);
    }
    FA fa{out, 20000, sms * 8};
    const float ms = timeit(launch_dfma, &fa, 3);
    const double tf = 2.0 * 8 * fa.iters * (double)fa.blocks * 256 / (ms * 1e-3) / 1e12;
    printf(", \"dfma_TFLOPs\": %.2f", tf);
    FA fm{out, 5000, sms * 8};
    const float msm = timeit(launch_dmma, &fm, 3);
    // per warp and iteration: 8 mma x 256 FMA
    const double tfm = 2.0 * 8 * 256 * fm.iters * (double)fm.blocks * (256 / 32) / (msm * 1e-3) / 1e12;
    printf(", \"dmma_m8n8k4_TFLOPs\": %.2f}\n", tfm);
    return 0;
}
