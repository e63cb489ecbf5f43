// Dependent-issue latency of the FP64 instructions the Tucker kernels chain together, in SM cycles (clock64 around
// a chain of N dependent operations by one warp).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int OP>
__global__ void probe(double* out, long long* cycles, double x0, double y0)
{
    __shared__ double sh[64];
    sh[threadIdx.x & 63] = x0;
    __syncthreads();
    double x = x0 + threadIdx.x * 1e-9, y = y0;
    double c[2] = {x, y};
    int idx = threadIdx.x & 31;
    constexpr int N = 2048;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = fma(x, y, 1e-9);                 // DFMA
        else if (OP == 1) x = x * y;                      // DMUL
        else if (OP == 2) x = 1.0 / x + 1.5;              // double division
        else if (OP == 3) x = sqrt(x) + 1.5;              // double square root
        else if (OP == 4) dmma884(c, x, y);               // DMMA, accumulator chain
        else if (OP == 5) { idx = (int)sh[idx] & 31; }    // LDS + F2I chain (shared-memory pointer chase)
        else if (OP == 6) x = (double)((float)x * 1.0001f + 1e-3f);   // F2F round trip + FFMA
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[OP] = (t1 - t0) / N;
    out[threadIdx.x] = x + c[0] + c[1] + idx;
}

int main()
{
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1024 * sizeof(double));
    cudaMallocManaged(&cyc, 8 * sizeof(long long));
    probe<0><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
    probe<1><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
    probe<2><<<1, 32>>>(out, cyc, 1.7, 0.9999999);
    probe<3><<<1, 32>>>(out, cyc, 1.7, 0.9999999);
    probe<4><<<1, 32>>>(out, cyc, 1.0000001, 1e-9);
    probe<5><<<1, 32>>>(out, cyc, 3.0, 1.0);
    probe<6><<<1, 32>>>(out, cyc, 1.0000001, 1.0);
    cudaDeviceSynchronize();
    printf("{\"what\": \"dependent-issue latency, SM cycles per operation, one warp\", \"dfma\": %lld, \"dmul\": %lld, \"ddiv\": %lld, \"dsqrt\": %lld, "
           "\"dmma_m8n8k4_accumulate\": %lld, \"lds_pointer_chase\": %lld, \"f2f_ffma_f2f\": %lld}\n",
           cyc[0], cyc[1], cyc[2], cyc[3], cyc[4], cyc[5], cyc[6]);
    return 0;
}
