// Does the FP64 tensor-core instruction (mma.sync.m8n8k4.f64) run beside the scalar DFMA pipe on B200, or on it?
// Times (a) DFMA only, (b) DMMA only, (c) both interleaved in every warp, (d) half the warps each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_vs_dfma dmma_vs_dfma.cu && ./dmma_vs_dfma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters, double a, double b)
{
    double f[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { f[i] = threadIdx.x * 1e-3 + i; c[i] = i; }
    const bool do_fma = MODE == 0 || MODE == 2 || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 0);
    const bool do_mma = MODE == 1 || MODE == 2 || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 1);
    for (int it = 0; it < iters; it++) {
        if (do_fma) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) f[i] = fma(f[i], a, b);         // 32 DFMA
        }
        if (do_mma) {
#pragma unroll
            for (int i = 0; i < 4; i++) dmma(c[2 * i], c[2 * i + 1], a, b);  // 4 DMMA = 4 x 256 FMA per warp = 32 per thread
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += f[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(double* d, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 512>>>(d, 100, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148, 512>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    double* d; cudaMalloc(&d, 148 * 512 * sizeof(double));
    const int iters = 20000;
    const double fmas = 148.0 * 512 * 32.0 * iters;      // FMAs per kind per launch (modes 0,1,2); mode 3: half of each
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters);
    printf("DFMA only      : %8.3f ms  %6.2f TFLOP/s\n", t0, 2 * fmas / t0 * 1e-9);
    printf("DMMA only      : %8.3f ms  %6.2f TFLOP/s\n", t1, 2 * fmas / t1 * 1e-9);
    printf("both, each warp: %8.3f ms  %6.2f TFLOP/s (sum)  -> %s\n", t2, 4 * fmas / t2 * 1e-9, t2 < 0.8 * (t0 + t1) ? "pipes overlap" : "one pipe");
    printf("half the warps : %8.3f ms  %6.2f TFLOP/s (sum)\n", t3, 2 * fmas / t3 * 1e-9);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
