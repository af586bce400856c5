// FP64 FMA on B200: dependent-issue latency and how many independent chains one warp / three warps per SM sub-partition
// need to reach the pipe rate.  Prints cycles per DFMA per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu && ./dfma_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b)
{
    double f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) f[i] = threadIdx.x * 1e-3 + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 64 / ILP; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) f[i] = fma(f[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(double* d, long long* dc, int threads)
{
    const int iters = 2000;
    k<ILP><<<148, threads>>>(d, dc, iters, 1.0000001, 1e-9);
    k<ILP><<<148, threads>>>(d, dc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per_warp = (double)c / (64.0 * iters);                    // cycles per DFMA of one warp
    const double per_smsp = per_warp / (threads / 128.0);                  // warps per sub-partition = threads / 128
    printf("  ILP %2d, %d warps/SMSP: %6.2f cycles per DFMA per warp, %5.2f per sub-partition\n", ILP, threads / 128, per_warp, per_smsp);
}

int main()
{
    double* d; long long* dc;
    cudaMalloc(&d, 148 * 1024 * sizeof(double)); cudaMalloc(&dc, 8);
    for (int threads : {128, 256, 384, 512}) {
        run<1>(d, dc, threads); run<2>(d, dc, threads); run<4>(d, dc, threads); run<8>(d, dc, threads); run<16>(d, dc, threads);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
