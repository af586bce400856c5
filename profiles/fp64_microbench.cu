// FP64 pipe characterisation on B200: DFMA throughput vs warps/SM and per-thread ILP.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64mb profiles/fp64_microbench.cu && /tmp/fp64mb
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double seed)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = seed + i;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps_per_sm, int sms, double* d)
{
    const int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm);
    const int blocks_per_sm = warps_per_sm * 32 / threads;
    const int blocks = sms * blocks_per_sm;
    const int iters = 20000 / ILP;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<blocks, threads>>>(d, iters, 1.0);
    cudaEventRecord(e0);
    k<ILP><<<blocks, threads>>>(d, iters, 2.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8 * ILP * (double)iters * blocks * threads;
    printf("warps/SM %2d ILP %d : %7.2f TFLOP/s\n", warps_per_sm, ILP, fl / (ms * 1e-3) / 1e12);
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, sizeof(double) * sms * 64 * 32 * 2);
    const int ws[] = {4, 8, 12, 16, 24, 32, 48, 64};
    for (int w : ws) { run<1>(w, sms, d); run<2>(w, sms, d); run<4>(w, sms, d); run<8>(w, sms, d); }
    return 0;
}
