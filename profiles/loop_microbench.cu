// How many warps does the (N,D)-merge loop need to saturate the FP64 pipe?  Same instruction pattern as the
// fused kernel's fast path: per component 2 LDS + 4 bins x (FMA, FMA, FMA, MUL), renorm every 16 components.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/loopmb profiles/loop_microbench.cu && /tmp/loopmb
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void renorm(double& N, double& D)
{
    const int hiD = __double2hiint(D);
    const int k = (hiD & 0x7ff00000) - 0x3ff00000;
    D = __hiloint2double(hiD - k, __double2loint(D));
    N = __hiloint2double(__double2hiint(N) - k, __double2loint(N));
}
template <int BPT>
__global__ void k(double* out, int ncomp, int reps)
{
    __shared__ double2 sc[512];
    __shared__ double a[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { sc[i] = make_double2(0.37 + 1e-3 * i, -0.2 - 1e-3 * i); a[i] = 0.5 + 1e-3 * i; }
    __syncthreads();
    double u[BPT], N[BPT], D[BPT];
#pragma unroll
    for (int j = 0; j < BPT; j++) { u[j] = 0.01 * (threadIdx.x * BPT + j); N[j] = 0; D[j] = 1; }
    for (int r = 0; r < reps; r++) {
        for (int k0 = 0; k0 + 16 <= ncomp; k0 += 16) {
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                const double2 p = sc[k0 + kk];
                const double aa = a[k0 + kk];
#pragma unroll
                for (int j = 0; j < BPT; j++) {
                    const double e = fma(u[j], p.x, p.y);
                    const double t = fma(e, e, aa);
                    N[j] = fma(N[j], t, D[j]);
                    D[j] *= t;
                }
            }
#pragma unroll
            for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < BPT; j++) s += N[j] / D[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int BPT>
void run(int warps_per_sm, int sms, double* d)
{
    const int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm);
    const int blocks = sms * (warps_per_sm * 32 / threads);
    const int ncomp = 96, reps = 400;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<BPT><<<blocks, threads>>>(d, ncomp, reps);
    cudaEventRecord(e0);
    k<BPT><<<blocks, threads>>>(d, ncomp, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double instr = 4.0 * BPT * ncomp * (double)reps * blocks * threads;   // FP64 lane-instructions
    printf("warps/SM %2d BPT %d : %6.2f T FP64 instr/s  (%.0f%% of 18.3)\n", warps_per_sm, BPT, instr / (ms * 1e-3) / 1e12, 100 * instr / (ms * 1e-3) / 18.3e12);
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, sizeof(double) * sms * 64 * 32 * 2);
    const int ws[] = {4, 8, 16, 24, 32, 48};
    for (int w : ws) { run<2>(w, sms, d); run<4>(w, sms, d); run<8>(w, sms, d); }
    return 0;
}
