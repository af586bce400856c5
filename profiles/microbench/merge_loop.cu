// The fused kernel's main loop in isolation: N/D merge of K shared-memory entries into 4 bins per thread
// (e = fma(u,s,c); t = fma(e,e,a); N = fma(N,t,D); D *= t), with 1, 2, 3 or 4 warps per SM sub-partition.
// Prints cycles per FP64 instruction per warp and per sub-partition (pipe rate: 2.0).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o merge_loop merge_loop.cu && ./merge_loop
#include <cstdio>
#include <cuda_runtime.h>

struct __align__(16) FastEntry { double s, c, a, pad; };
constexpr int K = 352;

__device__ __forceinline__ void renorm(double& N, double& D)
{
    const int hiD = __double2hiint(D);
    const int k = (hiD & 0x7ff00000) - 0x3ff00000;
    D = __hiloint2double(hiD - k, __double2loint(D));
    N = __hiloint2double(__double2hiint(N) - k, __double2loint(N));
}

template <int BPT, int VAR>
__global__ void __launch_bounds__(512, 1) k(double* out, long long* cyc, int reps)
{
    __shared__ FastEntry fast[K];
    for (int i = threadIdx.x; i < K; i += blockDim.x) { fast[i].s = 0.3 + 1e-3 * i; fast[i].c = -0.1 * i; fast[i].a = 0.05; fast[i].pad = 0; }
    double u[BPT], N[BPT], D[BPT];
#pragma unroll
    for (int j = 0; j < BPT; j++) { u[j] = threadIdx.x * 0.01 + j; N[j] = 0.0; D[j] = 1.0; }
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        for (int k0 = 0; k0 + 16 <= K; k0 += 16) {
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                const double2 p = *reinterpret_cast<const double2*>(&fast[k0 + kk].s);
                const double a = fast[k0 + kk].a;
#pragma unroll
                for (int j = 0; j < BPT; j++) {
                    if (VAR == 2) {            // e, t only: 2 instructions, no carried chain (accumulate into N to keep them live)
                        const double e = fma(u[j], p.x, p.y);
                        N[j] = fma(e, e, N[j]);
                    } else if (VAR == 3) {     // N, D chain only, t straight from the entry
                        N[j] = fma(N[j], a, D[j]);
                        D[j] *= a;
                    } else {
                        const double e = fma(u[j], p.x, p.y);
                        const double t = fma(e, e, a);
                        N[j] = fma(N[j], t, D[j]);
                        D[j] *= t;
                    }
                }
            }
            if (VAR != 1) {
#pragma unroll
                for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < BPT; j++) s += N[j] / D[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int BPT, int VAR>
void run(double* d, long long* dc, int threads)
{
    const int reps = 20;
    k<BPT, VAR><<<148, threads>>>(d, dc, reps);
    k<BPT, VAR><<<148, threads>>>(d, dc, reps);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, sizeof(c), cudaMemcpyDeviceToHost);
    const double n = ((VAR == 2 || VAR == 3) ? 2.0 : 4.0) * BPT * (K / 16 * 16) * reps;                     // FP64 instructions per thread
    printf("  variant %d, %d bins/thread, %d warps/SMSP: %6.2f cycles per FP64 instruction per warp, %5.2f per sub-partition\n", VAR, BPT, threads / 128,
           c / n, c / n / (threads / 128.0));
}

int main()
{
    double* d; long long* dc;
    cudaMalloc(&d, 148 * 1024 * sizeof(double)); cudaMalloc(&dc, 8);
    printf("variants: 0 = the kernel's loop, 1 = without the exponent renormalisation, 2 = e and t only, 3 = N/D chain only\n");
    for (int threads : {128, 384}) { run<2, 0>(d, dc, threads); run<4, 0>(d, dc, threads); run<8, 0>(d, dc, threads); run<4, 1>(d, dc, threads); run<4, 2>(d, dc, threads); run<4, 3>(d, dc, threads); }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
