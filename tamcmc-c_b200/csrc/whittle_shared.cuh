// whittle_shared.cuh -- device helpers shared by the two fused model + Whittle kernels (whittle.cu: persistent ring of
// producer / consumer warps; whittle_tiles.cu: one CTA per tile): reciprocal and exponent renormalisation of the merged
// fraction, the Gaussian envelope of models 0 / 1, the per-chain finalisation (likelihoods.cpp:17-40, model_def.cpp:394-405),
// the exchange step of a bin-sharded spectrum and the end-of-launch protocol of the last CTA.
#pragma once
#include "tamcmc_dev.h"
#include "kernels.h"
#include <cuda_runtime.h>
#include <math.h>

namespace {

// 1/x for the Whittle terms: hardware approximation (rel. error <= 2^-23) + ONE Newton step -> rel. error <= 2^-46 = 1.4e-14,
// four orders of magnitude inside the 1e-10 bar on logL (the model-spectrum entry uses a true division).  x = 0 -> inf.
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

// (N, D) *= 2^-k with k = exponent(D): exact, integer pipe only.  D > 0 always; N >= 0, and N == 0
// only while D == 1 (nothing merged yet), where k == 0.
__device__ __forceinline__ void renorm(double& N, double& D)
{
    const int hiD = __double2hiint(D);
    const int k = (hiD & 0x7ff00000) - 0x3ff00000;
    D = __hiloint2double(hiD - k, __double2loint(D));
    N = __hiloint2double(__double2hiint(N) - k, __double2loint(N));
}

// Gaussian envelope of the models without Lorentzians (ids 0, 1): |H| exp(-0.5 (x - nu)^2 / sigma^2), times the sinc^2
// leakage of Kallinger+2014 eq. 1 for model 0 (models.cpp:5693-5694, 5768-5772; noise_models.cpp:89-97).  Kept out of
// line: it is off the path of the Lorentzian models and must not cost them registers.
__device__ __noinline__ double gauss_envelope(const NoiseRec* nz, double x)
{
    const double d = x - nz->gnu;
    double g = nz->gH * exp(-(d * d) * nz->gk);
    if (nz->gauss == 2) {
        const double a = (0.5 * 3.14159265358979323846 * x) / nz->xnyq;
        const double eta = (x == 0.0) ? 1.0 : sin(a) / a;
        g *= eta * eta;
    }
    return g;
}

// One warp per (star, chain): sum of the per-tile partials in tile order (fixed shape: bitwise reproducible).  Each tile
// contributes sum(y/M) - ln(prod 1/M) = S - ln(m) - E ln 2;  likelihood_chi22p: f = -p*S_total with p truncated to long
// (model_def.cpp:399), divided by Tcoefs[m] (model_def.cpp:401).
__device__ void finalize_chains(const WhittleArgs& A, int warp, int lane, int nwarps)
{
    const double LN2 = 0.693147180559945309417232121458;
    for (int sc = warp; sc < A.nsc; sc += nwarps) {
        if (A.status[sc] != 0) continue;
        const int ntiles = A.stars[sc / A.Nchains].ntiles;
        const double* part = A.partial + 3 * (size_t)sc * A.tiles_stride;
        double acc = 0.0;
        for (int t = lane; t < ntiles; t += 32) acc += part[3 * t] - (log(part[3 * t + 1]) + part[3 * t + 2] * LN2);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
        if (lane == 0) {
            if (A.raw_sum || A.xworld > 1) A.out[sc] = acc;          // the local sum S: all-reduced by the caller / exchanged below
            else if (A.likelihood == 1) A.out[sc] = ((-acc) / 2) / A.Tcoefs[sc % A.Nchains];        // likelihoods.cpp:36-37, model_def.cpp:405
            else {
                const double pl = (double)(long long)A.p;
                A.out[sc] = (-pl * acc) / A.Tcoefs[sc % A.Nchains];
            }
        }
    }
}

// The exchange step of a bin-sharded spectrum, run by the last CTA of every rank once its chains' LOCAL sums S
// are in A.out: (1) the sums go into block [parity][rank] of the exchange buffer of every rank (peer stores over NVLink; own
// buffer included), (2) after a system-scope fence the rank's flag [parity][rank] of every buffer takes the new epoch value,
// (3) the CTA waits for all flags of its own buffer, (4) adds the ranks' sums in RANK ORDER -- every rank gets the same bits --
// and (5) applies the likelihood's factor and 1 / Tcoefs like finalize_chains.  Two parities: a rank one evaluation ahead never
// overwrites values a slower rank still has to read.  A peer that never shows up (its process died) ends the wait after ~2 s:
// NaN results and TAMCMC_ST_NONFINITE for every chain, instead of a kernel that spins for ever.
__device__ void exchange_and_finalize(const WhittleArgs& A, int tid, int NT)
{
    __shared__ unsigned int s_epoch, s_timeout;
    if (tid == 0) { const unsigned e = *A.xepoch + 1u; s_epoch = e ? e : 1u; s_timeout = 0u; }
    __syncthreads();
    const unsigned e = s_epoch;
    const int par = (int)(e & 1u), W = A.xworld, R = A.xrank;
    const size_t blk = ((size_t)par * TAMCMC_XCHG_MAX_WORLD + (size_t)R) * (size_t)A.xstride;
    for (int p = 0; p < W; p++) {
        double* dst = reinterpret_cast<double*>(A.xpeer[p]) + blk;
        for (int i = tid; i < A.nsc; i += NT) dst[i] = (A.status[i] != 0) ? 0.0 : A.out[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < W) {
        volatile unsigned int* f = reinterpret_cast<volatile unsigned int*>(reinterpret_cast<unsigned char*>(A.xpeer[tid]) + tamcmc_xchg_flag_offset(A.xstride))
                                   + par * TAMCMC_XCHG_MAX_WORLD + R;
        *f = e;
    }
    if (tid < W) {
        volatile unsigned int* f = reinterpret_cast<volatile unsigned int*>(reinterpret_cast<unsigned char*>(A.xpeer[R]) + tamcmc_xchg_flag_offset(A.xstride))
                                   + par * TAMCMC_XCHG_MAX_WORLD + tid;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*f != e) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 2000000000ull) { atomicExch(&s_timeout, 1u); break; }
        }
    }
    __threadfence_system();
    __syncthreads();
    const bool dead = s_timeout != 0u;
    const double* own = reinterpret_cast<const double*>(A.xpeer[R]) + (size_t)par * TAMCMC_XCHG_MAX_WORLD * (size_t)A.xstride;
    for (int i = tid; i < A.nsc; i += NT) {
        if (dead) { A.out[i] = nan(""); const_cast<int*>(A.status)[i] |= TAMCMC_ST_NONFINITE; continue; }
        if (A.status[i] != 0) continue;                                   // the expander already put NaN there
        double S = 0.0;
        for (int r = 0; r < W; r++) S += __ldcv(own + (size_t)r * A.xstride + i);
        if (A.likelihood == 1) A.out[i] = ((-S) / 2) / A.Tcoefs[i % A.Nchains];
        else A.out[i] = (-(double)(long long)A.p * S) / A.Tcoefs[i % A.Nchains];
    }
    __syncthreads();
    if (tid == 0) *A.xepoch = e;
}

// End of a launch, called by EVERY thread of every CTA once the CTA's tiles are done: the LAST CTA to arrive turns the per-tile
// partials into the per-chain results (finalize_chains, the exchange of a bin-sharded spectrum), mirrors them to the host if
// asked to, re-arms the work queue and publishes the launch epoch.
__device__ void finish_launch(const WhittleArgs& A, int tid, int NT, bool finalize = true)
{
    __shared__ unsigned int s_last;
    __syncthreads();
    if (tid == 0) { __threadfence(); s_last = (atomicAdd(&A.qctl->ctas_done, 1u) == gridDim.x - 1u) ? 1u : 0u; }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (finalize) finalize_chains(A, tid >> 5, tid & 31, NT / 32);
    if (A.xworld > 1 && !A.raw_sum) exchange_and_finalize(A, tid, NT);
    if (A.host_flag) {
        // host mirror: every writer fences its own stores at system scope, the barrier orders them before the flag
        __syncthreads();
        for (int i = tid; i < A.nsc; i += NT) { A.host_logL[i] = A.out[i]; A.host_status[i] = A.status[i]; }
        if (tid == 0) *A.host_overflow = 0u;
        __threadfence_system();
        __syncthreads();
    }
    if (tid == 0) {
        QueueCtl* q = A.qctl;
#pragma unroll
        for (int k = 0; k < TAMCMC_NBUCKETS; k++) q->count[k] = 0u;
        q->head = 0u; q->ctas_done = 0u; q->bg_count = 0u; q->bg_head = 0u;
        unsigned e = *A.epoch + 1u;
        e = e ? e : 1u;
        *A.epoch = e;                                   // next launch's ready-flag value (never 0)
        if (A.host_flag) *reinterpret_cast<volatile unsigned int*>(A.host_flag) = e;      // published last
    }
}

}  // namespace
