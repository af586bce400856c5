// expand.cu -- device-side parameter expander (sm_100a).
//
// One CTA per (star, chain).  Turns a raw TAMCMC parameter vector
// (params[Nparams] + plength[11], layout of tamcmc/sources/io_ms_global.cpp:1315-1398)
// into the flat tables the fused kernel consumes: one ModeRec per mode (with the
// bit-exact bin window of set_imin_imax), up to 7 CompRec per mode (nu_nlm, scaled
// width, inverse height), and one NoiseRec per chain.
//
// THIS FILE IS COMPILED WITH -fmad=false: every quantity that feeds a bin window
// (lin_interpol, a1 terms, set_imin_imax) must round exactly like the reference's
// scalar C++ does without FMA contraction (SURVEY.md 7 "Window bit-exactness").
//
// Reference functions restated here (new code, same arithmetic):
//   set_imin_imax            tamcmc/sources/build_lorentzian.cpp:595-676
//   nu_nlm of build_l_mode_* tamcmc/sources/build_lorentzian.cpp:48-348
//   amplitude_ratio / dmm    tamcmc/sources/function_rot.cpp:15-101
//   lin_interpol             tamcmc/sources/interpol.cpp:13-43
//   linfit / eta0_fct        tamcmc/sources/linfit.cpp:17-35, models.cpp:6065-6084
//   model_* unpacking        tamcmc/sources/models.cpp (line ranges at each family below)
#include "tamcmc_dev.h"
#include "kernels.h"
#include <cuda_runtime.h>
#include <math.h>

// Pslm(s,l,m) as double-double (hi, lo) and Qlm(l,m); filled by the host (capi.cu) with
// long double arithmetic that follows acoefs.cpp:51-110 and build_lorentzian.cpp:583-592.
__constant__ double c_Pslm_hi[7][4][7];
__constant__ double c_Pslm_lo[7][4][7];
__constant__ double c_Qlm[4][7];
// dmm(l, i, 0, beta) for i >= 0 (function_rot.cpp:76-88) with the integer factors tabulated by the host:
// coef[l][i][s] = combi(l, l-i-s) * combi(l, s) * (-1)^(l-i-s)   (INTEGER-division combi, function_rot.cpp:90-92)
// nnum[l][i]    = sqrt(factorial(l+i) * factorial(l-i)),  nden[l] = sqrt(factorial(l) * factorial(l))
__constant__ double c_dmm_coef[4][4][4];
__constant__ double c_dmm_nnum[4][4];
__constant__ double c_dmm_nden[4];

cudaError_t tamcmc_upload_dmm_tables(const double* coef, const double* nnum, const double* nden)
{
    cudaError_t e = cudaMemcpyToSymbol(c_dmm_coef, coef, sizeof(double) * 4 * 4 * 4);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_dmm_nnum, nnum, sizeof(double) * 4 * 4);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_dmm_nden, nden, sizeof(double) * 4);
}

cudaError_t tamcmc_upload_tables(const double* P_hi, const double* P_lo, const double* Q)
{
    cudaError_t e = cudaMemcpyToSymbol(c_Pslm_hi, P_hi, sizeof(double) * 7 * 4 * 7);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_Pslm_lo, P_lo, sizeof(double) * 7 * 4 * 7);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_Qlm, Q, sizeof(double) * 4 * 7);
}

namespace {

#ifdef TAMCMC_TRACE
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define ETRACE(slot) do { if (A.trace && threadIdx.x == 0 && blockIdx.x == 0) A.trace[(blockIdx.y ? 32 : 0) + (slot)] = gtime(); } while (0)
// stamp written by an arbitrary thread of CTA (0, 0) (sub-phases of one warp)
#define ETRACE_T(slot, thread) do { if (A.trace && threadIdx.x == (thread) && blockIdx.x == 0 && blockIdx.y == 0) A.trace[(slot)] = gtime(); } while (0)
#else
#define ETRACE(slot) do { } while (0)
#define ETRACE_T(slot, thread) do { } while (0)
#endif

__device__ __forceinline__ double Phi(int s, int l, int m) { return c_Pslm_hi[s][l][m + 3]; }
__device__ __forceinline__ double Plo(int s, int l, int m) { return c_Pslm_lo[s][l][m + 3]; }
__device__ __forceinline__ double Qlm(int l, int m) { return c_Qlm[l][m + 3]; }

// ---- double-double accumulator: emulates the reference's long double sums of
// a_j * Pslm(j,l,m) (build_lorentzian.cpp:222) to better than 2^-100 before the
// final rounding to double. ----
struct dd { double hi, lo; };
__device__ __forceinline__ dd dd_from(double a) { dd r; r.hi = a; r.lo = 0.0; return r; }
__device__ __forceinline__ dd dd_add(dd a, dd b)
{
    double s = a.hi + b.hi;
    double bb = s - a.hi;
    double err = (a.hi - (s - bb)) + (b.hi - bb);
    err += a.lo + b.lo;
    dd r; r.hi = s + err; r.lo = err - (r.hi - s);
    return r;
}
__device__ __forceinline__ dd dd_mul_d_dd(double a, double bhi, double blo)
{
    double p = a * bhi;
    double e = __fma_rn(a, bhi, -p);   // explicit: exact product error (unaffected by -fmad=false)
    e += a * blo;
    dd r; r.hi = p + e; r.lo = e - (r.hi - p);
    return r;
}

// x^n for a small non-negative integer n by repeated multiplication (the reference calls pow();
// the two agree to a few ulp, far inside the 1e-10 budget, and nothing here feeds a bin window)
__device__ __forceinline__ double ipow(double x, int n)
{
    double r = 1.0;
    for (int k = 0; k < n; k++) r *= x;
    return r;
}

// One entry of amplitude_ratio(l, inc) (function_rot.cpp:15-74): V(m=i) = d^l_{i,0}(inc)^2.  Following the four
// fill loops of function_rot, column l of the matrix holds dmm(l,|i|,0,+-beta) up to a sign, which the square
// removes; dmm's sum over s (function_rot.cpp:79-83) runs over the host-tabulated integer factors.
__device__ __noinline__ double d_amplitude_ratio_entry(int l, int i, double beta_deg)
{
    const double PI = 3.141592653589793238462643;
    const double angle = PI * beta_deg / 180.;
    const int a = i < 0 ? -i : i;
    double si, co;
    sincos(angle / 2., &si, &co);
    double sum = 0;
    for (int s = 0; s <= l - a; s++) {
        double var = c_dmm_coef[l][a][s];
        var = var * ipow(co, 2 * s + a) * ipow(si, 2 * l - 2 * s - a);
        sum = sum + var;
    }
    sum = sum * c_dmm_nnum[l][a];
    sum = sum / c_dmm_nden[l];
    return sum * sum;
}

// interpol.cpp:13-43
__device__ __noinline__ double d_lin_interpol(const double* x, const double* y, int Nx, double x_int)
{
    int i = 0;
    double a = 0, b = 0;
    if (x_int >= x[0] && x_int <= x[Nx - 1]) {
        while ((x_int < x[i] || x_int > x[i + 1]) && i < Nx - 2) i = i + 1;
        a = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
        b = y[i] - a * x[i];
    }
    if (x_int < x[0]) {
        a = (y[1] - y[0]) / (x[1] - x[0]);
        b = y[0] - a * x[0];
    }
    if (x_int > x[Nx - 1]) {
        a = (y[Nx - 1] - y[Nx - 2]) / (x[Nx - 1] - x[Nx - 2]);
        b = y[Nx - 2] - a * x[Nx - 2];
    }
    return a * x_int + b;
}

// linfit.cpp:17-35 with x = 0..n-1 (models.cpp:6065-6071), then eta0 = 3 pi / (rho G), rho = (Dnu/135.1)^2 rho_sun (models.cpp:6073-6084),
// with the four sums formed by a full warp (lane-strided partial sums, shuffle
// tree): the result differs from the serial order by rounding only (eta0 feeds nu_nlm, never a bin window).
__device__ __noinline__ double d_eta0_fct_warp(const double* fl0, int n_, int lane)
{
    double sx = 0, sy = 0;
    for (int i = lane; i < n_; i += 32) { sx += (double)i; sy += fl0[i]; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, d); sy += __shfl_xor_sync(0xffffffffu, sy, d); }
    const double n = (double)n_;
    const double mean_x = sx / n;
    double sty = 0, stt = 0;
    for (int i = lane; i < n_; i += 32) { const double t = (double)i - mean_x; sty += t * fl0[i]; stt += t * t; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { sty += __shfl_xor_sync(0xffffffffu, sty, d); stt += __shfl_xor_sync(0xffffffffu, stt, d); }
    (void)sy;
    const double Dnu_obs = sty / stt;
    const double G = 6.667e-8, Dnu_sun = 135.1, R_sun = 6.96342e5, M_sun = 1.98855e30;
    const double PI = 3.14159265358979323846;
    const double r5 = R_sun * 1e5;
    const double rho_sun = M_sun * 1e3 / (4 * PI * (r5 * r5 * r5) / 3);
    const double q = Dnu_obs / Dnu_sun;
    const double rho = (q * q) * rho_sun;
    return 3. * PI / (rho * G);
}

// build_lorentzian.cpp:595-649.  Non-exclusive ifs, last one wins.  Returns 0 on success.
__device__ __noinline__ int d_set_imin_imax(double x0, double xlast, int N, int l, double fc_l, double gamma_l,
                               double f_s, double c, double step, int* i0, int* i1)
{
    double p0 = nan(""), p1 = nan("");
    const double dl = (double)l;
    if (gamma_l >= 1 && f_s >= 1) {
        if (l != 0) { p0 = fc_l - c * (dl * f_s + gamma_l); p1 = fc_l + c * (dl * f_s + gamma_l); }
        else { p0 = fc_l - c * gamma_l * 2.2; p1 = fc_l + c * gamma_l * 2.2; }
    }
    if (gamma_l <= 1 && f_s >= 1) {
        if (l != 0) { p0 = fc_l - c * (dl * f_s + 1); p1 = fc_l + c * (dl * f_s + 1); }
        else { p0 = fc_l - c * 2.2; p1 = fc_l + c * 2.2; }
    }
    if (gamma_l >= 1 && f_s <= 1) {
        if (l != 0) { p0 = fc_l - c * (dl + gamma_l); p1 = fc_l + c * (dl + gamma_l); }
        else { p0 = fc_l - c * 2.2 * gamma_l; p1 = fc_l + c * 2.2 * gamma_l; }
    }
    if (gamma_l <= 1 && f_s <= 1) {
        if (l != 0) { p0 = fc_l - c * (dl + 1); p1 = fc_l + c * (dl + 1); }
        else { p0 = fc_l - c * 2.2; p1 = fc_l + c * 2.2; }
    }
    if ((p1 - step) < x0) p1 = x0 + c;
    if ((p0 + step) >= xlast) p0 = xlast - c;
    if (!(p0 == p0) || !(p1 == p1)) { *i0 = 0; *i1 = 0; return 1; }
    double f0 = floor((p0 - x0) / step);
    double f1 = ceil((p1 - x0) / step);
    f0 = fmin(fmax(f0, -2147483648.0), 2147483647.0);
    f1 = fmin(fmax(f1, -2147483648.0), 2147483647.0);
    int a = (int)f0, b = (int)f1;
    if (a < 0) a = 0;
    if (b > N) b = N;
    *i0 = a; *i1 = b;
    return (b - a <= 0) ? 1 : 0;
}

// per-chain quantities shared by all modes, staged in shared memory
struct Common {
    double ratios[4][7];   // amplitude_ratio(l, inc) or user ratios; [0][0] = 1
    double Vl[4];          // |V_l| (Vl[0] = 1)
    double eta0, trunc_c, asym;
    double a1, a3;         // Classic-type global splitting
    double a11, a12;       // a1l models
    double aterm[12];      // aj: [a1_0,a1_1,...,a6_0,a6_1]
    int do_amp;
    int status;
    int eta_from_fit;      // eta0 comes from the large-separation fit of warp 1 (phase 1b)
};

constexpr int EXP_THREADS = 512;                // 16 warps: the (mode, m) slot pass and the queue pass are latency-bound
constexpr int EXP_BATCH = 128;                 // modes expanded per pass
enum : unsigned char { SLOT_DEAD = 0, SLOT_FAST = 1, SLOT_WIDE = 2, SLOT_SLOW = 3, SLOT_NONFINITE = 4 };   // per (mode, m) slot
#ifndef TAMCMC_TILE_BASE_COST
#define TAMCMC_TILE_BASE_COST 26     // measured with the far-field folding on (trace build, C2): fixed phases of a tile / cost of one listed component
#endif
#ifndef TAMCMC_EDGE_COST
#define TAMCMC_EDGE_COST 1
#endif
constexpr int TILE_BASE_COST = TAMCMC_TILE_BASE_COST;             // per-tile fixed work in (component, bin)-pair units / 1024

// per-mode scratch between the three passes of a batch
struct ModeTmp {
    int l;
    int have;              // mode exists in this batch
    double fc, W, fsw;     // central frequency, width, splitting used by the window
    double f_s;            // splitting used by nu (a1etaa3 family)
    double H;              // common height (multiplied by the m-ratios), or < 0 when per-m heights are used
    double a[6];           // a1..a6 of this mode (aj family)
    double eta0;           // eta0 seen by this mode (eta_cm: take the chain's eta0, which warp 1 computes while pass A runs)
    int eta_cm;
    int hoff;              // model 13: offset of the per-m heights in the parameter vector
    int n;
    int i0, i1, bad;       // bit-exact window of set_imin_imax (pass A), bad != 0: imax - imin <= 0
    double sg;             // 2 / W, once per mode (2 * (1 / W) == 2 / W bit for bit: a scaling by two is exact)
};

// Called by ONE FULL WARP: lane k handles Harvey term k; live terms (tau != 0, noise_models.cpp:29) are compacted in
// term order with a ballot.  harvey_like(noise_params.array().abs(), ...) -- noise_models.cpp:15-39; models.cpp:2093-2100
__device__ __noinline__ void emit_noise(NoiseRec* out, const double* noise_params, int Nnoise, int Nharvey, int* status, int lane)
{
    NoiseRec& nr = *out;                     // filled in place (shared or global memory)
    if (Nharvey > TAMCMC_MAX_HARVEY) { if (lane == 0) atomicOr(status, TAMCMC_ST_BADCFG); Nharvey = TAMCMC_MAX_HARVEY; }
    double H = 0, tau = 0, pw = 0;
    if (lane < Nharvey) {
        H = fabs(noise_params[3 * lane]); tau = fabs(noise_params[3 * lane + 1]); pw = fabs(noise_params[3 * lane + 2]);
        if (!isfinite(H) || !isfinite(tau) || !isfinite(pw)) atomicOr(status, TAMCMC_ST_NONFINITE);
    }
    const bool live = (lane < Nharvey) && (tau != 0.0);
    const unsigned mk = __ballot_sync(0xffffffffu, live);
    const int nh = __popc(mk);
    if (live) {
        const int slot = __popc(mk & ((1u << lane) - 1u));
        nr.H[slot] = H;
        nr.lnsc[slot] = log((1e-3) * tau);
        nr.isc[slot] = (1e-3) * tau;
        nr.pw[slot] = pw;
        { double sn, cs; const double ang = 3.14159265358979323846 / fmax(pw, 1.0); sincos(ang, &sn, &cs); nr.spi[slot] = sn; nr.cpi[slot] = cs; }
        // generalised binomial coefficients C(p, q) = C(p, q-1) (p - q + 1) / q
        double b = 1.0;
        nr.binom[slot][0] = 1.0;
#pragma unroll
        for (int q = 1; q < TAMCMC_BG_TERMS; q++) { b = b * (pw - (double)(q - 1)) * (1.0 / (double)q); nr.binom[slot][q] = b; }
    }
    if (lane >= nh && lane < TAMCMC_MAX_HARVEY) { nr.H[lane] = 0; nr.lnsc[lane] = 0; nr.isc[lane] = 0; nr.pw[lane] = 0; nr.cpi[lane] = 0; nr.spi[lane] = 0; }
    if (lane == 0) {
        nr.nh = nh; nr.gauss = 0;
        nr.N0 = (Nnoise > 0) ? fabs(noise_params[Nnoise - 1]) : 0.0;
        if (!isfinite(nr.N0)) atomicOr(status, TAMCMC_ST_NONFINITE);
    }
}

// ------------------------------------------------------------------------------------------------
// Gaussian-envelope models (no Lorentzians): model_Harvey_Gaussian (models.cpp:5674-5725) and
// model_Kallinger2014_Gaussian (models.cpp:5728-5797).  Their background terms all have the shape H / (1 + (x/b)^c), so
// they ride the exact per-bin background path of the fused kernel; the Gaussian bump is one more per-bin term.
// ------------------------------------------------------------------------------------------------
// Kallinger+2014 super-Lorentzians of one chain (noise_models.cpp:115-126): amplitudes a_k, corner frequencies b_k, slopes c_k
struct KallingerPar { double a[3], b[3], c[3], N0; };

__device__ __noinline__ void kallinger_params(const double* p, KallingerPar& k)
{
    const double numax = fabs(p[15]), mu_numax = p[17];          // models.cpp:5742-5745
    k.a[0] = fabs(p[0] * pow(fabs(numax), p[1]));
    k.b[0] = fabs(p[2] * pow(fabs(numax + mu_numax), p[3]));
    k.c[0] = fabs(p[4]);
    k.a[1] = p[5];
    k.a[2] = p[6];
    k.b[1] = fabs(p[7] * pow(fabs(numax + mu_numax), p[8]));
    k.c[1] = fabs(p[9]);
    k.b[2] = fabs(p[10] * pow(fabs(numax + mu_numax), p[11]));
    k.c[2] = fabs(p[12]);
    k.N0 = fabs(p[13]);
}

// get_ksinorm (noise_models.cpp:65-84), first half: trapezoid sums of 1 / (1 + (x/b_k)^c_k) over the whole spectrum, one
// CTA per slice of TAMCMC_KSI_SLICE bins and chain, fixed reduction shape; the expander adds the slices in slice order.
// (x/b)^c = exp(c (ln x - ln b)) with ln x tabulated at create; x = 0 gives exp(-inf) = 0 like pow(0, c), c = 0 gives 1.
// 1/x to ~1 ulp: hardware approximation + two Newton steps (explicit FMAs: this file is compiled with -fmad=false)
__device__ __forceinline__ double rcp_newton(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

__global__ void __launch_bounds__(256) tamcmc_ksi_kernel(ExpandArgs A)
{
    const int sc = blockIdx.x, slice = blockIdx.y;
    const StarDesc& sd = A.stars[sc / A.Nchains];
    if (sd.model_id != TAMCMC_MODEL_ID_KALLINGER_GAUSS) return;
    if (A.active && !A.active[sc]) return;
    const int b0 = slice * A.ksi_slice_bins;
    if (b0 >= sd.Nloc) return;
    __shared__ double s_lnb[3], s_c[3], s_ib[3];
    __shared__ double s_red[8][3];
    if (threadIdx.x < 3) {
        // b_j = |k_j |numax + mu_numax|^s_j| (noise_models.cpp:117-121) through logarithms, one lane per term: the CTA waits for
        // this prologue, and three pow() + log() + divide in a row on one thread cost more than the slice's bins
        const double* p = A.params + (size_t)sc * A.params_stride;
        const int j = threadIdx.x, ik = (j == 0) ? 2 : (j == 1) ? 7 : 10;
        const double lnb = log(fabs(p[ik])) + p[ik + 1] * log(fabs(fabs(p[15]) + p[17]));
        s_lnb[j] = lnb; s_c[j] = fabs(p[ik + 2]); s_ib[j] = exp(-lnb);
    }
    __syncthreads();
    const double* lx = A.lnx + sd.off;
    const double* xs = A.x + sd.off;
    const int b1 = min(b0 + A.ksi_slice_bins, sd.Nloc);
    double acc[3] = {0.0, 0.0, 0.0};
    const bool ipow0 = (s_c[0] == 4.0 || s_c[0] == 2.0), ipow1 = (s_c[1] == 4.0 || s_c[1] == 2.0), ipow2 = (s_c[2] == 4.0 || s_c[2] == 2.0);
    const bool need_ln = !(ipow0 && ipow1 && ipow2);         // ln x is only read for non-integer slopes
    constexpr int U = 8;                                     // bins per thread in flight: the loads of a chunk are issued together
    for (int i0 = b0 + (int)threadIdx.x; i0 < b1; i0 += 256 * U) {
        double xv[U], lv[U];
#pragma unroll
        for (int q = 0; q < U; q++) {
            const int i = i0 + 256 * q;
            xv[q] = (i < b1) ? xs[i] : 0.0;
            lv[q] = (need_ln && i < b1) ? lx[i] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < U; q++) {
            const int i = i0 + 256 * q;
            const double w = (i >= b1) ? 0.0 : (i == 0 || i == sd.Nloc - 1) ? 0.5 : 1.0;      // no early exit: the 8 x 3 chains interleave
#pragma unroll
            for (int j = 0; j < 3; j++) {
                double z;
                if (s_c[j] == 4.0 || s_c[j] == 2.0) {          // the usual fixed slopes: plain products instead of exp
                    const double r = xv[q] * s_ib[j], r2 = r * r;
                    z = (s_c[j] == 4.0) ? r2 * r2 : r2;
                } else z = (s_c[j] == 0.0) ? 1.0 : exp(s_c[j] * (lv[q] - s_lnb[j]));
                acc[j] += (z < 1e300) ? w * rcp_newton(1.0 + z) : w / (1.0 + z);      // 1 + z in [1, 1e300]: the Newton form is exact enough; inf / NaN keep IEEE semantics
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], d);
    if ((threadIdx.x & 31) == 0) for (int j = 0; j < 3; j++) s_red[threadIdx.x >> 5][j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += s_red[w][threadIdx.x];
        A.ksi_part[((size_t)sc * A.ksi_slices + slice) * 3 + threadIdx.x] = t;
    }
}

// NoiseRec of the Kallinger2014 + Gaussian model (one warp; lanes 0..2 own one super-Lorentzian each):
//   ksi_k = b_k / (h * trapezoid sum)                      noise_models.cpp:80-82
//   term  = (ksi_k a_k^2 / b_k) / (1 + (x/b_k)^c_k)        noise_models.cpp:135-146
//   Gaussian |Amax| eta^2(x) exp(-0.5 (x - numax)^2 / sigma^2), eta = sinc leakage   models.cpp:5768-5772
// (the reference multiplies the noise by eta^2 into a discarded temporary, noise_models.cpp:148: only the Gaussian is attenuated)
__device__ __noinline__ void emit_kallinger(NoiseRec* out, const double* p, const double* ksi_part, int nslices, double h, double xnyq,
                                            int* status, int lane)
{
    NoiseRec& nr = *out;
    KallingerPar k;
    kallinger_params(p, k);
    if (lane < 3) {
        double sum = 0.0;
        for (int s = 0; s < nslices; s++) sum += ksi_part[3 * s + lane];
        const double ksi = k.b[lane] / (sum * h);
        const double H = (ksi * (k.a[lane] * k.a[lane])) / k.b[lane];
        nr.H[lane] = H;
        nr.lnsc[lane] = -log(k.b[lane]);
        nr.isc[lane] = 1.0 / k.b[lane];
        nr.pw[lane] = k.c[lane];
        nr.cpi[lane] = 0; nr.spi[lane] = 0;
        if (!isfinite(H) || !isfinite(nr.lnsc[lane]) || !isfinite(k.c[lane])) atomicOr(status, TAMCMC_ST_NONFINITE);
    } else if (lane < TAMCMC_MAX_HARVEY) { nr.H[lane] = 0; nr.lnsc[lane] = 0; nr.isc[lane] = 0; nr.pw[lane] = 0; nr.cpi[lane] = 0; nr.spi[lane] = 0; }
    if (lane == 0) {
        const double sig = fabs(p[16]);
        nr.nh = 3; nr.gauss = 2;
        nr.N0 = k.N0;
        nr.gH = fabs(p[14]); nr.gnu = fabs(p[15]); nr.gk = 0.5 / (sig * sig);
        nr.xnyq = xnyq;
        if (!isfinite(nr.N0) || !isfinite(nr.gH) || !isfinite(nr.gnu) || !isfinite(nr.gk)) atomicOr(status, TAMCMC_ST_NONFINITE);
    }
}

// NoiseRec of the Harvey + Gaussian model: params = [H1, tc1, p1, H2, tc2, p2, B0, Hgauss, nu_gauss, sigma] (models.cpp:5693-5701)
__device__ __noinline__ void emit_harvey_gauss(NoiseRec* out, const double* p, int* status, int lane)
{
    emit_noise(out, p, 7, 2, status, lane);
    if (lane == 0) {
        NoiseRec& nr = *out;
        const double sig = fabs(p[9]);
        nr.gauss = 1;
        nr.gH = fabs(p[7]); nr.gnu = p[8]; nr.gk = 0.5 / (sig * sig);
        nr.xnyq = 1.0;
        if (!isfinite(nr.gH) || !isfinite(nr.gnu) || !isfinite(nr.gk)) atomicOr(status, TAMCMC_ST_NONFINITE);
    }
}

// |params[n]/(pi*W)| (models.cpp:2032): long double in the reference, double here (<= 1 ulp apart)
__device__ __noinline__ double amp_to_height(double a, double W)
{
    const double PI = 3.141592653589793238462643383279502884;
    return fabs(a / (PI * W));
}

// nu_nlm of build_l_mode_a1etaa3 (build_lorentzian.cpp:143-145)
__device__ __forceinline__ double nu_a1etaa3(int l, int m, double fc, double f_s, double eta0, double a3)
{
    if (l == 0) return fc;
    const double fs6 = f_s * 1e-6;
    double v = fc * (1. + eta0 * (fs6 * fs6) * Qlm(l, m)) + (double)m * f_s;
    return __fma_rn(Phi(3, l, m), a3, v);   // long double product+sum in the reference: single rounding here
}

// nu_nlm of build_l_mode_aj (build_lorentzian.cpp:222-226)
__device__ __noinline__ double nu_aj(int l, int m, double fc, const double* a /*a1..a6*/, double eta0)
{
    if (l == 0) return fc;
    dd acc = dd_from(fc);
    for (int j = 1; j <= 6; j++)
        if (a[j - 1] != 0.0) acc = dd_add(acc, dd_mul_d_dd(a[j - 1], Phi(j, l, m), Plo(j, l, m)));
    double v = acc.hi;
    if (eta0 > 0) { const double a6_ = a[0] * 1e-6; v = v + fc * eta0 * Qlm(l, m) * (a6_ * a6_); }
    return v;
}

// Taylor series in u = x - xc of one Harvey-like term H / (1 + (tau' x)^p) (noise_models.cpp:30-31),
// valid on the tile when |u|/xc is small against the distance to the nearest singularity
// (x = 0 and (tau' x)^p = -1).  Returns false when the tile must use the per-bin exp() path.
// does the series of one term converge on the tile?  (*zc_out = (tau' xc)^p for the caller)
__device__ __noinline__ bool harvey_series_ok(double lnsc, double pw, double cpi, double spi, double xc, double lnxc, double umax, double* zc_out)
{
    if (!(xc > 0.0)) return false;
    const double L = lnsc + lnxc;
    const double tx = exp(L);            // tau' * xc
    const double zc = exp(pw * L);       // (tau' * xc)^p
    *zc_out = zc;
    if (!(zc < 1e30) || !(tx > 0.0)) return false;
    const double itx = 1.0 / tx;
    const double dr = cpi * itx - 1.0, di = spi * itx;
    const double rho = fmin(1.0, sqrt(dr * dr + di * di));
    return umax <= 0.04 * rho * xc;      // 0.04^NB = 1e-14 truncation
}

__device__ __noinline__ bool harvey_series(double H, double lnsc, double pw, double cpi, double spi, const double* binom,
                              double xc, double lnxc, double umax, double* out /*NB, accumulated*/)
{
    constexpr int NB = TAMCMC_BG_TERMS;
    double zc;
    if (!harvey_series_ok(lnsc, pw, cpi, spi, xc, lnxc, umax, &zc)) return false;
    double q[NB], g[NB];
    q[0] = 1.0 + zc;
#pragma unroll
    for (int k = 1; k < NB; k++) q[k] = zc * binom[k];
    g[0] = 1.0 / q[0];
#pragma unroll
    for (int n = 1; n < NB; n++) {
        double acc = 0.0;
#pragma unroll
        for (int k = 1; k <= n; k++) acc = fma(q[k], g[n - k], acc);
        g[n] = -g[0] * acc;
    }
    const double ixc = 1.0 / xc;
    double scl = H;
#pragma unroll
    for (int n = 0; n < NB; n++) { out[n] += g[n] * scl; scl *= ixc; }
    return true;
}

}  // namespace

// -------------------------------------------------------------------------------------------
// expand kernel: grid = nstars*Nchains CTAs, 128 threads.
// dynamic shared memory: params row [params_stride] doubles, then per-tile cost ints [max_tiles + 1]
// -------------------------------------------------------------------------------------------
// MODEL: the model id every star of the launch shares (the common ones are compiled on their own: the unpacking of the other
// families, a third of the kernel's code, drops out of the instruction stream of a launch that is bound by cold instruction
// fetches), or -1: read it from the star descriptor.
template <int MODEL>
__global__ void __launch_bounds__(EXP_THREADS, 1) tamcmc_expand_kernel(ExpandArgs A)
{
    extern __shared__ double sp[];             // this chain's parameter row, staged once
    // let the dependent fused kernel (launched with programmatic stream serialization) be scheduled right away: its
    // CTAs run their prologue and then block in griddepcontrol.wait until THIS grid has completed and flushed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int sc = blockIdx.x;                 // star*Nchains + chain
    const int star = sc / A.Nchains;
    // the first element of the parameter row this thread will stage: requested BEFORE the star descriptor is needed, so that the
    // two trips to memory overlap (both are cold: ~1 us each with L2 flushed)
    double p_first = 0.0;
    if (blockIdx.y == 0 && threadIdx.x >= 96 && (int)threadIdx.x - 96 < A.params_stride) p_first = A.params[(size_t)sc * A.params_stride + (threadIdx.x - 96)];
    const StarDesc sd = A.stars[star];
    ModeRec* modes = A.modes + (size_t)sc * A.modes_stride;
    CompRec* comps = A.comps + (size_t)sc * A.modes_stride * TAMCMC_MAX_COMP_PER_MODE;
    NoiseRec* noise = A.noise + sc;
    int* tcost = reinterpret_cast<int*>(sp + A.params_stride);   // [ntiles + 1] difference array -> cost
    int* tcover = tcost + (A.max_tiles + 2);                     // [ntiles + 1] difference array -> number of mode windows over the tile

    ETRACE(0);
    // ---- blockIdx.y >= 1: background CTAs.  One thread per tile of this chain builds the tile record
    // (origin, extent, Taylor series of the Harvey background); they run beside the mode CTAs. ----
    if (blockIdx.y > 0) {
        __shared__ NoiseRec s_nz;
        __shared__ int s_dummy;
        if (A.active && !A.active[sc]) return;
        const int* pl = sd.plength;
        const int bmodel = (MODEL == -1) ? sd.model_id : MODEL;
        const int o_noise = (bmodel == TAMCMC_MODEL_ID_MODE_TABLE) ? TAMCMC_MT_HDR : pl[0] + pl[1] + pl[2] + pl[3] + pl[4] + pl[5] + pl[6] + pl[7];
        if (threadIdx.x == 0) s_dummy = 0;
        __syncthreads();
        const bool kallinger = bmodel == TAMCMC_MODEL_ID_KALLINGER_GAUSS;      // always the exact per-bin background
        if (threadIdx.x < 32 && !kallinger) {
            if (bmodel == TAMCMC_MODEL_ID_HARVEY_GAUSS) emit_noise(&s_nz, A.params + (size_t)sc * A.params_stride, 7, 2, &s_dummy, threadIdx.x);
            else emit_noise(&s_nz, A.params + (size_t)sc * A.params_stride + o_noise, pl[8], (bmodel == 11 || bmodel == 14) ? 0 : (pl[8] - 1) / 3, &s_dummy, threadIdx.x);
        }
        __syncthreads();
        const int tile = (blockIdx.y - 1) * blockDim.x + threadIdx.x;
        if (tile >= sd.ntiles) return;
        const int lb0 = tile * sd.tile_bins;
        const int nvalid = min(sd.tile_bins, sd.Nloc - lb0);
        const double* xs = A.x + sd.off + lb0;
        TileRec tr;
        tr.xc = xs[nvalid >> 1];
        tr.umax = fmax(fabs(xs[0] - tr.xc), fabs(xs[sd.tile_bins - 1] - tr.xc));
        const double lnxc = A.lnx[sd.off + lb0 + (nvalid >> 1)];
        for (int k = 0; k < TAMCMC_BG_TERMS; k++) tr.bg[k] = 0.0;
        bool ok = !kallinger;
        for (int h = 0; ok && h < s_nz.nh; h++)
            ok = harvey_series(s_nz.H[h], s_nz.lnsc[h], s_nz.pw[h], s_nz.cpi[h], s_nz.spi[h], s_nz.binom[h], tr.xc, lnxc, tr.umax, tr.bg);
        TileRec* dst = A.tilerec + (size_t)sc * A.tiles_stride + tile;
        for (int k = 0; k < TAMCMC_BG_TERMS; k++) dst->bg[k] = tr.bg[k];
        dst->xc = tr.xc; dst->umax = tr.umax; dst->series_ok = ok ? 1 : 0;
        ETRACE(1);
        return;
    }

    __shared__ Common cm;
    __shared__ int s_status, s_status1b;     // s_status1b: raised by the scalar warps (phase 1b)
    __shared__ ModeTmp mt[EXP_BATCH];
    __shared__ int s_red[EXP_THREADS / 32], s_red2[EXP_THREADS / 32];
    __shared__ unsigned int s_bcnt[TAMCMC_NBUCKETS], s_bbase[TAMCMC_NBUCKETS], s_bgcnt, s_bgbase;
    static_assert(TAMCMC_NBUCKETS == (1 << TAMCMC_NBUCKETS_LOG2), "tile class is packed into the low bits");
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int* pl = sd.plength;
    const int Nmax = pl[0], lmax = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const int model = (MODEL == -1) ? sd.model_id : MODEL;
    const bool mode_table = (model == TAMCMC_MODEL_ID_MODE_TABLE);
    const int o_split = Nmax + lmax + Nf;
    const int o_width = o_split + Nsplit;
    const int o_noise = mode_table ? TAMCMC_MT_HDR : o_width + Nwidth;
    const int o_inc = o_noise + Nnoise;
    const int o_cfg = o_inc + Ninc;
    const int o_modes = TAMCMC_MT_HDR + Nnoise;      // mode table: first mode record
    const int ntiles = sd.ntiles;

    // Warps 0-2 ("scalar warps") do not stage anything: they compute the chain's m-height ratios, eta0 and the noise record --
    // long dependent FP64 chains (sincos, logarithms, divides; cold code) -- straight from the GLOBAL parameter row, starting at
    // once, while warps 3-15 stage the row in shared memory, file the scalars (phase 1a) and run pass A of the first batch.  The two
    // groups meet at the barrier in front of the per-component pass.
    const bool scalar_warp = warp < 3;
    const bool inactive = A.active && !A.active[sc];
    const double* gp = A.params + (size_t)sc * A.params_stride;
    auto stage_sync = []() { asm volatile("bar.sync 1, %0;" ::"n"(EXP_THREADS - 96) : "memory"); };      // warps 3..15 only
    const double* params = sp;
    const double* fl0_all = params + Nmax + lmax;
    const double* Wl0_all = params + o_width;
    const int nmodes = sd.nmodes_cap;
    if (scalar_warp) {
        if (tid == 64) s_status1b = 0;
        __syncwarp();
        if (!inactive) {
            if (tid < 15) {
                // warp 0, lanes 0..14: the 3+5+7 entries of amplitude_ratio(l, inc), l = 1..3 (or the ratio parameters)
                const int l = (tid < 3) ? 1 : (tid < 8) ? 2 : 3;
                const int i = tid - ((l == 1) ? 0 : (l == 2) ? 3 : 8) - l;       // m = -l..l
                const bool have = mode_table ? true : (model == 11) ? (pl[2 + l] >= 1) : (model == 14) ? false : (lmax >= l);
                if (have) {
                    if (mode_table) {
                        cm.ratios[l][i + l] = d_amplitude_ratio_entry(l, i, gp[1]);
                    } else if (model == 12) {
                        // models.cpp:2196-2214: m-ratios read from the "inclination" block, symmetric in m
                        const int base = (l == 1) ? 0 : (l == 2) ? 2 : 5;
                        cm.ratios[l][i + l] = fabs(gp[o_inc + base + (i < 0 ? -i : i)]);
                    } else if (model != 13) {
                        double inc;
                        if (model == 11) {       // models.cpp:3057-3058
                            const double PI = 3.141592653589793238462643383279502884;
                            inc = atan(gp[o_split + 4] / gp[o_split + 3]) * 180. / PI;
                        } else inc = gp[o_inc];
                        cm.ratios[l][i + l] = d_amplitude_ratio_entry(l, i, inc);
                    }
                }
                ETRACE_T(8, 14);
            } else if (tid >= 32 && tid < 64) {
                // warp 1: eta0 by the whole warp (models that use it), then lane 0 files the scalar parameters
                const bool need_eta = (model == 3 || model == 12 || model == 13 || model == 6 || model == 7 || model == 8) ||
                                      (model == 23 && gp[o_split + 12] == 1);
                const double eta_w = need_eta ? d_eta0_fct_warp(gp + Nmax + lmax, Nfl0, tid - 32) : 0.0;
                if (tid == 32 && need_eta) cm.eta0 = eta_w;          // need_eta == cm.eta_from_fit of phase 1a (same rule)
                ETRACE_T(9, 32);
            } else if (tid >= 64 && tid < 96) {
                // warp 2: Harvey-like background parameters, one lane per term
                if (model == TAMCMC_MODEL_ID_KALLINGER_GAUSS)
                    emit_kallinger(noise, gp, A.ksi_part + (size_t)sc * A.ksi_slices * 3, (sd.Nloc + A.ksi_slice_bins - 1) / A.ksi_slice_bins,
                                   sd.step, sd.xlast, &s_status1b, tid - 64);
                else if (model == TAMCMC_MODEL_ID_HARVEY_GAUSS) emit_harvey_gauss(noise, gp, &s_status1b, tid - 64);
                else emit_noise(noise, gp + o_noise, Nnoise, (model == 11 || model == 14) ? 0 : (Nnoise - 1) / 3, &s_status1b, tid - 64);
                ETRACE_T(10, 64);
            }
        }
    } else {
        if (tid == 96) s_status = inactive ? TAMCMC_ST_INACTIVE : 0;
        for (int k = tid - 96 + (EXP_THREADS - 96); k < A.params_stride; k += EXP_THREADS - 96) sp[k] = gp[k];
        if (tid - 96 < A.params_stride) sp[tid - 96] = p_first;
        for (int k = tid - 96; k <= ntiles; k += EXP_THREADS - 96) { tcost[k] = 0; tcover[k] = 0; }
        if (A.bgqueue && sd.nmodes_cap > 0) {
            // phase 3 looks at the centre and the ends of every tile no mode touches (does its background series converge?): start those
            // lines' way into L2 now, so that the look costs an L2 hit at the end of the kernel instead of a trip to HBM
            for (int t = tid - 96; t < ntiles; t += EXP_THREADS - 96) {
                const int lb0 = t * sd.tile_bins, nvalid = min(sd.tile_bins, sd.Nloc - lb0);
                const double* xs = A.x + sd.off + lb0;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(xs + (nvalid >> 1)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(xs));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(xs + sd.tile_bins - 1));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(A.lnx + sd.off + lb0 + (nvalid >> 1)));
            }
        }
        stage_sync();
        ETRACE_T(1, 96);
        // ---------------- phase 1a: the scalars pass A needs (one thread; everything here is a parameter read) ----------------
        if (!inactive) {
            if (tid >= 97 && tid <= 99) {
                const int l = tid - 96;
                const bool have = mode_table ? false : (model == 11 || model == 14) ? false : (lmax >= l);
                cm.Vl[l] = have ? fabs(params[Nmax + l - 1]) : 1.0;
            } else if (tid == 96) {
                cm.eta_from_fit = 0;
                cm.trunc_c = mode_table ? params[2] : params[o_cfg];
                cm.do_amp = mode_table ? 0 : (params[o_cfg + 1] != 0.0);
                cm.ratios[0][0] = 1.0;
                cm.Vl[0] = 1.0;
                cm.status = 0;
                double eta_other = 0.0;          // eta0 of the models that do not take it from the large-separation fit
                cm.a1 = cm.a3 = cm.a11 = cm.a12 = 0.0;
                switch (model) {
                case 3: case 12: case 13:    // models.cpp:2011-2016, 2219-2222, 2396-2399
                    cm.a1 = fabs(params[o_split]);
                    cm.eta_from_fit = 1;
                    cm.a3 = params[o_split + 2];
                    cm.asym = params[o_split + 5];
                    break;
                case 6:                      // models.cpp:87-91
                    cm.a11 = fabs(params[o_split]);
                    cm.a12 = fabs(params[o_split + 6]);
                    cm.eta_from_fit = 1;
                    cm.a3 = params[o_split + 2];
                    cm.asym = params[o_split + 5];
                    break;
                case 7: case 8:              // a1n / a1nl etaa3: models.cpp:290-296, 1075-1081 (splittings per radial order: pass A)
                    cm.eta_from_fit = 1;
                    cm.a3 = params[o_split + 2];
                    cm.asym = params[o_split + 5];
                    break;
                case 14:                     // model_MS_local_Hnlm: models.cpp:3242-3245
                    cm.a1 = fabs(params[o_split]);
                    eta_other = params[o_split + 1];
                    cm.a3 = params[o_split + 2];
                    cm.asym = params[o_split + 5];
                    break;
                case 11:                     // models.cpp:3059-3075 (Nvis plays the role of lmax)
                    cm.a1 = params[o_split + 3] * params[o_split + 3] + params[o_split + 4] * params[o_split + 4];
                    eta_other = params[o_split + 1];
                    cm.a3 = params[o_split + 2];
                    cm.asym = params[o_split + 5];
                    break;
                case 23:                     // models.cpp:1257-1270
                    for (int k = 0; k < 12; k++) cm.aterm[k] = params[o_split + k];
                    cm.asym = params[o_split + 13];
                    cm.eta_from_fit = (params[o_split + 12] == 1) ? 1 : 0;
                    break;
                case TAMCMC_MODEL_ID_KALLINGER_GAUSS: case TAMCMC_MODEL_ID_HARVEY_GAUSS:   // no modes: background + Gaussian envelope
                    cm.asym = 0.0;
                    break;
                case TAMCMC_MODEL_ID_MODE_TABLE:   // modes already resolved by a host expander (e.g. models.cpp:4788-4911)
                    cm.asym = params[3];
                    if (!(params[0] >= 0.0 && params[0] <= (double)sd.nmodes_cap)) cm.status = TAMCMC_ST_BADCFG;
                    break;
                default:
                    cm.status = TAMCMC_ST_BADCFG;
                    cm.asym = 0;
                    break;
                }
                if (!cm.eta_from_fit) cm.eta0 = eta_other;      // (the fit's value is written by warp 1: one writer either way)
                if (cm.status) atomicOr(&s_status, cm.status);
            }
        }
        stage_sync();
    }
    ETRACE_T(2, 96);
    // ---------------- phase 2: modes in batches of 128: pass A (one thread per mode), then ONE pass over the (mode, m) slots ----------------
    // (every thread must meet the same barriers: `inactive` comes from global memory and is the same for all of them; what the
    // two groups found -- unknown model, bad mode count, too many Harvey terms -- is looked at after the first barrier)
    bool run = !inactive;
    for (int base = 0; run && base < nmodes; base += EXP_BATCH) {
        if (!scalar_warp || base > 0) {
            const bool okA = !((s_status | (base > 0 ? s_status1b : 0)) & TAMCMC_ST_BADCFG);
            const int nmodes_live = !okA ? 0 : mode_table ? (int)params[0] : nmodes;     // mode table: per-chain mode count
            // ---- pass A: one thread per mode: degree, frequency, width, height rule, splittings ----
            {
                // threads 128..255 (warps 4-7): warps 0-2 are still busy with phase 1b during the first batch
                const int ta = tid - EXP_BATCH;
                const bool mine = ta >= 0 && ta < EXP_BATCH;
                const int j = base + ta;
                static_assert(2 * EXP_BATCH <= EXP_THREADS, "one scratch slot per mode of a batch, on warps 4..7");
                ModeTmp& t = mt[mine ? ta : 0];      // filled in place in shared memory (the other threads idle here)
                if (mine) t.have = 0;
                if (mode_table && mine && j < nmodes_live) {
                    // one optimum_lorentzian_calc_aj call of the host model function (e.g. models.cpp:4937, 4952, 4977, 5001)
                    const double* r = params + o_modes + TAMCMC_MT_STRIDE * j;
                    const int l = (int)r[0];
                    if (!(r[0] >= 0.0 && r[0] <= 3.0) || !(r[2] >= 0.0) || !(r[3] >= 0.0)) atomicOr(&s_status, (r[0] == r[0] && r[2] == r[2] && r[3] == r[3]) ? TAMCMC_ST_BADCFG : TAMCMC_ST_NONFINITE);
                    else {
                        t.have = 1; t.l = l; t.n = j; t.fc = r[1]; t.H = r[2]; t.W = r[3];
                        for (int k = 0; k < 6; k++) t.a[k] = r[4 + k];
                        t.eta0 = r[10]; t.eta_cm = 0; t.fsw = r[4]; t.f_s = 0.0;
                        t.hoff = o_modes + TAMCMC_MT_STRIDE * j + 11 + 3;     // extra[m] = params[hoff + m]
                    }
                } else if (!mode_table && mine && j < nmodes_live) {
                    int l, n;
                    if (model == 3 || model == 12 || model == 13 || model == 6 || model == 7 || model == 8) {
                        l = j % (lmax + 1); n = j / (lmax + 1);          // n-major, l interleaved (models.cpp:2026-2085)
                    } else {
                        // l-major: all l=0, then l=1, ... (models.cpp:1287-1376, 3082-3134)
                        int r = j; l = 0;
                        while (l < 3 && r >= pl[2 + l]) { r -= pl[2 + l]; l++; }
                        n = r;
                    }
                    const int o_fl = Nmax + lmax + (l >= 1 ? Nfl0 : 0) + (l >= 2 ? Nfl1 : 0) + (l >= 3 ? Nfl2 : 0);
                    const double fc = params[o_fl + n];
                    t.have = 1; t.l = l; t.n = n; t.fc = fc; t.hoff = -1; t.eta0 = 0.0; t.eta_cm = 1;     // cm.eta0 is read in pass B
                    for (int k = 0; k < 6; k++) t.a[k] = 0.0;
                    if (model == 14) {
                        // models.cpp:3256-3318: individual widths; heights H(n,l,|m|) at params[base_l + (l+1) n + |m|] with
                        // base_l = Nfl0, Nfl0+Nfl1, Nfl0+Nfl1+Nfl2 exactly as the reference indexes them (:3270, 3285, 3303)
                        const int idx = (l >= 1 ? Nfl0 : 0) + (l >= 2 ? Nfl1 : 0) + (l >= 3 ? Nfl2 : 0) + n;
                        t.W = fabs(params[o_width + idx]);
                        t.H = -1.0;
                        t.hoff = idx - n + (l + 1) * n;
                        t.f_s = cm.a1; t.fsw = cm.a1;
                    } else if (model == 11) {
                        // models.cpp:3082-3134: individual heights and widths per mode
                        const int idx = (l >= 1 ? Nfl0 : 0) + (l >= 2 ? Nfl1 : 0) + (l >= 3 ? Nfl2 : 0) + n;
                        t.W = fabs(params[o_width + idx]);
                        t.H = cm.do_amp ? amp_to_height(params[idx], t.W) : fabs(params[idx]);
                        t.f_s = cm.a1; t.fsw = cm.a1;
                    } else if (model == 23) {
                        // models.cpp:1287-1376
                        if (l == 0) {
                            t.W = fabs(Wl0_all[n]);
                            t.H = cm.do_amp ? amp_to_height(params[n], t.W) : fabs(params[n]);
                            t.eta0 = 0.0; t.eta_cm = 0;
                        } else {
                            t.W = fabs(d_lin_interpol(fl0_all, Wl0_all, Nmax, fc));
                            const double Hi = d_lin_interpol(fl0_all, params, Nmax, fc);
                            const double PI = 3.14159265358979323846;
                            t.H = cm.do_amp ? fabs(Hi / (PI * t.W) * cm.Vl[l]) : fabs(Hi * cm.Vl[l]);
                            const int na = 2 * l;            // l=1: a1,a2; l=2: a1..a4; l=3: a1..a6
                            for (int k = 0; k < na; k++) t.a[k] = cm.aterm[2 * k] + cm.aterm[2 * k + 1] * (fc * 1e-3);
                        }
                        t.f_s = 0.0;
                        t.fsw = t.a[0];                       // optimum_lorentzian_calc_aj: window uses a1 (build_lorentzian.cpp:513)
                    } else {
                        // Classic family and a1l: widths interpolated on the l=0 ladder, heights H[n]*V_l
                        t.W = (l == 0) ? fabs(Wl0_all[n]) : fabs(d_lin_interpol(fl0_all, Wl0_all, Nmax, fc));
                        if (model == 6) {                     // build_lorentzian.cpp:58-66, 383-396
                            t.f_s = (l == 0) ? 0.0 : (l == 1) ? cm.a11 : (l == 2) ? cm.a12 : (cm.a11 + cm.a12) / 2.;
                        } else if (model == 7 || model == 8) {
                            // splittings per radial order (models.cpp:290-291, 317-318; 1075-1076, 1099-1100)
                            const double a11 = fabs(params[o_split + 6 + n]);
                            const double a12 = (model == 8) ? fabs(params[o_split + 6 + Nmax + n]) : a11;
                            t.f_s = (l == 0) ? 0.0 : (l == 1) ? a11 : (l == 2) ? a12 : (a11 + a12) / 2.;
                        } else t.f_s = cm.a1;
                        t.fsw = t.f_s;
                        if (model == 13) {
                            // models.cpp:2409-2470: per-m heights from the parameter vector, |H|/(pi W) if do_amp
                            t.H = -1.0;
                            t.hoff = (l == 0) ? n : o_inc + (l + 1) * n;
                        } else {
                            if (l == 0) t.H = cm.do_amp ? amp_to_height(params[n], t.W) : fabs(params[n]);
                            else t.H = cm.do_amp ? amp_to_height(params[n], t.W) * cm.Vl[l] : fabs(params[n] * cm.Vl[l]);
                        }
                    }
                }
                ETRACE_T(11, EXP_BATCH + 5);
                // bit-exact window (build_lorentzian.cpp:595-649), one thread per mode
                if (mine && t.have) {
                    t.bad = d_set_imin_imax(sd.x0, sd.xlast, sd.Nglob, t.l, t.fc, t.W, t.fsw, cm.trunc_c, sd.step, &t.i0, &t.i1);
                    t.sg = 2.0 / t.W;
                }
            }
        }
        __syncthreads();
        if (base == 0) {
            // the two groups meet: from here on everybody sees the staged row, cm and both status words
            run = !((s_status | s_status1b) & TAMCMC_ST_BADCFG);
            if (!run) break;
        }
        ETRACE(3);
        // ---- per-component pass: 8 lanes per mode (lane k <-> m = k - l, lane 7 idle), so that a mode's records, its component
        // order (FAST first), the extent of its centres and its tile costs come out of ONE pass with warp votes instead of a second,
        // serial per-mode pass behind a barrier ----
        const double inv_step = 1.0 / sd.step, inv_T = 1.0 / (double)sd.tile_bins;       // scheduling weights only (never a window)
#pragma unroll 2
        for (int sl = tid; sl < EXP_BATCH * 8; sl += EXP_THREADS) {
            const int jj = sl >> 3, k = sl & 7;
            const ModeTmp& t = mt[jj];
            const int l = t.l, m = k - l;
            const bool live = t.have && k <= 2 * l;
            double nu = 0.0, h = 0.0, cs = 0.0, ia = 0.0;
            unsigned char cls = SLOT_DEAD;
            if (live) {
                const double eta0 = t.eta_cm ? cm.eta0 : t.eta0;
                if (model == 23) nu = nu_aj(l, m, t.fc, t.a, eta0);
                else if (mode_table) { nu = nu_aj(l, m, t.fc, t.a, eta0); if (l != 0) nu = nu + params[t.hoff + m]; }   // build_lorentzian.cpp:182-190
                else nu = nu_a1etaa3(l, m, t.fc, t.f_s, eta0, cm.a3);
                if (t.H >= 0.0) h = t.H * cm.ratios[l][k];
                else {
                    const double PI = 3.141592653589793238462643383279502884;
                    h = (l == 0) ? params[t.hoff] : params[t.hoff + (m < 0 ? -m : m)];
                    if (cm.do_amp) h = h / (PI * t.W);
                    h = fabs(h);
                }
                cs = t.sg / sqrt(h); ia = 1.0 / h;                   // scaled FAST form (used if the slot qualifies)
                // classification against the mode's window.  FAST: t' = (1+e^2)/A stays inside [1e-8, 1e16] over the whole
                // window, so 16 merges between two exponent renormalisations cannot leave the FP64 range.  WIDE (very narrow
                // modes, e.g. red-giant mixed modes far narrower than a bin): t' inside [1e-16, 1e32], same scaled form, but the
                // segments that hold such a mode renormalise every 4 merges.  Everything else live: general (SLOW) form.
                if (!isfinite(h) || !isfinite(nu)) cls = SLOT_NONFINITE;
                else if (h != 0.0 && t.W > 0.0) {               // else: contributes exactly 0 (gamma == 0: DESIGN.md deviations)
                    const double xlo = sd.x0 + (double)t.i0 * sd.step, xhi = sd.x0 + (double)t.i1 * sd.step;
                    const double emax = t.sg * fmax(fabs(xlo - nu), fabs(xhi - nu));
                    if (!(emax < 1e100)) cls = SLOT_NONFINITE;
                    else if (h >= 1e-8 && h <= 1e8 && emax < 1e4) cls = SLOT_FAST;
                    else if (h >= 1e-16 && h <= 1e16 && emax < 1e8) cls = SLOT_WIDE;
                    else cls = SLOT_SLOW;
                }
            }
            // votes of the mode's 8 lanes
            const int gsh = (lane >> 3) << 3;
            const bool isfast = (cls == SLOT_FAST || cls == SLOT_WIDE);
            const unsigned fastm = (__ballot_sync(0xffffffffu, isfast) >> gsh) & 0xffu;
            const unsigned slowm = (__ballot_sync(0xffffffffu, cls == SLOT_SLOW) >> gsh) & 0xffu;
            const unsigned widem = (__ballot_sync(0xffffffffu, cls == SLOT_WIDE) >> gsh) & 0xffu;
            const unsigned nonfm = (__ballot_sync(0xffffffffu, cls == SLOT_NONFINITE) >> gsh) & 0xffu;
            double numin = isfast ? nu : 1e300, numax = isfast ? nu : -1e300;          // extent of the FAST components' centres
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) { numin = fmin(numin, __shfl_xor_sync(0xffffffffu, numin, d)); numax = fmax(numax, __shfl_xor_sync(0xffffffffu, numax, d)); }
            const int j = base + jj;
            const int nf = __popc(fastm), ns = __popc(slowm);
            if (isfast || cls == SLOT_SLOW) {
                // FAST/WIDE components first (in m order), then the SLOW ones
                CompRec cr;
                cr.nu = nu; cr.m = m;
                if (isfast) { cr.flags = TAMCMC_CF_FAST; cr.s = cs; cr.a = ia; }
                else { cr.flags = TAMCMC_CF_SLOW; cr.s = t.sg; cr.a = h; }
                const int pos = isfast ? __popc(fastm & ((1u << k) - 1u)) : nf + __popc(slowm & ((1u << k) - 1u));
                comps[(size_t)j * TAMCMC_MAX_COMP_PER_MODE + pos] = cr;
            }
            if (k == 0 && !t.have && j < nmodes) {
                // unused slot of a mode table (or a rejected record): an empty record, so stale data is never listed
                ModeRec mr;
                mr.i0 = 0; mr.i1 = 0; mr.ncomp = 0; mr.nfast = 0; mr.l = 0; mr.pad = 0;
                mr.numin = 1.0; mr.numax = 0.0;
                mr.fc = 0; mr.gamma = 0; mr.qa = 0; mr.qb0 = 1; mr.qc = 0;
                modes[j] = mr;
            }
            if (k == 0 && t.have) {
                ModeRec mr;
                const int i0 = t.i0, i1 = t.i1, bad = t.bad;
                if (bad) atomicOr(&s_status, TAMCMC_ST_WINDOW);
                if (nonfm) atomicOr(&s_status, TAMCMC_ST_NONFINITE);
                mr.i0 = i0; mr.i1 = i1; mr.l = l; mr.fc = t.fc; mr.gamma = t.W;
                if (cm.asym != 0.0) {
                    mr.qa = cm.asym / t.fc;
                    const double k2 = 0.5 * t.W * cm.asym / t.fc;
                    mr.qc = k2 * k2;
                } else { mr.qa = 0.0; mr.qc = 0.0; }      // symmetric profiles: the fused kernel never reads q(x) of such a chain (asym_flag == 0)
                mr.qb0 = 1.0 - cm.asym;
                mr.pad = 0;
                if (!isfinite(t.fc) || !isfinite(t.W) || (cm.asym != 0.0 && (!isfinite(mr.qa) || !isfinite(mr.qc))))
                    atomicOr(&s_status, TAMCMC_ST_NONFINITE);
                const int wide = widem ? 1 : 0;
                mr.nfast = nf | (wide << 16); mr.ncomp = nf + ns;      // bit 16: the mode has WIDE fast components
                // numin > numax: the mode is never folded into a tile's far-field polynomial (no FAST components or WIDE dynamic range)
                const bool far_capable = nf > 0 && !wide;
                mr.numin = far_capable ? numin : 1.0; mr.numax = far_capable ? numax : 0.0;
                modes[j] = mr;
                // per-tile cost: difference array over the LOCAL tiles this window touches
                if (!bad && mr.ncomp > 0) {
                    const int lo = max(i0, sd.bin0) - sd.bin0, hi = min(i1, sd.bin0 + sd.Nloc) - sd.bin0;
                    if (hi > lo) {
                        int t0 = lo / sd.tile_bins, t1 = (hi - 1) / sd.tile_bins + 1;
                        atomicAdd(&tcover[t0], 1); atomicAdd(&tcover[t1], -1);      // every tile the window touches, near or far
                        if (far_capable && A.far_ratio > 0.0 && ns == 0) {
                            // the fused kernel merges this mode per bin only in the tiles whose centre lies within far_ratio half
                            // tiles of its components; elsewhere it costs nothing per bin (a scheduling weight, not a result)
                            const double T = (double)sd.tile_bins, Rb = A.far_ratio * 0.5 * T + 0.5 * T;
                            const double bmin = (numin - sd.x0) * inv_step - (double)sd.bin0, bmax = (numax - sd.x0) * inv_step - (double)sd.bin0;
                            const double tl = floor((bmin - Rb) * inv_T), th = ceil((bmax + Rb) * inv_T) + 1.0;
                            if (tl > (double)t0) t0 = (int)fmin(tl, (double)t1);
                            if (th < (double)t1) t1 = (int)fmax(th, (double)t0);
                        }
                        if (t1 > t0) {
                            atomicAdd(&tcost[t0], mr.ncomp);
                            atomicAdd(&tcost[t1], -mr.ncomp);
                        }
#if TAMCMC_EDGE_COST > 0
                        // a tile that holds a window edge merges the mode under masks (general entries): about twice a plain merge
                        {
                            const int e0 = lo / sd.tile_bins, e1 = (hi - 1) / sd.tile_bins;
                            if (lo % sd.tile_bins) { atomicAdd(&tcost[e0], TAMCMC_EDGE_COST * mr.ncomp); atomicAdd(&tcost[e0 + 1], -TAMCMC_EDGE_COST * mr.ncomp); }
                            if (hi % sd.tile_bins) { atomicAdd(&tcost[e1], TAMCMC_EDGE_COST * mr.ncomp); atomicAdd(&tcost[e1 + 1], -TAMCMC_EDGE_COST * mr.ncomp); }
                        }
#endif
                    }
                }
            }
        }
        ETRACE(4);
        __syncthreads();
    }

    __syncthreads();       // phase 1b has no barrier of its own when no batch ran (envelope models, masked chains)
    ETRACE(5);
    // ---------------- phase 3: tile costs -> heavy-first work queue ----------------
    const int status_all = s_status | s_status1b;
    run = !inactive && !(status_all & TAMCMC_ST_BADCFG);
    const bool enqueue = run && (status_all == 0);
    if (enqueue) {
        // inclusive scan of the difference array (serial per warp-strided chunk would need carries; ntiles is
        // a few hundred: one warp scans it in 32-wide steps with a running carry)
        if (warp == 0 || warp == 1) {
            // warp 0: cost of every tile; warp 1: how many mode windows lie over it (0 = a background-only tile)
            int* arr = (warp == 0) ? tcost : tcover;
            const int add = (warp == 0) ? TILE_BASE_COST : 0;
            int carry = 0;
            for (int t0 = 0; t0 < ntiles; t0 += 32) {
                const int t = t0 + lane;
                int v = (t < ntiles) ? arr[t] : 0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += o; }
                v += carry;
                if (t < ntiles) arr[t] = v + add;
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
        }
        __syncthreads();
        int mx = 0, nfree = 0;                      // heaviest tile; tiles no mode window touches
        for (int t = tid; t < ntiles; t += blockDim.x) { mx = max(mx, tcost[t]); nfree += (tcover[t] == 0); }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d)); nfree += __shfl_down_sync(0xffffffffu, nfree, d); }
        if (lane == 0) { s_red[warp] = mx; s_red2[warp] = nfree; }
        __syncthreads();
        mx = 0; nfree = 0;
        for (int w = 0; w < EXP_THREADS / 32; w++) { mx = max(mx, s_red[w]); nfree += s_red2[w]; }
        // cost class of every tile (sixteenths of the chain's heaviest tile, heaviest first) and its rank inside the class,
        // counted in shared memory; then ONE global atomic per class reserves the chain's slots in the queue
        if (tid < TAMCMC_NBUCKETS) s_bcnt[tid] = 0u;
        if (tid == TAMCMC_NBUCKETS) s_bgcnt = 0u;
        __syncthreads();
        // Background-only tiles (no window over them; the Lorentzian models with the chi(2,2p) likelihood) leave the ring: they go to
        // the second queue, which the fused kernel's consumer warps drain one tile per warp (whittle.cu, bg_phase).  Class
        // TAMCMC_NBUCKETS marks them below.
        // ... when they are at least a quarter of the chain's tiles: below that the ring absorbs them for less than the series test of
        // every such tile costs at the end of this kernel (C2: 14 of 163 tiles, +1.8 us on the expander for -0.5 us on the fused kernel)
        const bool bg_fast = A.bgqueue != nullptr && nmodes > 0 && 4 * nfree >= ntiles;
        const int rounds = (ntiles + blockDim.x - 1) / blockDim.x;
        for (int r = 0; r < rounds; r++) {
            const int t = r * blockDim.x + tid;
            if (t < ntiles) {
                bool bg = bg_fast && tcover[t] == 0;
                if (bg) {
                    // only tiles whose background SERIES converges take the warp-per-tile path (the exact per-bin terms near x = 0
                    // cost several times more per bin: one such tile on one warp would be the tail of the whole phase).  Same test
                    // on the same inputs as the background CTAs make for the tile record (blockIdx.y >= 1 above).
                    const int lb0 = t * sd.tile_bins, nvalid = min(sd.tile_bins, sd.Nloc - lb0);
                    const double* xs = A.x + sd.off + lb0;
                    const double xc = xs[nvalid >> 1];
                    const double umax = fmax(fabs(xs[0] - xc), fabs(xs[sd.tile_bins - 1] - xc));
                    const double lnxc = A.lnx[sd.off + lb0 + (nvalid >> 1)];
                    for (int h = 0; bg && h < noise->nh; h++) { double zc; bg = harvey_series_ok(noise->lnsc[h], noise->pw[h], noise->cpi[h], noise->spi[h], xc, lnxc, umax, &zc); }
                }
                if (bg) {
                    const unsigned rank = atomicAdd(&s_bgcnt, 1u);
                    tcost[t] = (int)((rank << (TAMCMC_NBUCKETS_LOG2 + 1)) | (unsigned)TAMCMC_NBUCKETS);
                } else {
                    const int q = (TAMCMC_NBUCKETS * (tcost[t] - TILE_BASE_COST)) / max(mx - TILE_BASE_COST, 1);
                    const int cls = min(max(TAMCMC_NBUCKETS - 1 - q, 0), TAMCMC_NBUCKETS - 1);
                    const unsigned rank = atomicAdd(&s_bcnt[cls], 1u);
                    tcost[t] = (int)((rank << (TAMCMC_NBUCKETS_LOG2 + 1)) | (unsigned)cls);      // ntiles <= 16384: rank fits
                }
            }
        }
        __syncthreads();
        if (tid < TAMCMC_NBUCKETS) s_bbase[tid] = s_bcnt[tid] ? atomicAdd(&A.qctl->count[tid], s_bcnt[tid]) : 0u;
        if (tid == TAMCMC_NBUCKETS) s_bgbase = s_bgcnt ? atomicAdd(&A.qctl->bg_count, s_bgcnt) : 0u;
        __syncthreads();
        for (int r = 0; r < rounds; r++) {
            const int t = r * blockDim.x + tid;
            if (t < ntiles) {
                const unsigned v = (unsigned)tcost[t], cls = v & (2u * TAMCMC_NBUCKETS - 1u), rank = v >> (TAMCMC_NBUCKETS_LOG2 + 1);
                const unsigned item = (unsigned)sc * (unsigned)A.tiles_stride + (unsigned)t;
                if (cls == TAMCMC_NBUCKETS) A.bgqueue[s_bgbase + rank] = item;
                else A.queue[(size_t)cls * A.qcap + s_bbase[cls] + rank] = item | ((A.mark_bgonly && tcover[t] == 0) ? 0x80000000u : 0u);
            }
        }
    }
    __syncthreads();
    ETRACE(6);
    if (tid == 0) {
        A.status[sc] = status_all;
        A.asym_flag[sc] = (!inactive && cm.asym != 0.0) ? 1 : 0;
        if (status_all != 0) A.out_logL[sc] = nan("");
    }
}

// Parallel-tempering swap of two adjacent chains on the device (MALA.cpp:397-461): one CTA; thread 0 decides, all threads
// exchange the two parameter rows.
__global__ void __launch_bounds__(256) tamcmc_pt_swap_kernel(double* params, double* logL, double* logPrior, const double* T, int stride,
                                                             int A, double u, int* swapped)
{
    __shared__ int s_do;
    const int B = A + 1;
    if (threadIdx.x == 0) {
        const double LA = logL[A], LB = logL[B];
        const double LA_TB = LA * T[A] / T[B], LB_TA = LB * T[B] / T[A];
        double r = exp(LA_TB + LB_TA - LA - LB);
        if (r > 1.0) r = 1.0;
        s_do = (u <= r) ? 1 : 0;                      // NaN: the comparison is false
        if (s_do) {
            logL[A] = LB_TA; logL[B] = LA_TB;
            if (logPrior) { const double t = logPrior[A]; logPrior[A] = logPrior[B]; logPrior[B] = t; }
        }
        if (swapped) *swapped = s_do;
    }
    __syncthreads();
    if (!s_do) return;
    double* a = params + (size_t)A * stride;
    double* b = params + (size_t)B * stride;
    for (int k = threadIdx.x; k < stride; k += blockDim.x) { const double t = a[k]; a[k] = b[k]; b[k] = t; }
}

cudaError_t tamcmc_launch_pt_swap(double* d_params_star, double* d_logL_star, double* d_logPrior_star, const double* d_Tcoefs, int stride,
                                  int A, double u, int* d_swapped, cudaStream_t st)
{
    tamcmc_pt_swap_kernel<<<1, 256, 0, st>>>(d_params_star, d_logL_star, d_logPrior_star, d_Tcoefs, stride, A, u, d_swapped);
    return cudaGetLastError();
}

cudaError_t tamcmc_expand_configure()
{
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(tamcmc_expand_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(tamcmc_expand_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(tamcmc_expand_kernel<23>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(tamcmc_expand_kernel<TAMCMC_MODEL_ID_MODE_TABLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
}

cudaError_t tamcmc_launch_expand(const ExpandArgs& a, int nblocks, cudaStream_t st)
{
    const size_t smem = sizeof(double) * (size_t)a.params_stride + sizeof(int) * 2 * (size_t)(a.max_tiles + 2);
    dim3 grid((unsigned)nblocks, 1u + (unsigned)((a.max_tiles + EXP_THREADS - 1) / EXP_THREADS), 1u);
    if (a.ksi_part) {       // some star runs the Kallinger2014 model: its normalisation sums come first
        dim3 kgrid((unsigned)nblocks, (unsigned)a.ksi_slices, 1u);
        tamcmc_ksi_kernel<<<kgrid, 256, 0, st>>>(a);
    }
    switch (a.uniform_model) {
    case 3: tamcmc_expand_kernel<3><<<grid, EXP_THREADS, smem, st>>>(a); break;
    case 23: tamcmc_expand_kernel<23><<<grid, EXP_THREADS, smem, st>>>(a); break;
    case TAMCMC_MODEL_ID_MODE_TABLE: tamcmc_expand_kernel<TAMCMC_MODEL_ID_MODE_TABLE><<<grid, EXP_THREADS, smem, st>>>(a); break;
    default: tamcmc_expand_kernel<-1><<<grid, EXP_THREADS, smem, st>>>(a); break;
    }
    return cudaGetLastError();
}
