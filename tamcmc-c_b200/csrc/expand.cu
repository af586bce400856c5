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
__device__ double d_amplitude_ratio_entry(int l, int i, double beta_deg)
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
__device__ double d_lin_interpol(const double* x, const double* y, int Nx, double x_int)
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

// linfit.cpp:17-35 with x = 0..n-1 (models.cpp:6065-6071), then models.cpp:6073-6084
__device__ double d_eta0_fct(const double* fl0, int n_)
{
    double sx = 0, sy = 0, sty = 0, stt = 0;
    const double n = (double)n_;
    for (int i = 0; i < n_; i++) sx += (double)i;
    for (int i = 0; i < n_; i++) sy += fl0[i];
    const double mean_x = sx / n;
    for (int i = 0; i < n_; i++) { double t = (double)i - mean_x; sty += t * fl0[i]; }
    for (int i = 0; i < n_; i++) { double t = (double)i - mean_x; stt += t * t; }
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
__device__ int d_set_imin_imax(double x0, double xlast, int N, int l, double fc_l, double gamma_l,
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
    double aterm[12];      // aj: [a1_0,a1_1,...,a6_0,a6_1]; ajAlm: [a1_0,a1_1,a3_0,a3_1,a5_0,a5_1,eps0,eps1,th0,dl]
    int do_amp;
    int status;
};

// one mode -> ModeRec + CompRecs.  nu[m+l], height[m+l] = H_l * V(m) for m=-l..l.
__device__ void emit_mode(const StarDesc& sd, ModeRec* mrec, CompRec* comps, int l, double fc, double gamma,
                          double fs_window, double c, double asym, const double* nu, const double* height,
                          int* status)
{
    ModeRec mr;
    int i0, i1;
    int bad = d_set_imin_imax(sd.x0, sd.xlast, sd.Nglob, l, fc, gamma, fs_window, c, sd.step, &i0, &i1);
    if (bad) atomicOr(status, TAMCMC_ST_WINDOW);
    mr.i0 = i0; mr.i1 = i1; mr.l = l; mr.fc = fc; mr.gamma = gamma;
    mr.qa = asym / fc;
    mr.qb0 = 1.0 - asym;
    { double k2 = 0.5 * gamma * asym / fc; mr.qc = k2 * k2; }
    mr.pad = 0.0;
    if (!isfinite(fc) || !isfinite(gamma) || (asym != 0.0 && (!isfinite(mr.qa) || !isfinite(mr.qc))))
        atomicOr(status, TAMCMC_ST_NONFINITE);

    const double sg = 2.0 / gamma;
    // largest |x - nu| inside the window, for the dynamic-range classification
    const double xlo = sd.x0 + (double)i0 * sd.step, xhi = sd.x0 + (double)i1 * sd.step;
    int nc = 0;
    for (int m = -l; m <= l; m++) {
        const double A = height[m + l];
        const double v = nu[m + l];
        if (!isfinite(A) || !isfinite(v)) { atomicOr(status, TAMCMC_ST_NONFINITE); continue; }
        if (A == 0.0 || !(gamma > 0.0)) continue;   // contributes exactly 0 (gamma==0: see DESIGN.md deviations)
        const double dmax = fmax(fabs(xlo - v), fabs(xhi - v));
        const double emax = sg * dmax;
        CompRec cr;
        cr.nu = v; cr.m = m;
        if (A >= 1e-20 && A <= 1e20 && emax < 1e5 && sg < 1e12) {
            cr.flags = TAMCMC_CF_FAST;
            cr.s = sg / sqrt(A);
            cr.a = 1.0 / A;
        } else {
            cr.flags = TAMCMC_CF_SLOW;
            cr.s = sg;
            cr.a = A;
        }
        comps[nc++] = cr;
    }
    mr.ncomp = nc;
    *mrec = mr;
}

__device__ void emit_noise(NoiseRec* out, const double* noise_params, int Nnoise, int Nharvey, int* status)
{
    // harvey_like(noise_params.array().abs(), ...) -- noise_models.cpp:15-39; models.cpp:2093-2100
    NoiseRec nr;
    int nh = 0;
    if (Nharvey > TAMCMC_MAX_HARVEY) { atomicOr(status, TAMCMC_ST_BADCFG); Nharvey = TAMCMC_MAX_HARVEY; }
    for (int k = 0; k < TAMCMC_MAX_HARVEY; k++) { nr.H[k] = 0; nr.lnsc[k] = 0; nr.pw[k] = 0; }
    for (int k = 0; k < Nharvey; k++) {
        const double H = fabs(noise_params[3 * k]);
        const double tau = fabs(noise_params[3 * k + 1]);
        const double pw = fabs(noise_params[3 * k + 2]);
        if (!isfinite(H) || !isfinite(tau) || !isfinite(pw)) atomicOr(status, TAMCMC_ST_NONFINITE);
        if (tau != 0.0) {                       // noise_models.cpp:29
            nr.H[nh] = H;
            nr.lnsc[nh] = log((1e-3) * tau);
            nr.pw[nh] = pw;
            nh++;
        }
    }
    nr.nh = nh; nr.pad = 0;
    nr.N0 = (Nnoise > 0) ? fabs(noise_params[Nnoise - 1]) : 0.0;
    if (!isfinite(nr.N0)) atomicOr(status, TAMCMC_ST_NONFINITE);
    *out = nr;
}

// |params[n]/(pi*W)| (models.cpp:2032): long double in the reference, double here (<= 1 ulp apart)
__device__ __forceinline__ double amp_to_height(double a, double W)
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
__device__ double nu_aj(int l, int m, double fc, const double* a /*a1..a6*/, double eta0)
{
    if (l == 0) return fc;
    dd acc = dd_from(fc);
    for (int j = 1; j <= 6; j++)
        if (a[j - 1] != 0.0) acc = dd_add(acc, dd_mul_d_dd(a[j - 1], Phi(j, l, m), Plo(j, l, m)));
    double v = acc.hi;
    if (eta0 > 0) { const double a6_ = a[0] * 1e-6; v = v + fc * eta0 * Qlm(l, m) * (a6_ * a6_); }
    return v;
}

}  // namespace

// -------------------------------------------------------------------------------------------
// expand kernel: grid = nstars*Nchains CTAs, 128 threads
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tamcmc_expand_kernel(ExpandArgs A)
{
    extern __shared__ double sp[];             // this chain's parameter row, staged once
    const int sc = blockIdx.x;                 // star*Nchains + chain
    const int star = sc / A.Nchains;
    const StarDesc sd = A.stars[star];
    ModeRec* modes = A.modes + (size_t)sc * A.modes_stride;
    CompRec* comps = A.comps + (size_t)sc * A.modes_stride * TAMCMC_MAX_COMP_PER_MODE;
    NoiseRec* noise = A.noise + sc;

    __shared__ Common cm;
    __shared__ int s_status;
    const int tid = threadIdx.x;
    const int* pl = sd.plength;
    const int Nmax = pl[0], lmax = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const int model = sd.model_id;
    const int o_split = Nmax + lmax + Nf;
    const int o_width = o_split + Nsplit;
    const int o_noise = o_width + Nwidth;
    const int o_inc = o_noise + Nnoise;
    const int o_cfg = o_inc + Ninc;

    if (tid == 0) {
        s_status = 0;
        if (A.active && !A.active[sc]) s_status = TAMCMC_ST_INACTIVE;
    }
    {
        const double* g = A.params + (size_t)sc * A.params_stride;
        for (int k = tid; k < A.params_stride; k += blockDim.x) sp[k] = g[k];
    }
    __syncthreads();
    const bool inactive = (s_status & TAMCMC_ST_INACTIVE) != 0;
    const double* params = sp;
    const double* fl0_all = params + Nmax + lmax;
    const double* Wl0_all = params + o_width;

    // ---------------- phase 1: per-chain common quantities, spread over three warps ----------------
    if (!inactive) {
        if (tid < 15) {
            // warp 0, lanes 0..14: the 3+5+7 entries of amplitude_ratio(l, inc), l = 1..3 (or the ratio parameters)
            const int l = (tid < 3) ? 1 : (tid < 8) ? 2 : 3;
            const int i = tid - ((l == 1) ? 0 : (l == 2) ? 3 : 8) - l;       // m = -l..l
            const bool have = (model == 11) ? (pl[2 + l] >= 1) : (lmax >= l);
            if (have) {
                if (model == 12) {
                    // models.cpp:2196-2214: m-ratios read from the "inclination" block, symmetric in m
                    const int base = (l == 1) ? 0 : (l == 2) ? 2 : 5;
                    cm.ratios[l][i + l] = fabs(params[o_inc + base + (i < 0 ? -i : i)]);
                } else if (model != 13) {
                    double inc;
                    if (model == 11) {       // models.cpp:3057-3058
                        const double PI = 3.141592653589793238462643383279502884;
                        inc = atan(params[o_split + 4] / params[o_split + 3]) * 180. / PI;
                    } else inc = params[o_inc];
                    cm.ratios[l][i + l] = d_amplitude_ratio_entry(l, i, inc);
                }
            }
        } else if (tid >= 16 && tid <= 18) {
            const int l = tid - 15;
            const bool have = (model == 11) ? (pl[2 + l] >= 1) : (lmax >= l);
            cm.Vl[l] = (have && model != 11) ? fabs(params[Nmax + l - 1]) : 1.0;
        } else if (tid == 32) {
            // warp 1, lane 0: scalar parameters and eta0
            cm.trunc_c = params[o_cfg];
            cm.do_amp = (params[o_cfg + 1] != 0.0);
            cm.ratios[0][0] = 1.0;
            cm.Vl[0] = 1.0;
            cm.status = 0;
            cm.a1 = cm.a3 = cm.a11 = cm.a12 = 0.0;
            switch (model) {
            case 3: case 12: case 13:    // models.cpp:2011-2016, 2219-2222, 2396-2399
                cm.a1 = fabs(params[o_split]);
                cm.eta0 = d_eta0_fct(fl0_all, Nfl0);
                cm.a3 = params[o_split + 2];
                cm.asym = params[o_split + 5];
                break;
            case 6:                      // models.cpp:87-91
                cm.a11 = fabs(params[o_split]);
                cm.a12 = fabs(params[o_split + 6]);
                cm.eta0 = d_eta0_fct(fl0_all, Nfl0);
                cm.a3 = params[o_split + 2];
                cm.asym = params[o_split + 5];
                break;
            case 11:                     // models.cpp:3059-3075 (Nvis plays the role of lmax)
                cm.a1 = params[o_split + 3] * params[o_split + 3] + params[o_split + 4] * params[o_split + 4];
                cm.eta0 = params[o_split + 1];
                cm.a3 = params[o_split + 2];
                cm.asym = params[o_split + 5];
                break;
            case 23:                     // models.cpp:1257-1270
                for (int k = 0; k < 12; k++) cm.aterm[k] = params[o_split + k];
                cm.asym = params[o_split + 13];
                cm.eta0 = (params[o_split + 12] == 1) ? d_eta0_fct(fl0_all, Nfl0) : 0.0;
                break;
            default:
                cm.status = TAMCMC_ST_BADCFG;
                cm.eta0 = 0; cm.asym = 0;
                break;
            }
            if (cm.status) atomicOr(&s_status, cm.status);
        } else if (tid == 64) {
            // warp 2, lane 0: Harvey-like background parameters
            emit_noise(noise, params + o_noise, Nnoise, (model == 11) ? 0 : (Nnoise - 1) / 3, &s_status);
        }
    }
    __syncthreads();

    // ---------------- phase 2: one thread per mode ----------------
    const int nmodes = sd.nmodes_cap;
    if (!inactive && !(s_status & TAMCMC_ST_BADCFG)) {
        for (int j = tid; j < nmodes; j += blockDim.x) {
            int l, n;
            double fc, W, H, fsw;
            double nu[7], hh[7];
            // reference call order -> mode index j
            if (model == 3 || model == 12 || model == 13 || model == 6) {
                l = j % (lmax + 1); n = j / (lmax + 1);          // n-major, l interleaved (models.cpp:2026-2085)
            } else {
                // l-major: all l=0, then l=1, ... (models.cpp:1287-1376, 3082-3134)
                int r = j; l = 0;
                while (l < 3 && r >= pl[2 + l]) { r -= pl[2 + l]; l++; }
                n = r;
            }
            const int o_fl = Nmax + lmax + (l >= 1 ? Nfl0 : 0) + (l >= 2 ? Nfl1 : 0) + (l >= 3 ? Nfl2 : 0);
            fc = params[o_fl + n];

            if (model == 11) {
                // models.cpp:3082-3134: individual heights and widths per mode
                const int idx = (l >= 1 ? Nfl0 : 0) + (l >= 2 ? Nfl1 : 0) + (l >= 3 ? Nfl2 : 0) + n;
                W = fabs(params[o_width + idx]);
                H = cm.do_amp ? amp_to_height(params[idx], W) : fabs(params[idx]);
                for (int m = -l; m <= l; m++) { nu[m + l] = nu_a1etaa3(l, m, fc, cm.a1, cm.eta0, cm.a3); hh[m + l] = H * cm.ratios[l][m + l]; }
                fsw = cm.a1;
            } else if (model == 23) {
                // models.cpp:1287-1376
                double a[6] = {0, 0, 0, 0, 0, 0};
                if (l == 0) {
                    W = fabs(Wl0_all[n]);
                    H = cm.do_amp ? amp_to_height(params[n], W) : fabs(params[n]);
                } else {
                    W = fabs(d_lin_interpol(fl0_all, Wl0_all, Nmax, fc));
                    const double Hi = d_lin_interpol(fl0_all, params, Nmax, fc);
                    const double PI = 3.14159265358979323846;
                    H = cm.do_amp ? fabs(Hi / (PI * W) * cm.Vl[l]) : fabs(Hi * cm.Vl[l]);
                    const int na = 2 * l;            // l=1: a1,a2; l=2: a1..a4; l=3: a1..a6
                    for (int k = 0; k < na; k++) a[k] = cm.aterm[2 * k] + cm.aterm[2 * k + 1] * (fc * 1e-3);
                }
                for (int m = -l; m <= l; m++) { nu[m + l] = nu_aj(l, m, fc, a, (l == 0) ? 0.0 : cm.eta0); hh[m + l] = H * cm.ratios[l][m + l]; }
                fsw = a[0];                           // optimum_lorentzian_calc_aj: window uses a1 (build_lorentzian.cpp:513)
            } else {
                // Classic family and a1l: widths interpolated on the l=0 ladder, heights H[n]*V_l
                W = (l == 0) ? fabs(Wl0_all[n]) : fabs(d_lin_interpol(fl0_all, Wl0_all, Nmax, fc));
                double f_s;
                if (model == 6) {                     // build_lorentzian.cpp:58-66, 383-396
                    f_s = (l == 0) ? 0.0 : (l == 1) ? cm.a11 : (l == 2) ? cm.a12 : (cm.a11 + cm.a12) / 2.;
                } else f_s = cm.a1;
                if (model == 13) {
                    // models.cpp:2409-2470: per-m heights from the parameter vector, |H|/(pi W) if do_amp
                    const double PI = 3.141592653589793238462643383279502884;
                    const int pos0 = (l + 1) * n;
                    for (int m = -l; m <= l; m++) {
                        double h = (l == 0) ? params[n] : params[o_inc + pos0 + (m < 0 ? -m : m)];
                        if (cm.do_amp) h = h / (PI * W);
                        hh[m + l] = fabs(h);
                    }
                } else {
                    if (l == 0) H = cm.do_amp ? amp_to_height(params[n], W) : fabs(params[n]);
                    else H = cm.do_amp ? amp_to_height(params[n], W) * cm.Vl[l] : fabs(params[n] * cm.Vl[l]);
                    for (int m = -l; m <= l; m++) hh[m + l] = H * cm.ratios[l][m + l];
                }
                for (int m = -l; m <= l; m++) nu[m + l] = nu_a1etaa3(l, m, fc, f_s, cm.eta0, cm.a3);
                fsw = f_s;
            }
            emit_mode(sd, modes + j, comps + (size_t)j * TAMCMC_MAX_COMP_PER_MODE, l, fc, W, fsw, cm.trunc_c, cm.asym, nu, hh, &s_status);
        }
    }
    __syncthreads();
    if (tid == 0) {
        A.status[sc] = s_status;
        A.asym_flag[sc] = (!inactive && cm.asym != 0.0) ? 1 : 0;
        if (s_status != 0) A.out_logL[sc] = nan("");
    }
}

cudaError_t tamcmc_launch_expand(const ExpandArgs& a, int nblocks, cudaStream_t st)
{
    tamcmc_expand_kernel<<<nblocks, 128, sizeof(double) * (size_t)a.params_stride, st>>>(a);
    return cudaGetLastError();
}
