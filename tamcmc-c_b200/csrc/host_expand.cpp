// host_expand.cpp -- host expanders: model functions whose mode list is resolved on the host hand the GPU path a
// MODE TABLE (include/tamcmc_gpu.h, TAMCMC_MODEL_MODE_TABLE).  Part of libtamcmc_gpu.so; plain C++ (no CUDA calls).
//
//   tamcmc_host_expand_ajAlm : the host half of model_MS_Global_ajAlm_HarveyLike (tamcmc/sources/models.cpp:1411-1746):
//                              parameter unpacking, interpolated widths/heights, a1/a3/a5/epsilon linear in nu, and either
//                              the direct Alm shift (decompose_Alm = -1, build_lorentzian.cpp:182-190) or the decomposition
//                              of the activity + centrifugal shifts into even a-coefficients (models.cpp:6110-6126,
//                              acoefs.cpp:190-256), exactly what that function passes to optimum_lorentzian_calc_aj/_ajAlm.
//   tamcmc_host_alm          : Alm(l, m, theta0, delta) of Gizon 2002 (external/Alm/Alm_cpp/activity.cpp:181-246) for the
//                              "gate" and "triangle" filters: 2 x integral of |Y_lm|^2 F(theta) sin(theta) over the
//                              northern band, with the same 64-point Gauss-Legendre rule in theta.  The reference's model
//                              interpolates precomputed grids of this integral with GSL (Alm_interp_iter_preinitialised);
//                              GSL and the grid archive are not part of this path -- callers that want the grid values pass
//                              their own callback.
#include "../../include/tamcmc_gpu.h"
#include "host_math.hpp"

#include <cmath>
#include <cstring>
#include <mutex>

namespace {

const double PI = 3.14159265358979323846;

struct GL64 {
    double x[64], w[64];
    GL64()
    {
        const int n = 64;
        for (int i = 0; i < n; i++) {
            double z = std::cos(PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
            for (int it = 0; it < 100; it++) {
                double p1 = 1.0, p2 = 0.0;
                for (int j = 0; j < n; j++) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
                pp = n * (z * p1 - p2) / (z * z - 1.0);
                const double z1 = z;
                z = z1 - p1 / pp;
                if (std::fabs(z - z1) < 1e-16) break;
            }
            x[i] = z; w[i] = 2.0 / ((1.0 - z * z) * pp * pp);
        }
    }
};
const GL64& gl64() { static GL64 g; return g; }

// |Y_lm(theta, .)|^2 for l <= 3 (what boost::math::spherical_harmonic_{r,i} give in activity.cpp:22-35)
double ylm2(int l, int m, double theta)
{
    const int am = m < 0 ? -m : m;
    const double mu = std::cos(theta), s = std::sin(theta);
    double P = 1.0;
    switch (l * 4 + am) {
    case 0: P = 1.0; break;
    case 4: P = mu; break;
    case 5: P = s; break;
    case 8: P = 0.5 * (3 * mu * mu - 1); break;
    case 9: P = 3 * mu * s; break;
    case 10: P = 3 * s * s; break;
    case 12: P = 0.5 * (5 * mu * mu * mu - 3 * mu); break;
    case 13: P = 1.5 * (5 * mu * mu - 1) * s; break;
    case 14: P = 15 * mu * s * s; break;
    case 15: P = 15 * s * s * s; break;
    default: return 0.0;
    }
    const double norm = (2 * l + 1) / (4 * PI) * (double)tamcmc_host::fact_i(l - am) / (double)tamcmc_host::fact_i(l + am);
    return norm * P * P;
}

// triangle_filter on [0, pi/2] (activity.cpp:96-121)
double triangle(double theta, double theta0, double delta)
{
    double F = 0.0;
    if (theta <= theta0 && theta >= 0 && (theta - (theta0 - delta / 2)) > 0) { const double a = 2 / delta; F = a * theta + (1 - a * theta0); }
    if (theta > theta0 && theta <= PI / 2 && (theta - (theta0 + delta / 2)) < 0) { const double a = -2 / delta; F = a * theta + (1 - a * theta0); }
    return F;
}

double builtin_alm(int l, int m, double theta0, double delta, int filter_code, void*)
{
    return tamcmc_host_alm(l, m, theta0, delta, filter_code);
}

}  // namespace

extern "C" {

double tamcmc_host_alm(int l, int m, double theta0, double delta, int filter_code)
{
    if (l < 0 || l > 3 || m < -l || m > l) return -10.0;                      // activity.cpp:240-243
    if (filter_code != 0 && filter_code != 2) return std::nan("");
    if (delta == 0) return 0.0;                                               // activity.cpp:196-198, 216-218
    double tmin = theta0 - delta / 2, tmax = theta0 + delta / 2;              // activity.cpp:184-194, 202-212
    if (tmin < 0) tmin = 0;
    if (tmax > PI / 2) tmax = PI / 2;
    const GL64& g = gl64();
    const double half = 0.5 * (tmax - tmin), mid = 0.5 * (tmax + tmin);
    double acc = 0.0;
    for (int i = 0; i < 64; i++) {
        const double th = mid + half * g.x[i];
        const double F = (filter_code == 0) ? 1.0 : triangle(th, theta0, delta);
        acc += g.w[i] * ylm2(l, m, th) * std::sin(th) * F;
    }
    return 2.0 * (2.0 * PI * half * acc);                                     // phi integral = 2 pi; x2: both hemispheres (activity.cpp:239)
}

int tamcmc_host_expand_ajAlm(const double* params, const int* plength, tamcmc_alm_fn alm, void* alm_user, int capacity,
                             double* row_out, int* nmodes_out)
{
    if (!params || !plength || !row_out || capacity < 1) return TAMCMC_ERR_ARG;
    const int Nmax = plength[0], lmax = plength[1], Nfl0 = plength[2], Nfl1 = plength[3], Nfl2 = plength[4], Nfl3 = plength[5];
    const int Nsplit = plength[6], Nwidth = plength[7], Nnoise = plength[8], Ninc = plength[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const int o_split = Nmax + lmax + Nf, o_width = o_split + Nsplit, o_noise = o_width + Nwidth, o_inc = o_noise + Nnoise;
    const int o_cfg = o_inc + Ninc;
    if (Nsplit < 12 || plength[10] < 4 || Nmax < 2 || lmax > 3 || Nnoise < 1) return TAMCMC_ERR_ARG;
    if (Nf > capacity) return TAMCMC_ERR_ARG;
    const double trunc_c = params[o_cfg];
    const bool do_amp = params[o_cfg + 1] != 0.0;
    const int decompose = (int)params[o_cfg + 2];
    const int filter_code = (int)params[o_cfg + 3];
    if (filter_code != 0 && filter_code != 2) return TAMCMC_ERR_MODEL;        // models.cpp:1444-1465 (gauss exits)
    if (decompose < -1 || decompose > 2) return TAMCMC_ERR_MODEL;             // models.cpp:1601-1603
    if (!alm) alm = builtin_alm;
    const long double pi = M_PI;
    const double* fl0_all = params + Nmax + lmax;
    const double* Wl0_all = params + o_width;
    const double* Hl0_all = params;
    const double* a1_terms = params + o_split;                                // models.cpp:1484-1497: a1, a3, a5, epsilon: (cte, slope)
    const double *a3_terms = a1_terms + 2, *a5_terms = a1_terms + 4, *eps_terms = a1_terms + 6;
    const double thetas[2] = {params[o_split + 8] * M_PI / 180., params[o_split + 9] * M_PI / 180.};
    const double eta0 = (params[o_split + 10] == 1) ? tamcmc_host::eta0_fct(fl0_all, Nfl0) : 0.0;
    const double asym = params[o_split + 11];
    const double inclination = params[o_inc];

    const int row_len = TAMCMC_MT_HEADER + Nnoise + TAMCMC_MT_STRIDE * capacity;
    std::memset(row_out, 0, sizeof(double) * (size_t)row_len);
    row_out[0] = Nf; row_out[1] = inclination; row_out[2] = trunc_c; row_out[3] = asym;
    for (int k = 0; k < Nnoise; k++) row_out[TAMCMC_MT_HEADER + k] = params[o_noise + k];
    double* rec = row_out + TAMCMC_MT_HEADER + Nnoise;
    int j = 0;
    for (int n = 0; n < Nfl0; n++, j++) {                                     // models.cpp:1533-1552
        double* r = rec + (size_t)TAMCMC_MT_STRIDE * j;
        const double W = std::fabs(Wl0_all[n]);
        r[0] = 0; r[1] = fl0_all[n]; r[3] = W;
        r[2] = do_amp ? (double)fabsl(params[n] / (pi * W)) : std::fabs(params[n]);
    }
    for (int l = 1; l <= 3; l++) {
        const int Nfl = (l == 1) ? Nfl1 : (l == 2) ? Nfl2 : Nfl3;
        const int off = Nmax + lmax + Nfl0 + (l >= 2 ? Nfl1 : 0) + (l >= 3 ? Nfl2 : 0);
        const double Vl = (lmax >= l) ? std::fabs(params[Nmax + l - 1]) : 0.0;
        for (int n = 0; n < Nfl; n++, j++) {                                  // models.cpp:1553-1720
            double* r = rec + (size_t)TAMCMC_MT_STRIDE * j;
            const double fl = params[off + n];
            const double W = std::fabs(tamcmc_host::lin_interpol(fl0_all, Wl0_all, Nmax, fl));
            const double Hi = tamcmc_host::lin_interpol(fl0_all, Hl0_all, Nmax, fl);
            const double H = do_amp ? (double)fabsl(Hi / (pi * W) * Vl) : std::fabs(Hi * Vl);
            const double a1 = a1_terms[0] + a1_terms[1] * (fl * 1e-3);
            const double a3 = (l >= 2) ? a3_terms[0] + a3_terms[1] * (fl * 1e-3) : 0;
            const double a5 = (l >= 3) ? a5_terms[0] + a5_terms[1] * (fl * 1e-3) : 0;
            const double eps = eps_terms[0] + eps_terms[1] * (fl * 1e-3);
            r[0] = l; r[1] = fl; r[2] = H; r[3] = W; r[4] = a1; r[6] = a3; r[8] = a5; r[10] = eta0;
            if (decompose == -1) {
                for (int m = -l; m <= l; m++) r[11 + 3 + m] = fl * eps * alm(l, m, thetas[0], thetas[1], filter_code, alm_user);
            } else {
                // decompose_Alm_fct_GSLgrid (models.cpp:6110-6126): fc, eta0, a1, epsilon arrive as long double there
                double nu[7], aj[6];
                const long double fc = fl, e0 = eta0, a1l = a1, el = eps;
                for (int m = -l; m <= l; m++) {
                    nu[m + l] = (double)fc;
                    if (e0 > 0) nu[m + l] = (double)(nu[m + l] + fc * e0 * tamcmc_host::Qlm(l, m) * powl(a1l * 1e-6, 2));
                    nu[m + l] = (double)(nu[m + l] + fc * el * alm(l, m, thetas[0], thetas[1], filter_code, alm_user));
                }
                tamcmc_host::eval_acoefs(l, nu, aj);
                r[5] = aj[1];                                                  // a2
                r[7] = (l >= 2 && decompose != 2) ? aj[3] : 0;                 // a4: models.cpp:1631-1664
                r[9] = (l >= 3 && decompose == 0) ? aj[5] : 0;                 // a6: models.cpp:1694-1720
            }
        }
    }
    if (nmodes_out) *nmodes_out = j;
    return TAMCMC_OK;
}

}  // extern "C"
