/*
 * tamcmc_oracle.c -- CPU oracle (test infrastructure only; see tamcmc_oracle.h).
 *
 * Plain-C restatement of the TAMCMC-C hot path.  Each function cites the
 * reference file:line it follows (paths relative to the upstream repo root).
 * Arithmetic follows the reference's operation order element by element;
 * `long double` is used exactly where the reference uses it.  Build with
 * -ffp-contract=off so that no FMA contraction changes the bin windows.
 *
 * Where the reference relies on Eigen's `.sum()` (likelihoods.cpp:23,
 * linfit.cpp:24-32) the summation ORDER inside Eigen is a vectorised tree that
 * is not reproducible without Eigen itself; the oracle sums left to right.
 * This changes results at the 1e-16 relative level only.
 */
#include "tamcmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------- */
/* acoefs.cpp                                                                */
/* ------------------------------------------------------------------------- */

/* tamcmc/sources/acoefs.cpp:19-49 */
long double orc_Hslm_Ritzoller1991(int s, int l, int m)
{
    const int L = l * (l + 1);
    const double dm = (double)m;
    long double Hsm = 0;
    if (s > 6) return -1;
    switch (s) {
    case 0: Hsm = 1; break;
    case 1: Hsm = 2 * m; break;
    case 2: Hsm = 6 * pow(dm, 2) - 2 * L; break;
    case 3: Hsm = 20 * pow(dm, 3) - 4 * (3 * L - 1) * m; break;
    case 4: Hsm = 70 * pow(dm, 4) - 10 * (6 * L - 5) * pow(dm, 2) + 6 * L * (L - 2); break;
    case 5: Hsm = 252 * pow(dm, 5) - 140 * (2 * L - 3) * pow(dm, 3) + (20 * L * (3 * L - 10) + 48) * m; break;
    case 6:
        Hsm = 924 * pow(dm, 6) - 420 * pow(dm, 4) * (3 * L - 7)
            + 84 * pow(dm, 2) * (5 * pow((double)L, 2) - 25 * L + 14)
            - 20 * L * (pow((double)L, 2) - 8 * L + 12);
        break;
    default: break;
    }
    return Hsm;
}

/* tamcmc/sources/acoefs.cpp:51-110.  s=1..3 are double-precision quotients
 * widened afterwards; s=4..6 are long double quotients H/c. */
long double orc_Pslm(int s, int l, int m)
{
    const double dm = (double)m, dl = (double)l;
    long double H, c, Ps = 0;
    if (s > 6) return 0;
    if (s == 0) Ps = l;
    if (s == 1) Ps = m;
    if (s == 2) {
        if (l > 0) Ps = (3 * pow(dm, 2) - l * (l + 1)) / (2 * l - 1);
        else Ps = 0;
    }
    if (s == 3) {
        if (l > 1) Ps = (5 * pow(dm, 3) - (3 * l * (l + 1) - 1) * m) / ((l - 1) * (2 * l - 1));
        else Ps = 0;
    }
    if (s == 4) {
        H = (35 * pow(dm, 4) - 5 * (6 * l * (l + 1) - 5) * pow(dm, 2)) + 3 * l * (l + 1) * (l * (l + 1) - 2);
        c = 2 * (l - 1) * (2 * l - 1) * (2 * l - 3);
        Ps = (c != 0) ? H / c : 0;
    }
    if (s == 5) {
        H = orc_Hslm_Ritzoller1991(s, l, m);
        c = 8 * (4 * pow(dl, 4) - 20 * pow(dl, 3) + 35 * pow(dl, 2) - 25 * l + 6);
        Ps = (c != 0) ? H / c : 0;
    }
    if (s == 6) {
        H = orc_Hslm_Ritzoller1991(s, l, m);
        c = 64 * pow(dl, 5) - 480 * pow(dl, 4) + 1360 * pow(dl, 3) - 1800 * pow(dl, 2) + 1096 * l - 240;
        Ps = (c != 0) ? H / c : 0;
    }
    return Ps;
}

/* tamcmc/sources/acoefs.cpp:112-145 (Tnlm), 147-180 (Snlm), 190-256 (eval_acoefs) */
void orc_eval_acoefs(int l, const double *nu, double aj[6])
{
    double t[3] = {0, 0, 0}, s[3] = {0, 0, 0};
    long double Num_a1, Den_a1;
    int k;
    for (k = 0; k < 6; k++) aj[k] = 0;
    if (l == 0 || l > 3) return;
    switch (l) {
    case 1:
        t[0] = (nu[2] - nu[0]) / 2;
        s[0] = (nu[0] + nu[2]) / 2 - nu[1];
        aj[0] = t[0];
        aj[1] = s[0] / 3;
        break;
    case 2:
        t[0] = (nu[3] - nu[1]) / 2;
        t[1] = (nu[4] - nu[0]) / 4;
        s[0] = (nu[1] + nu[3]) / 2 - nu[2];
        s[1] = (nu[0] + nu[4]) / 2 - nu[2];
        Num_a1 = t[0] + 4 * t[1];
        Den_a1 = 5;
        aj[0] = Num_a1 / Den_a1;
        aj[1] = (2 * s[1] - s[0]) / 7;
        aj[2] = (t[1] - t[0]) / 5;
        aj[3] = (s[1] - 4 * s[0]) / 70.;
        break;
    case 3:
        t[0] = (nu[4] - nu[2]) / 2;
        t[1] = (nu[5] - nu[1]) / 4;
        t[2] = (nu[6] - nu[0]) / 6;
        s[0] = (nu[2] + nu[4]) / 2 - nu[3];
        s[1] = (nu[1] + nu[5]) / 2 - nu[3];
        s[2] = (nu[0] + nu[6]) / 2 - nu[3];
        aj[0] = t[0] / 14 + 2 * t[1] / 7 + 9 * t[2] / 14;
        aj[2] = -t[0] / 9 - 2 * t[1] / 9 + t[2] / 3;
        aj[4] = t[2] / 42 + 5 * t[0] / 126 - 4 * t[1] / 63;
        aj[1] = (-15 * s[0] + 25 * s[2]) / 126;
        aj[3] = 13 * (s[0] - 7 * s[1] + 3 * s[2]) / 1001;
        aj[5] = (15 * s[0] - 6 * s[1] + s[2]) / 1386;
        break;
    }
}

/* ------------------------------------------------------------------------- */
/* function_rot.cpp                                                          */
/* ------------------------------------------------------------------------- */

/* tamcmc/sources/function_rot.cpp:94-101 (int result of a long product) */
int orc_factorial(int n)
{
    long f = 1;
    long i;
    for (i = 1; i <= n; i++) f = f * i;
    return (int)f;
}

/* tamcmc/sources/function_rot.cpp:90-92 -- INTEGER divisions of int factorials */
double orc_combi(int n, int r)
{
    return orc_factorial(n) / orc_factorial(n - r) / orc_factorial(r);
}

/* tamcmc/sources/function_rot.cpp:76-88 */
double orc_dmm(int l, int m1, int m2, double beta)
{
    double sum = 0, var = 0;
    long s;
    for (s = 0; s <= l - m1; s++) {
        var = orc_combi(l + m2, (int)(l - m1 - s)) * orc_combi(l - m2, (int)s) * pow(-1, (double)(l - m1 - s));
        var = var * pow(cos(beta / 2.), (double)(2 * s + m1 + m2)) * pow(sin(beta / 2.), (double)(2 * l - 2 * s - m1 - m2));
        sum = sum + var;
    }
    sum = sum * sqrt(orc_factorial(l + m1) * orc_factorial(l - m1));
    sum = sum / sqrt(orc_factorial(l + m2) * orc_factorial(l - m2));
    return sum;
}

/* tamcmc/sources/function_rot.cpp:44-74; mat is (2l+1)x(2l+1), row-major, mat[(i+l)*dim + (j+l)] */
void orc_function_rot(int l, double beta, double *mat)
{
    const int dim = 2 * l + 1;
    int i, j;
    for (i = 0; i < dim * dim; i++) mat[i] = 0;
    for (i = 0; i <= l; i++)
        for (j = -i; j <= i; j++)
            mat[(i + l) * dim + (j + l)] = orc_dmm(l, i, j, beta);
    for (i = -l; i <= 0; i++)
        for (j = i; j <= -i; j++)
            mat[(i + l) * dim + (j + l)] = mat[(-i + l) * dim + (-j + l)] * pow(-1, (double)(i - j));
    for (j = 0; j <= l; j++)
        for (i = -j; i <= j; i++)
            mat[(i + l) * dim + (j + l)] = orc_dmm(l, j, i, -beta);
    for (j = -l; j <= 0; j++)
        for (i = j; i <= -j; i++)
            mat[(i + l) * dim + (j + l)] = mat[(-i + l) * dim + (-j + l)] * pow(-1, (double)(i - j));
}

/* tamcmc/sources/function_rot.cpp:15-42 */
void orc_amplitude_ratio(int l, double beta_deg, double *V)
{
    const int dim = 2 * l + 1;
    const double PI = 3.141592653589793238462643;
    double angle = PI * beta_deg / 180.;
    double mat[49];
    int i;
    orc_function_rot(l, angle, mat);
    for (i = 0; i < dim; i++) {
        double v = mat[i * dim + l]; /* column l */
        V[i] = v * v;
    }
}

/* ------------------------------------------------------------------------- */
/* interpol.cpp / linfit.cpp / eta0                                          */
/* ------------------------------------------------------------------------- */

/* tamcmc/sources/interpol.cpp:13-43 */
double orc_lin_interpol(const double *x, const double *y, long Nx, double x_int)
{
    long i = 0;
    double a = 0, b = 0;
    if (x_int >= x[0] && x_int <= x[Nx - 1]) {
        while (x_int < x[i] || x_int > x[i + 1]) i = i + 1;
        if (i == 0 && (x_int < x[i] || x_int > x[i + 1])) i = i + 1;
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

/* tamcmc/sources/linfit.cpp:17-35 (Eigen sums restated left to right) */
void orc_linfit(const double *x, const double *y, long n_, double out[2])
{
    double sx = 0, sy = 0, sty = 0, stt = 0;
    double n = (double)n_;
    double mean_x;
    long i;
    for (i = 0; i < n_; i++) sx += x[i];
    for (i = 0; i < n_; i++) sy += y[i];
    mean_x = sx / n;
    for (i = 0; i < n_; i++) { double t = x[i] - mean_x; sty += t * y[i]; }
    for (i = 0; i < n_; i++) { double t = x[i] - mean_x; stt += t * t; }
    out[0] = sty / stt;
    out[1] = (sy - sx * out[0]) / n;
}

/* tamcmc/sources/models.cpp:6073-6084 */
double orc_eta0_fct_dnu(double Dnu_obs)
{
    const double G = 6.667e-8;
    const double Dnu_sun = 135.1;
    const double R_sun = 6.96342e5;
    const double M_sun = 1.98855e30;
    const double rho_sun = M_sun * 1e3 / (4 * M_PI * pow(R_sun * 1e5, 3) / 3);
    double rho, eta0;
    rho = pow(Dnu_obs / Dnu_sun, 2.) * rho_sun;
    eta0 = 3. * M_PI / (rho * G);
    return eta0;
}

/* tamcmc/sources/models.cpp:6065-6071 (LinSpaced(n,0,n-1) = 0,1,...,n-1) */
double orc_eta0_fct(const double *fl0_all, long n)
{
    double r[2];
    double *xfit = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    long i;
    for (i = 0; i < n; i++) xfit[i] = (double)i;
    orc_linfit(xfit, fl0_all, n, r);
    free(xfit);
    return orc_eta0_fct_dnu(r[0]);
}

/* ------------------------------------------------------------------------- */
/* build_lorentzian.cpp                                                      */
/* ------------------------------------------------------------------------- */

/* tamcmc/sources/build_lorentzian.cpp:583-592 */
double orc_Qlm(int l, int m)
{
    const long double Dnl = 2. / 3;
    double Q;
    Q = (l * (l + 1) - 3 * pow((double)m, 2)) / ((2 * l - 1) * (2 * l + 3));
    Q = Q * Dnl;
    return Q;
}

/* ---- window trace (oracle-only instrumentation) ---- */
static __thread int *tr_l = 0, *tr_i0 = 0, *tr_i1 = 0;
static __thread int tr_cap = 0, tr_n = 0;
void orc_trace_begin(int *l, int *i0, int *i1, int capacity) { tr_l = l; tr_i0 = i0; tr_i1 = i1; tr_cap = capacity; tr_n = 0; }
int orc_trace_end(void) { int n = tr_n; tr_l = tr_i0 = tr_i1 = 0; tr_cap = 0; tr_n = 0; return n; }
static void trace_push(int l, const int iv[2])
{
    if (tr_l && tr_n < tr_cap) { tr_l[tr_n] = l; tr_i0[tr_n] = iv[0]; tr_i1[tr_n] = iv[1]; }
    if (tr_l) tr_n++;
}

/* tamcmc/sources/build_lorentzian.cpp:595-676.  The four `if`s are NOT exclusive
 * (gamma_l==1 or f_s==1 satisfy two of them; the last one wins).  If none holds
 * (NaN inputs) pvals stays unset in the reference; the oracle reports a window error. */
int orc_set_imin_imax(const double *x, long N, int l, double fc_l, double gamma_l,
                      double f_s, double c, double step, int ivals[2])
{
    double p0 = NAN, p1 = NAN;
    if (gamma_l >= 1 && f_s >= 1) {
        if (l != 0) { p0 = fc_l - c * (l * f_s + gamma_l); p1 = fc_l + c * (l * f_s + gamma_l); }
        else { p0 = fc_l - c * gamma_l * 2.2; p1 = fc_l + c * gamma_l * 2.2; }
    }
    if (gamma_l <= 1 && f_s >= 1) {
        if (l != 0) { p0 = fc_l - c * (l * f_s + 1); p1 = fc_l + c * (l * f_s + 1); }
        else { p0 = fc_l - c * 2.2; p1 = fc_l + c * 2.2; }
    }
    if (gamma_l >= 1 && f_s <= 1) {
        if (l != 0) { p0 = fc_l - c * (l + gamma_l); p1 = fc_l + c * (l + gamma_l); }
        else { p0 = fc_l - c * 2.2 * gamma_l; p1 = fc_l + c * 2.2 * gamma_l; }
    }
    if (gamma_l <= 1 && f_s <= 1) {
        if (l != 0) { p0 = fc_l - c * (l + 1); p1 = fc_l + c * (l + 1); }
        else { p0 = fc_l - c * 2.2; p1 = fc_l + c * 2.2; }
    }
    if ((p1 - step) < x[0]) p1 = x[0] + c;
    if ((p0 + step) >= x[N - 1]) p0 = x[N - 1] - c;

    if (!(p0 == p0) || !(p1 == p1)) { ivals[0] = 0; ivals[1] = 0; return ORC_ERR_WINDOW; }
    {
        double f0 = floor((p0 - x[0]) / step);
        double f1 = ceil((p1 - x[0]) / step);
        /* the reference converts double -> int; keep that well defined for huge values */
        if (f0 < -2147483648.0) f0 = -2147483648.0;
        if (f0 > 2147483647.0) f0 = 2147483647.0;
        if (f1 < -2147483648.0) f1 = -2147483648.0;
        if (f1 > 2147483647.0) f1 = 2147483647.0;
        ivals[0] = (int)f0;
        ivals[1] = (int)f1;
    }
    if (ivals[0] < 0) ivals[0] = 0;
    if (ivals[1] > N) ivals[1] = (int)N;
    if (ivals[1] - ivals[0] <= 0) return ORC_ERR_WINDOW;
    return ORC_OK;
}

/* one m-component added to result[]; shared tail of every build_l_mode_* loop body
 * (e.g. build_lorentzian.cpp:143-158): profile, optional asymmetry, accumulation.
 * Pass structure (setConstant / one Eigen expression per statement) is kept. */
static void add_component(const double *x_l, long n, double nu, double HV, double fc_l, double asym,
                          double gamma_l, double *profile, double *tmp, double *tmp2,
                          double *asymetry, double *result)
{
    const double g2 = pow(gamma_l, 2);
    long i;
    for (i = 0; i < n; i++) tmp[i] = nu;                                   /* tmp.setConstant(nu) */
    for (i = 0; i < n; i++) { double d = x_l[i] - tmp[i]; profile[i] = d * d; }
    for (i = 0; i < n; i++) profile[i] = 4 * profile[i] / g2;
    if (asym == 0) {
        for (i = 0; i < n; i++) tmp[i] = 1;
        for (i = 0; i < n; i++) result[i] = result[i] + HV * (1.0 / (tmp[i] + profile[i]));
    } else {
        const double k2 = 0.5 * gamma_l * asym / fc_l;
        for (i = 0; i < n; i++) tmp[i] = 1;
        for (i = 0; i < n; i++) tmp2[i] = k2;
        for (i = 0; i < n; i++) {
            double w = tmp[i] + asym * (x_l[i] / fc_l - tmp[i]);
            asymetry[i] = w * w + tmp2[i] * tmp2[i];
        }
        for (i = 0; i < n; i++) tmp[i] = 1;
        for (i = 0; i < n; i++) result[i] = result[i] + HV * (asymetry[i] * (1.0 / (tmp[i] + profile[i])));
    }
}

typedef struct { double *profile, *tmp, *tmp2, *asymetry; } scratch_t;
static int scratch_alloc(scratch_t *s, long n)
{
    size_t b = sizeof(double) * (size_t)(n > 0 ? n : 1);
    s->profile = (double *)malloc(b); s->tmp = (double *)malloc(b);
    s->tmp2 = (double *)malloc(b); s->asymetry = (double *)malloc(b);
    return (s->profile && s->tmp && s->tmp2 && s->asymetry) ? 0 : 1;
}
static void scratch_free(scratch_t *s) { free(s->profile); free(s->tmp); free(s->tmp2); free(s->asymetry); }

/* tamcmc/sources/build_lorentzian.cpp:131-161 */
void orc_build_l_mode_a1etaa3(const double *x_l, long n, double H_l, double fc_l, double f_s,
                              double eta0, double a3, double asym, double gamma_l, int l,
                              const double *V, double *result)
{
    scratch_t s; int m; long i;
    scratch_alloc(&s, n);
    for (i = 0; i < n; i++) result[i] = 0;
    for (m = -l; m <= l; m++) {
        double nu;
        if (l != 0) nu = fc_l * (1. + eta0 * pow(f_s * 1e-6, 2) * orc_Qlm(l, m)) + m * f_s + orc_Pslm(3, l, m) * a3;
        else nu = fc_l;
        add_component(x_l, n, nu, H_l * V[m + l], fc_l, asym, gamma_l, s.profile, s.tmp, s.tmp2, s.asymetry, result);
    }
    scratch_free(&s);
}

/* tamcmc/sources/build_lorentzian.cpp:317-348 (per-m heights H_lm) */
void orc_build_l_mode_a1etaa3_v2(const double *x_l, long n, const double *H_lm, double fc_l,
                                 double f_s, double eta0, double a3, double asym,
                                 double gamma_l, int l, double *result)
{
    scratch_t s; int m; long i;
    scratch_alloc(&s, n);
    for (i = 0; i < n; i++) result[i] = 0;
    for (m = -l; m <= l; m++) {
        double nu;
        if (l != 0) nu = fc_l * (1. + eta0 * pow(f_s * 1e-6, 2) * orc_Qlm(l, m)) + m * f_s + orc_Pslm(3, l, m) * a3;
        else nu = fc_l;
        add_component(x_l, n, nu, H_lm[m + l], fc_l, asym, gamma_l, s.profile, s.tmp, s.tmp2, s.asymetry, result);
    }
    scratch_free(&s);
}

/* tamcmc/sources/build_lorentzian.cpp:48-86 */
void orc_build_l_mode_a1l_etaa3(const double *x_l, long n, double H_l, double fc_l, double f_s1,
                                double f_s2, double eta0, double a3, double asym,
                                double gamma_l, int l, const double *V, double *result)
{
    scratch_t s; int m; long i; double f_s = 0;
    if (l == 1) f_s = f_s1;
    if (l == 2) f_s = f_s2;
    if (l == 3) f_s = (f_s1 + f_s2) / 2.;
    scratch_alloc(&s, n);
    for (i = 0; i < n; i++) result[i] = 0;
    for (m = -l; m <= l; m++) {
        double nu;
        if (l != 0) nu = fc_l * (1. + eta0 * pow(f_s * 1e-6, 2) * orc_Qlm(l, m)) + m * f_s + orc_Pslm(3, l, m) * a3;
        else nu = fc_l;
        add_component(x_l, n, nu, H_l * V[m + l], fc_l, asym, gamma_l, s.profile, s.tmp, s.tmp2, s.asymetry, result);
    }
    scratch_free(&s);
}

/* tamcmc/sources/build_lorentzian.cpp:88-129 */
void orc_build_l_mode_a1l_a2a3(const double *x_l, long n, double H_l, double fc_l, double f_s1,
                               double f_s2, double a2, double a3, double asym, double gamma_l,
                               int l, const double *V, double *result)
{
    scratch_t s; int m; long i; double f_s = 0, a2_terms;
    if (l == 1) f_s = f_s1;
    if (l == 2) f_s = f_s2;
    if (l == 3) f_s = (f_s1 + f_s2) / 2.;
    scratch_alloc(&s, n);
    for (i = 0; i < n; i++) result[i] = 0;
    for (m = -l; m <= l; m++) {
        double nu;
        if (l != 0) {
            a2_terms = orc_Pslm(2, l, m) * a2;
            nu = fc_l + m * f_s + a2_terms + orc_Pslm(3, l, m) * a3;
        } else nu = fc_l;
        add_component(x_l, n, nu, H_l * V[m + l], fc_l, asym, gamma_l, s.profile, s.tmp, s.tmp2, s.asymetry, result);
    }
    scratch_free(&s);
}

/* nu_nlm of build_l_mode_aj, build_lorentzian.cpp:222-226 (long double accumulation, rounded on assignment) */
static double nu_nlm_aj(double fc_l, double a1, double a2, double a3, double a4, double a5, double a6,
                        double eta0, int l, int m)
{
    double nu = fc_l + a1 * orc_Pslm(1, l, m) + a2 * orc_Pslm(2, l, m) + a3 * orc_Pslm(3, l, m)
              + a4 * orc_Pslm(4, l, m) + a5 * orc_Pslm(5, l, m) + a6 * orc_Pslm(6, l, m);
    if (eta0 > 0) nu = nu + fc_l * eta0 * orc_Qlm(l, m) * pow(a1 * 1e-6, 2);
    return nu;
}

/* tamcmc/sources/build_lorentzian.cpp:208-246 */
void orc_build_l_mode_aj(const double *x_l, long n, double H_l, double fc_l, double a1, double a2,
                         double a3, double a4, double a5, double a6, double eta0, double asym,
                         double gamma_l, int l, const double *V, double *result)
{
    scratch_t s; int m; long i;
    scratch_alloc(&s, n);
    for (i = 0; i < n; i++) result[i] = 0;
    for (m = -l; m <= l; m++) {
        double nu = (l != 0) ? nu_nlm_aj(fc_l, a1, a2, a3, a4, a5, a6, eta0, l, m) : fc_l;
        add_component(x_l, n, nu, H_l * V[m + l], fc_l, asym, gamma_l, s.profile, s.tmp, s.tmp2, s.asymetry, result);
    }
    scratch_free(&s);
}

/* tamcmc/sources/build_lorentzian.cpp:163-206, with Alm(l,m,theta0,delta) supplied by the caller */
void orc_build_l_mode_ajAlm(const double *x_l, long n, double H_l, double fc_l, double a1, double a3,
                            double a5, double eta0, double epsilon_nl, const double *Alm_m,
                            double asym, double gamma_l, int l, const double *V, double *result)
{
    scratch_t s; int m; long i;
    scratch_alloc(&s, n);
    for (i = 0; i < n; i++) result[i] = 0;
    for (m = -l; m <= l; m++) {
        double nu;
        if (l != 0) {
            nu = fc_l + a1 * orc_Pslm(1, l, m) + a3 * orc_Pslm(3, l, m) + a5 * orc_Pslm(5, l, m);
            if (eta0 > 0) nu = nu + fc_l * eta0 * orc_Qlm(l, m) * pow(a1 * 1e-6, 2);
            nu = nu + fc_l * epsilon_nl * Alm_m[m + l];
        } else nu = fc_l;
        add_component(x_l, n, nu, H_l * V[m + l], fc_l, asym, gamma_l, s.profile, s.tmp, s.tmp2, s.asymetry, result);
    }
    scratch_free(&s);
}

/* y_out = y (full copy); y_out.segment += m0; return y_out -- build_lorentzian.cpp:448-457 */
static int replace_with_window_sum(double **y_io, long N, const int iv[2], const double *m0)
{
    double *y_out = (double *)malloc(sizeof(double) * (size_t)N);
    long i;
    if (!y_out) return ORC_ERR_ARG;
    memcpy(y_out, *y_io, sizeof(double) * (size_t)N);
    for (i = iv[0]; i < iv[1]; i++) y_out[i] = y_out[i] + m0[i - iv[0]];
    free(*y_io);
    *y_io = y_out;
    return ORC_OK;
}

/* tamcmc/sources/build_lorentzian.cpp:441-458 */
int orc_optimum_lorentzian_calc_a1etaa3(const double *x, long N, double **y_io, double H_l, double fc_l,
                                        double f_s, double eta0, double a3, double asym, double gamma_l,
                                        int l, const double *V, double step, double c)
{
    int iv[2], rc; long nw; double *x_l, *m0;
    rc = orc_set_imin_imax(x, N, l, fc_l, gamma_l, f_s, c, step, iv);
    trace_push(l, iv);
    if (rc) return rc;
    nw = iv[1] - iv[0];
    x_l = (double *)malloc(sizeof(double) * (size_t)nw);
    m0 = (double *)malloc(sizeof(double) * (size_t)nw);
    memcpy(x_l, x + iv[0], sizeof(double) * (size_t)nw);           /* x_l = x.segment(...) */
    orc_build_l_mode_a1etaa3(x_l, nw, H_l, fc_l, f_s, eta0, a3, asym, gamma_l, l, V, m0);
    rc = replace_with_window_sum(y_io, N, iv, m0);
    free(x_l); free(m0);
    return rc;
}

/* tamcmc/sources/build_lorentzian.cpp:525-541 */
int orc_optimum_lorentzian_calc_a1etaa3_v2(const double *x, long N, double **y_io, const double *H_lm,
                                           double fc_l, double f_s, double eta0, double a3, double asym,
                                           double gamma_l, int l, double step, double c)
{
    int iv[2], rc; long nw; double *x_l, *m0;
    rc = orc_set_imin_imax(x, N, l, fc_l, gamma_l, f_s, c, step, iv);
    trace_push(l, iv);
    if (rc) return rc;
    nw = iv[1] - iv[0];
    x_l = (double *)malloc(sizeof(double) * (size_t)nw);
    m0 = (double *)malloc(sizeof(double) * (size_t)nw);
    memcpy(x_l, x + iv[0], sizeof(double) * (size_t)nw);
    orc_build_l_mode_a1etaa3_v2(x_l, nw, H_lm, fc_l, f_s, eta0, a3, asym, gamma_l, l, m0);
    rc = replace_with_window_sum(y_io, N, iv, m0);
    free(x_l); free(m0);
    return rc;
}

static double fs_of_l(int l, double f_s1, double f_s2)
{
    switch (l) {
    case 0: return 0.;
    case 1: return f_s1;
    case 2: return f_s2;
    case 3: return (f_s1 + f_s2) / 2.;
    }
    return 0.;
}

/* tamcmc/sources/build_lorentzian.cpp:371-405 */
int orc_optimum_lorentzian_calc_a1l_etaa3(const double *x, long N, double **y_io, double H_l, double fc_l,
                                          double f_s1, double f_s2, double eta0, double a3, double asym,
                                          double gamma_l, int l, const double *V, double step, double c)
{
    int iv[2], rc; long nw; double *x_l, *m0;
    double f_s = fs_of_l(l, f_s1, f_s2);
    rc = orc_set_imin_imax(x, N, l, fc_l, gamma_l, f_s, c, step, iv);
    trace_push(l, iv);
    if (rc) return rc;
    nw = iv[1] - iv[0];
    x_l = (double *)malloc(sizeof(double) * (size_t)nw);
    m0 = (double *)malloc(sizeof(double) * (size_t)nw);
    memcpy(x_l, x + iv[0], sizeof(double) * (size_t)nw);
    orc_build_l_mode_a1l_etaa3(x_l, nw, H_l, fc_l, f_s1, f_s2, eta0, a3, asym, gamma_l, l, V, m0);
    rc = replace_with_window_sum(y_io, N, iv, m0);
    free(x_l); free(m0);
    return rc;
}

/* tamcmc/sources/build_lorentzian.cpp:408-438 */
int orc_optimum_lorentzian_calc_a1l_a2a3(const double *x, long N, double **y_io, double H_l, double fc_l,
                                         double f_s1, double f_s2, double a2, double a3, double asym,
                                         double gamma_l, int l, const double *V, double step, double c)
{
    int iv[2], rc; long nw; double *x_l, *m0;
    double f_s = fs_of_l(l, f_s1, f_s2);
    rc = orc_set_imin_imax(x, N, l, fc_l, gamma_l, f_s, c, step, iv);
    trace_push(l, iv);
    if (rc) return rc;
    nw = iv[1] - iv[0];
    x_l = (double *)malloc(sizeof(double) * (size_t)nw);
    m0 = (double *)malloc(sizeof(double) * (size_t)nw);
    memcpy(x_l, x + iv[0], sizeof(double) * (size_t)nw);
    orc_build_l_mode_a1l_a2a3(x_l, nw, H_l, fc_l, f_s1, f_s2, a2, a3, asym, gamma_l, l, V, m0);
    rc = replace_with_window_sum(y_io, N, iv, m0);
    free(x_l); free(m0);
    return rc;
}

/* tamcmc/sources/build_lorentzian.cpp:502-522 (window uses a1 as the splitting) */
int orc_optimum_lorentzian_calc_aj(const double *x, long N, double H_l, double fc_l, double a1, double a2,
                                   double a3, double a4, double a5, double a6, double eta0, double asym,
                                   double gamma_l, int l, const double *V, double step, double c,
                                   orc_Optim_L *out)
{
    int iv[2], rc; long nw; double *x_l;
    out->y = 0; out->i0 = 0; out->N = 0;
    rc = orc_set_imin_imax(x, N, l, fc_l, gamma_l, a1, c, step, iv);
    trace_push(l, iv);
    if (rc) return rc;
    nw = iv[1] - iv[0];
    x_l = (double *)malloc(sizeof(double) * (size_t)nw);
    out->y = (double *)malloc(sizeof(double) * (size_t)nw);
    memcpy(x_l, x + iv[0], sizeof(double) * (size_t)nw);
    orc_build_l_mode_aj(x_l, nw, H_l, fc_l, a1, a2, a3, a4, a5, a6, eta0, asym, gamma_l, l, V, out->y);
    out->i0 = iv[0];
    out->N = (int)nw;
    free(x_l);
    return ORC_OK;
}

/* tamcmc/sources/build_lorentzian.cpp:480-499 */
static int optimum_lorentzian_calc_ajAlm(const double *x, long N, double H_l, double fc_l, double a1,
                                         double a3, double a5, double eta0, double epsilon_nl,
                                         const double *Alm_m, double asym, double gamma_l, int l,
                                         const double *V, double step, double c, orc_Optim_L *out)
{
    int iv[2], rc; long nw; double *x_l;
    out->y = 0; out->i0 = 0; out->N = 0;
    rc = orc_set_imin_imax(x, N, l, fc_l, gamma_l, a1, c, step, iv);
    trace_push(l, iv);
    if (rc) return rc;
    nw = iv[1] - iv[0];
    x_l = (double *)malloc(sizeof(double) * (size_t)nw);
    out->y = (double *)malloc(sizeof(double) * (size_t)nw);
    memcpy(x_l, x + iv[0], sizeof(double) * (size_t)nw);
    orc_build_l_mode_ajAlm(x_l, nw, H_l, fc_l, a1, a3, a5, eta0, epsilon_nl, Alm_m, asym, gamma_l, l, V, out->y);
    out->i0 = iv[0];
    out->N = (int)nw;
    free(x_l);
    return ORC_OK;
}

/* model_final.segment(i0,N) += model_tmp.y  (e.g. models.cpp:1296-1298) */
static void add_block(double *model, orc_Optim_L *b)
{
    int i;
    for (i = 0; i < b->N; i++) model[b->i0 + i] = model[b->i0 + i] + b->y[i];
    free(b->y); b->y = 0;
}

/* ------------------------------------------------------------------------- */
/* noise_models.cpp / likelihoods.cpp                                        */
/* ------------------------------------------------------------------------- */

/* tamcmc/sources/noise_models.cpp:15-39.  noise_params = [H0,tc0,p0,...,N0] */
void orc_harvey_like(const double *noise_params, int n_noise, const double *x, long N,
                     double **y_io, int Nharvey)
{
    size_t b = sizeof(double) * (size_t)N;
    double *ones = (double *)malloc(b), *white = (double *)malloc(b), *tmp = (double *)malloc(b), *y_out = (double *)malloc(b);
    int cpt = 0, k; long i;
    for (i = 0; i < N; i++) white[i] = noise_params[n_noise - 1];
    memcpy(y_out, *y_io, b);
    for (k = 0; k < Nharvey; k++) {
        if (noise_params[cpt + 1] != 0) {
            const double sc = (1e-3) * noise_params[cpt + 1];
            const double pw = noise_params[cpt + 2];
            const double h = noise_params[cpt];
            for (i = 0; i < N; i++) tmp[i] = pow(sc * x[i], pw);
            for (i = 0; i < N; i++) ones[i] = 1;
            for (i = 0; i < N; i++) tmp[i] = h * (1.0 / (tmp[i] + ones[i]));
            for (i = 0; i < N; i++) y_out[i] = y_out[i] + tmp[i];
        }
        cpt = cpt + 3;
    }
    for (i = 0; i < N; i++) y_out[i] = y_out[i] + white[i];
    free(ones); free(white); free(tmp);
    free(*y_io);
    *y_io = y_out;
}

/* tamcmc/sources/likelihoods.cpp:17-28: two double reductions, widened, times -p */
long double orc_likelihood_chi22p(const double *y, const double *model, long N, long p)
{
    long double f;
    double s1 = 0, s2 = 0;
    long i;
    for (i = 0; i < N; i++) s1 += y[i] * (1.0 / model[i]);
    for (i = 0; i < N; i++) s2 += log(model[i]);
    f = s1 + s2;   /* double + double, as the reference's (…).sum() + (…).sum() */
    f = -p * f;
    return f;
}

/* tamcmc/sources/likelihoods.cpp:31-40 */
long double orc_likelihood_chi_square(const double *y, const double *model, const double *sigma, long N)
{
    long double f;
    double s = 0;
    long i;
    for (i = 0; i < N; i++) { double d = y[i] - model[i]; s += (d * d) * (1.0 / (sigma[i] * sigma[i])); }
    f = -s;
    f = f / 2;
    return f;
}

/* tamcmc/sources/model_def.cpp:390-401: p (double) truncated to long, result divided by Tcoefs[m] */
long double orc_call_likelihood_chi22p(const double *y, const double *model, long N, double p, double Tcoef)
{
    long double logL = orc_likelihood_chi22p(y, model, N, (long)p);
    return logL / Tcoef;
}

/* ------------------------------------------------------------------------- */
/* models.cpp                                                                */
/* ------------------------------------------------------------------------- */

static double *zeros(long N) { return (double *)calloc((size_t)(N > 0 ? N : 1), sizeof(double)); }
static void abs_copy(const double *src, int n, double *dst) { int i; for (i = 0; i < n; i++) dst[i] = fabs(src[i]); }

static const long double PI_L = 3.141592653589793238462643383279502884L;

/* tamcmc/sources/models.cpp:1943-2121 */
static int model_MS_Global_a1etaa3_HarveyLike_Classic(const double *params, const int *pl, const double *x, long N, double *out)
{
    const double step = x[1] - x[0];
    const long double pi = PI_L;
    const int Nmax = pl[0], lmax = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const int do_amp = (params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc + 1] != 0);
    double inclination, trunc_c;
    double r0[1] = {1}, r1[3], r2[5], r3[7];
    double Vl1 = 0, Vl2 = 0, Vl3 = 0, Hl0, Hl1, Hl2, Hl3, Wl0, Wl1, Wl2, Wl3, a1, eta0, a3, asym, fl0, fl1, fl2, fl3;
    const double *fl0_all = params + Nmax + lmax;
    const double *Wl0_all = params + Nmax + lmax + Nf + Nsplit;
    double *model = zeros(N), *noise_abs; long n; int rc = 0, Nharvey;

    trunc_c = params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc];
    inclination = params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise];
    if (lmax >= 1) { Vl1 = fabs(params[Nmax]); orc_amplitude_ratio(1, inclination, r1); }
    if (lmax >= 2) { Vl2 = fabs(params[Nmax + 1]); orc_amplitude_ratio(2, inclination, r2); }
    if (lmax >= 3) { Vl3 = fabs(params[Nmax + 2]); orc_amplitude_ratio(3, inclination, r3); }
    a1 = fabs(params[Nmax + lmax + Nf]);
    eta0 = orc_eta0_fct(fl0_all, Nfl0);
    a3 = params[Nmax + lmax + Nf + 2];
    asym = params[Nmax + lmax + Nf + 5];

    for (n = 0; n < Nmax && !rc; n++) {
        fl0 = fl0_all[n];
        Wl0 = fabs(Wl0_all[n]);
        if (do_amp) Hl0 = (double)fabsl(params[n] / (pi * Wl0)); else Hl0 = fabs(params[n]);
        rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl0, fl0, a1, eta0, a3, asym, Wl0, 0, r0, step, trunc_c);
        if (rc) break;
        if (lmax >= 1) {
            fl1 = params[Nmax + lmax + Nfl0 + n];
            Wl1 = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl1));
            if (do_amp) Hl1 = (double)(fabsl(params[n] / (pi * Wl1)) * Vl1); else Hl1 = fabs(params[n] * Vl1);
            rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl1, fl1, a1, eta0, a3, asym, Wl1, 1, r1, step, trunc_c);
            if (rc) break;
        }
        if (lmax >= 2) {
            fl2 = params[Nmax + lmax + Nfl0 + Nfl1 + n];
            Wl2 = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl2));
            if (do_amp) Hl2 = (double)(fabsl(params[n] / (pi * Wl2)) * Vl2); else Hl2 = fabs(params[n] * Vl2);
            rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl2, fl2, a1, eta0, a3, asym, Wl2, 2, r2, step, trunc_c);
            if (rc) break;
        }
        if (lmax >= 3) {
            fl3 = params[Nmax + lmax + Nfl0 + Nfl1 + Nfl2 + n];
            Wl3 = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl3));
            if (do_amp) Hl3 = (double)(fabsl(params[n] / (pi * Wl3)) * Vl3); else Hl3 = fabs(params[n] * Vl3);
            rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl3, fl3, a1, eta0, a3, asym, Wl3, 3, r3, step, trunc_c);
            if (rc) break;
        }
    }
    if (!rc) {
        noise_abs = (double *)malloc(sizeof(double) * (size_t)(Nnoise > 0 ? Nnoise : 1));
        abs_copy(params + Nmax + lmax + Nf + Nsplit + Nwidth, Nnoise, noise_abs);
        Nharvey = (Nnoise - 1) / 3;
        orc_harvey_like(noise_abs, Nnoise, x, N, &model, Nharvey);
        free(noise_abs);
        memcpy(out, model, sizeof(double) * (size_t)N);
    }
    free(model);
    return rc;
}

/* tamcmc/sources/models.cpp:2128-2336 (per-m height ratios are parameters; Hl*V still multiplies them) */
static int model_MS_Global_a1etaa3_HarveyLike_Classic_v2(const double *params, const int *pl, const double *x, long N, double *out)
{
    const double step = x[1] - x[0];
    const long double pi = PI_L;
    const int Nmax = pl[0], lmax = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const int do_amp = (params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc + 1] != 0);
    const int b = Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise;
    double trunc_c;
    double r0[1] = {1}, r1[3], r2[5], r3[7];
    double Vl1 = 0, Vl2 = 0, Vl3 = 0, Hl0, Hl1, Hl2, Hl3, Wl0, Wl1, Wl2, Wl3, a1, eta0, a3, asym, fl0, fl1, fl2, fl3;
    const double *fl0_all = params + Nmax + lmax;
    const double *Wl0_all = params + Nmax + lmax + Nf + Nsplit;
    double *model = zeros(N), *noise_abs; long n; int rc = 0, Nharvey;

    trunc_c = params[b + Ninc];
    if (lmax >= 1) Vl1 = fabs(params[Nmax]);
    if (lmax >= 2) Vl2 = fabs(params[Nmax + 1]);
    if (lmax >= 3) Vl3 = fabs(params[Nmax + 2]);
    /* models.cpp:2196-2214: the reference reads all nine slots whatever lmax is;
     * the oracle reads only the slots of existing degrees (unused otherwise). */
    if (lmax >= 1) { r1[0] = fabs(params[b + 1]); r1[1] = fabs(params[b]); r1[2] = fabs(params[b + 1]); }
    if (lmax >= 2) { r2[0] = fabs(params[b + 4]); r2[1] = fabs(params[b + 3]); r2[2] = fabs(params[b + 2]); r2[3] = fabs(params[b + 3]); r2[4] = fabs(params[b + 4]); }
    if (lmax >= 3) { r3[0] = fabs(params[b + 8]); r3[1] = fabs(params[b + 7]); r3[2] = fabs(params[b + 6]); r3[3] = fabs(params[b + 5]);
                     r3[4] = fabs(params[b + 6]); r3[5] = fabs(params[b + 7]); r3[6] = fabs(params[b + 8]); }
    a1 = fabs(params[Nmax + lmax + Nf]);
    eta0 = orc_eta0_fct(fl0_all, Nfl0);
    a3 = params[Nmax + lmax + Nf + 2];
    asym = params[Nmax + lmax + Nf + 5];

    for (n = 0; n < Nmax && !rc; n++) {
        fl0 = fl0_all[n];
        Wl0 = fabs(Wl0_all[n]);
        if (do_amp) Hl0 = (double)fabsl(params[n] / (pi * Wl0)); else Hl0 = fabs(params[n]);
        rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl0, fl0, a1, eta0, a3, asym, Wl0, 0, r0, step, trunc_c);
        if (rc) break;
        if (lmax >= 1) {
            fl1 = params[Nmax + lmax + Nfl0 + n];
            Wl1 = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl1));
            if (do_amp) Hl1 = (double)(fabsl(params[n] / (pi * Wl1)) * Vl1); else Hl1 = fabs(params[n] * Vl1);
            rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl1, fl1, a1, eta0, a3, asym, Wl1, 1, r1, step, trunc_c);
            if (rc) break;
        }
        if (lmax >= 2) {
            fl2 = params[Nmax + lmax + Nfl0 + Nfl1 + n];
            Wl2 = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl2));
            if (do_amp) Hl2 = (double)(fabsl(params[n] / (pi * Wl2)) * Vl2); else Hl2 = fabs(params[n] * Vl2);
            rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl2, fl2, a1, eta0, a3, asym, Wl2, 2, r2, step, trunc_c);
            if (rc) break;
        }
        if (lmax >= 3) {
            fl3 = params[Nmax + lmax + Nfl0 + Nfl1 + Nfl2 + n];
            Wl3 = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl3));
            if (do_amp) Hl3 = (double)(fabsl(params[n] / (pi * Wl3)) * Vl3); else Hl3 = fabs(params[n] * Vl3);
            rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl3, fl3, a1, eta0, a3, asym, Wl3, 3, r3, step, trunc_c);
            if (rc) break;
        }
    }
    if (!rc) {
        noise_abs = (double *)malloc(sizeof(double) * (size_t)(Nnoise > 0 ? Nnoise : 1));
        abs_copy(params + Nmax + lmax + Nf + Nsplit + Nwidth, Nnoise, noise_abs);
        Nharvey = (Nnoise - 1) / 3;
        orc_harvey_like(noise_abs, Nnoise, x, N, &model, Nharvey);
        free(noise_abs);
        memcpy(out, model, sizeof(double) * (size_t)N);
    }
    free(model);
    return rc;
}

/* tamcmc/sources/models.cpp:2338-2553 (heights H_nlm are parameters stored in the "inclination" block).
 * `Hl1=Hl1/(pi*Wl1)` divides a VectorXd by a long double scalar: Eigen converts the
 * scalar to double first, so that quotient is a double division. */
static int model_MS_Global_a1etaa3_HarveyLike_Classic_v3(const double *params, const int *pl, const double *x, long N, double *out)
{
    const double step = x[1] - x[0];
    const long double pi = PI_L;
    const int Nmax = pl[0], lmax = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const int do_amp = (params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc + 1] != 0);
    const int b = Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise;
    double trunc_c, Hl0[1], Hl1[3], Hl2[5], Hl3[7];
    double Wl0, Wl1, Wl2, Wl3, a1, eta0, a3, asym, fl0, fl1, fl2, fl3;
    const double *fl0_all = params + Nmax + lmax;
    const double *Wl0_all = params + Nmax + lmax + Nf + Nsplit;
    double *model = zeros(N), *noise_abs; long n, pos0; int rc = 0, Nharvey, k;

    trunc_c = params[b + Ninc];
    a1 = fabs(params[Nmax + lmax + Nf]);
    eta0 = orc_eta0_fct(fl0_all, Nfl0);
    a3 = params[Nmax + lmax + Nf + 2];
    asym = params[Nmax + lmax + Nf + 5];

    for (n = 0; n < Nmax && !rc; n++) {
        fl0 = fl0_all[n];
        Wl0 = fabs(Wl0_all[n]);
        if (do_amp) Hl0[0] = (double)fabsl(params[n] / (pi * Wl0)); else Hl0[0] = fabs(params[n]);
        rc = orc_optimum_lorentzian_calc_a1etaa3_v2(x, N, &model, Hl0, fl0, a1, eta0, a3, asym, Wl0, 0, step, trunc_c);
        if (rc) break;
        if (lmax >= 1) {
            fl1 = params[Nmax + lmax + Nfl0 + n];
            Wl1 = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl1));
            pos0 = 2 * n;
            Hl1[0] = params[b + pos0 + 1]; Hl1[1] = params[b + pos0]; Hl1[2] = params[b + pos0 + 1];
            if (do_amp) { const double d = (double)(pi * Wl1); for (k = 0; k < 3; k++) Hl1[k] = Hl1[k] / d; }
            for (k = 0; k < 3; k++) Hl1[k] = fabs(Hl1[k]);
            rc = orc_optimum_lorentzian_calc_a1etaa3_v2(x, N, &model, Hl1, fl1, a1, eta0, a3, asym, Wl1, 1, step, trunc_c);
            if (rc) break;
        }
        if (lmax >= 2) {
            fl2 = params[Nmax + lmax + Nfl0 + Nfl1 + n];
            Wl2 = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl2));
            pos0 = 3 * n;
            Hl2[0] = params[b + pos0 + 2]; Hl2[1] = params[b + pos0 + 1]; Hl2[2] = params[b + pos0];
            Hl2[3] = params[b + pos0 + 1]; Hl2[4] = params[b + pos0 + 2];
            if (do_amp) { const double d = (double)(pi * Wl2); for (k = 0; k < 5; k++) Hl2[k] = Hl2[k] / d; }
            for (k = 0; k < 5; k++) Hl2[k] = fabs(Hl2[k]);
            rc = orc_optimum_lorentzian_calc_a1etaa3_v2(x, N, &model, Hl2, fl2, a1, eta0, a3, asym, Wl2, 2, step, trunc_c);
            if (rc) break;
        }
        if (lmax >= 3) {
            fl3 = params[Nmax + lmax + Nfl0 + Nfl1 + Nfl2 + n];
            Wl3 = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl3));
            pos0 = 4 * n;
            Hl3[0] = params[b + pos0 + 3]; Hl3[1] = params[b + pos0 + 2]; Hl3[2] = params[b + pos0 + 1]; Hl3[3] = params[b + pos0];
            Hl3[4] = params[b + pos0 + 1]; Hl3[5] = params[b + pos0 + 2]; Hl3[6] = params[b + pos0 + 3];
            if (do_amp) { const double d = (double)(pi * Wl3); for (k = 0; k < 7; k++) Hl3[k] = Hl3[k] / d; }
            for (k = 0; k < 7; k++) Hl3[k] = fabs(Hl3[k]);
            rc = orc_optimum_lorentzian_calc_a1etaa3_v2(x, N, &model, Hl3, fl3, a1, eta0, a3, asym, Wl3, 3, step, trunc_c);
            if (rc) break;
        }
    }
    if (!rc) {
        noise_abs = (double *)malloc(sizeof(double) * (size_t)(Nnoise > 0 ? Nnoise : 1));
        abs_copy(params + Nmax + lmax + Nf + Nsplit + Nwidth, Nnoise, noise_abs);
        Nharvey = (Nnoise - 1) / 3;
        orc_harvey_like(noise_abs, Nnoise, x, N, &model, Nharvey);
        free(noise_abs);
        memcpy(out, model, sizeof(double) * (size_t)N);
    }
    free(model);
    return rc;
}

/* tamcmc/sources/models.cpp:25-215 */
static int model_MS_Global_a1l_etaa3_HarveyLike(const double *params, const int *pl, const double *x, long N, double *out)
{
    const double step = x[1] - x[0];
    const long double pi = PI_L;
    const int Nmax = pl[0], lmax = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const int do_amp = (params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc + 1] != 0);
    double inclination, trunc_c;
    double r0[1] = {1}, r1[3], r2[5], r3[7];
    double Vl1 = 0, Vl2 = 0, Vl3 = 0, Hl0, Hl1, Hl2, Hl3, Wl0, Wl1, Wl2, Wl3, a11, a12, eta0, a3, asym, fl0, fl1, fl2, fl3;
    const double *fl0_all = params + Nmax + lmax;
    const double *Wl0_all = params + Nmax + lmax + Nf + Nsplit;
    double *model = zeros(N), *noise_abs; long n; int rc = 0, Nharvey;

    trunc_c = params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc];
    inclination = params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise];
    if (lmax >= 1) { Vl1 = fabs(params[Nmax]); orc_amplitude_ratio(1, inclination, r1); }
    if (lmax >= 2) { Vl2 = fabs(params[Nmax + 1]); orc_amplitude_ratio(2, inclination, r2); }
    if (lmax >= 3) { Vl3 = fabs(params[Nmax + 2]); orc_amplitude_ratio(3, inclination, r3); }
    a11 = fabs(params[Nmax + lmax + Nf]);
    a12 = fabs(params[Nmax + lmax + Nf + 6]);
    eta0 = orc_eta0_fct(fl0_all, Nfl0);
    a3 = params[Nmax + lmax + Nf + 2];
    asym = params[Nmax + lmax + Nf + 5];

    for (n = 0; n < Nmax && !rc; n++) {
        fl0 = fl0_all[n];
        Wl0 = fabs(Wl0_all[n]);
        if (do_amp) Hl0 = (double)fabsl(params[n] / (pi * Wl0)); else Hl0 = fabs(params[n]);
        rc = orc_optimum_lorentzian_calc_a1l_etaa3(x, N, &model, Hl0, fl0, a11, a12, eta0, a3, asym, Wl0, 0, r0, step, trunc_c);
        if (rc) break;
        if (lmax >= 1) {
            fl1 = params[Nmax + lmax + Nfl0 + n];
            Wl1 = orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl1);
            Wl1 = fabs(Wl1);
            if (do_amp) Hl1 = (double)(fabsl(params[n] / (pi * Wl1)) * Vl1); else Hl1 = fabs(params[n] * Vl1);
            rc = orc_optimum_lorentzian_calc_a1l_etaa3(x, N, &model, Hl1, fl1, a11, a12, eta0, a3, asym, Wl1, 1, r1, step, trunc_c);
            if (rc) break;
        }
        if (lmax >= 2) {
            fl2 = params[Nmax + lmax + Nfl0 + Nfl1 + n];
            Wl2 = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl2));
            if (do_amp) Hl2 = (double)(fabsl(params[n] / (pi * Wl2)) * Vl2); else Hl2 = fabs(params[n] * Vl2);
            rc = orc_optimum_lorentzian_calc_a1l_etaa3(x, N, &model, Hl2, fl2, a11, a12, eta0, a3, asym, Wl2, 2, r2, step, trunc_c);
            if (rc) break;
        }
        if (lmax >= 3) {
            fl3 = params[Nmax + lmax + Nfl0 + Nfl1 + Nfl2 + n];
            Wl3 = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl3));
            if (do_amp) Hl3 = (double)(fabsl(params[n] / (pi * Wl3)) * Vl3); else Hl3 = fabs(params[n] * Vl3);
            rc = orc_optimum_lorentzian_calc_a1l_etaa3(x, N, &model, Hl3, fl3, a11, a12, eta0, a3, asym, Wl3, 3, r3, step, trunc_c);
            if (rc) break;
        }
    }
    if (!rc) {
        noise_abs = (double *)malloc(sizeof(double) * (size_t)(Nnoise > 0 ? Nnoise : 1));
        abs_copy(params + Nmax + lmax + Nf + Nsplit + Nwidth, Nnoise, noise_abs);
        Nharvey = (Nnoise - 1) / 3;
        orc_harvey_like(noise_abs, Nnoise, x, N, &model, Nharvey);
        free(noise_abs);
        memcpy(out, model, sizeof(double) * (size_t)N);
    }
    free(model);
    return rc;
}

/* model_MS_Global_a1n_etaa3_HarveyLike (models.cpp:217-407, id 7), model_MS_Global_a1nl_etaa3_HarveyLike (:1003-1193, id 8),
 * model_MS_Global_a1n_a2a3_HarveyLike (:409-607, id 18), model_MS_Global_a1nl_a2a3_HarveyLike (:805-1001, id 19): the loop of
 * model_MS_Global_a1l_etaa3_HarveyLike with the splittings read per radial order:
 *   a11[n] = |params[split+6+n]|;  a12[n] = a11[n] (a1n) or |params[split+6+Nmax+n]| (a1nl);
 *   a2a3 variants: no eta term, a2[n] = params[split+6+Nmax+n] (a1n) / params[split+6+2Nmax+n] (a1nl), build_l_mode_a1l_a2a3 */
static int model_MS_Global_a1x_HarveyLike(const double *params, const int *pl, const double *x, long N, double *out, int id)
{
    const double step = x[1] - x[0];
    const long double pi = PI_L;
    const int Nmax = pl[0], lmax = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const int o_split = Nmax + lmax + Nf;
    const int do_amp = (params[o_split + Nsplit + Nwidth + Nnoise + Ninc + 1] != 0);
    const int a2a3 = (id == 18 || id == 19), nl = (id == 8 || id == 19);
    const double trunc_c = params[o_split + Nsplit + Nwidth + Nnoise + Ninc];
    const double inclination = params[o_split + Nsplit + Nwidth + Nnoise];
    double ratios[4][7], Vl[4] = {1, 0, 0, 0};
    const double *fl0_all = params + Nmax + lmax;
    const double *Wl0_all = params + o_split + Nsplit;
    const double eta0 = a2a3 ? 0.0 : orc_eta0_fct(fl0_all, Nfl0);
    const double a3 = params[o_split + 2], asym = params[o_split + 5];
    double *model = zeros(N), *noise_abs; long n; int rc = 0, Nharvey, l;
    ratios[0][0] = 1;
    for (l = 1; l <= lmax && l <= 3; l++) { Vl[l] = fabs(params[Nmax + l - 1]); orc_amplitude_ratio(l, inclination, ratios[l]); }
    for (n = 0; n < Nmax && !rc; n++) {
        const double a11 = fabs(params[o_split + 6 + n]);
        const double a12 = nl ? fabs(params[o_split + 6 + Nmax + n]) : a11;
        const double a2 = a2a3 ? params[o_split + 6 + (nl ? 2 : 1) * Nmax + n] : 0.0;
        for (l = 0; l <= lmax && l <= 3 && !rc; l++) {
            const int off = Nmax + lmax + (l >= 1 ? Nfl0 : 0) + (l >= 2 ? Nfl1 : 0) + (l >= 3 ? Nfl2 : 0);
            const double fl = params[off + n];
            const double Wl = (l == 0) ? fabs(Wl0_all[n]) : fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl));
            double Hl;
            if (l == 0) Hl = do_amp ? (double)fabsl(params[n] / (pi * Wl)) : fabs(params[n]);
            else Hl = do_amp ? (double)(fabsl(params[n] / (pi * Wl)) * Vl[l]) : fabs(params[n] * Vl[l]);
            if (a2a3) rc = orc_optimum_lorentzian_calc_a1l_a2a3(x, N, &model, Hl, fl, a11, a12, (l == 0) ? 0.0 : a2, (l == 0) ? 0.0 : a3, asym, Wl, l, ratios[l], step, trunc_c);
            else rc = orc_optimum_lorentzian_calc_a1l_etaa3(x, N, &model, Hl, fl, a11, a12, eta0, a3, asym, Wl, l, ratios[l], step, trunc_c);
        }
    }
    if (!rc) {
        noise_abs = (double *)malloc(sizeof(double) * (size_t)(Nnoise > 0 ? Nnoise : 1));
        abs_copy(params + o_split + Nsplit + Nwidth, Nnoise, noise_abs);
        Nharvey = (Nnoise - 1) / 3;
        orc_harvey_like(noise_abs, Nnoise, x, N, &model, Nharvey);
        free(noise_abs);
        memcpy(out, model, sizeof(double) * (size_t)N);
    }
    free(model);
    return rc;
}

/* tamcmc/sources/models.cpp:3012-3196 */
static int model_MS_local_basic(const double *params, const int *pl, const double *x, long N, double *out)
{
    const double step = x[1] - x[0];
    const long double pi = PI_L;
    const int Nmax = pl[0], Nvis = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const double trunc_c = params[Nmax + Nvis + Nf + Nsplit + Nwidth + Nnoise + Ninc];
    const int do_amp = (params[Nmax + Nvis + Nf + Nsplit + Nwidth + Nnoise + Ninc + 1] != 0);
    double inclination, a1, eta0, a3, asym, fl, Wl, Hl;
    double r0[1] = {1}, r1[3], r2[5], r3[7];
    double *model = zeros(N), *noise_abs; long n; int rc = 0;

    inclination = atan(params[Nmax + Nvis + Nf + 4] / params[Nmax + Nvis + Nf + 3]);
    inclination = (double)(inclination * 180. / pi);
    a1 = pow(params[Nmax + Nvis + Nf + 3], 2) + pow(params[Nmax + Nvis + Nf + 4], 2);
    if (Nfl1 >= 1) orc_amplitude_ratio(1, inclination, r1);
    if (Nfl2 >= 1) orc_amplitude_ratio(2, inclination, r2);
    if (Nfl3 >= 1) orc_amplitude_ratio(3, inclination, r3);
    eta0 = params[Nmax + Nvis + Nf + 1];
    a3 = params[Nmax + Nvis + Nf + 2];
    asym = params[Nmax + Nvis + Nf + 5];

    for (n = 0; n < Nfl0 && !rc; n++) {
        fl = params[Nmax + Nvis + n];
        Wl = fabs(params[Nmax + Nvis + Nf + Nsplit + n]);
        if (do_amp) Hl = (double)fabsl(params[n] / (pi * Wl)); else Hl = fabs(params[n]);
        rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl, fl, a1, eta0, a3, asym, Wl, 0, r0, step, trunc_c);
    }
    for (n = 0; n < Nfl1 && !rc; n++) {
        fl = params[Nmax + Nvis + Nfl0 + n];
        Wl = fabs(params[Nmax + Nvis + Nf + Nsplit + Nfl0 + n]);
        if (do_amp) Hl = (double)fabsl(params[Nfl0 + n] / (pi * Wl)); else Hl = fabs(params[Nfl0 + n]);
        rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl, fl, a1, eta0, a3, asym, Wl, 1, r1, step, trunc_c);
    }
    for (n = 0; n < Nfl2 && !rc; n++) {
        fl = params[Nmax + Nvis + Nfl0 + Nfl1 + n];
        Wl = fabs(params[Nmax + Nvis + Nf + Nsplit + Nfl0 + Nfl1 + n]);
        if (do_amp) Hl = (double)fabsl(params[Nfl0 + Nfl1 + n] / (pi * Wl)); else Hl = fabs(params[Nfl0 + Nfl1 + n]);
        rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl, fl, a1, eta0, a3, asym, Wl, 2, r2, step, trunc_c);
    }
    for (n = 0; n < Nfl3 && !rc; n++) {
        fl = params[Nmax + Nvis + Nfl0 + Nfl1 + Nfl2 + n];
        Wl = fabs(params[Nmax + Nvis + Nf + Nsplit + Nfl0 + Nfl1 + Nfl2 + n]);
        if (do_amp) Hl = (double)fabsl(params[Nfl0 + Nfl1 + Nfl2 + n] / (pi * Wl)); else Hl = fabs(params[Nfl0 + Nfl1 + Nfl2 + n]);
        rc = orc_optimum_lorentzian_calc_a1etaa3(x, N, &model, Hl, fl, a1, eta0, a3, asym, Wl, 3, r3, step, trunc_c);
    }
    if (!rc) {
        noise_abs = (double *)malloc(sizeof(double) * (size_t)(Nnoise > 0 ? Nnoise : 1));
        abs_copy(params + Nmax + Nvis + Nf + Nsplit + Nwidth, Nnoise, noise_abs);
        orc_harvey_like(noise_abs, Nnoise, x, N, &model, 0);   /* Nharvey=0: white noise only (models.cpp:3168) */
        free(noise_abs);
        memcpy(out, model, sizeof(double) * (size_t)N);
    }
    free(model);
    return rc;
}

/* tamcmc/sources/models.cpp:3198-3343: model_MS_local_Hnlm.  Heights are H(n,l,|m|), symmetric in m, stored
 * (l+1) per mode at params[base_l + (l+1) n + |m|] with base_l = Nfl0, Nfl0+Nfl1, Nfl0+Nfl1+Nfl2 for l = 1, 2, 3 exactly as the
 * reference indexes them (models.cpp:3270-3272, 3285-3289, 3303-3309); white noise only (Nharvey = 0). */
static int model_MS_local_Hnlm(const double *params, const int *pl, const double *x, long N, double *out)
{
    const double step = x[1] - x[0];
    const long double pi = PI_L;
    const int Nmax = pl[0], Nvis = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const double trunc_c = params[Nmax + Nvis + Nf + Nsplit + Nwidth + Nnoise + Ninc];
    const int do_amp = (params[Nmax + Nvis + Nf + Nsplit + Nwidth + Nnoise + Ninc + 1] != 0);
    const double a1 = fabs(params[Nmax + Nvis + Nf]);
    const double eta0 = params[Nmax + Nvis + Nf + 1], a3 = params[Nmax + Nvis + Nf + 2], asym = params[Nmax + Nvis + Nf + 5];
    double *model = zeros(N), *noise_abs; long n; int rc = 0, l, m;
    const int Nfl[4] = {Nfl0, Nfl1, Nfl2, Nfl3};
    int f_off = 0, h_off = 0;
    for (l = 0; l <= 3 && !rc; l++) {
        for (n = 0; n < Nfl[l] && !rc; n++) {
            const double fl = params[Nmax + Nvis + f_off + n];
            const double Wl = fabs(params[Nmax + Nvis + Nf + Nsplit + f_off + n]);
            const int pos0 = (l + 1) * (int)n;
            double Hlm[7];
            for (m = -l; m <= l; m++) {
                double h = params[h_off + pos0 + (m < 0 ? -m : m)];
                if (do_amp) h = (l == 0) ? (double)(h / (pi * Wl)) : (double)(h / (pi * Wl));
                Hlm[m + l] = fabs(h);
            }
            rc = orc_optimum_lorentzian_calc_a1etaa3_v2(x, N, &model, Hlm, fl, a1, eta0, a3, asym, Wl, l, step, trunc_c);
        }
        f_off += Nfl[l];
        h_off += Nfl[l];      /* the reference indexes the l>=2 heights from Nfl0+Nfl1(+Nfl2), not from Nfl0+2*Nfl1(+3*Nfl2): models.cpp:3285-3289, 3303-3309 */
    }
    if (!rc) {
        noise_abs = (double *)malloc(sizeof(double) * (size_t)(Nnoise > 0 ? Nnoise : 1));
        abs_copy(params + Nmax + Nvis + Nf + Nsplit + Nwidth, Nnoise, noise_abs);
        orc_harvey_like(noise_abs, Nnoise, x, N, &model, 0);   /* Nharvey = 0 (models.cpp:3329) */
        free(noise_abs);
        memcpy(out, model, sizeof(double) * (size_t)N);
    }
    free(model);
    return rc;
}

/* tamcmc/sources/models.cpp:1195-1408.  The four omp-parallel loops (l=0, then 1, 2, 3)
 * are run serially in index order. */
static int model_MS_Global_aj_HarveyLike(const double *params, const int *pl, const double *x, long N, double *out)
{
    const double step = x[1] - x[0];
    const long double pi = M_PI;
    const int Nmax = pl[0], lmax = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const double trunc_c = params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc];
    const int do_amp = (params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc + 1] != 0);
    double inclination;
    double r0[1] = {1}, r1[3], r2[5], r3[7];
    double Vl1 = 0, Vl2 = 0, Vl3 = 0, Hl, Wl, fl, a1, a2, a3, a4, a5, a6, eta0, asym;
    const double *fl0_all = params + Nmax + lmax;
    const double *Wl0_all = params + Nmax + lmax + Nf + Nsplit;
    const double *Hl0_all = params;
    const double *a1_terms = params + Nmax + lmax + Nf;
    const double *a2_terms = a1_terms + 2, *a3_terms = a1_terms + 4, *a4_terms = a1_terms + 6;
    const double *a5_terms = a1_terms + 8, *a6_terms = a1_terms + 10;
    double *model = zeros(N), *noise_abs; int n, rc = 0, Nharvey;
    orc_Optim_L blk;

    inclination = params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise];
    if (lmax >= 1) { Vl1 = fabs(params[Nmax]); orc_amplitude_ratio(1, inclination, r1); }
    if (lmax >= 2) { Vl2 = fabs(params[Nmax + 1]); orc_amplitude_ratio(2, inclination, r2); }
    if (lmax >= 3) { Vl3 = fabs(params[Nmax + 2]); orc_amplitude_ratio(3, inclination, r3); }
    asym = params[Nmax + lmax + Nf + 13];
    if (params[Nmax + lmax + Nf + 12] == 1) eta0 = orc_eta0_fct(fl0_all, Nfl0); else eta0 = 0;

    for (n = 0; n < Nfl0 && !rc; n++) {
        fl = fl0_all[n];
        Wl = fabs(Wl0_all[n]);
        if (do_amp) Hl = (double)fabsl(params[n] / (pi * Wl)); else Hl = fabs(params[n]);
        rc = orc_optimum_lorentzian_calc_aj(x, N, Hl, fl, 0, 0, 0, 0, 0, 0, 0, asym, Wl, 0, r0, step, trunc_c, &blk);
        if (!rc) add_block(model, &blk);
    }
    for (n = 0; n < Nfl1 && !rc; n++) {
        fl = params[Nmax + lmax + Nfl0 + n];
        Wl = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl));
        if (do_amp) Hl = (double)fabsl(orc_lin_interpol(fl0_all, Hl0_all, Nmax, fl) / (pi * Wl) * Vl1);
        else Hl = fabs(orc_lin_interpol(fl0_all, Hl0_all, Nmax, fl) * Vl1);
        a1 = a1_terms[0] + a1_terms[1] * (fl * 1e-3);
        a2 = a2_terms[0] + a2_terms[1] * (fl * 1e-3);
        rc = orc_optimum_lorentzian_calc_aj(x, N, Hl, fl, a1, a2, 0, 0, 0, 0, eta0, asym, Wl, 1, r1, step, trunc_c, &blk);
        if (!rc) add_block(model, &blk);
    }
    for (n = 0; n < Nfl2 && !rc; n++) {
        fl = params[Nmax + lmax + Nfl0 + Nfl1 + n];
        Wl = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl));
        if (do_amp) Hl = (double)fabsl(orc_lin_interpol(fl0_all, Hl0_all, Nmax, fl) / (pi * Wl) * Vl2);
        else Hl = fabs(orc_lin_interpol(fl0_all, Hl0_all, Nmax, fl) * Vl2);
        a1 = a1_terms[0] + a1_terms[1] * (fl * 1e-3);
        a2 = a2_terms[0] + a2_terms[1] * (fl * 1e-3);
        a3 = a3_terms[0] + a3_terms[1] * (fl * 1e-3);
        a4 = a4_terms[0] + a4_terms[1] * (fl * 1e-3);
        rc = orc_optimum_lorentzian_calc_aj(x, N, Hl, fl, a1, a2, a3, a4, 0, 0, eta0, asym, Wl, 2, r2, step, trunc_c, &blk);
        if (!rc) add_block(model, &blk);
    }
    for (n = 0; n < Nfl3 && !rc; n++) {
        fl = params[Nmax + lmax + Nfl0 + Nfl1 + Nfl2 + n];
        Wl = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl));
        if (do_amp) Hl = (double)fabsl(orc_lin_interpol(fl0_all, Hl0_all, Nmax, fl) / (pi * Wl) * Vl3);
        else Hl = fabs(orc_lin_interpol(fl0_all, Hl0_all, Nmax, fl) * Vl3);
        a1 = a1_terms[0] + a1_terms[1] * (fl * 1e-3);
        a2 = a2_terms[0] + a2_terms[1] * (fl * 1e-3);
        a3 = a3_terms[0] + a3_terms[1] * (fl * 1e-3);
        a4 = a4_terms[0] + a4_terms[1] * (fl * 1e-3);
        a5 = a5_terms[0] + a5_terms[1] * (fl * 1e-3);
        a6 = a6_terms[0] + a6_terms[1] * (fl * 1e-3);
        rc = orc_optimum_lorentzian_calc_aj(x, N, Hl, fl, a1, a2, a3, a4, a5, a6, eta0, asym, Wl, 3, r3, step, trunc_c, &blk);
        if (!rc) add_block(model, &blk);
    }
    if (!rc) {
        noise_abs = (double *)malloc(sizeof(double) * (size_t)(Nnoise > 0 ? Nnoise : 1));
        abs_copy(params + Nmax + lmax + Nf + Nsplit + Nwidth, Nnoise, noise_abs);
        Nharvey = (Nnoise - 1) / 3;
        orc_harvey_like(noise_abs, Nnoise, x, N, &model, Nharvey);
        free(noise_abs);
        memcpy(out, model, sizeof(double) * (size_t)N);
    }
    free(model);
    return rc;
}

/* models.cpp:6110-6126 decompose_Alm_fct_GSLgrid with the Alm value supplied by `alm`.
 * fc_l, eta0, a1, epsilon_nl arrive as long double in the reference. */
static void decompose_Alm(int l, long double fc_l, long double eta0, long double a1, long double epsilon_nl,
                          const double thetas[2], int filter_code, orc_alm_fn alm, void *user, double aj[6])
{
    double nu[7]; int m;
    for (m = -l; m <= l; m++) {
        nu[m + l] = (double)fc_l;
        if (eta0 > 0) nu[m + l] = (double)(nu[m + l] + fc_l * eta0 * orc_Qlm(l, m) * powl(a1 * 1e-6, 2));
        nu[m + l] = (double)(nu[m + l] + fc_l * epsilon_nl * alm(l, m, thetas[0], thetas[1], filter_code, user));
    }
    orc_eval_acoefs(l, nu, aj);
}

/* tamcmc/sources/models.cpp:1411-1746 */
static int model_MS_Global_ajAlm_HarveyLike(const double *params, const int *pl, const double *x, long N, double *out,
                                            orc_alm_fn alm, void *user)
{
    const double step = x[1] - x[0];
    const long double pi = M_PI;
    const int Nmax = pl[0], lmax = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const int cfg = Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc;
    const double trunc_c = params[cfg];
    const int do_amp = (params[cfg + 1] != 0);
    const int decompose = (int)params[cfg + 2];
    const int filter_code = (int)params[cfg + 3];
    double inclination, thetas[2], aj[6], Alm_m[7];
    double r0[1] = {1}, r1[3], r2[5], r3[7];
    double Vl1 = 0, Vl2 = 0, Vl3 = 0, Hl, Wl, fl, a1, a3, a5, eta0, eps, asym;
    const double *fl0_all = params + Nmax + lmax;
    const double *Wl0_all = params + Nmax + lmax + Nf + Nsplit;
    const double *Hl0_all = params;
    const double *a1_terms = params + Nmax + lmax + Nf;
    const double *a3_terms = a1_terms + 2, *a5_terms = a1_terms + 4, *eps_terms = a1_terms + 6;
    double *model; int n, m, rc = 0, Nharvey, l;
    const double *V; double Vl;
    orc_Optim_L blk;

    if (filter_code != 0 && filter_code != 2) return ORC_ERR_MODEL;      /* models.cpp:1444-1465 */
    if (decompose < -1 || decompose > 2) return ORC_ERR_MODEL;           /* models.cpp:1601-1603 */
    if (!alm) return ORC_ERR_ARG;
    model = zeros(N);
    thetas[0] = params[Nmax + lmax + Nf + 8] * M_PI / 180.;
    thetas[1] = params[Nmax + lmax + Nf + 9] * M_PI / 180.;
    asym = params[Nmax + lmax + Nf + 11];
    inclination = params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise];
    if (lmax >= 1) { Vl1 = fabs(params[Nmax]); orc_amplitude_ratio(1, inclination, r1); }
    if (lmax >= 2) { Vl2 = fabs(params[Nmax + 1]); orc_amplitude_ratio(2, inclination, r2); }
    if (lmax >= 3) { Vl3 = fabs(params[Nmax + 2]); orc_amplitude_ratio(3, inclination, r3); }
    if (params[Nmax + lmax + Nf + 10] == 1) eta0 = orc_eta0_fct(fl0_all, Nfl0); else eta0 = 0;

    for (n = 0; n < Nfl0 && !rc; n++) {
        fl = fl0_all[n];
        Wl = fabs(Wl0_all[n]);
        if (do_amp) Hl = (double)fabsl(params[n] / (pi * Wl)); else Hl = fabs(params[n]);
        rc = orc_optimum_lorentzian_calc_aj(x, N, Hl, fl, 0, 0, 0, 0, 0, 0, 0, asym, Wl, 0, r0, step, trunc_c, &blk);
        if (!rc) add_block(model, &blk);
    }
    for (l = 1; l <= 3 && !rc; l++) {
        const int Nfl = (l == 1) ? Nfl1 : (l == 2) ? Nfl2 : Nfl3;
        const int off = Nmax + lmax + Nfl0 + (l >= 2 ? Nfl1 : 0) + (l >= 3 ? Nfl2 : 0);
        V = (l == 1) ? r1 : (l == 2) ? r2 : r3;
        Vl = (l == 1) ? Vl1 : (l == 2) ? Vl2 : Vl3;
        for (n = 0; n < Nfl && !rc; n++) {
            fl = params[off + n];
            Wl = fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fl));
            if (do_amp) Hl = (double)fabsl(orc_lin_interpol(fl0_all, Hl0_all, Nmax, fl) / (pi * Wl) * Vl);
            else Hl = fabs(orc_lin_interpol(fl0_all, Hl0_all, Nmax, fl) * Vl);
            a1 = a1_terms[0] + a1_terms[1] * (fl * 1e-3);
            a3 = (l >= 2) ? a3_terms[0] + a3_terms[1] * (fl * 1e-3) : 0;
            a5 = (l >= 3) ? a5_terms[0] + a5_terms[1] * (fl * 1e-3) : 0;
            eps = eps_terms[0] + eps_terms[1] * (fl * 1e-3);
            if (decompose == -1) {
                for (m = -l; m <= l; m++) Alm_m[m + l] = alm(l, m, thetas[0], thetas[1], filter_code, user);
                rc = optimum_lorentzian_calc_ajAlm(x, N, Hl, fl, a1, a3, a5, eta0, eps, Alm_m, asym, Wl, l, V, step, trunc_c, &blk);
            } else {
                double a2, a4, a6;
                decompose_Alm(l, fl, eta0, a1, eps, thetas, filter_code, alm, user, aj);
                a2 = aj[1];
                /* which even coefficients reach the profile: models.cpp:1567-1600 (l=1), 1631-1664 (l=2), 1694-1720 (l=3) */
                a4 = (l >= 2 && decompose != 2) ? aj[3] : 0;
                a6 = (l >= 3 && decompose == 0) ? aj[5] : 0;
                rc = orc_optimum_lorentzian_calc_aj(x, N, Hl, fl, a1, a2, a3, a4, a5, a6, eta0, asym, Wl, l, V, step, trunc_c, &blk);
            }
            if (!rc) add_block(model, &blk);
        }
    }
    if (!rc) {
        double *noise_abs = (double *)malloc(sizeof(double) * (size_t)(Nnoise > 0 ? Nnoise : 1));
        abs_copy(params + Nmax + lmax + Nf + Nsplit + Nwidth, Nnoise, noise_abs);
        Nharvey = (Nnoise - 1) / 3;
        orc_harvey_like(noise_abs, Nnoise, x, N, &model, Nharvey);
        free(noise_abs);
        memcpy(out, model, sizeof(double) * (size_t)N);
    }
    free(model);
    return rc;
}

/* Model_def::call_model switch, tamcmc/sources/model_def.cpp:220-388 */
/* ------------------------------------------------------------------------------------------------
 * Gaussian-envelope models (no Lorentzians)
 * ------------------------------------------------------------------------------------------------ */
/* tamcmc/sources/models.cpp:5674-5725: model_Harvey_Gaussian.  params = [H1, tc1, p1, H2, tc2, p2, B0, Hgauss, nu, sigma];
 * Gaussian first (:5693-5694), then harvey_like on |params[0..6]| with Nharvey = 2 (:5698-5700). */
static int model_Harvey_Gaussian(const double *params, const double *x, long N, double *out)
{
    double *model = (double *)malloc(sizeof(double) * (size_t)N), noise_abs[7];
    const double sig2 = pow(fabs(params[9]), 2);
    long i; int k;
    for (i = 0; i < N; i++) { const double d = x[i] - params[8]; model[i] = -0.5 * (d * d) / sig2; }
    for (i = 0; i < N; i++) model[i] = fabs(params[7]) * exp(model[i]);
    for (k = 0; k < 7; k++) noise_abs[k] = fabs(params[k]);
    orc_harvey_like(noise_abs, 7, x, N, &model, 2);
    memcpy(out, model, sizeof(double) * (size_t)N);
    free(model);
    return 0;
}

/* tamcmc/sources/noise_models.cpp:65-84: get_ksinorm, trapezoid rule over the whole (equally spaced) x */
static double get_ksinorm(double b, double c, const double *x, long N)
{
    double integral = 0.0;
    const double h = x[1] - x[0];
    long i;
    for (i = 0; i < N; i++) {
        const double term = 1.0 / (1.0 + pow(x[i] / b, c));
        if (i == 0 || i == N - 1) integral += 0.5 * term;
        else integral += term;
    }
    integral *= h;
    return b / integral;
}

/* tamcmc/sources/models.cpp:5728-5797 + noise_models.cpp:87-150: model_Kallinger2014_Gaussian.
 * params = [k_a, s_a, k_b0, s_b0, c0, a1, a2, k1, s1, c1, k2, s2, c2, N0, Amax, numax, sigma, mu_numax].
 * Gaussian times the sinc^2 leakage (:5768-5772); white noise and three normalised super-Lorentzians (noise_models.cpp:128-146).
 * The last statement of Kallinger2014 (`Power.cwiseProduct(eta_squared);`, noise_models.cpp:148) discards its result: the noise
 * is NOT attenuated.  (The reference also forces outparams and rewrites params.model on every call, :5764 -- a file side effect
 * outside this path.) */
static int model_Kallinger2014_Gaussian(const double *params, const double *x, long N, double *out)
{
    const double Amax = fabs(params[14]), numax = fabs(params[15]), sig_numax = fabs(params[16]), mu_numax = params[17];
    const double *np = params;
    double x_nyquist = x[0];
    double a0, b0, c0, a1, a2, b1, b2, c1, c2, N0, ksi0, ksi1, ksi2, sig2, f0, f1, f2;
    long i;
    for (i = 1; i < N; i++) if (x[i] > x_nyquist) x_nyquist = x[i];
    sig2 = pow(fabs(sig_numax), 2);
    for (i = 0; i < N; i++) {
        /* eta_squared_Kallinger2014, noise_models.cpp:87-97 */
        double eta = sin(0.5 * M_PI * x[i] / x_nyquist) / (0.5 * M_PI * x[i] / x_nyquist);
        double g;
        if (i == 0 && x[0] == 0) eta = 1;
        g = -0.5 * ((x[i] - numax) * (x[i] - numax)) / sig2;
        out[i] = fabs(Amax) * (eta * eta) * exp(g);
    }
    a0 = fabs(np[0] * pow(fabs(numax), np[1]));
    b0 = fabs(np[2] * pow(fabs(numax + mu_numax), np[3]));
    c0 = fabs(np[4]);
    a1 = np[5];
    a2 = np[6];
    b1 = fabs(np[7] * pow(fabs(numax + mu_numax), np[8]));
    b2 = fabs(np[10] * pow(fabs(numax + mu_numax), np[11]));
    c1 = fabs(np[9]);
    c2 = fabs(np[12]);
    N0 = fabs(np[13]);
    ksi0 = get_ksinorm(b0, c0, x, N);
    ksi1 = get_ksinorm(b1, c1, x, N);
    ksi2 = get_ksinorm(b2, c2, x, N);
    f0 = ksi0 * pow(a0, 2) / b0; f1 = ksi1 * pow(a1, 2) / b1; f2 = ksi2 * pow(a2, 2) / b2;
    for (i = 0; i < N; i++) {
        double P = out[i] + N0;
        P = P + f0 * (1.0 / (pow(x[i] / b0, c0) + 1.0));
        P = P + f1 * (1.0 / (pow(x[i] / b1, c1) + 1.0));
        P = P + f2 * (1.0 / (pow(x[i] / b2, c2) + 1.0));
        out[i] = P;
    }
    return 0;
}

int orc_call_model(int model_id, const double *params, const int *plength, const double *x, long N,
                   double *model_out, orc_alm_fn alm, void *alm_user)
{
    if (N < 2) return ORC_ERR_ARG;
    switch (model_id) {
    case ORC_MODEL_KALLINGER2014_GAUSSIAN: return N < 3 ? ORC_ERR_ARG : model_Kallinger2014_Gaussian(params, x, N, model_out);
    case ORC_MODEL_HARVEY_GAUSSIAN:      return model_Harvey_Gaussian(params, x, N, model_out);
    case ORC_MODEL_MS_GLOBAL_CLASSIC:    return model_MS_Global_a1etaa3_HarveyLike_Classic(params, plength, x, N, model_out);
    case ORC_MODEL_MS_GLOBAL_CLASSIC_V2: return model_MS_Global_a1etaa3_HarveyLike_Classic_v2(params, plength, x, N, model_out);
    case ORC_MODEL_MS_GLOBAL_CLASSIC_V3: return model_MS_Global_a1etaa3_HarveyLike_Classic_v3(params, plength, x, N, model_out);
    case ORC_MODEL_MS_GLOBAL_A1L_ETAA3:  return model_MS_Global_a1l_etaa3_HarveyLike(params, plength, x, N, model_out);
    case ORC_MODEL_MS_GLOBAL_A1N_ETAA3: case ORC_MODEL_MS_GLOBAL_A1NL_ETAA3:
        return model_MS_Global_a1x_HarveyLike(params, plength, x, N, model_out, model_id);
    /* ORC_MODEL_MS_GLOBAL_A1N_A2A3 (18) and ORC_MODEL_MS_GLOBAL_A1NL_A2A3 (19) compute a model and then print "not tested yet"
     * and exit in the reference (models.cpp:599-603, 993-997): unusable there, ORC_ERR_MODEL here */
    case ORC_MODEL_MS_LOCAL_BASIC:       return model_MS_local_basic(params, plength, x, N, model_out);
    case ORC_MODEL_MS_LOCAL_HNLM:        return model_MS_local_Hnlm(params, plength, x, N, model_out);
    case ORC_MODEL_MS_GLOBAL_AJ:         return model_MS_Global_aj_HarveyLike(params, plength, x, N, model_out);
    case ORC_MODEL_MS_GLOBAL_AJALM:      return model_MS_Global_ajAlm_HarveyLike(params, plength, x, N, model_out, alm, alm_user);
    default: return ORC_ERR_MODEL;
    }
}

/* One MCMC-step worth of likelihood evaluations: threads over chains as MALA.cpp:648-668,
 * each chain = call_model + call_likelihood (model_def.cpp:473-474). */
int orc_eval_chains(int model_id, const double *params, int Nparams, const int *plength,
                    const double *x, const double *y, long N, int Nchains, const double *Tcoefs, double p,
                    double *logL_out, int nthreads)
{
    int rc_all = 0, c;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (c = 0; c < Nchains; c++) {
        double *model = (double *)malloc(sizeof(double) * (size_t)N);
        int rc = orc_call_model(model_id, params + (size_t)c * Nparams, plength, x, N, model, 0, 0);
        if (rc) {
            logL_out[c] = NAN;
#pragma omp critical
            rc_all = rc;
        } else {
            logL_out[c] = (double)orc_call_likelihood_chi22p(y, model, N, p, Tcoefs[c]);
        }
        free(model);
    }
    return rc_all;
}

/* call_model + call_likelihood case 1 (chi_square, model_def.cpp:403-406): likelihood_chi_square(y, model, sigma_y) / Tcoefs[m] */
int orc_eval_chains_chi_square(int model_id, const double *params, int Nparams, const int *plength, const double *x,
                               const double *y, const double *sigma, long N, int Nchains, const double *Tcoefs,
                               double *logL_out, int nthreads)
{
    int rc_all = 0, c;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (c = 0; c < Nchains; c++) {
        double *model = (double *)malloc(sizeof(double) * (size_t)N);
        int rc = orc_call_model(model_id, params + (size_t)c * Nparams, plength, x, N, model, 0, 0);
        if (rc) {
            logL_out[c] = NAN;
#pragma omp critical
            rc_all = rc;
        } else logL_out[c] = (double)(orc_likelihood_chi_square(y, model, sigma, N) / Tcoefs[c]);
        free(model);
    }
    return rc_all;
}

/* ------------------------------------------------------------------------- */
/* generic mode table (include/tamcmc_gpu.h: TAMCMC_MODEL_MODE_TABLE)         */
/* ------------------------------------------------------------------------- */
/* What every model function of the aj family does AFTER it has resolved its mode list on the host: one
 * optimum_lorentzian_calc_aj call per mode (e.g. models.cpp:4931-5006 for model_RGB_asympt_aj_AppWidth_HarveyLike_v4,
 * models.cpp:1287-1376 for model_MS_Global_aj_HarveyLike), accumulated into model_final, then
 * harvey_like(noise_params.array().abs(), ...) (models.cpp:5011-5017).  A mode whose per-m extra shifts are not all
 * zero follows build_l_mode_ajAlm's order instead (build_lorentzian.cpp:182-190): nu = fc + a1 P1 + a3 P3 + a5 P5,
 * + centrifugal term, + extra[m] (= fc*epsilon_nl*Alm(l,m), supplied by the caller).
 * row = [nmodes, inclination, trunc_c, asym, noise[Nnoise], nmodes x {l, fc, H, W, a1..a6, eta0, extra[-3..3], 0, 0}] */
#define ORC_MT_HDR 4
#define ORC_MT_STRIDE 20
int orc_mode_table_model(const double *row, int Nnoise, int step_mode, const double *x, long N, double *out)
{
    const int nmodes = (int)row[0];
    const double inclination = row[1], trunc_c = row[2], asym = row[3];
    const double step = step_mode ? x[2] - x[1] : x[1] - x[0];     /* models.cpp:4714 vs models.cpp:1952 */
    const double *modes = row + ORC_MT_HDR + Nnoise;
    double ratios[4][7];
    double *model = zeros(N);
    int j, l, m, rc = 0, Nharvey;
    ratios[0][0] = 1;
    for (l = 1; l <= 3; l++) orc_amplitude_ratio(l, inclination, ratios[l]);
    for (j = 0; j < nmodes && !rc; j++) {
        const double *r = modes + (size_t)j * ORC_MT_STRIDE;
        orc_Optim_L blk;
        int has_extra = 0;
        l = (int)r[0];
        if (l < 0 || l > 3) { rc = ORC_ERR_ARG; break; }
        for (m = 0; m < 7; m++) if (r[11 + m] != 0) has_extra = 1;
        if (!has_extra) {
            rc = orc_optimum_lorentzian_calc_aj(x, N, r[2], r[1], r[4], r[5], r[6], r[7], r[8], r[9], r[10], asym, r[3], l,
                                                ratios[l], step, trunc_c, &blk);
        } else {
            /* epsilon_nl * Alm_m * fc == extra: pass fc-normalised values so the product restores extra[m] */
            double Alm_m[7];
            for (m = -l; m <= l; m++) Alm_m[m + l] = r[11 + 3 + m];
            {
                int iv[2]; long nw; double *x_l;
                blk.y = 0; blk.i0 = 0; blk.N = 0;
                rc = orc_set_imin_imax(x, N, l, r[1], r[3], r[4], trunc_c, step, iv);
                trace_push(l, iv);
                if (!rc) {
                    scratch_t sc; long i;
                    nw = iv[1] - iv[0];
                    x_l = (double *)malloc(sizeof(double) * (size_t)nw);
                    blk.y = (double *)malloc(sizeof(double) * (size_t)nw);
                    memcpy(x_l, x + iv[0], sizeof(double) * (size_t)nw);
                    scratch_alloc(&sc, nw);
                    for (i = 0; i < nw; i++) blk.y[i] = 0;
                    for (m = -l; m <= l; m++) {
                        double nu;
                        if (l != 0) {
                            nu = r[1] + r[4] * orc_Pslm(1, l, m) + r[6] * orc_Pslm(3, l, m) + r[8] * orc_Pslm(5, l, m);
                            if (r[10] > 0) nu = nu + r[1] * r[10] * orc_Qlm(l, m) * pow(r[4] * 1e-6, 2);
                            nu = nu + Alm_m[m + l];
                        } else nu = r[1];
                        add_component(x_l, nw, nu, r[2] * ratios[l][m + l], r[1], asym, r[3], sc.profile, sc.tmp, sc.tmp2, sc.asymetry, blk.y);
                    }
                    scratch_free(&sc);
                    free(x_l);
                    blk.i0 = iv[0]; blk.N = (int)nw;
                }
            }
        }
        if (!rc) add_block(model, &blk);
    }
    if (!rc) {
        double *noise_abs = (double *)malloc(sizeof(double) * (size_t)(Nnoise > 0 ? Nnoise : 1));
        abs_copy(row + ORC_MT_HDR, Nnoise, noise_abs);
        Nharvey = (Nnoise - 1) / 3;
        orc_harvey_like(noise_abs, Nnoise, x, N, &model, Nharvey);
        free(noise_abs);
        memcpy(out, model, sizeof(double) * (size_t)N);
    }
    free(model);
    return rc;
}

int orc_mode_table_eval_chains(const double *rows, int row_stride, int Nnoise, int step_mode, const double *x, const double *y,
                               long N, int Nchains, const double *Tcoefs, double p, double *logL_out, int nthreads)
{
    int rc_all = 0, c;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (c = 0; c < Nchains; c++) {
        double *model = (double *)malloc(sizeof(double) * (size_t)N);
        int rc = orc_mode_table_model(rows + (size_t)c * row_stride, Nnoise, step_mode, x, N, model);
        if (rc) {
            logL_out[c] = NAN;
#pragma omp critical
            rc_all = rc;
        } else logL_out[c] = (double)orc_call_likelihood_chi22p(y, model, N, p, Tcoefs[c]);
        free(model);
    }
    return rc_all;
}

/* ------------------------------------------------------------------------- */
/* best-effort CPU variant (second CPU baseline line of bench.py only)       */
/* ------------------------------------------------------------------------- */
/* Same per-element arithmetic as the reference-faithful path (each mode's m-components are summed in m
 * order into a local value that is then added to the model, then the Harvey terms, then the Whittle sum),
 * but window-only, single pass per mode, no full-vector copies and no temporaries.  Implemented for
 * model_MS_Global_a1etaa3_HarveyLike_Classic (models.cpp:1943-2121); other models use the faithful path. */
static int classic_fast(const double *params, const int *pl, const double *x, const double *y, long N,
                        double *model, double p, double Tcoef, double *logL)
{
    const double step = x[1] - x[0];
    const long double pi = PI_L;
    const int Nmax = pl[0], lmax = pl[1], Nfl0 = pl[2], Nfl1 = pl[3], Nfl2 = pl[4], Nfl3 = pl[5];
    const int Nsplit = pl[6], Nwidth = pl[7], Nnoise = pl[8], Ninc = pl[9];
    const int Nf = Nfl0 + Nfl1 + Nfl2 + Nfl3;
    const int do_amp = (params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc + 1] != 0);
    const double trunc_c = params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc];
    const double inclination = params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise];
    const double *fl0_all = params + Nmax + lmax, *Wl0_all = params + Nmax + lmax + Nf + Nsplit;
    const double a1 = fabs(params[Nmax + lmax + Nf]), a3 = params[Nmax + lmax + Nf + 2], asym = params[Nmax + lmax + Nf + 5];
    const double eta0 = orc_eta0_fct(fl0_all, Nfl0);
    double ratios[4][7], Vl[4] = {1, 0, 0, 0};
    long n, i; int l, m, iv[2], rc, k;
    ratios[0][0] = 1;
    for (l = 1; l <= lmax && l <= 3; l++) { Vl[l] = fabs(params[Nmax + l - 1]); orc_amplitude_ratio(l, inclination, ratios[l]); }
    for (i = 0; i < N; i++) model[i] = 0;
    for (n = 0; n < Nmax; n++)
        for (l = 0; l <= lmax; l++) {
            const int off = Nmax + lmax + (l >= 1 ? Nfl0 : 0) + (l >= 2 ? Nfl1 : 0) + (l >= 3 ? Nfl2 : 0);
            const double fc = params[off + n];
            const double W = (l == 0) ? fabs(Wl0_all[n]) : fabs(orc_lin_interpol(fl0_all, Wl0_all, Nmax, fc));
            double H, nu[7], hv[7];
            if (l == 0) H = do_amp ? (double)fabsl(params[n] / (pi * W)) : fabs(params[n]);
            else H = do_amp ? (double)(fabsl(params[n] / (pi * W)) * Vl[l]) : fabs(params[n] * Vl[l]);
            rc = orc_set_imin_imax(x, N, l, fc, W, a1, trunc_c, step, iv);
            if (rc) return rc;
            for (m = -l; m <= l; m++) {
                nu[m + l] = (l != 0) ? fc * (1. + eta0 * pow(a1 * 1e-6, 2) * orc_Qlm(l, m)) + m * a1 + orc_Pslm(3, l, m) * a3 : fc;
                hv[m + l] = H * ratios[l][m + l];
            }
            {
                const double g2 = pow(W, 2), k2 = 0.5 * W * asym / fc;
                for (i = iv[0]; i < iv[1]; i++) {
                    double r = 0;
                    for (k = 0; k <= 2 * l; k++) {
                        const double d = x[i] - nu[k];
                        const double prof = 4 * (d * d) / g2;
                        if (asym == 0) r = r + hv[k] * (1.0 / (1 + prof));
                        else { const double w = 1 + asym * (x[i] / fc - 1); r = r + hv[k] * ((w * w + k2 * k2) * (1.0 / (1 + prof))); }
                    }
                    model[i] = model[i] + r;
                }
            }
        }
    {
        const double *np_ = params + Nmax + lmax + Nf + Nsplit + Nwidth;
        const int Nharvey = (Nnoise - 1) / 3;
        const double white = fabs(np_[Nnoise - 1]);
        double s1 = 0, s2 = 0;
        for (i = 0; i < N; i++) {
            double v = model[i];
            for (k = 0; k < Nharvey; k++)
                if (np_[3 * k + 1] != 0) v = v + fabs(np_[3 * k]) * (1.0 / (pow((1e-3) * fabs(np_[3 * k + 1]) * x[i], fabs(np_[3 * k + 2])) + 1));
            v = v + white;
            s1 += y[i] * (1.0 / v);
            s2 += log(v);
        }
        *logL = (double)(((long double)(-(long)p) * (long double)(s1 + s2)) / Tcoef);
    }
    return ORC_OK;
}

int orc_eval_chains_fast(int model_id, const double *params, int Nparams, const int *plength,
                         const double *x, const double *y, long N, int Nchains, const double *Tcoefs, double p,
                         double *logL_out, int nthreads)
{
    int rc_all = 0, c;
    if (model_id != ORC_MODEL_MS_GLOBAL_CLASSIC)
        return orc_eval_chains(model_id, params, Nparams, plength, x, y, N, Nchains, Tcoefs, p, logL_out, nthreads);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (c = 0; c < Nchains; c++) {
        double *model = (double *)malloc(sizeof(double) * (size_t)N);
        int rc = classic_fast(params + (size_t)c * Nparams, plength, x, y, N, model, p, Tcoefs[c], &logL_out[c]);
        if (rc) {
            logL_out[c] = NAN;
#pragma omp critical
            rc_all = rc;
        }
        free(model);
    }
    return rc_all;
}
