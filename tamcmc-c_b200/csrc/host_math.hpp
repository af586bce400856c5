// host_math.hpp -- scalar host-side restatements shared by capi.cu (constant tables of the device expander) and
// host_expand.cpp (host expanders that turn a model's parameter vector into a mode table).
// New code, same arithmetic as the reference functions cited at each definition (paths relative to the upstream repo).
#pragma once
#include <cmath>

namespace tamcmc_host {

// Ritzwoller & Lavely (1991) polynomials normalised so that Pslm(l)=l (Schou, Christensen-Dalsgaard & Thompson 1994),
// as used by tamcmc/sources/acoefs.cpp:51-110 (H for s = 5, 6: acoefs.cpp:19-49)
inline long double Hslm(int s, int l, int m)
{
    const double L = (double)(l * (l + 1)), M = (double)m;
    switch (s) {
    case 5: return 252 * std::pow(M, 5) - 140 * (2 * L - 3) * std::pow(M, 3) + (20 * L * (3 * L - 10) + 48) * M;
    case 6: return 924 * std::pow(M, 6) - 420 * std::pow(M, 4) * (3 * L - 7) + 84 * std::pow(M, 2) * (5 * L * L - 25 * L + 14)
                 - 20 * L * (L * L - 8 * L + 12);
    }
    return 0;
}
inline long double Pslm(int s, int l, int m)
{
    const double M = (double)m, dl = (double)l;
    const int LL = l * (l + 1);
    long double H, c;
    switch (s) {
    case 1: return m;
    case 2: return (l > 0) ? (long double)((3 * M * M - LL) / (2 * l - 1)) : 0.0L;
    case 3: return (l > 1) ? (long double)((5 * M * M * M - (3 * LL - 1) * M) / ((l - 1) * (2 * l - 1))) : 0.0L;
    case 4:
        H = (35 * std::pow(M, 4) - 5 * (6 * LL - 5) * M * M) + 3 * LL * (LL - 2);
        c = 2 * (l - 1) * (2 * l - 1) * (2 * l - 3);
        return (c != 0) ? H / c : 0.0L;
    case 5:
        H = Hslm(5, l, m);
        c = 8 * (4 * std::pow(dl, 4) - 20 * std::pow(dl, 3) + 35 * dl * dl - 25 * dl + 6);
        return (c != 0) ? H / c : 0.0L;
    case 6:
        H = Hslm(6, l, m);
        c = 64 * std::pow(dl, 5) - 480 * std::pow(dl, 4) + 1360 * std::pow(dl, 3) - 1800 * dl * dl + 1096 * dl - 240;
        return (c != 0) ? H / c : 0.0L;
    }
    return 0.0L;
}
// tamcmc/sources/build_lorentzian.cpp:583-592 (the 2/3 factor is held as long double there)
inline double Qlm(int l, int m)
{
    const long double Dnl = 2. / 3;
    double Q = (l * (l + 1) - 3 * (double)m * (double)m) / ((2 * l - 1) * (2 * l + 3));
    Q = (double)(Q * Dnl);
    return Q;
}
// tamcmc/sources/function_rot.cpp:90-101: int factorial and combi with INTEGER divisions
inline int fact_i(int n) { long f = 1; for (long i = 1; i <= n; i++) f *= i; return (int)f; }
inline int combi_i(int n, int r) { return fact_i(n) / fact_i(n - r) / fact_i(r); }

// tamcmc/sources/interpol.cpp:13-43
inline double lin_interpol(const double* x, const double* y, int Nx, double x_int)
{
    int i = 0;
    double a = 0, b = 0;
    if (x_int >= x[0] && x_int <= x[Nx - 1]) {
        while ((x_int < x[i] || x_int > x[i + 1]) && i < Nx - 2) i = i + 1;
        a = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
        b = y[i] - a * x[i];
    }
    if (x_int < x[0]) { a = (y[1] - y[0]) / (x[1] - x[0]); b = y[0] - a * x[0]; }
    if (x_int > x[Nx - 1]) { a = (y[Nx - 1] - y[Nx - 2]) / (x[Nx - 1] - x[Nx - 2]); b = y[Nx - 2] - a * x[Nx - 2]; }
    return a * x_int + b;
}

// linfit with x = 0..n-1 (tamcmc/sources/linfit.cpp:17-35, models.cpp:6065-6071) then eta0 (models.cpp:6073-6084)
inline double eta0_fct(const double* fl0, int n_)
{
    double sx = 0, sy = 0, sty = 0, stt = 0;
    const double n = (double)n_;
    for (int i = 0; i < n_; i++) sx += (double)i;
    for (int i = 0; i < n_; i++) sy += fl0[i];
    const double mean_x = sx / n;
    for (int i = 0; i < n_; i++) { const double t = (double)i - mean_x; sty += t * fl0[i]; }
    for (int i = 0; i < n_; i++) { const double t = (double)i - mean_x; stt += t * t; }
    const double Dnu_obs = sty / stt;
    const double G = 6.667e-8, Dnu_sun = 135.1, R_sun = 6.96342e5, M_sun = 1.98855e30;
    const double PI = 3.14159265358979323846;
    const double r5 = R_sun * 1e5;
    const double rho_sun = M_sun * 1e3 / (4 * PI * (r5 * r5 * r5) / 3);
    const double q = Dnu_obs / Dnu_sun;
    const double rho = (q * q) * rho_sun;
    return 3. * PI / (rho * G);
}

// analytical a-coefficients a1..a6 from the 2l+1 split frequencies nu[m+l] (tamcmc/sources/acoefs.cpp:112-256)
inline void eval_acoefs(int l, const double* nu, double aj[6])
{
    for (int k = 0; k < 6; k++) aj[k] = 0.0;
    if (l < 1 || l > 3) return;
    double T[3] = {0, 0, 0}, S[3] = {0, 0, 0};
    for (int k = 1; k <= l; k++) {
        T[k - 1] = (nu[l + k] - nu[l - k]) / (2 * k);
        S[k - 1] = (nu[l - k] + nu[l + k]) / 2 - nu[l];
    }
    if (l == 1) { aj[0] = T[0]; aj[1] = S[0] / 3; }
    else if (l == 2) {
        const long double num = T[0] + 4 * T[1];
        aj[0] = (double)(num / 5);
        aj[1] = (2 * S[1] - S[0]) / 7;
        aj[2] = (T[1] - T[0]) / 5;
        aj[3] = (S[1] - 4 * S[0]) / 70.;
    } else {
        aj[0] = T[0] / 14 + 2 * T[1] / 7 + 9 * T[2] / 14;
        aj[2] = -T[0] / 9 - 2 * T[1] / 9 + T[2] / 3;
        aj[4] = T[2] / 42 + 5 * T[0] / 126 - 4 * T[1] / 63;
        aj[1] = (-15 * S[0] + 25 * S[2]) / 126;
        aj[3] = 13 * (S[0] - 7 * S[1] + 3 * S[2]) / 1001;
        aj[5] = (15 * S[0] - 6 * S[1] + S[2]) / 1386;
    }
}

}  // namespace tamcmc_host
