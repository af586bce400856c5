// alm_grid.cpp -- the Alm activity term from precomputed grids, as model_MS_Global_ajAlm_HarveyLike uses it
// (tamcmc/sources/models.cpp:1411-1746 through build_l_mode_ajAlm, build_lorentzian.cpp:189, and
// decompose_Alm_fct_GSLgrid, models.cpp:6110-6126).  Part of libtamcmc_gpu.so; plain host C++ (no CUDA calls).
//
// Reference pieces restated here (new code, same arithmetic / same file format):
//   loadGridData             external/Alm/Alm_cpp/bilinear_interpol.cpp:27-99   gzip'ed text: "x=..", "y=..", "z=", rows of z
//   loadAllData              external/Alm/Alm_cpp/Alm_interpol.cpp:11-79        files <dir>/<ftype>/A<l><m+l>.gz, m = 0..l
//   flatten_grid/init_2dgrid external/Alm/Alm_cpp/bilinear_interpol.cpp:101-124 z[j*nx + i], gsl_interp2d_bicubic
//   interpolate_core         external/Alm/Alm_cpp/bilinear_interpol.cpp:127-135 gsl_interp2d_eval_e
//   Alm_interp_iter_preinitialised  external/Alm/Alm_cpp/Alm_interpol.cpp:188-348  grid A<l><|m|>
//   Config::Config           tamcmc/sources/config.cpp:77-147                   "gate" and "triangle" grids, loaded once
//   make_Alm_grid / saveAlm / writeToFile / linspace_vec
//                            external/Alm/Alm_cpp/make_grids.cpp:17-131, gzip_compress.cpp:41-70, linspace.cpp:8-27 (GridMaker)
//
// THIRD-PARTY ALGORITHM.  The interpolation itself lives in GSL (un-vendored, version unpinned: find_package(GSL) in the
// reference's CMakeLists.txt:88).  What is restated is GSL 2.x's published algorithm for gsl_interp2d_bicubic
// (interp2d/bicubic.c) on top of gsl_interp_cspline (interpolation/cspline.c):
//   * init: zx = d/dx of the NATURAL cubic spline through each grid row, evaluated at the nodes; zy likewise along each
//     column; zxy = d/dx of the natural cubic spline through each row of zy;
//   * eval: the bicubic Hermite patch of the cell [x_i, x_i+1] x [y_j, y_j+1] that interpolates z, zx, zy, zxy at its four
//     corners (16 coefficients in t = (x - x_i)/dx, u = (y - y_j)/dy);
//   * gsl_interp2d_eval_e refuses points outside the grid (GSL_EDOM; the reference does not catch it: GSL's default
//     handler aborts the program).  Here: NaN.
#include "../../include/tamcmc_gpu.h"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <sys/stat.h>

namespace {

struct Grid {
    int nx = 0, ny = 0;
    std::vector<double> x, y, z, zx, zy, zxy;     // z[j*nx + i]: x = theta0 (columns of the file), y = delta (rows)
    bool ok = false;
};

// ---- the file: gzip'ed text (zlib reads plain text transparently as well) ----
bool read_all(const std::string& path, std::string& out)
{
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 15];
    int n;
    out.clear();
    while ((n = gzread(f, buf, sizeof(buf))) > 0) out.append(buf, (size_t)n);
    gzclose(f);
    return n == 0;
}

// stringstream(str) >> double of the reference's str_to_dbl (string_handler.cpp:457-463): a token that does not parse
// ("x=0", "y=0": the axis label is glued to the first value) yields 0
double token_to_dbl(const std::string& t)
{
    const char* s = t.c_str();
    char* e = nullptr;
    const double v = std::strtod(s, &e);
    return (e == s) ? 0.0 : v;
}

void split_ws(const std::string& line, std::vector<std::string>& out)
{
    out.clear();
    size_t i = 0;
    while (i < line.size()) {
        while (i < line.size() && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r')) i++;
        size_t j = i;
        while (j < line.size() && !(line[j] == ' ' || line[j] == '\t' || line[j] == '\r')) j++;
        if (j > i) out.push_back(line.substr(i, j - i));
        i = j;
    }
}

// loadGridData (bilinear_interpol.cpp:27-99): the line that starts with "x=" gives the columns, "y=" the rows, "z=" is a
// label, every other line is one row of z with n_cols values
bool parse_grid(const std::string& text, Grid& g)
{
    std::vector<std::string> tok;
    std::vector<std::vector<double>> rows;
    size_t p = 0;
    while (p < text.size()) {
        size_t q = text.find('\n', p);
        if (q == std::string::npos) q = text.size();
        const std::string line = text.substr(p, q - p);
        p = q + 1;
        split_ws(line, tok);
        if (tok.empty()) continue;
        const std::string key = tok[0].substr(0, tok[0].find('='));
        const bool labelled = tok[0].find('=') != std::string::npos;
        if (labelled && key == "x") { g.x.clear(); for (const auto& t : tok) g.x.push_back(token_to_dbl(t)); }
        else if (labelled && key == "y") { g.y.clear(); for (const auto& t : tok) g.y.push_back(token_to_dbl(t)); }
        else if (labelled && key == "z") continue;
        else {
            if (g.x.empty() || tok.size() < g.x.size()) return false;
            std::vector<double> r(g.x.size());
            for (size_t i = 0; i < g.x.size(); i++) r[i] = token_to_dbl(tok[i]);
            rows.push_back(r);
        }
    }
    g.nx = (int)g.x.size(); g.ny = (int)g.y.size();
    if (g.nx < 2 || g.ny < 2 || (int)rows.size() < g.ny) return false;
    g.z.assign((size_t)g.nx * g.ny, 0.0);
    for (int j = 0; j < g.ny; j++)
        for (int i = 0; i < g.nx; i++) g.z[(size_t)j * g.nx + i] = rows[j][i];       // flatten_grid
    for (int i = 1; i < g.nx; i++) if (!(g.x[i] > g.x[i - 1])) return false;         // GSL requires strictly increasing axes
    for (int j = 1; j < g.ny; j++) if (!(g.y[j] > g.y[j - 1])) return false;
    return true;
}

// Derivative at every node of the natural cubic spline through (xa, ya) -- gsl_interp_cspline's init (c[0] = c[n-1] = 0,
// symmetric tridiagonal system for the interior c) followed by gsl_spline_eval_deriv at the nodes.
void natural_spline_node_derivs(const double* xa, const double* ya, int n, double* d)
{
    std::vector<double> c((size_t)n, 0.0);
    const int N = n - 2;                                   // unknowns c[1..n-2]
    if (N == 1) {
        const double h0 = xa[1] - xa[0], h1 = xa[2] - xa[1];
        c[1] = 3.0 * ((ya[2] - ya[1]) / h1 - (ya[1] - ya[0]) / h0) / (2.0 * (h0 + h1));
    } else if (N > 1) {
        std::vector<double> diag((size_t)N), off((size_t)N), g((size_t)N), alpha((size_t)N), gamma((size_t)N), zz((size_t)N);
        for (int i = 0; i < N; i++) {
            const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
            const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
            off[i] = h_ip1;
            diag[i] = 2.0 * (h_ip1 + h_i);
            g[i] = 3.0 * ((ya[i + 2] - ya[i + 1]) * g_ip1 - (ya[i + 1] - ya[i]) * g_i);
        }
        // LDL^T of the symmetric tridiagonal matrix, forward and back substitution
        alpha[0] = diag[0]; gamma[0] = off[0] / alpha[0];
        for (int i = 1; i < N - 1; i++) { alpha[i] = diag[i] - off[i - 1] * gamma[i - 1]; gamma[i] = off[i] / alpha[i]; }
        alpha[N - 1] = diag[N - 1] - off[N - 2] * gamma[N - 2];
        zz[0] = g[0];
        for (int i = 1; i < N; i++) zz[i] = g[i] - gamma[i - 1] * zz[i - 1];
        for (int i = 0; i < N; i++) zz[i] /= alpha[i];
        c[N] = zz[N - 1];
        for (int i = N - 2; i >= 0; i--) c[i + 1] = zz[i] - gamma[i] * c[i + 2];
    }
    for (int i = 0; i < n - 1; i++) {
        const double dx = xa[i + 1] - xa[i], dy = ya[i + 1] - ya[i];
        d[i] = dy / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;                         // b_i: the slope at the left end of interval i
    }
    {
        const int i = n - 2;                                                         // the last node is the right end of the last interval
        const double dx = xa[i + 1] - xa[i], dy = ya[i + 1] - ya[i];
        const double b = dy / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0, dd = (c[i + 1] - c[i]) / (3.0 * dx);
        d[n - 1] = b + dx * (2.0 * c[i] + 3.0 * dd * dx);
    }
}

void bicubic_init(Grid& g)
{
    const int nx = g.nx, ny = g.ny;
    g.zx.assign((size_t)nx * ny, 0.0); g.zy.assign((size_t)nx * ny, 0.0); g.zxy.assign((size_t)nx * ny, 0.0);
    std::vector<double> col((size_t)ny), dcol((size_t)ny);
    for (int j = 0; j < ny; j++) natural_spline_node_derivs(g.x.data(), &g.z[(size_t)j * nx], nx, &g.zx[(size_t)j * nx]);
    for (int i = 0; i < nx; i++) {
        for (int j = 0; j < ny; j++) col[j] = g.z[(size_t)j * nx + i];
        natural_spline_node_derivs(g.y.data(), col.data(), ny, dcol.data());
        for (int j = 0; j < ny; j++) g.zy[(size_t)j * nx + i] = dcol[j];
    }
    for (int j = 0; j < ny; j++) natural_spline_node_derivs(g.x.data(), &g.zy[(size_t)j * nx], nx, &g.zxy[(size_t)j * nx]);
}

// index i with a[i] <= v < a[i+1]; the last node belongs to the last interval (gsl_interp_bsearch over [0, n-1])
int find_cell(const std::vector<double>& a, double v)
{
    int lo = 0, hi = (int)a.size() - 1;
    while (hi > lo + 1) { const int mid = (hi + lo) / 2; if (a[mid] > v) hi = mid; else lo = mid; }
    return lo;
}

double bicubic_eval(const Grid& g, double x, double y)
{
    if (!(x >= g.x.front() && x <= g.x.back() && y >= g.y.front() && y <= g.y.back())) return std::nan("");      // GSL_EDOM
    const int xi = find_cell(g.x, x), yi = find_cell(g.y, y), nx = g.nx;
    const double dx = g.x[xi + 1] - g.x[xi], dy = g.y[yi + 1] - g.y[yi];
    const double t = (x - g.x[xi]) / dx, u = (y - g.y[yi]) / dy;
    // cubic Hermite basis: value at 0, value at 1, slope at 0, slope at 1 (slopes in cell units)
    const double t2 = t * t, t3 = t2 * t, u2 = u * u, u3 = u2 * u;
    const double ht[4] = {2 * t3 - 3 * t2 + 1, -2 * t3 + 3 * t2, t3 - 2 * t2 + t, t3 - t2};
    const double hu[4] = {2 * u3 - 3 * u2 + 1, -2 * u3 + 3 * u2, u3 - 2 * u2 + u, u3 - u2};
    double r = 0.0;
    for (int b = 0; b < 2; b++)
        for (int a = 0; a < 2; a++) {
            const size_t k = (size_t)(yi + b) * nx + (xi + a);
            r += g.z[k] * ht[a] * hu[b] + g.zx[k] * dx * ht[2 + a] * hu[b] + g.zy[k] * dy * ht[a] * hu[2 + b]
               + g.zxy[k] * dx * dy * ht[2 + a] * hu[2 + b];
        }
    return r;
}

bool is_dir(const std::string& p)
{
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

const char* ftype_name(int filter_code) { return filter_code == 0 ? "gate" : filter_code == 2 ? "triangle" : nullptr; }

thread_local std::string g_grid_error;

}  // namespace

// [filter slot 0 = gate, 1 = triangle][l - 1][|m|]
struct tamcmc_alm_grids { Grid g[2][3][4]; };

extern "C" {

const char* tamcmc_alm_grids_last_error(void) { return g_grid_error.c_str(); }

int tamcmc_alm_grids_load(const char* grid_dir, tamcmc_alm_grids** out)
{
    if (!grid_dir || !out) return TAMCMC_ERR_ARG;
    *out = nullptr;
    tamcmc_alm_grids* G = new tamcmc_alm_grids();
    for (int f = 0; f < 2; f++) {
        const std::string dir = std::string(grid_dir) + "/" + (f == 0 ? "gate" : "triangle");
        if (!is_dir(dir)) { g_grid_error = "Invalid grid directory or file type: " + dir; delete G; return TAMCMC_ERR_ARG; }
        for (int l = 1; l <= 3; l++)
            for (int m = 0; m <= l; m++) {
                // only m >= 0 is read: file A<l><m+l>.gz (Alm_interpol.cpp:20-33)
                const std::string file = dir + "/A" + std::to_string(l) + std::to_string(m + l) + ".gz";
                std::string text;
                Grid& g = G->g[f][l - 1][m];
                if (!read_all(file, text)) { g_grid_error = "Grid file does not exist or cannot be read: " + file; delete G; return TAMCMC_ERR_ARG; }
                if (!parse_grid(text, g)) { g_grid_error = "Malformed grid file: " + file; delete G; return TAMCMC_ERR_ARG; }
                bicubic_init(g);
                g.ok = true;
            }
    }
    *out = G;
    return TAMCMC_OK;
}

void tamcmc_alm_grids_free(tamcmc_alm_grids* G) { delete G; }

// tamcmc_alm_fn: Alm_interp_iter_preinitialised (Alm_interpol.cpp:188-348).  theta0, delta in radians.
double tamcmc_alm_grids_eval(int l, int m, double theta0, double delta, int filter_code, void* grids)
{
    const tamcmc_alm_grids* G = static_cast<const tamcmc_alm_grids*>(grids);
    if (l <= 0 || l > 3) return -9998;                                               // Alm_interpol.cpp:194-202
    const int am = m < 0 ? -m : m;
    if (am > l) return -9998;                                                        // "Invalid lm combination"
    if (!G || (filter_code != 0 && filter_code != 2)) return std::nan("");
    const Grid& g = G->g[filter_code == 0 ? 0 : 1][l - 1][am];
    if (!g.ok) return std::nan("");
    return bicubic_eval(g, theta0, delta);
}

int tamcmc_alm_grids_shape(const tamcmc_alm_grids* G, int filter_code, int l, int m, int* nx, int* ny)
{
    if (!G || l < 1 || l > 3 || m < -l || m > l || (filter_code != 0 && filter_code != 2)) return TAMCMC_ERR_ARG;
    const Grid& g = G->g[filter_code == 0 ? 0 : 1][l - 1][m < 0 ? -m : m];
    if (nx) *nx = g.nx;
    if (ny) *ny = g.ny;
    return TAMCMC_OK;
}

int tamcmc_alm_grids_nodes(const tamcmc_alm_grids* G, int filter_code, int l, int m, double* x, double* y, double* z)
{
    if (!G || l < 1 || l > 3 || m < -l || m > l || (filter_code != 0 && filter_code != 2)) return TAMCMC_ERR_ARG;
    const Grid& g = G->g[filter_code == 0 ? 0 : 1][l - 1][m < 0 ? -m : m];
    if (x) std::memcpy(x, g.x.data(), sizeof(double) * g.x.size());
    if (y) std::memcpy(y, g.y.data(), sizeof(double) * g.y.size());
    if (z) std::memcpy(z, g.z.data(), sizeof(double) * g.z.size());
    return TAMCMC_OK;
}

// GridMaker (do_grids.cpp:63-76, make_grids.cpp:38-131): the 2l+1 grids A<l><m+l>.gz of every l <= lmax for one filter, in
// <out_dir>/<ftype>/, on theta0 in [theta_min, theta_max] x delta in [delta_min, delta_max] with ceil(range / resol) nodes
// per axis (linspace_vec keeps both end points), Alm = 0 where delta < 0.001.  Values are printed like the reference's
// ofstream << double does (6 significant digits) and the text is gzip'ed.
int tamcmc_alm_grids_make(const char* out_dir, int filter_code, int lmax, double resol, double theta_min, double theta_max,
                          double delta_min, double delta_max)
{
    const char* ftype = ftype_name(filter_code);
    if (!out_dir || !ftype || lmax < 1 || lmax > 3 || !(resol > 0)) return TAMCMC_ERR_ARG;
    if (theta_min < 0 || theta_max > M_PI || theta_max <= theta_min) return TAMCMC_ERR_ARG;      // make_grids.cpp:43-48
    if (delta_min < 0 || delta_max > M_PI || delta_max <= delta_min) return TAMCMC_ERR_ARG;      // make_grids.cpp:49-54
    const int Ntheta = (int)std::ceil((theta_max - theta_min) / resol), Ndelta = (int)std::ceil((delta_max - delta_min) / resol);
    if (Ntheta < 2 || Ndelta < 2) return TAMCMC_ERR_ARG;
    auto linspace = [](double a, double b, int n) {                                              // linspace.cpp:8-27
        std::vector<double> v;
        const double d = (double)(((long double)b - (long double)a) / (n - 1));
        for (int i = 0; i < n - 1; i++) v.push_back((double)((long double)a + (long double)d * i));
        v.push_back(b);
        return v;
    };
    const std::vector<double> theta = linspace(theta_min, theta_max, Ntheta), delta = linspace(delta_min, delta_max, Ndelta);
    mkdir(out_dir, 0700);
    const std::string dir = std::string(out_dir) + "/" + ftype;
    mkdir(dir.c_str(), 0700);
    if (!is_dir(dir)) { g_grid_error = "Could not create output directory " + dir; return TAMCMC_ERR_ARG; }
    const double delta_limit = 0.001;
    char num[64];
    for (int l = 1; l <= lmax; l++)
        for (int m = -l; m <= l; m++) {
            std::string text = "x=";
            for (double v : theta) { std::snprintf(num, sizeof(num), "%g ", v); text += num; }
            text += "\ny=";
            for (double v : delta) { std::snprintf(num, sizeof(num), "%g ", v); text += num; }
            text += "\nz=\n";
            for (int j = 0; j < Ndelta; j++) {
                for (int i = 0; i < Ntheta; i++) {
                    const double r = (delta[j] >= delta_limit) ? tamcmc_host_alm(l, m, theta[i], delta[j], filter_code) : 0.0;
                    std::snprintf(num, sizeof(num), "%g ", r);
                    text += num;
                }
                text += "\n";
            }
            const std::string file = dir + "/A" + std::to_string(l) + std::to_string(m + l) + ".gz";
            gzFile f = gzopen(file.c_str(), "wb");
            if (!f) { g_grid_error = "Could not open output file " + file; return TAMCMC_ERR_ARG; }
            const int w = gzwrite(f, text.data(), (unsigned)text.size());
            gzclose(f);
            if (w != (int)text.size()) { g_grid_error = "Gzip compression error: " + file; return TAMCMC_ERR_ARG; }
        }
    return TAMCMC_OK;
}

}  // extern "C"
