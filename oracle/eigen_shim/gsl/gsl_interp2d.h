// Stand-in for the part of GSL the reference's Alm grid path uses (GSL is not available in the build container):
// gsl_interp2d_{alloc,init,eval_e,free} with gsl_interp2d_bicubic, gsl_interp_accel_{alloc,free}.
// TEST INFRASTRUCTURE ONLY: lets external/Alm/Alm_cpp/bilinear_interpol.cpp and Alm_interpol.cpp compile where they lie.
// It follows GSL 2.x's published algorithm (interp2d/bicubic.c over interpolation/cspline.c): node derivatives zx, zy, zxy
// from NATURAL cubic splines along rows / columns / rows of zy, then the 16-coefficient bicubic patch written out term by
// term as bicubic_eval does.  It is NOT GSL: the bicubic value of the ajAlm model stays "parity unpinned against GSL
// itself" (DESIGN.md 4); what this pins is everything the reference's own sources do around it (file parsing, grid choice,
// axis order, units), and the product's independent Hermite-basis implementation against this term-by-term one.
#ifndef ORACLE_GSL_STUB
#define ORACLE_GSL_STUB
#include <cmath>
#include <cstddef>
#include <vector>
struct gsl_interp2d_type { const char* name; };
static const gsl_interp2d_type gsl_interp2d_bicubic_obj = {"bicubic"};
static const gsl_interp2d_type gsl_interp2d_bilinear_obj = {"bilinear"};
static const gsl_interp2d_type* const gsl_interp2d_bicubic = &gsl_interp2d_bicubic_obj;
static const gsl_interp2d_type* const gsl_interp2d_bilinear = &gsl_interp2d_bilinear_obj;
struct gsl_interp2d {
    const gsl_interp2d_type* type;
    double xmin, xmax, ymin, ymax;
    size_t xsize, ysize;
    std::vector<double> zx, zy, zxy;
};
struct gsl_interp_accel { size_t cache; };
enum { GSL_SUCCESS = 0, GSL_EDOM = 1 };
inline gsl_interp_accel* gsl_interp_accel_alloc() { return new gsl_interp_accel{0}; }
inline void gsl_interp_accel_free(gsl_interp_accel* a) { delete a; }
inline gsl_interp2d* gsl_interp2d_alloc(const gsl_interp2d_type* T, size_t xsize, size_t ysize)
{
    gsl_interp2d* p = new gsl_interp2d();
    p->type = T; p->xsize = xsize; p->ysize = ysize;
    return p;
}
inline void gsl_interp2d_free(gsl_interp2d* p) { delete p; }
namespace gsl_shim_detail {
// cspline_init (natural boundary) + gsl_spline_eval_deriv at every node
inline void cspline_node_derivs(const double* xa, const double* ya, size_t n, double* out)
{
    std::vector<double> c(n, 0.0);
    const size_t sys = n - 2;
    if (sys == 1) c[1] = 3.0 * ((ya[2] - ya[1]) / (xa[2] - xa[1]) - (ya[1] - ya[0]) / (xa[1] - xa[0])) / (2.0 * (xa[2] - xa[0]));
    else if (sys > 1) {
        // Thomas algorithm on the tridiagonal system  h_i c_i + 2 (h_i + h_{i+1}) c_{i+1} + h_{i+1} c_{i+2} = 3 (dy_{i+1}/h_{i+1} - dy_i/h_i)
        std::vector<double> lo(sys), di(sys), up(sys), rhs(sys);
        for (size_t i = 0; i < sys; i++) {
            const double h0 = xa[i + 1] - xa[i], h1 = xa[i + 2] - xa[i + 1];
            lo[i] = h0; di[i] = 2.0 * (h0 + h1); up[i] = h1;
            rhs[i] = 3.0 * ((ya[i + 2] - ya[i + 1]) / h1 - (ya[i + 1] - ya[i]) / h0);
        }
        for (size_t i = 1; i < sys; i++) { const double w = lo[i] / di[i - 1]; di[i] -= w * up[i - 1]; rhs[i] -= w * rhs[i - 1]; }
        c[sys] = rhs[sys - 1] / di[sys - 1];
        for (size_t i = sys - 1; i-- > 0;) c[i + 1] = (rhs[i] - up[i] * c[i + 2]) / di[i];
    }
    for (size_t k = 0; k < n; k++) {
        const size_t i = (k < n - 1) ? k : n - 2;                    // interval that gsl_interp_accel_find returns for xa[k]
        const double dx = xa[i + 1] - xa[i], dy = ya[i + 1] - ya[i], delx = xa[k] - xa[i];
        const double b = dy / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0, cc = c[i], d = (c[i + 1] - c[i]) / (3.0 * dx);
        out[k] = b + delx * (2.0 * cc + 3.0 * d * delx);
    }
}
inline size_t bsearch(const double* a, double x, size_t lo, size_t hi)
{
    while (hi > lo + 1) { const size_t i = (hi + lo) / 2; if (a[i] > x) hi = i; else lo = i; }
    return lo;
}
}  // namespace gsl_shim_detail
#define GSL_SHIM_IDX(i, j) ((j) * xsize + (i))
inline int gsl_interp2d_init(gsl_interp2d* p, const double xa[], const double ya[], const double za[], size_t xsize, size_t ysize)
{
    p->xmin = xa[0]; p->xmax = xa[xsize - 1]; p->ymin = ya[0]; p->ymax = ya[ysize - 1];
    p->zx.assign(xsize * ysize, 0.0); p->zy.assign(xsize * ysize, 0.0); p->zxy.assign(xsize * ysize, 0.0);
    std::vector<double> v(xsize > ysize ? xsize : ysize), d(v.size());
    for (size_t j = 0; j < ysize; j++) {
        for (size_t i = 0; i < xsize; i++) v[i] = za[GSL_SHIM_IDX(i, j)];
        gsl_shim_detail::cspline_node_derivs(xa, v.data(), xsize, d.data());
        for (size_t i = 0; i < xsize; i++) p->zx[GSL_SHIM_IDX(i, j)] = d[i];
    }
    for (size_t i = 0; i < xsize; i++) {
        for (size_t j = 0; j < ysize; j++) v[j] = za[GSL_SHIM_IDX(i, j)];
        gsl_shim_detail::cspline_node_derivs(ya, v.data(), ysize, d.data());
        for (size_t j = 0; j < ysize; j++) p->zy[GSL_SHIM_IDX(i, j)] = d[j];
    }
    for (size_t j = 0; j < ysize; j++) {
        for (size_t i = 0; i < xsize; i++) v[i] = p->zy[GSL_SHIM_IDX(i, j)];
        gsl_shim_detail::cspline_node_derivs(xa, v.data(), xsize, d.data());
        for (size_t i = 0; i < xsize; i++) p->zxy[GSL_SHIM_IDX(i, j)] = d[i];
    }
    return GSL_SUCCESS;
}
inline int gsl_interp2d_eval_e(const gsl_interp2d* p, const double xarr[], const double yarr[], const double zarr[], double x, double y,
                               gsl_interp_accel*, gsl_interp_accel*, double* z)
{
    if (x < p->xmin || x > p->xmax || y < p->ymin || y > p->ymax) { *z = std::nan(""); return GSL_EDOM; }    // real GSL: error handler (abort)
    const size_t xsize = p->xsize, ysize = p->ysize;
    const size_t xi = gsl_shim_detail::bsearch(xarr, x, 0, xsize - 1), yi = gsl_shim_detail::bsearch(yarr, y, 0, ysize - 1);
    const double xmin = xarr[xi], xmax = xarr[xi + 1], ymin = yarr[yi], ymax = yarr[yi + 1];
    const double zminmin = zarr[GSL_SHIM_IDX(xi, yi)], zminmax = zarr[GSL_SHIM_IDX(xi, yi + 1)];
    const double zmaxmin = zarr[GSL_SHIM_IDX(xi + 1, yi)], zmaxmax = zarr[GSL_SHIM_IDX(xi + 1, yi + 1)];
    const double dx = xmax - xmin, dy = ymax - ymin;
    const double t = (x - xmin) / dx, u = (y - ymin) / dy, dt = 1. / dx, du = 1. / dy;
    const double zxminmin = p->zx[GSL_SHIM_IDX(xi, yi)] / dt, zxminmax = p->zx[GSL_SHIM_IDX(xi, yi + 1)] / dt;
    const double zxmaxmin = p->zx[GSL_SHIM_IDX(xi + 1, yi)] / dt, zxmaxmax = p->zx[GSL_SHIM_IDX(xi + 1, yi + 1)] / dt;
    const double zyminmin = p->zy[GSL_SHIM_IDX(xi, yi)] / du, zyminmax = p->zy[GSL_SHIM_IDX(xi, yi + 1)] / du;
    const double zymaxmin = p->zy[GSL_SHIM_IDX(xi + 1, yi)] / du, zymaxmax = p->zy[GSL_SHIM_IDX(xi + 1, yi + 1)] / du;
    const double zxyminmin = p->zxy[GSL_SHIM_IDX(xi, yi)] / (dt * du), zxyminmax = p->zxy[GSL_SHIM_IDX(xi, yi + 1)] / (dt * du);
    const double zxymaxmin = p->zxy[GSL_SHIM_IDX(xi + 1, yi)] / (dt * du), zxymaxmax = p->zxy[GSL_SHIM_IDX(xi + 1, yi + 1)] / (dt * du);
    const double t0 = 1, t1 = t, t2 = t * t, t3 = t * t2, u0 = 1, u1 = u, u2 = u * u, u3 = u * u2;
    double r = 0, v;
    v = zminmin; r += v * t0 * u0;
    v = zyminmin; r += v * t0 * u1;
    v = -3 * zminmin + 3 * zminmax - 2 * zyminmin - zyminmax; r += v * t0 * u2;
    v = 2 * zminmin - 2 * zminmax + zyminmin + zyminmax; r += v * t0 * u3;
    v = zxminmin; r += v * t1 * u0;
    v = zxyminmin; r += v * t1 * u1;
    v = -3 * zxminmin + 3 * zxminmax - 2 * zxyminmin - zxyminmax; r += v * t1 * u2;
    v = 2 * zxminmin - 2 * zxminmax + zxyminmin + zxyminmax; r += v * t1 * u3;
    v = -3 * zminmin + 3 * zmaxmin - 2 * zxminmin - zxmaxmin; r += v * t2 * u0;
    v = -3 * zyminmin + 3 * zymaxmin - 2 * zxyminmin - zxymaxmin; r += v * t2 * u1;
    v = 9 * zminmin - 9 * zmaxmin + 9 * zmaxmax - 9 * zminmax + 6 * zxminmin + 3 * zxmaxmin - 3 * zxmaxmax - 6 * zxminmax + 6 * zyminmin
        - 6 * zymaxmin - 3 * zymaxmax + 3 * zyminmax + 4 * zxyminmin + 2 * zxymaxmin + zxymaxmax + 2 * zxyminmax;
    r += v * t2 * u2;
    v = -6 * zminmin + 6 * zmaxmin - 6 * zmaxmax + 6 * zminmax - 4 * zxminmin - 2 * zxmaxmin + 2 * zxmaxmax + 4 * zxminmax - 3 * zyminmin
        + 3 * zymaxmin + 3 * zymaxmax - 3 * zyminmax - 2 * zxyminmin - zxymaxmin - zxymaxmax - 2 * zxyminmax;
    r += v * t2 * u3;
    v = 2 * zminmin - 2 * zmaxmin + zxminmin + zxmaxmin; r += v * t3 * u0;
    v = 2 * zyminmin - 2 * zymaxmin + zxyminmin + zxymaxmin; r += v * t3 * u1;
    v = -6 * zminmin + 6 * zmaxmin - 6 * zmaxmax + 6 * zminmax - 3 * zxminmin - 3 * zxmaxmin + 3 * zxmaxmax + 3 * zxminmax - 4 * zyminmin
        + 4 * zymaxmin + 2 * zymaxmax - 2 * zyminmax - 2 * zxyminmin - 2 * zxymaxmin - zxymaxmax - zxyminmax;
    r += v * t3 * u2;
    v = 4 * zminmin - 4 * zmaxmin + 4 * zmaxmax - 4 * zminmax + 2 * zxminmin + 2 * zxmaxmin - 2 * zxmaxmax - 2 * zxminmax + 2 * zyminmin
        - 2 * zymaxmin - 2 * zymaxmax + 2 * zyminmax + zxyminmin + zxymaxmin + zxymaxmax + zxyminmax;
    r += v * t3 * u3;
    *z = r;
    return GSL_SUCCESS;
}
#undef GSL_SHIM_IDX
#endif
