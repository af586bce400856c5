// rgb_device.cu -- device half of the red-giant expander (BASELINE configs C1 / C4): tamcmc_gpu_rgb_expand.
//
// Replaces, for ALL chains of a step in one call, the two loops that are 99 % of model_RGB_asympt_aj_*Width_HarveyLike_v4's
// set-up (tamcmc/sources/models.cpp:4684-4927, 4334-4556) and that the reference runs per chain under OpenMP:
//   * the (p mode, g mode) pair loop of the asymptotic mixed-mode solver (external/ARMM/solver_mm.cpp:558-573 -> solver_mm :326-449):
//     tamcmc_rgb_search_kernel (one warp per segment of a p mode's band: where the sign changes are) and tamcmc_rgb_pairs_kernel (every sign
//     change evaluated exactly for every g mode, minimum per sign change) -- rgb_solver.cuh;
//   * the normalisation of the zeta function: the maximum of the zeta sums over a 4-year-resolution grid
//     (external/ARMM/bump_DP.cpp:126-163): tamcmc_rgb_ksi_max_kernel + tamcmc_rgb_ksi_top_kernel locate it (one thread per grid frequency),
//     the host evaluates it there like the reference.
// The host does what is left (host_rgb.cpp: prepare -> [device] -> finish): unpacking, l=0 widths, the p / g mode lists and the long
// double grid set-up in front; filter / sort / unique of the solutions, bias spline, zeta at the ~50 mixed modes, heights, widths,
// splittings and the mode-table rows behind.  Rows are written where the caller says -- normally the pinned staging block of the
// evaluation context (tamcmc_gpu_params_staging), so tamcmc_gpu_eval reads them without another copy.
// A chain the device flags (rgb_solver.cuh: an exact zero, an unexpected shape, ...) is solved by the host code of the same library
// (tamcmc_host_expand_rgb_v4's own path), reported in path_out.  Compiled with -fmad=false: the solver's arithmetic must not contract.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/tamcmc_gpu.h"
#include "rgb_solver.cuh"

using namespace tamcmc_rgb;

namespace {

constexpr unsigned long long SLOT_EMPTY = ~0ull;
constexpr int TOP_CAP = 14;
struct OutHdr { unsigned long long norm_bits; int count, flag, ntop, top[TOP_CAP], pad_; };     // per chain, 80 bytes
constexpr int REC_CAP = 128;         // sign-change records per band (two or three per segment)


// ---- phase 1 (rgb_solver.cuh: pair_segment + local_search), one WARP per segment.  The scalar code bisects; a warp probes 32 points of
// the bracket at once (f rises on it: the negative probes are a prefix), so a 1000-point stretch takes two rounds instead of ten dependent
// evaluations.  Same decisions, same records (the host emulation runs the scalar code; the rows must come out identical).
#define RGB_FULL 0xffffffffu

// the l with f(l) < 0 < f(l+1) on a stretch where f rises from f(l0) < 0 to f(h0) > 0; -1 with a flag when a probe is zero, not finite,
// or the signs are not a prefix
template <class F>
__device__ int warp_ksection(F&& S, int l, int h, int lane, int& flag)
{
    while (h - l > 1) {
        const int span = h - l - 1;
        const bool all = span > 32;
        const int probe = all ? l + (int)(((long)(h - l) * (lane + 1)) / 33) : l + 1 + lane;
        const bool valid = all || lane < span;
        const double v = valid ? S(probe) : 1.0;
        const bool bad = valid && (!(v == v) || v == 0.0 || fabs(v) > 1.0e300);
        if (__any_sync(RGB_FULL, bad)) { flag |= RGB_FLAG_ZERO; return -1; }
        const unsigned neg = __ballot_sync(RGB_FULL, valid && v < 0.0);
        const int nneg = __popc(neg), nvalid = all ? 32 : span;
        if (neg != ((nneg == 32) ? 0xffffffffu : ((1u << nneg) - 1u))) { flag |= RGB_FLAG_SHAPE; return -1; }
        const int below = __shfl_sync(RGB_FULL, probe, (nneg > 0) ? nneg - 1 : 0);
        const int above = __shfl_sync(RGB_FULL, probe, (nneg < 32) ? nneg : 31);
        if (nneg > 0) l = below;
        if (nneg < nvalid) h = above;
    }
    return l;
}

// values of f at up to 7 points, one lane each, returned to every lane; false with a flag when one is zero or not finite
template <class F>
__device__ bool warp_values(F&& S, const int* pts, int np, double* fv, int lane, int& flag)
{
    const double mine = (lane < np) ? S(pts[lane]) : 1.0;
    const bool bad = lane < np && (!(mine == mine) || mine == 0.0 || fabs(mine) > 1.0e300);
    if (__any_sync(RGB_FULL, bad)) { flag |= (mine == 0.0) ? RGB_FLAG_ZERO : RGB_FLAG_NONFINITE; return false; }
#pragma unroll
    for (int k = 0; k < 7; k++) fv[k] = __shfl_sync(RGB_FULL, mine, k);
    return true;
}

__device__ void warp_local(const Band& B, double inv_g, int idx, Record* out, int* nrec, int lane, int& flag)
{
    Record R;
    R.idx = idx; R.kind = 0; R.i = 0;
    if (!local_grid_ext(band_nu(B, idx), B.resol2, B.Dh, B.Dl, R.lo, R.hi, R.n)) { flag |= RGB_FLAG_EXT; return; }
    const int Nx = R.n;
    if (Nx < 2) return;
    const double lo = R.lo, hi = R.hi, lstep = (hi - lo) / (double)(Nx - 1);
    auto y = [&](int i) { return (i == Nx - 1) ? hi : lo + (double)i * lstep; };
    auto S = [&](int i) { return pmg_sign<TrigLib, TrigCR>(B, inv_g, y(i), flag); };
    const double uA = u_of(B, inv_g, y(0)), uB = u_of(B, inv_g, y(Nx - 1));
    const double mA = floor(uA - 0.5), mB = ceil(uB - 0.5);
    const bool manypoles = !(mA - mB < 1.0);
    int pts[7], np = 0;
    pts[np++] = 0;
    if (!manypoles && mA >= mB) {
        const double nup = nu_of_u(B, inv_g, mA + 0.5);
        const int il = (int)floor((nup - lo) / lstep);
        for (int k = il - 1; k <= il + 2; k++) if (k > pts[np - 1] && k <= Nx - 1) pts[np++] = k;
    }
    if (Nx - 1 > pts[np - 1]) pts[np++] = Nx - 1;
    double fv[7];
    if (!warp_values(S, pts, np, fv, lane, flag)) return;
    const double X0 = fv[0], XN = fv[np - 1];
    bool emit = false;
    if (0.0 > XN) { R.kind = 2; R.i = Nx - 2; emit = true; }                 // the last assignment of lin_interpol wins
    else if (0.0 < X0) { R.kind = 1; R.i = 0; emit = true; }
    else {
        if (manypoles) { flag |= RGB_FLAG_POLES; return; }
        for (int s = 0; s + 1 < np && !emit; s++) {
            const int pa = pts[s], pb = pts[s + 1];
            const double fa = fv[s], fb = fv[s + 1];
            if (pb == pa + 1) { if (fa < 0.0 && fb > 0.0) { R.i = pa; emit = true; } }
            else if (fa < 0.0 && fb > 0.0) {
                const int l = warp_ksection(S, pa, pb, lane, flag);
                if (l < 0) return;
                R.i = l; emit = true;
            } else if (fa > 0.0 && fb < 0.0) { flag |= RGB_FLAG_SHAPE; return; }
        }
        if (!emit) { flag |= RGB_FLAG_SHAPE; return; }
        if (R.i > Nx - 2) R.i = Nx - 2;
        R.kind = 0;
    }
    if (lane == 0) {
        const int k = atomicAdd(nrec, 1);
        if (k < REC_CAP) out[k] = R; else flag |= RGB_FLAG_OVERFLOW;
    }
}

__global__ void __launch_bounds__(512) tamcmc_rgb_search_kernel(const Band* __restrict__ bands, Record* __restrict__ recs, int* __restrict__ nrec,
                                                                OutHdr* __restrict__ hdr)
{
    const Band B = bands[blockIdx.x];
    if (B.nband == 0 || B.rep_inv_g == 0.0) return;
    const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5), nwarps = (int)(blockDim.x >> 5);
    const double inv_g = B.rep_inv_g;
    int flag = 0;
    double m_hi = 0, nu0 = 0, bstep = 0;
    const int nseg = pair_segments(B, inv_g, m_hi, nu0, bstep, flag);
    Record* out = recs + (size_t)blockIdx.x * REC_CAP;
    const int nb = B.nband, npoles = nseg - 1;
    auto ip_of = [&](int k) { return (int)floor((nu_of_u(B, inv_g, (m_hi - (double)k) + 0.5) - nu0) / bstep); };
    auto S = [&](int i) { return pmg_sign<TrigLib, TrigCR>(B, inv_g, band_nu(B, i), flag); };
    for (int j = (int)blockIdx.y * nwarps + warp; j < nseg; j += nwarps * (int)gridDim.y) {       // (two CTAs per band: ~one segment per warp)
        int pts[7], np = 0, start = 0;
        if (j > 0) { start = ip_of(j - 1) + 2; if (start < 0) start = 0; if (start > nb - 1) start = nb - 1; }
        pts[np++] = start;
        if (j < npoles) {
            const int ip = ip_of(j);
            for (int k = ip - 1; k <= ip + 2; k++) if (k > pts[np - 1] && k <= nb - 1) pts[np++] = k;
        } else if (nb - 1 > pts[np - 1]) pts[np++] = nb - 1;
        double fv[7];
        if (!warp_values(S, pts, np, fv, lane, flag)) break;
        bool stop = false;
        for (int s = 0; s + 1 < np && !stop; s++) {
            const int pa = pts[s], pb = pts[s + 1];
            const double fa = fv[s], fb = fv[s + 1];
            int idx = -1;
            if (pb == pa + 1) {
                if ((fb > 0.0 && fa < 0.0) || (fb < 0.0 && fa > 0.0)) idx = pa;            // the two tests of sign_change; no value is zero here
            } else {
                if (fa > 0.0 && fb < 0.0) { flag |= RGB_FLAG_SHAPE; stop = true; }
                else if (fa < 0.0 && fb > 0.0) { idx = warp_ksection(S, pa, pb, lane, flag); if (idx < 0) stop = true; }
            }
            if (idx >= 0) warp_local(B, inv_g, idx, out, nrec + blockIdx.x, lane, flag);
        }
        if (stop) break;
    }
    if (flag) atomicOr(&hdr[B.chain].flag, flag);
}

// phase 2: every record of a band for every g mode of the band, `lanes` threads per (p mode, g mode) pair.  The solution of record r goes
// to slot r of the band with atomicMin on its bit pattern (frequencies are positive: the order of the bits is the order of the values): the
// reference keeps the smallest of every cluster of solutions (sort + unique with a tolerance of two bins, solver_mm.cpp:575-590) -- the
// minimum per record (= per sign change of the coarse grid) is that value whenever a cluster does not straddle two coarse grid cells, and
// the host's sort + unique merges the slots when it does.  [bands][REC_CAP] slots and the record counts go back to the host as they are.
__global__ void __launch_bounds__(128) tamcmc_rgb_pairs_kernel(const Band* __restrict__ bands, const Pair* __restrict__ pairs, int npairs, int lanes,
                                                                const Record* __restrict__ recs, const int* __restrict__ nrec,
                                                                unsigned long long* __restrict__ slots, OutHdr* __restrict__ hdr)
{
    const int t = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
    const int pair = t / lanes, lane = t % lanes;
    if (pair >= npairs) return;
    const Pair Q = pairs[pair];
    int n = nrec[Q.band];
    if (n > REC_CAP) n = REC_CAP;
    if (lane >= n) return;
    const Band B = bands[Q.band];
    int flag = 0;
    for (int r = lane; r < n; r += lanes) {
        const Record R = recs[(size_t)Q.band * REC_CAP + r];
        double sol;
        if (record_eval<TrigCR>(B, Q.inv_g, Q.nu_g, R, sol, flag)) {
            if (sol > 0.0) atomicMin(slots + (size_t)Q.band * REC_CAP + r, (unsigned long long)__double_as_longlong(sol));
            else flag |= RGB_FLAG_NONFINITE;
        }
    }
    if (flag) atomicOr(&hdr[B.chain].flag, flag);
}

__device__ __forceinline__ double fast_rcp(double d)          // 1 / d to ~1 ulp: hardware seed + two Newton steps
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = __fma_rn(-d, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-d, r, 1.0);
    return __fma_rn(r, e, r);
}

// WHERE the zeta sums (bump_DP.cpp:126-163) are largest on the 4-year-resolution grid.  One term of the sum over (np, ng) is
//   ksi_fct1 (bump_DP.cpp:46-64) = 1 / (1 + (nd / qd) (cu2 / cd2)) = A / (A + B),  A = qd cd2 (per p mode), B = nd cu2 (per g mode),
//   cu2 = cos^2(pi 1e6 (1/nu - 1/nu_g) / DPl),  1/nu_g = (ng + alpha) DPl 1e-6:
// cos^2 has period pi, so cu2 -- and with the common DPl, B -- is the SAME for every g mode up to rounding: the reference adds Lg copies of
// each p mode's term.  The kernel evaluates Lg * sum_p A_p / (A_p + B_0) (Lp + 1 cosines and Lp reciprocals per grid point instead of
// Lp + Lg and 3 Lp Lg): good to ~1e-13 relative, which is enough to say where the maximum is; the value that normalises the zeta function
// is then summed by the host, term by term like the reference, at the grid points the next kernel lists.
__global__ void __launch_bounds__(128) tamcmc_rgb_ksi_max_kernel(const KsiHdr* __restrict__ hdrs, const double* __restrict__ kp,
                                                                  const double* __restrict__ kg, double* __restrict__ vals, OutHdr* __restrict__ hdr)
{
    const KsiHdr H = hdrs[blockIdx.y];
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if ((int)(blockIdx.x * blockDim.x) >= H.Ndata) return;
    double total = 0.0;
    if (i < H.Ndata) {
        // Eigen::VectorXd::LinSpaced(Ndata, fmin, fmax)[i]
        const double v = (H.Ndata == 1 || i == H.Ndata - 1) ? H.fmax : H.fmin + (double)i * ((H.fmax - H.fmin) / (double)(H.Ndata - 1));
        const double inv = 1.0 / v, sq = 1e-6 * (v * v);
        const double inv_g0 = kg[2 * (size_t)H.off_g], DPl0 = kg[2 * (size_t)H.off_g + 1];
        const double cu = cos((H.c_up * (inv - inv_g0)) / DPl0);
        const double B = (sq * DPl0) * (cu * cu);
        for (int p = 0; p < H.Lp; p++) {
            const double* P = kp + 3 * ((size_t)H.off_p + p);
            const double c = cos((H.pi_d * (v - P[0])) / P[1]);
            const double A = P[2] * (c * c);
            total += A * fast_rcp(A + B);
        }
        total *= (double)H.Lg;
        vals[H.val_off + i] = total;
    }
    // the maximum does not depend on the order it is taken in; a NaN anywhere must survive (its bit pattern is above every finite value's)
    unsigned long long b = (unsigned long long)__double_as_longlong(total);
    if (total != total) b = 0x7ff8000000000000ull;
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o); b = (t > b) ? t : b; }
    if ((threadIdx.x & 31) == 0) atomicMax(&hdr[H.chain].norm_bits, b);
}

// the grid points whose (fast) sum is within 1e-9 of the maximum: the host evaluates those exactly and takes the maximum of that
__global__ void __launch_bounds__(128) tamcmc_rgb_ksi_top_kernel(const KsiHdr* __restrict__ hdrs, const double* __restrict__ vals, OutHdr* __restrict__ hdr)
{
    const KsiHdr H = hdrs[blockIdx.y];
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= H.Ndata) return;
    const double mx = __longlong_as_double((long long)hdr[H.chain].norm_bits);
    const double v = vals[H.val_off + i];
    if (v >= mx * (1.0 - 1e-9)) {
        const int k = atomicAdd(&hdr[H.chain].ntop, 1);
        if (k < TOP_CAP) hdr[H.chain].top[k] = i;
    }
}

std::string g_err;

#define RGB_CUDA(x)                                                                                                   \
    do {                                                                                                              \
        cudaError_t e_ = (x);                                                                                         \
        if (e_ != cudaSuccess) { g_err = std::string(#x) + ": " + cudaGetErrorString(e_); return TAMCMC_ERR_CUDA; }   \
    } while (0)

size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

}  // namespace

// One group of chains in flight: its own buffers and streams, so that the host can finish one group while the device solves the other
struct RgbGroup {
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev_in = nullptr, ev_ksi = nullptr;
    char* h_in = nullptr; char* d_in = nullptr; size_t in_cap = 0;
    char* h_out = nullptr; char* d_out = nullptr;          // [max_chains] OutHdr
    char* h_res = nullptr; char* d_res = nullptr; size_t res_cap = 0;     // [bands] record counts, then [bands][REC_CAP] solution slots
    std::vector<int> band_lo, band_hi;                     // per chain: its bands in `task`
    double* d_vals = nullptr; size_t vals_cap = 0;        // zeta sums over the normalisation grids
    char* d_recs = nullptr; size_t recs_cap = 0;          // [bands][REC_CAP] records
    DeviceTask task;
    int c0 = 0, c1 = 0;                                    // chains [c0, c1) of the call
    bool launched = false;
};

struct tamcmc_gpu_rgb {
    int device = 0, max_chains = 0;
    size_t out_bytes = 0;
    RgbGroup grp[2];
    std::vector<Prep*> preps;
    std::vector<DeviceTask> chain_task;    // one per chain, filled by the chain's own thread, then concatenated per group
    std::vector<int> on_device;
    double last_ms[4] = {0, 0, 0, 0};      // prepare, device (enqueue + first wait), finish, total of the last call (host clock)
    long chains_total = 0, chains_host = 0; // chain set-ups asked for / handed to the host solver since create
};

namespace {

// chains [G.c0, G.c1): concatenate their tasks, copy them in, launch the kernels and the copy back; nothing waits here
int rgb_enqueue(tamcmc_gpu_rgb* h, RgbGroup& G)
{
    DeviceTask& T = G.task;
    T.clear();
    G.launched = false;
    G.band_lo.assign((size_t)h->max_chains, 0); G.band_hi.assign((size_t)h->max_chains, 0);
    for (int c = G.c0; c < G.c1; c++) {
        if (!h->on_device[(size_t)c]) continue;
        const DeviceTask& Tc = h->chain_task[(size_t)c];
        const int band0 = (int)T.bands.size(), p0 = (int)(T.kp.size() / 3), g0 = (int)(T.kg.size() / 2);
        for (Band B : Tc.bands) { B.slot_off += T.nslots; T.bands.push_back(B); }
        for (Pair Q : Tc.pairs) { Q.band += band0; T.pairs.push_back(Q); }
        for (KsiHdr K : Tc.ksi) { K.off_p += p0; K.off_g += g0; K.val_off += T.nvals; T.ksi.push_back(K); }
        T.kp.insert(T.kp.end(), Tc.kp.begin(), Tc.kp.end());
        T.kg.insert(T.kg.end(), Tc.kg.begin(), Tc.kg.end());
        T.nslots += Tc.nslots; T.nvals += Tc.nvals;
        G.band_lo[(size_t)c] = band0; G.band_hi[(size_t)c] = (int)T.bands.size();
    }
    if (T.ksi.empty()) return TAMCMC_OK;
    const size_t hdr_bytes = align16((size_t)h->max_chains * sizeof(OutHdr));
    size_t off[6];
    off[0] = 0;
    off[1] = off[0] + align16(T.bands.size() * sizeof(Band));
    off[2] = off[1] + align16(T.pairs.size() * sizeof(Pair));
    off[3] = off[2] + align16(T.ksi.size() * sizeof(KsiHdr));
    off[4] = off[3] + align16(T.kp.size() * 8);
    off[5] = off[4] + align16(T.kg.size() * 8);
    if (off[5] > G.in_cap) {
        if (G.d_in) cudaFree(G.d_in);
        if (G.h_in) cudaFreeHost(G.h_in);
        G.d_in = nullptr; G.h_in = nullptr;
        G.in_cap = off[5] + off[5] / 2;
        RGB_CUDA(cudaMalloc((void**)&G.d_in, G.in_cap));
        RGB_CUDA(cudaMallocHost((void**)&G.h_in, G.in_cap));
    }
    const int nbands = (int)T.bands.size();
    const size_t cnt_bytes = align16((size_t)nbands * 4), res_bytes = cnt_bytes + (size_t)nbands * REC_CAP * 8;
    if (res_bytes > G.res_cap) {
        if (G.d_res) cudaFree(G.d_res);
        if (G.h_res) cudaFreeHost(G.h_res);
        G.d_res = nullptr; G.h_res = nullptr;
        G.res_cap = res_bytes + res_bytes / 2;
        RGB_CUDA(cudaMalloc((void**)&G.d_res, G.res_cap));
        RGB_CUDA(cudaMallocHost((void**)&G.h_res, G.res_cap));
    }
    if ((size_t)T.nvals > G.vals_cap) {
        if (G.d_vals) cudaFree(G.d_vals);
        G.d_vals = nullptr;
        G.vals_cap = (size_t)T.nvals + (size_t)T.nvals / 2 + 1024;
        RGB_CUDA(cudaMalloc((void**)&G.d_vals, G.vals_cap * 8));
    }
    std::memcpy(G.h_in + off[0], T.bands.data(), T.bands.size() * sizeof(Band));
    std::memcpy(G.h_in + off[1], T.pairs.data(), T.pairs.size() * sizeof(Pair));
    std::memcpy(G.h_in + off[2], T.ksi.data(), T.ksi.size() * sizeof(KsiHdr));
    std::memcpy(G.h_in + off[3], T.kp.data(), T.kp.size() * 8);
    std::memcpy(G.h_in + off[4], T.kg.data(), T.kg.size() * 8);
    RGB_CUDA(cudaMemcpyAsync(G.d_in, G.h_in, off[5], cudaMemcpyHostToDevice, G.stream));
    RGB_CUDA(cudaMemsetAsync(G.d_out, 0, hdr_bytes, G.stream));
    RGB_CUDA(cudaEventRecord(G.ev_in, G.stream));
    OutHdr* d_hdr = (OutHdr*)G.d_out;
    {   // the zeta normalisation on its own stream: it does not depend on the pair loop
        int maxN = 0;
        for (const KsiHdr& K : T.ksi) if (K.Ndata > maxN) maxN = K.Ndata;
        RGB_CUDA(cudaStreamWaitEvent(G.stream2, G.ev_in, 0));
        const dim3 grid((unsigned)((maxN + 127) / 128), (unsigned)T.ksi.size());
        tamcmc_rgb_ksi_max_kernel<<<grid, 128, 0, G.stream2>>>((const KsiHdr*)(G.d_in + off[2]), (const double*)(G.d_in + off[3]),
                                                                (const double*)(G.d_in + off[4]), G.d_vals, d_hdr);
        tamcmc_rgb_ksi_top_kernel<<<grid, 128, 0, G.stream2>>>((const KsiHdr*)(G.d_in + off[2]), G.d_vals, d_hdr);
        RGB_CUDA(cudaEventRecord(G.ev_ksi, G.stream2));
    }
    if (res_bytes) {
        RGB_CUDA(cudaMemsetAsync(G.d_res, 0, cnt_bytes, G.stream));
        RGB_CUDA(cudaMemsetAsync(G.d_res + cnt_bytes, 0xff, res_bytes - cnt_bytes, G.stream));
    }
    if (!T.pairs.empty()) {
        const int npairs = (int)T.pairs.size();
        int lanes = 8;
        for (const Band& B : T.bands) if (2 * B.nseg_est + 2 > lanes) lanes = 2 * B.nseg_est + 2;       // root + pole per segment
        lanes = (lanes + 7) & ~7;
        if (lanes > REC_CAP) lanes = REC_CAP;
        const size_t rec_bytes = (size_t)nbands * REC_CAP * sizeof(Record);
        if (rec_bytes > G.recs_cap) {
            if (G.d_recs) cudaFree(G.d_recs);
            G.d_recs = nullptr;
            G.recs_cap = rec_bytes + rec_bytes / 2;
            RGB_CUDA(cudaMalloc((void**)&G.d_recs, G.recs_cap));
        }
        int* d_nrec = (int*)G.d_res;
        unsigned long long* d_slots = (unsigned long long*)(G.d_res + cnt_bytes);
        Record* d_rec = (Record*)G.d_recs;
        tamcmc_rgb_search_kernel<<<dim3((unsigned)nbands, 2), 512, 0, G.stream>>>((const Band*)(G.d_in + off[0]), d_rec, d_nrec, d_hdr);
        const long nthreads = (long)npairs * lanes;
        tamcmc_rgb_pairs_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, G.stream>>>(
            (const Band*)(G.d_in + off[0]), (const Pair*)(G.d_in + off[1]), npairs, lanes, d_rec, d_nrec, d_slots, d_hdr);
    }
    RGB_CUDA(cudaGetLastError());
    RGB_CUDA(cudaStreamWaitEvent(G.stream, G.ev_ksi, 0));
    RGB_CUDA(cudaMemcpyAsync(G.h_out, G.d_out, hdr_bytes, cudaMemcpyDeviceToHost, G.stream));
    if (res_bytes) RGB_CUDA(cudaMemcpyAsync(G.h_res, G.d_res, res_bytes, cudaMemcpyDeviceToHost, G.stream));
    G.launched = true;
    return TAMCMC_OK;
}

// wait for the group's results, then the rest of the model function for its chains (host); flagged chains are solved by the host code
int rgb_collect(tamcmc_gpu_rgb* h, RgbGroup& G, int model_id, const double* params, int params_stride, const int* plength, double step,
                int capacity, double* rows_out, int row_stride, int* nmodes_out, int* status_out, int* path_out)
{
    if (G.launched) RGB_CUDA(cudaStreamSynchronize(G.stream));
    const OutHdr* h_hdr = (const OutHdr*)G.h_out;
    const int nbands = (int)G.task.bands.size();
    const int* h_nrec = (const int*)G.h_res;
    const unsigned long long* h_slots = (const unsigned long long*)(G.h_res + align16((size_t)nbands * 4));
    std::vector<int>& dev = h->on_device;
    const int c0 = G.c0, c1 = G.c1;
    std::vector<double> norm((size_t)h->max_chains, -1.0);
    // one parallel region, three stages with a barrier between them
    std::vector<std::pair<int, int>> blocks;
#pragma omp parallel
    {
        // 3a, one chain per thread: the exact zeta normalisation at the grid points the device found, and the mixed-mode frequencies
#pragma omp for schedule(dynamic, 1)
        for (int c = c0; c < c1; c++) {
            if (status_out[c] != TAMCMC_OK) continue;
            if (dev[(size_t)c] && !G.launched) dev[(size_t)c] = 0;
            if (!dev[(size_t)c]) continue;
            const OutHdr& O = h_hdr[c];
            if (O.flag != 0 || O.ntop < 1) { if (path_out) path_out[c] = O.flag ? O.flag : RGB_FLAG_NONFINITE; dev[(size_t)c] = 0; continue; }
            if (O.ntop <= TOP_CAP) norm[(size_t)c] = ksi_norm_at(h->preps[(size_t)c], O.top, O.ntop);      // else: finish() takes the maximum itself
            std::vector<double> cand;                      // the solutions of the chain's bands: one per record the ratio test kept
            for (int b = G.band_lo[(size_t)c]; b < G.band_hi[(size_t)c]; b++) {
                const int n = h_nrec[b] < REC_CAP ? h_nrec[b] : REC_CAP;
                for (int r = 0; r < n; r++) {
                    const unsigned long long v = h_slots[(size_t)b * REC_CAP + r];
                    if (v != SLOT_EMPTY) { double d; std::memcpy(&d, &v, 8); cand.push_back(d); }
                }
            }
            status_out[c] = finish_modes(h->preps[(size_t)c], true, cand.data(), (int)cand.size());
            if (path_out) path_out[c] = 0;
        }
        // 3b, the zeta sums at the mixed modes: blocks of 8 frequencies of all chains over all threads
#pragma omp single
        {
            for (int c = c0; c < c1; c++)
                if (dev[(size_t)c] && status_out[c] == TAMCMC_OK)
                    for (int b = 0; b < ksi_blocks(h->preps[(size_t)c]); b++) blocks.push_back({c, b});
        }
#pragma omp for schedule(dynamic, 1)
        for (long k = 0; k < (long)blocks.size(); k++) ksi_block_compute(h->preps[(size_t)blocks[(size_t)k].first], blocks[(size_t)k].second);
        // 3c, one chain per thread: heights, widths, splittings, rows
#pragma omp for schedule(dynamic, 1)
        for (int c = c0; c < c1; c++) {
            if (status_out[c] != TAMCMC_OK) continue;
            double* row = rows_out + (size_t)c * row_stride;
            int nm = 0;
            if (dev[(size_t)c]) {
                ksi_done(h->preps[(size_t)c]);
                status_out[c] = finish(h->preps[(size_t)c], true, nullptr, 0, norm[(size_t)c], capacity, row, &nm);
            } else {
                status_out[c] = prepare(h->preps[(size_t)c], model_id, params + (size_t)c * params_stride, plength, step, false);
                if (status_out[c] == TAMCMC_OK) status_out[c] = finish(h->preps[(size_t)c], false, nullptr, 0, -1.0, capacity, row, &nm);
            }
            if (nmodes_out) nmodes_out[c] = nm;
        }
    }
    for (int c = c0; c < c1; c++) { h->chains_total++; if (!dev[(size_t)c] && status_out[c] == TAMCMC_OK) h->chains_host++; }
    return TAMCMC_OK;
}

}  // namespace

extern "C" {

const char* tamcmc_gpu_rgb_last_error(void) { return g_err.c_str(); }

void tamcmc_gpu_rgb_destroy(tamcmc_gpu_rgb* h);

// Replaces: nothing the reference has as one call -- the per-chain OpenMP fan-out of generate_model (model_def.cpp:466-482) entering
// model_RGB_asympt_aj_*Width_HarveyLike_v4 once per chain; here the set-up of all chains of a step is one batched device solve.
int tamcmc_gpu_rgb_create(tamcmc_gpu_rgb** out, int device, int max_chains)
{
    if (!out || max_chains < 1) return TAMCMC_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { g_err = "no usable CUDA device"; return TAMCMC_ERR_CUDA; }
    RGB_CUDA(cudaSetDevice(device));
    tamcmc_gpu_rgb* h = new tamcmc_gpu_rgb();
    h->device = device; h->max_chains = max_chains;
    h->out_bytes = align16((size_t)max_chains * sizeof(OutHdr));
    cudaError_t e = cudaSuccess;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);        // (greatest priority = lowest number)
    for (int g = 0; g < 2; g++) {
        RgbGroup& G = h->grp[g];
        const int prio = (g == 0) ? prio_hi : prio_lo;            // the group the host waits for first goes first on the device
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&G.stream, cudaStreamNonBlocking, prio);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&G.stream2, cudaStreamNonBlocking, prio);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&G.ev_in, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&G.ev_ksi, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc((void**)&G.d_out, h->out_bytes);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&G.h_out, h->out_bytes);
    }
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); tamcmc_gpu_rgb_destroy(h); return TAMCMC_ERR_CUDA; }
    for (int c = 0; c < max_chains; c++) h->preps.push_back(prep_new());
    *out = h;
    return TAMCMC_OK;
}

void tamcmc_gpu_rgb_destroy(tamcmc_gpu_rgb* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (Prep* p : h->preps) prep_free(p);
    for (RgbGroup& G : h->grp) {
        if (G.d_in) cudaFree(G.d_in);
        if (G.h_in) cudaFreeHost(G.h_in);
        if (G.d_out) cudaFree(G.d_out);
        if (G.h_out) cudaFreeHost(G.h_out);
        if (G.d_res) cudaFree(G.d_res);
        if (G.h_res) cudaFreeHost(G.h_res);
        if (G.d_recs) cudaFree(G.d_recs);
        if (G.d_vals) cudaFree(G.d_vals);
        if (G.ev_in) cudaEventDestroy(G.ev_in);
        if (G.ev_ksi) cudaEventDestroy(G.ev_ksi);
        if (G.stream) cudaStreamDestroy(G.stream);
        if (G.stream2) cudaStreamDestroy(G.stream2);
    }
    delete h;
}

// params: [nchains][params_stride] parameter vectors in the layout of models 25 / 27 (tamcmc_host_expand_rgb_v4); rows_out:
// [nchains][row_stride] mode-table rows of `capacity` modes (row_stride >= TAMCMC_MT_HEADER + Nnoise + TAMCMC_MT_STRIDE * capacity).
// status_out[c]: what tamcmc_host_expand_rgb_v4 would have returned for chain c; path_out[c] (may be NULL): 0 = solved on the device,
// otherwise the rgb_solver.cuh flag bits (or -1: not exportable) that sent the chain to the host solver.  Returns TAMCMC_ERR_CUDA /
// TAMCMC_ERR_ARG for a failed call, else TAMCMC_OK (per-chain outcomes are in status_out).
// From 16 chains on they go through in two groups: both are enqueued at once, and the host finishes the first group's rows (sort + unique,
// zeta at the mixed modes, heights / widths / splittings) while the device solves the second.
int tamcmc_gpu_rgb_expand(tamcmc_gpu_rgb* h, int model_id, const double* params, int params_stride, const int* plength, double step,
                          int nchains, int capacity, double* rows_out, int row_stride, int* nmodes_out, int* status_out, int* path_out)
{
    if (!h || !params || !plength || !rows_out || !status_out || nchains < 1 || nchains > h->max_chains || capacity < 1) return TAMCMC_ERR_ARG;
    if (row_stride < TAMCMC_MT_HEADER + plength[8] + TAMCMC_MT_STRIDE * capacity) return TAMCMC_ERR_ARG;
    RGB_CUDA(cudaSetDevice(h->device));
    timespec t0, t1, t2, t3;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    // ---- stage 1 (host, one chain per thread): everything in front of the pair loop, and the chain's device task ----
    h->on_device.assign((size_t)nchains, 0);
    if ((int)h->chain_task.size() < nchains) h->chain_task.resize((size_t)nchains);
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < nchains; c++) {
        status_out[c] = prepare(h->preps[(size_t)c], model_id, params + (size_t)c * params_stride, plength, step, true);
        if (path_out) path_out[c] = -1;
        if (nmodes_out) nmodes_out[c] = 0;
        DeviceTask& Tc = h->chain_task[(size_t)c];
        Tc.clear();
        if (status_out[c] == TAMCMC_OK && export_task(h->preps[(size_t)c], c, Tc)) h->on_device[(size_t)c] = 1;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    // ---- stage 2 (device): pair loop + zeta normalisation, both groups enqueued back to back ----
    // (measured at 10 chains: the second group's enqueue costs the host what the overlap saves -- one group below 16 chains)
    const int ngroups = (nchains >= 16) ? 2 : 1;
    const int split = (ngroups == 2) ? (nchains + 1) / 2 : nchains;
    h->grp[0].c0 = 0; h->grp[0].c1 = split; h->grp[1].c0 = split; h->grp[1].c1 = nchains;
    for (int g = 0; g < ngroups; g++) { const int rc = rgb_enqueue(h, h->grp[g]); if (rc) return rc; }
    // ---- stage 3 (host): group by group, as the results arrive ----
    bool first = true;
    for (int g = 0; g < ngroups; g++) {
        RgbGroup& G = h->grp[g];
        if (first && G.launched) { RGB_CUDA(cudaStreamSynchronize(G.stream)); clock_gettime(CLOCK_MONOTONIC, &t2); first = false; }
        const int rc = rgb_collect(h, G, model_id, params, params_stride, plength, step, capacity, rows_out, row_stride, nmodes_out, status_out, path_out);
        if (rc) return rc;
    }
    if (first) clock_gettime(CLOCK_MONOTONIC, &t2);
    clock_gettime(CLOCK_MONOTONIC, &t3);
    auto ms = [](const timespec& a, const timespec& b) { return 1e3 * (double)(b.tv_sec - a.tv_sec) + 1e-6 * (double)(b.tv_nsec - a.tv_nsec); };
    h->last_ms[0] = ms(t0, t1); h->last_ms[1] = ms(t1, t2); h->last_ms[2] = ms(t2, t3); h->last_ms[3] = ms(t0, t3);
    return TAMCMC_OK;
}

// host-clock milliseconds of the stages of the last tamcmc_gpu_rgb_expand: prepare, device (copies + kernels + sync), finish, total
void tamcmc_gpu_rgb_timings(const tamcmc_gpu_rgb* h, double out[4])
{
    for (int k = 0; k < 4; k++) out[k] = h ? h->last_ms[k] : 0.0;
}

// how many chain set-ups this handle was asked for, and how many of them the device flagged and the host solver of the library produced
// (reported per call in path_out as well): the device path is the one that runs, and a caller can show it
void tamcmc_gpu_rgb_counts(const tamcmc_gpu_rgb* h, long* chains_total, long* chains_host)
{
    if (chains_total) *chains_total = h ? h->chains_total : 0;
    if (chains_host) *chains_host = h ? h->chains_host : 0;
}

// TEST HOOK (no GPU needed): the segment decomposition of rgb_solver.cuh run on the host for ONE chain, with glibc's tan / atan
// (exact_trig = 0: must reproduce tamcmc_host_expand_rgb_v4 bit for bit) or the correctly rounded ones of dd_math.cuh (1: what the
// device computes), solutions reduced to the minimum per slot like the kernel does.  Every band index's local grid from the error-free
// emulation of the reference's long double arithmetic is compared with the real thing (flag 128 on any difference).
// Not a product path: tamcmc_gpu_rgb_expand never calls it.
int tamcmc_host_rgb_expand_emulated(int model_id, const double* params, const int* plength, double step, int capacity, double* row_out,
                                    int* nmodes_out, int exact_trig, int* flags_out)
{
    if (!params || !plength || !row_out || capacity < 1) return TAMCMC_ERR_ARG;
    Prep* P = prep_new();
    int rc = prepare(P, model_id, params, plength, step, true);
    DeviceTask T;
    if (rc == TAMCMC_OK && !export_task(P, 0, T)) { prep_free(P); if (flags_out) *flags_out = -1; return TAMCMC_ERR_MODEL; }
    int flag = 0;
    std::vector<double> cand;
    if (rc == TAMCMC_OK) {
        std::vector<double> slots((size_t)T.nslots, -1.0);
        for (const Band& B : T.bands)
            for (int i = 0; i < B.nband; i++) {
                double lo, hi, rlo, rhi; int n;
                const bool ok = local_grid_ext(band_nu(B, i), B.resol2, B.Dh, B.Dl, lo, hi, n);
                const long rn = local_grid_reference(P, band_nu(B, i), &rlo, &rhi);
                if (!ok) { flag |= RGB_FLAG_EXT; continue; }
                if (lo != rlo || hi != rhi || !((rn < 2 && n < 2) || rn == (long)n)) flag |= 128;
            }
        std::vector<std::vector<Record>> recs(T.bands.size());
        for (size_t b = 0; b < T.bands.size(); b++) {
            const Band& B = T.bands[b];
            if (B.nband == 0 || B.rep_inv_g == 0.0) continue;
            double m_hi = 0, nu0 = 0, bstep = 0;
            const int nseg = pair_segments(B, B.rep_inv_g, m_hi, nu0, bstep, flag);
            for (int j = 0; j < nseg; j++) {
                auto emit = [&](const Record& R) { recs[b].push_back(R); };
                if (exact_trig) pair_segment<TrigLib, TrigCR>(B, B.rep_inv_g, j, nseg, m_hi, nu0, bstep, emit, flag);
                else pair_segment<TrigLib, TrigLib>(B, B.rep_inv_g, j, nseg, m_hi, nu0, bstep, emit, flag);
            }
            if ((int)recs[b].size() > REC_CAP) flag |= RGB_FLAG_OVERFLOW;
        }
        for (const Pair& Q : T.pairs) {
            const Band& B = T.bands[(size_t)Q.band];
            for (const Record& R : recs[(size_t)Q.band]) {
                double sol;
                const bool got = exact_trig ? record_eval<TrigCR>(B, Q.inv_g, Q.nu_g, R, sol, flag) : record_eval<TrigLib>(B, Q.inv_g, Q.nu_g, R, sol, flag);
                if (got) { double& v = slots[(size_t)(B.slot_off + R.idx)]; if (v < 0.0 || sol < v) v = sol; }
            }
        }
        for (double v : slots) if (v >= 0.0) cand.push_back(v);
        rc = finish(P, true, cand.data(), (int)cand.size(), -1.0, capacity, row_out, nmodes_out);
    }
    if (flags_out) *flags_out = flag;
    prep_free(P);
    return rc;
}

}  // extern "C"
