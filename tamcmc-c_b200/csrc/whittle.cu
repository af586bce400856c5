// whittle.cu -- fused power-spectrum model + Whittle chi^2(2 dof) log-likelihood (sm_100a, FP64).
//
// Persistent, warp-specialised kernel.  Each CTA has one PRODUCER warp and 8 CONSUMER warps:
//
//  * the producer pops (star, chain, tile) work items from the heavy-first queue the expander built,
//    classifies the chain's modes against the tile, writes the tile-local component list into one of two
//    shared-memory segments, builds the Taylor series of the Harvey background for the tile, and fetches the
//    tile's x and y (2 x 8 KB) with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) -- all one tile
//    ahead of the consumers, synchronised with full/empty mbarriers;
//  * the 256 consumer threads own 4 bins each.  The model spectrum M never exists in HBM: every thread keeps
//    the running Lorentzian sum of a bin as ONE fraction N/D,
//        sum_k A_k / (1 + 4 (x - nu_k)^2 / Gamma_k^2)  =  N / D ,
//    merging a component with 2 FMAs for its scaled denominator t' = (1 + e^2)/A_k and 1 FMA + 1 MUL for
//    (N, D) <- (N t' + D, D t').  That replaces the reference's FP64 divide per (component, bin)
//    (build_lorentzian.cpp:151: cwiseInverse) by 4 FP64-pipe instructions, with one divide per bin at the
//    end.  Exponents of (N, D) are renormalised with 4 integer ops per bin every 16 components.  The background
//    (noise_models.cpp:15-39) is a 9th-degree polynomial per tile (or exact exp() per bin near x = 0), the
//    Whittle terms y/M + ln M (likelihoods.cpp:23) are reduced in registers, by warp shuffles and a
//    fixed-shape block tree; the last CTA to finish a chain sums its per-tile partials in index order, so
//    results are bitwise reproducible run to run whatever the scheduling.
//
// Mode windows (ModeRec.i0/i1) come bit-exact from the expander; a mode whose window covers the whole tile
// takes the mask-free fast path, a mode that only partly overlaps it takes the masked general path.
#include "tamcmc_dev.h"
#include "kernels.h"
#include <cuda_runtime.h>
#include <math.h>

namespace {

constexpr int NC = TAMCMC_CONSUMERS;                       // consumer threads
constexpr int NT = TAMCMC_THREADS;                         // + producer warp
constexpr int BPT = TAMCMC_BINS_PER_THREAD;
constexpr int TILE = TAMCMC_TILE;
constexpr int GROUP = 16;                                  // fast components merged between two renormalisations
constexpr int NB = TAMCMC_BG_TERMS;
constexpr int CAPF = 352;                                  // fast components per segment
constexpr int CAPG = 24;                                   // general entries per segment
constexpr int CAPH = 64;                                   // mode headers per segment (asym fast path)
constexpr int PLCAP = 480;                                 // modes classified per producer pass (multiple of 96)

struct ModeHdr { double qa, qb, qc; int begin, count; };
struct GenEntry { double s, c, aadd, num, qa, qb, qc; int lo, hi; };

enum { SEG_FIRST = 1, SEG_LAST = 2, SEG_DONE = 4, SEG_BGSERIES = 8, SEG_ASYM = 16 };

struct __align__(16) Segment {
    double x[TILE];          // TMA destination (first segment of a tile)
    double y[TILE];
    double2 sc[CAPF];        // fast list: {s', c'} with e' = fma(u, s', c')
    double a[CAPF];          // fast list: 1/A
    GenEntry gen[CAPG];
    ModeHdr hdr[CAPH];
    double bg[NB];           // background Taylor coefficients in u (incl. nothing of N0)
    double xc, N0;
    int nfast, ngen, nhdr, flags;
    int sc_index, tile, nvalid, lb0;
    long long off;           // offset of the tile in the concatenated arrays
};

struct Smem {
    Segment seg[2];
    unsigned long long full[2], empty[2];
    int4 plist[PLCAP];       // producer scratch: modes overlapping the tile
    double red[NC / 32];
    int is_last;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok = 0;
    const unsigned addr = smem_u32(bar);
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    }
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); }

// (N, D) *= 2^-k with k = exponent(D): exact, integer pipe only.  D > 0 always; N >= 0, and N == 0
// only while D == 1 (nothing merged yet), where k == 0.
__device__ __forceinline__ void renorm(double& N, double& D)
{
    const int hiD = __double2hiint(D);
    const int k = (hiD & 0x7ff00000) - 0x3ff00000;
    D = __hiloint2double(hiD - k, __double2loint(D));
    N = __hiloint2double(__double2hiint(N) - k, __double2loint(N));
}

// ------------------------------------------------------------------------------------------------
// producer warp
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned pop_item(const WhittleArgs& A, int lane)
{
    unsigned idx = 0;
    if (lane == 0) idx = atomicAdd(&A.qctl->head, 1u);
    return __shfl_sync(0xffffffffu, idx, 0);
}

// publish the open segment: header by lane 0, then every lane arrives on the full barrier; the first
// segment of a tile also carries the TMA bulk loads of x and y
__device__ __forceinline__ void publish_segment(Smem& sm, int b, int lane, bool first, int flags, int nf, int ng, int nhd,
                                                int sc, int tile, int nvalid, int lb0, long long off, double xc, double N0,
                                                const double* xs, const double* ys)
{
    Segment* sg = &sm.seg[b];
    if (lane == 0) {
        sg->nfast = nf; sg->ngen = ng; sg->nhdr = nhd; sg->flags = flags;
        sg->sc_index = sc; sg->tile = tile; sg->nvalid = nvalid; sg->lb0 = lb0; sg->off = off; sg->xc = xc; sg->N0 = N0;
    }
    __syncwarp();
    if (first && lane == 0) {
        mbar_arrive_expect_tx(&sm.full[b], 2u * TILE * sizeof(double));
        tma_load_1d(sg->x, xs, TILE * sizeof(double), &sm.full[b]);
        tma_load_1d(sg->y, ys, TILE * sizeof(double), &sm.full[b]);
    } else mbar_arrive(&sm.full[b]);
}

__device__ void producer_loop(const WhittleArgs& A, Smem& sm, int lane)
{
    unsigned use[2] = {0, 0};         // how many times each segment buffer has been filled
    int b = 0;
    const unsigned nheavy = A.qctl->count[0], nlight = A.qctl->count[1], ntot = nheavy + nlight;
    const unsigned below = (1u << lane) - 1u;
    unsigned idx = pop_item(A, lane);
    for (;;) {
        if (idx >= ntot) {
            if (use[b]) mbar_wait(&sm.empty[b], (use[b] - 1) & 1);
            if (lane == 0) sm.seg[b].flags = SEG_DONE;
            __syncwarp();
            mbar_arrive(&sm.full[b]);
            return;
        }
        const unsigned item = (idx < nheavy) ? A.queue[idx] : A.queue[A.qcap + (idx - nheavy)];
        idx = pop_item(A, lane);                      // next item: the atomic's latency hides behind this tile
        const int sc = (int)(item / (unsigned)A.tiles_stride);
        const int tile = (int)(item - (unsigned)sc * (unsigned)A.tiles_stride);
        const StarDesc* sd = A.stars + sc / A.Nchains;
        const TileRec* tr = A.tilerec + item;
        // one round trip: star fields, tile record, chain flags
        const long long soff = sd->off;
        const int Nloc = sd->Nloc, bin0 = sd->bin0, nmodes = sd->nmodes_cap;
        const double xc = tr->xc;
        const int series_ok = tr->series_ok;
        const double bgk = (lane < NB) ? tr->bg[lane] : 0.0;
        const bool asym = A.asym_flag[sc] != 0;
        const double N0 = A.noise[sc].N0;
        const ModeRec* modes = A.modes + (size_t)sc * A.modes_stride;
        const CompRec* comps = A.comps + (size_t)sc * A.modes_stride * TAMCMC_MAX_COMP_PER_MODE;

        const int lb0 = tile * TILE;
        const int g0 = bin0 + lb0;
        const int nvalid = min(TILE, Nloc - lb0);
        const int gend = g0 + nvalid;
        const long long off = soff + lb0;
        const double* xs = A.x + off;
        const double* ys = A.y + off;

        bool first = true;
        int nf = 0, ng = 0, nhd = 0;          // fill levels of the open segment (warp-uniform)
        if (use[b]) mbar_wait(&sm.empty[b], (use[b] - 1) & 1);

        for (int chunk = 0; chunk < nmodes; chunk += PLCAP) {
            // ---- pass 1: classify the chain's modes against the tile (16-byte headers), ordered compaction ----
            int nlist = 0;
            const int cend = min(nmodes, chunk + PLCAP);
            for (int base = chunk; base < cend; base += 96) {
                int4 h[3];
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const int mi = base + 32 * r + lane;
                    h[r] = (mi < cend) ? *reinterpret_cast<const int4*>(modes + mi) : make_int4(0, 0, 0, 0);
                }
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const int mi = base + 32 * r + lane;
                    // h = {i0, i1, ncomp, nfast}
                    const bool ov = (h[r].z > 0) && (h[r].x < gend) && (h[r].y > g0);
                    const unsigned mk = __ballot_sync(0xffffffffu, ov);
                    if (ov) {
                        const int full = (h[r].x <= g0 && h[r].y >= gend) ? 1 : 0;
                        sm.plist[nlist + __popc(mk & below)] = make_int4(mi | (full << 30), h[r].x, h[r].y, h[r].z | (h[r].w << 8));
                    }
                    nlist += __popc(mk);
                }
            }
            __syncwarp();
            // ---- pass 2: emit the listed modes, 32 (or 3) at a time ----
            for (int r0 = 0; r0 < nlist; r0 += 32) {
                const bool have = (r0 + lane) < nlist;
                const int4 pe = have ? sm.plist[r0 + lane] : make_int4(0, 0, 0, 0);
                const int mi = pe.x & 0x3fffffff;
                const bool full = ((pe.x >> 30) & 1) != 0;
                const int ncomp = have ? (pe.w & 0xff) : 0;
                const int nfast = (have && full) ? (pe.w >> 8) : 0;
                const int ngen = ncomp - nfast;
                int sub_lo = 0;
                while (sub_lo < 32) {
                    int sub_hi = 32;
                    bool mine = lane >= sub_lo;
                    int tf = mine ? nfast : 0, tg = mine ? ngen : 0, th = (mine && nfast > 0) ? 1 : 0;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        tf += __shfl_xor_sync(0xffffffffu, tf, d);
                        tg += __shfl_xor_sync(0xffffffffu, tg, d);
                        th += __shfl_xor_sync(0xffffffffu, th, d);
                    }
                    if (tg > CAPG) {
                        // too many general entries for one segment: 3 modes at a time (3 x 7 <= CAPG)
                        sub_hi = sub_lo + 3;
                        mine = lane >= sub_lo && lane < sub_hi;
                        tf = mine ? nfast : 0; tg = mine ? ngen : 0; th = (mine && nfast > 0) ? 1 : 0;
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) {
                            tf += __shfl_xor_sync(0xffffffffu, tf, d);
                            tg += __shfl_xor_sync(0xffffffffu, tg, d);
                            th += __shfl_xor_sync(0xffffffffu, th, d);
                        }
                    }
                    if (nf + tf > CAPF || ng + tg > CAPG || nhd + th > CAPH) {
                        // close the open segment (not the last of the tile), continue in the other buffer
                        publish_segment(sm, b, lane, first, (first ? SEG_FIRST : 0) | (asym ? SEG_ASYM : 0), nf, ng, nhd,
                                        sc, tile, nvalid, lb0, off, xc, N0, xs, ys);
                        use[b]++; b ^= 1; first = false;
                        if (use[b]) mbar_wait(&sm.empty[b], (use[b] - 1) & 1);
                        nf = ng = nhd = 0;
                    }
                    Segment* sg = &sm.seg[b];
                    // exclusive offsets inside the segment
                    const int mf = mine ? nfast : 0, mg = mine ? ngen : 0, mh = (mine && nfast > 0) ? 1 : 0;
                    int of = mf, og = mg, oh = mh;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int a1 = __shfl_up_sync(0xffffffffu, of, d), a2 = __shfl_up_sync(0xffffffffu, og, d), a3 = __shfl_up_sync(0xffffffffu, oh, d);
                        if (lane >= d) { of += a1; og += a2; oh += a3; }
                    }
                    of = nf + of - mf; og = ng + og - mg; oh = nhd + oh - mh;
                    if (mine && ncomp > 0) {
                        const CompRec* cp = comps + (size_t)mi * TAMCMC_MAX_COMP_PER_MODE;
                        double qa = 0.0, qb = 1.0, qc = 0.0;
                        if (asym || ngen > 0) {
                            const ModeRec* mr = modes + mi;
                            qa = mr->qa; qb = mr->qb0 + xc * qa; qc = mr->qc;
                        }
                        if (nfast > 0) {
                            ModeHdr hh; hh.qa = qa; hh.qb = qb; hh.qc = qc; hh.begin = of; hh.count = nfast;
                            sg->hdr[oh] = hh;
                        }
                        // components in two batches of up to 4: the loads of a batch are independent
#pragma unroll
                        for (int k0 = 0; k0 < 8; k0 += 4) {
                            double cnu[4], cs[4], ca[4];
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) {
                                const int k = k0 + kk;
                                if (k < ncomp) { cnu[kk] = cp[k].nu; cs[kk] = cp[k].s; ca[kk] = cp[k].a; }
                            }
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) {
                                const int k = k0 + kk;
                                if (k < ncomp) {
                                    const double cc = -(cnu[kk] - xc) * cs[kk];
                                    if (k < nfast) {
                                        sg->sc[of + k] = make_double2(cs[kk], cc);
                                        sg->a[of + k] = ca[kk];
                                    } else {
                                        // components are stored FAST-first; a FAST one lands here only on a window edge
                                        const bool ff = k < (pe.w >> 8);
                                        GenEntry ge;
                                        ge.s = cs[kk]; ge.c = cc; ge.aadd = ff ? ca[kk] : 1.0; ge.num = ff ? 1.0 : ca[kk];
                                        ge.qa = qa; ge.qb = qb; ge.qc = qc;
                                        ge.lo = pe.y - g0; ge.hi = pe.z - g0;
                                        sg->gen[og + (k - nfast)] = ge;
                                    }
                                }
                            }
                        }
                    }
                    nf += tf; ng += tg; nhd += th;
                    sub_lo = sub_hi;
                }
            }
            __syncwarp();
        }

        // last segment of the tile: background record, then publish
        if (lane < NB) sm.seg[b].bg[lane] = bgk;
        publish_segment(sm, b, lane, first, (first ? SEG_FIRST : 0) | SEG_LAST | (asym ? SEG_ASYM : 0) | (series_ok ? SEG_BGSERIES : 0),
                        nf, ng, nhd, sc, tile, nvalid, lb0, off, xc, N0, xs, ys);
        use[b]++; b ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------
// consumer warps
// ------------------------------------------------------------------------------------------------
template <bool WRITE_MODEL>
__device__ void consumer_loop(const WhittleArgs& A, Smem& sm, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    unsigned use[2] = {0, 0};
    int b = 0;
    double u[BPT], N[BPT], D[BPT], yv[BPT];
#pragma unroll
    for (int j = 0; j < BPT; j++) { u[j] = 0; N[j] = 0; D[j] = 1; yv[j] = 0; }

    for (;;) {
        mbar_wait(&sm.full[b], use[b] & 1);
        use[b]++;
        const Segment& sg = sm.seg[b];
        const int flags = sg.flags;
        if (flags & SEG_DONE) return;
        const bool asym = (flags & SEG_ASYM) != 0;

        if (flags & SEG_FIRST) {
            // this thread's 4 bins: b(j) = 2*tid + 512*(j>>1) + (j&1), read as 128-bit pairs
            const double xc = sg.xc;
#pragma unroll
            for (int pj = 0; pj < BPT / 2; pj++) {
                const double2 v = *reinterpret_cast<const double2*>(&sg.x[2 * tid + 2 * NC * pj]);
                const double2 w = *reinterpret_cast<const double2*>(&sg.y[2 * tid + 2 * NC * pj]);
                u[2 * pj] = v.x - xc; u[2 * pj + 1] = v.y - xc;
                yv[2 * pj] = w.x; yv[2 * pj + 1] = w.y;
            }
#pragma unroll
            for (int j = 0; j < BPT; j++) { N[j] = 0.0; D[j] = 1.0; }
        }

        // ---------- fast path: windows cover the whole tile, no masks ----------
        const int tot_fast = sg.nfast;
        if (!asym) {
            int k = 0;
            for (; k + GROUP <= tot_fast; k += GROUP) {
#pragma unroll
                for (int kk = 0; kk < GROUP; kk++) {
                    const double2 p = sg.sc[k + kk];
                    const double a = sg.a[k + kk];
#pragma unroll
                    for (int j = 0; j < BPT; j++) {
                        const double e = fma(u[j], p.x, p.y);
                        const double t = fma(e, e, a);
                        N[j] = fma(N[j], t, D[j]);
                        D[j] *= t;
                    }
                }
#pragma unroll
                for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
            }
            for (; k < tot_fast; k++) {
                const double2 p = sg.sc[k];
                const double a = sg.a[k];
#pragma unroll
                for (int j = 0; j < BPT; j++) {
                    const double e = fma(u[j], p.x, p.y);
                    const double t = fma(e, e, a);
                    N[j] = fma(N[j], t, D[j]);
                    D[j] *= t;
                }
            }
#pragma unroll
            for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
        } else {
            // asymmetric Lorentzians (build_lorentzian.cpp:153-157): every component of a mode is
            // multiplied by q(x) = (1 + asym (x/fc - 1))^2 + (Gamma asym / (2 fc))^2
            int since = 0;
            const int nhdr = sg.nhdr;
            for (int m = 0; m < nhdr; m++) {
                const ModeHdr h = sg.hdr[m];
                double q[BPT];
#pragma unroll
                for (int j = 0; j < BPT; j++) { const double w = fma(u[j], h.qa, h.qb); q[j] = fma(w, w, h.qc); }
                for (int k = h.begin; k < h.begin + h.count; k++) {
                    const double2 p = sg.sc[k];
                    const double a = sg.a[k];
#pragma unroll
                    for (int j = 0; j < BPT; j++) {
                        const double e = fma(u[j], p.x, p.y);
                        const double t = fma(e, e, a);
                        N[j] = fma(N[j], t, q[j] * D[j]);
                        D[j] *= t;
                    }
                }
                since += h.count;
                if (since + TAMCMC_MAX_COMP_PER_MODE > GROUP) {
                    since = 0;
#pragma unroll
                    for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
        }

        // ---------- general path: window edges and extreme-dynamic-range components.  The tile is cut in
        // 512-bin segments (one per register pair of every thread): a segment the window covers is merged
        // unmasked, a segment holding a window edge under a per-bin mask, others are skipped; all three
        // decisions are CTA-uniform. ----------
        const int tot_gen = sg.ngen;
        for (int g = 0; g < tot_gen; g++) {
            const GenEntry ge = sg.gen[g];
#pragma unroll
            for (int pj = 0; pj < BPT / 2; pj++) {
                const int s0 = 2 * NC * pj, s1 = s0 + 2 * NC;
                if (ge.hi <= s0 || ge.lo >= s1) continue;
                const bool whole = (ge.lo <= s0 && ge.hi >= s1);
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int j = 2 * pj + r;
                    const int bb = 2 * tid + s0 + r;
                    const bool in = whole || ((bb >= ge.lo) && (bb < ge.hi));
                    const double e = fma(u[j], ge.s, ge.c);
                    const double t = fma(e, e, ge.aadd);
                    double nd = ge.num * D[j];
                    if (asym) { const double w = fma(u[j], ge.qa, ge.qb); nd *= fma(w, w, ge.qc); }
                    const double Nn = fma(N[j], t, nd);
                    const double Dn = D[j] * t;
                    if (in) { N[j] = Nn; D[j] = Dn; }
                    renorm(N[j], D[j]);
                }
            }
        }

        if (flags & SEG_LAST) {
            const int sc = sg.sc_index, tile = sg.tile, nvalid = sg.nvalid, lb0 = sg.lb0;
            const double N0 = sg.N0;
            const NoiseRec* nz = A.noise + sc;
            double bgv[BPT];
            if (flags & SEG_BGSERIES) {
                double cf[NB];
#pragma unroll
                for (int k = 0; k < NB; k++) cf[k] = sg.bg[k];
#pragma unroll
                for (int j = 0; j < BPT; j++) {
                    double acc = cf[NB - 1];
#pragma unroll
                    for (int k = NB - 2; k >= 0; k--) acc = fma(acc, u[j], cf[k]);
                    bgv[j] = acc + N0;
                }
            } else {
                // near x = 0 / near a singularity of a term: evaluate every bin exactly and merge the terms into
                // the same fraction: (1e-3 tau x)^p = exp(p (ln(1e-3 tau) + ln x)); the clamp keeps D finite
                const double* lx = A.lnx + sg.off;
                const int nh = nz->nh;
                double lnx[BPT];
#pragma unroll
                for (int pj = 0; pj < BPT / 2; pj++) {
                    const double2 v = *reinterpret_cast<const double2*>(lx + 2 * tid + 2 * NC * pj);
                    lnx[2 * pj] = v.x; lnx[2 * pj + 1] = v.y;
                }
                for (int h = 0; h < nh; h++) {
                    const double H = nz->H[h], ls = nz->lnsc[h], pw = nz->pw[h];
#pragma unroll
                    for (int j = 0; j < BPT; j++) {
                        const double arg = fmin(pw * (ls + lnx[j]), 70.0);
                        const double z = (pw == 0.0) ? 1.0 : exp(arg);
                        const double t = 1.0 + z;
                        N[j] = fma(N[j], t, H * D[j]);
                        D[j] *= t;
                    }
                }
#pragma unroll
                for (int j = 0; j < BPT; j++) bgv[j] = N0;
            }

            // ---------- M = N/D + background; Whittle terms ----------
            double s1 = 0.0, prod = 1.0;
#pragma unroll
            for (int j = 0; j < BPT; j++) {
                const int bb = 2 * tid + 2 * NC * (j >> 1) + (j & 1);
                const double num = fma(bgv[j], D[j], N[j]);
                if (WRITE_MODEL) { if (bb < nvalid) A.model_out[lb0 + bb] = num / D[j]; }
                if (bb < nvalid) {
                    const double minv = D[j] / num;           // 1/M_i
                    s1 = fma(yv[j], minv, s1);                // y_i / M_i
                    prod *= minv;                             // ln M_i summed as -ln(prod)
                }
            }
            double v = s1 - log(prod);
            const int ntiles = A.stars[sc / A.Nchains].ntiles;
            // every consumer has read its segment data: hand the buffer back before the reduction
            mbar_arrive(&sm.empty[b]);

            // ---------- deterministic block reduction over the 8 consumer warps ----------
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
            if (lane == 0) sm.red[warp] = v;
            consumer_sync();
            if (tid == 0) {
                double t = sm.red[0];
#pragma unroll
                for (int w = 1; w < NC / 32; w++) t += sm.red[w];
                A.partial[(size_t)sc * A.tiles_stride + tile] = t;
                __threadfence();
                const unsigned int ticket = atomicAdd(&A.counters[sc], 1u);
                sm.is_last = (ticket == (unsigned int)(ntiles - 1));
            }
            consumer_sync();
            if (sm.is_last) {
                // last CTA to finish a tile of this (star, chain): sum the per-tile partials in index order
                __threadfence();
                const volatile double* part = A.partial + (size_t)sc * A.tiles_stride;
                double acc = 0.0;
                for (int t = tid; t < ntiles; t += NC) acc += part[t];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
                consumer_sync();                          // sm.red reads above are complete
                if (lane == 0) sm.red[warp] = acc;
                consumer_sync();
                if (tid == 0) {
                    double S = sm.red[0];
#pragma unroll
                    for (int w = 1; w < NC / 32; w++) S += sm.red[w];
                    if (A.raw_sum) A.out[sc] = S;
                    else {
                        // likelihood_chi22p: f = -p*S with p truncated to long (model_def.cpp:399), / Tcoefs[m] (:401)
                        const double pl = (double)(long long)A.p;
                        A.out[sc] = (-pl * S) / A.Tcoefs[sc % A.Nchains];
                    }
                    A.counters[sc] = 0u;                  // ready for the next launch
                }
            }
            consumer_sync();                              // sm.red / sm.is_last reusable
        } else {
            mbar_arrive(&sm.empty[b]);
        }
        b ^= 1;
    }
}

template <bool WRITE_MODEL>
__global__ void __launch_bounds__(NT, TAMCMC_MIN_CTAS) tamcmc_whittle_kernel(WhittleArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&sm.full[0], 32); mbar_init(&sm.full[1], 32);
        mbar_init(&sm.empty[0], NC); mbar_init(&sm.empty[1], NC);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= NC) producer_loop(A, sm, tid - NC);
    else consumer_loop<WRITE_MODEL>(A, sm, tid);
}

__global__ void tamcmc_lnx_kernel(const double* __restrict__ x, double* __restrict__ lnx, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lnx[i] = log(x[i]);
}

// ---- DFMA roofline microbenchmark: 8 independent FMA chains per thread ----
__global__ void __launch_bounds__(256) tamcmc_dfma_kernel(double* out, int iters, double seed)
{
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace

cudaError_t tamcmc_whittle_configure(int* grid_ctas)
{
    const int smem = (int)sizeof(Smem);
    cudaError_t e = cudaFuncSetAttribute(tamcmc_whittle_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tamcmc_whittle_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tamcmc_whittle_kernel<false>, NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    *grid_ctas = sms * per_sm;      // persistent: one CTA per resident slot
    return cudaSuccess;
}

cudaError_t tamcmc_launch_whittle(const WhittleArgs& a, int grid_ctas, bool write_model, cudaStream_t st)
{
    const size_t smem = sizeof(Smem);
    if (write_model) tamcmc_whittle_kernel<true><<<grid_ctas, NT, smem, st>>>(a);
    else tamcmc_whittle_kernel<false><<<grid_ctas, NT, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t tamcmc_launch_lnx(const double* x, double* lnx, long long n, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    tamcmc_lnx_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, lnx, n);
    return cudaGetLastError();
}

cudaError_t tamcmc_fp64_peak(double* tflops, float* ms_out, int iters)
{
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const int blocks = sms * 8, threads = 256;
    double* d = nullptr;
    e = cudaMalloc(&d, sizeof(double) * (size_t)blocks * threads);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) tamcmc_dfma_kernel<<<blocks, threads>>>(d, iters, 1.0 + w);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        tamcmc_dfma_kernel<<<blocks, threads>>>(d, iters, 2.0 + r);
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return cudaSuccess;
}
