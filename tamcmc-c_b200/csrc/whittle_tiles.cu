// whittle_tiles.cu -- fused power-spectrum model + Whittle log-likelihood, ALL WARPS OF A CTA ON ONE TILE (sm_100a, FP64).
//
// Same inputs (the expander's ModeRec / CompRec / NoiseRec / TileRec tables and its heaviest-first work queue), same per-tile
// partial sums and same end-of-launch protocol as the producer / consumer ring of whittle.cu -- a different schedule.  With
// the far-field folding a tile of 1536 bins merges only ~10 components per bin, and what bounds the ring is
// the tile BUILD: one producer warp classifies the chain's modes and expands ~250 far components in series per tile, four
// such warps per SM (~3.5 us per tile and SM), while the FP64 pipe is ~20 % busy.  Here a CTA of 384 threads (12 warps, TWO
// CTAs resident per SM) owns a tile from start to end and spreads the build over all its threads:
//
//   0. persistent CTAs walk the work queue with a fixed stride (item b, b + grid, ...: the queue is sorted by cost, so every
//      CTA gets the same mix).  ONE ITEM AHEAD, warp 0 looks up the next item, reads its star / tile / noise records into a
//      shared-memory context and thread 0 starts TMA bulk copies (cp.async.bulk + mbarrier complete_tx) of the next chain's
//      mode headers and component records into the second table stage: no thread waits for a dependent global round trip
//      between two tiles -- the only global loads on the critical path are the tile's own x and y, requested first.
//   1. lists: 96 threads classify one mode each against the tile with the expander's bit-exact windows (window covers the
//      tile -> mask-free fast entries; window edge / extreme dynamic range -> general entries with [lo, hi); every component
//      >= far_ratio half-tiles away -> far entries), a block-wide scan places the entries, and every (mode, component) slot of
//      the staged tables is filed by its own thread.  More than 96 modes, or more entries than the lists hold: several passes.
//   2. far:   192 threads expand the far components in Taylor series about the tile centre (Chebyshev-U recurrence), one
//      shared-memory column each; a fixed-shape reduction adds the columns to the tile polynomial.
//   3. merge: every thread owns 4 bins and carries their Lorentzian sum as ONE fraction N/D (4 FP64 instructions per
//      component and bin, exponents renormalised with integer operations) -- the loops of whittle.cu.
//   4. epilogue: tile polynomial (background series + far field) or exact per-bin background, Whittle terms, fixed-shape
//      block reduction -> partial[sc][tile]; the last CTA of the launch finalises the chains (finish_launch).
// Results are bitwise reproducible run to run: nothing depends on which CTA runs where or when.
#include "whittle_shared.cuh"
#include <cstddef>

namespace {

constexpr int NC = TAMCMC_CONSUMERS;                       // threads per CTA (4 bins each for a full-size tile)
constexpr int BPT_MAX = TAMCMC_BINS_PER_THREAD;
constexpr int TILE_MAX = TAMCMC_TILE;
constexpr int GROUP = 16;                                  // fast components merged between two renormalisations
constexpr int GROUP_WIDE = 4;                              // ... in passes that hold WIDE-range components
constexpr int NB = TAMCMC_BG_TERMS;
constexpr int NFAR = TAMCMC_FAR_TERMS;
constexpr int TMB = 96;                                    // modes classified per pass (one thread each)
constexpr int SLOT_ROUNDS = 2;                             // (mode, component) slots per thread of a pass
constexpr int TCAPF = 640;                                 // fast entries (from the front) + far entries (from the back) per pass
constexpr int TCAPG = 64;                                  // general entries per pass
constexpr int FW = 192;                                    // threads that expand far components
constexpr int RED_LANES = 16;                              // threads that add up one coefficient of the far field
constexpr unsigned QENT_NONE = 0xffffffffu;                // "no such item" (a real entry never has all bits set: bit 31 is a flag)
static_assert(NC % 32 == 0 && TMB % 32 == 0 && TMB <= NC && FW <= NC && FW % RED_LANES == 0, "thread roles");
static_assert(NFAR * RED_LANES <= NC, "one group of 16 threads per coefficient");
static_assert(TMB * TAMCMC_MAX_COMP_PER_MODE <= SLOT_ROUNDS * NC, "every (mode, component) slot of a pass has a thread");
static_assert(TAMCMC_MAX_COMP_PER_MODE <= TCAPG && TAMCMC_MAX_COMP_PER_MODE <= TCAPF, "a single mode always fits a pass");
static_assert(sizeof(ModeRec) % 16 == 0 && sizeof(CompRec) % 16 == 0, "TMA bulk copies move multiples of 16 bytes");

#ifndef TAMCMC_TILES_MIN_CTAS
#define TAMCMC_TILES_MIN_CTAS 2                            // resident CTAs per SM the kernel is compiled for (register cap 65536 / (384 x this))
#endif

// per-mode record between the classification and the slot threads of a pass
struct __align__(16) ModeInfo {
    int of, og, ofar;        // first fast / general / far entry of the mode
    unsigned counts;         // ncomp | nfast << 8 | nfar << 16 | nfast_rec << 24; 0: the mode is not listed in this pass
    int lo, hi;              // window in tile-local bins, clamped to the tile
    int wbit, pad;
    double qa, qb, qc;       // asymmetry factor q(u) = (qa u + qb)^2 + qc
};

// one table stage: the TMA destination of a pass's mode headers and component records; once the pass's entries are filed the
// same bytes hold the far-field coefficient columns of that pass
union __align__(128) Stage {
    struct { ModeRec modes[TMB]; CompRec comps[TMB * TAMCMC_MAX_COMP_PER_MODE]; } raw;
    double facc[NFAR][FW];
};

// context of a work item as warp 0 stages it one item ahead: 8-byte words copied from the item's records
constexpr int CTX_TR = 0;                                  // TileRec (14 words)
constexpr int CTX_SD = 16;                                 // StarDesc words 0..5: off | Nloc, Nglob | bin0, tile0 | ntiles, tile_bins | model_id, Nparams | nmodes_cap, ..
constexpr int CTX_NZ = 22;                                 // NoiseRec words 0..1: nh, gauss | N0
constexpr int CTX_ASYM = 24, CTX_QENT = 25, CTX_QNEXT = 26;   // asym flag, the item's queue entry, the FOLLOWING item's queue entry
static_assert(sizeof(TileRec) == 8 * 14 && offsetof(TileRec, xc) == 80 && offsetof(TileRec, series_ok) == 96, "context layout");
static_assert(offsetof(StarDesc, off) == 0 && offsetof(StarDesc, Nloc) == 8 && offsetof(StarDesc, bin0) == 16 && offsetof(StarDesc, nmodes_cap) == 40, "context layout");
static_assert(offsetof(NoiseRec, gauss) == 4 && offsetof(NoiseRec, N0) == 8, "context layout");

struct __align__(128) TSmem {
    Stage st[2];
    double2 f_sc[TCAPF];                 // fast entries: e' = fma(u, s, c) ...
    double f_a[TCAPF];                   // ... t' = fma(e', e', a)
    ModeInfo minfo[TMB];
    ModeHdr hdr[TMB];                    // asymmetric profiles: q(u) of a mode and its run of fast entries
    GenEntry gen[TCAPG];
    double far_q[TMB][3];                // asymmetric profiles: q(u) = Q0 + Q1 u + Q2 u^2 of the pass's far modes
    unsigned char far_mode[TCAPF];       // asymmetric profiles: mode (index within the pass) of every far entry
    double poly[NFAR];                   // tile polynomial in u = x - xc: background series + far field
    double red_s[NC / 32], red_m[NC / 32];
    int red_e[NC / 32];
    unsigned wtot[TMB / 32][2];          // per-warp totals of the classification scan
    unsigned cut_tot[2];                 // totals of a pass that was cut short
    unsigned long long ctxraw[2][32];    // item contexts (this item / the next one)
    unsigned long long full[2];          // mbarriers of the two table stages
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
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
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(4000u) : "memory");
    }
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// TMA bulk copies of the mode headers and component records [base, base + TMB) of chain sc into a table stage (thread 0).
// The row of a chain always has modes_stride records, so the copy never depends on the star's own mode count.
__device__ __forceinline__ void issue_tables(const WhittleArgs& A, TSmem& sm, int stage, int sc, int base)
{
    const int nbc = min(TMB, A.modes_stride - base);
    const size_t m0 = (size_t)sc * A.modes_stride + base;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the stage was last written with ordinary stores (far-field columns)
    mbar_arrive_expect_tx(&sm.full[stage], (unsigned)nbc * (unsigned)(sizeof(ModeRec) + TAMCMC_MAX_COMP_PER_MODE * sizeof(CompRec)));
    tma_load_1d(sm.st[stage].raw.modes, A.modes + m0, (unsigned)nbc * (unsigned)sizeof(ModeRec), &sm.full[stage]);
    tma_load_1d(sm.st[stage].raw.comps, A.comps + m0 * TAMCMC_MAX_COMP_PER_MODE, (unsigned)nbc * (unsigned)(TAMCMC_MAX_COMP_PER_MODE * sizeof(CompRec)), &sm.full[stage]);
}

// queue entry of work item idx (called by a full warp; cum_inc: the lane's inclusive prefix of the cost-class counts)
__device__ __forceinline__ unsigned queue_entry(const WhittleArgs& A, unsigned idx, unsigned cum_inc, int lane)
{
    const int bucket = __popc(__ballot_sync(0xffffffffu, lane < TAMCMC_NBUCKETS && idx >= cum_inc));
    const unsigned cbase = __shfl_sync(0xffffffffu, cum_inc, (bucket + 31) & 31);      // prefix of the class before
    return A.queue[(size_t)bucket * A.qcap + (idx - (bucket ? cbase : 0u))];
}

// this lane's word of the context of the item with queue entry qent
__device__ __forceinline__ unsigned long long ctx_word(const WhittleArgs& A, unsigned qent, int lane)
{
    const unsigned item = qent & 0x7fffffffu;
    const int sc = (int)(item / (unsigned)A.tiles_stride);
    unsigned long long v = 0ull;
    if (lane < CTX_TR + 14) v = reinterpret_cast<const unsigned long long*>(A.tilerec + item)[lane - CTX_TR];
    else if (lane >= CTX_SD && lane < CTX_SD + 6) v = reinterpret_cast<const unsigned long long*>(A.stars + sc / A.Nchains)[lane - CTX_SD];
    else if (lane >= CTX_NZ && lane < CTX_NZ + 2) v = reinterpret_cast<const unsigned long long*>(A.noise + sc)[lane - CTX_NZ];
    else if (lane == CTX_ASYM) v = (unsigned long long)(unsigned)A.asym_flag[sc];
    else if (lane == CTX_QENT) v = qent;
    return v;
}

// Taylor coefficients in u of one far component 1 / ((s u + c)^2 + a), written to (FIRST) or added to this thread's column of
// the pass's coefficient table (stride FW doubles):
// f_0 = h, f_1 = p h, f_{k+1} = p f_k - q f_{k-1} with h = 1/(c^2 + a), p = -2 s c h, q = s^2 h (whittle.cu, far_series).
template <bool ASYM, bool FIRST>
__device__ __forceinline__ void far_series_col(double* col, double s, double c, double a, double Q0 = 1.0, double Q1 = 0.0, double Q2 = 0.0)
{
    const double w2 = fma(c, c, a);
    double h;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(h) : "d"(w2));
    h = fma(fma(-w2, h, 1.0), h, h);
    h = fma(fma(-w2, h, 1.0), h, h);          // two Newton steps: relative error ~1e-16
    const double sh = s * h;
    const double p = -2.0 * c * sh, q = s * sh;
    double f0 = h, f1 = p * h;
    {
        const double v0 = ASYM ? Q0 * f0 : f0, v1 = ASYM ? fma(Q0, f1, Q1 * f0) : f1;
        if (FIRST) { col[0] = v0; col[FW] = v1; } else { col[0] += v0; col[FW] += v1; }
    }
#pragma unroll
    for (int t = 2; t < NFAR; t++) {
        const double f2 = fma(p, f1, -(q * f0));
        const double v = ASYM ? fma(Q0, f2, fma(Q1, f1, Q2 * f0)) : f2;
        if (FIRST) col[FW * t] = v; else col[FW * t] += v;
        f0 = f1; f1 = f2;
    }
}

template <bool WRITE_MODEL, int BPT>
__global__ void __launch_bounds__(NC, TAMCMC_TILES_MIN_CTAS) tamcmc_whittle_tiles_kernel(WhittleArgs A)
{
    constexpr int TILE = NC * BPT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TSmem& sm = *reinterpret_cast<TSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&sm.full[0], 1); mbar_init(&sm.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // lane k < NBUCKETS keeps the inclusive prefix sum of the cost-class counts of the work queue
    static_assert(TAMCMC_NBUCKETS <= 32, "one lane per cost class");
    unsigned cum_inc = (lane < TAMCMC_NBUCKETS) ? A.qctl->count[lane] : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, cum_inc, d); if (lane >= d) cum_inc += o; }
    const unsigned ntot = __shfl_sync(0xffffffffu, cum_inc, 31);
    const unsigned G = gridDim.x;
    unsigned idx = blockIdx.x;

    if (idx < ntot) {
        // ---- prologue: the first item's context (the only look-up a CTA waits for) ----
        unsigned qent_ahead = QENT_NONE;         // warp 0: queue entry of the item after the next one (requested an item ahead)
        if (warp == 0) {
            const unsigned q0 = queue_entry(A, idx, cum_inc, lane);
            const unsigned q1 = (idx + G < ntot) ? queue_entry(A, idx + G, cum_inc, lane) : QENT_NONE;
            unsigned long long w = ctx_word(A, q0, lane);
            if (lane == CTX_QNEXT) w = q1;
            sm.ctxraw[0][lane] = w;
            if (idx + 2u * G < ntot) qent_ahead = queue_entry(A, idx + 2u * G, cum_inc, lane);
        }
        __syncthreads();                         // mbarriers initialised, context 0 in place
        unsigned ph0 = 0u, ph1 = 0u;             // parities of the two table stages
        unsigned unit = 0u;                      // passes done so far (stage = unit & 1)
        int pf_sc = -1, pf_base = 0;             // tables in flight (or landed) in the stage of the NEXT pass
        int cslot = 0;

        for (;; idx += G, cslot ^= 1) {
            // ---- the item ----
            const unsigned long long* cw = sm.ctxraw[cslot];
            const TileRec* tr = reinterpret_cast<const TileRec*>(cw + CTX_TR);
            const unsigned qent = (unsigned)cw[CTX_QENT], qnext = (unsigned)cw[CTX_QNEXT];
            const bool bg_only = (qent >> 31) != 0u;
            const unsigned item = qent & 0x7fffffffu;
            const int sc = (int)(item / (unsigned)A.tiles_stride);
            const int tile = (int)(item - (unsigned)sc * (unsigned)A.tiles_stride);
            const long long soff = (long long)cw[CTX_SD];
            const int Nloc = (int)(unsigned)cw[CTX_SD + 1], bin0 = (int)(unsigned)cw[CTX_SD + 2];
            const int nmodes = bg_only ? 0 : (int)(unsigned)cw[CTX_SD + 5];
            const double xc = tr->xc, umax = tr->umax;
            const int series_ok = tr->series_ok;
            const bool asym = cw[CTX_ASYM] != 0ull;
            const int lb0 = tile * TILE;
            const int nvalid = min(TILE, Nloc - lb0);
            const long long off = soff + lb0;
            const int g0 = bin0 + lb0, gend = g0 + nvalid;
            // the next item (if any): its chain, and whether it has mode tables to prefetch
            const bool have_next = qnext != QENT_NONE;
            const int next_sc = have_next ? (int)((qnext & 0x7fffffffu) / (unsigned)A.tiles_stride) : -1;
            const bool next_tables = have_next && !(qnext >> 31);

            // this thread's bins: b(j) = 2*tid + 2*NC*(j>>1) + (j&1), read as 128-bit pairs
            double u[BPT], N[BPT], D[BPT];
#pragma unroll
            for (int j = 0; j < BPT; j++) { N[j] = 0.0; D[j] = 1.0; u[j] = 0.0; }
            bool have_u = false;             // u still holds x
            bool requested = false;          // this item's global loads are under way
            unsigned long long w_next = 0ull;
            unsigned q_after = QENT_NONE;
            // The item's global loads.  x is requested here and first used by the merge loops (u = x - xc waits there); y is only needed
            // by the epilogue: its lines start their way into L2 and the registers are claimed there.  Warp 0, one item ahead: the
            // records of the next item (stored into the other context slot at the end of this item) and the queue entry of the item
            // after it -- none of these loads is waited for before this item's work is done.  All of them are issued AFTER thread 0
            // has started the table copies of the first pass (the proxy fence in front of a bulk copy waits for the thread's
            // outstanding loads: behind them it would cost a trip to DRAM on every tile).
            auto request_item_loads = [&]() {
                requested = true;
#pragma unroll
                for (int pj = 0; pj < BPT / 2; pj++) {
                    const double2 v = __ldg(reinterpret_cast<const double2*>(A.x + off + 2 * tid + 2 * NC * pj));
                    u[2 * pj] = v.x; u[2 * pj + 1] = v.y;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(A.y + off + 2 * tid + 2 * NC * pj));
                }
                if (warp == 0) {
                    if (have_next) w_next = ctx_word(A, qnext, lane);
                    q_after = qent_ahead;
                    qent_ahead = (idx + 3u * G < ntot) ? queue_entry(A, idx + 3u * G, cum_inc, lane) : QENT_NONE;
                }
            };

            // tile polynomial: background series; the far field is added per pass by the owner thread of each coefficient
            if (tid < NFAR * RED_LANES && (tid & (RED_LANES - 1)) == 0) {
                const int t = tid / RED_LANES;
                sm.poly[t] = (t < NB && series_ok) ? tr->bg[t] : 0.0;
            }
            const bool far_on = A.far_ratio > 0.0 && nmodes > 0;
            const double farR = A.far_ratio * umax;
            int any_far = 0;

            // ---- passes over the chain's modes ----
            for (int base = 0; base < nmodes;) {
                const int nb = min(TMB, nmodes - base);
                const int stg = (int)(unit & 1u);
                Stage& st = sm.st[stg];
                // the tables of this pass: prefetched during the previous pass / item, or (first pass of a CTA, a pass cut short, an item
                // behind a background-only tile) requested now
                if (!(pf_sc == sc && pf_base == base)) {
                    if (pf_sc >= 0) { if (stg) { mbar_wait(&sm.full[1], ph1); ph1 ^= 1u; } else { mbar_wait(&sm.full[0], ph0); ph0 ^= 1u; } }      // not these tables: let them land
                    __syncthreads();
                    if (tid == 0) issue_tables(A, sm, stg, sc, base);
                }
                // what comes after this pass goes into the other stage now (everybody left that stage at the end of the previous pass)
                {
                    int n_sc = -1, n_base = 0;
                    if (base + nb < nmodes) { n_sc = sc; n_base = base + nb; }
                    else if (next_tables) { n_sc = next_sc; n_base = 0; }
                    if (tid == 0 && n_sc >= 0) issue_tables(A, sm, stg ^ 1, n_sc, n_base);
                    pf_sc = n_sc; pf_base = n_base;
                }
                if (!requested) request_item_loads();
                if (stg) { mbar_wait(&sm.full[1], ph1); ph1 ^= 1u; } else { mbar_wait(&sm.full[0], ph0); ph0 ^= 1u; }
                unit++;

                const ModeRec* modes = st.raw.modes;          // records [base, base + nb) of the chain
                int ncomp = 0, nfast = 0, ngen = 0, i0 = 0, i1 = 0, nfast_rec = 0, wbit = 0, nfar = 0, hh = 0, mwide = 0;
                double qa = 0.0, qb = 1.0, qc = 0.0;
                if (tid < nb) {
                    const int4 h = *reinterpret_cast<const int4*>(modes + tid);     // {i0, i1, ncomp, nfast | wide << 16}
                    if (h.z > 0 && h.x < gend && h.y > g0) {
                        ncomp = h.z; nfast_rec = h.w & 0xffff; i0 = h.x; i1 = h.y;
                        nfast = (h.x <= g0 && h.y >= gend) ? nfast_rec : 0;
                        ngen = ncomp - nfast;
                        wbit = (h.w >> 16) & 1;
                        if (far_on && nfast > 0) {
                            const double2 fz = *reinterpret_cast<const double2*>(&modes[tid].numin);
                            if (fz.x <= fz.y && ((fz.x - xc) >= farR || (xc - fz.y) >= farR)) { nfar = nfast; nfast = 0; }
                        }
                        mwide = (nfast > 0) ? wbit : 0;
                        hh = (asym && nfast > 0) ? 1 : 0;
                        if (asym || ngen > 0) { qa = modes[tid].qa; qb = modes[tid].qb0 + xc * qa; qc = modes[tid].qc; }
                    }
                }
                // block-wide inclusive scan of the per-mode counts (two packed words; every total <= 96 x 7 = 672 < 2^16)
                unsigned w0 = (unsigned)nfast | ((unsigned)ngen << 16);
                unsigned w1 = (unsigned)nfar | ((unsigned)hh << 16) | ((unsigned)mwide << 24);
                if (warp < TMB / 32) {
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const unsigned o0 = __shfl_up_sync(0xffffffffu, w0, d), o1 = __shfl_up_sync(0xffffffffu, w1, d);
                        if (lane >= d) { w0 += o0; w1 += o1; }
                    }
                    if (lane == 31) { sm.wtot[warp][0] = w0; sm.wtot[warp][1] = w1; }
                }
                if (!__syncthreads_or(ncomp > 0)) { base += nb; continue; }       // no mode of this pass touches the tile
                unsigned tot0 = 0u, tot1 = 0u;                                       // totals of the whole pass
#pragma unroll
                for (int w = 0; w < TMB / 32; w++) {
                    const unsigned a0 = sm.wtot[w][0], a1 = sm.wtot[w][1];
                    if (w < warp) { w0 += a0; w1 += a1; }
                    tot0 += a0; tot1 += a1;
                }
                const unsigned e0 = w0 - ((unsigned)nfast | ((unsigned)ngen << 16));      // exclusive prefixes of this thread's mode
                const unsigned e1 = w1 - ((unsigned)nfar | ((unsigned)hh << 16) | ((unsigned)mwide << 24));
                int cut = nb;
                if ((tot0 & 0xffffu) + (tot1 & 0xffffu) > (unsigned)TCAPF || (tot0 >> 16) > (unsigned)TCAPG) {
                    // the lists do not hold the whole pass: keep the longest run of leading modes that fits (a single mode always does)
                    const bool fits = tid < nb && (w0 & 0xffffu) + (w1 & 0xffffu) <= (unsigned)TCAPF && (w0 >> 16) <= (unsigned)TCAPG;
                    cut = __syncthreads_count(fits);
                    if (tid == cut - 1) { sm.cut_tot[0] = w0; sm.cut_tot[1] = w1; }
                    __syncthreads();
                    tot0 = sm.cut_tot[0]; tot1 = sm.cut_tot[1];
                }
                const int tf = (int)(tot0 & 0xffffu), tg = (int)(tot0 >> 16), tfar = (int)(tot1 & 0xffffu), th = (int)((tot1 >> 16) & 0xffu);
                const bool seg_wide = (tot1 >> 24) != 0u;
                if (tid < nb) {
                    // what the slot threads need to place this mode's components
                    ModeInfo m;
                    const bool listed = tid < cut && ncomp > 0;
                    m.of = (int)(e0 & 0xffffu); m.og = (int)(e0 >> 16); m.ofar = (int)(e1 & 0xffffu);
                    m.lo = max(i0 - g0, 0); m.hi = min(i1 - g0, TILE);
                    m.counts = listed ? ((unsigned)ncomp | ((unsigned)nfast << 8) | ((unsigned)nfar << 16) | ((unsigned)nfast_rec << 24)) : 0u;
                    m.wbit = wbit; m.pad = 0;
                    m.qa = qa; m.qb = qb; m.qc = qc;
                    sm.minfo[tid] = m;
                    if (listed && hh) { ModeHdr mh; mh.qa = qa; mh.qb = qb; mh.qc = qc; mh.begin = m.of; mh.count = nfast; sm.hdr[(e1 >> 16) & 0xffu] = mh; }
                    if (listed && asym && nfar > 0) { sm.far_q[tid][0] = fma(qb, qb, qc); sm.far_q[tid][1] = 2.0 * qa * qb; sm.far_q[tid][2] = qa * qa; }
                }
                __syncthreads();
                // every slot thread files its component: fast entry, far entry (from the back of the fast arrays) or general entry
#pragma unroll
                for (int r = 0; r < SLOT_ROUNDS; r++) {
                    const int sl = tid + NC * r;
                    if (sl < nb * TAMCMC_MAX_COMP_PER_MODE) {
                        const int ml = sl / TAMCMC_MAX_COMP_PER_MODE, k = sl - ml * TAMCMC_MAX_COMP_PER_MODE;
                        const unsigned cnt = sm.minfo[ml].counts;
                        const int m_ncomp = (int)(cnt & 0xffu), m_nfast = (int)((cnt >> 8) & 0xffu), m_nfar = (int)((cnt >> 16) & 0xffu), m_nfast_rec = (int)(cnt >> 24);
                        if (k < m_ncomp) {
                            const ModeInfo& m = sm.minfo[ml];
                            const CompRec& cr = st.raw.comps[sl];
                            const double cs = cr.s, ca = cr.a;
                            const double cc = -(cr.nu - xc) * cs;
                            if (k < m_nfar) {
                                const int fe = TCAPF - 1 - (m.ofar + k);
                                sm.f_sc[fe] = make_double2(cs, cc); sm.f_a[fe] = ca;
                                if (asym) sm.far_mode[fe] = (unsigned char)ml;
                            } else if (k < m_nfast) { sm.f_sc[m.of + k] = make_double2(cs, cc); sm.f_a[m.of + k] = ca; }
                            else {
                                // components are stored FAST-first; a FAST one lands here only on a window edge
                                const bool ff = k < m_nfast_rec;
                                GenEntry ge;
                                ge.s = cs; ge.c = cc; ge.aadd = ff ? ca : 1.0; ge.num = ff ? 1.0 : ca;
                                // window in tile-local bins, clamped to the tile; bit 30 of hi: the entry needs an exponent
                                // renormalisation after every merge (general form, or a WIDE-range mode)
                                ge.qa = m.qa; ge.qb = m.qb; ge.qc = m.qc; ge.lo = m.lo;
                                ge.hi = m.hi | ((!ff || m.wbit) ? (1 << 30) : 0);
                                sm.gen[m.og + (k - (m_nfast + m_nfar))] = ge;
                            }
                        }
                    }
                }
                __syncthreads();                       // the lists of this pass are complete; the staged tables are dead from here
                double (*facc)[FW] = st.facc;          // ... and hold the far-field columns of the pass
                // ---- 3. far field of this pass ----
                if (tfar > 0) {
                    any_far = 1;
                    if (tid < FW) {
                        // column `tid` of the coefficient table: the thread's first component writes it, further ones (a pass with more
                        // than 192 far components) add to it; columns >= tfar are never read
                        double* col = &facc[0][tid];
                        for (int e = tid; e < tfar; e += FW) {
                            const int fe = TCAPF - 1 - e;
                            const double2 sc_ = sm.f_sc[fe];
                            const double ca_ = sm.f_a[fe];
                            if (!asym) {
                                if (e == tid) far_series_col<false, true>(col, sc_.x, sc_.y, ca_); else far_series_col<false, false>(col, sc_.x, sc_.y, ca_);
                            } else {
                                const double* Q = sm.far_q[sm.far_mode[fe]];
                                if (e == tid) far_series_col<true, true>(col, sc_.x, sc_.y, ca_, Q[0], Q[1], Q[2]);
                                else far_series_col<true, false>(col, sc_.x, sc_.y, ca_, Q[0], Q[1], Q[2]);
                            }
                        }
                    }
                    __syncthreads();
                    if (tid < NFAR * RED_LANES) {
                        // coefficient t: 16 threads add the columns p, p + 16, ... in order, then a shuffle tree (fixed shape)
                        const int t = tid / RED_LANES, p = tid & (RED_LANES - 1);
                        const int ncol = min(tfar, FW);
                        double s = 0.0;
                        for (int col = p; col < ncol; col += RED_LANES) s += facc[t][col];
    #pragma unroll
                        for (int d = RED_LANES / 2; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                        if (p == 0) sm.poly[t] += s;
                    }
                }

                // ---- merge this pass's components into the bins' fractions ----
                if (!have_u) {
                    have_u = true;
    #pragma unroll
                    for (int j = 0; j < BPT; j++) u[j] -= xc;
                }
                if (!asym) {
                    int k = 0;
                    if (!seg_wide) {
                        for (; k + GROUP <= tf; k += GROUP) {
    #pragma unroll
                            for (int kk = 0; kk < GROUP; kk++) {
                                const double2 p = sm.f_sc[k + kk];
                                const double a = sm.f_a[k + kk];
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
    #pragma unroll 4
                        for (; k < tf; k++) {
                            const double2 p = sm.f_sc[k];
                            const double a = sm.f_a[k];
    #pragma unroll
                            for (int j = 0; j < BPT; j++) {
                                const double e = fma(u[j], p.x, p.y);
                                const double t = fma(e, e, a);
                                N[j] = fma(N[j], t, D[j]);
                                D[j] *= t;
                            }
                        }
                    } else {
                        // WIDE-range components in this pass: renormalise every GROUP_WIDE merges
                        for (; k < tf; k += GROUP_WIDE) {
    #pragma unroll
                            for (int kk = 0; kk < GROUP_WIDE; kk++) {
                                if (k + kk < tf) {
                                    const double2 p = sm.f_sc[k + kk];
                                    const double a = sm.f_a[k + kk];
    #pragma unroll
                                    for (int j = 0; j < BPT; j++) {
                                        const double e = fma(u[j], p.x, p.y);
                                        const double t = fma(e, e, a);
                                        N[j] = fma(N[j], t, D[j]);
                                        D[j] *= t;
                                    }
                                }
                            }
    #pragma unroll
                            for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
                        }
                    }
    #pragma unroll
                    for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
                } else {
                    // asymmetric Lorentzians (build_lorentzian.cpp:153-157): every component of a mode is
                    // multiplied by q(x) = (1 + asym (x/fc - 1))^2 + (Gamma asym / (2 fc))^2
                    int since = 0;
                    for (int m = 0; m < th; m++) {
                        const ModeHdr h = sm.hdr[m];
                        double q[BPT];
    #pragma unroll
                        for (int j = 0; j < BPT; j++) { const double w = fma(u[j], h.qa, h.qb); q[j] = fma(w, w, h.qc); }
                        for (int k = h.begin; k < h.begin + h.count; k++) {
                            const double2 p = sm.f_sc[k];
                            const double a = sm.f_a[k];
    #pragma unroll
                            for (int j = 0; j < BPT; j++) {
                                const double e = fma(u[j], p.x, p.y);
                                const double t = fma(e, e, a);
                                N[j] = fma(N[j], t, q[j] * D[j]);
                                D[j] *= t;
                            }
                        }
                        since += h.count;
                        if (since + TAMCMC_MAX_COMP_PER_MODE > (seg_wide ? TAMCMC_MAX_COMP_PER_MODE : GROUP)) {
                            since = 0;
    #pragma unroll
                            for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
                        }
                    }
    #pragma unroll
                    for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
                }
                // general entries: window edges and extreme-dynamic-range components.  Every warp owns a contiguous run of 64 bins per
                // register pair (bins 2*tid + 2*NC*pj + r): a run the window covers is merged unmasked, a run holding a window edge
                // under a per-bin mask, the others are skipped; the three-way decision is warp-uniform.
                {
                    int gsince[BPT / 2];                       // plain merges of each register pair since its last renormalisation
    #pragma unroll
                    for (int pj = 0; pj < BPT / 2; pj++) gsince[pj] = 0;
                    for (int g = 0; g < tg; g++) {
                        const GenEntry ge = sm.gen[g];
                        const int ghi = ge.hi & 0xffff;
                        const bool heavy = (ge.hi >> 30) & 1;
    #pragma unroll
                        for (int pj = 0; pj < BPT / 2; pj++) {
                            const int r0 = 2 * NC * pj + 64 * warp, r1 = r0 + 64;
                            if (ghi <= r0 || ge.lo >= r1) continue;
                            const bool whole = (ge.lo <= r0 && ghi >= r1);
                            if (whole && !heavy) {
    #pragma unroll
                                for (int r = 0; r < 2; r++) {
                                    const int j = 2 * pj + r;
                                    const double e = fma(u[j], ge.s, ge.c);
                                    const double t = fma(e, e, ge.aadd);
                                    double nd = D[j];                       // FAST form: num == 1
                                    if (asym) { const double w = fma(u[j], ge.qa, ge.qb); nd *= fma(w, w, ge.qc); }
                                    N[j] = fma(N[j], t, nd);
                                    D[j] *= t;
                                }
                                if (++gsince[pj] >= GROUP - 1) { gsince[pj] = 0; renorm(N[2 * pj], D[2 * pj]); renorm(N[2 * pj + 1], D[2 * pj + 1]); }
                            } else {
    #pragma unroll
                                for (int r = 0; r < 2; r++) {
                                    const int j = 2 * pj + r;
                                    const int bb = 2 * tid + 2 * NC * pj + r;
                                    const bool in = whole || ((bb >= ge.lo) && (bb < ghi));
                                    const double e = fma(u[j], ge.s, ge.c);
                                    const double t = fma(e, e, ge.aadd);
                                    double nd = ge.num * D[j];
                                    if (asym) { const double w = fma(u[j], ge.qa, ge.qb); nd *= fma(w, w, ge.qc); }
                                    const double Nn = fma(N[j], t, nd);
                                    const double Dn = D[j] * t;
                                    if (in) { N[j] = Nn; D[j] = Dn; }
                                    renorm(N[j], D[j]);
                                }
                                gsince[pj] = 0;
                            }
                        }
                    }
                    if (tg) {
    #pragma unroll
                        for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
                    }
                }
                base += cut;
                __syncthreads();      // lists, stage and (last pass) tile polynomial: everybody is done with / can see them
            }

            // ---- epilogue ----
            if (!requested) request_item_loads();
            if (!have_u) {
#pragma unroll
                for (int j = 0; j < BPT; j++) u[j] -= xc;
            }
            if (nmodes == 0) __syncthreads();            // (no pass ran: the tile polynomial written above becomes visible here)
            double yv[BPT];
#pragma unroll
            for (int pj = 0; pj < BPT / 2; pj++) {
                const double2 w = __ldg(reinterpret_cast<const double2*>(A.y + off + 2 * tid + 2 * NC * pj));
                yv[2 * pj] = w.x; yv[2 * pj + 1] = w.y;
            }
            // what only the epilogue needs comes from the staged context here, not from registers held across the passes
            const double N0 = __longlong_as_double((long long)cw[CTX_NZ + 1]);
            const bool gauss = (unsigned)(cw[CTX_NZ] >> 32) != 0u;
            const NoiseRec* nz = A.noise + sc;
            double bgv[BPT];
            if (any_far) {
                double acc[BPT];
    #pragma unroll
                for (int j = 0; j < BPT; j++) acc[j] = sm.poly[NFAR - 1];
    #pragma unroll
                for (int k = NFAR - 2; k >= 0; k--) {
                    const double ck = sm.poly[k];
    #pragma unroll
                    for (int j = 0; j < BPT; j++) acc[j] = fma(acc[j], u[j], ck);
                }
    #pragma unroll
                for (int j = 0; j < BPT; j++) bgv[j] = acc[j] + N0;
            } else if (series_ok) {
                double acc[BPT];
    #pragma unroll
                for (int j = 0; j < BPT; j++) acc[j] = sm.poly[NB - 1];
    #pragma unroll
                for (int k = NB - 2; k >= 0; k--) {
                    const double ck = sm.poly[k];
    #pragma unroll
                    for (int j = 0; j < BPT; j++) acc[j] = fma(acc[j], u[j], ck);
                }
    #pragma unroll
                for (int j = 0; j < BPT; j++) bgv[j] = acc[j] + N0;
            } else {
    #pragma unroll
                for (int j = 0; j < BPT; j++) bgv[j] = N0;
            }
            // (with series_ok == 0 the polynomial holds the far field alone: the exact background terms follow)
            if (!series_ok) {
                // near x = 0 / near a singularity of a term: evaluate every bin exactly and merge the terms into
                // the same fraction: (1e-3 tau x)^p = exp(p (ln(1e-3 tau) + ln x)); the clamp keeps D finite
                const double* lx = A.lnx + off;
                const int nh = nz->nh;
                double lnx[BPT];
    #pragma unroll
                for (int pj = 0; pj < BPT / 2; pj++) {
                    const double2 v = *reinterpret_cast<const double2*>(lx + 2 * tid + 2 * NC * pj);
                    lnx[2 * pj] = v.x; lnx[2 * pj + 1] = v.y;
                }
                for (int h = 0; h < nh; h++) {
                    const double H = nz->H[h], ls = nz->lnsc[h], pw = nz->pw[h], isc = nz->isc[h];
                    const bool p4 = (pw == 4.0), p2 = (pw == 2.0);     // the usual fixed slopes: plain products instead of exp
    #pragma unroll
                    for (int j = 0; j < BPT; j++) {
                        double z;
                        if (p4 || p2) {
                            const double r = (u[j] + xc) * isc, r2 = r * r;
                            z = fmin(p4 ? r2 * r2 : r2, 2.5e30);
                        } else {
                            const double arg = fmin(pw * (ls + lnx[j]), 70.0);
                            z = (pw == 0.0) ? 1.0 : exp(arg);
                        }
                        const double t = 1.0 + z;
                        N[j] = fma(N[j], t, H * D[j]);
                        D[j] *= t;
                    }
                }
            }
            if (gauss) {
    #pragma unroll
                for (int j = 0; j < BPT; j++) bgv[j] += gauss_envelope(nz, u[j] + xc);
            }

            // M = N/D + background; Whittle terms.  y_i/M_i is summed; ln M_i is carried as the product of the 1/M_i split
            // exactly into mantissa and integer exponent (the per-chain finalisation takes one log per tile).
            double s1 = 0.0, pm = 1.0;
            int pe = 0;
            int bad = 0;                     // sign / NaN watch of 1/M_i: OR of the high words
            if (A.likelihood == 1) {
                // chi_square (likelihoods.cpp:31-40): sum of (y - M)^2 / sigma_y^2; the weights 1/sigma^2 were formed once at create
                const double* wg = A.wsig + off;
    #pragma unroll
                for (int pj = 0; pj < BPT / 2; pj++) {
                    const double2 w = *reinterpret_cast<const double2*>(wg + 2 * tid + 2 * NC * pj);
    #pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const int j = 2 * pj + r;
                        const int bb = 2 * tid + 2 * NC * pj + r;
                        const double num = fma(bgv[j], D[j], N[j]);
                        const double M = num / D[j];
                        if (WRITE_MODEL) { if (bb < nvalid) A.model_out[lb0 + bb] = M; }
                        if (bb < nvalid) { const double d = yv[j] - M; s1 = fma(d * d, r ? w.y : w.x, s1); }
                    }
                }
            } else {
    #pragma unroll
                for (int j = 0; j < BPT; j++) {
                    const int bb = 2 * tid + 2 * NC * (j >> 1) + (j & 1);
                    const double num = fma(bgv[j], D[j], N[j]);
                    if (WRITE_MODEL) { if (bb < nvalid) A.model_out[lb0 + bb] = num / D[j]; }
                    if (bb < nvalid) {
                        const double minv = D[j] * fast_rcp(num); // 1/M_i
                        s1 = fma(yv[j], minv, s1);                // y_i / M_i
                        const int hi = __double2hiint(minv);
                        const int k = (hi & 0x7ff00000) - 0x3ff00000;
                        pm *= __hiloint2double(hi - k, __double2loint(minv));
                        pe += k >> 20;
                        bad |= hi;
                    }
                }
            }
            // one non-positive model bin makes the reference's logL NaN (likelihoods.cpp:23, MALA.cpp:522): the sign bit of any
            // 1/M_i poisons the tile (whittle.cu, consumer_loop)
            if (bad < 0) pm = nan("");
            // fixed-shape block reduction: shuffle tree per warp (<= 4 mantissas in [1,2) per thread: 2^128 at most), mantissa
            // product back into [1,2), then the 12 warps in warp order
    #pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                s1 += __shfl_down_sync(0xffffffffu, s1, d);
                pm *= __shfl_down_sync(0xffffffffu, pm, d);
                pe += __shfl_down_sync(0xffffffffu, pe, d);
            }
            if (lane == 0) {
                const int hi = __double2hiint(pm);
                const int k = (hi & 0x7ff00000) - 0x3ff00000;
                sm.red_s[warp] = s1;
                sm.red_m[warp] = (pm == pm) ? __hiloint2double(hi - k, __double2loint(pm)) : pm;
                sm.red_e[warp] = pe + (k >> 20);
            }
            // warp 0 files the next item's context (its loads were issued when this item began)
            if (warp == 0 && have_next) { if (lane == CTX_QNEXT) w_next = q_after; sm.ctxraw[cslot ^ 1][lane] = w_next; }
            __syncthreads();
            if (tid == 0) {
                double S = 0.0, Pm = 1.0;
                int Pe = 0;
#pragma unroll
                for (int w = 0; w < NC / 32; w++) { S += sm.red_s[w]; Pm *= sm.red_m[w]; Pe += sm.red_e[w]; }
                const int hi = __double2hiint(Pm);
                const int k = (hi & 0x7ff00000) - 0x3ff00000;
                double* part = A.partial + 3 * ((size_t)sc * A.tiles_stride + tile);
                part[0] = S;
                part[1] = (Pm == Pm) ? __hiloint2double(hi - k, __double2loint(Pm)) : Pm;
                part[2] = (double)(Pe + (k >> 20));
            }
            if (!have_next) break;
        }
    }
    finish_launch(A, tid, NC);
}

}  // namespace

static int g_tiles_per_sm = 0, g_tiles_sms = 0;

template <bool WM, int BPT>
static cudaError_t tiles_configure_one(int* per_sm_out)
{
    const int smem = (int)sizeof(TSmem);
    cudaError_t e = cudaFuncSetAttribute(tamcmc_whittle_tiles_kernel<WM, BPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tamcmc_whittle_tiles_kernel<WM, BPT>, NC, smem);
    if (e != cudaSuccess) return e;
    if (per_sm_out) *per_sm_out = per_sm;
    return cudaSuccess;
}

cudaError_t tamcmc_whittle_tiles_configure(int* ctas_per_sm)
{
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    if ((e = tiles_configure_one<false, BPT_MAX>(&per_sm)) != cudaSuccess) return e;
    if ((e = tiles_configure_one<true, BPT_MAX>(nullptr)) != cudaSuccess) return e;
    if ((e = tiles_configure_one<false, BPT_MAX / 2>(nullptr)) != cudaSuccess) return e;
    if ((e = tiles_configure_one<true, BPT_MAX / 2>(nullptr)) != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    g_tiles_per_sm = per_sm; g_tiles_sms = sms;
    if (ctas_per_sm) *ctas_per_sm = per_sm;
    return cudaSuccess;
}

// persistent grid: one CTA per resident slot, never more CTAs than the launch can have work items
cudaError_t tamcmc_launch_whittle_tiles(const WhittleArgs& a, unsigned nitems_max, bool write_model, int tile_bins, cudaStream_t st)
{
    unsigned grid = (unsigned)(g_tiles_per_sm > 0 ? g_tiles_per_sm * g_tiles_sms : 148);
    if (nitems_max < grid) grid = nitems_max;
    if (grid == 0u) grid = 1u;                   // the end-of-launch protocol needs one CTA
    const size_t smem = sizeof(TSmem);
    if (tile_bins == TILE_MAX) {
        if (write_model) tamcmc_whittle_tiles_kernel<true, BPT_MAX><<<grid, NC, smem, st>>>(a);
        else tamcmc_whittle_tiles_kernel<false, BPT_MAX><<<grid, NC, smem, st>>>(a);
    } else if (tile_bins == TILE_MAX / 2) {
        if (write_model) tamcmc_whittle_tiles_kernel<true, BPT_MAX / 2><<<grid, NC, smem, st>>>(a);
        else tamcmc_whittle_tiles_kernel<false, BPT_MAX / 2><<<grid, NC, smem, st>>>(a);
    } else return cudaErrorInvalidValue;
    return cudaGetLastError();
}
