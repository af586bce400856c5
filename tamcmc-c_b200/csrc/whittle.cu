// whittle.cu -- fused power-spectrum model + Whittle chi^2(2 dof) log-likelihood (sm_100a, FP64).
//
// Persistent, warp-specialised kernel.  Each CTA has 4 PRODUCER warps and 12 CONSUMER warps:
//
//  * every producer warp owns ONE slot of the shared-memory ring.  It pops a (star, chain, tile) work item from the
//    heaviest-first queue the expander built, classifies the chain's modes against the tile (bit-exact windows from
//    the expander) and writes the tile-local component list straight into its slot: window covers the tile ->
//    mask-free FastEntry, window edge or extreme dynamic range -> GenEntry with [lo,hi); it fetches the tile's x and
//    y (2 x 12 KB) with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) and publishes the slot on its
//    `full` mbarrier.  Four producers overlap the global-memory latency of four tiles; a list that exceeds the
//    slot's capacities is streamed through the same slot as several segments.  FAR-FIELD FOLDING: a mode whose components
//    all lie >= far_ratio tile half-widths from the tile centre (and whose window covers the tile) is not listed: the
//    producer expands its components in Taylor series about the tile centre (Chebyshev-U recurrence, far_series) and adds
//    the coefficients to the tile's background polynomial, so the consumers pay one longer Horner evaluation per bin
//    instead of 4 FP64 instructions per far component and bin (truncation <= 5e-13 of a folded component);
//  * the 384 consumer threads own 4 bins each and take the slots round-robin, one tile at a time.  The model
//    spectrum M never exists in HBM: every thread keeps the running Lorentzian sum of a bin as ONE fraction N/D,
//        sum_k A_k / (1 + 4 (x - nu_k)^2 / Gamma_k^2)  =  N / D ,
//    merging a component with 2 FMAs for its scaled denominator t' = (1 + e^2)/A_k and 1 FMA + 1 MUL for
//    (N, D) <- (N t' + D, D t').  That replaces the reference's FP64 divide per (component, bin)
//    (build_lorentzian.cpp:151: cwiseInverse) by 4 FP64-pipe instructions, with one reciprocal per bin at the
//    end.  Exponents of (N, D) are renormalised with 4 integer ops per bin every 16 components.  The background
//    (noise_models.cpp:15-39) is a 9th-degree polynomial per tile (or exact exp() per bin near x = 0; 19th degree once far
//    modes are folded into it), the
//    Whittle terms y/M + ln M (likelihoods.cpp:23) are summed per thread; the thread stores its three sums in the
//    slot's scratch and releases the slot, and the slot's PRODUCER warp reduces the 384 triples (fixed shape) into
//    one partial per tile once the empty barrier has handed the slot back.  The last CTA to finish sums the per-tile
//    partials of every chain in tile order, so results are bitwise reproducible run to run whatever the scheduling.
//    The Gaussian-envelope models (ids 0, 1: no Lorentzians) ride the same path with empty component lists.
#include "whittle_shared.cuh"

namespace {

constexpr int NC = TAMCMC_CONSUMERS;                       // consumer threads
constexpr int NT = TAMCMC_THREADS;                         // + producer warp
constexpr int BPT_MAX = TAMCMC_BINS_PER_THREAD;            // bins per consumer thread of a full-size tile
constexpr int TILE_MAX = TAMCMC_TILE;                      // the kernel is instantiated for tiles of TILE_MAX and TILE_MAX / 2 bins
                                                           // (BPT = 4 and 2): small spectra use the smaller tile so that more SMs get work
constexpr int GROUP = 16;                                  // fast components merged between two renormalisations
constexpr int GROUP_WIDE = 4;                              // ... in segments that hold WIDE-range components
constexpr int NB = TAMCMC_BG_TERMS;
constexpr int NFAR = TAMCMC_FAR_TERMS;                     // terms of the tile polynomial when far Lorentzians are folded into it
static_assert(NFAR >= NB && NFAR <= 32, "one producer lane per coefficient");
constexpr int CAPF = TAMCMC_CAPF;                          // fast entries per segment
constexpr int CAPG = TAMCMC_CAPG;                          // general entries per segment
constexpr int CAPH = TAMCMC_CAPH;                          // mode headers per segment (asym fast path)
constexpr int NBUF = TAMCMC_PRODUCERS;                     // ring slots = producer warps
constexpr int NPROD = NBUF;
#ifdef TAMCMC_WARP_ARRIVE
constexpr int EMPTY_COUNT = NC / 32;                       // one arrival per consumer warp (lane 0): measured 1.3 % SLOWER on C2
#else
constexpr int EMPTY_COUNT = NC;                            // every consumer thread arrives on the slot's `empty` barrier
#endif

enum { SEG_FIRST = 1, SEG_LAST = 2, SEG_DONE = 4, SEG_BGSERIES = 8, SEG_ASYM = 16, SEG_GAUSS = 32, SEG_WIDE = 64, SEG_FAR = 128 };

template <int TILE>
struct __align__(16) Segment {
    double x[TILE];          // TMA destinations: spectrum tile (first segment of a tile) ...
    double y[TILE];
    FastEntry fast[CAPF];    // ... and the tile's component lists, written by the slot's producer warp
    ModeHdr hdr[CAPH];
    GenEntry gen[CAPG];
    double bg[NFAR];         // tile polynomial in u = x - xc: Harvey background series (NB terms) + far Lorentzians (NFAR terms)
    double xc, N0;
    int nfast, ngen, nhdr, flags;
    int sc_index, tile, nvalid, lb0;
    long long off;           // offset of the tile in the concatenated arrays
};

template <int TILE>
struct Smem {
    Segment<TILE> seg[NBUF];
    unsigned long long full[NBUF], empty[NBUF];
    // per-thread Whittle sums of the slot's tile, left by the consumers for the slot's producer warp (tile_reduce)
    double red_s[NBUF][NC];
    double red_m[NBUF][NC];
    int red_e[NBUF][NC];
    double far_acc[NBUF][NFAR][33];   // per-lane partial sums of the far-field coefficients of the tile a producer is listing
    int far_list[NBUF][32 * TAMCMC_MAX_COMP_PER_MODE];   // far components of the mode batch being listed (indices into the chain's CompRec table)
    double far_q[NBUF][32][3];        // asymmetric profiles: q(u) = Q0 + Q1 u + Q2 u^2 of the batch's far modes (one per lane)
    volatile int cons_cur;            // slot the consumers are working on (-1 before the first tile)
    unsigned int done_mask;           // producers that have drained the queue
};

static_assert(sizeof(Smem<TILE_MAX>) <= 232448, "the ring must fit the 227 KB of shared memory a CTA can opt into");

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
// Blocks in hardware for up to `MBAR_SUSPEND_NS` per try (suspendTimeHint): a waiting warp issues one instruction group
// every few microseconds instead of spinning and taking issue slots from the FP64 loops of the consumer warps.
constexpr unsigned MBAR_SUSPEND_NS = 4000;
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok = 0;
    const unsigned addr = smem_u32(bar);
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(MBAR_SUSPEND_NS) : "memory");
    }
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#ifdef TAMCMC_TRACE
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TRACE(slot, val) do { if (A.trace && (slot) < 64) A.trace[(size_t)blockIdx.x * 64 + (slot)] = (val); } while (0)
// per-phase cycle accounting of consumer warp 0 (slots 48..57 of the CTA's trace row)
#ifndef TAMCMC_TRACE_TID
#define TAMCMC_TRACE_TID 0      // consumer thread whose phases are accounted (e.g. 256 = warp 8)
#endif
// -DTAMCMC_TRACE_GANTT: lane 0 of EVERY consumer warp of CTA 0 also logs (clock64 << 4 | phase) at each phase boundary
// (trace words 65536 + 512 warp + event): profiles/trace_gantt.py draws which warps overlap which phases
#ifdef TAMCMC_TRACE_GANTT
#define GANTT(k) do { if (blockIdx.x == 0 && (tid & 31) == 0 && A.trace && gantt_n < 512) A.trace[65536 + 512 * (tid >> 5) + gantt_n++] = ((unsigned long long)clock64() << 4) | (unsigned)(k); } while (0)
#define GANTT_DECL int gantt_n = 0;
#else
#define GANTT(k) do { } while (0)
#define GANTT_DECL
#endif
#define PHASE_DECL GANTT_DECL long long ph_t = clock64(); long long ph_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PHASE(k) do { GANTT(k); if (tid == TAMCMC_TRACE_TID) { const long long now_ = clock64(); ph_acc[k] += now_ - ph_t; ph_t = now_; } } while (0)
#define PHASE_FLUSH do { if (tid == TAMCMC_TRACE_TID) for (int k_ = 0; k_ < 10; k_++) TRACE(48 + k_, (unsigned long long)ph_acc[k_]); } while (0)
#else
#define TRACE(slot, val) do { } while (0)
#define PHASE_DECL
#define PHASE(k) do { } while (0)
#define PHASE_FLUSH do { } while (0)
#endif
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); }

// ------------------------------------------------------------------------------------------------
// producer warps
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ int warp_excl_scan(int v, int lane)
{
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += o; }
    return x - v;
}

__device__ __forceinline__ unsigned pop_item(const WhittleArgs& A, int lane)
{
    unsigned idx = 0;
    if (lane == 0) idx = atomicAdd(&A.qctl->head, 1u);
    return __shfl_sync(0xffffffffu, idx, 0);
}

// first slot after `c` (cyclically) whose producer has not drained the queue; `w` itself is never in the mask
__device__ __forceinline__ int next_live(int c, unsigned done_mask)
{
    int n = (c + 1) % NBUF;
#pragma unroll
    for (int k = 0; k < NBUF; k++) { if (!((done_mask >> n) & 1u)) break; n = (n + 1) % NBUF; }
    return n;
}

// ------------------------------------------------------------------------------------------------
// tile reduction, run by the slot's PRODUCER warp once the consumers have released the slot: the 384 per-thread sums
// (sum y/M, mantissa product and exponent sum of prod 1/M) become the tile's partial.  The consumers only store their
// three values and go on to the next tile -- no shuffle tree, no atomics, no combine on the FP64-bound warps.  Fixed
// shape (lane l takes threads l, l+32, ...; then a shuffle tree): bitwise reproducible whatever the scheduling.
// ------------------------------------------------------------------------------------------------
template <int TILE>
__device__ __forceinline__ void tile_reduce(const WhittleArgs& A, Smem<TILE>& sm, int w, int lane)
{
    const Segment<TILE>& sg = sm.seg[w];
    double s1 = 0.0, pm = 1.0;
    int pe = 0;
#pragma unroll
    for (int i = 0; i < NC / 32; i++) {
        s1 += sm.red_s[w][lane + 32 * i];
        pm *= sm.red_m[w][lane + 32 * i];          // each is a product of <= 4 mantissas in [1,2): 12 of them stay below 2^48
        pe += sm.red_e[w][lane + 32 * i];
    }
    {
        const int hi = __double2hiint(pm);
        const int k = (hi & 0x7ff00000) - 0x3ff00000;
        pm = __hiloint2double(hi - k, __double2loint(pm));
        pe += k >> 20;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, d);
        pm *= __shfl_down_sync(0xffffffffu, pm, d);     // 32 mantissas in [1,2) stay below 2^32
        pe += __shfl_down_sync(0xffffffffu, pe, d);
    }
    if (lane == 0) {
        const int hi = __double2hiint(pm);
        const int k = (hi & 0x7ff00000) - 0x3ff00000;
        double* part = A.partial + 3 * ((size_t)sg.sc_index * A.tiles_stride + sg.tile);
        part[0] = s1; part[1] = __hiloint2double(hi - k, __double2loint(pm)); part[2] = (double)(pe + (k >> 20));
    }
    __syncwarp();
}

// Taylor coefficients in u of one far component 1 / ((s u + c)^2 + a), added to this lane's partial sums (stride 33 doubles):
// f_0 = h, f_1 = p h, f_{k+1} = p f_k - q f_{k-1} with h = 1/(c^2 + a), p = -2 s c h, q = s^2 h (see producer_loop).
template <bool ASYM>
__device__ __forceinline__ void far_series(double* facc, double s, double c, double a, double Q0 = 1.0, double Q1 = 0.0, double Q2 = 0.0)
{
    const double w2 = fma(c, c, a);
    double h;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(h) : "d"(w2));
    h = fma(fma(-w2, h, 1.0), h, h);
    h = fma(fma(-w2, h, 1.0), h, h);          // two Newton steps: relative error ~1e-16
    const double sh = s * h;
    const double p = -2.0 * c * sh, q = s * sh;
    double f0 = h, f1 = p * h;
    if (ASYM) {
        // times the asymmetry factor of the mode (build_lorentzian.cpp:153-157), a quadratic in u: the Cauchy product, truncated alike
        facc[0] += Q0 * f0; facc[33] += fma(Q0, f1, Q1 * f0);
    } else { facc[0] += f0; facc[33] += f1; }
#pragma unroll
    for (int t = 2; t < NFAR; t++) {
        const double f2 = fma(p, f1, -(q * f0));
        if (ASYM) facc[33 * t] += fma(Q0, f2, fma(Q1, f1, Q2 * f0));
        else facc[33 * t] += f2;
        f0 = f1; f1 = f2;
    }
}

template <int TILE>
__device__ void producer_loop(const WhittleArgs& A, Smem<TILE>& sm, int w, int lane)
{
    // lane k < NBUCKETS keeps the inclusive prefix sum of the cost-class counts: item idx lives in the class whose
    // prefix first exceeds it
    static_assert(TAMCMC_NBUCKETS <= 32, "one lane per cost class");
    unsigned cum_inc = (lane < TAMCMC_NBUCKETS) ? A.qctl->count[lane] : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, cum_inc, d); if (lane >= d) cum_inc += o; }
    const unsigned ntot = __shfl_sync(0xffffffffu, cum_inc, 31);
    // Work distribution.  The queue is sorted heaviest-first (16 cost classes).
    //  * first item of every producer: STATIC, in snake order over the 4 x gridDim heaviest items (producer w of CTA k
    //    takes w*grid + k for even w, w*grid + grid-1-k for odd w), so every CTA starts with the same total weight;
    //  * afterwards: dynamic pops, but a producer pops only while its slot is among the next LOOK live slots after the
    //    one the consumers work on (LOOK = 2, 1 for the last ~4 items per CTA): items are claimed late, by the CTAs
    //    that are actually ahead, which is what keeps the tail of the persistent grid short.
    const unsigned G = gridDim.x;
    const unsigned nstatic = (unsigned)NPROD * G;
    const unsigned endgame_from = (ntot > 4u * G) ? ntot - 4u * G : 0u;
    Segment<TILE>* const sg = &sm.seg[w];
    unsigned long long* const full = &sm.full[w];
    unsigned long long* const empty = &sm.empty[w];
    unsigned use = 0;                 // segments published in this slot so far
    bool first_item = true;
    bool pending = false;             // the slot holds a finished tile whose per-thread sums are still to be reduced
    unsigned last_idx = 0;
    for (;;) {
        unsigned idx;
        if (first_item) {
            idx = (unsigned)w * G + ((w & 1) ? (G - 1u - blockIdx.x) : blockIdx.x);
            first_item = false;
        } else {
            // The next item is claimed as soon as the look-ahead rule allows, BEFORE the slot has been handed back: the pop, the
            // queue entry and the tile / star records (three dependent round trips) are then in flight while the consumers finish
            // the slot's tile; the wait for the slot comes where the slot is first written (below).  The finished tile's sums are
            // reduced further down, under the header round trip of the next tile: the consumers rewrite the slot's scratch only at
            // the end of the tile this producer is about to list.
#ifdef TAMCMC_CLAIM_AFTER_RELEASE
            if (use) mbar_wait(empty, (use - 1) & 1);          // claim nothing while this slot still holds a tile
#endif
            const int look = (last_idx + nstatic >= endgame_from) ? A.look_end : A.look;
            for (;;) {
                const int cur = sm.cons_cur;
                const unsigned dm = *reinterpret_cast<volatile unsigned int*>(&sm.done_mask);
                int n = next_live(cur, dm), hops = 1;
                while (n != w && hops < look) { n = next_live(n, dm); hops++; }
                if (n == w) break;
                __nanosleep(256);
            }
            idx = nstatic + pop_item(A, lane);
        }
        last_idx = idx;
        if (idx >= ntot) {
            if (use) mbar_wait(empty, (use - 1) & 1);
            if (pending) { tile_reduce<TILE>(A, sm, w, lane); pending = false; }
            if (lane == 0) { sg->flags = SEG_DONE; atomicOr(&sm.done_mask, 1u << w); mbar_arrive(full); mbar_arrive(full); }
            return;
        }
        const int bucket = __popc(__ballot_sync(0xffffffffu, lane < TAMCMC_NBUCKETS && idx >= cum_inc));
        const unsigned cbase = __shfl_sync(0xffffffffu, cum_inc, (bucket + 31) & 31);      // prefix of the class before
        const unsigned item = A.queue[(size_t)bucket * A.qcap + (idx - (bucket ? cbase : 0u))];
        const int sc = (int)(item / (unsigned)A.tiles_stride);
        const int tile = (int)(item - (unsigned)sc * (unsigned)A.tiles_stride);
        const StarDesc* sd = A.stars + sc / A.Nchains;
        const TileRec* tr = A.tilerec + item;
        // one round trip: star fields, tile record, chain flags
        const long long soff = sd->off;
        const int Nloc = sd->Nloc, bin0 = sd->bin0, nmodes = sd->nmodes_cap;
        const double xc = tr->xc;
        const int series_ok = tr->series_ok;
        const double umax = tr->umax;
        const double bgk = (lane < NB && series_ok) ? tr->bg[lane] : 0.0;
        const bool asym = A.asym_flag[sc] != 0;
        const double N0 = A.noise[sc].N0;
        const int gauss = A.noise[sc].gauss ? SEG_GAUSS : 0;      // Gaussian-envelope models (ids 0, 1)
        const int lb0 = tile * TILE;
        const int nvalid = min(TILE, Nloc - lb0);
        const long long off = soff + lb0;
        const int g0 = bin0 + lb0, gend = g0 + nvalid;
        const ModeRec* modes = A.modes + (size_t)sc * A.modes_stride;
        const CompRec* comps = A.comps + (size_t)sc * A.modes_stride * TAMCMC_MAX_COMP_PER_MODE;

        // Far field.  A symmetric Lorentzian in the scaled FAST form, 1 / ((s u + c)^2 + a), is analytic in u with poles at
        // distance |w|/s, w = c - i sqrt(a), from the tile centre.  When every component of a mode lies >= far_ratio * umax
        // from the centre (and the window covers the tile), its Taylor series in u converges like far_ratio^-k on the tile:
        //     f_k = (-s/|w|)^k U_k(c/|w|) / |w|^2        (U_k: Chebyshev polynomials of the second kind),
        // i.e. f_0 = h, f_1 = p h, f_{k+1} = p f_k - q f_{k-1} with h = 1/(c^2 + a), p = -2 s c h, q = s^2 h.  Truncated
        // after NFAR terms the relative error of the component is <= (NFAR+1) ratio^-NFAR ((1+1/ratio)/(1-1/ratio))^2
        // (1e-13 for 16 terms at ratio 8).  Such a mode costs NFAR producer-side steps per component instead of 4 FP64
        // instructions per component AND BIN; its coefficients join the background polynomial of the tile.
        // (a partial last tile qualifies too: its padding repeats the last x, so umax bounds |u| of every bin)
        const bool far_on = A.far_ratio > 0.0 && nmodes > 0;
        const double farR = A.far_ratio * umax;
        double* const facc = &sm.far_acc[w][0][lane];
        if (far_on) {
#pragma unroll
            for (int k = 0; k < NFAR; k++) facc[33 * k] = 0.0;
        }
        int any_far = 0;
        bool first = true;
        int cf = 0, cg = 0, ch = 0, seg_wide = 0;
        // ---- open the first segment: the slot must have been released; x and y start streaming in ----
        if (use) mbar_wait(empty, (use - 1) & 1);
        if (lane == 0) {
            mbar_arrive_expect_tx(full, 2u * TILE * (unsigned)sizeof(double));
            tma_load_1d(sg->x, A.x + off, TILE * sizeof(double), full);
            tma_load_1d(sg->y, A.y + off, TILE * sizeof(double), full);
        }
        // mode headers are fetched one batch ahead: the header round trip of batch k+1 overlaps the component-record round
        // trip of batch k (a tile-list build is a chain of dependent loads, and the first one is every CTA's start-up)
        int4 hnext = make_int4(0, 0, 0, 0);
        double2 fnext = make_double2(1.0, 0.0);
        // (the first batch is fetched for every lane below the table's stride, not below nmodes: the header round trip then
        // starts with the star / tile-record round trip instead of after it; lanes >= nmodes are ignored below)
        if (lane < A.modes_stride) {
            hnext = *reinterpret_cast<const int4*>(modes + lane);     // {i0, i1, ncomp, nfast | wide << 16}
            if (A.far_ratio > 0.0) fnext = *reinterpret_cast<const double2*>(&modes[lane].numin);
        }
        // the previous tile of this slot: its per-thread sums become the tile partial while the loads above are in flight
        // (tile_reduce reads the slot's sc_index / tile fields: nothing above has rewritten them yet)
        if (pending) { tile_reduce<TILE>(A, sm, w, lane); pending = false; }
        for (int base = 0; base < nmodes; base += 32) {
            const int mi = base + lane;
            int ncomp = 0, nfast = 0, ngen = 0, i0 = 0, i1 = 0, nfast_rec = 0, mwide = 0, wbit = 0, nfar = 0;
            const int4 h = hnext;
            const double2 fz = fnext;
            if (mi + 32 < nmodes) {
                hnext = *reinterpret_cast<const int4*>(modes + mi + 32);
                if (far_on) fnext = *reinterpret_cast<const double2*>(&modes[mi + 32].numin);
            }
            if (mi < nmodes) {
                if (h.z > 0 && h.x < gend && h.y > g0) {
                    ncomp = h.z; nfast_rec = h.w & 0xffff; i0 = h.x; i1 = h.y;
                    nfast = (h.x <= g0 && h.y >= gend) ? nfast_rec : 0;
                    ngen = ncomp - nfast;
                    wbit = (h.w >> 16) & 1;
                    if (far_on && nfast > 0 && fz.x <= fz.y && ((fz.x - xc) >= farR || (xc - fz.y) >= farR)) { nfar = nfast; nfast = 0; }
                    mwide = (nfast > 0) ? wbit : 0;
                }
            }
#ifdef TAMCMC_COMP_PREFETCH
            // (measured 3 % SLOWER on C2 and on 32 stars per launch: left out of the default build)
            // the component records this lane's mode will need (general/fast entries below, far series via the list) start
            // their way into L1 now: the scans, the list and the barrier below then overlap the L2 round trip
            if (ncomp > 0) {
                const char* p0 = reinterpret_cast<const char*>(comps + (size_t)mi * TAMCMC_MAX_COMP_PER_MODE);
                const char* p1 = p0 + ncomp * (int)sizeof(CompRec) - 1;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p0));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p1));
                if (ncomp > 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(p0 + 128));
            }
#endif
            // ONE inclusive scan of the four per-lane counts packed in 8-bit fields (each total <= 32 x 7 = 224): offsets and
            // totals of the fast entries, general entries, asym headers and far components of the batch
            const int hh = (asym && nfast > 0) ? 1 : 0;
            const unsigned packed = (unsigned)nfast | ((unsigned)ngen << 8) | ((unsigned)hh << 16) | ((unsigned)nfar << 24);
            unsigned incl = packed;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
            const unsigned tot = __shfl_sync(0xffffffffu, incl, 31), excl = incl - packed;
            // far components of this batch: spread evenly over the lanes (a mode's 2l+1 components would otherwise run
            // serially on its lane); their records are fetched now and turned into series after the entries below
            const int tfar = (int)(tot >> 24);
            constexpr int FR = 1;       // far records fetched ahead per lane (2 spills at the 128-register cap)
            double fr_nu[FR], fr_s[FR], fr_a[FR];
            if (tfar) {
                any_far = 1;
                const int ofar = (int)(excl >> 24);
                for (int k = 0; k < nfar; k++) sm.far_list[w][ofar + k] = mi * TAMCMC_MAX_COMP_PER_MODE + k;
                if (asym && nfar > 0) {
                    const ModeRec* mr = modes + mi;
                    const double qa = mr->qa, qb = mr->qb0 + xc * qa;
                    sm.far_q[w][lane][0] = fma(qb, qb, mr->qc); sm.far_q[w][lane][1] = 2.0 * qa * qb; sm.far_q[w][lane][2] = qa * qa;
                }
                __syncwarp();
#pragma unroll
                for (int r = 0; r < FR; r++) {
                    const int e = lane + 32 * r;
                    if (e < tfar) { const CompRec* c = comps + sm.far_list[w][e]; fr_nu[r] = c->nu; fr_s[r] = c->s; fr_a[r] = c->a; }
                }
            }
            if (!__any_sync(0xffffffffu, ncomp > 0)) continue;
            const bool whole_batch = (int)((tot >> 8) & 255u) <= CAPG;     // the usual case: the batch is listed in one go
            int sub_lo = 0;
            while (sub_lo < 32) {
                int sub_hi = 32;
                bool mine = lane >= sub_lo;
                int tf, tg, th;
                if (whole_batch) { tf = (int)(tot & 255u); tg = (int)((tot >> 8) & 255u); th = (int)((tot >> 16) & 255u); }
                else {
                    // too many general entries for one segment: 3 modes at a time (3 x 7 <= CAPG)
                    tg = warp_sum(mine ? ngen : 0);
                    if (tg > CAPG) { sub_hi = sub_lo + 3; mine = lane >= sub_lo && lane < sub_hi; tg = warp_sum(mine ? ngen : 0); }
                    tf = warp_sum(mine ? nfast : 0); th = warp_sum(mine ? hh : 0);
                }
                if (cf + tf > CAPF || cg + tg > CAPG || ch + th > CAPH) {
                    // ---- the slot is full: publish this segment, wait for the consumers to release the slot, go on ----
                    if (lane == 0) {
                        sg->nfast = cf; sg->ngen = cg; sg->nhdr = ch;
                        sg->flags = (first ? SEG_FIRST : 0) | (asym ? SEG_ASYM : 0) | (series_ok ? SEG_BGSERIES : 0) | (seg_wide ? SEG_WIDE : 0) | gauss;
                        sg->sc_index = sc; sg->tile = tile; sg->nvalid = nvalid; sg->lb0 = lb0; sg->off = off; sg->xc = xc; sg->N0 = N0;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full);
                    use++;
                    first = false; cf = cg = ch = 0; seg_wide = 0;
                    mbar_wait(empty, (use - 1) & 1);
                    if (lane == 0) mbar_arrive(full);          // this segment carries no TMA bytes
                }
                seg_wide |= __any_sync(0xffffffffu, mine && mwide) ? 1 : 0;
                const int mf = mine ? nfast : 0, mg = mine ? ngen : 0, mh = mine ? hh : 0;
                int of, og, oh;
                if (whole_batch) { of = cf + (int)(excl & 255u); og = cg + (int)((excl >> 8) & 255u); oh = ch + (int)((excl >> 16) & 255u); }
                else { of = cf + warp_excl_scan(mf, lane); og = cg + warp_excl_scan(mg, lane); oh = ch + warp_excl_scan(mh, lane); }
                if (mine && ncomp > 0) {
                    const CompRec* cp = comps + (size_t)mi * TAMCMC_MAX_COMP_PER_MODE;
                    double qa = 0.0, qb = 1.0, qc = 0.0;
                    if (asym || ngen > 0) { const ModeRec* mr = modes + mi; qa = mr->qa; qb = mr->qb0 + xc * qa; qc = mr->qc; }
                    if (hh) { ModeHdr m; m.qa = qa; m.qb = qb; m.qc = qc; m.begin = of; m.count = nfast; sg->hdr[oh] = m; }
                    const int nlead = nfast + nfar;          // leading components that are not general entries in this tile
#pragma unroll
                    for (int k0 = 0; k0 < 8; k0 += 4) {
                        double cnu[4], cs[4], ca[4];
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) { const int k = k0 + kk; if (k >= nfar && k < ncomp) { cnu[kk] = cp[k].nu; cs[kk] = cp[k].s; ca[kk] = cp[k].a; } }
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) {
                            const int k = k0 + kk;
                            if (k >= nfar && k < ncomp) {
                                const double cc = -(cnu[kk] - xc) * cs[kk];
                                if (k < nfast) { FastEntry fe; fe.s = cs[kk]; fe.c = cc; fe.a = ca[kk]; fe.pad = 0.0; sg->fast[of + k] = fe; }
                                else {
                                    // components are stored FAST-first; a FAST one lands here only on a window edge
                                    const bool ff = k < nfast_rec;
                                    GenEntry ge;
                                    ge.s = cs[kk]; ge.c = cc; ge.aadd = ff ? ca[kk] : 1.0; ge.num = ff ? 1.0 : ca[kk];
                                    // window in tile-local bins, clamped to the tile; bit 30 of hi: the entry needs an exponent
                                    // renormalisation after every merge (general form, or a WIDE-range mode)
                                    ge.qa = qa; ge.qb = qb; ge.qc = qc; ge.lo = max(i0 - g0, 0);
                                    ge.hi = min(i1 - g0, TILE) | ((!ff || wbit) ? (1 << 30) : 0);
                                    sg->gen[og + (k - nlead)] = ge;
                                }
                            }
                        }
                    }
                }
                cf += tf; cg += tg; ch += th;
                sub_lo = sub_hi;
            }
            if (tfar) {
                if (!asym) {
#pragma unroll
                    for (int r = 0; r < FR; r++)
                        if (lane + 32 * r < tfar) far_series<false>(facc, fr_s[r], -(fr_nu[r] - xc) * fr_s[r], fr_a[r]);
                    for (int e = lane + 32 * FR; e < tfar; e += 32) {
                        const CompRec* c = comps + sm.far_list[w][e];
                        const double cs_ = c->s;
                        far_series<false>(facc, cs_, -(c->nu - xc) * cs_, c->a);
                    }
                } else {
                    for (int e = lane; e < tfar; e += 32) {
                        const int idx = sm.far_list[w][e];
                        const CompRec* c = comps + idx;
                        const double* Q = sm.far_q[w][idx / TAMCMC_MAX_COMP_PER_MODE - base];
                        double cnu_, cs_, ca_;
                        static_assert(FR == 1, "the first round reads the one record fetched ahead");
                        if (e < 32 * FR) { cnu_ = fr_nu[0]; cs_ = fr_s[0]; ca_ = fr_a[0]; } else { cnu_ = c->nu; cs_ = c->s; ca_ = c->a; }
                        far_series<true>(facc, cs_, -(cnu_ - xc) * cs_, ca_, Q[0], Q[1], Q[2]);
                    }
                }
                __syncwarp();          // the list is rewritten by the next batch
            }
        }
        // ---- last segment of the tile (possibly empty: a tile no mode touches still has its background and Whittle terms)
        if (lane == 0) {
            sg->nfast = cf; sg->ngen = cg; sg->nhdr = ch;
            sg->flags = (first ? SEG_FIRST : 0) | SEG_LAST | (asym ? SEG_ASYM : 0) | (series_ok ? SEG_BGSERIES : 0) | (seg_wide ? SEG_WIDE : 0) | gauss
                      | (any_far ? SEG_FAR : 0);
            sg->sc_index = sc; sg->tile = tile; sg->nvalid = nvalid; sg->lb0 = lb0; sg->off = off; sg->xc = xc; sg->N0 = N0;
        }
        {
            // coefficient k of the tile polynomial: background series + the 32 per-lane far-field sums, in lane order (fixed shape)
            double ck = bgk;
            if (any_far) {
                __syncwarp();
                if (lane < NFAR) {
                    const double* row = &sm.far_acc[w][lane][0];
                    double fa = 0.0, fb = 0.0, fc = 0.0, fd = 0.0;     // four interleaved chains, fixed association
#pragma unroll
                    for (int j = 0; j < 32; j += 4) { fa += row[j]; fb += row[j + 1]; fc += row[j + 2]; fd += row[j + 3]; }
                    ck += (fa + fb) + (fc + fd);
                }
            }
            if (lane < NFAR) sg->bg[lane] = ck;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(full);             // 2 arrivals per phase: the opening one (with the TMA byte count) and this
        use++;
        pending = true;
    }
}

// ------------------------------------------------------------------------------------------------
// background-only tiles: ONE WARP PER TILE, straight from global memory
// ------------------------------------------------------------------------------------------------
// A tile no mode window touches has no component list: its model is the background alone.  Sent through the ring it still
// pays the whole per-tile protocol (claim, records, TMA, barriers, 12 warps in step: ~1.5-2 us for ~90 FP64 instructions per
// thread), and a long spectrum is mostly such tiles (BASELINE config C3: 85 % of 6510).  The expander lists them in a second
// queue instead (expand.cu, phase 3) and every consumer warp drains that queue before it enters the ring: a warp pops a tile,
// reads its x, y with coalesced 128-bit loads (48 bins per lane: deep instruction-level parallelism, no barrier, no shared
// memory), evaluates the background like the ring's epilogue does -- tile polynomial, or the exact per-bin terms where the series
// does not converge -- forms the Whittle terms and leaves the tile's partial.  The producers build the first ring tiles
// meanwhile, so the phase also fills the kernel's start-up.  Fixed order inside a warp: results stay bitwise reproducible.
struct BgArgs {      // what bg_phase reads, by value: a reference to the kernel's parameter block would force a local copy of all of it
    const unsigned int* bgqueue; QueueCtl* qctl; const StarDesc* stars; const TileRec* tilerec; const NoiseRec* noise;
    const double *x, *y, *lnx; double *partial, *model_out; int tiles_stride, Nchains;
};

// Whittle terms of one bin of a background-only tile: y / M and 1 / M split into mantissa and exponent (see the ring's epilogue)
struct BgSums {
    double s1 = 0.0, pm = 1.0;
    int pe = 0, bad = 0;
    __device__ __forceinline__ void add(double y, double minv)
    {
        s1 = fma(y, minv, s1);
        const int hi = __double2hiint(minv);
        const int k = (hi & 0x7ff00000) - 0x3ff00000;
        pm *= __hiloint2double(hi - k, __double2loint(minv));
        pe += k >> 20;
        bad |= hi;
    }
    __device__ __forceinline__ void fold()          // the mantissa product back into [1, 2)
    {
        const int hi = __double2hiint(pm);
        const int k = (hi & 0x7ff00000) - 0x3ff00000;
        pm = __hiloint2double(hi - k, __double2loint(pm));
        pe += k >> 20;
    }
};

template <bool WRITE_MODEL, int TILE>
__device__ __noinline__ void bg_phase(const BgArgs A, int lane)
{
    const unsigned nbg = A.qctl->bg_count;            // final: the expander grid completed before this grid started
    for (;;) {
        unsigned idx = 0;
        if (lane == 0) idx = atomicAdd(&A.qctl->bg_head, 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= nbg) return;
        const unsigned item = A.bgqueue[idx];
        const int sc = (int)(item / (unsigned)A.tiles_stride);
        const int tile = (int)(item - (unsigned)sc * (unsigned)A.tiles_stride);
        const StarDesc* sd = A.stars + sc / A.Nchains;
        const TileRec* tr = A.tilerec + item;
        const NoiseRec* nz = A.noise + sc;
        const int lb0 = tile * TILE;
        const long long off = sd->off + lb0;
        const int nvalid = min(TILE, sd->Nloc - lb0);
        const double xc = tr->xc, N0 = nz->N0;
        const double* __restrict__ xs = A.x + off + 2 * lane;
        const double* __restrict__ ys = A.y + off + 2 * lane;
        BgSums S;
        if (tr->series_ok) {
            // ---- tile polynomial (the usual case).  4 pairs per lane and array in a chunk; the NEXT chunk's loads are in flight
            // while this one is evaluated (16 x 512 B per warp) ----
            double cf[NB];
#pragma unroll
            for (int k = 0; k < NB; k++) cf[k] = tr->bg[k];
            constexpr int CH = 4;
            double2 xv[CH], yv[CH], xn[CH], yn[CH];
#pragma unroll
            for (int c = 0; c < CH; c++) { xv[c] = __ldg(reinterpret_cast<const double2*>(xs + 64 * c)); yv[c] = __ldg(reinterpret_cast<const double2*>(ys + 64 * c)); }
#pragma unroll 1
            for (int k0 = 0; k0 < TILE / 64; k0 += CH) {
                if (k0 + CH < TILE / 64) {
#pragma unroll
                    for (int c = 0; c < CH; c++) {
                        xn[c] = __ldg(reinterpret_cast<const double2*>(xs + 64 * (k0 + CH + c)));
                        yn[c] = __ldg(reinterpret_cast<const double2*>(ys + 64 * (k0 + CH + c)));
                    }
                }
#pragma unroll
                for (int j = 0; j < 2 * CH; j++) {
                    const int bb = 2 * lane + 64 * (k0 + (j >> 1)) + (j & 1);
                    const double u = ((j & 1) ? xv[j >> 1].y : xv[j >> 1].x) - xc;
                    double acc = cf[NB - 1];
#pragma unroll
                    for (int k = NB - 2; k >= 0; k--) acc = fma(acc, u, cf[k]);
                    const double num = acc + N0;
                    if (WRITE_MODEL) { if (bb < nvalid) A.model_out[lb0 + bb] = num; }
                    if (bb < nvalid) S.add((j & 1) ? yv[j >> 1].y : yv[j >> 1].x, fast_rcp(num));
                }
                S.fold();
#pragma unroll
                for (int c = 0; c < CH; c++) { xv[c] = xn[c]; yv[c] = yn[c]; }
            }
        } else {
            // ---- the exact per-bin terms where the series does not converge (near x = 0), merged into one fraction like the ring's
            // epilogue; one term at a time over the chunk's bins, so its parameters are read once per chunk ----
            const double* __restrict__ lx = A.lnx + off + 2 * lane;
            const int nh = nz->nh;
            constexpr int CH = 2;
#pragma unroll 1
            for (int k0 = 0; k0 < TILE / 64; k0 += CH) {
                double x[2 * CH], lnx[2 * CH], Nn[2 * CH], Dn[2 * CH];
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const double2 v = __ldg(reinterpret_cast<const double2*>(xs + 64 * (k0 + c)));
                    const double2 w = __ldg(reinterpret_cast<const double2*>(lx + 64 * (k0 + c)));
                    x[2 * c] = v.x; x[2 * c + 1] = v.y; lnx[2 * c] = w.x; lnx[2 * c + 1] = w.y;
                }
#pragma unroll
                for (int j = 0; j < 2 * CH; j++) { Nn[j] = 0.0; Dn[j] = 1.0; }
                for (int h = 0; h < nh; h++) {
                    const double H = nz->H[h], ls = nz->lnsc[h], pw = nz->pw[h], isc = nz->isc[h];
                    const bool p4 = (pw == 4.0), p2 = (pw == 2.0);
#pragma unroll
                    for (int j = 0; j < 2 * CH; j++) {
                        double z;
                        if (p4 || p2) { const double q = ((x[j] - xc) + xc) * isc, q2 = q * q; z = fmin(p4 ? q2 * q2 : q2, 2.5e30); }     // (u + xc like the ring)
                        else { const double arg = fmin(pw * (ls + lnx[j]), 70.0); z = (pw == 0.0) ? 1.0 : exp(arg); }
                        const double t = 1.0 + z;
                        Nn[j] = fma(Nn[j], t, H * Dn[j]);
                        Dn[j] *= t;
                    }
                }
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const double2 yv = __ldg(reinterpret_cast<const double2*>(ys + 64 * (k0 + c)));
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const int j = 2 * c + r, bb = 2 * lane + 64 * (k0 + c) + r;
                        const double num = fma(N0, Dn[j], Nn[j]);       // (D <= (1 + 2.5e30)^8: no renormalisation needed; N may be exactly 0)
                        if (WRITE_MODEL) { if (bb < nvalid) A.model_out[lb0 + bb] = num / Dn[j]; }
                        if (bb < nvalid) S.add(r ? yv.y : yv.x, Dn[j] * fast_rcp(num));
                    }
                }
                S.fold();
            }
        }
        double s1 = S.s1, pm = (S.bad < 0) ? nan("") : S.pm;       // a non-positive model bin: NaN like log() of it (likelihoods.cpp:23)
        int pe = S.pe;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s1 += __shfl_down_sync(0xffffffffu, s1, d);
            pm *= __shfl_down_sync(0xffffffffu, pm, d);     // 32 mantissas in [1, 2) stay below 2^32
            pe += __shfl_down_sync(0xffffffffu, pe, d);
        }
        if (lane == 0) {
            const int hi = __double2hiint(pm);
            const int k = (hi & 0x7ff00000) - 0x3ff00000;
            double* part = A.partial + 3 * ((size_t)sc * A.tiles_stride + tile);
            part[0] = s1; part[1] = __hiloint2double(hi - k, __double2loint(pm)); part[2] = (double)(pe + (k >> 20));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// consumer warps
// ------------------------------------------------------------------------------------------------
template <bool WRITE_MODEL, int BPT>
__device__ void consumer_loop(const WhittleArgs& A, Smem<NC * BPT>& sm, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    unsigned use[NBUF];
    for (int i = 0; i < NBUF; i++) use[i] = 0;
    int b = 0;
    unsigned done = 0;                  // producers (slots) that have delivered SEG_DONE
    double u[BPT], N[BPT], D[BPT], yv[BPT];
#pragma unroll
    for (int j = 0; j < BPT; j++) { u[j] = 0; N[j] = 0; D[j] = 1; yv[j] = 0; }

#ifdef TAMCMC_TRACE
    int tslot = 2;
    if (tid == 0) { TRACE(0, gtime()); unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); TRACE(63, (unsigned long long)smid + 1); }
#endif
    // De-phasing: the 12 consumer warps do the same work on every tile, so left alone they run in lock step and the FP64
    // pipe idles while ALL of them are in the latency-bound epilogue.  The three warps of each SM sub-partition start
    // `stagger_ns` apart; the offset persists (bounded by the ring depth), so one warp's epilogue overlaps the others'
    // main loops.
    // (applied after the FIRST tile has arrived: every warp waits for that one, an offset taken earlier would be lost there)
    bool staggered = !(A.stagger_ns > 0 && warp >= 4);
    PHASE_DECL
    for (;;) {
#ifdef TAMCMC_TRACE
        if (tid == 0) TRACE(tslot, gtime());        // begin waiting for a segment
#endif
        PHASE(9);
        mbar_wait(&sm.full[b], use[b] & 1);
        use[b]++;
        const Segment<NC * BPT>& sg = sm.seg[b];
        const int flags = sg.flags;
        PHASE(0);
#ifdef TAMCMC_TRACE
        if (tid == 0) { TRACE(tslot + 1, gtime()); tslot += 2; if (flags & SEG_DONE) TRACE(1, gtime()); }
#endif
        if (flags & SEG_DONE) {
            // this producer has drained the queue: drop its slot from the round-robin
            done |= 1u << b;
            if (done == (1u << NBUF) - 1u) { PHASE_FLUSH; return; }
            do { b = (b + 1 == NBUF) ? 0 : b + 1; } while ((done >> b) & 1u);
            continue;
        }
        const bool asym = (flags & SEG_ASYM) != 0;
        if ((flags & SEG_FIRST) && tid == 0) sm.cons_cur = b;
        if (!staggered) { staggered = true; __nanosleep((unsigned)A.stagger_ns * (unsigned)(warp >> 2)); }

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

        PHASE(1);
        // ---------- fast path: windows cover the whole tile, no masks ----------
        const int tot_fast = sg.nfast;
        if (!asym) {
            int k = 0;
            if (!(flags & SEG_WIDE)) {
                for (; k + GROUP <= tot_fast; k += GROUP) {
#pragma unroll
                    for (int kk = 0; kk < GROUP; kk++) {
                        const double2 p = *reinterpret_cast<const double2*>(&sg.fast[k + kk].s);
                        const double a = sg.fast[k + kk].a;
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
                for (; k < tot_fast; k++) {
                    const double2 p = *reinterpret_cast<const double2*>(&sg.fast[k].s);
                    const double a = sg.fast[k].a;
#pragma unroll
                    for (int j = 0; j < BPT; j++) {
                        const double e = fma(u[j], p.x, p.y);
                        const double t = fma(e, e, a);
                        N[j] = fma(N[j], t, D[j]);
                        D[j] *= t;
                    }
                }
            } else {
                // WIDE-range components in this segment: renormalise every GROUP_WIDE merges
                for (; k < tot_fast; k += GROUP_WIDE) {
#pragma unroll
                    for (int kk = 0; kk < GROUP_WIDE; kk++) {
                        if (k + kk < tot_fast) {
                            const double2 p = *reinterpret_cast<const double2*>(&sg.fast[k + kk].s);
                            const double a = sg.fast[k + kk].a;
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
            const int nhdr = sg.nhdr;
            for (int m = 0; m < nhdr; m++) {
                const ModeHdr h = sg.hdr[m];
                double q[BPT];
#pragma unroll
                for (int j = 0; j < BPT; j++) { const double w = fma(u[j], h.qa, h.qb); q[j] = fma(w, w, h.qc); }
                for (int k = h.begin; k < h.begin + h.count; k++) {
                    const double2 p = *reinterpret_cast<const double2*>(&sg.fast[k].s);
                    const double a = sg.fast[k].a;
#pragma unroll
                    for (int j = 0; j < BPT; j++) {
                        const double e = fma(u[j], p.x, p.y);
                        const double t = fma(e, e, a);
                        N[j] = fma(N[j], t, q[j] * D[j]);
                        D[j] *= t;
                    }
                }
                since += h.count;
                if (since + TAMCMC_MAX_COMP_PER_MODE > ((flags & SEG_WIDE) ? TAMCMC_MAX_COMP_PER_MODE : GROUP)) {
                    since = 0;
#pragma unroll
                    for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
        }

        // ---------- general path: window edges and extreme-dynamic-range components.  Every warp owns a
        // contiguous run of 64 bins per register pair (bins 2*tid + 2*NC*pj + r): a run the window covers is merged unmasked, a run holding
        // a window edge under a per-bin mask, the others are skipped; the three-way decision is warp-uniform. ----------
        PHASE(2);
        const int tot_gen = sg.ngen;
        int gsince[BPT / 2];                       // plain merges of each register pair since its last renormalisation
#pragma unroll
        for (int pj = 0; pj < BPT / 2; pj++) gsince[pj] = 0;
        for (int g = 0; g < tot_gen; g++) {
            const GenEntry ge = sg.gen[g];          // (prefetching the next entry measured slower)
            const int ghi = ge.hi & 0xffff;
            const bool heavy = (ge.hi >> 30) & 1;
#pragma unroll
            for (int pj = 0; pj < BPT / 2; pj++) {
                const int w0 = 2 * NC * pj + 64 * warp, w1 = w0 + 64;
                if (ghi <= w0 || ge.lo >= w1) continue;
                const bool whole = (ge.lo <= w0 && ghi >= w1);
                if (whole && !heavy) {
                    // the window covers this run and the component is in the FAST range: a plain merge
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
        if (tot_gen) {
#pragma unroll
            for (int j = 0; j < BPT; j++) renorm(N[j], D[j]);
        }

        PHASE(3);
#ifdef TAMCMC_TRACE
        // per-tile record of CTA row 1024 + blockIdx.x: list sizes of the segment whose wait ended at stamp tslot - 1
        if (tid == 0 && A.trace && (tslot >> 1) - 2 < 64 && (tslot >> 1) >= 2)
            A.trace[(size_t)(1024 + blockIdx.x) * 64 + (tslot >> 1) - 2] = (unsigned long long)sg.nfast | ((unsigned long long)sg.ngen << 16)
                | ((unsigned long long)sg.nhdr << 32) | ((unsigned long long)(flags & 0xff) << 48) | ((unsigned long long)(sg.tile & 0xff) << 56);
#endif
        if (flags & SEG_LAST) {
            const int sc = sg.sc_index, tile = sg.tile, nvalid = sg.nvalid, lb0 = sg.lb0;
            const double N0 = sg.N0;
            const NoiseRec* nz = A.noise + sc;
            double bgv[BPT];
            if (flags & SEG_FAR) {
                // background series + far Lorentzians: NFAR terms
                double acc[BPT];
#pragma unroll
                for (int j = 0; j < BPT; j++) acc[j] = sg.bg[NFAR - 1];
#pragma unroll
                for (int k = NFAR - 2; k >= 0; k--) {
                    const double ck = sg.bg[k];
#pragma unroll
                    for (int j = 0; j < BPT; j++) acc[j] = fma(acc[j], u[j], ck);
                }
#pragma unroll
                for (int j = 0; j < BPT; j++) bgv[j] = acc[j] + N0;
            } else if (flags & SEG_BGSERIES) {
                double cf[NB];
#pragma unroll
                for (int k = 0; k < NB; k++) cf[k] = sg.bg[k];
#pragma unroll
                for (int j = 0; j < BPT; j++) {
                    double acc = cf[NB - 1];
#pragma unroll
                    for (int k = NB - 2; k >= 0; k--) acc = fma(acc, u[j], cf[k]);      // Horner (Estrin's scheme measured slower:
                                                                                        // the epilogue is FP64-issue bound, not latency bound)
                    bgv[j] = acc + N0;
                }
            } else {
#pragma unroll
                for (int j = 0; j < BPT; j++) bgv[j] = N0;
            }
            if (!(flags & SEG_BGSERIES)) {
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
                const double xc = sg.xc;
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

            if (flags & SEG_GAUSS) {
                const double xc = sg.xc;
#pragma unroll
                for (int j = 0; j < BPT; j++) bgv[j] += gauss_envelope(nz, u[j] + xc);
            }

            // ---------- M = N/D + background; Whittle terms.  y_i/M_i is summed; ln M_i is carried as the
            // product of the 1/M_i split exactly into mantissa and integer exponent (no log in this kernel:
            // the per-chain finalisation takes one log per tile). ----------
            PHASE(4);
            double s1 = 0.0, pm = 1.0;
            int pe = 0;
            int bad = 0;                     // sign / NaN watch of 1/M_i: OR of the high words (see below)
            if (A.likelihood == 1) {
                // chi_square (likelihoods.cpp:31-40): sum of (y - M)^2 / sigma_y^2; the weights 1/sigma^2 were formed once at create
                const double* wg = A.wsig + sg.off;
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
            } else
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
            // The reference takes log() of every model bin (likelihoods.cpp:23): ONE non-positive M_i makes its logL NaN and
            // the proposal is rejected (MALA.cpp:522).  The product of the 1/M_i would hide an even number of negative factors,
            // so the sign bit of any of them poisons the tile (one integer OR per bin; heights and noise terms are abs()
            // values, so this only triggers on caller-supplied mode tables or data that defeat that).
            if (bad < 0) pm = nan("");
            // ---------- tile completion: every thread leaves its three sums in the slot's scratch; the slot's producer
            // warp reduces them (tile_reduce) after the empty barrier has handed the slot back.  Nobody waits for
            // anybody: this warp goes straight on to the next tile. ----------
            PHASE(5);
            sm.red_s[b][tid] = s1; sm.red_m[b][tid] = pm; sm.red_e[b][tid] = pe;
            PHASE(6);
            PHASE(7);
            __syncwarp();
            if (EMPTY_COUNT == NC || lane == 0) mbar_arrive(&sm.empty[b]);     // release: the slot and the sums go back to its producer warp
            PHASE(8);
            do { b = (b + 1 == NBUF) ? 0 : b + 1; } while ((done >> b) & 1u);      // next tile: next live slot
        } else {
            __syncwarp();
            if (EMPTY_COUNT == NC || lane == 0) mbar_arrive(&sm.empty[b]);     // more segments of this tile follow in the SAME slot
        }
    }
}

// Full-size tiles: one CTA per SM.  Half-size tiles: compiled for TWO resident CTAs per SM (64 registers per thread, 2 x 107 KB
// of shared memory): the two CTAs work on different tiles, so one CTA's latency-bound phases (tile epilogue, window edges,
// start-up) overlap the other's FP64 main loop.
template <bool WRITE_MODEL, int BPT>
#ifdef TAMCMC_MAXNREG
__global__ void __maxnreg__(TAMCMC_MAXNREG) tamcmc_whittle_kernel(WhittleArgs A)
#else
__global__ void __launch_bounds__(NT, (BPT == BPT_MAX) ? TAMCMC_MIN_CTAS : TAMCMC_MIN_CTAS_HALF) tamcmc_whittle_kernel(WhittleArgs A)
#endif
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<NC * BPT>& sm = *reinterpret_cast<Smem<NC * BPT>*>(smem_raw);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NBUF; i++) { mbar_init(&sm.full[i], 2); mbar_init(&sm.empty[i], EMPTY_COUNT); }
        sm.cons_cur = -1; sm.done_mask = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Programmatic dependent launch: this grid may have been scheduled while the expand kernel was still running (its
    // prologue above overlaps the expander's tail); everything below reads the expander's output.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid >= NC) producer_loop<NC * BPT>(A, sm, (tid - NC) >> 5, tid & 31);
    else {
        if (A.bgqueue) {                                                  // background-only tiles first, one per warp
#ifdef TAMCMC_TRACE
            if ((tid & 31) == 0 && (tid >> 5) < 2) TRACE(58 + 2 * (tid >> 5), gtime());      // warps 0, 1: begin / end of the background phase
#endif
            BgArgs b;
            b.bgqueue = A.bgqueue; b.qctl = A.qctl; b.stars = A.stars; b.tilerec = A.tilerec; b.noise = A.noise;
            b.x = A.x; b.y = A.y; b.lnx = A.lnx; b.partial = A.partial; b.model_out = A.model_out;
            b.tiles_stride = A.tiles_stride; b.Nchains = A.Nchains;
            bg_phase<WRITE_MODEL, NC * BPT>(b, tid & 31);
#ifdef TAMCMC_TRACE
            if ((tid & 31) == 0 && (tid >> 5) < 2) TRACE(59 + 2 * (tid >> 5), gtime());
#endif
        }
        consumer_loop<WRITE_MODEL, BPT>(A, sm, tid);
    }

    finish_launch(A, tid, NT);
}

__global__ void tamcmc_lnx_kernel(const double* __restrict__ x, double* __restrict__ lnx, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lnx[i] = log(x[i]);
}

__global__ void tamcmc_wsig_kernel(const double* __restrict__ sig, double* __restrict__ w, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = 1.0 / (sig[i] * sig[i]);          // sigma.array().square().cwiseInverse(), likelihoods.cpp:36
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

template <bool WM, int BPT>
static cudaError_t configure_one(int sms, int* grid)
{
    const int smem = (int)sizeof(Smem<NC * BPT>);
    cudaError_t e = cudaFuncSetAttribute(tamcmc_whittle_kernel<WM, BPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tamcmc_whittle_kernel<WM, BPT>, NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (grid) *grid = sms * per_sm;      // persistent: one CTA per resident slot
    return cudaSuccess;
}

cudaError_t tamcmc_whittle_configure(int* grid_full, int* grid_half)
{
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    if ((e = configure_one<false, BPT_MAX>(sms, grid_full)) != cudaSuccess) return e;
    if ((e = configure_one<true, BPT_MAX>(sms, nullptr)) != cudaSuccess) return e;
    if ((e = configure_one<false, BPT_MAX / 2>(sms, grid_half)) != cudaSuccess) return e;
    return configure_one<true, BPT_MAX / 2>(sms, nullptr);
}

cudaError_t tamcmc_launch_whittle(const WhittleArgs& a, int grid_ctas, bool write_model, int tile_bins, cudaStream_t st, bool pdl)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid_ctas, 1, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    WhittleArgs args = a;
    if (tile_bins == TILE_MAX) {
        cfg.dynamicSmemBytes = sizeof(Smem<TILE_MAX>);
        if (write_model) return cudaLaunchKernelEx(&cfg, tamcmc_whittle_kernel<true, BPT_MAX>, args);
        return cudaLaunchKernelEx(&cfg, tamcmc_whittle_kernel<false, BPT_MAX>, args);
    }
    if (tile_bins == TILE_MAX / 2) {
        cfg.dynamicSmemBytes = sizeof(Smem<TILE_MAX / 2>);
        if (write_model) return cudaLaunchKernelEx(&cfg, tamcmc_whittle_kernel<true, BPT_MAX / 2>, args);
        return cudaLaunchKernelEx(&cfg, tamcmc_whittle_kernel<false, BPT_MAX / 2>, args);
    }
    return cudaErrorInvalidValue;
}

cudaError_t tamcmc_launch_lnx(const double* x, double* lnx, long long n, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    tamcmc_lnx_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, lnx, n);
    return cudaGetLastError();
}

cudaError_t tamcmc_launch_wsig(double* w_inout, long long n, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    tamcmc_wsig_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w_inout, w_inout, n);
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
