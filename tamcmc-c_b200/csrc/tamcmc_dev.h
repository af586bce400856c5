// tamcmc_dev.h -- device-side data layout shared by the expander, the fused
// model+Whittle kernel and the C-ABI implementation.  Not part of the public ABI.
//
// HBM layout per context (see DESIGN.md "Data layout in HBM"):
//   x, y, lnx        : one concatenated FP64 array each over all stars' local bins (tile-padded)
//   params           : [nstars][Nchains][Nparams_max] FP64, row-major (uploaded every step)
//   modes / comps    : [nstars][Nchains][max_modes] ModeRec, [..][max_modes*7] CompRec
//   noise            : [nstars][Nchains] NoiseRec
//   queue            : [16][nstars*Nchains*max_tiles] uint32 work items in 16 cost classes, heaviest first
//   partial          : [nstars][Nchains][max_tiles] FP64 per-tile partial sums
//   out              : [nstars*Nchains] FP64 logL followed by [nstars*Nchains] int32 status
#pragma once
#include <cstdint>

#define TAMCMC_MAX_COMP_PER_MODE 7   // l <= 3 -> 2l+1 <= 7 (reference: acoefs.cpp supports l<=3)
#define TAMCMC_MAX_HARVEY 8
#define TAMCMC_BG_TERMS 10           // Taylor coefficients of the Harvey background per tile
#ifndef TAMCMC_FAR_TERMS
#define TAMCMC_FAR_TERMS 20          // coefficients of the per-tile polynomial once FAR Lorentzians are folded into it (whittle.cu)
#endif
#define TAMCMC_FAR_RATIO_DEFAULT 5.0 // a mode is FAR from a tile when every component centre is >= ratio * (tile half-width) away from the tile centre
#ifndef TAMCMC_TILE
#define TAMCMC_TILE 1536             // bins per tile
#endif
#ifndef TAMCMC_CONSUMERS
#define TAMCMC_CONSUMERS 384         // consumer threads per CTA (4 bins per thread); ONE persistent CTA per SM
#endif
#ifndef TAMCMC_PRODUCERS
#define TAMCMC_PRODUCERS 4           // producer warps per CTA = slots of the shared-memory ring
#endif
#define TAMCMC_THREADS (TAMCMC_CONSUMERS + 32 * TAMCMC_PRODUCERS)
#define TAMCMC_BINS_PER_THREAD (TAMCMC_TILE / TAMCMC_CONSUMERS)
#define TAMCMC_XCHG_MAX_WORLD 8      // ranks of one bin-sharded spectrum (the GPUs of one NVSwitch box)
#define TAMCMC_MAX_TILES 16384       // tiles per star the expander's cost scan supports (16.7M bins)
#ifndef TAMCMC_MIN_CTAS
#define TAMCMC_MIN_CTAS 1            // resident CTAs per SM the fused kernel is compiled for (full-size tiles)
#endif
#ifndef TAMCMC_MIN_CTAS_HALF
#define TAMCMC_MIN_CTAS_HALF 1       // ... half-size tiles (2 was measured: the 64-register cap spills in the main loop, 20 % slower)
#endif

// generic mode table (public id TAMCMC_MODEL_MODE_TABLE, include/tamcmc_gpu.h): a parameter row is
//   [nmodes, inclination, trunc_c, asym, noise[Nnoise], nmodes x {l, fc, H, W, a1..a6, eta0, extra[-3..3], 0, 0}]
#define TAMCMC_MODEL_ID_MODE_TABLE 1000
// Gaussian-envelope models (no Lorentzians): models.cpp:5728-5797 and 5674-5725
#define TAMCMC_MODEL_ID_KALLINGER_GAUSS 0
#define TAMCMC_MODEL_ID_HARVEY_GAUSS 1
#define TAMCMC_KSI_SLICE 2048        // bins per CTA of the normalisation pre-pass (get_ksinorm, noise_models.cpp:65-84)
#define TAMCMC_MT_HDR 4
#define TAMCMC_MT_STRIDE 20

// per-chain status bits (device -> host)
#define TAMCMC_ST_OK 0
#define TAMCMC_ST_WINDOW 1      // set_imin_imax: imax-imin <= 0 (reference: exit, build_lorentzian.cpp:650-665)
#define TAMCMC_ST_NONFINITE 2   // NaN/Inf in a derived mode quantity
#define TAMCMC_ST_BADCFG 4      // unsupported configuration value (e.g. filter_code, decompose_Alm)
#define TAMCMC_ST_INACTIVE 8    // active_mask[chain]==0: prior short-circuit (model_def.cpp:476-480)

// component flags
#define TAMCMC_CF_FAST 1        // scaled form t' = (1+e^2)/A, merged with 2 FP64 ops
#define TAMCMC_CF_SLOW 2        // general form (A, t), merged with 3 FP64 ops (extreme dynamic range)

struct __align__(16) ModeRec {
    // first 16 bytes = everything the tile classification needs (one 128-bit load)
    int i0, i1;          // GLOBAL bin window [i0, i1)  (set_imin_imax, bit-exact)
    int ncomp;           // number of live components (height != 0); FAST ones are stored first
    int nfast;           // bits 0-15: leading components in the scaled FAST form; bit 16: some of them are WIDE-range
                         // (the segments that list this mode renormalise every 4 merges instead of every 16)
    // next 16 bytes: what the far-field test of a tile needs (second 128-bit load, issued with the first)
    double numin, numax; // smallest / largest nu_nlm of the FAST components (+inf / -inf ... see expand.cu: never FAR when numin > numax)
    int l;               // degree
    int pad;
    double fc;           // central frequency fc_l
    double gamma;        // width
    double qa;           // asym / fc
    double qb0;          // 1 - asym                    (w(x) = qb0 + x*qa)
    double qc;           // (0.5*gamma*asym/fc)^2
};

struct __align__(8) CompRec {
    double nu;           // nu_nlm
    double s;            // FAST: 2/(gamma*sqrt(A)) ; SLOW: 2/gamma
    double a;            // FAST: 1/A               ; SLOW: A
    int flags; int m;
};

struct NoiseRec {
    int nh;                              // live Harvey terms (tau != 0)
    int gauss;                           // envelope models: 0 none, 1 Gaussian, 2 Gaussian times the sinc^2 leakage (Kallinger+2014 eq. 1)
    double N0;                           // white noise
    double gH, gnu, gk;                  // Gaussian envelope gH * exp(-(x - gnu)^2 * gk), gk = 0.5 / sigma^2 (models.cpp:5693-5694, 5771-5772)
    double xnyq;                         // leakage: eta = sin(a)/a, a = 0.5 pi x / x_nyquist (noise_models.cpp:89-97)
    double H[TAMCMC_MAX_HARVEY];         // heights
    double lnsc[TAMCMC_MAX_HARVEY];      // ln(1e-3 * tau)
    double isc[TAMCMC_MAX_HARVEY];       // 1e-3 * tau itself: integer exponents (2, 4) are evaluated as plain products
    double pw[TAMCMC_MAX_HARVEY];        // exponents
    double cpi[TAMCMC_MAX_HARVEY];       // cos(pi/p), sin(pi/p): direction of the nearest pole of 1/(1+z^p)
    double spi[TAMCMC_MAX_HARVEY];
    double binom[TAMCMC_MAX_HARVEY][TAMCMC_BG_TERMS];   // generalised binomial coefficients C(p, k)
};

struct StarDesc {
    long long off;       // offset of this star's local bins in the concatenated x/y/lnx arrays
    int Nloc;            // local bins held by this context
    int Nglob;           // bins of the whole spectrum (window clipping)
    int bin0;            // global index of local bin 0
    int tile0;           // first flat tile index
    int ntiles;
    int tile_bins;       // bins per tile of this context: TAMCMC_TILE, or TAMCMC_TILE / 2 for small problems
    int model_id;
    int Nparams;
    int nmodes_cap;      // modes per chain (capacity == exact count for the MS families)
    int plength[11];
    double x0, xlast;    // global x[0], x[N-1]
    double step;         // x[1]-x[0] (MS models, models.cpp:1952) or x[2]-x[1] (RGB v4, models.cpp:4714)
};

// ---- per-tile component lists (built by the producer warps of the fused kernel in their ring slots) ----
#define TAMCMC_CAPF 352              // fast entries per shared-memory segment
#define TAMCMC_CAPH 64               // mode headers per segment (asymmetric profiles)
#define TAMCMC_CAPG 24               // general entries per segment

struct __align__(16) FastEntry { double s, c, a, pad; };                 // e' = fma(u, s, c); t' = fma(e', e', a)
struct __align__(16) ModeHdr { double qa, qb, qc; int begin, count; };   // asym fast path: q(u) and its fast entries
struct __align__(16) GenEntry { double s, c, aadd, num, qa, qb, qc; int lo, hi; };

// per (star, chain, tile) record, written by the background CTAs of the expand launch
struct __align__(16) TileRec {
    double bg[TAMCMC_BG_TERMS];   // Taylor coefficients in u = x - xc of sum_h H_h/(1+(tau_h x)^p_h)
    double xc;                    // tile-local origin x[tile centre]
    double umax;                  // max |x - xc| over the tile
    int series_ok;                // 0: the tile must evaluate the background exactly per bin
    int pad[3];
};

// work queue header: count[k] items in cost bucket k (k = 0 heaviest), head = pop cursor
#define TAMCMC_NBUCKETS 16           // work-queue cost classes, heaviest first (fine classes = near-sorted pops = short tail)
#define TAMCMC_NBUCKETS_LOG2 4
// Zero between evaluations: the last CTA of the fused kernel to finish resets it.
// bg_count / bg_head: the second queue, of BACKGROUND-ONLY tiles (no mode window touches them): the fused kernel's consumer
// warps drain it one tile per WARP straight from global memory before they enter the ring (whittle.cu, bg_phase).
struct QueueCtl { unsigned int count[TAMCMC_NBUCKETS]; unsigned int head; unsigned int ctas_done; unsigned int bg_count; unsigned int bg_head; };
