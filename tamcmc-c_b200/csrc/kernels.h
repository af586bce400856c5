// kernels.h -- launch interfaces between capi.cu and the kernel translation units.
#pragma once
#include "tamcmc_dev.h"
#include <cuda_runtime.h>

struct ExpandArgs {
    const StarDesc* stars;
    const double* params;            // [nstars*Nchains][params_stride]
    const unsigned char* active;     // [nstars*Nchains] or nullptr
    ModeRec* modes;                  // [nstars*Nchains][modes_stride]
    CompRec* comps;                  // [nstars*Nchains][modes_stride*7]
    NoiseRec* noise;                 // [nstars*Nchains]
    int* status;                     // [nstars*Nchains]
    int* asym_flag;                  // [nstars*Nchains]
    double* out_logL;                // [nstars*Nchains] (NaN written for non-OK chains)
    unsigned int* queue;             // [2][qcap] work items sc*tiles_stride + tile: heavy tiles, then light tiles
    TileRec* tilerec;                // [nstars*Nchains][tiles_stride]
    const double* x;                 // concatenated spectra (tile centres / extents)
    const double* lnx;
    QueueCtl* qctl;                  // all-zero at launch (reset by the previous fused-kernel launch)
    unsigned int qcap;
    int Nchains;
    int params_stride;
    int modes_stride;
    int tiles_stride;
    int max_tiles;                   // largest ntiles of any star (sizes the dynamic shared memory)
    unsigned long long* trace;       // profiling aid (TAMCMC_TRACE builds)
    double* ksi_part;                // [nstars*Nchains][ksi_slices][3] partial sums of get_ksinorm (Kallinger2014 model only), or nullptr
    int ksi_slices;
    int ksi_slice_bins;              // bins per pre-pass CTA: the smallest power-of-two multiple of TAMCMC_KSI_SLICE that fits the grid in one wave
    double far_ratio;                // far-field folding of the fused kernel (0 = off): only used to weigh the tiles of the work queue
    unsigned int* bgqueue;           // [qcap] work items of the background-only tiles (QueueCtl.bg_count of them); nullptr: every tile goes
                                     // through the ring (chi_square likelihood, TAMCMC_GPU_BG_FAST=0)
    int uniform_model;               // the model id all stars of the context share, or -1 (selects a model-specialised expander build)
    int mark_bgonly;                 // 1: bit 31 of a queue entry marks a tile no mode window touches (the one-CTA-per-tile kernel of
                                     // whittle_tiles.cu skips the mode lists of such a tile)
};

struct WhittleArgs {
    const StarDesc* stars;
    const double* x;                 // concatenated local bins, tile-padded
    const double* y;
    const double* lnx;
    const double* wsig;              // chi_square likelihood: 1/sigma_y^2 per bin (same layout as x); nullptr for chi(2,2p)
    const ModeRec* modes;
    const CompRec* comps;
    const NoiseRec* noise;
    const int* asym_flag;
    const double* Tcoefs;            // [Nchains]
    const unsigned int* queue;
    const unsigned int* bgqueue;     // background-only tiles (see ExpandArgs)
    const TileRec* tilerec;
    QueueCtl* qctl;
    unsigned int qcap;
    double* partial;                 // [nstars*Nchains][tiles_stride][3]: sum y/M, mantissa and exponent of prod 1/M
    double* out;                     // [nstars*Nchains]: tempered logL, or raw sum S if raw_sum
    double* model_out;               // WRITE_MODEL: [Nloc] of the (single) evaluated star/chain
    double p;                        // likelihood parameter (truncated to long like model_def.cpp:399)
    int Nchains;
    int modes_stride;
    int tiles_stride;
    unsigned long long* trace;       // profiling aid (builds with -DTAMCMC_TRACE): [grid][64] globaltimer stamps
    unsigned int* epoch;             // device launch counter (never 0), bumped by the last CTA of the fused kernel: the value
                                     // the host-mirror flag publishes
    // host mirror (mapped pinned memory, device addresses; null = none): the last CTA copies the results there and then
    // publishes the new epoch value in *host_flag, so the host can pick them up without a D2H copy or a stream sync
    double* host_logL; int* host_status; unsigned int* host_overflow; unsigned int* host_flag;
    const int* status;               // [nstars*Nchains] per-chain status written by the expand kernel
    int nsc;                         // nstars*Nchains
    int stagger_ns;                  // start offset between the consumer warps of one SM sub-partition (de-phasing)
    int look, look_end;              // producer look-ahead in tiles (1 or 2): steady state / last ~4 items per CTA
    int likelihood;                  // 0: chi(2,2p)  S = sum(ln M + y/M);  1: chi_square  S = sum((y-M)^2/sigma^2)
    int raw_sum;                     // 1: out = S = sum(ln M + y/M) over LOCAL bins (bin-sharded contexts)
    double far_ratio;                // > 0: modes whose components all lie >= far_ratio * umax from a tile's centre are folded into
                                     // the tile's polynomial instead of being merged per bin (whittle.cu); 0 = every component per bin
    // bin-sharded contexts with an attached exchange (capi.cu, tamcmc_gpu_exchange_*): the last CTA of every rank writes its chains'
    // local sums into the exchange buffer of EVERY rank over NVLink peer memory, waits for the other ranks' flags, adds the sums
    // in rank order (bitwise identical on every rank) and applies -p S / T -- the path's one exchange step, no collective launch
    int xworld, xrank;               // xworld <= 1: no exchange
    unsigned long long xpeer[TAMCMC_XCHG_MAX_WORLD];   // device addresses of the ranks' exchange buffers as mapped in THIS process
    unsigned int* xepoch;            // this rank's exchange counter (device): every rank advances it once per evaluation
    int xstride;                     // doubles per (parity, rank) block of an exchange buffer (>= nsc)
};

// exchange buffer of one rank: double S[2][TAMCMC_XCHG_MAX_WORLD][xstride], then unsigned flag[2][TAMCMC_XCHG_MAX_WORLD]
__host__ __device__ inline size_t tamcmc_xchg_flag_offset(int xstride) { return sizeof(double) * 2u * TAMCMC_XCHG_MAX_WORLD * (size_t)xstride; }
__host__ __device__ inline size_t tamcmc_xchg_bytes(int xstride) { return tamcmc_xchg_flag_offset(xstride) + sizeof(unsigned int) * 2u * TAMCMC_XCHG_MAX_WORLD; }

cudaError_t tamcmc_upload_tables(const double* P_hi, const double* P_lo, const double* Q);
cudaError_t tamcmc_upload_dmm_tables(const double* coef, const double* nnum, const double* nden);
cudaError_t tamcmc_launch_pt_swap(double* d_params_star, double* d_logL_star, double* d_logPrior_star, const double* d_Tcoefs, int stride,
                                 int A, double u, int* d_swapped, cudaStream_t st);
cudaError_t tamcmc_expand_configure();
cudaError_t tamcmc_launch_expand(const ExpandArgs& a, int nblocks, cudaStream_t st);
cudaError_t tamcmc_whittle_configure(int* grid_full, int* grid_half);   // one-time function attributes; persistent grid sizes for the two tile sizes
// pdl: programmatic dependent launch behind the expand kernel on the same stream
// tile_bins: TAMCMC_TILE or TAMCMC_TILE / 2 (the context's tile size, StarDesc.tile_bins)
cudaError_t tamcmc_launch_whittle(const WhittleArgs& a, int grid_ctas, bool write_model, int tile_bins, cudaStream_t st, bool pdl);
// one CTA per tile (whittle_tiles.cu): nitems_max = upper bound of the work items the expander can enqueue (CTAs beyond the
// queue's actual length only take part in the end-of-launch protocol)
cudaError_t tamcmc_whittle_tiles_configure(int* ctas_per_sm);
cudaError_t tamcmc_launch_whittle_tiles(const WhittleArgs& a, unsigned nitems_max, bool write_model, int tile_bins, cudaStream_t st);
cudaError_t tamcmc_launch_wsig(double* sigma_in_weights_out, long long n, cudaStream_t st);   // in place: w = 1/sigma^2
cudaError_t tamcmc_launch_lnx(const double* x, double* lnx, long long n, cudaStream_t st);
// DFMA throughput microbenchmark: returns achieved FP64 TFLOP/s (2 flops per DFMA)
cudaError_t tamcmc_fp64_peak(double* tflops, float* ms, int iters);
