// capi.cu -- implementation of the C ABI declared in include/tamcmc_gpu.h.
//
// Owns the device memory, the stream and the launch sequence
//     H2D(params) -> expand kernel -> fused model+Whittle kernel -> D2H(logL, status).
// No CPU evaluation path exists here: without a CUDA device every entry point fails.
#include "../../include/tamcmc_gpu.h"
#include "kernels.h"
#include "tamcmc_dev.h"
#include "host_math.hpp"

#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_last_error;

int fail_cuda(cudaError_t e, const char* what)
{
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return TAMCMC_ERR_CUDA;
}
#define CK(call)                                            \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

using tamcmc_host::Pslm;
using tamcmc_host::Qlm;
using tamcmc_host::fact_i;
using tamcmc_host::combi_i;

bool g_tables_uploaded[64] = {false};
int g_grid_ctas[64] = {0}, g_grid_ctas_half[64] = {0};

int upload_tables(int device)
{
    if (device >= 0 && device < 64 && g_tables_uploaded[device]) return TAMCMC_OK;
    double hi[7][4][7], lo[7][4][7], Q[4][7];
    std::memset(hi, 0, sizeof(hi)); std::memset(lo, 0, sizeof(lo)); std::memset(Q, 0, sizeof(Q));
    for (int s = 1; s <= 6; s++)
        for (int l = 0; l <= 3; l++)
            for (int m = -l; m <= l; m++) {
                const long double P = Pslm(s, l, m);
                const double h = (double)P;
                hi[s][l][m + 3] = h;
                lo[s][l][m + 3] = (double)(P - (long double)h);
            }
    for (int l = 0; l <= 3; l++)
        for (int m = -l; m <= l; m++) Q[l][m + 3] = Qlm(l, m);
    CK(tamcmc_upload_tables(&hi[0][0][0], &lo[0][0][0], &Q[0][0]));
    {
        // integer factors of dmm(l, i, 0, beta), i >= 0 (function_rot.cpp:76-88)
        double coef[4][4][4], nnum[4][4], nden[4];
        std::memset(coef, 0, sizeof(coef)); std::memset(nnum, 0, sizeof(nnum)); std::memset(nden, 0, sizeof(nden));
        for (int l = 0; l <= 3; l++) {
            nden[l] = std::sqrt((double)(fact_i(l) * fact_i(l)));
            for (int i = 0; i <= l; i++) {
                nnum[l][i] = std::sqrt((double)(fact_i(l + i) * fact_i(l - i)));
                for (int s2 = 0; s2 <= l - i; s2++)
                    coef[l][i][s2] = (double)combi_i(l, l - i - s2) * (double)combi_i(l, s2) * (((l - i - s2) & 1) ? -1.0 : 1.0);
            }
        }
        CK(tamcmc_upload_dmm_tables(&coef[0][0][0], &nnum[0][0], &nden[0]));
    }
    CK(tamcmc_expand_configure());
    { int per_sm = 0; CK(tamcmc_whittle_tiles_configure(&per_sm)); }
    { int g = 0, gh = 0; CK(tamcmc_whittle_configure(&g, &gh)); const int d = (device < 64 && device >= 0) ? device : 0; g_grid_ctas[d] = g; g_grid_ctas_half[d] = gh; }
    if (device >= 0 && device < 64) g_tables_uploaded[device] = true;
    return TAMCMC_OK;
}

int modes_of(int model_id, const int* pl)
{
    switch (model_id) {
    case 3: case 6: case 7: case 8: case 12: case 13: return pl[0] * (pl[1] + 1);
    // 18 / 19 (a1n / a1nl a2a3) print "not tested yet" and exit in the reference (models.cpp:599-603, 993-997): ERR_MODEL
    case 11: case 14: case 23: return pl[2] + pl[3] + pl[4] + pl[5];
    case TAMCMC_MODEL_ID_MODE_TABLE: return pl[0];
    case TAMCMC_MODEL_ID_KALLINGER_GAUSS: case TAMCMC_MODEL_ID_HARVEY_GAUSS: return 0;      // background + Gaussian envelope only
    }
    return -1;
}

}  // namespace

struct tamcmc_gpu_ctx {
    int device = 0;
    int nstars = 0, Nchains = 0;
    int params_stride = 0, modes_stride = 0, tiles_stride = 0, total_tiles = 0;
    long long total_bins_padded = 0;
    double p = 1.0;
    std::vector<StarDesc> h_stars;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    // device
    StarDesc* d_stars = nullptr;
    unsigned int* d_queue = nullptr;
    unsigned int* d_bgqueue = nullptr;        // background-only tiles (nullptr: every tile goes through the ring)
    TileRec* d_tilerec = nullptr;
    unsigned int* d_epoch = nullptr;
    unsigned long long* d_trace = nullptr;   // profiling aid (TAMCMC_TRACE builds)
    QueueCtl* d_qctl = nullptr;
    unsigned int qcap = 0;
    int grid_ctas = 0;
    int max_tiles = 0;
    int tile_bins = TAMCMC_TILE;     // TAMCMC_TILE, or half of it when the full-size tiles would leave most SMs idle
    double *d_x = nullptr, *d_y = nullptr, *d_lnx = nullptr, *d_wsig = nullptr;
    int likelihood = 0;
    double* d_params = nullptr;
    unsigned char* d_active = nullptr;
    ModeRec* d_modes = nullptr;
    CompRec* d_comps = nullptr;
    NoiseRec* d_noise = nullptr;
    int* d_asym = nullptr;
    double* d_Tcoefs = nullptr;
    double* d_partial = nullptr;
    int pending = 0;                         // tamcmc_gpu_eval_begin issued, tamcmc_gpu_eval_end not yet: 1 zero-copy, 2 copy engine
    double* d_ksi = nullptr;                 // get_ksinorm slice sums (Kallinger2014 model), [SC][ksi_slices][3]
    int ksi_slices = 0, ksi_slice_bins = TAMCMC_KSI_SLICE, ksi_maxN = 0;
    void* d_out = nullptr;          // [SC] double logL then [SC] int status
    double* d_model = nullptr;      // max Nloc
    // pinned host staging
    double* h_params = nullptr;
    unsigned char* h_active = nullptr;
    void* h_out = nullptr;
    QueueCtl* h_qctl = nullptr;
    // zero-copy host path of tamcmc_gpu_eval: parameters are read by the expander straight from mapped pinned memory and
    // the last CTA of the fused kernel writes results + a completion flag back into mapped pinned memory
    bool zero_copy = true;           // TAMCMC_GPU_NO_ZEROCOPY=1 selects the DMA path (H2D + D2H copies + stream sync)
    bool stage_params = false;       // zero-copy results, but the parameter block is copied to the device first (large rows)
    unsigned char* h_mirror = nullptr;   // [SC] double logL | [SC] int status | pad to 64 | (reserved word) | pad to 128 | uint flag
    double* dm_logL = nullptr; int* dm_status = nullptr; unsigned int* dm_overflow = nullptr; unsigned int* dm_flag = nullptr;
    double* dh_params = nullptr; unsigned char* dh_active = nullptr;    // device addresses of h_params / h_active
    unsigned int epoch_host = 1;     // mirrors the device epoch: advanced once per fused-kernel launch
    // CUDA graphs of the device-side sequence (memset, expand, tile lists, fused kernel, finalize), keyed by the
    // buffer pointers of the call
    struct GraphEntry { const double* p; const unsigned char* a; double* o; int raw; bool prof; bool mirror; cudaGraphExec_t exec; };
    GraphEntry graphs[4] = {};
    int ngraphs = 0;
    int stagger_ns = 0;               // TAMCMC_GPU_STAGGER_NS (tuning aid)
    int look = 3, look_end = 3;       // producer look-ahead in tiles (TAMCMC_GPU_LOOK / TAMCMC_GPU_LOOK_END = 1 .. producers - 1; tuning aid)
    double far_ratio = TAMCMC_FAR_RATIO_DEFAULT;   // far-field folding (TAMCMC_GPU_FAR_RATIO; 0 = off)
    bool use_graphs = true;
    int uniform_model = -1;          // model id shared by all stars (-1: mixed): the expander has builds specialised for 3, 23 and the mode table
    bool use_tiles = false;          // TAMCMC_GPU_KERNEL=tiles: the all-warps-on-one-tile schedule of whittle_tiles.cu instead of the
                                     // producer / consumer ring of whittle.cu (same results; measured slower, profiles/r2/NOTES.md)
    unsigned int nitems_max = 0;     // work items a launch can have: sum over the stars of ntiles x Nchains
    bool use_pdl = false;            // programmatic dependent launch expand -> fused kernel: measured no gain inside a CUDA graph
                                     // (profiles/r1/NOTES.md); TAMCMC_GPU_PDL=1 enables it
    // exchange of a bin-sharded spectrum (tamcmc_gpu_exchange_*): peer-mapped buffers of all ranks, this rank's counter
    int xworld = 1, xrank = 0, xstride = 0;
    void* x_own = nullptr;                    // this rank's exchange buffer (cudaMalloc; exported through cudaIpc)
    void* x_peer[TAMCMC_XCHG_MAX_WORLD] = {};  // every rank's buffer as mapped here (x_peer[xrank] == x_own)
    bool x_ipc[TAMCMC_XCHG_MAX_WORLD] = {};    // opened with cudaIpcOpenMemHandle (closed at destroy)
    unsigned int* d_xepoch = nullptr;
    // measurement
    bool profiling = false;
    long nlaunch_prof = 0;
    double expand_ms = 0, whittle_ms = 0;
    long launches = 0;
    long pairs_last = -1;

    size_t mirror_flag_off() const { return (((size_t)SC() * 12 + 63) / 64) * 64 + 64; }
    const double* hm_logL() const { return reinterpret_cast<const double*>(h_mirror); }
    const int* hm_status() const { return reinterpret_cast<const int*>(h_mirror + (size_t)SC() * 8); }
    volatile unsigned int* hm_overflow() const { return reinterpret_cast<volatile unsigned int*>(h_mirror + mirror_flag_off() - 64); }
    volatile unsigned int* hm_flag() const { return reinterpret_cast<volatile unsigned int*>(h_mirror + mirror_flag_off()); }
    int SC() const { return nstars * Nchains; }
    double* d_logL() const { return reinterpret_cast<double*>(d_out); }
    int* d_status() const { return reinterpret_cast<int*>(reinterpret_cast<double*>(d_out) + nstars * Nchains); }
    size_t out_bytes() const { return (size_t)SC() * (sizeof(double) + sizeof(int)); }
};

namespace {

ExpandArgs make_expand_args(tamcmc_gpu_ctx* c, const double* d_params, const unsigned char* d_active, double* d_logL)
{
    ExpandArgs a;
    a.stars = c->d_stars; a.params = d_params; a.active = d_active;
    a.modes = c->d_modes; a.comps = c->d_comps; a.noise = c->d_noise;
    a.status = c->d_status(); a.asym_flag = c->d_asym; a.out_logL = d_logL;
    a.queue = c->d_queue; a.qctl = c->d_qctl; a.qcap = c->qcap;
    a.tilerec = c->d_tilerec; a.x = c->d_x; a.lnx = c->d_lnx;
    a.Nchains = c->Nchains; a.params_stride = c->params_stride; a.modes_stride = c->modes_stride;
    a.tiles_stride = c->tiles_stride; a.max_tiles = c->max_tiles; a.trace = c->d_trace ? c->d_trace + 64 * 2048 : nullptr;
    a.ksi_part = c->d_ksi; a.ksi_slices = c->ksi_slices; a.ksi_slice_bins = c->ksi_slice_bins;
    a.far_ratio = c->far_ratio;
    a.bgqueue = c->use_tiles ? nullptr : c->d_bgqueue;
    a.mark_bgonly = c->use_tiles ? 1 : 0;
    a.uniform_model = c->uniform_model;
    return a;
}

WhittleArgs make_whittle_args(tamcmc_gpu_ctx* c, double* d_out, int raw_sum, bool mirror)
{
    WhittleArgs a;
    a.stars = c->d_stars;
    a.x = c->d_x; a.y = c->d_y; a.lnx = c->d_lnx; a.wsig = c->d_wsig; a.likelihood = c->likelihood;
    a.modes = c->d_modes; a.comps = c->d_comps; a.noise = c->d_noise;
    a.asym_flag = c->d_asym; a.Tcoefs = c->d_Tcoefs;
    a.queue = c->d_queue; a.bgqueue = c->use_tiles ? nullptr : c->d_bgqueue; a.qctl = c->d_qctl; a.qcap = c->qcap; a.tilerec = c->d_tilerec;
    a.partial = c->d_partial;
    a.out = d_out; a.model_out = c->d_model;
    a.p = c->p; a.Nchains = c->Nchains; a.modes_stride = c->modes_stride; a.tiles_stride = c->tiles_stride;
    a.raw_sum = raw_sum; a.trace = c->d_trace;
    a.epoch = c->d_epoch;
    a.look = c->look; a.look_end = c->look_end; a.stagger_ns = c->stagger_ns; a.far_ratio = c->far_ratio;
    a.status = c->d_status(); a.nsc = c->SC();
    // the host mirror costs a system-scope fence at the end of the launch: only the host-buffer entry point asks for it
    a.host_logL = mirror ? c->dm_logL : nullptr; a.host_status = mirror ? c->dm_status : nullptr;
    a.host_overflow = mirror ? c->dm_overflow : nullptr; a.host_flag = mirror ? c->dm_flag : nullptr;
    a.xworld = c->xworld; a.xrank = c->xrank; a.xepoch = c->d_xepoch; a.xstride = c->xstride;
    for (int r = 0; r < TAMCMC_XCHG_MAX_WORLD; r++) a.xpeer[r] = reinterpret_cast<unsigned long long>(c->x_peer[r]);
    return a;
}

// expand + fused kernel (tile lists, model, Whittle sums, per-chain finalisation, queue re-arm), enqueued on `st`
int enqueue_sequence(tamcmc_gpu_ctx* c, const double* d_params, const unsigned char* d_active, double* d_logL,
                     int raw_sum, cudaStream_t st, bool prof, bool capturing, bool mirror)
{
    ExpandArgs ea = make_expand_args(c, d_params, d_active, d_logL);
    WhittleArgs wa = make_whittle_args(c, d_logL, raw_sum, mirror);
    // profiling: CUDA events between the kernels.  Inside a captured graph they become event-record NODES
    // (cudaEventRecordExternal), so the durations are those of the replayed graph -- the configuration that is benchmarked --
    // without the host launch latency a per-kernel cudaEventRecord pair on a stream adds to a kernel this short.
#define REC(k) (capturing ? cudaEventRecordWithFlags(c->ev[k], st, cudaEventRecordExternal) : cudaEventRecord(c->ev[k], st))
    if (prof) CK(REC(0));
    CK(tamcmc_launch_expand(ea, c->SC(), st));
    if (prof) CK(REC(1));
    if (c->use_tiles) CK(tamcmc_launch_whittle_tiles(wa, c->nitems_max, false, c->tile_bins, st));
    else CK(tamcmc_launch_whittle(wa, c->grid_ctas, false, c->tile_bins, st, c->use_pdl && !prof));
    if (prof) CK(REC(2));
#undef REC
    return TAMCMC_OK;
}

// after a launch that failed part-way (e.g. the expander went out, the fused kernel did not): read the device's launch
// epoch back so that the next zero-copy evaluation waits for the right flag value
void resync_epoch(tamcmc_gpu_ctx* c)
{
    unsigned int e = 0;
    if (cudaStreamSynchronize(c->stream) == cudaSuccess && cudaMemcpy(&e, c->d_epoch, sizeof(e), cudaMemcpyDeviceToHost) == cudaSuccess && e) c->epoch_host = e;
    cudaMemsetAsync(c->d_qctl, 0, sizeof(QueueCtl), c->stream);      // an expander without its fused kernel leaves the queue armed
}

// One evaluation on `st`: the kernels are replayed as one CUDA graph (with event-record nodes between them while profiling).
int launch_eval(tamcmc_gpu_ctx* c, const double* d_params, const unsigned char* d_active, double* d_logL,
                int raw_sum, cudaStream_t st, bool mirror = false)
{
    // the host's copy of the launch epoch (what the last CTA of THIS launch will publish) advances only once the launch has
    // been accepted: a failed capture / instantiate / launch leaves host and device counters in step
    auto launched_ok = [c]() { c->launches += (c->d_ksi ? 3 : 2) + (c->use_tiles ? 1 : 0); const unsigned e = c->epoch_host + 1u; c->epoch_host = e ? e : 1u; };
    const bool prof = c->profiling && st == c->stream;
    if (!c->use_graphs) { const int rc = enqueue_sequence(c, d_params, d_active, d_logL, raw_sum, st, prof, false, mirror); if (rc == TAMCMC_OK) launched_ok(); else resync_epoch(c); return rc; }
    for (int i = 0; i < c->ngraphs; i++) {
        const tamcmc_gpu_ctx::GraphEntry& g = c->graphs[i];
        if (g.p == d_params && g.a == d_active && g.o == d_logL && g.raw == raw_sum && g.prof == prof && g.mirror == mirror) { CK(cudaGraphLaunch(g.exec, st)); launched_ok(); return TAMCMC_OK; }
    }
    if (c->ngraphs == 4) {           // evict the oldest
        cudaGraphExecDestroy(c->graphs[0].exec);
        for (int i = 1; i < 4; i++) c->graphs[i - 1] = c->graphs[i];
        c->ngraphs = 3;
    }
    cudaGraph_t graph = nullptr;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_sequence(c, d_params, d_active, d_logL, raw_sum, c->stream, prof, true, mirror);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail_cuda(e, "cudaStreamEndCapture");
    tamcmc_gpu_ctx::GraphEntry g;
    g.p = d_params; g.a = d_active; g.o = d_logL; g.raw = raw_sum; g.prof = prof; g.mirror = mirror; g.exec = nullptr;
    e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGraphInstantiate");
    c->graphs[c->ngraphs++] = g;
    CK(cudaGraphLaunch(g.exec, st));
    launched_ok();
    return TAMCMC_OK;
}

int collect_profile(tamcmc_gpu_ctx* c)
{
    if (!c->profiling) return TAMCMC_OK;
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, c->ev[0], c->ev[1]));
    CK(cudaEventElapsedTime(&b, c->ev[1], c->ev[2]));
    c->expand_ms += a; c->whittle_ms += b; c->nlaunch_prof++;
    return TAMCMC_OK;
}

// after an expand launch that is NOT followed by the fused kernel (debug entries): re-arm the queue like its last CTA does
int reset_queue(tamcmc_gpu_ctx* c)
{
    CK(cudaMemsetAsync(c->d_qctl, 0, sizeof(QueueCtl), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return TAMCMC_OK;
}

int status_to_rc(const int* st, int n)
{
    int rc = TAMCMC_OK;
    for (int i = 0; i < n; i++) {
        if (st[i] & TAMCMC_ST_BADCFG) return TAMCMC_ERR_MODEL;
        if (st[i] & TAMCMC_ST_WINDOW) rc = TAMCMC_ERR_WINDOW;
        else if ((st[i] & TAMCMC_ST_NONFINITE) && rc == TAMCMC_OK) rc = TAMCMC_ERR_NONFINITE;
    }
    return rc;
}

// evaluate one parameter row as chain 0 of `star` (all other chains masked); used by the debug entries
int expand_single(tamcmc_gpu_ctx* c, int star, const double* row)
{
    if (!c || !row || star < 0 || star >= c->nstars || c->pending) return TAMCMC_ERR_ARG;      // not re-entrant: one evaluation in flight
    const int SC = c->SC();
    std::memset(c->h_params, 0, sizeof(double) * (size_t)SC * c->params_stride);
    std::memset(c->h_active, 0, (size_t)SC);
    const int sc = star * c->Nchains;
    std::memcpy(c->h_params + (size_t)sc * c->params_stride, row, sizeof(double) * (size_t)c->h_stars[star].Nparams);
    c->h_active[sc] = 1;
    CK(cudaMemcpyAsync(c->d_params, c->h_params, sizeof(double) * (size_t)SC * c->params_stride, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_active, c->h_active, (size_t)SC, cudaMemcpyHostToDevice, c->stream));
    ExpandArgs ea = make_expand_args(c, c->d_params, c->d_active, c->d_logL());
    CK(tamcmc_launch_expand(ea, SC, c->stream));
    c->launches += 1;
    return TAMCMC_OK;
}

}  // namespace

extern "C" {

int tamcmc_gpu_abi_version(void) { return TAMCMC_GPU_ABI_VERSION; }

const char* tamcmc_gpu_last_error(void) { return g_last_error.c_str(); }

const char* tamcmc_gpu_strerror(int s)
{
    switch (s) {
    case TAMCMC_OK: return "ok";
    case TAMCMC_ERR_ARG: return "invalid argument";
    case TAMCMC_ERR_MODEL: return "model id unknown, obsolete or not on the GPU path";
    case TAMCMC_ERR_CUDA: return "CUDA error (no CPU fallback exists)";
    case TAMCMC_ERR_WINDOW: return "set_imin_imax: imax - imin <= 0 for some chain";
    case TAMCMC_ERR_NONFINITE: return "non-finite mode quantity for some chain";
    case TAMCMC_ERR_LIKELIHOOD: return "likelihood id unknown (model_def.cpp:405-416)";
    case TAMCMC_ERR_POOL: return "reserved";
    }
    return "unknown status";
}

int tamcmc_gpu_create(int device, int nstars, const tamcmc_gpu_star* stars, int Nchains,
                      const double* Tcoefs, double p, int likelihood_id, tamcmc_gpu_ctx** out)
{
    if (!out) return TAMCMC_ERR_ARG;
    *out = nullptr;
    if (nstars <= 0 || !stars || Nchains <= 0 || !Tcoefs) return TAMCMC_ERR_ARG;
    if (likelihood_id != TAMCMC_LIKELIHOOD_CHI22P && likelihood_id != TAMCMC_LIKELIHOOD_CHI_SQUARE) return TAMCMC_ERR_LIKELIHOOD;
    int ndev = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
        if (ndev <= 0 || device < 0 || device >= ndev) { g_last_error = "no usable CUDA device"; return TAMCMC_ERR_CUDA; }
    }
    CK(cudaSetDevice(device));
    { int rc = upload_tables(device); if (rc) return rc; }

    tamcmc_gpu_ctx* c = new tamcmc_gpu_ctx();
    c->device = device; c->nstars = nstars; c->Nchains = Nchains; c->p = p; c->likelihood = likelihood_id;
    if (const char* e = std::getenv("TAMCMC_GPU_NO_GRAPH")) c->use_graphs = !(e[0] == '1');
    if (const char* e = std::getenv("TAMCMC_GPU_KERNEL")) c->use_tiles = (std::strcmp(e, "tiles") == 0);
    if (const char* e = std::getenv("TAMCMC_GPU_PDL")) c->use_pdl = (e[0] == '1');
    if (const char* e = std::getenv("TAMCMC_GPU_STAGGER_NS")) c->stagger_ns = std::atoi(e);
    if (const char* e = std::getenv("TAMCMC_GPU_LOOK")) { const int v = std::atoi(e); if (v >= 1 && v < TAMCMC_PRODUCERS) c->look = v; }
    if (const char* e = std::getenv("TAMCMC_GPU_LOOK_END")) { const int v = std::atoi(e); if (v >= 1 && v < TAMCMC_PRODUCERS) c->look_end = v; }
    if (const char* e = std::getenv("TAMCMC_GPU_FAR_RATIO")) { const double v = std::atof(e); c->far_ratio = (v >= 4.0 && v <= 1e6) ? v : 0.0; }
    c->h_stars.resize(nstars);
    {
        // tile size: full-size tiles unless they would give fewer than ~2 work items per SM (spectra of a few thousand
        // bins: the red-giant slices of BASELINE configs C1/C4); TAMCMC_GPU_TILE overrides
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        long long items = 0;
        for (int s = 0; s < nstars; s++) items += (long long)Nchains * ((stars[s].N + TAMCMC_TILE - 1) / TAMCMC_TILE);
        c->tile_bins = (items >= 2LL * sms) ? TAMCMC_TILE : TAMCMC_TILE / 2;
        if (const char* e = std::getenv("TAMCMC_GPU_TILE")) { const int t = std::atoi(e); if (t == TAMCMC_TILE || t == TAMCMC_TILE / 2) c->tile_bins = t; }
    }
    const int TB = c->tile_bins;
    long long off = 0; int tiles = 0; int maxN = 0;
    for (int s = 0; s < nstars; s++) {
        const tamcmc_gpu_star& in = stars[s];
        StarDesc& sd = c->h_stars[s];
        const int nm = modes_of(in.model_id, in.plength);
        if (nm < 0) { delete c; return TAMCMC_ERR_MODEL; }
        int need = 0;
        for (int k = 0; k < 11; k++) { if (in.plength[k] < 0) { delete c; return TAMCMC_ERR_ARG; } need += in.plength[k]; }
        const bool mode_table = in.model_id == TAMCMC_MODEL_ID_MODE_TABLE;
        if (mode_table) {
            // plength = [capacity (modes per chain), step_mode, 0,0,0,0,0,0, Nnoise, 0, 0]
            need = TAMCMC_MT_HDR + in.plength[8] + TAMCMC_MT_STRIDE * in.plength[0];
            if (in.plength[1] > 1 || in.plength[8] < 1) { delete c; return TAMCMC_ERR_ARG; }
            for (int k = 2; k < 11; k++) if (k != 8 && in.plength[k] != 0) { delete c; return TAMCMC_ERR_ARG; }
            if (in.plength[1] == 1 && (in.N < 3 || in.N_global > 0)) { delete c; return TAMCMC_ERR_ARG; }
        }
        const bool envelope = in.model_id == TAMCMC_MODEL_ID_KALLINGER_GAUSS || in.model_id == TAMCMC_MODEL_ID_HARVEY_GAUSS;
        if (envelope) {
            // fixed parameter positions, plength is not read (models.cpp:5693-5701, 5741-5745); the Kallinger normalisation
            // integrates over the WHOLE spectrum (noise_models.cpp:65-84), so that model does not take a bin-range slice
            need = (in.model_id == TAMCMC_MODEL_ID_KALLINGER_GAUSS) ? 18 : 10;
            if (in.model_id == TAMCMC_MODEL_ID_KALLINGER_GAUSS && in.N_global > 0) { delete c; return TAMCMC_ERR_ARG; }
        }
        if (!in.x || !in.y || in.N < 2 || in.N > 2000000000L || in.Nparams < need || (nm == 0 && !envelope)) { delete c; return TAMCMC_ERR_ARG; }
        if (in.model_id == 3 || in.model_id == 6 || in.model_id == 7 || in.model_id == 8 || in.model_id == 12 || in.model_id == 13) {
            // the reference indexes fl_l[n] for n < Nmax and l <= lmax (models.cpp:2026-2075)
            for (int l = 0; l <= in.plength[1] && l <= 3; l++)
                if (in.plength[2 + l] < in.plength[0]) { delete c; return TAMCMC_ERR_ARG; }
            if (in.plength[1] > 3 || in.plength[0] < 2) { delete c; return TAMCMC_ERR_ARG; }
            // per-radial-order splittings live in the splitting block (models.cpp:229, 421, 817, 1015)
            const int per_n = (in.model_id == 7) ? 1 : (in.model_id == 8) ? 2 : 0;
            if (in.plength[6] < 6 + per_n * in.plength[0]) { delete c; return TAMCMC_ERR_ARG; }
        }
        if ((in.model_id == 23) && (in.plength[0] < 2 || in.plength[2] < 2)) { delete c; return TAMCMC_ERR_ARG; }
        sd.off = off;
        sd.Nloc = (int)in.N;
        const bool sharded = in.N_global > 0;
        sd.Nglob = sharded ? (int)in.N_global : (int)in.N;
        sd.bin0 = sharded ? (int)in.bin_offset : 0;
        sd.x0 = sharded ? in.x_first : in.x[0];
        sd.xlast = sharded ? in.x_last : in.x[in.N - 1];
        sd.step = sharded ? (in.x_second - in.x_first) : (in.x[1] - in.x[0]);
        if (mode_table && in.plength[1] == 1) sd.step = in.x[2] - in.x[1];      // RGB v4 models: models.cpp:4714
        if (sharded && (in.bin_offset < 0 || in.bin_offset + in.N > in.N_global)) { delete c; return TAMCMC_ERR_ARG; }
        sd.ntiles = (sd.Nloc + TB - 1) / TB;
        sd.tile_bins = TB;
        sd.tile0 = tiles;
        sd.model_id = in.model_id;
        sd.Nparams = in.Nparams;
        sd.nmodes_cap = nm;
        for (int k = 0; k < 11; k++) sd.plength[k] = envelope ? 0 : in.plength[k];
        if (in.model_id == TAMCMC_MODEL_ID_KALLINGER_GAUSS) c->ksi_maxN = std::max(c->ksi_maxN, sd.Nloc);
        tiles += sd.ntiles;
        off += (long long)sd.ntiles * TB;
        if (in.Nparams > c->params_stride) c->params_stride = in.Nparams;
        if (nm > c->modes_stride) c->modes_stride = nm;
        if (sd.ntiles > c->tiles_stride) c->tiles_stride = sd.ntiles;
        if (sd.ntiles > TAMCMC_MAX_TILES) { delete c; return TAMCMC_ERR_ARG; }
        if (sd.Nloc > maxN) maxN = sd.Nloc;
    }
    c->total_tiles = tiles;
    c->uniform_model = c->h_stars.empty() ? -1 : c->h_stars[0].model_id;
    for (const StarDesc& sd : c->h_stars) if (sd.model_id != c->uniform_model) c->uniform_model = -1;
    if (const char* e = std::getenv("TAMCMC_GPU_GENERIC_EXPANDER")) if (e[0] == '1') c->uniform_model = -1;
    c->nitems_max = (unsigned int)((size_t)tiles * (size_t)Nchains);
    if ((size_t)nstars * (size_t)Nchains * (size_t)c->tiles_stride >= 0x80000000ull) c->use_tiles = false;      // bit 31 of a queue entry is a flag there
    if (c->modes_stride < 1) c->modes_stride = 1;       // envelope models only: keep the mode tables non-empty
    c->max_tiles = c->tiles_stride;
    // the expander stages one parameter row + the per-tile cost array in (at most 96 KB of) shared memory
    if (sizeof(double) * (size_t)c->params_stride + sizeof(int) * 2 * (size_t)(c->max_tiles + 2) > 96u * 1024u) { delete c; return TAMCMC_ERR_ARG; }
    if ((size_t)nstars * (size_t)Nchains * (size_t)c->tiles_stride > 0xfffffff0ull) { delete c; return TAMCMC_ERR_ARG; }
    c->total_bins_padded = off;
    const int SC = c->SC();

#define CKC(call)                                                                  \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) { int rc__ = fail_cuda(e__, #call); tamcmc_gpu_destroy(c); return rc__; } \
    } while (0)

    CKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 3; i++) CKC(cudaEventCreate(&c->ev[i]));
    CKC(cudaMalloc(&c->d_stars, sizeof(StarDesc) * nstars));
    c->qcap = (unsigned int)((size_t)SC * (size_t)c->tiles_stride);
    CKC(cudaMalloc(&c->d_queue, sizeof(unsigned int) * TAMCMC_NBUCKETS * (size_t)c->qcap));
    {
        // background-only tiles bypass the ring for the chi(2,2p) likelihood (TAMCMC_GPU_BG_FAST=0 keeps them in it)
        bool bg_fast = likelihood_id == TAMCMC_LIKELIHOOD_CHI22P;
        if (const char* e = std::getenv("TAMCMC_GPU_BG_FAST")) bg_fast = bg_fast && !(e[0] == '0');
        if (bg_fast) CKC(cudaMalloc(&c->d_bgqueue, sizeof(unsigned int) * (size_t)c->qcap));
    }
    CKC(cudaMalloc(&c->d_qctl, sizeof(QueueCtl)));
    CKC(cudaMalloc(&c->d_epoch, sizeof(unsigned int)));
    { const unsigned int one = 1u; CKC(cudaMemcpy(c->d_epoch, &one, sizeof(one), cudaMemcpyHostToDevice)); }
    CKC(cudaMalloc(&c->d_tilerec, sizeof(TileRec) * (size_t)c->qcap));
#ifdef TAMCMC_TRACE
    CKC(cudaMalloc(&c->d_trace, sizeof(unsigned long long) * 64 * 4096));
    CKC(cudaMemset(c->d_trace, 0, sizeof(unsigned long long) * 64 * 4096));
#endif
    CKC(cudaMemset(c->d_qctl, 0, sizeof(QueueCtl)));
    c->grid_ctas = (c->tile_bins == TAMCMC_TILE) ? g_grid_ctas[device < 64 ? device : 0] : g_grid_ctas_half[device < 64 ? device : 0];
    CKC(cudaMalloc(&c->d_x, sizeof(double) * off));
    CKC(cudaMalloc(&c->d_y, sizeof(double) * off));
    CKC(cudaMalloc(&c->d_lnx, sizeof(double) * off));
    CKC(cudaMalloc(&c->d_params, sizeof(double) * ((size_t)SC * c->params_stride + 64)));
    CKC(cudaMemset(c->d_params, 0, sizeof(double) * ((size_t)SC * c->params_stride + 64)));
    CKC(cudaMalloc(&c->d_active, (size_t)SC));
    CKC(cudaMalloc(&c->d_modes, sizeof(ModeRec) * (size_t)SC * c->modes_stride));
    CKC(cudaMalloc(&c->d_comps, sizeof(CompRec) * (size_t)SC * c->modes_stride * TAMCMC_MAX_COMP_PER_MODE));
    CKC(cudaMalloc(&c->d_noise, sizeof(NoiseRec) * (size_t)SC));
    CKC(cudaMalloc(&c->d_asym, sizeof(int) * (size_t)SC));
    CKC(cudaMalloc(&c->d_Tcoefs, sizeof(double) * Nchains));
    CKC(cudaMalloc(&c->d_partial, sizeof(double) * 3 * (size_t)SC * c->tiles_stride));
    if (c->ksi_maxN > 0) {
        // at most ~2 waves of pre-pass CTAs (4 x 256 threads per SM resident); measured flat between 2048 and 4096 bins per CTA, slower at 8192 when the batch allows it: their prologue, not their bins, is the cost
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        while ((long long)SC * ((c->ksi_maxN + c->ksi_slice_bins - 1) / c->ksi_slice_bins) > 8LL * sms && c->ksi_slice_bins < (1 << 20)) c->ksi_slice_bins *= 2;
        if (const char* e = std::getenv("TAMCMC_GPU_KSI_SLICE")) { const int v = std::atoi(e); if (v >= 256 && v <= (1 << 20)) c->ksi_slice_bins = v; }
        c->ksi_slices = (c->ksi_maxN + c->ksi_slice_bins - 1) / c->ksi_slice_bins;
    }
    if (c->ksi_slices > 0) {
        CKC(cudaMalloc(&c->d_ksi, sizeof(double) * 3 * (size_t)SC * c->ksi_slices));
        CKC(cudaMemset(c->d_ksi, 0, sizeof(double) * 3 * (size_t)SC * c->ksi_slices));
    }
    CKC(cudaMalloc(&c->d_out, c->out_bytes()));
    CKC(cudaMemset(c->d_out, 0, c->out_bytes()));
    CKC(cudaMalloc(&c->d_model, sizeof(double) * (size_t)maxN));
    CKC(cudaHostAlloc(&c->h_params, sizeof(double) * (size_t)SC * c->params_stride, cudaHostAllocMapped));
    CKC(cudaHostAlloc(&c->h_active, (size_t)SC, cudaHostAllocMapped));
    CKC(cudaHostAlloc(&c->h_mirror, c->mirror_flag_off() + 64, cudaHostAllocMapped));
    std::memset(c->h_mirror, 0, c->mirror_flag_off() + 64);
    {
        void* dp = nullptr;
        CKC(cudaHostGetDevicePointer(&dp, c->h_params, 0)); c->dh_params = reinterpret_cast<double*>(dp);
        CKC(cudaHostGetDevicePointer(&dp, c->h_active, 0)); c->dh_active = reinterpret_cast<unsigned char*>(dp);
        CKC(cudaHostGetDevicePointer(&dp, c->h_mirror, 0));
        unsigned char* dm = reinterpret_cast<unsigned char*>(dp);
        c->dm_logL = reinterpret_cast<double*>(dm);
        c->dm_status = reinterpret_cast<int*>(dm + (size_t)SC * 8);
        c->dm_overflow = reinterpret_cast<unsigned int*>(dm + c->mirror_flag_off() - 64);
        c->dm_flag = reinterpret_cast<unsigned int*>(dm + c->mirror_flag_off());
    }
    // Host entry: small parameter blocks (C2: 11 KB) are read by the expander straight from mapped pinned memory; a large one
    // (red-giant mode tables: 10 chains x 2000 doubles = 160 KB of dependent PCIe reads) goes through ONE cudaMemcpyAsync into
    // device memory instead, results still come back through the mapped mirror + flag (no D2H copy, no stream sync)
    c->stage_params = sizeof(double) * (size_t)SC * c->params_stride > (size_t)32 * 1024;
    if (const char* e = std::getenv("TAMCMC_GPU_STAGE_PARAMS")) c->stage_params = (e[0] == '1');
    if (const char* e = std::getenv("TAMCMC_GPU_NO_ZEROCOPY")) c->zero_copy = !(e[0] == '1');
    CKC(cudaMallocHost(&c->h_out, c->out_bytes()));
    CKC(cudaMallocHost(&c->h_qctl, sizeof(QueueCtl)));

    // upload spectra (padded to a multiple of the tile: x pad = last x, y pad = 0)
    {
        std::vector<double> hx((size_t)off), hy((size_t)off, 0.0);
        for (int s = 0; s < nstars; s++) {
            const StarDesc& sd = c->h_stars[s];
            std::memcpy(&hx[(size_t)sd.off], stars[s].x, sizeof(double) * (size_t)sd.Nloc);
            std::memcpy(&hy[(size_t)sd.off], stars[s].y, sizeof(double) * (size_t)sd.Nloc);
            for (long long i = sd.Nloc; i < (long long)sd.ntiles * TB; i++) hx[(size_t)(sd.off + i)] = stars[s].x[sd.Nloc - 1];
            // the far-field folding bounds |x - xc| inside a tile by the tile's end points: a frequency axis that is not
            // monotone (the reference's windows assume x0 + i*step, build_lorentzian.cpp:645-646) turns the folding off
            bool up = true, down = true;
            for (int i = 1; i < sd.Nloc; i++) { up = up && stars[s].x[i] >= stars[s].x[i - 1]; down = down && stars[s].x[i] <= stars[s].x[i - 1]; }
            if (!up && !down) c->far_ratio = 0.0;
        }
        CKC(cudaMemcpy(c->d_x, hx.data(), sizeof(double) * off, cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(c->d_y, hy.data(), sizeof(double) * off, cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(c->d_stars, c->h_stars.data(), sizeof(StarDesc) * nstars, cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(c->d_Tcoefs, Tcoefs, sizeof(double) * Nchains, cudaMemcpyHostToDevice));
        CKC(tamcmc_launch_lnx(c->d_x, c->d_lnx, off, c->stream));
        c->launches += 1;
        if (likelihood_id == TAMCMC_LIKELIHOOD_CHI_SQUARE) {
            // weights 1/sigma_y^2 (pad: 1); a missing sigma_y column means sigma = 1 (config.cpp:367-374)
            std::vector<double> hs((size_t)off, 1.0);
            for (int s = 0; s < nstars; s++)
                if (stars[s].sigma_y) std::memcpy(&hs[(size_t)c->h_stars[s].off], stars[s].sigma_y, sizeof(double) * (size_t)c->h_stars[s].Nloc);
            CKC(cudaMalloc(&c->d_wsig, sizeof(double) * off));
            CKC(cudaMemcpy(c->d_wsig, hs.data(), sizeof(double) * off, cudaMemcpyHostToDevice));
            CKC(tamcmc_launch_wsig(c->d_wsig, off, c->stream));
            c->launches += 1;
        }
        CKC(cudaStreamSynchronize(c->stream));
    }
#undef CKC
    *out = c;
    return TAMCMC_OK;
}

void tamcmc_gpu_destroy(tamcmc_gpu_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_stars); cudaFree(c->d_queue); cudaFree(c->d_bgqueue); cudaFree(c->d_qctl); cudaFree(c->d_epoch); cudaFree(c->d_tilerec); cudaFree(c->d_trace); cudaFree(c->d_x); cudaFree(c->d_y); cudaFree(c->d_lnx); cudaFree(c->d_wsig);
    cudaFree(c->d_params); cudaFree(c->d_active); cudaFree(c->d_modes); cudaFree(c->d_comps); cudaFree(c->d_noise);
    cudaFree(c->d_asym); cudaFree(c->d_Tcoefs); cudaFree(c->d_partial); cudaFree(c->d_ksi); cudaFree(c->d_out);
    cudaFree(c->d_model);
    for (int r = 0; r < TAMCMC_XCHG_MAX_WORLD; r++) if (c->x_ipc[r] && c->x_peer[r]) cudaIpcCloseMemHandle(c->x_peer[r]);
    cudaFree(c->x_own); cudaFree(c->d_xepoch);
    if (c->h_params) cudaFreeHost(c->h_params);
    if (c->h_active) cudaFreeHost(c->h_active);
    if (c->h_mirror) cudaFreeHost(c->h_mirror);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->h_qctl) cudaFreeHost(c->h_qctl);
    for (int i = 0; i < c->ngraphs; i++) cudaGraphExecDestroy(c->graphs[i].exec);
    for (int i = 0; i < 3; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int tamcmc_gpu_params_stride(const tamcmc_gpu_ctx* c) { return c ? c->params_stride : 0; }
int tamcmc_gpu_nstars(const tamcmc_gpu_ctx* c) { return c ? c->nstars : 0; }
int tamcmc_gpu_nchains(const tamcmc_gpu_ctx* c) { return c ? c->Nchains : 0; }
long tamcmc_gpu_launch_count(const tamcmc_gpu_ctx* c) { return c ? c->launches : 0; }

double* tamcmc_gpu_params_staging(tamcmc_gpu_ctx* c, int* stride_out)
{
    if (!c) return nullptr;
    if (stride_out) *stride_out = c->params_stride;
    return c->h_params;
}

int tamcmc_gpu_eval_begin(tamcmc_gpu_ctx* c, const double* params, const unsigned char* active_mask)
{
    if (!c || !params || c->pending) return TAMCMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
    const int SC = c->SC();
    const size_t pbytes = sizeof(double) * (size_t)SC * c->params_stride;
    if (params != c->h_params) std::memcpy(c->h_params, params, pbytes);      // (rows built in place: tamcmc_gpu_params_staging)
    if (active_mask) std::memcpy(c->h_active, active_mask, (size_t)SC);
    if (c->zero_copy) {
        // ---- zero-copy: no DMA copies, no stream synchronisation.  The expander reads the rows over PCIe; the last CTA
        // of the fused kernel writes logL/status into the mapped mirror and publishes the launch's epoch in the flag ----
        const double* dp = c->dh_params;
        const unsigned char* da = active_mask ? c->dh_active : nullptr;
        if (c->stage_params) {
            CK(cudaMemcpyAsync(c->d_params, c->h_params, pbytes, cudaMemcpyHostToDevice, c->stream));
            dp = c->d_params;
            if (active_mask) { CK(cudaMemcpyAsync(c->d_active, c->h_active, (size_t)SC, cudaMemcpyHostToDevice, c->stream)); da = c->d_active; }
        }
        { int rc = launch_eval(c, dp, da, c->d_logL(), 0, c->stream, true); if (rc) return rc; }
        c->pending = 1;
    } else {
        CK(cudaMemcpyAsync(c->d_params, c->h_params, pbytes, cudaMemcpyHostToDevice, c->stream));
        const unsigned char* d_act = nullptr;
        if (active_mask) {
            CK(cudaMemcpyAsync(c->d_active, c->h_active, (size_t)SC, cudaMemcpyHostToDevice, c->stream));
            d_act = c->d_active;
        }
        { int rc = launch_eval(c, c->d_params, d_act, c->d_logL(), 0, c->stream); if (rc) return rc; }
        CK(cudaMemcpyAsync(c->h_out, c->d_out, c->out_bytes(), cudaMemcpyDeviceToHost, c->stream));
        c->pending = 2;
    }
    return TAMCMC_OK;
}

int tamcmc_gpu_eval_end(tamcmc_gpu_ctx* c, double* logL_out, int* status_out)
{
    if (!c || !logL_out || !c->pending) return TAMCMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
    const int SC = c->SC();
    const int mode = c->pending;
    c->pending = 0;
    const int* st = nullptr;
    if (mode == 1) {
        const unsigned int expected = c->epoch_host;
        volatile unsigned int* flag = c->hm_flag();
        unsigned long spins = 0;
        while (*flag != expected) {
            if ((++spins & 0x3fffu) == 0) {              // every 16k polls make sure the device is still healthy
                const cudaError_t e = cudaStreamQuery(c->stream);
                if (e == cudaSuccess) { if (*flag != expected) { g_last_error = "fused kernel finished without publishing its results"; return TAMCMC_ERR_CUDA; } break; }
                if (e != cudaErrorNotReady) return fail_cuda(e, "cudaStreamQuery (fused kernel)");
            }
            // back-off: a few thousand pause-spins cover the usual 50-100 us evaluation; after that the core is yielded between
            // polls, so a long batched evaluation (or many contexts polled from an OpenMP team) does not burn a host core each
            if (spins < 4096u) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            } else std::this_thread::yield();
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        if (c->profiling) CK(cudaStreamSynchronize(c->stream));
        { int rc = collect_profile(c); if (rc) return rc; }
        std::memcpy(logL_out, c->hm_logL(), sizeof(double) * (size_t)SC);
        st = c->hm_status();
    } else {
        CK(cudaStreamSynchronize(c->stream));
        { int rc = collect_profile(c); if (rc) return rc; }
        std::memcpy(logL_out, c->h_out, sizeof(double) * (size_t)SC);
        st = reinterpret_cast<const int*>(reinterpret_cast<const double*>(c->h_out) + SC);
    }
    if (status_out) std::memcpy(status_out, st, sizeof(int) * (size_t)SC);
    c->pairs_last = -1;
    return status_to_rc(st, SC);
}

int tamcmc_gpu_eval(tamcmc_gpu_ctx* c, const double* params, const unsigned char* active_mask,
                    double* logL_out, int* status_out)
{
    if (!c || !params || !logL_out) return TAMCMC_ERR_ARG;
    { int rc = tamcmc_gpu_eval_begin(c, params, active_mask); if (rc) return rc; }
    return tamcmc_gpu_eval_end(c, logL_out, status_out);
}

int tamcmc_gpu_eval_device(tamcmc_gpu_ctx* c, const double* d_params, const unsigned char* d_active,
                           double* d_logL, int raw_sum, void* stream)
{
    if (!c || !d_params || !d_logL || c->pending) return TAMCMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : c->stream;
    return launch_eval(c, d_params, d_active, d_logL, raw_sum, st);
}

int tamcmc_gpu_model(tamcmc_gpu_ctx* c, int star, const double* params_row, double* model_out)
{
    if (!model_out) return TAMCMC_ERR_ARG;
    { int rc = expand_single(c, star, params_row); if (rc) return rc; }
    const StarDesc& sd = c->h_stars[star];
    WhittleArgs wa = make_whittle_args(c, c->d_logL(), 0, false);
    { const cudaError_t e = c->use_tiles ? tamcmc_launch_whittle_tiles(wa, c->nitems_max, true, c->tile_bins, c->stream)
                                         : tamcmc_launch_whittle(wa, c->grid_ctas, true, c->tile_bins, c->stream, false); if (e != cudaSuccess) { resync_epoch(c); return fail_cuda(e, "tamcmc_launch_whittle"); } }
    c->launches += 1;
    { const unsigned e = c->epoch_host + 1u; c->epoch_host = e ? e : 1u; }       // only after the launch was accepted
    CK(cudaMemcpyAsync(c->h_out, c->d_out, c->out_bytes(), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const int SC = c->SC();
    const int* st = reinterpret_cast<const int*>(reinterpret_cast<const double*>(c->h_out) + SC);
    const int s0 = st[star * c->Nchains];
    if (s0 & TAMCMC_ST_BADCFG) return TAMCMC_ERR_MODEL;
    if (s0 & TAMCMC_ST_WINDOW) return TAMCMC_ERR_WINDOW;
    if (s0 & TAMCMC_ST_NONFINITE) return TAMCMC_ERR_NONFINITE;
    CK(cudaMemcpy(model_out, c->d_model, sizeof(double) * (size_t)sd.Nloc, cudaMemcpyDeviceToHost));
    return TAMCMC_OK;
}

int tamcmc_gpu_windows(tamcmc_gpu_ctx* c, int star, const double* params_row, int cap,
                       int* nmodes, int* l, int* imin, int* imax)
{
    if (!nmodes || !l || !imin || !imax || cap < 0) return TAMCMC_ERR_ARG;
    { int rc = expand_single(c, star, params_row); if (rc) return rc; }
    const StarDesc& sd = c->h_stars[star];
    std::vector<ModeRec> mr((size_t)sd.nmodes_cap);
    const size_t sc = (size_t)star * c->Nchains;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(mr.data(), c->d_modes + sc * c->modes_stride, sizeof(ModeRec) * mr.size(), cudaMemcpyDeviceToHost));
    int st = 0;
    CK(cudaMemcpy(&st, c->d_status() + sc, sizeof(int), cudaMemcpyDeviceToHost));
    { int rc = reset_queue(c); if (rc) return rc; }
    int nlive = sd.nmodes_cap;
    if (sd.model_id == TAMCMC_MODEL_ID_MODE_TABLE && params_row[0] >= 0 && params_row[0] <= sd.nmodes_cap) nlive = (int)params_row[0];
    *nmodes = nlive;
    for (int i = 0; i < nlive && i < cap; i++) { l[i] = mr[i].l; imin[i] = mr[i].i0; imax[i] = mr[i].i1; }
    if (st & TAMCMC_ST_BADCFG) return TAMCMC_ERR_MODEL;
    if (st & TAMCMC_ST_WINDOW) return TAMCMC_ERR_WINDOW;
    if (st & TAMCMC_ST_NONFINITE) return TAMCMC_ERR_NONFINITE;
    return TAMCMC_OK;
}

int tamcmc_gpu_components(tamcmc_gpu_ctx* c, int star, const double* params_row, int cap,
                          int* ncomp, int* mode_index, int* m, double* nu, double* height, double* width)
{
    if (!ncomp || cap < 0) return TAMCMC_ERR_ARG;
    { int rc = expand_single(c, star, params_row); if (rc) return rc; }
    const StarDesc& sd = c->h_stars[star];
    const size_t sc = (size_t)star * c->Nchains;
    std::vector<ModeRec> mr((size_t)sd.nmodes_cap);
    std::vector<CompRec> cr((size_t)sd.nmodes_cap * TAMCMC_MAX_COMP_PER_MODE);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(mr.data(), c->d_modes + sc * c->modes_stride, sizeof(ModeRec) * mr.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cr.data(), c->d_comps + sc * c->modes_stride * TAMCMC_MAX_COMP_PER_MODE, sizeof(CompRec) * cr.size(), cudaMemcpyDeviceToHost));
    { int rc = reset_queue(c); if (rc) return rc; }
    int n = 0;
    for (int i = 0; i < sd.nmodes_cap; i++)
        for (int k = 0; k < mr[i].ncomp; k++) {
            const CompRec& q = cr[(size_t)i * TAMCMC_MAX_COMP_PER_MODE + k];
            if (n < cap) {
                if (mode_index) mode_index[n] = i;
                if (m) m[n] = q.m;
                if (nu) nu[n] = q.nu;
                if (height) height[n] = (q.flags & TAMCMC_CF_FAST) ? 1.0 / q.a : q.a;
                if (width) width[n] = mr[i].gamma;
            }
            n++;
        }
    *ncomp = n;
    return TAMCMC_OK;
}

long tamcmc_gpu_pairs_last(tamcmc_gpu_ctx* c)
{
    if (!c) return -1;
    if (c->pairs_last >= 0) return c->pairs_last;
    if (cudaSetDevice(c->device) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    const int SC = c->SC();
    std::vector<ModeRec> mr((size_t)SC * c->modes_stride);
    std::vector<int> st((size_t)SC);
    if (cudaMemcpy(mr.data(), c->d_modes, sizeof(ModeRec) * mr.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (cudaMemcpy(st.data(), c->d_status(), sizeof(int) * SC, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    long P = 0;
    for (int sc = 0; sc < SC; sc++) {
        if (st[sc]) continue;
        const StarDesc& sd = c->h_stars[sc / c->Nchains];
        for (int i = 0; i < sd.nmodes_cap; i++) {
            const ModeRec& r = mr[(size_t)sc * c->modes_stride + i];
            long lo = r.i0 > sd.bin0 ? r.i0 : sd.bin0, hi = r.i1 < sd.bin0 + sd.Nloc ? r.i1 : sd.bin0 + sd.Nloc;
            if (hi > lo) P += (long)(2 * r.l + 1) * (hi - lo);
        }
    }
    c->pairs_last = P;
    return P;
}

int tamcmc_gpu_set_profiling(tamcmc_gpu_ctx* c, int on)
{
    if (!c) return TAMCMC_ERR_ARG;
    c->profiling = on != 0;
    c->nlaunch_prof = 0; c->expand_ms = 0; c->whittle_ms = 0;
    return TAMCMC_OK;
}

int tamcmc_gpu_get_kernel_ms(tamcmc_gpu_ctx* c, long* nlaunch, double* expand_ms_total, double* whittle_ms_total)
{
    if (!c) return TAMCMC_ERR_ARG;
    if (nlaunch) *nlaunch = c->nlaunch_prof;
    if (expand_ms_total) *expand_ms_total = c->expand_ms;
    if (whittle_ms_total) *whittle_ms_total = c->whittle_ms;
    return TAMCMC_OK;
}

int tamcmc_gpu_debug_trace(tamcmc_gpu_ctx* c, unsigned long long* out, int nctas)
{
    if (!c || !out || nctas <= 0 || nctas > 4096) return TAMCMC_ERR_ARG;
    if (!c->d_trace) return TAMCMC_ERR_ARG;      // library not built with -DTAMCMC_TRACE
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out, c->d_trace + (nctas == 1 ? 64 * 2048 : 0), sizeof(unsigned long long) * 64 * (size_t)nctas, cudaMemcpyDeviceToHost));
    CK(cudaMemset(c->d_trace, 0, sizeof(unsigned long long) * 64 * 4096));
    return TAMCMC_OK;
}

int tamcmc_gpu_pt_swap_device(tamcmc_gpu_ctx* c, int star, int A, double u, double* d_params, double* d_logL, double* d_logPrior,
                              int* d_swapped, void* stream)
{
    if (!c || !d_params || !d_logL || star < 0 || star >= c->nstars || A < 0 || A + 1 >= c->Nchains) return TAMCMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
    const size_t row0 = (size_t)star * c->Nchains;
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : c->stream;
    CK(tamcmc_launch_pt_swap(d_params + row0 * c->params_stride, d_logL + row0, d_logPrior ? d_logPrior + row0 : nullptr, c->d_Tcoefs,
                             c->params_stride, A, u, d_swapped, st));
    c->launches += 1;
    return TAMCMC_OK;
}

// ---- exchange step of a bin-sharded spectrum over NVLink peer memory ----
int tamcmc_gpu_exchange_create(tamcmc_gpu_ctx* c, void* handle_out)
{
    if (!c || !handle_out || c->pending) return TAMCMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == TAMCMC_XCHG_HANDLE_BYTES, "handle size of the ABI");
    if (!c->x_own) {
        c->xstride = ((c->SC() + 7) / 8) * 8;
        CK(cudaMalloc(&c->x_own, tamcmc_xchg_bytes(c->xstride)));
        CK(cudaMemset(c->x_own, 0, tamcmc_xchg_bytes(c->xstride)));
        CK(cudaMalloc(&c->d_xepoch, sizeof(unsigned int)));
        CK(cudaMemset(c->d_xepoch, 0, sizeof(unsigned int)));
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->x_own));
    std::memcpy(handle_out, &h, sizeof(h));
    return TAMCMC_OK;
}

static int exchange_finish_attach(tamcmc_gpu_ctx* c, int rank, int world)
{
    c->xrank = rank; c->xworld = world;
    // the kernel parameters change: graphs captured without the exchange are dropped
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < c->ngraphs; i++) cudaGraphExecDestroy(c->graphs[i].exec);
    c->ngraphs = 0;
    return TAMCMC_OK;
}

int tamcmc_gpu_exchange_attach(tamcmc_gpu_ctx* c, int rank, int world, const void* handles)
{
    if (!c || !handles || !c->x_own || world < 1 || world > TAMCMC_XCHG_MAX_WORLD || rank < 0 || rank >= world || c->pending) return TAMCMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
    for (int r = 0; r < world; r++) {
        if (r == rank) { c->x_peer[r] = c->x_own; c->x_ipc[r] = false; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * TAMCMC_XCHG_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->x_peer[r] = p; c->x_ipc[r] = true;
    }
    return exchange_finish_attach(c, rank, world);
}

int tamcmc_gpu_exchange_attach_ptrs(tamcmc_gpu_ctx* c, int rank, int world, void* const* bufs)
{
    if (!c || !bufs || !c->x_own || world < 1 || world > TAMCMC_XCHG_MAX_WORLD || rank < 0 || rank >= world || c->pending) return TAMCMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
    for (int r = 0; r < world; r++) { c->x_peer[r] = (r == rank) ? c->x_own : bufs[r]; c->x_ipc[r] = false; }
    return exchange_finish_attach(c, rank, world);
}

void* tamcmc_gpu_exchange_buffer(tamcmc_gpu_ctx* c) { return c ? c->x_own : nullptr; }

int tamcmc_gpu_sync(tamcmc_gpu_ctx* c)
{
    if (!c) return TAMCMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return collect_profile(c);
}

int tamcmc_gpu_fp64_peak(int device, double* tflops)
{
    if (!tflops) return TAMCMC_ERR_ARG;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
    if (ndev <= 0 || device < 0 || device >= ndev) { g_last_error = "no usable CUDA device"; return TAMCMC_ERR_CUDA; }
    CK(cudaSetDevice(device));
    float ms = 0;
    CK(tamcmc_fp64_peak(tflops, &ms, 4096));
    return TAMCMC_OK;
}

}  // extern "C"
