// host_rgb.hpp -- internal interface between the red-giant host expander (host_rgb.cpp) and the device solver (rgb_device.cu):
// tamcmc_host_expand_rgb_v4 in stages (prepare -> [pair loop + zeta normalisation] -> finish) and the flat task the device works on.
#pragma once
#include <vector>

namespace tamcmc_rgb {

// one p mode of one chain: the coarse grid nu[i] = lo + i gstep (last = hi) of solver_mm (external/ARMM/solver_mm.cpp:355-362), the band
// [i_lo, i_lo + nband) of it where p - g can change sign, the constants of p - g, the constants of the local grids (2 resol; resol * factor
// as the reference forms it in long double, split into two doubles), 1 / nu_g of the band's first g mode (the search of phase 1 runs with
// it), where the band's solution slots start (one per band index), and an estimate of the number of segments (poles of the tangent + 1)
struct Band { double nu_p, Dnu, lo, hi, gstep, DPl, q, resol2, Dh, Dl, rep_inv_g; int n, i_lo, nband, slot_off, chain, nseg_est; };
struct Pair { double inv_g, nu_g; int band, pad_; };                               // one (p mode, g mode): 1 / nu_g (double division), nu_g
struct KsiHdr { double fmin, fmax, c_up, pi_d; int Lp, Lg, Ndata, off_p, off_g, chain, val_off, pad_; };   // the zeta normalisation of one chain (bump_DP.cpp:126-163)
struct DeviceTask {
    std::vector<Band> bands;
    std::vector<Pair> pairs;
    int nslots = 0;                          // solution slots of all bands (sum of nband)
    int nvals = 0;                           // grid points of all zeta normalisations (sum of Ndata)
    std::vector<KsiHdr> ksi;
    std::vector<double> kp;                  // per p mode: nu_p, Dnu_p, q Dnu_p
    std::vector<double> kg;                  // per g mode: 1 / nu_g, DPl
    void clear() { bands.clear(); pairs.clear(); ksi.clear(); kp.clear(); kg.clear(); nslots = 0; nvals = 0; }
};

struct Prep;
Prep* prep_new();
void prep_free(Prep*);
int prepare(Prep*, int model_id, const double* params, const int* plength, double step, bool defer_solve);
bool export_task(const Prep*, int chain, DeviceTask&);
long local_grid_reference(const Prep*, double nu_idx, double* lo, double* hi);     // test hook: the long double original
int finish_modes(Prep*, bool deferred, const double* cand, int ncand);
int ksi_blocks(const Prep*);
void ksi_block_compute(Prep*, int block);
void ksi_done(Prep*);
double ksi_norm_at(const Prep*, const int* idx, int n);
int finish(Prep*, bool deferred, const double* cand, int ncand, double norm, int capacity, double* row_out, int* nmodes_out);

}  // namespace tamcmc_rgb
