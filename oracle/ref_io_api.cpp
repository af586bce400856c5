// ref_io_api.cpp -- C entry points over the reference's OWN .model reader and parameter-vector builder
// (tamcmc/sources/io_ms_global.cpp: read_MCMC_file_MS_Global :27-360, build_init_MS_Global :362-1400), compiled where they
// lie under /root/reference against oracle/eigen_shim.  TEST INFRASTRUCTURE ONLY (oracle/Makefile, target `refio`): the pin
// of tamcmc-c_b200/model_setup.py and the generator of tests/golden/reference_ms_global_init.json.
#include <Eigen/Dense>
#include <cstring>
#include <string>
#include <vector>
#include "data.h"
#include "io_ms_global.h"

extern "C" int refio_build_init_ms_global(const char* path, double resol, int cap, int* n_out, double* inputs, int* relax, double* priors /*[4][cap]*/,
                                          int* plength /*[11]*/, double* extra_priors /*[10]*/, char* names /*[cap][64]*/, char* prior_names /*[cap][32]*/,
                                          char* model_fullname /*[128]*/)
{
    const MCMC_files mf = read_MCMC_file_MS_Global(std::string(path), 0);
    const Input_Data in = build_init_MS_Global(mf, 0, resol);
    const int n = (int)in.inputs.size();
    *n_out = n;
    if (n > cap) return 1;
    for (int i = 0; i < n; i++) {
        inputs[i] = in.inputs[i];
        relax[i] = in.relax[i];
        for (int k = 0; k < 4; k++) priors[(size_t)k * cap + i] = in.priors(k, i);
        std::strncpy(names + (size_t)i * 64, in.inputs_names[(size_t)i].c_str(), 63); names[(size_t)i * 64 + 63] = 0;
        std::strncpy(prior_names + (size_t)i * 32, in.priors_names[(size_t)i].c_str(), 31); prior_names[(size_t)i * 32 + 31] = 0;
    }
    for (int k = 0; k < 11; k++) plength[k] = (k < in.plength.size()) ? in.plength[k] : 0;
    for (int k = 0; k < 10; k++) extra_priors[k] = (k < in.extra_priors.size()) ? in.extra_priors[k] : 0.0;
    std::strncpy(model_fullname, in.model_fullname.c_str(), 127); model_fullname[127] = 0;
    return 0;
}

#include "io_asymptotic.h"
// the red-giant dialect: read_MCMC_file_asymptotic + build_init_asymptotic (tamcmc/sources/io_asymptotic.cpp:27-875)
extern "C" int refio_build_init_asymptotic(const char* path, double resol, int cap, int* n_out, double* inputs, int* relax, double* priors /*[4][cap]*/,
                                           int* plength /*[11]*/, double* extra_priors /*[10]*/, char* names /*[cap][64]*/, char* prior_names /*[cap][32]*/,
                                           char* model_fullname /*[128]*/)
{
    const MCMC_files mf = read_MCMC_file_asymptotic(std::string(path), 0);
    const Input_Data in = build_init_asymptotic(mf, 0, resol);
    const int n = (int)in.inputs.size();
    *n_out = n;
    if (n > cap) return 1;
    for (int i = 0; i < n; i++) {
        inputs[i] = in.inputs[i];
        relax[i] = in.relax[i];
        for (int k = 0; k < 4; k++) priors[(size_t)k * cap + i] = in.priors(k, i);
        std::strncpy(names + (size_t)i * 64, in.inputs_names[(size_t)i].c_str(), 63); names[(size_t)i * 64 + 63] = 0;
        std::strncpy(prior_names + (size_t)i * 32, in.priors_names[(size_t)i].c_str(), 31); prior_names[(size_t)i * 32 + 31] = 0;
    }
    for (int k = 0; k < 11; k++) plength[k] = (k < in.plength.size()) ? in.plength[k] : 0;
    for (int k = 0; k < 10; k++) extra_priors[k] = (k < in.extra_priors.size()) ? in.extra_priors[k] : 0.0;
    std::strncpy(model_fullname, in.model_fullname.c_str(), 127); model_fullname[127] = 0;
    return 0;
}

#include "io_local.h"
// the local-fit dialect: read_MCMC_file_local + build_init_local (tamcmc/sources/io_local.cpp:25-327, 329-1176); slice_ind selects the
// '*' frequency range of the file that is analysed
extern "C" int refio_build_init_local(const char* path, int slice_ind, double resol, int cap, int* n_out, double* inputs, int* relax, double* priors /*[4][cap]*/,
                                      int* plength /*[11]*/, double* extra_priors /*[10]*/, char* names /*[cap][64]*/, char* prior_names /*[cap][32]*/,
                                      char* model_fullname /*[128]*/)
{
    const MCMC_files mf = read_MCMC_file_local(std::string(path), slice_ind, 0);
    const Input_Data in = build_init_local(mf, 0, resol);
    const int n = (int)in.inputs.size();
    *n_out = n;
    if (n > cap) return 1;
    for (int i = 0; i < n; i++) {
        inputs[i] = in.inputs[i];
        relax[i] = in.relax[i];
        for (int k = 0; k < 4; k++) priors[(size_t)k * cap + i] = in.priors(k, i);
        std::strncpy(names + (size_t)i * 64, in.inputs_names[(size_t)i].c_str(), 63); names[(size_t)i * 64 + 63] = 0;
        std::strncpy(prior_names + (size_t)i * 32, in.priors_names[(size_t)i].c_str(), 31); prior_names[(size_t)i * 32 + 31] = 0;
    }
    for (int k = 0; k < 11; k++) plength[k] = (k < in.plength.size()) ? in.plength[k] : 0;
    for (int k = 0; k < 10; k++) extra_priors[k] = (k < in.extra_priors.size()) ? in.extra_priors[k] : 0.0;
    std::strncpy(model_fullname, in.model_fullname.c_str(), 127); model_fullname[127] = 0;
    return 0;
}
