// outputs.hpp -- the reference's binary chain outputs, written by the C++ driver side (host only, header only).
//
//   <prefix>.hdr            ASCII metadata: '#' comments and '! key= values' lines
//   <prefix>_chain-<k>.bin  raw float64, one row of Nvars values per kept sample, one file per chain
//
// Same files as Outputs::write_bin_params (tamcmc/sources/outputs.cpp:1231-1334), so the reference's post-processing tools
// (bin2txt, getstats) read a run of mcmc_driver.hpp unchanged; tamcmc-c_b200/formats.py holds the Python reader/writer of the
// same format.  Errors are returned, never fatal (the reference exits when a file cannot be opened, outputs.cpp:1322-1327).
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

namespace tamcmc {
namespace outputs {

// A row vector the way Eigen's operator<< prints `v.transpose()` (outputs.cpp:1270, 1274, 1289): stream default precision
// (6 significant digits), every coefficient right-aligned to the widest one, single-space separators.
template <class T>
inline std::string eigen_row(const std::vector<T>& v, bool integers)
{
    std::vector<std::string> txt;
    size_t w = 0;
    for (const T& x : v) {
        char buf[64];
        if (integers) std::snprintf(buf, sizeof buf, "%ld", (long)x);
        else std::snprintf(buf, sizeof buf, "%g", (double)x);
        txt.emplace_back(buf);
        w = std::max(w, txt.back().size());
    }
    std::string out;
    for (size_t i = 0; i < txt.size(); i++) { if (i) out += " "; out += std::string(w - txt[i].size(), ' ') + txt[i]; }
    return out;
}

struct ParamsMeta {
    long Nsamples = 0;                       // samples the run was asked for
    int Nchains = 0;
    std::vector<int> relax, plength;         // one relax flag per parameter (variables and constants)
    std::vector<std::string> cons_names;     // {"None"}: no constant (the reference then writes -1 as the value)
    std::vector<double> cons_values;
    std::vector<std::string> var_names;
};

inline std::string params_header_text(const ParamsMeta& m, long Nsamples_done)
{
    std::string s;
    s += "# This is the header file of the BINARY output file for the model parameters \n";
    s += "# This file contains values for vars[0:Nchains-1][ 0:Nvars-1]. Each matrix is in a different file, indexed by the chain number\n";
    s += "! Nsamples= " + std::to_string(m.Nsamples) + "\n";
    s += "! Nchains= " + std::to_string(m.Nchains) + "\n";
    s += "! Nsamples_done=" + std::to_string(Nsamples_done) + "\n";
    s += "! Nvars= " + std::to_string(m.var_names.size()) + "\n";
    s += "! Ncons= " + std::to_string(m.cons_names.size()) + "\n";
    s += "! relax= " + eigen_row(m.relax, true) + "\n";
    s += "! plength= " + eigen_row(m.plength, true) + "\n";
    s += "! constant_names= ";
    for (const auto& n : m.cons_names) s += n + "   ";
    s += "\n! constant_values= ";
    s += (!m.cons_names.empty() && m.cons_names[0] == "None") ? std::string("-1") : eigen_row(m.cons_values, false);
    s += "\n! variable_names=";
    for (const auto& n : m.var_names) s += n + "   ";
    s += "\n";
    return s;
}

// Writes (first == true: header + fresh chain files) or appends (first == false: chain files only, like the reference's
// buffered writes after the first) `nrows` samples held as vars[row][chain][var] (row-major, contiguous).  0 on success.
inline int write_params(const std::string& prefix, const ParamsMeta& m, const double* vars, long nrows, long Nsamples_done, bool first)
{
    const size_t nv = m.var_names.size();
    if (first) {
        std::ofstream h((prefix + ".hdr").c_str());
        if (!h.is_open()) return -1;
        h << params_header_text(m, Nsamples_done);
    }
    std::vector<double> row(nv);
    for (int c = 0; c < m.Nchains; c++) {
        std::ofstream f((prefix + "_chain-" + std::to_string(c) + ".bin").c_str(),
                        first ? (std::ofstream::binary | std::ofstream::trunc) : (std::ofstream::binary | std::ofstream::app));
        if (!f.is_open()) return -2;
        for (long i = 0; i < nrows; i++)
            f.write(reinterpret_cast<const char*>(vars + ((size_t)i * m.Nchains + c) * nv), (std::streamsize)(nv * sizeof(double)));
        f.flush();
        if (!f.good()) return -3;
    }
    return 0;
}

// ---- restore files (Outputs::write_buffer_restore, outputs.cpp:863-1027): <dir>/<id>_restore_<phase>_{1,2,3}.dat ----
// What Driver::restore_variables / restore_proposal need, with the reference's names.  `*_mean` are the averages over the last
// buffer (do_restore_proposal_mean, MALA.cpp:206-224).
struct RestoreState {
    int Nchains = 0, Nvars = 0;
    long iteration = 0;
    std::vector<std::string> variable_names;
    std::vector<double> vars, vars_mean;             // [Nchains][Nvars]
    std::vector<double> sigmas, sigmas_mean;         // [Nchains]
    std::vector<double> mus, mus_mean;               // [Nchains][Nvars]
    std::vector<double> covarmats, covarmats_mean;   // [Nchains][Nvars][Nvars]
};

// Parses ONE restore file into `st` (keys it does not hold are left alone).  0 on success, <0 when the file cannot be opened or
// a block does not have the size the header announces.
inline int read_restore_file(const std::string& path, RestoreState& st)
{
    std::ifstream f(path.c_str());
    if (!f.is_open()) return -1;
    auto numbers = [](const std::string& s, std::vector<double>& out) {
        const char* p = s.c_str();
        char* e = nullptr;
        for (;;) { const double v = std::strtod(p, &e); if (e == p) break; out.push_back(v); p = e; }
    };
    std::vector<double>* cur = nullptr;
    std::string line;
    while (std::getline(f, line)) {
        const size_t b = line.find_first_not_of(" \t\r");
        if (b == std::string::npos || line[b] == '#') continue;
        if (line[b] == '!') {
            const size_t eq = line.find('=', b);
            if (eq == std::string::npos) continue;
            std::string key = line.substr(b + 1, eq - b - 1);
            key.erase(0, key.find_first_not_of(' ')); key.erase(key.find_last_not_of(' ') + 1);
            const std::string val = line.substr(eq + 1);
            cur = nullptr;
            if (key == "Nchains") st.Nchains = std::atoi(val.c_str());
            else if (key == "Nvars") st.Nvars = std::atoi(val.c_str());
            else if (key == "iteration") st.iteration = std::atol(val.c_str());
            else if (key == "variable_names") {
                st.variable_names.clear();
                size_t p = 0;
                while ((p = val.find_first_not_of(' ', p)) != std::string::npos) { const size_t q = val.find(' ', p); st.variable_names.push_back(val.substr(p, q - p)); if (q == std::string::npos) break; p = q; }
            }
            else if (key == "vars") cur = &st.vars;
            else if (key == "vars_mean") cur = &st.vars_mean;
            else if (key == "sigmas") cur = &st.sigmas;
            else if (key == "sigmas_mean") cur = &st.sigmas_mean;
            else if (key == "mus") cur = &st.mus;
            else if (key == "mus_mean") cur = &st.mus_mean;
            else if (key == "covarmats") cur = &st.covarmats;
            else if (key == "covarmats_mean") cur = &st.covarmats_mean;
            if (cur) { cur->clear(); numbers(val, *cur); }
        } else if (line[b] == '*') {
            continue;                                      // chain marker of a covariance block: the rows follow in chain order
        } else if (cur) numbers(line, *cur);
    }
    const size_t nc = (size_t)st.Nchains, nv = (size_t)st.Nvars;
    auto ok = [](const std::vector<double>& v, size_t n) { return v.empty() || v.size() == n; };
    if (!ok(st.vars, nc * nv) || !ok(st.vars_mean, nc * nv) || !ok(st.sigmas, nc) || !ok(st.sigmas_mean, nc) || !ok(st.mus, nc * nv) ||
        !ok(st.mus_mean, nc * nv) || !ok(st.covarmats, nc * nv * nv) || !ok(st.covarmats_mean, nc * nv * nv)) return -2;
    return 0;
}

// Writes the three restore files of `st` in the reference's layout (outputs.cpp:863-1027): every matrix row printed on its own
// (Eigen's `.row(i)`: aligned per row, 6 significant digits).  0 on success, -1 when a file cannot be opened.
inline int write_restore(const std::string& dir, const std::string& star_id, const std::string& phase, const RestoreState& st)
{
    const size_t nc = (size_t)st.Nchains, nv = (size_t)st.Nvars;
    auto row = [&](const std::vector<double>& v, size_t off, size_t n) { return eigen_row(std::vector<double>(v.begin() + (long)off, v.begin() + (long)(off + n)), false) + "\n"; };
    auto rows = [&](const std::vector<double>& v, size_t first, size_t nrows) { std::string t; for (size_t r = 0; r < nrows; r++) t += row(v, (first + r) * nv, nv); return t; };
    auto head = [&](int n, const char* what, const char* a, const char* b) {
        std::string t = "# This is an output file containing what is required to restore a run to its last saved position \n";
        t += "# File number: " + std::to_string(n) + " \n" + what + "# Use this if you wish to: \n";
        t += std::string("#       (1) complete a finished job that requires more samples ==> set erase_old_file=0 and ") + a + " \n";
        t += std::string("#       (2) restart a finished job by ignoring old samples (e.g. ignoring a Burn-in) ==> set erase_old_file=1 and ") + b + " \n";
        t += std::string("#       (3) terminate an unfinished job which failed to finished (e.g. due to computer unexpected shutdown) ==> set erase_old_file=0 and ") + a + " \n";
        t += "! Nchains= " + std::to_string(st.Nchains) + "\n! Nvars= " + std::to_string(st.Nvars) + "\n! iteration=" + std::to_string(st.iteration) + "\n! variable_names=";
        for (const auto& nm : st.variable_names) t += nm + "   ";
        return t + "\n";
    };
    std::string t[3];
    t[0] = head(1, "# Contains the last values for the variables vars[0:Nchain-1]. vars_mean denotes averaged values of Nbuffer \n", "do_restore_[X]=1", "do_restore_proposal=1")
         + "! vars= \n" + rows(st.vars, 0, nc) + "! vars_mean= \n" + rows(st.vars_mean, 0, nc);
    t[1] = head(2, "# Contains the last values of (a) sigmas[0:Nchains-1] and (b) mus[0:Nchains-1, 0:Nvars-1].  sigmas_mean and mus_mean denotes averaged values of Nbuffer\n", "do_restore=1", "do_restore=1")
         + "! sigmas= " + eigen_row(st.sigmas, false) + "\n! mus= \n" + rows(st.mus, 0, nc)
         + "! sigmas_mean= " + eigen_row(st.sigmas_mean, false) + "\n! mus_mean= \n" + rows(st.mus_mean, 0, nc);
    t[2] = head(3, "# Contains the last value of the covariance matrix covarmats[0:Nchains-1, 0:Nvars-1, 0:Nvars-1]. covarmats_mean denotes the averaged values over Nbuffer\n", "do_restore=1", "do_restore=1");
    for (int pass = 0; pass < 2; pass++) {
        const std::vector<double>& C = pass ? st.covarmats_mean : st.covarmats;
        t[2] += pass ? "! covarmats_mean= \n" : "! covarmats= \n";
        for (size_t c = 0; c < nc; c++) t[2] += "*" + std::to_string(c) + "\n" + rows(C, c * nv, nv);
    }
    for (int n = 0; n < 3; n++) {
        std::ofstream f((dir + "/" + star_id + "_restore_" + phase + "_" + std::to_string(n + 1) + ".dat").c_str());
        if (!f.is_open()) return -1;
        f << t[n];
    }
    return 0;
}

inline int read_restore(const std::string& dir, const std::string& star_id, const std::string& phase, RestoreState& st)
{
    for (int n = 1; n <= 3; n++) {
        const int rc = read_restore_file(dir + "/" + star_id + "_restore_" + phase + "_" + std::to_string(n) + ".dat", st);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace outputs
}  // namespace tamcmc
