// config.hpp -- the reference's .cfg control file on the C++ driver side (host only, header only).
//
// Line rules of Config::format_line / read_cfg_file (tamcmc/sources/config.cpp:1062-1110, 1223-1500): '!Group:' opens a group,
// '#' starts a comment line, every other line is `key=value; free text` with the value ending at the FIRST ';' (a line without
// one is an error in the reference), numbers are read with strtod (leading number, trailing text ignored), lists are comma
// separated.  tamcmc-c_b200/formats.py:read_cfg is the same reader in Python; tests compare the two.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "mcmc_driver.hpp"

namespace tamcmc {
namespace config {

using Groups = std::map<std::string, std::map<std::string, std::string>>;

inline std::string trim(const std::string& s)
{
    const size_t b = s.find_first_not_of(" \t\r\n");
    if (b == std::string::npos) return "";
    return s.substr(b, s.find_last_not_of(" \t\r\n") - b + 1);
}

// 0 on success; -1: cannot open; >0: number of the first line whose value no ';' terminates (config.cpp:1090-1096)
inline int read_cfg(const std::string& path, Groups& out)
{
    std::ifstream f(path.c_str());
    if (!f.is_open()) return -1;
    std::string line, group;
    int n = 0;
    while (std::getline(f, line)) {
        n++;
        const std::string s = trim(line);
        if (s.empty() || s[0] == '#') continue;
        if (s[0] == '!') { group = trim(s.substr(1, s.find(':') == std::string::npos ? std::string::npos : s.find(':') - 1)); out[group]; continue; }
        if (s == "/END") break;
        const size_t semi = s.find(';');
        if (semi == std::string::npos) return n;
        const std::string body = trim(s.substr(0, semi));
        const size_t eq = body.find('=');
        if (eq == std::string::npos || group.empty()) continue;
        out[group][trim(body.substr(0, eq))] = trim(body.substr(eq + 1));
    }
    return 0;
}

inline double number(const std::string& raw) { return std::strtod(raw.c_str(), nullptr); }
inline std::vector<long> integer_list(const std::string& raw)
{
    std::vector<long> v;
    size_t p = 0;
    while (p <= raw.size()) {
        const size_t q = raw.find(',', p);
        const std::string tok = trim(raw.substr(p, q == std::string::npos ? std::string::npos : q - p));
        if (!tok.empty()) v.push_back((long)number(tok));
        if (q == std::string::npos) break;
        p = q + 1;
    }
    return v;
}

// The !MALA group into a DriverConfig (keys that are absent keep the defaults; `epsilon2` is the reference's MALA.epsi2,
// config.cpp:1276-1279).  false: periods_learn does not have one entry fewer than Nt_learn (config_default.cfg:18).
inline bool apply_mala(const Groups& g, DriverConfig& cfg)
{
    const auto it = g.find("MALA");
    if (it == g.end()) return true;
    const auto& m = it->second;
    auto has = [&](const char* k) { return m.find(k) != m.end(); };
    if (has("Nchains")) cfg.Nchains = (int)number(m.at("Nchains"));
    if (has("lambda_temp")) cfg.lambda_temp = number(m.at("lambda_temp"));
    if (has("c0")) cfg.c0 = number(m.at("c0"));
    if (has("epsilon1")) cfg.epsilon1 = number(m.at("epsilon1"));
    if (has("epsilon2")) cfg.epsi2 = number(m.at("epsilon2"));
    if (has("A1")) cfg.A1 = number(m.at("A1"));
    if (has("target_acceptance")) cfg.target_acceptance = number(m.at("target_acceptance"));
    if (has("dN_mixing")) cfg.dN_mixing = (long)number(m.at("dN_mixing"));
    if (has("Nt_learn")) cfg.Nt_learn = integer_list(m.at("Nt_learn"));
    if (has("periods_learn")) cfg.periods_learn = integer_list(m.at("periods_learn"));
    return cfg.periods_learn.size() + 1 == cfg.Nt_learn.size();
}

}  // namespace config
}  // namespace tamcmc
