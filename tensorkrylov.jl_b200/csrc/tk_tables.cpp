// tk_tables.cpp -- host-side exponential-sum tables and spectral schedule helpers.
//
// Restates (does not copy) the reference's ApproximationData logic:
//   rank selection    src/approximation.jl:44-84  (CSV lookup, kappa rounded DOWN to one digit)
//   coefficient files src/approximation.jl:119-147
//   nonsymmetric sinc quadrature  src/approximation.jl:86-107, 150-158
//   analytic Laplace spectrum     src/eigenvalues.jl:247-265
#include "../../include/tensorkrylov_b200.h"
#include "tk_host.h"

#include <dirent.h>
#include <sys/stat.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

namespace tk {

struct Tables {
    std::vector<double> R;
    std::vector<int> ranks;
    std::vector<double> err;  // [row][rank column]
    std::map<std::tuple<int, int, int>, std::pair<std::vector<double>, std::vector<double>>> coeffs;
    bool loaded = false;
};

static Tables g_tables;

static bool is_dir(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

static int load_packed(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return set_error(TK_ETABLE, "cannot open table file %s", path.c_str());
    char magic[8];
    f.read(magic, 8);
    if (!f || std::memcmp(magic, "TKXSUM01", 8) != 0) return set_error(TK_ETABLE, "%s is not a packed table file", path.c_str());
    int32_t nrows = 0, nranks = 0;
    f.read(reinterpret_cast<char*>(&nrows), 4);
    f.read(reinterpret_cast<char*>(&nranks), 4);
    if (nrows <= 0 || nranks <= 0 || nrows > 100000 || nranks > 4096) return set_error(TK_ETABLE, "corrupt table header");
    Tables t;
    t.R.resize(nrows);
    t.err.resize((size_t)nrows * nranks);
    std::vector<int32_t> ranks(nranks);
    f.read(reinterpret_cast<char*>(t.R.data()), 8 * (size_t)nrows);
    f.read(reinterpret_cast<char*>(t.err.data()), 8 * (size_t)nrows * nranks);
    f.read(reinterpret_cast<char*>(ranks.data()), 4 * (size_t)nranks);
    t.ranks.assign(ranks.begin(), ranks.end());
    int32_t nfiles = 0, pad = 0;
    f.read(reinterpret_cast<char*>(&nfiles), 4);
    f.read(reinterpret_cast<char*>(&pad), 4);
    for (int i = 0; i < nfiles; ++i) {
        int32_t hdr[4];
        f.read(reinterpret_cast<char*>(hdr), 16);
        if (!f || hdr[0] <= 0 || hdr[0] > 4096) return set_error(TK_ETABLE, "corrupt coefficient record %d", i);
        std::vector<double> om(hdr[0]), al(hdr[0]);
        f.read(reinterpret_cast<char*>(om.data()), 8 * (size_t)hdr[0]);
        f.read(reinterpret_cast<char*>(al.data()), 8 * (size_t)hdr[0]);
        t.coeffs[std::make_tuple(hdr[0], hdr[1], hdr[2])] = std::make_pair(std::move(om), std::move(al));
    }
    if (!f) return set_error(TK_ETABLE, "truncated table file %s", path.c_str());
    t.loaded = true;
    g_tables = std::move(t);
    return 0;
}

static double parse_cell(const std::string& s) {
    if (s == "Inf" || s == "inf" || s == "--") return std::numeric_limits<double>::infinity();
    return std::strtod(s.c_str(), nullptr);
}

// The reference's directory: output_data/tabelle_complete.csv plus files 1_xkTT.D_E whose lines are "value {name}".
static int load_dir(const std::string& dir) {
    Tables t;
    std::ifstream csv(dir + "/output_data/tabelle_complete.csv");
    if (!csv) return set_error(TK_ETABLE, "no output_data/tabelle_complete.csv under %s", dir.c_str());
    std::string line;
    bool header = true;
    while (std::getline(csv, line)) {
        if (line.empty()) continue;
        std::vector<std::string> cells;
        std::stringstream ss(line);
        std::string cell;
        while (std::getline(ss, cell, ',')) cells.push_back(cell);
        if (header) {
            for (size_t c = 1; c < cells.size(); ++c) t.ranks.push_back(std::atoi(cells[c].c_str()));
            header = false;
            continue;
        }
        if (cells.size() != t.ranks.size() + 1) return set_error(TK_ETABLE, "ragged CSV row");
        t.R.push_back(std::strtod(cells[0].c_str(), nullptr));
        for (size_t c = 1; c < cells.size(); ++c) t.err.push_back(parse_cell(cells[c]));
    }
    DIR* dp = opendir(dir.c_str());
    if (!dp) return set_error(TK_ETABLE, "cannot list %s", dir.c_str());
    while (dirent* ent = readdir(dp)) {
        int tt = 0, digit = 0, order = 0, used = 0;
        if (std::sscanf(ent->d_name, "1_xk%2d.%d_%d%n", &tt, &digit, &order, &used) != 3) continue;
        if (ent->d_name[used] != '\0') continue;
        std::ifstream f(dir + "/" + ent->d_name);
        std::vector<double> vals;
        while (std::getline(f, line)) {
            size_t a = line.find_first_not_of(" \t\r");
            if (a == std::string::npos) continue;
            vals.push_back(std::strtod(line.c_str() + a, nullptr));
        }
        if ((int)vals.size() < 2 * tt) { closedir(dp); return set_error(TK_ETABLE, "short coefficient file %s", ent->d_name); }
        t.coeffs[std::make_tuple(tt, digit, order)] =
            std::make_pair(std::vector<double>(vals.begin(), vals.begin() + tt), std::vector<double>(vals.begin() + tt, vals.begin() + 2 * tt));
    }
    closedir(dp);
    t.loaded = true;
    g_tables = std::move(t);
    return 0;
}

int tables_sym_lookup(double kappa, double tol, int* t_out, int* digit_out, int* order_out, const double** omega,
                      const double** alpha) {
    if (!g_tables.loaded) return set_error(TK_ETABLE, "exponential-sum tables not loaded (tk_tables_load)");
    if (!(kappa >= 1.0) || !std::isfinite(kappa)) return set_error(TK_EINVAL, "condition number %g is not a finite value >= 1", kappa);
    // parse_condition, approximation.jl:109-116
    int order = (int)std::floor(std::log10(kappa));
    int digit = (int)std::floor(kappa / std::pow(10.0, (double)order));
    int row = -1;
    for (int guard = 0; guard < 1000 && row < 0; ++guard) {   // getclosestrow + the while loop, :56-76
        const double want = digit * std::pow(10.0, (double)order);
        for (size_t r = 0; r < g_tables.R.size(); ++r)
            if (g_tables.R[r] == want) { row = (int)r; break; }
        if (row < 0) digit += 1;
    }
    if (row < 0) return set_error(TK_ETABLE, "condition number %g is outside the table", kappa);
    const size_t nr = g_tables.ranks.size();
    int best = -1;
    for (size_t c = 0; c < nr; ++c)                           // mask = tol .>= row; minimum rank, :79-82
        if (tol >= g_tables.err[(size_t)row * nr + c] && (best < 0 || g_tables.ranks[c] < best)) best = g_tables.ranks[c];
    if (best < 0) return set_error(TK_ETABLE, "no tabulated rank reaches tol %g at R = %d e%d", tol, digit, order);
    auto it = g_tables.coeffs.find(std::make_tuple(best, digit, order));
    if (it == g_tables.coeffs.end()) return set_error(TK_ETABLE, "missing coefficient file 1_xk%02d.%d_%d", best, digit, order);
    *t_out = best; *digit_out = digit; *order_out = order;
    *omega = it->second.first.data();
    *alpha = it->second.second.data();
    return 0;
}

// exponential_sum_parameters! on its own (approximation.jl:119-147): the coefficient file of a GIVEN rank in the
// row the condition number selects, with the tabulated error of that cell.
int tables_sym_rank(double kappa, int rank, const double** omega, const double** alpha, double* err) {
    if (!g_tables.loaded) return set_error(TK_ETABLE, "exponential-sum tables not loaded (tk_tables_load)");
    if (!(kappa >= 1.0) || !std::isfinite(kappa)) return set_error(TK_EINVAL, "condition number %g is not a finite value >= 1", kappa);
    int order = (int)std::floor(std::log10(kappa));
    int digit = (int)std::floor(kappa / std::pow(10.0, (double)order));
    int row = -1;
    for (int guard = 0; guard < 1000 && row < 0; ++guard) {
        const double want = digit * std::pow(10.0, (double)order);
        for (size_t r = 0; r < g_tables.R.size(); ++r)
            if (g_tables.R[r] == want) { row = (int)r; break; }
        if (row < 0) digit += 1;
    }
    if (row < 0) return set_error(TK_ETABLE, "condition number %g is outside the table", kappa);
    auto it = g_tables.coeffs.find(std::make_tuple(rank, digit, order));
    if (it == g_tables.coeffs.end()) return set_error(TK_ETABLE, "no coefficient file 1_xk%02d.%d_%d", rank, digit, order);
    const size_t nr = g_tables.ranks.size();
    if (err) {
        *err = INFINITY;
        for (size_t c = 0; c < nr; ++c)
            if (g_tables.ranks[c] == rank) *err = g_tables.err[(size_t)row * nr + c];
    }
    *omega = it->second.first.data();
    *alpha = it->second.second.data();
    return 0;
}

void laplace_extremes(int d, long long n, int k, double* lmin, double* lmax) {
    // laplace_eigenvalue / analytic_eigenvalues, eigenvalues.jl:247-265
    const double h = 1.0 / (double)(n + 1);
    const double inv_h2 = 1.0 / (h * h);
    const double s1 = std::sin(1.0 * M_PI * (1.0 / (2.0 * (k + 1))));
    const double sk = std::sin((double)k * M_PI * (1.0 / (2.0 * (k + 1))));
    *lmin = 4.0 * inv_h2 * (s1 * s1) * d;
    *lmax = 4.0 * inv_h2 * (sk * sk) * d;
}

int nonsym_coefficients(double lambda_min, double tol, std::vector<double>& omega, std::vector<double>& alpha, int* rank_out) {
    if (!(lambda_min > 0.0) || !(tol > 0.0)) return set_error(TK_EINVAL, "lambda_min and tol must be positive");
    int rank = 1;                                             // compute_rank!, approximation.jl:95-107
    auto bound = [&](int r) { return 2.75 * (1.0 / lambda_min) * std::exp(-M_PI * std::sqrt(r / 2.0)); };
    while (bound(rank) > tol) {
        if (++rank > 100000) return set_error(TK_EINVAL, "nonsymmetric rank bound does not reach tol");
    }
    const double h = M_PI * (1.0 / std::sqrt((double)rank));   // :150-158
    omega.clear(); alpha.clear();
    for (int j = -rank; j <= rank; ++j) {
        alpha.push_back(std::log(std::exp(j * h) + std::sqrt(1.0 + std::exp(2.0 * j * h))));
        omega.push_back(h * (1.0 / std::sqrt(1.0 + std::exp(-2.0 * j * h))));
    }
    *rank_out = rank;
    return 0;
}

}  // namespace tk

extern "C" {

int tk_tables_load(const char* path) {
    if (!path) return tk::set_error(TK_EINVAL, "null path");
    std::string p(path);
    if (tk::is_dir(p)) return tk::load_dir(p);
    return tk::load_packed(p);
}

int tk_tables_sym_lookup(double kappa, double tol, int32_t* t, int32_t* first_digit, int32_t* order, double* omega, double* alpha) {
    int tt, dg, od;
    const double *om, *al;
    int rc = tk::tables_sym_lookup(kappa, tol, &tt, &dg, &od, &om, &al);
    if (rc) return rc;
    if (t) *t = tt;
    if (first_digit) *first_digit = dg;
    if (order) *order = od;
    if (omega) std::memcpy(omega, om, 8 * (size_t)tt);
    if (alpha) std::memcpy(alpha, al, 8 * (size_t)tt);
    return 0;
}

int tk_tables_sym_rank(double kappa, int32_t rank, double* omega, double* alpha, double* err) {
    const double *om, *al;
    int rc = tk::tables_sym_rank(kappa, rank, &om, &al, err);
    if (rc) return rc;
    if (omega) std::memcpy(omega, om, 8 * (size_t)rank);
    if (alpha) std::memcpy(alpha, al, 8 * (size_t)rank);
    return 0;
}

int tk_nonsym_coefficients(double lambda_min, double tol, int32_t cap, int32_t* rank, int32_t* nterms, double* omega, double* alpha) {
    std::vector<double> om, al;
    int r;
    int rc = tk::nonsym_coefficients(lambda_min, tol, om, al, &r);
    if (rc) return rc;
    if (rank) *rank = r;
    if (nterms) *nterms = (int32_t)om.size();
    if ((int)om.size() > cap) return tk::set_error(TK_EINVAL, "need room for %d terms, cap is %d", (int)om.size(), cap);
    if (omega) std::memcpy(omega, om.data(), 8 * om.size());
    if (alpha) std::memcpy(alpha, al.data(), 8 * al.size());
    return 0;
}

int tk_laplace_extremes(int32_t d, int64_t n, int32_t k, double* lambda_min, double* lambda_max) {
    if (d < 1 || n < 1 || k < 1 || !lambda_min || !lambda_max) return tk::set_error(TK_EINVAL, "bad arguments");
    tk::laplace_extremes(d, n, k, lambda_min, lambda_max);
    return 0;
}

}  // extern "C"
