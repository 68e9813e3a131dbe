// tk_host.h -- host-side helpers shared by the translation units of libtensorkrylov_b200.
#pragma once
#include <vector>

namespace tk {

// stores a printf-formatted message for tk_last_error() and returns `code`
int set_error(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

int tables_sym_lookup(double kappa, double tol, int* t, int* digit, int* order, const double** omega, const double** alpha);
int tables_sym_rank(double kappa, int rank, const double** omega, const double** alpha, double* err);
void laplace_extremes(int d, long long n, int k, double* lmin, double* lmax);
// (min, max) eigenvalue of every leading minor of the ld x ld column-major block `lead`, k = 1..nmax, at out[2k], out[2k+1]
int minor_extremes(const std::vector<double>& lead, int ld, int nmax, int kind, std::vector<double>& out);
int nonsym_coefficients(double lambda_min, double tol, std::vector<double>& omega, std::vector<double>& alpha, int* rank);

}  // namespace tk
