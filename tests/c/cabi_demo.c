/* Pure-C client of libtensorkrylov_b200.so: the call sequence a Julia `ccall` wrapper performs, with Julia's data
 * layouts (SparseMatrixCSC with 1-based Int64 indices, Vector{Float64}).  Prints the ConvergenceData vectors so the
 * Python test can compare them with the ctypes path.  Usage: cabi_demo <tables> <d> <n> <nmax> <tol> */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tensorkrylov_b200.h"

#define CHECK(call)                                                         \
    do {                                                                    \
        int rc_ = (call);                                                   \
        if (rc_ != 0) {                                                     \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, tk_last_error()); \
            return 1;                                                       \
        }                                                                   \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const char* tables = argv[1];
    const int d = atoi(argv[2]), n = atoi(argv[3]), nmax = atoi(argv[4]);
    const double tol = atof(argv[5]);

    /* A_s = (n+1)^2 tridiag(-1, 2, -1) as SparseMatrixCSC (tensor_struct.jl:48-57) */
    const double h2 = (double)(n + 1) * (double)(n + 1);
    int64_t* colptr = malloc(sizeof(int64_t) * (n + 1));
    int64_t* rowval = malloc(sizeof(int64_t) * 3 * n);
    double* nzval = malloc(sizeof(double) * 3 * n);
    int64_t nnz = 0;
    for (int j = 0; j < n; ++j) {
        colptr[j] = nnz + 1;
        if (j > 0) { rowval[nnz] = j; nzval[nnz++] = -h2; }
        rowval[nnz] = j + 1; nzval[nnz++] = 2.0 * h2;
        if (j < n - 1) { rowval[nnz] = j + 2; nzval[nnz++] = -h2; }
    }
    colptr[n] = nnz + 1;

    /* b_s: one deterministic vector for all modes, normalised (system.jl:5-11, 33-37) */
    double* b = malloc(sizeof(double) * n);
    double nb = 0.0;
    for (int i = 0; i < n; ++i) { b[i] = 0.5 + 0.5 * sin(1.0 + 0.37 * i); nb += b[i] * b[i]; }
    nb = 1.0 / sqrt(nb);
    for (int i = 0; i < n; ++i) b[i] *= nb;

    int64_t* ns = malloc(sizeof(int64_t) * d);
    for (int s = 0; s < d; ++s) ns[s] = n;
    tk_handle* h = NULL;
    CHECK(tk_tables_load(tables));
    CHECK(tk_create(&h, d, ns, nmax, TK_SYM, TK_LAPLACE, TK_LANCZOS_REORTH, TK_FLAG_REFERENCE_H1, 0, 0, 1, NULL));
    CHECK(tk_set_operator_csc(h, 0, n, colptr, rowval, nzval));
    for (int s = 1; s < d; ++s) CHECK(tk_share_operator(h, s, 0));
    for (int s = 0; s < d; ++s) CHECK(tk_set_rhs(h, s, b, n));
    CHECK(tk_schedule_laplace(h, tol));

    double* relres = malloc(sizeof(double) * nmax);
    double* projres = malloc(sizeof(double) * nmax);
    double* orth = malloc(sizeof(double) * nmax);
    int32_t status = -1, term_k = 0;
    int64_t niter = 0;
    CHECK(tk_solve(h, tol, &status, &niter, &term_k, relres, projres, orth));
    printf("status %d niter %lld term_k %d\n", status, (long long)niter, term_k);
    for (int k = 0; k < nmax; ++k) printf("%d %.17g %.17g %.17g\n", k + 1, relres[k], projres[k], orth[k]);
    int32_t t = 0;
    CHECK(tk_solution_rank(h, &t));
    double* lambda = malloc(sizeof(double) * (t > 0 ? t : 1));
    double* fmat = malloc(sizeof(double) * n * (t > 0 ? t : 1));
    CHECK(tk_get_solution(h, d - 1, lambda, t, fmat, (int64_t)n * t, 1));
    printf("solution t %d lambda0 %.17g f00 %.17g\n", t, lambda[0], fmat[0]);
    tk_destroy(h);
    CHECK(tk_release_cache());
    return 0;
}
