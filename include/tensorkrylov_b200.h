/*
 * tensorkrylov_b200.h -- C-ABI of libtensorkrylov_b200.so
 *
 * B200-native (sm_100a CUDA) implementation of ONE path of thbake/TensorKrylov.jl:
 * the tensorized Krylov solve  A x = b,  A = sum_s I x..x A_s x..x I,  b = b_1 x..x b_d
 * (`tensorkrylov!`, src/tensor_krylov_method.jl:36-125, entered through
 * `solve_tensorized_system`, src/system.jl:65-83).
 *
 * The reference is pure Julia and has NO FFI; this header is the boundary a
 * `ccall` wrapper binds (see INTEGRATION.md for the Julia stub).  The cut is
 * around the whole iteration loop: operators, right-hand sides and the
 * exponential-sum schedule go in once, the loop runs on the device with no host
 * arithmetic between iterations, and the ConvergenceData histories plus the
 * Kruskal solution come back once.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer; the
 *     handle owns all device memory (and its NCCL communicator);
 *   - all matrices are Julia layout: column-major Float64; sparse operators are
 *     Julia `SparseMatrixCSC` verbatim (1-based Int64 colptr/rowval);
 *   - mode indices `s` and iteration indices `k` in this header are 0-based for
 *     modes and 1-based for k (k is the reference's loop variable);
 *   - every function returns 0 on success, a negative TK_E* code otherwise;
 *     tk_last_error() holds the message (thread-local);
 *   - a handle is single-owner and blocking, like the single-threaded reference.
 */
#ifndef TENSORKRYLOV_B200_H
#define TENSORKRYLOV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tk_handle tk_handle;

/* Instance tags, src/tensor_struct.jl:83-85 */
enum { TK_SYM = 0, TK_NONSYM = 1 };
/* MatrixGallery tags in declaration order, src/tensor_struct.jl:18-23 */
enum { TK_LAPLACE_DENSE = 0, TK_LAPLACE = 1, TK_CONVDIFF = 2, TK_EIGVALMAT = 3, TK_RANDSPD = 4, TK_GENERIC = 5 };
/* TensorDecomposition tags, src/decompositions.jl:120-176 */
enum { TK_LANCZOS = 0, TK_LANCZOS_REORTH = 1, TK_ARNOLDI = 2 };
/* exit status of tk_solve: the three exits of tensorkrylov! (tensor_krylov_method.jl:85-96, 108-118, 122) + NaN guard */
enum { TK_CONVERGED = 0, TK_NMAX = 1, TK_BREAKDOWN = 2, TK_NAN = 3, TK_RUNNING = -1 };

/* tk_create flags */
enum {
    /* Every mode uses exp(gamma*H_1) of mode 1, as the reference does
     * (utils.jl:509-521, tensor_struct.jl:257-260).  Without this flag each mode
     * exponentiates its own H_s (the mathematically intended variant). */
    TK_FLAG_REFERENCE_H1 = 1,
    /* Benchmark mode: never stop on r_comp<0 or on tol; run exactly nmax-1 iterations. */
    TK_FLAG_FIXED_ITERATIONS = 2,
    /* Record CUDA events around the Krylov-step kernels (read with tk_get_timing). */
    TK_FLAG_TIME_KERNELS = 4,
    /* ... and around every other kernel of the iteration as well. */
    TK_FLAG_TIME_ALL = 8
};

enum {
    TK_EINVAL = -1, TK_ECUDA = -2, TK_ENOMEM = -3, TK_ESTATE = -4, TK_ETABLE = -5, TK_ENCCL = -6, TK_EUNSUPPORTED = -7
};

const char* tk_last_error(void);
int tk_version(void);
int tk_device_count(int* count);

/* ---- exponential-sum tables: replaces ApproximationData / compute_rank! /
 * exponential_sum_parameters! (src/approximation.jl:6-175) and the
 * analytic Laplace spectrum (src/eigenvalues.jl:247-265).  Host-side, no GPU needed. */

/* `path` is either the packed file written by tools/pack_tables.py or the reference's
 * coefficients_data/ directory as shipped (approximation.jl:44-54, 119-147). */
int tk_tables_load(const char* path);
/* approximation.jl:65-84 + 119-147: kappa -> (t, omega[t], alpha[t]); omega/alpha need room for 63. */
int tk_tables_sym_lookup(double kappa, double tol, int32_t* t, int32_t* first_digit, int32_t* order,
                         double* omega, double* alpha);
/* exponential_sum_parameters! alone (approximation.jl:119-147): the coefficient file of a GIVEN rank in the row kappa
 * selects, and the tabulated error of that cell; TK_ETABLE if the cell has no file ('--' in the table). */
int tk_tables_sym_rank(double kappa, int32_t rank, double* omega, double* alpha, double* err);
/* approximation.jl:86-107 + 150-158: writes 2*rank+1 terms; returns TK_EINVAL if cap is too small. */
int tk_nonsym_coefficients(double lambda_min, double tol, int32_t cap, int32_t* rank, int32_t* nterms,
                           double* omega, double* alpha);
/* eigenvalues.jl:247-265 */
int tk_laplace_extremes(int32_t d, int64_t n, int32_t k, double* lambda_min, double* lambda_max);

/* ---- multi-GPU bootstrap (one process per GPU; modes are block-partitioned over ranks).
 * Rank 0 calls tk_comm_unique_id and ships the 128 bytes to the other ranks. */
int tk_comm_unique_id(void* out128);

/* ---- handle lifetime.  Replaces the TensorLanczos/TensorLanczosReorth/TensorArnoldi
 * constructors (decompositions.jl:120-176), except that V_s is n x (nmax+1), not n x (n+1).
 * n[d]: order of every A_s (all equal in this version).  device: CUDA ordinal.
 * rank/world/unique_id: world==1 -> unique_id may be NULL. */
int tk_create(tk_handle** out, int32_t d, const int64_t* n, int32_t nmax, int32_t instance,
              int32_t matrixclass, int32_t variant, int32_t flags, int32_t device,
              int32_t rank, int32_t world, const void* unique_id);
void tk_destroy(tk_handle* h);
/* Destroyed handles leave their device blocks in a process-level cache so the next tk_create of the same shape
 * costs no cudaMalloc (the reference allocates its whole state per call; so does a drop-in).  This frees the cache. */
int tk_release_cache(void);
/* the modes [first, first+count) this rank owns */
int tk_local_modes(const tk_handle* h, int32_t* first, int32_t* count);
/* needed = 1 if this rank wants the operator and right-hand side of global mode s: its own modes, plus mode 0
 * under TK_FLAG_REFERENCE_H1 (every rank advances its own copy of mode 1's Krylov recurrence instead of
 * receiving H_1 over the wire).  Feeding modes a rank does not need is allowed and ignored. */
int tk_needs_mode(const tk_handle* h, int32_t s, int32_t* needed);

/* ---- inputs (global mode index s; calls for modes another rank owns are ignored) */
/* KroneckerMatrix.M[s] as SparseMatrixCSC (tensor_struct.jl:168-212) */
int tk_set_operator_csc(tk_handle* h, int32_t s, int64_t n, const int64_t* colptr, const int64_t* rowval,
                        const double* nzval);
/* dense column-major; uplo = 'L' reads only the lower triangle (Symmetric(.,:L), tensor_struct.jl:77), 'F' full */
int tk_set_operator_dense(tk_handle* h, int32_t s, int64_t n, const double* a, char uplo);
/* the reference aliases one matrix object d times (tensor_struct.jl:208-210) */
int tk_share_operator(tk_handle* h, int32_t s_dst, int32_t s_src);
/* the same for every mode this rank holds at once: KroneckerMatrix{U}(A_1, d) fills all d slots with one object
   (tensor_struct.jl:208-210); one call instead of d - 1 */
int tk_share_operator_all(tk_handle* h, int32_t s_src);
/* b_s as handed to tensorkrylov! (already normalised by TensorizedSystem, system.jl:33-37) */
int tk_set_rhs(tk_handle* h, int32_t s, const double* b, int64_t n);
/* one vector for all local modes: random_rhs, system.jl:5-11 */
int tk_set_rhs_all(tk_handle* h, const double* b, int64_t n);
/* per-iteration exp-sum data, k = 2..nmax: what the two update_data! calls
 * (tensor_krylov_method.jl:72-73) produce.  lambda_min is spectraldata.lambda_min[k]. */
int tk_set_schedule(tk_handle* h, int32_t k, double lambda_min, int32_t t, const double* alpha, const double* omega);
/* fills k = 2..nmax for Laplace / SymInstance from the loaded tables (eigenvalues.jl:335 + approximation.jl:160-168) */
int tk_schedule_laplace(tk_handle* h, double tol);
/* the same for every (instance, matrix class) the reference has an extreme_eigvals method for (eigenvalues.jl:335-350):
 * Laplace analytic; RandSPD, and every NonSymInstance class, from the eigenvalues of the leading k x k minors of the
 * operator fed for mode 0 (computed here on the host threads, cached per operator); EigValMat from its diagonal.
 * Needs the tables (Sym) and mode 0's operator on THIS rank (feed mode 0 to every rank; it is ignored otherwise). */
int tk_schedule(tk_handle* h, double tol);
/* the spectral kernel of tk_schedule on its own: (min, max) eigenvalue of the leading k x k minors, k = 1..nmax, of a
 * column-major ld x ld block; out[2k], out[2k+1] (out holds 2*(nmax+1) doubles; general != 0: minimum only, max = NaN) */
int tk_minor_extremes(const double* lead, int32_t ld, int32_t nmax, int32_t general, double* out);

/* ---- the solve: tensorkrylov! (tensor_krylov_method.jl:36-125).
 * relres/projres/orth: nmax doubles each, written like ConvergenceData
 * (convergence.jl:11-20: entry 1 stays 1.0; entry k is iteration k).  niter is
 * ConvergenceData.niterations: nmax, or k-1 after a breakdown at k; term_k is the
 * iteration the loop left at.  The caller mirrors resize! on breakdown. */
int tk_solve(tk_handle* h, double tol, int32_t* status, int64_t* niter, int32_t* term_k,
             double* relres, double* projres, double* orth);

/* the KruskalTensor x returned on convergence (basis_tensor_mul!, utils.jl:478-488):
 * lambda[t], fmat = V_s[:,1:k] * Y_s, n x t column-major, of the iteration the loop left at.  Valid after
 * TK_CONVERGED, or after any exit when force != 0 (the iterate of the last iteration).  tk_solution_rank gives t;
 * every getter takes the capacity of the caller's buffers (in elements) and refuses buffers that are too small. */
int tk_solution_rank(tk_handle* h, int32_t* t);
int tk_get_solution(tk_handle* h, int32_t s, double* lambda, int32_t lambda_cap, double* fmat, int64_t fmat_cap,
                    int32_t force);
/* all local modes at once: fmat = [count][n*t], mode-major, each block n x t column-major.  One launch computes a
 * chunk of modes while the previous chunk crosses PCIe; a destination from tk_alloc_host is written by DMA directly. */
int tk_get_solution_all(tk_handle* h, double* lambda, int32_t lambda_cap, double* fmat, int64_t fmat_cap, int32_t force);
/* the same into DEVICE memory of the handle's GPU (for callers that keep x on the device): one launch, no copy */
int tk_get_solution_device(tk_handle* h, double* lambda, int32_t lambda_cap, double* fmat_dev, int64_t fmat_cap,
                           int32_t force);
/* page-locked host memory for results (and inputs) that should cross PCIe without a staging copy */
int tk_alloc_host(void** out, int64_t bytes);
int tk_free_host(void* p);

/* ---- test-only single-phase entry points and state readers (parity/debug) */
int tk_begin(tk_handle* h);                 /* orthonormalize!(decomp, b) + initialize_compressed_rhs: k = 1 */
int tk_step_bases(tk_handle* h, int32_t k); /* orthonormalize!(decomp, k) + update_rhs!  (orthogonal_bases.jl:162-180) */
int tk_compress(tk_handle* h, int32_t k);   /* eigensolve + CP assembly = solve_compressed_system */
int tk_residual(tk_handle* h, int32_t k, double tol, double* out8);
        /* out8 = {||Hy||^2, <Hy,b>, ||b~||^2, boundary term, r_comp, r_norm, t, lambda_min} */
int tk_get_H(tk_handle* h, int32_t s, double* H /* (nmax+1)^2 col-major */);
int tk_get_V(tk_handle* h, int32_t s, int32_t col /* 1-based */, double* v /* n */);
int tk_get_bt(tk_handle* h, int32_t s, double* bt /* nmax+1 */);
int tk_get_Y(tk_handle* h, int32_t s, int32_t k, double* Y /* k x t col-major */, int32_t* t);
int tk_get_eig(tk_handle* h, int32_t s, int32_t k, double* theta /* k */, double* Q /* k x k col-major */);
int tk_get_orth_state(tk_handle* h, int32_t s, double* S /* running ||V'V - I||_F^2 */, int32_t* fallbacks);

/* batched symmetric tridiagonal eigensolver on its own (kernel 2): nb problems of order k;
 * diag[nb][k], sub[nb][k-1] -> theta[nb][k], Q[nb][k*k] col-major (may be NULL). */
int tk_tridiag_eig_batched(int32_t device, int32_t nb, int32_t k, const double* diag, const double* sub,
                           double* theta, double* Q, int32_t* fallbacks /* problems redone by the QL fallback, may be NULL */);

/* Enqueueing.  tk_solve cuts the loop into segments of consecutive iterations; the device status word ends the
 * solve (every later kernel returns at once), the host only stops enqueueing, two segments behind.  The second
 * tk_solve of a handle with an unchanged configuration records each segment as a CUDA graph and replays it from
 * then on (environment TK_GRAPH = 0 never, 2 from the first solve).  TK_FLAG_TIME_* solves use stream launches. */

/* which: 0 = 3-term Lanczos step, 1 = orthogonality-monitor Gram row, 2 = Arnoldi/MGS step,
 * 3 = eigensolver, 4 = CP assembly + Gram, 5 = cross-mode combine, 6 = the whole last tk_solve
 * (CUDA events on the handle's stream, first enqueue to last kernel), 7 = from the last tk_timing_mark to the end of
 * the last tk_solve (one device-side window around several solves).  Sums over the last tk_solve. */
/* records the start of a timed region on the handle's stream (read it back with which = 7) */
int tk_timing_mark(tk_handle* h);
int tk_get_timing(tk_handle* h, int32_t which, double* ms_total, int64_t* launches, double* algorithmic_bytes);
int tk_launch_count(tk_handle* h, int64_t* launches);
/* rows k0..k1 of the per-iteration record of the last solve, 8 doubles per iteration:
 * {||Hy||^2, <Hy,b>, ||b~||^2, boundary term, r_comp, r_norm, t, lambda_min} (the terms of utils.jl:393, 441) */
int tk_get_detail(tk_handle* h, int32_t k0, int32_t k1, double* out);
/* how the last tk_solve was enqueued: CUDA graph launches (0 = direct stream launches), host time spent recording
 * graphs, whether the cross-GPU exchange went through peer-mapped memory (1) or NCCL (0), number of segments */
int tk_get_solve_info(tk_handle* h, int32_t* graphs_launched, double* graph_build_ms, int32_t* peer_exchange,
                      int32_t* segments);

#ifdef __cplusplus
}
#endif
#endif
