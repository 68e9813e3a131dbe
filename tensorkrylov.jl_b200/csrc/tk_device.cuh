// tk_device.cuh -- shared device-side definitions for libtensorkrylov_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

namespace tk {

namespace cg = cooperative_groups;

constexpr int ST_RUNNING = -1;  // == TK_RUNNING
constexpr int ST_CONVERGED = 0, ST_NMAX = 1, ST_BREAKDOWN = 2, ST_NAN = 3;

constexpr int OP_DIA = 0, OP_CSR = 1, OP_DENSE = 2;
constexpr int MAX_DIAG = 9;
constexpr double SQRT_EPS = 1.4901161193847656e-08;  // sqrt(eps(Float64)), orthogonal_bases.jl:123

// One coefficient matrix A_s in device memory.
//  DIA  : diag[j*ld + i] = A[i, i+offs[j]]  (offs ascending, zero outside the matrix)
//  CSR  : rowptr/colidx (0-based int32), val; column indices ascending inside a row
//  DENSE: column-major n x n
struct OpDesc {
    int type;
    int ndiag;
    int offs[MAX_DIAG];
    long long ld;
    const double* diag;
    const int* rowptr;
    const int* colidx;
    const double* val;
    const double* dense;
    int constd;               // DIA: every diagonal holds one value on its in-matrix part ...
    double cval[MAX_DIAG];    // ... which is cval[j] (Toeplitz operators like the 1D Laplacian)
};

// Krylov state of the modes this GPU owns.  k is the reference's 1-based iteration:
// v_k is column k-1 of V_s; T[s][0][k-1] = H[k,k], T[s][1][k-1] = H[k+1,k], T[s][2][k-1] = H[k,k+1].
struct KrylovParams {
    int n;             // order of A_s
    int ncol;          // nmax + 1
    long long ldv;     // column stride of V (multiple of 16 doubles)
    long long vstride; // mode stride of V = ncol * ldv
    double* V;         // [dl][ncol][ldv]
    const double* b;   // [dl][ldv]
    double* T;         // [dl][3][ncol]   tridiagonal part of H (Lanczos variants)
    double* Hd;        // [dl][ncol*ncol] column-major dense H (Arnoldi), else nullptr
    double* bt;        // [dl][ncol]      compressed right-hand side  b~_s = V_s' b_s
    double* g;         // [dl][ncol]      newest Gram row  g_j = v_j . v_new
    double* S;         // [dl]            running ||V'V - I||_F^2 over the columns built so far
    double* orthS;     // [ncol]          S of global mode 0 after column c (c = index+1)
    int* fallbacks;    // [dl]            number of MGS fallbacks taken (LanczosReorth)
    const OpDesc* ops;
    const int* mode_op;  // [dl]
    const int* status;   // device status word
    int* snap;           // [2] launch-uniform snapshots of the status word for the cluster kernels
    int mode0_local;     // local index of global mode 0, or -1 if another rank owns it
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA, result in every thread.  Fixed association order -> bitwise reproducible.
// scratch: >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = (lane < nw) ? scratch[lane] : 0.0;
    return warp_sum(r);
}

// Two sums at once (same barriers).  scratch: >= 64 doubles.
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { scratch[w] = a; scratch[32 + w] = b; }
    __syncthreads();
    a = warp_sum((lane < nw) ? scratch[lane] : 0.0);
    b = warp_sum((lane < nw) ? scratch[32 + lane] : 0.0);
}

// Row i of A_s times v.  Accumulates in ascending column order, like the reference's CSC
// product (SparseArrays mul!, called at orthogonal_bases.jl:20,45,103).
__device__ __forceinline__ double apply_row(const OpDesc& op, const double* __restrict__ v, int i, int n) {
    double acc = 0.0;
    if (op.type == OP_DIA) {
        for (int j = 0; j < op.ndiag; ++j) {
            const int c = i + op.offs[j];
            if (c >= 0 && c < n) acc = fma(__ldg(op.diag + (long long)j * op.ld + i), v[c], acc);
        }
    } else if (op.type == OP_CSR) {
        const int p1 = __ldg(op.rowptr + i + 1);
        for (int p = __ldg(op.rowptr + i); p < p1; ++p) acc = fma(__ldg(op.val + p), v[__ldg(op.colidx + p)], acc);
    } else {
        const double* a = op.dense + i;
        for (int j = 0; j < n; ++j) acc = fma(__ldg(a + (long long)j * n), v[j], acc);
    }
    return acc;
}

// Run/skip decision of a whole CTA.  finalize flips the status word from another stream while kernels of later
// iterations are in flight, so threads that each read the live word could disagree and part of a CTA would miss the
// barriers below; thread 0 reads it once and every thread takes that reading.
__device__ __forceinline__ bool cta_running(const int* status) {
    __shared__ int st_s;
    if (threadIdx.x == 0) st_s = *reinterpret_cast<const volatile int*>(status);
    __syncthreads();
    return st_s == ST_RUNNING;
}

// Solve-wide control words.  The host writes the pinned copy before it launches the first segment of a solve and
// reset_kernel brings it into device memory, so launches recorded in a CUDA graph do not carry the tolerance or the
// solve counter as kernel arguments.  finalize writes the exit back to the pinned copy.
struct SolveCtl {
    double tol;
    long long epoch;        // solve counter * (nmax + 2): base of the per-iteration flags of the peer exchange
    int status;             // host copy only: exit status once the solve has ended, else ST_RUNNING
    int term_k;
    long long niter;
};

// Rows tid, tid + THREADS, ... of A_s times v for a thread that keeps EPT rows in registers.  Same arithmetic and
// accumulation order as apply_row, but the descriptor is read once and, for DIA operators, the loops over rows and
// diagonals are fully unrolled so every load of the product is in flight at once: with apply_row's data-dependent
// loops the 8 rows x 4 diagonals of the convection-diffusion operator cost 32 serial memory latencies (19 us of a
// 51 us Arnoldi step at C4, ncu source view).
template <int EPT, int THREADS>
__device__ __forceinline__ void apply_rows(const OpDesc& op, const double* __restrict__ v, int n, int tid, double (&out)[EPT]) {
    if (op.type == OP_DIA) {
        const int nd = op.ndiag;
        const long long ld = op.ld;
        const double* __restrict__ diag = op.diag;
        int offs[MAX_DIAG];
#pragma unroll
        for (int j = 0; j < MAX_DIAG; ++j) offs[j] = op.offs[j];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int i = tid + e * THREADS;
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < MAX_DIAG; ++j) {
                const int c = i + offs[j];
                if (j < nd && i < n && c >= 0 && c < n) acc = fma(__ldg(diag + (long long)j * ld + i), v[c], acc);
            }
            out[e] = acc;
        }
    } else {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int i = tid + e * THREADS;
            out[e] = (i < n) ? apply_row(op, v, i, n) : 0.0;
        }
    }
}

// 16-byte streaming load that does not allocate in L1 (the V panel is read once per launch).
__device__ __forceinline__ double2 ld_stream2(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

__device__ __forceinline__ double ld_stream1(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

}  // namespace tk
