// tk_expm.cuh -- compressed solve for NonSymInstance: Y_s[:,j] = exp(gamma_j * H) * b~_s with H the k x k upper
// Hessenberg matrix of the Arnoldi process (tensor_krylov_method.jl:10-34; utils.jl:509-511 calls Julia's dense
// exp(::Matrix), a scaling-and-squaring Pade approximant).
//
// Here: scaling and squaring around a degree-16 Taylor polynomial evaluated by Paterson-Stockmeyer, which needs
// only matrix products (6 + s per matrix), so the whole batch (one matrix per exp-sum term j and spectral class)
// runs through one batched FP64 GEMM kernel:
//     s  = max(0, ceil(log2(||gamma H||_1 / theta))),  A = gamma H / 2^s,   theta = 0.8  (||A||^17/17! < 2^-53)
//     A2 = A A, A3 = A2 A, A4 = A2 A2
//     P  = c12 I + c13 A + c14 A2 + c15 A3 + c16 A4
//     P  = (c8  I + c9  A + c10 A2 + c11 A3) + A4 P
//     P  = (c4  I + c5  A + c6  A2 + c7  A3) + A4 P
//     P  = (c0  I + c1  A + c2  A2 + c3  A3) + A4 P          c_i = 1/i!
//     P  = P P   (s times; matrices with fewer squarings sit the remaining passes out)
#pragma once
#include "tk_device.cuh"

namespace tk {

constexpr int EX_SLOTS = 6;          // A, A2, A3, A4, P, Q(ping-pong)
constexpr double EX_THETA = 0.8;

struct ExpmParams {
    int k, ld, t, ncls, ncol;
    long long mslot;                 // ld*ld
    double* W;                       // [ncls*t][EX_SLOTS][mslot]
    int* nsq;                        // [ncls*t] squarings per matrix
    int* where;                      // [ncls*t] slot (4 or 5) holding the finished exponential
    const double* Hd;                // dense H per mode [.. ][ncol*ncol] col-major, or nullptr
    const double* T;                 // tridiagonal part per mode [..][3][ncol] (used when Hd == nullptr)
    const int* cls_mode;             // [ncls] local mode slot whose H defines class c
    const double* alpha;             // [t]
    double lam_inv;
    const int* status;
};

__device__ __forceinline__ double hess_entry(const ExpmParams& p, int slot, int r, int c) {
    if (p.Hd) return p.Hd[(long long)slot * p.ncol * p.ncol + (long long)c * p.ncol + r];
    const double* T = p.T + (long long)slot * 3 * p.ncol;
    if (r == c) return T[r];
    if (r == c + 1) return T[p.ncol + c];       // H[c+2,c+1] (1-based) sub-diagonal
    if (c == r + 1) return T[2 * p.ncol + r];   // H[r+1,r+2] super-diagonal
    return 0.0;
}

// one CTA per matrix m = c*t + j: 1-norm of gamma_j H, scaling, A <- gamma_j H / 2^s
__global__ void __launch_bounds__(256) expm_setup_kernel(ExpmParams p) {
    if (!cta_running(p.status)) return;
    __shared__ double scratch[32];
    const int m = blockIdx.x, c = m / p.t, j = m % p.t, k = p.k, ld = p.ld;
    const int slot = p.cls_mode[c];
    const double gamma = -p.alpha[j] * p.lam_inv;         // tensor_krylov_method.jl:27
    double best = 0.0;
    for (int col = 0; col < k; ++col) {
        double acc = 0.0;
        for (int r = threadIdx.x; r < k; r += blockDim.x) acc += fabs(hess_entry(p, slot, r, col));
        acc = block_sum(acc, scratch);
        best = fmax(best, acc);
    }
    const double nrm = fabs(gamma) * best;
    int s = 0;
    if (nrm > EX_THETA) s = (int)ceil(log2(nrm / EX_THETA));
    s = max(s, 0);
    const double scale = gamma * exp2((double)-s);
    double* A = p.W + (long long)m * EX_SLOTS * p.mslot;
    for (int idx = threadIdx.x; idx < ld * ld; idx += blockDim.x) {
        const int r = idx % ld, col = idx / ld;
        A[idx] = (r < k && col < k) ? scale * hess_entry(p, slot, r, col) : 0.0;
    }
    if (threadIdx.x == 0) {
        p.nsq[m] = s;
        p.where[m] = 4 + (s & 1);                         // squarings ping-pong between slots 4 and 5
    }
}

// Batched C = A*B (+ c0 I + c1 X1 + c2 X2 + c3 X3 + c4 X4), all k x k column-major with leading dimension ld.
// grid = (tiles_n, tiles_m, batch); 256 threads, 64x64 tile, 4x4 per thread, K step 16 through shared memory.
struct GemmJob {
    int a, b, c;              // slots of A, B, C inside a matrix record; for squaring passes a = b = src, c = dst
    int comb;                 // 1: add the Taylor block c0..c4 below
    double c0, c1, c2, c3, c4;
    int sq_step;              // >= 0: squaring pass number; matrix m takes part iff nsq[m] > sq_step, and reads
                              //       slot 4 + (sq_step & 1), writes slot 4 + ((sq_step + 1) & 1)
};

// One 64x64 tile of C = A*B (+ Taylor block), 256 threads, 4x4 outputs per thread, K step 16 through shared memory.
__device__ __forceinline__ void gemm_tile(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                                          const double* rec, long long mslot, int k, int ld, int row0, int col0, int comb,
                                          double c0, double c1, double c2, double c3, double c4, double (*As)[65],
                                          double (*Bs)[65]) {
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;   // 16 x 16 threads, each 4 x 4 outputs
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
    for (int kk = 0; kk < k; kk += 16) {
        // A tile: rows row0..row0+63, cols kk..kk+15  (column-major -> coalesced along rows)
        for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
            const int r = idx % 64, cc = idx / 64;
            const int gr = row0 + r, gc = kk + cc;
            As[cc][r] = (gr < k && gc < k) ? __ldcg(A + (long long)gc * ld + gr) : 0.0;
        }
        // B tile: rows kk..kk+15, cols col0..col0+63
        for (int idx = threadIdx.x; idx < 16 * 64; idx += 256) {
            const int r = idx % 16, cc = idx / 16;
            const int gr = kk + r, gc = col0 + cc;
            Bs[r][cc] = (gr < k && gc < k) ? __ldcg(B + (long long)gc * ld + gr) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[q][tx + 16 * i];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) b[jj] = Bs[q][ty + 16 * jj];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fma(a[i], b[jj], acc[i][jj]);
        }
        __syncthreads();
    }
    const double* X1 = rec;                     // A
    const double* X2 = rec + 1 * mslot;         // A2
    const double* X3 = rec + 2 * mslot;         // A3
    const double* X4 = rec + 3 * mslot;         // A4
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int gr = row0 + tx + 16 * i, gc = col0 + ty + 16 * jj;
            if (gr < k && gc < k) {
                const long long off = (long long)gc * ld + gr;
                double v = acc[i][jj];
                if (comb) {
                    v += c1 * X1[off] + c2 * X2[off] + c3 * X3[off];
                    if (c4 != 0.0) v += c4 * X4[off];
                    if (gr == gc) v += c0;
                }
                C[off] = v;
            }
        }
}

__global__ void __launch_bounds__(256) expm_gemm_kernel(ExpmParams p, GemmJob job) {
    if (!cta_running(p.status)) return;
    const int m = blockIdx.z;
    int sa = job.a, sb = job.b, sc = job.c;
    if (job.sq_step >= 0) {
        if (p.nsq[m] <= job.sq_step) return;
        sa = sb = 4 + (job.sq_step & 1);
        sc = 4 + ((job.sq_step + 1) & 1);
    }
    double* rec = p.W + (long long)m * EX_SLOTS * p.mslot;
    __shared__ double As[16][65];
    __shared__ double Bs[16][65];
    gemm_tile(rec + (long long)sa * p.mslot, rec + (long long)sb * p.mslot, rec + (long long)sc * p.mslot, rec, p.mslot, p.k,
              p.ld, blockIdx.y * 64, blockIdx.x * 64, job.comb, job.c0, job.c1, job.c2, job.c3, job.c4, As, Bs);
}

// ------------------------------------------------------------------------------------------
// The whole exponential of one matrix in ONE launch: a thread-block cluster of TILES x TILES CTAs per matrix, CTA
// (ty, tx) owns the 64 x 64 tile (ty, tx) of every product; the products are separated by cluster barriers (the
// operands live in global memory / L2).  Every matrix does exactly its own number of squarings.  Replaces ~7 + s
// dependent tiny launches per iteration; k <= 64*TILES, cluster sizes 1, 4, 9 and 16 (the last two non-portable).
// ------------------------------------------------------------------------------------------
template <int TILES>
__global__ void __launch_bounds__(256) expm_fused_kernel(ExpmParams p, double f0, double f1, double f2, double f3,
                                                         double f4, double f5, double f6, double f7, double f8, double f9,
                                                         double f10, double f11, double f12, double f13, double f14,
                                                         double f15, double f16) {
    constexpr int CL = TILES * TILES;
    if (CL > 1) {
        // The status word can flip (finalize_body, another stream) while this launch is starting.  The CTAs of a
        // cluster meet at cluster barriers, so they must all take the same run/skip decision: CTA 0 of the cluster
        // reads the word and the others take its copy through distributed shared memory.
        __shared__ int st_sh;
        cg::cluster_group cl = cg::this_cluster();
        if (cl.block_rank() == 0 && threadIdx.x == 0) st_sh = *p.status;
        cl.sync();
        const int st = *cl.map_shared_rank(&st_sh, 0);
        cl.sync();                       // CTA 0 keeps its copy alive until every peer has read it
        if (st != ST_RUNNING) return;
    } else if (!cta_running(p.status)) {
        return;
    }
    __shared__ double As[16][65];
    __shared__ double Bs[16][65];
    __shared__ double scratch[32];
    const int m = blockIdx.x / CL, ct = blockIdx.x % CL;
    const int row0 = (ct / TILES) * 64, col0 = (ct % TILES) * 64;
    const int c = m / p.t, j = m % p.t, k = p.k, ld = p.ld;
    const int slot = p.cls_mode[c];
    double* rec = p.W + (long long)m * EX_SLOTS * p.mslot;
    auto csync = [&]() {
        __threadfence();
        if (CL > 1) cg::this_cluster().sync(); else __syncthreads();
    };
    // ---- scaling: every CTA computes the 1-norm (k^2 cached reads), then writes its own tile of A
    const double gamma = -p.alpha[j] * p.lam_inv;
    double best = 0.0;
    for (int col = 0; col < k; ++col) {
        double acc = 0.0;
        for (int r = threadIdx.x; r < k; r += blockDim.x) acc += fabs(hess_entry(p, slot, r, col));
        acc = block_sum(acc, scratch);
        best = fmax(best, acc);
    }
    const double nrm = fabs(gamma) * best;
    int s = 0;
    if (nrm > EX_THETA) s = max((int)ceil(log2(nrm / EX_THETA)), 0);
    const double scale = gamma * exp2((double)-s);
    for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
        const int r = row0 + idx % 64, col = col0 + idx / 64;
        if (r < ld && col < ld) rec[(long long)col * ld + r] = (r < k && col < k) ? scale * hess_entry(p, slot, r, col) : 0.0;
    }
    if (ct == 0 && threadIdx.x == 0) { p.nsq[m] = s; p.where[m] = 4 + (s & 1); }
    csync();
    auto G = [&](int a, int b, int cdst, int comb, double c0, double c1, double c2, double c3, double c4) {
        gemm_tile(rec + (long long)a * p.mslot, rec + (long long)b * p.mslot, rec + (long long)cdst * p.mslot, rec, p.mslot,
                  k, ld, row0, col0, comb, c0, c1, c2, c3, c4, As, Bs);
    };
    G(0, 0, 1, 0, 0, 0, 0, 0, 0);                 // A2
    csync();
    G(1, 0, 2, 0, 0, 0, 0, 0, 0);                 // A3
    G(1, 1, 3, 0, 0, 0, 0, 0, 0);                 // A4
    csync();
    // top of the Horner scheme: slot 5 <- f12 I + f13 A + f14 A2 + f15 A3 + f16 A4 (own tile only)
    for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
        const int r = row0 + idx % 64, col = col0 + idx / 64;
        if (r < k && col < k) {
            const long long off = (long long)col * ld + r;
            double v = f13 * rec[off] + f14 * rec[p.mslot + off] + f15 * rec[2 * p.mslot + off] + f16 * rec[3 * p.mslot + off];
            if (r == col) v += f12;
            rec[5 * p.mslot + off] = v;
        }
    }
    csync();
    G(3, 5, 4, 1, f8, f9, f10, f11, 0.0);
    csync();
    G(3, 4, 5, 1, f4, f5, f6, f7, 0.0);
    csync();
    G(3, 5, 4, 1, f0, f1, f2, f3, 0.0);
    csync();
    for (int sq = 0; sq < s; ++sq) {
        const int src = 4 + (sq & 1), dst = 4 + ((sq + 1) & 1);
        G(src, src, dst, 0, 0, 0, 0, 0, 0);
        csync();
    }
}

// slot 5 <- c12 I + c13 A + c14 A2 + c15 A3 + c16 A4   (start of the Horner scheme in A4; the three Horner products
// then go 5 -> 4 -> 5 -> 4, leaving the Taylor polynomial in slot 4 where the squarings start)
__global__ void __launch_bounds__(256) expm_top_kernel(ExpmParams p, double c0, double c1, double c2, double c3, double c4) {
    if (!cta_running(p.status)) return;
    const int m = blockIdx.y, k = p.k, ld = p.ld;
    double* rec = p.W + (long long)m * EX_SLOTS * p.mslot;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < ld * ld; idx += gridDim.x * blockDim.x) {
        const int r = idx % ld, c = idx / ld;
        double v = 0.0;
        if (r < k && c < k) {
            v = c1 * rec[idx] + c2 * rec[p.mslot + idx] + c3 * rec[2 * p.mslot + idx] + c4 * rec[3 * p.mslot + idx];
            if (r == c) v += c0;
        }
        rec[5 * p.mslot + idx] = v;
    }
}

// Y_s[r][j] = sum_c E_{cls(s), j}[r, c] * b~_s[c]     (expA * b[s], utils.jl:517); Y row-major [k][tld].
__global__ void __launch_bounds__(256) expm_apply_kernel(ExpmParams p, int per_mode, const double* bt, double* Y,
                                                         long long ystride, int tld) {
    if (!cta_running(p.status)) return;
    extern __shared__ double bsm[];
    const int s = blockIdx.x, j = blockIdx.y, k = p.k, ld = p.ld;
    const int m = (per_mode ? s : 0) * p.t + j;
    const double* E = p.W + (long long)m * EX_SLOTS * p.mslot + (long long)p.where[m] * p.mslot;
    const double* b = bt + (long long)s * p.ncol;
    for (int c = threadIdx.x; c < k; c += blockDim.x) bsm[c] = b[c];
    __syncthreads();
    for (int r = threadIdx.x; r < k; r += blockDim.x) {
        double acc = 0.0;
        for (int c = 0; c < k; ++c) acc = fma(E[(long long)c * ld + r], bsm[c], acc);
        Y[(long long)s * ystride + (long long)r * tld + j] = acc;
    }
}

}  // namespace tk
