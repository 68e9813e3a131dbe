// tk_compress.cuh -- kernels (2) (3) (4): compressed solve and residual estimate.
//
//   tridiag_eig_kernel    H_k = Q diag(theta) Q'            replaces the t dense exp(gamma*Symmetric(H)) calls (utils.jl:509-511)
//   assemble_cp_kernel    Y_s[:,j] = Q exp(gamma_j theta) Q' b~_s   (tensor_krylov_method.jl:10-34, utils.jl:513-521)
//   gram_blocks_kernel    Z_s = H_s Y_s, Y'Y, Y'Z, Z'Z, ||b~_s||^2  (utils.jl:186-204, 229-253, 285-288)
//   combine_chunk_kernel  product over modes in R[e,h]/(e^2,h^2)    (MVnorm utils.jl:280-324, boundary term :428-437,
//   finalize_kernel       tensorinnerprod :332-369, r_comp :393, status :395 and tensor_krylov_method.jl:99-118)
#pragma once
#include "tk_device.cuh"

namespace tk {

// ------------------------------------------------------------------------------------------
// Kernel (2): symmetric tridiagonal eigensolver, implicit-shift QL with Wilkinson shifts.
// grid = (problems, row blocks).  The scalar recurrence that generates the Givens rotations is run
// redundantly by every warp on a private copy of (d, e) in shared memory, so no barrier of any kind is
// needed: a CTA owns `rows_per_blk` rows of Q (shared memory, one row per thread) and applies each
// rotation to its rows as it is generated, carrying the shared column in a register.  Row blocks of
// one problem never talk to each other.  The loads of the next (d, e) pair are issued one rotation
// ahead and 1/sqrt replaces sqrt + division on the dependent chain.
// in : T[prob*tstride + i] = diagonal, T[prob*tstride + ncol + i] = sub-diagonal (H[i+2,i+1])
// out: theta[prob*thstride + i], Q[prob*qstride + i*ldq + r] = component r of eigenvector i
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tridiag_eig_kernel(const double* __restrict__ T, long long tstride, int ncol,
                                                          int k, int rows_per_blk, int ldz, double* theta, int thstride,
                                                          double* Q, long long qstride, int ldq, const int* status,
                                                          int* fail) {
    if (status && *status != ST_RUNNING) return;
    extern __shared__ double smem[];
    const int prob = blockIdx.x;
    const int row0 = blockIdx.y * rows_per_blk;
    const int nrows = min(rows_per_blk, k - row0);
    const int lrow = threadIdx.x, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5, lane = threadIdx.x & 31;
    double* d = smem + (size_t)warp * 2 * k;
    double* e = d + k;
    double* Z = smem + (size_t)nwarp * 2 * k;          // Z[i*ldz + lrow]
    const double* Td = T + (long long)prob * tstride;
    for (int i = lane; i < k; i += 32) {
        d[i] = Td[i];
        e[i] = (i < k - 1) ? Td[ncol + i] : 0.0;
    }
    const bool active = lrow < nrows;
    if (active)
        for (int i = 0; i < k; ++i) Z[(size_t)i * ldz + lrow] = (i == row0 + lrow) ? 1.0 : 0.0;
    __syncwarp();
    const double EPS = 2.220446049250313e-16;
    bool failed = false;
    for (int l = 0; l < k; ++l) {
        int iter = 0;
        while (true) {
            int m = l;
            for (; m < k - 1; ++m) {
                const double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= EPS * dd) break;
            }
            if (m == l) break;
            if (++iter > 80) { failed = true; break; }
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = sqrt(fma(g, g, 1.0));
            g = d[m] - d[l] + e[l] / (g + copysign(r, g));
            double s = 1.0, c = 1.0, pp = 0.0;
            int last = m;
            double zc = active ? Z[(size_t)m * ldz + lrow] : 0.0;
            bool underflow = false;
            double e_i = e[m - 1], d_i = d[m - 1], d_ip1 = d[m];
            for (int i = m - 1; i >= l; --i) {
                // operands of the NEXT rotation: never written during this sweep, so they can be fetched now
                const double e_n = (i > l) ? e[i - 1] : 0.0, d_n = (i > l) ? d[i - 1] : 0.0;
                const double zi = active ? Z[(size_t)i * ldz + lrow] : 0.0;
                const double f = s * e_i, b = c * e_i;
                const double h2 = fma(f, f, g * g);
                if (h2 == 0.0) {
                    e[i + 1] = 0.0;
                    d[i + 1] = d_ip1 - pp;
                    e[m] = 0.0;
                    underflow = true;
                    break;
                }
                const double rinv = rsqrt(h2);
                r = h2 * rinv;
                e[i + 1] = r;
                s = f * rinv;
                c = g * rinv;
                g = d_ip1 - pp;
                r = fma(d_i - g, s, 2.0 * c * b);
                pp = s * r;
                d[i + 1] = g + pp;
                g = fma(c, r, -b);
                if (active) {
                    Z[(size_t)(i + 1) * ldz + lrow] = fma(s, zi, c * zc);
                    zc = fma(c, zi, -s * zc);
                }
                last = i;
                e_i = e_n; d_ip1 = d_i; d_i = d_n;
            }
            if (active) Z[(size_t)last * ldz + lrow] = zc;
            if (underflow) continue;
            d[l] -= pp;
            e[l] = g;
            e[m] = 0.0;
        }
        if (failed) break;
    }
    __syncwarp();
    if (blockIdx.y == 0)
        for (int i = threadIdx.x; i < k; i += blockDim.x) theta[(long long)prob * thstride + i] = d[i];
    if (failed && threadIdx.x == 0 && fail) atomicExch(fail, 1);
    if (active) {
        double* Qg = Q + (long long)prob * qstride + row0 + lrow;
        for (int i = 0; i < k; ++i) Qg[(long long)i * ldq] = Z[(size_t)i * ldz + lrow];
    }
}

struct CompressParams {
    int k, t, tld, ncol;
    int per_mode;            // 0: every mode uses class 0 (reference, utils.jl:509); 1: mode s uses class s
    const double* theta; int thstride;
    const double* Q; long long qstride; int ldq;
    const double* bt;        // [dl][ncol]
    const double* alpha;     // [t]
    const double* omega;     // [t]
    double lam_inv;          // inv(lambda_min)
    double* Y;               // [dl][ystride], row-major k x tld
    double* Z;
    long long ystride;
    const double* T;         // [dl][3][ncol]
    const double* Hd;        // dense Hessenberg (Arnoldi) or nullptr
    double* E;               // [dl][3*t*t]: Y'Y, Y'Z, Z'Z  (entry (i,j) at j*t+i)
    long long estride;
    double* bb;              // [dl] ||b~_s[1:k]||^2
    const int* status;
};

// ------------------------------------------------------------------------------------------
// Kernel (3): Y_s = Q (exp(theta gamma') o c 1'),  c = Q' b~_s[1:k].  One CTA per mode.
// ------------------------------------------------------------------------------------------
constexpr int ASM_TJ = 16;

__global__ void __launch_bounds__(256) assemble_cp_kernel(CompressParams p) {
    if (*p.status != ST_RUNNING) return;
    extern __shared__ double smem[];
    const int s = blockIdx.x, k = p.k, t = p.t;
    const int cls = p.per_mode ? s : 0;
    const double* Q = p.Q + (long long)cls * p.qstride;
    const double* theta = p.theta + (long long)cls * p.thstride;
    const double* bt = p.bt + (long long)s * p.ncol;
    double* csm = smem;            // k
    double* thsm = smem + k;       // k
    double* F = smem + 2 * k;      // k x ASM_TJ
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = warp; i < k; i += nwarp) {
        const double* qi = Q + (long long)i * p.ldq;
        double acc = 0.0;
        for (int r = lane; r < k; r += 32) acc = fma(qi[r], bt[r], acc);
        acc = warp_sum(acc);
        if (lane == 0) { csm[i] = acc; thsm[i] = theta[i]; }
    }
    __syncthreads();
    double* Y = p.Y + (long long)s * p.ystride;
    for (int j0 = 0; j0 < t; j0 += ASM_TJ) {
        const int tj = min(ASM_TJ, t - j0);
        for (int idx = threadIdx.x; idx < k * ASM_TJ; idx += blockDim.x) {
            const int i = idx / ASM_TJ, jj = idx % ASM_TJ;
            double f = 0.0;
            if (jj < tj) {
                const double gamma = -p.alpha[j0 + jj] * p.lam_inv;   // tensor_krylov_method.jl:27
                f = csm[i] * exp(gamma * thsm[i]);
            }
            F[idx] = f;
        }
        __syncthreads();
        for (int r = threadIdx.x; r < k; r += blockDim.x) {
            double acc[ASM_TJ];
#pragma unroll
            for (int jj = 0; jj < ASM_TJ; ++jj) acc[jj] = 0.0;
            for (int i = 0; i < k; ++i) {
                const double q = Q[(long long)i * p.ldq + r];
                const double* Fi = F + i * ASM_TJ;
#pragma unroll
                for (int jj = 0; jj < ASM_TJ; ++jj) acc[jj] = fma(q, Fi[jj], acc[jj]);
            }
            double* Yr = Y + (long long)r * p.tld + j0;
#pragma unroll
            for (int jj = 0; jj < ASM_TJ; ++jj)
                if (jj < tj) Yr[jj] = acc[jj];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Kernel (4a): per-mode blocks of the estimator.  Z = H[1:k,1:k] Y uses the FULL H view, not the
// Symmetric wrapper (utils.jl:247), so the un-symmetric entry H[k-1,k] left by an MGS fallback
// is honoured through T[2].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_blocks_kernel(CompressParams p) {
    if (*p.status != ST_RUNNING) return;
    __shared__ double scratch[32];
    const int s = blockIdx.x, k = p.k, t = p.t, tld = p.tld;
    const double* Y = p.Y + (long long)s * p.ystride;
    double* Z = p.Z + (long long)s * p.ystride;
    const double* T = p.T + (long long)s * 3 * p.ncol;
    if (p.Hd == nullptr) {
        for (int idx = threadIdx.x; idx < k * t; idx += blockDim.x) {
            const int r = idx / t, j = idx % t;
            double acc = 0.0;
            if (r > 0) acc = T[p.ncol + (r - 1)] * Y[(long long)(r - 1) * tld + j];          // H[r,r-1]
            acc = fma(T[r], Y[(long long)r * tld + j], acc);                                 // H[r,r]
            if (r < k - 1) acc = fma(T[2 * p.ncol + r], Y[(long long)(r + 1) * tld + j], acc);  // H[r,r+1]
            Z[(long long)r * tld + j] = acc;
        }
    } else {
        const double* H = p.Hd + (long long)s * p.ncol * p.ncol;
        for (int idx = threadIdx.x; idx < k * t; idx += blockDim.x) {
            const int r = idx / t, j = idx % t;
            double acc = 0.0;
            for (int c = max(0, r - 1); c < k; ++c) acc = fma(H[(long long)c * p.ncol + r], Y[(long long)c * tld + j], acc);
            Z[(long long)r * tld + j] = acc;
        }
    }
    __syncthreads();
    double* E = p.E + (long long)s * p.estride;
    const int tt = t * t;
    for (int pidx = threadIdx.x; pidx < tt; pidx += blockDim.x) {
        const int i = pidx % t, j = pidx / t;
        double l = 0.0, x = 0.0, lz = 0.0;
        for (int r = 0; r < k; ++r) {
            const double yi = Y[(long long)r * tld + i], yj = Y[(long long)r * tld + j];
            const double zi = Z[(long long)r * tld + i], zj = Z[(long long)r * tld + j];
            l = fma(yi, yj, l);
            x = fma(yi, zj, x);     // X[i,j] = (Y'Z)[i,j]
            lz = fma(zi, zj, lz);
        }
        E[pidx] = l;
        E[tt + pidx] = x;
        E[2 * tt + pidx] = lz;
    }
    const double* bt = p.bt + (long long)s * p.ncol;
    double acc = 0.0;
    for (int r = threadIdx.x; r < k; r += blockDim.x) acc = fma(bt[r], bt[r], acc);
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) p.bb[s] = acc;
}

// ------------------------------------------------------------------------------------------
// Kernel (4b): cross-mode combine.  For each (i,j) the per-mode element
//     a_q = Ly_q + e X_q[i,j] + h X_q[j,i] + e h Lz_q[i,j]      in R[e,h]/(e^2,h^2)
// is multiplied over the modes; the e*h coefficient of the product is the (i,j) integrand of
// MVnorm (utils.jl:296-318).  The same product with a second nilpotent g carries the boundary
// term (utils.jl:428-437) and, on the first rows, tensorinnerprod (utils.jl:351-362).
// This replaces the reference's O(d^3 t^2) loops by O(d t^2) and is associative, so a chunk of
// modes -> one partial, and partials merge across chunks and across GPUs.
// partial layout: P0 | Pe | Ph | Peh | Pg (t*t each) | v0 | v1 (t each) | bb | orthS | ...
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) combine_chunk_kernel(CompressParams p, int dl, int chunk_modes, int chunk_base,
                                                            double* partials, long long pstride, const double* orthS,
                                                            int mode0_local) {
    if (*p.status != ST_RUNNING) return;
    const int t = p.t, tt = t * t, k = p.k, tld = p.tld;
    const int q0 = blockIdx.x * chunk_modes - chunk_base, q1 = min(dl, q0 + chunk_modes);
    const int qb = max(q0, 0);
    double* P = partials + (long long)blockIdx.x * pstride;
    for (int pidx = threadIdx.x; pidx < tt; pidx += blockDim.x) {
        const int i = pidx % t, j = pidx / t, pT = i * t + j;
        double P0 = 1.0, Pe = 0.0, Ph = 0.0, Peh = 0.0, Pg = 0.0;
        for (int q = qb; q < q1; ++q) {
            const double* E = p.E + (long long)q * p.estride;
            const double* Yk = p.Y + (long long)q * p.ystride + (long long)(k - 1) * tld;   // delta^q = Y_q[k,:]
            const double hq = p.T[(long long)q * 3 * p.ncol + p.ncol + (k - 1)];             // H_q[k+1,k]
            const double L = E[pidx], Xij = E[tt + pidx], Xji = E[tt + pT], Lz = E[2 * tt + pidx];
            const double gq = (hq * hq) * (Yk[i] * Yk[j]);
            Peh = fma(Peh, L, fma(Pe, Xji, fma(Ph, Xij, P0 * Lz)));
            Pe = fma(Pe, L, P0 * Xij);
            Ph = fma(Ph, L, P0 * Xji);
            Pg = fma(Pg, L, P0 * gq);
            P0 *= L;
        }
        P[pidx] = P0; P[tt + pidx] = Pe; P[2 * tt + pidx] = Ph; P[3 * tt + pidx] = Peh; P[4 * tt + pidx] = Pg;
    }
    for (int i = threadIdx.x; i < t; i += blockDim.x) {
        double v0 = 1.0, v1 = 0.0;
        for (int q = qb; q < q1; ++q) {
            const double y1 = p.Y[(long long)q * p.ystride + i], z1 = p.Z[(long long)q * p.ystride + i];
            v1 = fma(v1, y1, v0 * z1);
            v0 *= y1;
        }
        P[5 * tt + i] = v0;
        P[5 * tt + t + i] = v1;
    }
    if (threadIdx.x == 0) {
        double bb = 1.0;
        for (int q = qb; q < q1; ++q) bb *= p.bb[q];
        P[5 * tt + 2 * t] = bb;
        P[5 * tt + 2 * t + 1] = (mode0_local >= qb && mode0_local < q1) ? orthS[k - 1] : -1.0;
    }
}

struct FinalizeParams {
    int k, t, nmax, nparts, fixed_iterations;
    long long pstride;
    const double* partials;   // [nparts][pstride], merged in order
    const double* omega;
    double lam_inv, lambda_min, tol;
    const double* bnorm;      // device scalar kronprodnorm(b)
    double* relres; double* projres; double* orth;   // [nmax] ConvergenceData vectors
    double* detail;           // [nmax+1][8]
    int* status; long long* niter; int* term_k;
};

__global__ void __launch_bounds__(256) finalize_kernel(FinalizeParams f) {
    if (*f.status != ST_RUNNING) return;
    __shared__ double scratch[32];
    const int t = f.t, tt = t * t;
    double hy2 = 0.0, bnd = 0.0;
    for (int pidx = threadIdx.x; pidx < tt; pidx += blockDim.x) {
        const int i = pidx % t, j = pidx / t;
        if (i < j) continue;                       // lower triangles only (i >= j)
        double P0 = 1.0, Pe = 0.0, Ph = 0.0, Peh = 0.0, Pg = 0.0;
        for (int c = 0; c < f.nparts; ++c) {
            const double* B = f.partials + (long long)c * f.pstride;
            const double B0 = B[pidx], Be = B[tt + pidx], Bh = B[2 * tt + pidx], Beh = B[3 * tt + pidx], Bg = B[4 * tt + pidx];
            Peh = fma(Peh, B0, fma(Pe, Bh, fma(Ph, Be, P0 * Beh)));
            Pe = fma(Pe, B0, P0 * Be);
            Ph = fma(Ph, B0, P0 * Bh);
            Pg = fma(Pg, B0, P0 * Bg);
            P0 *= B0;
        }
        const double li = f.lam_inv * f.omega[i], lj = f.lam_inv * f.omega[j];   // y.lambda, tensor_krylov_method.jl:23
        const double w = (i == j) ? 1.0 : 2.0;
        hy2 = fma(w * (li * lj), Peh, hy2);
        bnd = fma(w * (li * lj), Pg, bnd);
    }
    hy2 = block_sum(hy2, scratch);
    bnd = block_sum(bnd, scratch);
    double ip = 0.0;
    for (int i = threadIdx.x; i < t; i += blockDim.x) {
        double v0 = 1.0, v1 = 0.0;
        for (int c = 0; c < f.nparts; ++c) {
            const double* B = f.partials + (long long)c * f.pstride + 5 * tt;
            v1 = fma(v1, B[i], v0 * B[t + i]);
            v0 *= B[i];
        }
        ip = fma(f.lam_inv * f.omega[i], v1, ip);
    }
    ip = block_sum(ip, scratch);
    if (threadIdx.x == 0) {
        double bb = 1.0, orthS = 0.0;
        for (int c = 0; c < f.nparts; ++c) {
            const double* B = f.partials + (long long)c * f.pstride + 5 * tt + 2 * t;
            bb *= B[0];
            if (B[1] >= 0.0) orthS = B[1];
        }
        const double b_norm = *f.bnorm;
        const double hyb = ip * b_norm;                        // utils.jl:365
        const double r_comp = hy2 - 2.0 * hyb + bb;            // utils.jl:393
        const int k = f.k;
        double r_norm = sqrt(bnd + r_comp);                    // utils.jl:441
        double* D = f.detail + (long long)k * 8;
        D[0] = hy2; D[1] = hyb; D[2] = bb; D[3] = bnd; D[4] = r_comp; D[5] = r_norm; D[6] = (double)t; D[7] = f.lambda_min;
        int st = ST_RUNNING;
        if (r_comp != r_comp || bnd != bnd) {
            st = ST_NAN;
        } else if (r_comp < 0.0 && !f.fixed_iterations) {
            st = ST_BREAKDOWN;                                 // utils.jl:395 -> tensor_krylov_method.jl:85-96
            *f.niter = k - 1;
        } else {
            if (r_comp < 0.0) r_norm = sqrt(fmax(bnd + r_comp, 0.0));
            const double rel = r_norm / b_norm;                // tensor_krylov_method.jl:99
            f.relres[k - 1] = rel;
            f.projres[k - 1] = r_comp;
            f.orth[k - 1] = sqrt(orthS);                       // orthogonality_loss(V_1[:,1:k], k), :103
            if (!f.fixed_iterations && rel < f.tol) st = ST_CONVERGED;
            else if (k == f.nmax) st = ST_NMAX;
        }
        if (st != ST_RUNNING) {
            *f.term_k = k;
            __threadfence();
            *f.status = st;
        }
    }
}

// ------------------------------------------------------------------------------------------
// x.fmat[s] = V_s[:,1:k] Y_s   (basis_tensor_mul!, utils.jl:478-488); n x t column-major out.
// ------------------------------------------------------------------------------------------
constexpr int BM_TJ = 8;
__global__ void __launch_bounds__(256) basis_mul_kernel(const double* __restrict__ V, long long ldv, int n, int k,
                                                        const double* __restrict__ Y, int tld, int t, double* X) {
    extern __shared__ double ysm[];   // k x BM_TJ
    const int j0 = blockIdx.y * BM_TJ, tj = min(BM_TJ, t - j0);
    for (int idx = threadIdx.x; idx < k * BM_TJ; idx += blockDim.x) {
        const int c = idx / BM_TJ, jj = idx % BM_TJ;
        ysm[idx] = (jj < tj) ? Y[(long long)c * tld + j0 + jj] : 0.0;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc[BM_TJ];
#pragma unroll
    for (int jj = 0; jj < BM_TJ; ++jj) acc[jj] = 0.0;
    for (int c = 0; c < k; ++c) {
        const double v = V[(long long)c * ldv + i];
#pragma unroll
        for (int jj = 0; jj < BM_TJ; ++jj) acc[jj] = fma(v, ysm[c * BM_TJ + jj], acc[jj]);
    }
#pragma unroll
    for (int jj = 0; jj < BM_TJ; ++jj)
        if (jj < tj) X[(long long)(j0 + jj) * n + i] = acc[jj];
}

}  // namespace tk
