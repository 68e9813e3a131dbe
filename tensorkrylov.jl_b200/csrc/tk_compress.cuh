// tk_compress.cuh -- kernels (2) (3) (4): compressed solve and residual estimate.
//
//   tridiag_eig_bisect_kernel / tridiag_eig_kernel   H_k = Q diag(theta) Q'   replaces the t dense
//                         exp(gamma*Symmetric(H)) calls (utils.jl:509-511); bisection + twisted factorisation, QL fallback
//   assemble_cp_kernel    Y_s[:,j] = Q exp(gamma_j theta) Q' b~_s   (tensor_krylov_method.jl:10-34, utils.jl:513-521)
//   gram_blocks_kernel    Z_s = H_s Y_s, Y'Y, Y'Z, Z'Z, ||b~_s||^2  (utils.jl:186-204, 229-253, 285-288);
//                         gram_z_kernel / gram_e_kernel: the same over many CTAs per mode (few modes, many terms)
//   combine_chunk_kernel  product over modes in R[e,h]/(e^2,h^2)    (MVnorm utils.jl:280-324, boundary term :428-437,
//                         tensorinnerprod :332-369); its last CTA exchanges the merged partial with the peer GPUs and runs
//   finalize_body         r_comp :393, status :395 and tensor_krylov_method.jl:99-118  (finalize_kernel: NCCL fallback only)
//   basis_mul_all_kernel  x.fmat[s] = V_s[:,1:k] Y_s for all modes in one launch   (basis_tensor_mul!, utils.jl:478-488)
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
                                                          int* fail, const int* need) {
    if (status && !cta_running(status)) return;
    if (need && !need[blockIdx.x]) return;      // fallback role: only the problems the bisection kernel flagged
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

// ------------------------------------------------------------------------------------------
// Kernel (2), primary variant: bisection + twisted factorisation, O(k) latency instead of the O(k^2) dependent
// rotation chain of QL (B200's dependent-FP64 latency makes that chain ~230 ns per rotation).
//   one CTA per problem, thread j <-> eigenvalue j (ascending)
//   1. scale T by a power of two to norm <= 1; Gershgorin interval
//   2. eigenvalue j by (tpe+1)-way multisection on the Sturm count, tpe <= 8 threads per eigenvalue with one
//      interior point each.  The count uses the division-free three-term recurrence
//      p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2} (one dependent FMA per step), rescaled by a power of two
//   3. eigenvector j from the twisted factorisation of T - theta_j I (Parlett-Dhillon): forward LDL' and backward
//      UDU' sweeps (two interleaved division chains), twist at argmin |gamma_i|, then two one-multiply recurrences
//   4. eigenvalues closer than 1e-3 ||T|| are re-orthogonalised against each other (modified Gram-Schmidt inside the
//      cluster, one warp per cluster), as LAPACK's dstein does
//   5. problems with gaps below 1e-9 ||T||, very large clusters or non-finite vectors raise need[prob]; the QL
//      kernel, launched right behind, recomputes exactly those problems (and returns at once for the others)
// scratch / Q: [prob][ldq*ldq]; during the kernel both are used as [component i][eigen index j] planes.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double pow2i(int ex) {   // 2^ex for -1022 <= ex <= 1023
    return __hiloint2double((1023 + ex) << 20, 0);
}

// TPE threads cooperate on one eigenvalue during the multisection (one interior point each: the dependent
// Sturm chains of a pass then run in different lanes/warps instead of back to back in one in-order warp).
template <int MAXT>
__global__ void __launch_bounds__(MAXT) tridiag_eig_bisect_kernel(const double* __restrict__ T, long long tstride,
                                                                  int ncol, int k, int tpe, double* theta, int thstride,
                                                                  double* Q, double* scratch, long long qstride,
                                                                  int ldq, const int* status, int* need) {
    if (status && !cta_running(status)) return;
    extern __shared__ double smem[];
    __shared__ double red[32];
    __shared__ double tile[32][33];
    __shared__ int nclus_s, bad_s;
    const int prob = blockIdx.x, nthr = blockDim.x;
    double* d = smem;             // k   scaled diagonal
    double* e = smem + k;         // k   scaled sub-diagonal (e[k-1] = 0)
    double* e2 = smem + 2 * k;    // k
    double* th = smem + 3 * k;    // k   scaled eigenvalues
    int* cstart = reinterpret_cast<int*>(smem + 4 * k);   // cluster boundaries (pairs)
    const double* Td = T + (long long)prob * tstride;
    double* A = scratch + (long long)prob * qstride;   // plane [i*ldq + j]: D+ then z
    double* B = Q + (long long)prob * qstride;         // plane [i*ldq + j]: D-; finally the transposed output
    if (threadIdx.x == 0) { nclus_s = 0; bad_s = 0; }
    // ---- 1. norm and scaling
    double lo_g = 1e300, hi_g = -1e300;
    for (int i = threadIdx.x; i < k; i += nthr) {
        const double di = Td[i];
        const double el = (i > 0) ? fabs(Td[ncol + i - 1]) : 0.0, er = (i < k - 1) ? fabs(Td[ncol + i]) : 0.0;
        lo_g = fmin(lo_g, di - el - er);
        hi_g = fmax(hi_g, di + el + er);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo_g = fmin(lo_g, __shfl_xor_sync(0xffffffffu, lo_g, o));
        hi_g = fmax(hi_g, __shfl_xor_sync(0xffffffffu, hi_g, o));
    }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = lo_g; }
    __syncthreads();
    lo_g = red[0];
    for (int w = 1; w < (nthr + 31) / 32; ++w) lo_g = fmin(lo_g, red[w]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = hi_g; }
    __syncthreads();
    hi_g = red[0];
    for (int w = 1; w < (nthr + 31) / 32; ++w) hi_g = fmax(hi_g, red[w]);
    const double nrm = fmax(fabs(lo_g), fabs(hi_g));
    int ex = 0;
    if (nrm > 0.0 && nrm < 1e300) ex = ((__double2hiint(nrm) >> 20) & 0x7ff) - 1023 + 1;   // nrm < 2^ex
    ex = max(-1000, min(1000, ex));
    const double sc = pow2i(-ex), unsc = pow2i(ex);
    for (int i = threadIdx.x; i < k; i += nthr) {
        d[i] = Td[i] * sc;
        const double ei = (i < k - 1) ? Td[ncol + i] * sc : 0.0;
        e[i] = ei;
        e2[i] = ei * ei;
    }
    __syncthreads();
    const double EPS = 2.220446049250313e-16;
    // ---- 2. multisection: thread = (eigenvalue jj, interior point m); (tpe+1)-way split per pass
    {
        const int jj = threadIdx.x / tpe, m = threadIdx.x % tpe;
        const bool act = jj < k;
        const int jc = act ? jj : k - 1;
        double lo = lo_g * sc - 4.0 * EPS * k - 1e-300, hi = hi_g * sc + 4.0 * EPS * k + 1e-300;
        // interval shrinks by (tpe+1) per pass; stop at 2^-57 of the initial width
        int passes = 57;
        if (tpe == 2) passes = 36; else if (tpe == 4) passes = 25; else if (tpe == 8) passes = 18;
        for (int pass = 0; pass < passes; ++pass) {
            const double w = (hi - lo) / (double)(tpe + 1);
            const double x = lo + (m + 1) * w;
            double pm = 1.0, p = d[0] - x;
            bool neg = (p < 0.0) || (p == 0.0);
            int cnt = neg ? 1 : 0;
            for (int i = 1; i < k; ++i) {
                const double pn = fma(d[i] - x, p, -(e2[i - 1] * pm));
                const bool ng = (pn < 0.0) || (pn == 0.0 && !neg);
                cnt += (ng != neg) ? 1 : 0;
                neg = ng;
                pm = p;
                p = pn;
                if ((i & 7) == 7) {
                    const double big = fmax(fabs(p), fabs(pm));
                    if (big > 0.0 && big < 1e300) {
                        const int pe = ((__double2hiint(big) >> 20) & 0x7ff) - 1023;
                        if (pe > 64 || pe < -64) {
                            const double f = pow2i(max(-1000, min(1000, -pe)));
                            p *= f; pm *= f;
                        }
                    }
                }
            }
            // eigenvalue jc lies right of every point with cnt <= jc and left of every point with cnt > jc
            double nlo = (cnt <= jc) ? x : lo, nhi = (cnt > jc) ? x : hi;
            for (int o = 1; o < tpe; o <<= 1) {
                nlo = fmax(nlo, __shfl_xor_sync(0xffffffffu, nlo, o));
                nhi = fmin(nhi, __shfl_xor_sync(0xffffffffu, nhi, o));
            }
            lo = nlo; hi = nhi;
        }
        if (act && m == 0) th[jj] = 0.5 * (lo + hi);
    }
    __syncthreads();
    // ---- 3. twisted factorisation: thread j <-> eigenvector j
    const int j = threadIdx.x;
    if (j < k) {
        const double t0 = th[j];
        double dp = d[0] - t0, dm = d[k - 1] - t0;
        if (dp == 0.0) dp = 1e-300;
        if (dm == 0.0) dm = 1e-300;
        A[(long long)0 * ldq + j] = dp;
        B[(long long)(k - 1) * ldq + j] = dm;
        for (int i = 0; i < k - 1; ++i) {
            const int ib = k - 2 - i;
            dp = (d[i + 1] - t0) - e2[i] / dp;
            dm = (d[ib] - t0) - e2[ib] / dm;
            if (dp == 0.0) dp = 1e-300;
            if (dm == 0.0) dm = 1e-300;
            A[(long long)(i + 1) * ldq + j] = dp;
            B[(long long)ib * ldq + j] = dm;
        }
        // twist index r = argmin |gamma_i|, gamma_i = D+_i + D-_i - (d_i - theta); 4 rows in flight
        int r = 0;
        double best = 1e300;
        int i = 0;
        for (; i + 3 < k; i += 4) {
            const double a0 = A[(long long)i * ldq + j], a1 = A[(long long)(i + 1) * ldq + j],
                         a2 = A[(long long)(i + 2) * ldq + j], a3 = A[(long long)(i + 3) * ldq + j];
            const double b0 = B[(long long)i * ldq + j], b1 = B[(long long)(i + 1) * ldq + j],
                         b2 = B[(long long)(i + 2) * ldq + j], b3 = B[(long long)(i + 3) * ldq + j];
            const double g0 = fabs(a0 + b0 - (d[i] - t0)), g1 = fabs(a1 + b1 - (d[i + 1] - t0)),
                         g2 = fabs(a2 + b2 - (d[i + 2] - t0)), g3 = fabs(a3 + b3 - (d[i + 3] - t0));
            if (g0 < best) { best = g0; r = i; }
            if (g1 < best) { best = g1; r = i + 1; }
            if (g2 < best) { best = g2; r = i + 2; }
            if (g3 < best) { best = g3; r = i + 3; }
        }
        for (; i < k; ++i) {
            const double gam = fabs(A[(long long)i * ldq + j] + B[(long long)i * ldq + j] - (d[i] - t0));
            if (gam < best) { best = gam; r = i; }
        }
        // z_r = 1; upward z_i = -(e_i / D+_i) z_{i+1}; downward z_{i+1} = -(e_i / D-_{i+1}) z_i.
        // The divisions do not depend on z: four are issued together, then the four dependent multiplies.
        double z = 1.0, nrm2 = 1.0;
        i = r - 1;
        for (; i - 3 >= 0; i -= 4) {
            const double l0 = e[i] / A[(long long)i * ldq + j], l1 = e[i - 1] / A[(long long)(i - 1) * ldq + j],
                         l2 = e[i - 2] / A[(long long)(i - 2) * ldq + j], l3 = e[i - 3] / A[(long long)(i - 3) * ldq + j];
            const double z0 = -l0 * z, z1 = -l1 * z0, z2 = -l2 * z1, z3 = -l3 * z2;
            A[(long long)i * ldq + j] = z0; A[(long long)(i - 1) * ldq + j] = z1;
            A[(long long)(i - 2) * ldq + j] = z2; A[(long long)(i - 3) * ldq + j] = z3;
            nrm2 += z0 * z0 + z1 * z1 + z2 * z2 + z3 * z3;
            z = z3;
        }
        for (; i >= 0; --i) {
            z = -(e[i] / A[(long long)i * ldq + j]) * z;
            A[(long long)i * ldq + j] = z;
            nrm2 = fma(z, z, nrm2);
        }
        z = 1.0;
        i = r;
        for (; i + 3 < k - 1; i += 4) {
            const double u0 = e[i] / B[(long long)(i + 1) * ldq + j], u1 = e[i + 1] / B[(long long)(i + 2) * ldq + j],
                         u2 = e[i + 2] / B[(long long)(i + 3) * ldq + j], u3 = e[i + 3] / B[(long long)(i + 4) * ldq + j];
            const double z0 = -u0 * z, z1 = -u1 * z0, z2 = -u2 * z1, z3 = -u3 * z2;
            A[(long long)(i + 1) * ldq + j] = z0; A[(long long)(i + 2) * ldq + j] = z1;
            A[(long long)(i + 3) * ldq + j] = z2; A[(long long)(i + 4) * ldq + j] = z3;
            nrm2 += z0 * z0 + z1 * z1 + z2 * z2 + z3 * z3;
            z = z3;
        }
        for (; i < k - 1; ++i) {
            z = -(e[i] / B[(long long)(i + 1) * ldq + j]) * z;
            A[(long long)(i + 1) * ldq + j] = z;
            nrm2 = fma(z, z, nrm2);
        }
        A[(long long)r * ldq + j] = 1.0;
        const double inv = rsqrt(nrm2);
        if (!(nrm2 > 0.0) || !(nrm2 < 1e300) || inv != inv) atomicExch(&bad_s, 1);
        i = 0;
        for (; i + 3 < k; i += 4) {
            const double a0 = A[(long long)i * ldq + j], a1 = A[(long long)(i + 1) * ldq + j],
                         a2 = A[(long long)(i + 2) * ldq + j], a3 = A[(long long)(i + 3) * ldq + j];
            A[(long long)i * ldq + j] = a0 * inv; A[(long long)(i + 1) * ldq + j] = a1 * inv;
            A[(long long)(i + 2) * ldq + j] = a2 * inv; A[(long long)(i + 3) * ldq + j] = a3 * inv;
        }
        for (; i < k; ++i) A[(long long)i * ldq + j] *= inv;
        theta[(long long)prob * thstride + j] = t0 * unsc;
    }
    __threadfence_block();
    __syncthreads();
    // ---- 4. clusters
    if (threadIdx.x == 0) {
        int nc = 0, start = 0;
        for (int c = 1; c <= k; ++c) {
            if (c == k || th[c] - th[c - 1] > 1e-3) {
                if (c - start > 1) {
                    cstart[2 * nc] = start; cstart[2 * nc + 1] = c; ++nc;
                    if (c - start > 48) bad_s = 1;
                }
                start = c;
            } else if (th[c] - th[c - 1] < 1e-9) {
                bad_s = 1;
            }
        }
        nclus_s = nc;
    }
    __syncthreads();
    const int nclus = nclus_s;
    if (nclus > 0 && !bad_s) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = nthr >> 5;
        for (int c = warp; c < nclus; c += nwarp) {
            const int c0 = cstart[2 * c], c1 = cstart[2 * c + 1];
            for (int a = c0 + 1; a < c1; ++a) {
                for (int b = c0; b < a; ++b) {
                    double dot = 0.0;
                    for (int i = lane; i < k; i += 32) dot = fma(A[(long long)i * ldq + b], A[(long long)i * ldq + a], dot);
                    dot = warp_sum(dot);
                    for (int i = lane; i < k; i += 32) A[(long long)i * ldq + a] = fma(-dot, A[(long long)i * ldq + b], A[(long long)i * ldq + a]);
                    __syncwarp();
                }
                double nn = 0.0;
                for (int i = lane; i < k; i += 32) nn = fma(A[(long long)i * ldq + a], A[(long long)i * ldq + a], nn);
                nn = warp_sum(nn);
                const double inv = rsqrt(nn);
                if (!(nn > 1e-8)) { if (lane == 0) atomicExch(&bad_s, 1); }
                for (int i = lane; i < k; i += 32) A[(long long)i * ldq + a] *= inv;
                __syncwarp();
            }
        }
    }
    __threadfence_block();
    __syncthreads();
    if (threadIdx.x == 0 && need) need[prob] = bad_s;
    // ---- transpose into the output layout Q[j*ldq + i]
    const int nt = (k + 31) / 32;
    for (int tt = 0; tt < nt * nt; ++tt) {
        const int ti = tt / nt, tj = tt % nt;      // components ti*32.., eigen indices tj*32..
        for (int idx = threadIdx.x; idx < 1024; idx += nthr) {
            const int r = idx >> 5, c = idx & 31;
            const int i = ti * 32 + r, jj = tj * 32 + c;
            tile[r][c] = (i < k && jj < k) ? A[(long long)i * ldq + jj] : 0.0;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < 1024; idx += nthr) {
            const int r = idx >> 5, c = idx & 31;          // r: eigen index in tile, c: component in tile
            const int jj = tj * 32 + r, i = ti * 32 + c;
            if (i < k && jj < k) B[(long long)jj * ldq + i] = tile[c][r];
        }
        __syncthreads();
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

__device__ void gram_blocks_body(const CompressParams& p, double* scratch);

__global__ void __launch_bounds__(256, 3) assemble_cp_kernel(CompressParams p) {
    if (!cta_running(p.status)) return;
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
    // same CTA, same mode: Z_s = H_s Y_s and the Gram blocks (kernel (4a)) while Y_s is still hot in L1/L2
    __shared__ double scratch_g[32];
    gram_blocks_body(p, scratch_g);
}

// ------------------------------------------------------------------------------------------
// Kernel (4a): per-mode blocks of the estimator.  Z = H[1:k,1:k] Y uses the FULL H view, not the
// Symmetric wrapper (utils.jl:247), so the un-symmetric entry H[k-1,k] left by an MGS fallback
// is honoured through T[2].
// ------------------------------------------------------------------------------------------
__device__ void gram_blocks_body(const CompressParams& p, double* scratch) {
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

__global__ void __launch_bounds__(256) gram_blocks_kernel(CompressParams p) {
    if (!cta_running(p.status)) return;
    __shared__ double scratch[32];
    gram_blocks_body(p, scratch);
}

// The same two phases as separate launches with grid = (modes, tiles): one CTA per mode starves the GPU when there
// are few modes and many exp-sum terms (NonSymInstance at n = 200: d = 5 modes, t = 111 terms, k = 150 -- 12 321
// Gram entries and 16 650 entries of Z per mode on ONE CTA took 0.9 ms per iteration).
__global__ void __launch_bounds__(256) gram_z_kernel(CompressParams p) {
    if (!cta_running(p.status)) return;
    const int s = blockIdx.x, k = p.k, t = p.t, tld = p.tld;
    const double* Y = p.Y + (long long)s * p.ystride;
    double* Z = p.Z + (long long)s * p.ystride;
    const int idx = blockIdx.y * blockDim.x + threadIdx.x;
    if (idx >= k * t) return;
    const int r = idx / t, j = idx % t;
    double acc = 0.0;
    if (p.Hd == nullptr) {
        const double* T = p.T + (long long)s * 3 * p.ncol;
        if (r > 0) acc = T[p.ncol + (r - 1)] * Y[(long long)(r - 1) * tld + j];
        acc = fma(T[r], Y[(long long)r * tld + j], acc);
        if (r < k - 1) acc = fma(T[2 * p.ncol + r], Y[(long long)(r + 1) * tld + j], acc);
    } else {
        const double* H = p.Hd + (long long)s * p.ncol * p.ncol;
        for (int c = max(0, r - 1); c < k; ++c) acc = fma(H[(long long)c * p.ncol + r], Y[(long long)c * tld + j], acc);
    }
    Z[(long long)r * tld + j] = acc;
}

__global__ void __launch_bounds__(256) gram_e_kernel(CompressParams p) {
    if (!cta_running(p.status)) return;
    __shared__ double scratch[32];
    const int s = blockIdx.x, k = p.k, t = p.t, tld = p.tld, tt = t * t;
    const double* Y = p.Y + (long long)s * p.ystride;
    const double* Z = p.Z + (long long)s * p.ystride;
    double* E = p.E + (long long)s * p.estride;
    const int pidx = blockIdx.y * blockDim.x + threadIdx.x;
    if (pidx < tt) {
        const int i = pidx % t, j = pidx / t;
        double l = 0.0, x = 0.0, lz = 0.0;
        for (int r = 0; r < k; ++r) {          // same order as gram_blocks_body: bit-identical blocks
            const double yi = Y[(long long)r * tld + i], yj = Y[(long long)r * tld + j];
            const double zi = Z[(long long)r * tld + i], zj = Z[(long long)r * tld + j];
            l = fma(yi, yj, l);
            x = fma(yi, zj, x);
            lz = fma(zi, zj, lz);
        }
        E[pidx] = l;
        E[tt + pidx] = x;
        E[2 * tt + pidx] = lz;
    }
    if (blockIdx.y == 0) {
        const double* bt = p.bt + (long long)s * p.ncol;
        double acc = 0.0;
        for (int r = threadIdx.x; r < k; r += blockDim.x) acc = fma(bt[r], bt[r], acc);
        acc = block_sum(acc, scratch);
        if (threadIdx.x == 0) p.bb[s] = acc;
    }
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
struct FinalizeParams {
    int k, t, nmax, nparts, fixed_iterations;
    long long pstride;
    const double* partials;   // [nparts][pstride], merged in order
    const double* omega;
    double lam_inv, lambda_min;
    const SolveCtl* ctl;      // device copy: tolerance of this solve
    SolveCtl* hctl;           // pinned host copy: receives the exit (status, term_k, niterations)
    double* relres; double* projres; double* orth;   // [nmax] ConvergenceData vectors
    double* detail;           // [nmax+1][8]
    int* status; long long* niter; int* term_k;
};

// Exchange of the merged partials between the GPUs of one box without a collective launch: every rank stores its
// merged partial straight into every peer's receive buffer (peer-mapped device memory, NVLink) and then publishes
// one flag per peer with system-scope release; the same CTA then acquires the flags of all ranks and runs the final
// merge.  recv[r] is rank r's buffer [2][world][slot_stride] (two iterations deep: a rank can be at most one
// iteration ahead of the slowest one, because its next exchange needs everybody's flag of this one), flag[r] is
// rank r's array of `world` counters; the value published for iteration k of a solve is ctl->epoch + k, which only
// ever grows, so nothing has to be reset between iterations or solves.
constexpr int PX_MAX = 8;
struct PeerExchange {
    int world, rank;
    long long slot_stride;
    double* recv[PX_MAX];
    unsigned long long* flag[PX_MAX];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Final merge over the GPUs' partials, r_comp (utils.jl:393), the exits of the loop body
// (tensor_krylov_method.jl:85-118) and the ConvergenceData entries of iteration k.  Whole CTA.
__device__ void finalize_body(const FinalizeParams& f, double* scratch) {
    const int t = f.t, tt = t * t;
    double hy2 = 0.0, bnd = 0.0;
    for (int pidx = threadIdx.x; pidx < tt; pidx += blockDim.x) {
        const int i = pidx % t, j = pidx / t;
        if (i < j) continue;                       // lower triangles only (i >= j)
        double P0 = 1.0, Pe = 0.0, Ph = 0.0, Peh = 0.0, Pg = 0.0;
        for (int c = 0; c < f.nparts; ++c) {
            const double* B = f.partials + (long long)c * f.pstride;
            const double B0 = __ldcg(B + pidx), Be = __ldcg(B + tt + pidx), Bh = __ldcg(B + 2 * tt + pidx),
                         Beh = __ldcg(B + 3 * tt + pidx), Bg = __ldcg(B + 4 * tt + pidx);
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
            const double b0 = __ldcg(B + i), b1 = __ldcg(B + t + i);
            v1 = fma(v1, b0, v0 * b1);
            v0 *= b0;
        }
        ip = fma(f.lam_inv * f.omega[i], v1, ip);
    }
    ip = block_sum(ip, scratch);
    if (threadIdx.x == 0) {
        double bb = 1.0, orthS = 0.0, bn2 = 1.0;
        for (int c = 0; c < f.nparts; ++c) {
            const double* B = f.partials + (long long)c * f.pstride + 5 * tt + 2 * t;
            bb *= __ldcg(B);
            const double o = __ldcg(B + 1);
            if (o >= 0.0) orthS = o;
            bn2 *= __ldcg(B + 2);
        }
        const double b_norm = sqrt(bn2);                       // kronprodnorm(b), tensor_struct.jl:277-281
        const double hyb = ip * b_norm;                        // utils.jl:365
        const double r_comp = hy2 - 2.0 * hyb + bb;            // utils.jl:393
        const int k = f.k;
        double r_norm = sqrt(bnd + r_comp);                    // utils.jl:441
        double* D = f.detail + (long long)k * 8;
        D[0] = hy2; D[1] = hyb; D[2] = bb; D[3] = bnd; D[4] = r_comp; D[5] = r_norm; D[6] = (double)t; D[7] = f.lambda_min;
        int st = ST_RUNNING;
        long long nit = f.nmax;
        if (r_comp != r_comp || bnd != bnd) {
            st = ST_NAN;
        } else if (r_comp < 0.0 && !f.fixed_iterations) {
            st = ST_BREAKDOWN;                                 // utils.jl:395 -> tensor_krylov_method.jl:85-96
            nit = k - 1;
            *f.niter = nit;
        } else {
            if (r_comp < 0.0) r_norm = sqrt(fmax(bnd + r_comp, 0.0));
            const double rel = r_norm / b_norm;                // tensor_krylov_method.jl:99
            f.relres[k - 1] = rel;
            f.projres[k - 1] = r_comp;
            f.orth[k - 1] = sqrt(orthS);                       // orthogonality_loss(V_1[:,1:k], k), :103
            if (!f.fixed_iterations && rel < f.ctl->tol) st = ST_CONVERGED;
            else if (k == f.nmax) st = ST_NMAX;
        }
        if (st != ST_RUNNING) {
            *f.term_k = k;
            __threadfence();
            *f.status = st;
            if (f.hctl) {                                      // the host polls the pinned copy between segments
                f.hctl->term_k = k;
                f.hctl->niter = nit;
                __threadfence_system();
                *reinterpret_cast<volatile int*>(&f.hctl->status) = st;
            }
        }
    }
}

// The CTA that finishes last (ticket counter) merges this GPU's chunk partials, in chunk order, into ONE partial:
// that is what crosses NVLink (18 KB per GPU at t = 21 instead of one partial per chunk).  With fin_here it then
// goes on to the final merge in the same CTA -- directly on one GPU, after the peer exchange (px.world > 1) on
// several -- so an iteration needs no separate finalize launch and no collective launch.
__global__ void __launch_bounds__(256) combine_chunk_kernel(CompressParams p, int dl, int chunk_modes, int chunk_base,
                                                            double* partials, long long pstride, const double* orthS,
                                                            int mode0_local, const double* bnorm2, unsigned int* ticket,
                                                            double* merged, int fin_here, FinalizeParams f, PeerExchange px) {
    if (!cta_running(p.status)) return;
    __shared__ unsigned int my_ticket;
    __shared__ double scratch[32];
    __shared__ int timed_out;
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
        double bb = 1.0, bn2 = 1.0;
        for (int q = qb; q < q1; ++q) { bb *= p.bb[q]; bn2 *= bnorm2[q]; }
        P[5 * tt + 2 * t] = bb;
        P[5 * tt + 2 * t + 1] = (mode0_local >= qb && mode0_local < q1) ? orthS[k - 1] : -1.0;
        P[5 * tt + 2 * t + 2] = bn2;                    // prod_s b_s.b_s of the chunk, for kronprodnorm(b)
    }
    // ---- last CTA: ordered merge of all chunk partials of this GPU
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) my_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    if (my_ticket != gridDim.x - 1) return;
    __threadfence();
    const int nparts = gridDim.x;
    for (int pidx = threadIdx.x; pidx < tt; pidx += blockDim.x) {
        double P0 = 1.0, Pe = 0.0, Ph = 0.0, Peh = 0.0, Pg = 0.0;
        for (int c = 0; c < nparts; ++c) {
            const double* B = partials + (long long)c * pstride;
            const double B0 = __ldcg(B + pidx), Be = __ldcg(B + tt + pidx), Bh = __ldcg(B + 2 * tt + pidx),
                         Beh = __ldcg(B + 3 * tt + pidx), Bg = __ldcg(B + 4 * tt + pidx);
            Peh = fma(Peh, B0, fma(Pe, Bh, fma(Ph, Be, P0 * Beh)));
            Pe = fma(Pe, B0, P0 * Be);
            Ph = fma(Ph, B0, P0 * Bh);
            Pg = fma(Pg, B0, P0 * Bg);
            P0 *= B0;
        }
        merged[pidx] = P0; merged[tt + pidx] = Pe; merged[2 * tt + pidx] = Ph; merged[3 * tt + pidx] = Peh; merged[4 * tt + pidx] = Pg;
    }
    for (int i = threadIdx.x; i < t; i += blockDim.x) {
        double v0 = 1.0, v1 = 0.0;
        for (int c = 0; c < nparts; ++c) {
            const double* B = partials + (long long)c * pstride + 5 * tt;
            const double b0 = __ldcg(B + i), b1 = __ldcg(B + t + i);
            v1 = fma(v1, b0, v0 * b1);
            v0 *= b0;
        }
        merged[5 * tt + i] = v0;
        merged[5 * tt + t + i] = v1;
    }
    if (threadIdx.x == 0) {
        double bb = 1.0, oS = -1.0, bn2 = 1.0;
        for (int c = 0; c < nparts; ++c) {
            const double* B = partials + (long long)c * pstride + 5 * tt + 2 * t;
            bb *= __ldcg(B);
            const double o = __ldcg(B + 1);
            if (o >= 0.0) oS = o;
            bn2 *= __ldcg(B + 2);
        }
        merged[5 * tt + 2 * t] = bb;
        merged[5 * tt + 2 * t + 1] = oS;
        merged[5 * tt + 2 * t + 2] = bn2;
        *ticket = 0u;
    }
    if (!fin_here) return;
    __syncthreads();
    if (px.world > 1) {
        // ---- peer exchange: my merged partial into slot `rank` of everybody's receive buffer (parity k & 1)
        const unsigned long long want = (unsigned long long)f.ctl->epoch + (unsigned long long)k;
        const long long half = (long long)px.world * px.slot_stride;
        for (int r = 0; r < px.world; ++r) {
            double* dst = px.recv[r] + (long long)(k & 1) * half + (long long)px.rank * px.slot_stride;
            for (long long i = threadIdx.x; i < pstride; i += blockDim.x) dst[i] = merged[i];
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < px.world) st_release_sys(px.flag[threadIdx.x] + px.rank, want);
        if (threadIdx.x == 0) timed_out = 0;
        __syncthreads();
        if (threadIdx.x < px.world) {
            const unsigned long long* mine = px.flag[px.rank] + threadIdx.x;
            const long long t0 = clock64();
            while (ld_acquire_sys(mine) < want) {
                if (clock64() - t0 > 20000000000LL) { timed_out = 1; break; }   // ~10 s: a peer died; do not hang the GPU
                __nanosleep(64);
            }
        }
        __syncthreads();
        if (timed_out) {
            if (threadIdx.x == 0) {
                *f.term_k = k;
                __threadfence();
                *f.status = ST_NAN;
                if (f.hctl) { f.hctl->term_k = k; f.hctl->niter = f.nmax; __threadfence_system(); *reinterpret_cast<volatile int*>(&f.hctl->status) = ST_NAN; }
            }
            return;
        }
        __threadfence_system();
        f.partials = px.recv[px.rank] + (long long)(k & 1) * half;
        f.pstride = px.slot_stride;
        f.nparts = px.world;
    }
    finalize_body(f, scratch);
}

__global__ void __launch_bounds__(256) finalize_kernel(FinalizeParams f) {
    if (!cta_running(f.status)) return;
    __shared__ double scratch[32];
    finalize_body(f, scratch);
}

// ------------------------------------------------------------------------------------------
// x.fmat[s] = V_s[:,1:k] Y_s   (basis_tensor_mul!, utils.jl:478-488) for a range of modes in ONE launch.
// grid = (row tiles, modes, column passes); a thread owns one row of the n x t result and keeps TJ of its columns
// in registers, so with t <= TJ the basis panel V_s[:,1:k] is read exactly once (k*n*8 bytes per mode) and the
// result written once (t*n*8).  Y_s (k x t, a few KB) sits in shared memory and is read as broadcasts.
// X: [mode][t][n], i.e. n x t column-major per mode (Julia's Matrix{Float64}).
// ------------------------------------------------------------------------------------------
template <int TJ>
__global__ void __launch_bounds__(256) basis_mul_all_kernel(const double* __restrict__ V, long long vstride, long long ldv,
                                                            int n, int k, const double* __restrict__ Y, long long ystride,
                                                            int tld, int t, double* __restrict__ X, int mode0) {
    extern __shared__ double ysm[];   // k x TJ
    const int sl = blockIdx.y, s = mode0 + sl;
    const int j0 = blockIdx.z * TJ, tj = min(TJ, t - j0);
    const double* Ys = Y + (long long)s * ystride;
    for (int idx = threadIdx.x; idx < k * TJ; idx += blockDim.x) {
        const int c = idx / TJ, jj = idx % TJ;
        ysm[idx] = (jj < tj) ? Ys[(long long)c * tld + j0 + jj] : 0.0;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* Vs = V + (long long)s * vstride + i;
    double acc[TJ];
#pragma unroll
    for (int jj = 0; jj < TJ; ++jj) acc[jj] = 0.0;
    int c = 0;
    for (; c + 3 < k; c += 4) {          // four basis columns in flight per thread
        const double v0 = ld_stream1(Vs + (long long)c * ldv), v1 = ld_stream1(Vs + (long long)(c + 1) * ldv),
                     v2 = ld_stream1(Vs + (long long)(c + 2) * ldv), v3 = ld_stream1(Vs + (long long)(c + 3) * ldv);
        const double* y0 = ysm + c * TJ;
#pragma unroll
        for (int jj = 0; jj < TJ; ++jj) acc[jj] = fma(v0, y0[jj], acc[jj]);
#pragma unroll
        for (int jj = 0; jj < TJ; ++jj) acc[jj] = fma(v1, y0[TJ + jj], acc[jj]);
#pragma unroll
        for (int jj = 0; jj < TJ; ++jj) acc[jj] = fma(v2, y0[2 * TJ + jj], acc[jj]);
#pragma unroll
        for (int jj = 0; jj < TJ; ++jj) acc[jj] = fma(v3, y0[3 * TJ + jj], acc[jj]);
    }
    for (; c < k; ++c) {
        const double v = ld_stream1(Vs + (long long)c * ldv);
#pragma unroll
        for (int jj = 0; jj < TJ; ++jj) acc[jj] = fma(v, ysm[c * TJ + jj], acc[jj]);
    }
    double* Xs = X + ((long long)sl * t + j0) * n + i;
#pragma unroll
    for (int jj = 0; jj < TJ; ++jj)
        if (jj < tj) Xs[(long long)jj * n] = acc[jj];
}

}  // namespace tk
