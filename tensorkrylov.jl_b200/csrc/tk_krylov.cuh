// tk_krylov.cuh -- kernel family (1): one launch advances the Krylov basis of every mode.
//
//   reset_kernel          status words, ConvergenceData = ones, control words (start of a solve)
//   init_basis_kernel     V[:,1] = b/||b||, b~[1]                (decompositions.jl:112-118, utils.jl:456-464)
//   lanczos_ttr_kernel    3-term recurrence step                (orthogonal_bases.jl:39-67)
//   lanczos_ttr_bulk_kernel   the same step for banded operators, inputs fetched by cp.async.bulk (TMA)
//   gram_row_kernel       g_j = v_j . v_{k+1}, j = 1..k+1       (the only NEW row of V'V; orthogonal_bases.jl:119,250-257)
//   monitor_body          loss test + MGS fallback, run by the last gram_row CTA of a mode (orthogonal_bases.jl:119-131)
//   arnoldi_bgs_kernel    two-sweep Gram-Schmidt step, 4 columns per reduction round (orthogonal_bases.jl:15-37)
//   arnoldi_mgs_reg_kernel / arnoldi_mgs_kernel   the same step in strict MGS order (long modes)
#pragma once
#include "tk_device.cuh"

namespace tk {

// ------------------------------------------------------------------------------------------
// k = 0: normalise b into the first basis vector and start b~.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) init_basis_kernel(KrylovParams p, double* bnorm2) {
    __shared__ double scratch[32];
    const int s = blockIdx.x, n = p.n;
    const double* b = p.b + (long long)s * p.ldv;
    double* v1 = p.V + (long long)s * p.vstride;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fma(b[i], b[i], acc);
    const double bb = block_sum(acc, scratch);
    const double inv = 1.0 / sqrt(bb);                 // inv(norm(b)) .* b
    acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = inv * b[i];
        v1[i] = x;
        acc = fma(x, b[i], acc);
    }
    const double bt0 = block_sum(acc, scratch);
    if (threadIdx.x == 0) {
        p.bt[(long long)s * p.ncol] = bt0;             // b~_s[1] = v_1 . b_s
        bnorm2[s] = bb;                                // for kronprodnorm(b), tensor_struct.jl:271-281
        p.S[s] = 0.0;
        p.fallbacks[s] = 0;
    }
}

// ------------------------------------------------------------------------------------------
// Start of a solve: status words, ConvergenceData(nmax) = ones (convergence.jl:11-20), the ticket counters (a solve
// that ended early may have left kernels half-skipped), and the solve-wide control words from the pinned host copy.
// One launch instead of a dozen small copies, and recordable in a CUDA graph.
// ------------------------------------------------------------------------------------------
struct ResetParams {
    int* status4;                 // [0] live status word, [1..2] snapshots of the cluster kernels
    int* term_k; int* eigfail; long long* niter;
    double* relres; double* projres; double* orth; int nmax;
    unsigned int* tickets; int ntickets; unsigned int* ticket1;
    SolveCtl* ctl; const SolveCtl* hctl;
};

__global__ void __launch_bounds__(256) reset_kernel(ResetParams r) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int i = tid; i < r.nmax; i += nthr) { r.relres[i] = 1.0; r.projres[i] = 1.0; r.orth[i] = 1.0; }
    for (int i = tid; i < r.ntickets; i += nthr) r.tickets[i] = 0u;
    if (tid == 0) {
        r.status4[0] = ST_RUNNING; r.status4[1] = ST_RUNNING; r.status4[2] = ST_RUNNING; r.status4[3] = 0;
        *r.term_k = 0; *r.eigfail = 0; *r.niter = r.nmax;
        if (r.ticket1) *r.ticket1 = 0u;
        r.ctl->tol = r.hctl->tol;
        r.ctl->epoch = r.hctl->epoch;
    }
}

// ------------------------------------------------------------------------------------------
// 3-term Lanczos step k for every mode.  CPM CTAs (one thread-block cluster) share a mode and
// split the rows; the three reductions go through distributed shared memory.
//   u = A v_k - beta_{k-1} v_{k-1};  H[k,k] = u.v_k;  v^ = u - H[k,k] v_k;  beta = ||v^||
//   v_{k+1} = v^ / beta (zeros if beta == 0);  H[k+1,k] = H[k,k+1] = beta;  b~[k+1] = v_{k+1}.b = (v^.b)/beta
// Algorithmic HBM bytes per mode: (ndiag + 4) * 8 * n  (diagonals, v_k, v_{k-1}, b read; v_{k+1} written).
// ------------------------------------------------------------------------------------------
// The status word is flipped by finalize_body (the last CTA of the combine kernel) on another stream while 3-term steps of later iterations are in
// flight.  The CTAs of a cluster exchange partial sums through each other's shared memory, so all of them must take
// the same run/skip decision: they read a snapshot that only the 3-term kernels themselves write (launch k reads
// slot k&1 and refreshes slot (k+1)&1 for the next launch on the same stream), never the live word.
__device__ __forceinline__ bool ttr_running(const KrylovParams& p, int k) {
    const int st = p.snap[k & 1];
    if (blockIdx.x == 0 && threadIdx.x == 0) p.snap[(k + 1) & 1] = (st != ST_RUNNING) ? st : *p.status;
    return st == ST_RUNNING;
}

template <int CPM>
__device__ __forceinline__ double cluster_sum(double blockval, double* slot) {
    if (CPM == 1) return blockval;
    cg::cluster_group cl = cg::this_cluster();
    if (threadIdx.x == 0) *slot = blockval;
    cl.sync();
    double tot = 0.0;
#pragma unroll
    for (int r = 0; r < CPM; ++r) tot += *cl.map_shared_rank(slot, r);
    return tot;
}

// DIA row with a compile-time number of diagonals: lets the compiler unroll and batch the loads of the row loop
template <int ND>
__device__ __forceinline__ double apply_row_dia(const OpDesc& op, const double* __restrict__ v, int i, int n) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
        const int c = i + op.offs[j];
        if (c >= 0 && c < n) acc = fma(__ldg(op.diag + (long long)j * op.ld + i), v[c], acc);
    }
    return acc;
}

template <int CPM, int ND>   // ND > 0: every operator is DIA with exactly ND diagonals
__global__ void __launch_bounds__(512) lanczos_ttr_kernel(KrylovParams p, int k) {
    if (!ttr_running(p, k)) return;
    extern __shared__ double smem[];
    __shared__ double scratch[64];
    __shared__ double slots[4];
    const int s = blockIdx.x / CPM, part = blockIdx.x % CPM, n = p.n;
    const int chunk = (((n + CPM - 1) / CPM) + 1) & ~1;
    const int lo = part * chunk, hi = min(n, lo + chunk);
    double* u = smem;  // rows [lo, hi)
    const OpDesc& op = p.ops[p.mode_op[s]];
    double* Vs = p.V + (long long)s * p.vstride;
    const double* vk = Vs + (long long)(k - 1) * p.ldv;
    const double* vkm1 = (k >= 2) ? Vs + (long long)(k - 2) * p.ldv : nullptr;
    double* vnew = Vs + (long long)k * p.ldv;
    double* T = p.T + (long long)s * 3 * p.ncol;
    const double beta_prev = (k >= 2) ? T[2 * p.ncol + (k - 2)] : 0.0;  // H[k-1,k]  (decompositions.jl:78)

    double acc = 0.0;
    if (ND > 0) {
#pragma unroll 2
        for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            double ui = apply_row_dia<(ND > 0 ? ND : 1)>(op, vk, i, n);
            if (vkm1) ui -= beta_prev * vkm1[i];
            u[i - lo] = ui;
            acc = fma(ui, vk[i], acc);
        }
    } else {
        for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            double ui = apply_row(op, vk, i, n);
            if (vkm1) ui -= beta_prev * vkm1[i];
            u[i - lo] = ui;
            acc = fma(ui, vk[i], acc);
        }
    }
    const double alpha = cluster_sum<CPM>(block_sum(acc, scratch), &slots[0]);

    // second and last reduction round: ||v^||^2 and v^.b together (b~[k+1] = v_{k+1}.b = (v^.b)/beta)
    const double* b = p.b + (long long)s * p.ldv;
    double accb = 0.0;
    acc = 0.0;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double w = u[i - lo] - alpha * vk[i];
        u[i - lo] = w;
        acc = fma(w, w, acc);
        accb = fma(w, b[i], accb);
    }
    block_sum2(acc, accb, scratch);
    double beta2 = acc, vb = accb;
    if (CPM > 1) {
        cg::cluster_group cl = cg::this_cluster();
        if (threadIdx.x == 0) { slots[1] = acc; slots[2] = accb; }
        cl.sync();
        beta2 = 0.0; vb = 0.0;
#pragma unroll
        for (int r = 0; r < CPM; ++r) {
            const double* rs = cl.map_shared_rank(slots, r);
            beta2 += rs[1];
            vb += rs[2];
        }
    }
    const double beta = sqrt(beta2);
    const double inv = (beta == 0.0) ? 0.0 : 1.0 / beta;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) vnew[i] = inv * u[i - lo];
    const double btn = inv * vb;
    if (part == 0 && threadIdx.x == 0) {
        T[k - 1] = alpha;                   // H[k,k]
        T[p.ncol + (k - 1)] = beta;         // H[k+1,k]
        T[2 * p.ncol + (k - 1)] = beta;     // H[k,k+1]   update_subdiagonals!, decompositions.jl:180-186
        p.bt[(long long)s * p.ncol + k] = btn;
    }
    if (CPM > 1) cg::this_cluster().sync();  // keep every CTA's slots alive until all peers have read them
}

// ------------------------------------------------------------------------------------------
// Bulk-copy variant of the 3-term step for banded operators (DIA, every |offset| <= TTR_HALO).
// The slices of v_k (with a halo), v_{k-1} and -- WITHB only -- b a CTA needs are fetched by cp.async.bulk copies that
// complete on one mbarrier, so a CTA has its whole input (16 or 24 bytes per row) in flight from its first
// instruction without holding registers for it; the three passes then run out of shared memory and the only
// global traffic left in them is the diagonals (L2-resident, or none at all when every diagonal is constant:
// CONSTD) and the store of v_{k+1}.  Same thread-to-row map and reduction order as lanczos_ttr_kernel, so for
// equal (CPM, blockDim) the two kernels give bit-identical results.
// ------------------------------------------------------------------------------------------
constexpr int TTR_HALO = 2;   // doubles on each side of the v_k slice: keeps every copy 16-byte aligned

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// Wait for phase `parity` of an mbarrier; traps instead of hanging if the copies never land.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_addr(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

constexpr int TTR_RPT = 10;   // rows per thread of the bulk kernel (its u / v^ slice lives in registers)

template <int CPM, int ND, bool CONSTD, int RPT, bool WITHB>
__global__ void __launch_bounds__(512) lanczos_ttr_bulk_kernel(KrylovParams p, int k) {
    // WITHB = false: b~_s[k+1] = v_{k+1} . b_s is left to the Gram-row kernel that follows (it has v_{k+1} . v_1 and
    // v_1 = b_s / |b_s|): this kernel then moves 24 instead of 32 bytes per row and keeps two slices instead of three
    // in shared memory (40 KB at 2 500 rows: 4 CTAs per SM instead of 3).
    // the snapshot is only consumed after the bulk loads have been issued (and have landed): its latency is off the
    // critical path, and a skipped launch merely fetches slices it does not use
    const bool running = ttr_running(p, k);
    extern __shared__ __align__(16) double smem[];
    __shared__ double scratch[64];
    __shared__ double slots[2 + 2 * CPM];                     // [2+2r..] second-round partials pushed by rank r
    __shared__ double aslots[CPM];                            // first-round partials pushed by rank r
    __shared__ __align__(8) uint64_t bar;
    const int s = blockIdx.x / CPM, part = blockIdx.x % CPM, n = p.n;
    const int chunk = (((n + CPM - 1) / CPM) + 1) & ~1;
    const int lo = part * chunk, hi = min(n, lo + chunk);     // host guarantees lo < n for every part
    const double* vks = smem + TTR_HALO;                      // row i of v_k at vks[i - lo]
    const double* us = smem + chunk + 2 * TTR_HALO;           // v_{k-1} slice
    const double* bs = us + chunk;
    const OpDesc& op = p.ops[p.mode_op[s]];
    double* Vs = p.V + (long long)s * p.vstride;
    const double* vk = Vs + (long long)(k - 1) * p.ldv;
    const double* vkm1 = (k >= 2) ? Vs + (long long)(k - 2) * p.ldv : nullptr;
    const double* b = p.b + (long long)s * p.ldv;
    double* vnew = Vs + (long long)k * p.ldv;
    double* T = p.T + (long long)s * 3 * p.ncol;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // ranges in rows, even at both ends (ldv is a multiple of 16 and columns are padded up to it)
        const int he = (int)min((long long)((hi + 1) & ~1), p.ldv);
        const int g0 = max(lo - TTR_HALO, 0), g1 = (int)min((long long)he + TTR_HALO, p.ldv);
        const uint32_t bytes_v = (uint32_t)(g1 - g0) * 8u, bytes_s = (uint32_t)(he - lo) * 8u;
        const uint32_t total = bytes_v + bytes_s * ((vkm1 ? 1u : 0u) + (WITHB ? 1u : 0u));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(&bar)), "r"(total) : "memory");
        bulk_load(smem + (g0 - lo + TTR_HALO), vk + g0, bytes_v, &bar);
        if (vkm1) bulk_load(smem + chunk + 2 * TTR_HALO, vkm1 + lo, bytes_s, &bar);
        if (WITHB) bulk_load(smem + 2 * chunk + 2 * TTR_HALO, b + lo, bytes_s, &bar);
    }
    // A CTA may only touch a peer's shared memory once that peer has started: arrive on the cluster barrier now,
    // wait for it right before the first remote store (by then every peer has long arrived: no time is spent there).
    if (CPM > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    int offs[ND];
    double cval[ND];
#pragma unroll
    for (int j = 0; j < ND; ++j) { offs[j] = op.offs[j]; cval[j] = CONSTD ? op.cval[j] : 0.0; }
    const double beta_prev = (k >= 2) ? T[2 * p.ncol + (k - 2)] : 0.0;  // H[k-1,k]  (decompositions.jl:78)
    __syncthreads();                                          // the barrier is initialised before anyone polls it
    mbar_wait(&bar, 0);
    if (CPM > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    if (!running) return;

    // u = A v_k - beta_{k-1} v_{k-1} for rows lo + threadIdx.x + r * blockDim.x, kept in registers
    double u[RPT];
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int li = threadIdx.x + r * blockDim.x, i = lo + li;
        double ui = 0.0;
        if (i < hi) {
#pragma unroll
            for (int j = 0; j < ND; ++j) {
                const int c = i + offs[j];
                if ((unsigned)c < (unsigned)n)
                    ui = fma(CONSTD ? cval[j] : __ldg(op.diag + (long long)j * op.ld + i), vks[li + offs[j]], ui);
            }
            if (vkm1) ui -= beta_prev * us[li];
            acc = fma(ui, vks[li], acc);
        }
        u[r] = ui;
    }
    // first reduction round, pushed like the second one: every CTA stores its partial into all peers, so after the
    // barrier each CTA reads only its own shared memory (no remote-load latency behind the barrier); summed in rank order
    double alpha = block_sum(acc, scratch);
    if (CPM > 1) {
        cg::cluster_group cl = cg::this_cluster();
        if (threadIdx.x < CPM) *(cl.map_shared_rank(aslots, threadIdx.x) + part) = alpha;
        cl.sync();
        alpha = 0.0;
#pragma unroll
        for (int r = 0; r < CPM; ++r) alpha += aslots[r];
    }

    double accb = 0.0;
    acc = 0.0;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int li = threadIdx.x + r * blockDim.x;
        if (lo + li < hi) {
            const double w = u[r] - alpha * vks[li];
            u[r] = w;
            acc = fma(w, w, acc);
            if (WITHB) accb = fma(w, bs[li], accb);
        }
    }
    if (WITHB) block_sum2(acc, accb, scratch);
    else acc = block_sum(acc, scratch);
    double beta2 = acc, vb = accb;
    if (CPM > 1) {
        // push the two partials into every CTA of the cluster; after the barrier each CTA only reads its own
        // shared memory, so no CTA can exit while a peer still needs it
        cg::cluster_group cl = cg::this_cluster();
        if (threadIdx.x < CPM) {
            double* dst = cl.map_shared_rank(slots, threadIdx.x) + 2 + 2 * part;
            dst[0] = acc; dst[1] = accb;
        }
        cl.sync();
        beta2 = 0.0; vb = 0.0;
#pragma unroll
        for (int r = 0; r < CPM; ++r) { beta2 += slots[2 + 2 * r]; vb += slots[3 + 2 * r]; }
    }
    const double beta = sqrt(beta2);
    const double inv = (beta == 0.0) ? 0.0 : 1.0 / beta;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int li = threadIdx.x + r * blockDim.x;
        if (lo + li < hi) vnew[lo + li] = inv * u[r];
    }
    const double btn = inv * vb;
    if (part == 0 && threadIdx.x == 0) {
        T[k - 1] = alpha;                   // H[k,k]
        T[p.ncol + (k - 1)] = beta;         // H[k+1,k]
        T[2 * p.ncol + (k - 1)] = beta;     // H[k,k+1]   update_subdiagonals!, decompositions.jl:180-186
        if (WITHB) p.bt[(long long)s * p.ncol + k] = btn;
    }
}

// ------------------------------------------------------------------------------------------
// CTA-wide modified Gram-Schmidt step k (orthogonal_bases.jl:15-37) for mode s.
// v: working vector of n doubles (shared or global scratch); every thread owns the rows
// i = tid, tid + blockDim, ... so the dot/axpy sequence needs no barrier besides the reduction.
// hcol[0..k-1] = H[1:k,k], hcol[k] = H[k+1,k].  Writes V[:,k+1] and b~[k+1].
// ------------------------------------------------------------------------------------------
__device__ void mgs_step_cta(const KrylovParams& p, int s, int k, double* v, double* hcol, double* scratch) {
    const int n = p.n;
    const OpDesc& op = p.ops[p.mode_op[s]];
    double* Vs = p.V + (long long)s * p.vstride;
    const double* vk = Vs + (long long)(k - 1) * p.ldv;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] = apply_row(op, vk, i, n);
    for (int pass = 0; pass < 2; ++pass) {
        for (int c = 0; c < k; ++c) {
            const double* col = Vs + (long long)c * p.ldv;
            double acc = 0.0;
            for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fma(v[i], col[i], acc);
            const double h = block_sum(acc, scratch);
            if (threadIdx.x == 0) hcol[c] = pass ? hcol[c] + h : h;
            for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] = fma(-h, col[i], v[i]);
        }
    }
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fma(v[i], v[i], acc);
    const double beta = sqrt(block_sum(acc, scratch));
    const double inv = 1.0 / beta;  // no zero-norm guard in the reference (:35-36)
    double* vnew = Vs + (long long)k * p.ldv;
    const double* b = p.b + (long long)s * p.ldv;
    acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = v[i] * inv;
        vnew[i] = x;
        v[i] = x;
        acc = fma(x, b[i], acc);
    }
    const double btn = block_sum(acc, scratch);
    if (threadIdx.x == 0) {
        hcol[k] = beta;
        p.bt[(long long)s * p.ncol + k] = btn;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Fold the new Gram row into S = ||V'V - I||_F^2 and, for TensorLanczosReorth, run the MGS fallback when
// sqrt(S) > sqrt(eps) (orthogonal_bases.jl:119-131).  newcol = 0-based index of the newest column (= k for step k).
// Executed by one whole CTA: the gram_row CTA of the mode that finishes last (ticket counter).
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void monitor_body(const KrylovParams& p, int s, int newcol, int reorth, double* hcol, double* v,
                                          double* scratch) {
    const int n = p.n, k = newcol;  // k = reference step index
    double* g = p.g + (long long)s * p.ncol;
    double acc = 0.0;
    for (int j = threadIdx.x; j < newcol; j += blockDim.x) { const double gj = __ldcg(g + j); acc = fma(gj, gj, acc); }
    double off2 = block_sum(acc, scratch);
    double dd = __ldcg(g + newcol) - 1.0;
    double Snew = p.S[s] + 2.0 * off2 + dd * dd;
    if (reorth && sqrt(Snew) > SQRT_EPS) {
        mgs_step_cta(p, s, k, v, hcol, scratch);
        double* T = p.T + (long long)s * 3 * p.ncol;
        if (threadIdx.x == 0) {
            T[k - 1] = hcol[k - 1];                             // H[k,k]
            if (k >= 2) T[2 * p.ncol + (k - 2)] = hcol[k - 2];  // H[k-1,k] keeps the MGS value; H[1:k-2,k] .= 0 (:129)
            T[p.ncol + (k - 1)] = hcol[k];                      // beta = H[k+1,k] (:127)
            T[2 * p.ncol + (k - 1)] = hcol[k];                  // update_subdiagonals! (:137)
            p.fallbacks[s] += 1;
        }
        // Gram row of the replaced column (v holds the new v_{k+1})
        const double* Vs = p.V + (long long)s * p.vstride;
        off2 = 0.0;
        for (int j = 0; j <= newcol; ++j) {
            const double* col = Vs + (long long)j * p.ldv;
            acc = 0.0;
            for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fma(v[i], col[i], acc);
            const double gj = block_sum(acc, scratch);
            if (threadIdx.x == 0) g[j] = gj;
            if (j < newcol) off2 = fma(gj, gj, off2);
            else dd = gj - 1.0;
        }
        Snew = p.S[s] + 2.0 * off2 + dd * dd;
    }
    if (threadIdx.x == 0) {
        p.S[s] = Snew;
        if (s == p.mode0_local) p.orthS[newcol] = Snew;
    }
}

// ------------------------------------------------------------------------------------------
// Orthogonality monitor: the newest row of the Gram matrix V'V.
//   g[s][j] = V_s[:,j] . V_s[:,ncols-1],  j = 0..ncols-1
// The reference forms the whole (k+1)x(k+1) Gram matrix with dgemm at every step in every mode
// (orthogonal_bases.jl:119); only this row is new, the rest is carried in S[s].
// grid = (column chunks, modes); the warps of a CTA stream two columns at a time (16-byte loads, 8 in flight per
// lane) against the new vector.  HBM-bound: 8*n bytes per column.
// The new vector (80 KB at n = 10^4) is read through L1 (w_in_smem = 0: the column streams bypass L1 with
// .L1::no_allocate, so it stays resident after the first pass) or staged in shared memory first (w_in_smem = 1).
// L1 is the default from 64 modes per GPU on: a staging phase is a synchronised burst of L2 reads during which the
// HBM streams of the whole wave stand still (measured, chunks grid: 0.92 -> 0.96 of the HBM peak at 1 024 modes,
// 0.75 -> 0.79 at 128; with 32 modes the staged form is faster).
// ------------------------------------------------------------------------------------------
constexpr int GRAM_PSTRIDE = 16;   // partial sums per column (>= warps per column)

template <int U, int THREADS>   // U = 16-byte loads in flight per lane and column (two columns are streamed at once)
__global__ void __launch_bounds__(THREADS, (THREADS == 256 ? 2 : 1)) gram_row_kernel(KrylovParams p, int ncols, int cols_per_cta, int mode_base,
                                                       int w_in_smem, int wpc, int monitor /* -1 none, 0, 1 = reorth */,
                                                       unsigned int* tickets, double* vscratch, int bt_from_g) {
    // bt_from_g = 1: b~_s[ncols] (1-based) = v_ncols . b_s is taken from this row's first entry: v_1 = b_s / b~_s[1]
    // (init_basis_kernel), so v_ncols . b_s = b~_s[1] (v_ncols . v_1) up to one rounding of a quantity that is itself
    // rounding noise (the columns are orthogonal).  The bulk 3-term kernel then never reads b_s (see there).
    // wpc = warps that share one column: every column is cut into wpc contiguous segments so all warps of the
    // CTA stream equal amounts, whatever the number of columns.
    if (!cta_running(p.status)) return;
    extern __shared__ double smem[];
    const int s = mode_base + blockIdx.y, n = p.n;
    const int c0 = blockIdx.x * cols_per_cta, c1 = min(ncols, c0 + cols_per_cta);
    const double* Vs = p.V + (long long)s * p.vstride;
    const double* wg = Vs + (long long)(ncols - 1) * p.ldv;
    const int nq = n >> 1;
    __shared__ double scratch[32];
    __shared__ unsigned int my_ticket;
    double* part = smem;                                  // [cols_per_cta][GRAM_PSTRIDE] partial sums
    double* hcol = smem + (size_t)cols_per_cta * GRAM_PSTRIDE;              // ncol doubles (MGS fallback)
    double* wsm = hcol + ((p.ncol + 1) & ~1);
    if (w_in_smem) {
        const double2* w2g = reinterpret_cast<const double2*>(wg);
        double2* s2 = reinterpret_cast<double2*>(wsm);
        int q = threadIdx.x;
        for (; q + 3 * (int)blockDim.x < nq; q += 4 * blockDim.x) {
            const double2 t0 = w2g[q], t1 = w2g[q + blockDim.x], t2 = w2g[q + 2 * blockDim.x], t3 = w2g[q + 3 * blockDim.x];
            s2[q] = t0; s2[q + blockDim.x] = t1; s2[q + 2 * blockDim.x] = t2; s2[q + 3 * blockDim.x] = t3;
        }
        for (; q < nq; q += blockDim.x) s2[q] = w2g[q];
        if ((n & 1) && threadIdx.x == 0) wsm[n - 1] = wg[n - 1];
        __syncthreads();
    }
    const double* w = w_in_smem ? wsm : wg;
    const double2* w2 = reinterpret_cast<const double2*>(w);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int seg = warp % wpc, group = warp / wpc, ngroups = nwarp / wpc;
    const int seglen = (((nq + wpc - 1) / wpc) + 31) & ~31;
    const int q0 = seg * seglen, q1 = min(nq, q0 + seglen);
    for (int j = c0 + group; j < c1; j += 2 * ngroups) {
        const int jb = j + ngroups;
        const bool two = jb < c1;
        const double2* ca = reinterpret_cast<const double2*>(Vs + (long long)j * p.ldv);
        const double2* cb = reinterpret_cast<const double2*>(Vs + (long long)(two ? jb : j) * p.ldv);
        double a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { a[u] = 0.0; b[u] = 0.0; }
        int q = q0 + lane;
        for (; q + 32 * (U - 1) < q1; q += 32 * U) {
            double2 x[U], z[U];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = ld_stream2(ca + q + 32 * u);
#pragma unroll
            for (int u = 0; u < U; ++u) z[u] = ld_stream2(cb + q + 32 * u);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double2 y = w2[q + 32 * u];
                a[u] = fma(x[u].x, y.x, a[u]); a[u] = fma(x[u].y, y.y, a[u]);
                b[u] = fma(z[u].x, y.x, b[u]); b[u] = fma(z[u].y, y.y, b[u]);
            }
        }
        for (; q < q1; q += 32) {
            const double2 x0 = ld_stream2(ca + q), z0 = ld_stream2(cb + q);
            const double2 y0 = w2[q];
            a[0] = fma(x0.x, y0.x, a[0]); a[0] = fma(x0.y, y0.y, a[0]);
            b[0] = fma(z0.x, y0.x, b[0]); b[0] = fma(z0.y, y0.y, b[0]);
        }
        if ((n & 1) && seg == wpc - 1 && lane == 0) {
            a[0] = fma(Vs[(long long)j * p.ldv + n - 1], w[n - 1], a[0]);
            if (two) b[0] = fma(Vs[(long long)jb * p.ldv + n - 1], w[n - 1], b[0]);
        }
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) { sa += a[u]; sb += b[u]; }
        sa = warp_sum(sa);
        sb = warp_sum(sb);
        if (lane == 0) {
            part[(j - c0) * GRAM_PSTRIDE + seg] = sa;
            if (two) part[(jb - c0) * GRAM_PSTRIDE + seg] = sb;
        }
    }
    __syncthreads();
    double* g = p.g + (long long)s * p.ncol;
    for (int j = c0 + threadIdx.x; j < c1; j += blockDim.x) {
        double acc = 0.0;
        for (int sgi = 0; sgi < wpc; ++sgi) acc += part[(j - c0) * GRAM_PSTRIDE + sgi];
        g[j] = acc;
        if (bt_from_g && j == 0) p.bt[(long long)s * p.ncol + (ncols - 1)] = acc * p.bt[(long long)s * p.ncol];
    }
    if (monitor < 0) return;
    // the CTA of this mode that finishes last folds the row into S (and runs the MGS fallback if needed)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) my_ticket = atomicAdd(tickets + s, 1u);
    __syncthreads();
    if (my_ticket != gridDim.x - 1) return;
    __threadfence();
    double* v = w_in_smem ? wsm : vscratch + (long long)s * p.ldv;
    monitor_body(p, s, ncols - 1, monitor, hcol, v, scratch);
    if (threadIdx.x == 0) tickets[s] = 0u;
}

// ------------------------------------------------------------------------------------------
// Register-resident Arnoldi step: same two-pass MGS in the reference's order, but the working vector lives in
// registers (EPT rows per thread), each basis column is fetched ONCE per pass (the dot and the update use the same
// registers), the next PD columns are already in flight while the current one is being reduced, and the block reduction
// needs a single barrier (two alternating scratch rows).  The step is a chain of 2k dependent reductions, so the
// latency of one reduction (~0.3 us) is what bounds it, not bandwidth.
// ------------------------------------------------------------------------------------------
template <int EPT, int THREADS, int PD>   // PD: basis columns kept in flight ahead of the one being reduced (0, 1 or 3)
__global__ void __launch_bounds__(THREADS) arnoldi_mgs_reg_kernel(KrylovParams p, int k) {
    if (!cta_running(p.status)) return;
    extern __shared__ double hcol[];                    // k+1
    __shared__ double scr[2][32];
    const int s = blockIdx.x, n = p.n, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    constexpr int NB = PD + 1;                          // register sets for basis columns
    const OpDesc& op = p.ops[p.mode_op[s]];
    double* Vs = p.V + (long long)s * p.vstride;
    const double* vk = Vs + (long long)(k - 1) * p.ldv;
    double v[EPT], colr[NB][EPT];
    apply_rows<EPT, THREADS>(op, vk, n, tid, v);
    int buf = 0;
    auto reduce = [&](double x) -> double {
        x = warp_sum(x);
        if (lane == 0) scr[buf][warp] = x;
        __syncthreads();
        double r = (lane < NW) ? scr[buf][lane] : 0.0;
        buf ^= 1;
        return warp_sum(r);
    };
    for (int pass = 0; pass < 2; ++pass) {
        // software pipeline: columns c .. c+PD are resident in registers when column c is reduced
#pragma unroll
        for (int q = 0; q < PD; ++q) {
            const double* col = Vs + (long long)min(q, k - 1) * p.ldv;
#pragma unroll
            for (int e = 0; e < EPT; ++e) { const int i = tid + e * THREADS; colr[q][e] = (i < n) ? col[i] : 0.0; }
        }
        for (int c0 = 0; c0 < k; c0 += NB) {
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const int c = c0 + q;
                if (c < k) {                                            // uniform across the CTA
                    // fetch column c+PD into the set that was freed by column c-1 (or the spare set at q = 0)
                    const int cf = c + PD;
                    {
                        const int slot = (q + PD) % NB;
                        const double* col = Vs + (long long)min(cf, k - 1) * p.ldv;
#pragma unroll
                        for (int e = 0; e < EPT; ++e) { const int i = tid + e * THREADS; colr[slot][e] = (i < n) ? col[i] : 0.0; }
                    }
                    double acc = 0.0;
#pragma unroll
                    for (int e = 0; e < EPT; ++e) acc = fma(v[e], colr[q][e], acc);
                    const double h = reduce(acc);
                    if (tid == 0) hcol[c] = pass ? hcol[c] + h : h;
#pragma unroll
                    for (int e = 0; e < EPT; ++e) v[e] = fma(-h, colr[q][e], v[e]);
                }
            }
        }
    }
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) acc = fma(v[e], v[e], acc);
    const double beta = sqrt(reduce(acc));
    const double inv = 1.0 / beta;
    double* vnew = Vs + (long long)k * p.ldv;
    const double* b = p.b + (long long)s * p.ldv;
    acc = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int i = tid + e * THREADS;
        if (i < n) {
            const double x = v[e] * inv;
            vnew[i] = x;
            acc = fma(x, b[i], acc);
        }
    }
    const double btn = reduce(acc);
    __syncthreads();
    double* Hs = p.Hd + (long long)s * p.ncol * p.ncol + (long long)(k - 1) * p.ncol;
    for (int c = tid; c < k; c += THREADS) Hs[c] = hcol[c];
    if (tid == 0) {
        Hs[k] = beta;
        p.bt[(long long)s * p.ncol + k] = btn;
        double* T = p.T + (long long)s * 3 * p.ncol;
        T[k - 1] = hcol[k - 1];
        T[p.ncol + (k - 1)] = beta;
    }
}

// ------------------------------------------------------------------------------------------
// Blocked form of the Arnoldi step.  Strict MGS is a chain of 2k dependent CTA-wide reductions (~0.35 us each on
// B200: two shuffle trees, a barrier and the dependent FP64 adds), which bounds arnoldi_mgs_reg_kernel at ~0.3 of
// the HBM roofline however fast the columns arrive.  Here the basis columns are taken B at a time: the B projections
// of a block are computed against the SAME working vector in one reduction round (classical Gram-Schmidt inside the
// block), then subtracted in column order; blocks follow each other as in MGS, and the whole sweep runs twice like
// the reference's (orthogonal_bases.jl:22-33).  Against strict MGS the projection on column i of a block differs by
// sum_{j<i in block} h_j (v_j . v_i): the basis of TensorArnoldi has gone through two passes at every step, so
// v_j . v_i = O(eps) and the difference is O(eps ||A||) per entry of H -- rounding level, far inside the 1e-11 parity
// bound (tests/test_gpu_parity.py::test_arnoldi_steps, test_gpu_baseline_sizes.py C4).  The MGS fallback of
// TensorLanczosReorth, whose basis HAS lost orthogonality when it fires, keeps the strict order (mgs_step_cta).
// The chain is 2k/B reductions, the next block is in flight while the current one is reduced: HBM-bound.
// ------------------------------------------------------------------------------------------
template <int EPT, int THREADS, int B>
__global__ void __launch_bounds__(THREADS) arnoldi_bgs_kernel(KrylovParams p, int k) {
    if (!cta_running(p.status)) return;
    extern __shared__ double hcol[];                    // k+1
    __shared__ double scr[2][B][32];
    const int s = blockIdx.x, n = p.n, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    const OpDesc& op = p.ops[p.mode_op[s]];
    double* Vs = p.V + (long long)s * p.vstride;
    const double* vk = Vs + (long long)(k - 1) * p.ldv;
    double v[EPT], ca[B][EPT], cb[B][EPT];
    apply_rows<EPT, THREADS>(op, vk, n, tid, v);
    int buf = 0;
    auto load_block = [&](double (&dst)[B][EPT], int c0) {
#pragma unroll
        for (int q = 0; q < B; ++q) {
            const double* col = Vs + (long long)min(c0 + q, k - 1) * p.ldv;
#pragma unroll
            for (int e = 0; e < EPT; ++e) { const int i = tid + e * THREADS; dst[q][e] = (i < n) ? col[i] : 0.0; }
        }
    };
    auto process = [&](double (&blk)[B][EPT], int c0, int pass) {
        double h[B];
#pragma unroll
        for (int q = 0; q < B; ++q) {
            double a0 = 0.0, a1 = 0.0;                  // two chains per column: halves the dependent-FMA latency
#pragma unroll
            for (int e = 0; e < EPT; e += 2) {
                a0 = fma(v[e], blk[q][e], a0);
                if (e + 1 < EPT) a1 = fma(v[e + 1], blk[q][e + 1], a1);
            }
            h[q] = a0 + a1;
        }
        if (B == 4) {
            // The four sums are reduced TOGETHER: each butterfly step halves the number of values a lane carries
            // (lanes trade the columns they give up), so the warp stage costs 6 shuffles instead of 20 and ends with the
            // total of column c in the 8 lanes c*8..c*8+7; the CTA stage reduces the NW warp partials of a column inside
            // that group of lanes (3 shuffles) and hands the four totals to every lane (4 shuffles).  The sweep is bound
            // by instruction issue (ncu: one CTA per mode, 8 warps per SM), and the shuffle trees were a third of it.
            const bool hi = (lane & 16) != 0, h8 = (lane & 8) != 0;
            double k0 = hi ? h[2] : h[0], k1 = hi ? h[3] : h[1];
            k0 += __shfl_xor_sync(0xffffffffu, hi ? h[0] : h[2], 16);
            k1 += __shfl_xor_sync(0xffffffffu, hi ? h[1] : h[3], 16);
            double kk = h8 ? k1 : k0;
            kk += __shfl_xor_sync(0xffffffffu, h8 ? k0 : k1, 8);
            kk += __shfl_xor_sync(0xffffffffu, kk, 4);
            kk += __shfl_xor_sync(0xffffffffu, kk, 2);
            kk += __shfl_xor_sync(0xffffffffu, kk, 1);
            if ((lane & 7) == 0) scr[buf][lane >> 3][warp] = kk;
            __syncthreads();
            double t2 = 0.0;
#pragma unroll
            for (int w = (lane & 7); w < NW; w += 8) t2 += scr[buf][lane >> 3][w];
            t2 += __shfl_xor_sync(0xffffffffu, t2, 4);
            t2 += __shfl_xor_sync(0xffffffffu, t2, 2);
            t2 += __shfl_xor_sync(0xffffffffu, t2, 1);
#pragma unroll
            for (int q = 0; q < B; ++q) h[q] = __shfl_sync(0xffffffffu, t2, q * 8);
        } else {
#pragma unroll
            for (int q = 0; q < B; ++q) h[q] = warp_sum(h[q]);
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < B; ++q) scr[buf][q][warp] = h[q];
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < B; ++q) h[q] = warp_sum((lane < NW) ? scr[buf][q][lane] : 0.0);
        }
        buf ^= 1;
        // columns past the k-th are clamped copies of it: projecting with h = 0 leaves the working vector unchanged,
        // so the tail block needs no branch
#pragma unroll
        for (int q = 0; q < B; ++q) {
            if (c0 + q >= k) h[q] = 0.0;
            else if (tid == 0) hcol[c0 + q] = pass ? hcol[c0 + q] + h[q] : h[q];
#pragma unroll
            for (int e = 0; e < EPT; ++e) v[e] = fma(-h[q], blk[q][e], v[e]);
        }
    };
    for (int pass = 0; pass < 2; ++pass) {
        load_block(ca, 0);
        for (int c0 = 0; c0 < k; c0 += 2 * B) {
            // The next block is requested UNCONDITIONALLY (column indices are clamped to k-1) before the current one is
            // reduced: behind a branch the compiler sinks the loads below the dot products of the current block, and
            // every block then pays the full memory latency (measured: 3.5 us per block instead of ~1).
            load_block(cb, c0 + B);
            process(ca, c0, pass);
            if (c0 + B < k) {
                load_block(ca, c0 + 2 * B);
                process(cb, c0 + B, pass);
            }
        }
    }
    auto reduce1 = [&](double x) -> double {
        x = warp_sum(x);
        if (lane == 0) scr[buf][0][warp] = x;
        __syncthreads();
        const double r = (lane < NW) ? scr[buf][0][lane] : 0.0;
        buf ^= 1;
        return warp_sum(r);
    };
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) acc = fma(v[e], v[e], acc);
    const double beta = sqrt(reduce1(acc));
    const double inv = 1.0 / beta;                      // no zero-norm guard in the reference (:35-36)
    double* vnew = Vs + (long long)k * p.ldv;
    const double* b = p.b + (long long)s * p.ldv;
    acc = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int i = tid + e * THREADS;
        if (i < n) {
            const double x = v[e] * inv;
            vnew[i] = x;
            acc = fma(x, b[i], acc);
        }
    }
    const double btn = reduce1(acc);
    __syncthreads();
    double* Hs = p.Hd + (long long)s * p.ncol * p.ncol + (long long)(k - 1) * p.ncol;
    for (int c = tid; c < k; c += THREADS) Hs[c] = hcol[c];
    if (tid == 0) {
        Hs[k] = beta;
        p.bt[(long long)s * p.ncol + k] = btn;
        double* T = p.T + (long long)s * 3 * p.ncol;
        T[k - 1] = hcol[k - 1];
        T[p.ncol + (k - 1)] = beta;
    }
}

// ------------------------------------------------------------------------------------------
// Arnoldi step k for every mode: one CTA per mode runs the two-pass MGS and stores column k of
// the Hessenberg matrix.  Algorithmic HBM bytes per mode: (16 k + 8 (ndiag + 3)) n.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) arnoldi_mgs_kernel(KrylovParams p, int k, double* vscratch) {
    if (!cta_running(p.status)) return;
    extern __shared__ double smem[];
    __shared__ double scratch[32];
    const int s = blockIdx.x;
    double* hcol = smem;
    double* v = vscratch ? vscratch + (long long)s * p.ldv : smem + p.ncol;
    mgs_step_cta(p, s, k, v, hcol, scratch);
    double* Hs = p.Hd + (long long)s * p.ncol * p.ncol + (long long)(k - 1) * p.ncol;  // column k
    for (int c = threadIdx.x; c <= k; c += blockDim.x) Hs[c] = hcol[c];
    if (threadIdx.x == 0) {
        double* T = p.T + (long long)s * 3 * p.ncol;
        T[k - 1] = hcol[k - 1];
        T[p.ncol + (k - 1)] = hcol[k];  // H[k+1,k], read by the residual's boundary term
    }
}

}  // namespace tk
