// tk_api.cu -- C-ABI of libtensorkrylov_b200.so: handle, inputs, the device-resident solve loop.
//
// The loop body follows tensorkrylov! (src/tensor_krylov_method.jl:63-120) phase by phase:
//   step_bases  : orthonormalize!(decomp, k) + update_rhs!            kernels (1)
//   compress    : solve_compressed_system                            kernels (2) (3)
//   residual    : residualnorm! + the convergence/breakdown decision  kernels (4)
// All state stays in HBM; the host only enqueues launches and polls the device status word a
// few iterations behind the GPU.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges cost nothing unless a profiler is attached

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

extern "C" char** environ;

#include "../../include/tensorkrylov_b200.h"
#include "tk_compress.cuh"
#include "tk_expm.cuh"
#include "tk_host.h"
#include "tk_krylov.cuh"

namespace tk {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define TK_CUDA(call)                                                                                     \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return tk::set_error(TK_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define TK_TRY(call)            \
    do {                        \
        int rc__ = (call);      \
        if (rc__ != 0) return rc__; \
    } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL is only needed for world > 1: bind it lazily so a single-GPU process never loads it.
// ---------------------------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_bind() {
    if (g_nccl.lib) return 0;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return set_error(TK_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
#define TK_SYM(field, name)                                                            \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name));         \
    if (!g_nccl.field) return set_error(TK_ENCCL, "libnccl has no symbol %s", name);
    TK_SYM(GetUniqueId, "ncclGetUniqueId")
    TK_SYM(CommInitRank, "ncclCommInitRank")
    TK_SYM(CommDestroy, "ncclCommDestroy")
    TK_SYM(AllGather, "ncclAllGather")
    TK_SYM(Broadcast, "ncclBroadcast")
    TK_SYM(GetErrorString, "ncclGetErrorString")
#undef TK_SYM
    g_nccl.lib = lib;
    return 0;
}

// One NCCL communicator per unique id and process, shared by every handle created with that id (a unique id
// may initialise only one communicator; solves come and go, the communicator stays).
struct CommEntry { ncclComm_t comm; int rank, world, refs; };
static std::map<std::string, CommEntry> g_comms;

#define TK_NCCL(call)                                                                              \
    do {                                                                                           \
        ncclResult_t r__ = (call);                                                                 \
        if (r__ != ncclSuccess) return tk::set_error(TK_ENCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r__)); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Device memory: a process-level cache of freed blocks keyed by (device, bytes).  The reference's entry point
// builds its whole state per call (decompositions.jl:127-174); a drop-in does the same, and cudaMalloc/cudaFree
// of multi-GB bases would otherwise cost more than the solve.  tk_release_cache() returns everything.
// ---------------------------------------------------------------------------------------------
static std::map<std::pair<int, size_t>, std::vector<void*>> g_pool;
static size_t g_pool_bytes = 0;
static std::mutex g_mutex;   // guards the process-level tables (block cache, communicators, kernel attributes)

static void pool_trim() {
    for (auto& kv : g_pool) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(kv.first.first);
        for (void* q : kv.second) cudaFree(q);
        cudaSetDevice(cur);
    }
    g_pool.clear();
    g_pool_bytes = 0;
}

static void evict_parked_handles();   // parked solvers give their blocks back to the cache (defined below)

static cudaError_t pool_alloc(void** out, size_t bytes) {
    std::unique_lock<std::mutex> lock(g_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    auto it = g_pool.find(std::make_pair(dev, bytes));
    if (it != g_pool.end() && !it->second.empty()) {
        *out = it->second.back();
        it->second.pop_back();
        g_pool_bytes -= bytes;
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        lock.unlock();
        evict_parked_handles();
        lock.lock();
        pool_trim();
        e = cudaMalloc(out, bytes);
    }
    return e;
}

static void pool_free(void* q, size_t bytes) {
    std::lock_guard<std::mutex> lock(g_mutex);
    int dev = 0;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, q) == cudaSuccess) dev = attr.device; else cudaGetLastError();
    g_pool[std::make_pair(dev, bytes)].push_back(q);
    g_pool_bytes += bytes;
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t count = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) pool_free(p, count * sizeof(T));
        p = nullptr;
        count = 0;
    }
    int alloc(size_t n, bool zero = true) {
        release();
        if (n == 0) n = 1;
        cudaError_t e = pool_alloc(reinterpret_cast<void**>(&p), n * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            return set_error(TK_ENOMEM, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
        }
        count = n;
        if (zero) {
            e = cudaMemset(p, 0, n * sizeof(T));
            if (e != cudaSuccess) return set_error(TK_ECUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
        }
        return 0;
    }
};

// Peer-exchange buffers of one (communicator, slot size): receive buffer and flags of this rank plus the mapped views of
// every peer's.  Setting them up is collective (two all-gathers, IPC export / open), so they belong to the process,
// not to a handle: a solver created later on the same communicator with the same slot size borrows them (a drop-in
// caller creates a handle per solve).  The flag values only ever grow, across handles too: `epoch` is the running
// base, advanced by nmax + 2 per solve -- every rank runs the same sequence of solves, so all ranks agree on it.
struct ExchangeCtx {
    DevBuf<double> recv;
    DevBuf<unsigned long long> flags;
    PeerExchange px;
    bool ready = false;
    long long epoch = 0;
};
static std::map<std::pair<std::string, long long>, std::unique_ptr<ExchangeCtx>> g_exchange;

// NVTX range over a phase of the host-side driver (solve, segment, graph recording, solution fetch)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct HostOp {
    int type = OP_DIA;
    int ndiag = 0;
    int offs[MAX_DIAG] = {0};
    long long ld = 0;
    long long nnz = 0;
    DevBuf<double> vals;   // diag / val / dense
    DevBuf<int> rowptr, colidx;
    bool constd = false;       // DIA with one value per diagonal (Toeplitz): the 3-term kernel needs no operator loads
    double cval[MAX_DIAG] = {0};
    double norm_bound = 0.0;   // sqrt(||A||_1 ||A||_inf) >= ||A||_2, bounds the squarings of the Hessenberg exponential
    double bytes_per_row() const {  // operator bytes streamed per row by one SpMV
        if (type == OP_DIA) return 8.0 * ndiag;
        if (type == OP_CSR) return ld > 0 ? (12.0 * nnz + 4.0 * (ld + 1)) / (double)ld : 0.0;
        return 8.0 * (double)ld;
    }
};

struct SchedEntry {
    bool set = false;
    double lambda_min = 0.0;
    int t = 0;
    size_t off = 0;  // into the alpha/omega pools
};

enum { TM_TTR = 0, TM_GRAM = 1, TM_MGS = 2, TM_EIG = 3, TM_ASM = 4, TM_COMBINE = 5, TM_SOLVE = 6, TM_REGION = 7, TM_KINDS = 8 };

}  // namespace tk

using namespace tk;

// Streams, events and the pinned status ring of a handle.  Creating and destroying them (cudaMallocHost/cudaFreeHost
// in particular) costs far more than a solve, so destroyed handles park the bundle in a process-level free list.
struct tk_resources {
    int device = 0;
    cudaStream_t s_main = nullptr, s_asm = nullptr, s_eig[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t step_ev[8], eig_ev[8], asm_ev[8], seg_ev[4], join_ev[10], ev_fork, ev_solve[2], ev_region;
    SolveCtl* hctl = nullptr;           // pinned + mapped: tolerance/epoch in, exit status out
    SolveCtl* hctl_dev = nullptr;       // the same block as the device sees it
    cudaStream_t s_copy = nullptr;      // device -> host copies of the solution, overlapped with its computation
    cudaEvent_t copy_ev[2], fill_ev[2];
    double* stage[2] = {nullptr, nullptr};   // pinned staging buffers for results that go to pageable memory
    size_t stage_bytes = 0;
    std::vector<cudaEvent_t> ev_pool;   // timing events, grown on demand
};
static std::vector<tk_resources*> g_res_free;
static std::map<size_t, std::vector<void*>> g_host_pool;   // parked page-locked blocks of tk_alloc_host, by size
static std::map<void*, size_t> g_host_live;
static std::map<std::string, void*> g_ipc_open;            // peer buffers mapped through CUDA IPC, by exported handle

static int acquire_resources(int device, tk_resources** out) {
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        for (size_t i = 0; i < g_res_free.size(); ++i)
            if (g_res_free[i]->device == device) {
                *out = g_res_free[i];
                g_res_free.erase(g_res_free.begin() + i);
                return 0;
            }
    }
    std::unique_ptr<tk_resources> r(new tk_resources());
    r->device = device;
    TK_CUDA(cudaStreamCreateWithFlags(&r->s_main, cudaStreamNonBlocking));
    TK_CUDA(cudaStreamCreateWithFlags(&r->s_asm, cudaStreamNonBlocking));
    for (auto& st : r->s_eig) TK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    TK_CUDA(cudaStreamCreateWithFlags(&r->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 8; ++i) {
        TK_CUDA(cudaEventCreateWithFlags(&r->step_ev[i], cudaEventDisableTiming));
        TK_CUDA(cudaEventCreateWithFlags(&r->eig_ev[i], cudaEventDisableTiming));
        TK_CUDA(cudaEventCreateWithFlags(&r->asm_ev[i], cudaEventDisableTiming));
    }
    for (auto& e : r->seg_ev) TK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : r->join_ev) TK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : r->copy_ev) TK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : r->fill_ev) TK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    TK_CUDA(cudaEventCreateWithFlags(&r->ev_fork, cudaEventDisableTiming));
    TK_CUDA(cudaEventCreate(&r->ev_solve[0]));
    TK_CUDA(cudaEventCreate(&r->ev_solve[1]));
    TK_CUDA(cudaEventCreate(&r->ev_region));
    TK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&r->hctl), sizeof(SolveCtl), cudaHostAllocMapped));
    TK_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&r->hctl_dev), r->hctl, 0));
    std::memset(r->hctl, 0, sizeof(SolveCtl));
    *out = r.release();
    return 0;
}

struct tk_handle {
    tk_resources* res = nullptr;
    ~tk_handle() {
        for (auto& sg : segs) if (sg.exec) cudaGraphExecDestroy(sg.exec);
        if (res) {                      // also on a failed tk_create
            res->ev_pool.swap(ev_pool);
            std::lock_guard<std::mutex> lock(g_mutex);
            g_res_free.push_back(res);
        }
    }
    int d = 0, dl = 0, first = 0, nmax = 0, ncol = 0, n = 0;
    int dk = 0;          // modes the Krylov kernels advance: dl, plus a shadow copy of global mode 0 (slot dl) when the
                         // reference's H_1-for-all-modes rule is on and another rank owns mode 0
    int eig_slot = 0;    // local slot whose T feeds class 0 under TK_FLAG_REFERENCE_H1
    int instance = 0, matrixclass = 0, variant = 0, flags = 0, device = 0, rank = 0, world = 1, sm_count = 148;
    bool gram_does_bt = false;     // the last 3-term launch left b~[k+1] to the Gram-row kernel that follows it
    long long ldv = 0;
    int per_mode = 0, ncls = 1;
    bool use_expm = false;   // compressed solve through the dense exponential (NonSymInstance, and the EigValMat class)
    int chunk_modes = 16, nchunks = 1, chunk_base = 0;
    cudaStream_t stream = nullptr;    // Krylov-step kernels (1)
    cudaStream_t stream2 = nullptr;   // CP assembly + residual (3)(4): runs behind the Krylov steps, concurrently
    // eigensolver (2): eig(k) only needs step k, so consecutive k run concurrently on NEIG round-robin streams
    // (one SM each) and overlap the Krylov steps and the assembly of earlier iterations
    // (as many as the ring is deep: at k ~ 256 one bisection eigensolve takes ~1 ms on its one SM, and with few short
    // modes -- config 2 -- the solve is bound by how many of them are in flight: 137 ms of eigensolves per 43 ms solve)
    static constexpr int NEIG = 8, NBUF = 8;
    cudaStream_t stream3[NEIG] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // Dependencies between the streams: one event per kind and iteration (ring of 8), valid inside the segment that
    // recorded it -- segments fork from and join into `stream`, so anything older is already complete
    struct EvSlot { cudaEvent_t ev = nullptr; int k = -1, seg = -1; };
    EvSlot step_slot[8], eigdone_slot[8], asm_slot[8];
    int cur_seg = -1;
    ncclComm_t comm = nullptr;

    // The iteration loop is cut into segments of consecutive iterations; a segment is enqueued either directly or as
    // one CUDA graph launch (recorded once per configuration).  steps k0..k1 advance the Krylov bases, chains c0..c1
    // are the compressed solve + residual estimate of those iterations; the chains of the last few steps of a
    // segment are deferred into the next one so the Krylov stream never waits for them at a boundary.
    struct Segment {
        int k0 = 2, k1 = 1, c0 = 2, c1 = 1;
        bool first = false;
        cudaGraphExec_t exec = nullptr;
        unsigned long long epoch = 0;
        long long launches = 0;
    };
    std::vector<Segment> segs;
    unsigned long long sig = 0;          // signature of everything recorded launches carry as arguments (prepare)
    unsigned long long last_solve_sig = 0;
    std::string shape_key;               // tk_create arguments: a parked handle is revived by an identical tk_create
    long long solve_count = 0;
    SolveCtl* hctl = nullptr;            // pinned (host view)
    SolveCtl* hctl_dev = nullptr;        // pinned (device view)
    DevBuf<SolveCtl> ctl_d;
    double graph_build_ms = 0.0;         // host time spent recording + instantiating graphs in the last tk_solve
    int graphs_launched = 0;

    // peer exchange of the merged partials (world > 1): receive buffer + flags of this rank, peers' mapped views
    PeerExchange px;
    bool px_ready = false;
    ExchangeCtx* xc = nullptr;           // process-level buffers behind px (not owned)
    std::string comm_key;                // the NCCL unique id this handle's communicator was made from

    // Krylov state
    DevBuf<double> V, b, T, Hd, bt, g, S, orthS, bnorm2, vscratch;
    DevBuf<int> fallbacks, mode_op_d, status_d, term_k_d, eigfail_d;
    DevBuf<long long> niter_d;
    DevBuf<OpDesc> ops_d;
    std::vector<std::unique_ptr<HostOp>> ops;
    std::vector<int> mode_op;   // local mode -> op index (-1 unset)
    std::vector<char> rhs_set;
    bool ops_dirty = true;
    std::vector<double> a1_lead;   // leading nmax x nmax block of global mode 0's operator (spectral data, eigenvalues.jl:276-282)

    // schedule
    std::vector<SchedEntry> sched;
    std::vector<double> alpha_pool, omega_pool;
    DevBuf<double> alpha_d, omega_d;
    bool sched_dirty = true;
    int tmax = 0;

    // compressed solve / residual
    DevBuf<double> theta, Q, Y, Z, E, bbm, partials, gathered, relres_d, projres_d, orth_d, detail_d;
    DevBuf<double> eig_scratch;         // second k x k plane per eigenproblem (bisection kernel), same ring as Q
    DevBuf<int> eig_need;               // per problem: 1 if the QL fallback must recompute it
    DevBuf<unsigned int> ticket_d;      // last-CTA-done counter of combine_chunk_kernel
    DevBuf<unsigned int> tickets;       // per mode: last-CTA-done counter of gram_row_kernel
    DevBuf<double> merged;              // this GPU's merged partial (what the all-gather ships)
    DevBuf<double> exW;                 // NonSymInstance: workspace of the batched matrix exponential
    DevBuf<int> ex_nsq, ex_where, cls_mode_d;
    int ex_ld = 0;
    int ring_depth = NBUF;              // iterations whose spectral data (theta/Q or the exponentials) may be in flight
    int ldq = 0;
    long long ystride = 0, estride = 0, pstride_max = 0;
    bool work_ready = false;
    int last_k = 0, last_t = 0, last_tld = 0;
    double last_lam_inv = 0.0;
    bool begun = false;

    // timing
    struct Timed { int kind; cudaEvent_t a, b; double bytes; };
    std::vector<Timed> timed;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t ev_solve[2] = {nullptr, nullptr};
    cudaEvent_t ev_region = nullptr;   // tk_timing_mark: start of a multi-solve timed region
    size_t ev_used = 0;
    double tm_ms[TM_KINDS] = {0}, tm_bytes[TM_KINDS] = {0};
    long long tm_launches[TM_KINDS] = {0};
    long long launches = 0;

    KrylovParams kp() const {
        KrylovParams p;
        p.n = n; p.ncol = ncol; p.ldv = ldv; p.vstride = (long long)ncol * ldv;
        p.V = V.p; p.b = b.p; p.T = T.p; p.Hd = Hd.p; p.bt = bt.p; p.g = g.p; p.S = S.p; p.orthS = orthS.p;
        p.fallbacks = fallbacks.p; p.ops = ops_d.p; p.mode_op = mode_op_d.p; p.status = status_d.p; p.snap = status_d.p + 1;
        p.mode0_local = (first == 0 && dl > 0) ? 0 : -1;
        return p;
    }
};

// Solvers released by tk_destroy are parked whole (state buffers, workspaces, recorded graphs, streams) and revived by
// a tk_create with identical arguments: the reference's calling convention builds its whole state per solve
// (decompositions.jl:127-174), and a drop-in caller that does the same should not pay allocation, workspace set-up and
// graph recording every time.  Inputs are NOT kept: a revived handle must be fed operators, right-hand sides and the
// schedule like a new one.  TK_HANDLE_CACHE = number of parked handles (default 2, 0 = off).
static std::vector<tk_handle*> g_parked;

namespace tk {

static void evict_parked_handles() {
    std::vector<tk_handle*> gone;
    { std::lock_guard<std::mutex> lock(g_mutex); gone.swap(g_parked); }
    for (tk_handle* q : gone) delete q;        // blocks go back to the cache (pool_free takes the lock itself)
}

// developer knobs for kernel tuning experiments (unset in production)
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

static int check_mode(const tk_handle* h, int s, bool* local) {
    if (!h) return set_error(TK_EINVAL, "null handle");
    if (s < 0 || s >= h->d) return set_error(TK_EINVAL, "mode %d out of range [0,%d)", s, h->d);
    *local = (s >= h->first && s < h->first + h->dl);
    return 0;
}

// local slots that hold global mode s: its own slot, and the shadow slot for mode 0
static int slots_of(const tk_handle* h, int s, int out[2]) {
    int n = 0;
    if (s >= h->first && s < h->first + h->dl) out[n++] = s - h->first;
    if (s == 0 && h->dk > h->dl) out[n++] = h->dl;
    return n;
}

static cudaEvent_t next_event(tk_handle* h) {
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->ev_pool.push_back(e);
    }
    return h->ev_pool[h->ev_used++];
}

struct TimedScope {
    tk_handle* h; int kind; cudaEvent_t a = nullptr; double bytes; cudaStream_t st;
    TimedScope(tk_handle* h_, int kind_, double bytes_, cudaStream_t st_) : h(h_), kind(kind_), bytes(bytes_), st(st_) {
        const bool on = (h->flags & TK_FLAG_TIME_ALL) || ((h->flags & TK_FLAG_TIME_KERNELS) && kind <= TM_MGS);
        if (on) { a = next_event(h); cudaEventRecord(a, st); }
    }
    ~TimedScope() {
        if (a) { cudaEvent_t b = next_event(h); cudaEventRecord(b, st); h->timed.push_back({kind, a, b, bytes}); }
    }
};

static int upload_ops(tk_handle* h) {
    if (!h->ops_dirty) return 0;
    for (int s = 0; s < h->dk; ++s)
        if (h->mode_op[s] < 0) return set_error(TK_ESTATE, "operator of mode %d not set", s < h->dl ? h->first + s : 0);
    std::vector<OpDesc> descs(h->ops.size());
    for (size_t i = 0; i < h->ops.size(); ++i) {
        const HostOp& o = *h->ops[i];
        OpDesc& dsc = descs[i];
        std::memset(&dsc, 0, sizeof(dsc));
        dsc.type = o.type; dsc.ndiag = o.ndiag; dsc.ld = o.ld;
        std::memcpy(dsc.offs, o.offs, sizeof(o.offs));
        dsc.diag = o.type == OP_DIA ? o.vals.p : nullptr;
        dsc.val = o.type == OP_CSR ? o.vals.p : nullptr;
        dsc.dense = o.type == OP_DENSE ? o.vals.p : nullptr;
        dsc.rowptr = o.rowptr.p; dsc.colidx = o.colidx.p;
        dsc.constd = o.constd ? 1 : 0;
        std::memcpy(dsc.cval, o.cval, sizeof(o.cval));
    }
    TK_TRY(h->ops_d.alloc(std::max<size_t>(descs.size(), 1), false));
    TK_CUDA(cudaMemcpy(h->ops_d.p, descs.data(), descs.size() * sizeof(OpDesc), cudaMemcpyHostToDevice));
    TK_CUDA(cudaMemcpy(h->mode_op_d.p, h->mode_op.data(), h->dk * sizeof(int), cudaMemcpyHostToDevice));
    h->ops_dirty = false;
    return 0;
}

static int upload_schedule(tk_handle* h) {
    if (!h->sched_dirty) return 0;
    int tmax = 0;
    for (int k = 2; k <= h->nmax; ++k) {
        if (!h->sched[k].set) return set_error(TK_ESTATE, "schedule entry k=%d not set (tk_set_schedule)", k);
        tmax = std::max(tmax, h->sched[k].t);
    }
    TK_TRY(h->alpha_d.alloc(std::max<size_t>(h->alpha_pool.size(), 1), false));
    TK_TRY(h->omega_d.alloc(std::max<size_t>(h->omega_pool.size(), 1), false));
    TK_CUDA(cudaMemcpy(h->alpha_d.p, h->alpha_pool.data(), 8 * h->alpha_pool.size(), cudaMemcpyHostToDevice));
    TK_CUDA(cudaMemcpy(h->omega_d.p, h->omega_pool.data(), 8 * h->omega_pool.size(), cudaMemcpyHostToDevice));
    if (tmax != h->tmax) h->work_ready = false;
    h->tmax = tmax;
    h->sched_dirty = false;
    return 0;
}

static int setup_peer_exchange(tk_handle* h);

static int alloc_work(tk_handle* h) {
    if (h->work_ready) return 0;
    const int kmax = h->nmax, tmax = std::max(h->tmax, 1);
    const int tld = (tmax + 3) & ~3;
    h->ldq = (h->ncol + 1) & ~1;
    if (!h->use_expm) {
        // ring: eig(k+1..) overlaps assembly(k).  Q and the bisection scratch are ncls * ldq^2 doubles per slot, so
        // with per-mode eigenproblems and a large nmax the ring is shortened until it fits in 4 GB
        h->ring_depth = std::min<int>(tk_handle::NBUF, std::max(1, env_int("TK_RING_DEPTH", tk_handle::NBUF)));
        const double slot = 2.0 * 8.0 * (double)h->ncls * h->ldq * h->ldq;
        while (h->ring_depth > 1 && slot * h->ring_depth > 4e9) h->ring_depth >>= 1;
    }
    const size_t depth = h->use_expm ? 1 : (size_t)h->ring_depth;
    TK_TRY(h->theta.alloc(depth * h->ncls * h->ncol));
    TK_TRY(h->Q.alloc(depth * h->ncls * h->ldq * h->ldq));
    TK_TRY(h->eig_scratch.alloc(depth * h->ncls * h->ldq * h->ldq, false));
    TK_TRY(h->eig_need.alloc(depth * h->ncls));
    h->ystride = (long long)kmax * tld;
    TK_TRY(h->Y.alloc((size_t)h->dl * h->ystride));
    TK_TRY(h->Z.alloc((size_t)h->dl * h->ystride));
    h->estride = 3LL * tmax * tmax;
    TK_TRY(h->E.alloc((size_t)h->dl * h->estride));
    TK_TRY(h->bbm.alloc(h->dl));
    h->pstride_max = 5LL * tmax * tmax + 2LL * tmax + 8;
    TK_TRY(h->partials.alloc((size_t)h->nchunks * h->pstride_max));
    TK_TRY(h->ticket_d.alloc(1));
    TK_TRY(h->merged.alloc((size_t)h->pstride_max));
    if (h->world > 1) {
        TK_TRY(h->gathered.alloc((size_t)h->world * h->pstride_max));
        TK_TRY(setup_peer_exchange(h));
    }
    if (h->use_expm) {
        h->ex_ld = (h->nmax + 3) & ~3;
        // exponentials of up to ring_depth iterations are in flight (they only depend on the Krylov step); with one
        // matrix class the ring is cheap, with per-mode classes it is kept at depth 1
        h->ring_depth = h->ncls == 1 ? tk_handle::NBUF : 1;
        const size_t nmat = (size_t)h->ncls * tmax * h->ring_depth;
        const size_t bytes = nmat * EX_SLOTS * (size_t)h->ex_ld * h->ex_ld * 8;
        size_t free_b = 0, total_b = 0;
        TK_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if (bytes > free_b + g_pool_bytes)
            return set_error(TK_ENOMEM, "matrix-exponential workspace needs %.1f GB (%zu matrices of order %d); use TK_FLAG_REFERENCE_H1 or fewer modes per GPU",
                             bytes / 1e9, nmat, h->nmax);
        TK_TRY(h->exW.alloc(nmat * EX_SLOTS * (size_t)h->ex_ld * h->ex_ld, false));
        TK_TRY(h->ex_nsq.alloc(nmat));
        TK_TRY(h->ex_where.alloc(nmat));
        std::vector<int> cm(h->ncls);
        for (int c = 0; c < h->ncls; ++c) cm[c] = h->per_mode ? c : h->eig_slot;
        TK_TRY(h->cls_mode_d.alloc(h->ncls, false));
        TK_CUDA(cudaMemcpy(h->cls_mode_d.p, cm.data(), sizeof(int) * cm.size(), cudaMemcpyHostToDevice));
    }
    h->work_ready = true;
    return 0;
}

static size_t smem_limit(const tk_handle* h) {
    (void)h;
    return 200 * 1024;
}

// Opt a kernel in to more than 48 KB of dynamic shared memory -- once per kernel and size, not once per launch.
template <typename K>
static int allow_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return 0;
    static std::map<std::pair<int, const void*>, size_t> granted;
    std::lock_guard<std::mutex> lock(g_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& have = granted[std::make_pair(dev, reinterpret_cast<const void*>(kernel))];
    if (bytes > have) {
        TK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return 0;
}

// --- kernel (1) launches -------------------------------------------------------------------
// Operator bytes one all-mode SpMV has to fetch from HBM, averaged per mode-row: an operator object aliased by
// several modes (the reference aliases ONE matrix d times, tensor_struct.jl:208-210) is fetched once, not per mode.
static double op_bytes_per_row(const tk_handle* h) {
    std::set<int> distinct;
    for (int s = 0; s < h->dk; ++s) distinct.insert(h->mode_op[s]);
    double acc = 0.0;
    for (int o : distinct) acc += h->ops[o]->bytes_per_row();
    return h->dk ? acc / h->dk : 0.0;
}

template <int CPM, int ND>
static int launch_ttr_t(tk_handle* h, int k, int threads, size_t smem) {
    auto kernel = lanczos_ttr_kernel<CPM, ND>;
    TK_TRY(allow_smem(kernel, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h->dk * CPM);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CPM; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CPM > 1 ? 1 : 0;
    KrylovParams p = h->kp();
    TK_CUDA(cudaLaunchKernelEx(&cfg, kernel, p, k));
    h->launches++;
    return 0;
}

template <int CPM, int ND, bool CONSTD, int RPT, bool WITHB>
static int launch_ttr_bulk_t(tk_handle* h, int k, int threads, size_t smem) {
    auto kernel = lanczos_ttr_bulk_kernel<CPM, ND, CONSTD, RPT, WITHB>;
    TK_TRY(allow_smem(kernel, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h->dk * CPM);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (CPM > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CPM; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    KrylovParams p = h->kp();
    TK_CUDA(cudaLaunchKernelEx(&cfg, kernel, p, k));
    h->launches++;
    return 0;
}

// Bulk-copy 3-term step (lanczos_ttr_bulk_kernel): banded DIA operators with 3 or 4 diagonals inside the halo.
// Returns 1 when the configuration is not eligible and the generic kernel must run.
static int launch_ttr_bulk(tk_handle* h, int k, int nd) {
    if (nd != 3 && nd != 4) return 1;
    if (!env_int("TK_TTR_BULK", 1)) return 1;
    bool constd = true;
    for (int s = 0; s < h->dk; ++s) {
        const HostOp& o = *h->ops[h->mode_op[s]];
        for (int j = 0; j < o.ndiag; ++j)
            if (o.offs[j] < -TTR_HALO || o.offs[j] > TTR_HALO) return 1;
        constd = constd && o.constd;
    }
    if (env_int("TK_TTR_NOCONST", 0)) constd = false;
    // CTAs per mode: slices of <= 2560 rows (40 KB of shared memory -> 4 CTAs per SM; 60 KB -> 3 with the b slice),
    // at least 64 rows each
    int cpm = 1;
    while (cpm < 8 && h->n / cpm > 2560) cpm *= 2;
    while (cpm < 8 && (long long)h->dk * cpm < 296 && h->n / (2 * cpm) >= 512) cpm *= 2;
    if (env_int("TK_TTR_CPM", 0)) cpm = env_int("TK_TTR_CPM", 0);
    if (cpm != 1 && cpm != 2 && cpm != 4 && cpm != 8) return 1;
    const int chunk = (((h->n + cpm - 1) / cpm) + 1) & ~1;
    if ((long long)(cpm - 1) * chunk >= h->n || chunk < 64) return 1;      // every CTA of a cluster owns rows
    // with full orthogonalisation the Gram-row kernel that follows supplies b~[k+1] (TK_TTR_NOB = 0: compute it here)
    const bool withb = !(h->variant == TK_LANCZOS_REORTH && env_int("TK_TTR_NOB", 1));
    const size_t smem = ((size_t)(withb ? 3 : 2) * chunk + 2 * TTR_HALO) * 8;
    if (smem > smem_limit(h)) return 1;
    // every thread keeps TTR_RPT rows of its slice in registers (5 and 20 rows per thread measured slower)
    int threads = std::max(64, (((chunk + TTR_RPT - 1) / TTR_RPT) + 31) & ~31);
    if (env_int("TK_TTR_THREADS", 0)) threads = std::max(threads, env_int("TK_TTR_THREADS", 0));
    if (threads > 512) return 1;
    h->gram_does_bt = !withb;
#define TK_BULK_ND(C, N)                                                                                 \
    do {                                                                                                 \
        if (withb) return constd ? launch_ttr_bulk_t<C, N, true, TTR_RPT, true>(h, k, threads, smem)     \
                                 : launch_ttr_bulk_t<C, N, false, TTR_RPT, true>(h, k, threads, smem);   \
        return constd ? launch_ttr_bulk_t<C, N, true, TTR_RPT, false>(h, k, threads, smem)               \
                      : launch_ttr_bulk_t<C, N, false, TTR_RPT, false>(h, k, threads, smem);             \
    } while (0)
#define TK_BULK_CASE(C)                                                                                  \
    case C:                                                                                              \
        if (nd == 3) TK_BULK_ND(C, 3);                                                                   \
        TK_BULK_ND(C, 4);
    switch (cpm) {
        TK_BULK_CASE(1)
        TK_BULK_CASE(2)
        TK_BULK_CASE(4)
        TK_BULK_CASE(8)
    }
#undef TK_BULK_CASE
#undef TK_BULK_ND
    h->gram_does_bt = false;
    return 1;
}

static int launch_ttr(tk_handle* h, int k) {
    const double bytes = (op_bytes_per_row(h) + 32.0) * (double)h->n * h->dk;
    TimedScope ts(h, TM_TTR, bytes, h->stream);
    h->gram_does_bt = false;
    // CTAs per mode (one cluster): at least 2 whenever a slice keeps >= 2048 rows (measured best on B200 at
    // n = 10^4), more when the modes alone cannot fill the machine or a slice would not fit in shared memory
    int cpm = 1;
    while (cpm < 8 && ((size_t)h->n * 8 / cpm > smem_limit(h) ||
                       (h->n / (2 * cpm) >= 2048 && cpm < 2) ||
                       (h->n / (2 * cpm) >= 512 && (long long)h->dk * cpm < 296)))
        cpm *= 2;
    if (env_int("TK_TTR_CPM", 0)) cpm = env_int("TK_TTR_CPM", 0);
    const int chunk = (((h->n + cpm - 1) / cpm) + 1) & ~1;
    const size_t smem = (size_t)chunk * 8;
    if (smem > smem_limit(h)) return set_error(TK_EUNSUPPORTED, "n = %d is too large for the 3-term step kernel", h->n);
    int threads = chunk >= 2048 ? 512 : 256;
    if (env_int("TK_TTR_THREADS", 0)) threads = env_int("TK_TTR_THREADS", 0);
    // all operators DIA with the same small number of diagonals -> kernels with the diagonal loop unrolled
    int nd = -1;
    for (int s = 0; s < h->dk && nd != 0; ++s) {
        const HostOp& o = *h->ops[h->mode_op[s]];
        const int v = (o.type == OP_DIA && (o.ndiag == 3 || o.ndiag == 4)) ? o.ndiag : 0;
        nd = (nd == -1 || nd == v) ? v : 0;
    }
    if (env_int("TK_TTR_GENERIC", 0)) nd = 0;
    {
        const int rc = launch_ttr_bulk(h, k, nd);
        if (h->gram_does_bt) ts.bytes -= 8.0 * (double)h->n * h->dk;      // b_s is not read at all
        if (rc != 1) return rc;
    }
#define TK_TTR_CASE(C)                                                        \
    case C:                                                                   \
        if (nd == 3) return launch_ttr_t<C, 3>(h, k, threads, smem);          \
        if (nd == 4) return launch_ttr_t<C, 4>(h, k, threads, smem);          \
        return launch_ttr_t<C, 0>(h, k, threads, smem);
    switch (cpm) {
        TK_TTR_CASE(1)
        TK_TTR_CASE(2)
        TK_TTR_CASE(4)
        default:
            if (nd == 3) return launch_ttr_t<8, 3>(h, k, threads, smem);
            if (nd == 4) return launch_ttr_t<8, 4>(h, k, threads, smem);
            return launch_ttr_t<8, 0>(h, k, threads, smem);
    }
#undef TK_TTR_CASE
}

// Gram row of the newest column for `nmodes` modes starting at local mode `base`
// Gram row of the newest column for `nmodes` modes starting at local mode `base`, followed (same launch, last CTA of
// each mode) by the monitor: monitor = 0 plain bookkeeping, 1 = with the LanczosReorth MGS fallback.
static int launch_gram(tk_handle* h, int ncols, int base, int nmodes, int monitor, cudaStream_t st = nullptr,
                       bool bt_from_g = false) {
    if (nmodes <= 0) return 0;
    if (!st) st = h->stream;
    const int threads = env_int("TK_GRAM_THREADS", 256) == 512 ? 512 : 256, nwarp = threads / 32;
    // warps per column: long columns are cut into segments so all warps of a CTA stream the same amount
    int wpc = h->n >= 8192 ? 8 : h->n >= 4096 ? 4 : h->n >= 2048 ? 2 : 1;
    if (env_int("TK_GRAM_WPC", 0)) wpc = env_int("TK_GRAM_WPC", 0);
    wpc = std::min(wpc, nwarp);
    const int maxcpc = env_int("TK_GRAM_CPC", 32);
    const int gran = nwarp / wpc;                    // columns a CTA processes per pass
    // chunks per mode: at most maxcpc columns per CTA (the new vector is re-staged once per CTA); among the
    // admissible counts pick the one whose grid fills whole waves of 2 CTAs x 148 SMs best
    const int lo_ch = (ncols + maxcpc - 1) / maxcpc, hi_ch = std::max(lo_ch, std::min((ncols + gran - 1) / gran, 48));
    int nchunks = lo_ch;
    double best = -1.0;
    for (int c = lo_ch; c <= hi_ch; ++c) {
        const double ctas = (double)c * nmodes, waves = std::ceil(ctas / 296.0);
        const int cpc_c = (ncols + c - 1) / c;
        const double score = ctas / (waves * 296.0) / (1.0 + 1.0 / cpc_c);
        if (score > best + 1e-9) { best = score; nchunks = c; }
    }
    if (env_int("TK_GRAM_CHUNKS", 0)) nchunks = env_int("TK_GRAM_CHUNKS", 0);
    int cpc = (ncols + nchunks - 1) / nchunks;
    nchunks = (ncols + cpc - 1) / cpc;
    const size_t fixed = ((size_t)cpc * GRAM_PSTRIDE + ((h->ncol + 1) & ~1)) * 8;
    size_t smem = fixed + (size_t)h->n * 8;
    // new vector through L1 (from 64 modes per GPU on) or staged in shared memory; TK_GRAM_WSMEM = 0 / 1 force
    const int wsm_env = env_int("TK_GRAM_WSMEM", -1);
    const bool w_smem = smem <= smem_limit(h) && (wsm_env >= 0 ? wsm_env != 0 : h->dk < 64);
    if (!w_smem) {
        smem = fixed;
        if (monitor > 0 && !h->vscratch.p) TK_TRY(h->vscratch.alloc((size_t)h->dk * h->ldv));
    }
    const double bytes = 8.0 * (double)h->n * (double)ncols * nmodes;
    const int U = env_int("TK_GRAM_U", 4);
    TimedScope ts(h, TM_GRAM, bytes, st);
#define TK_GRAM_LAUNCH(UU, TT)                                                                                   \
    do {                                                                                                         \
        TK_TRY(allow_smem(gram_row_kernel<UU, TT>, smem));                                                       \
        gram_row_kernel<UU, TT><<<dim3(nchunks, nmodes), TT, smem, st>>>(h->kp(), ncols, cpc, base,       \
                                                                                w_smem ? 1 : 0, wpc, monitor,    \
                                                                                h->tickets.p, h->vscratch.p,     \
                                                                                bt_from_g ? 1 : 0);              \
    } while (0)
    if (threads == 512) {
        if (U == 8) TK_GRAM_LAUNCH(8, 512); else if (U == 2) TK_GRAM_LAUNCH(2, 512); else TK_GRAM_LAUNCH(4, 512);
    } else {
        if (U == 8) TK_GRAM_LAUNCH(8, 256); else if (U == 2) TK_GRAM_LAUNCH(2, 256); else TK_GRAM_LAUNCH(4, 256);
    }
#undef TK_GRAM_LAUNCH
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return 0;
}

static int mgs_smem(tk_handle* h, size_t* smem, double** vscr) {
    size_t need = ((size_t)h->ncol + (size_t)h->n) * 8;
    if (need <= smem_limit(h)) {
        *smem = need; *vscr = nullptr;
    } else {
        if (!h->vscratch.p) TK_TRY(h->vscratch.alloc((size_t)h->dk * h->ldv));
        *smem = (size_t)h->ncol * 8; *vscr = h->vscratch.p;
    }
    return 0;
}

static int launch_arnoldi(tk_handle* h, int k) {
    const double bytes = (16.0 * k + op_bytes_per_row(h) + 24.0) * (double)h->n * h->dk;
    TimedScope ts(h, TM_MGS, bytes, h->stream);
    KrylovParams p = h->kp();
    const size_t hsm = (size_t)h->ncol * 8;
    const bool reg_ok = env_int("TK_MGS_REG", 1) != 0;
#define TK_MGS_LAUNCH(E, TH, PF)                                                      \
    do {                                                                              \
        arnoldi_mgs_reg_kernel<E, TH, PF><<<h->dk, TH, hsm, h->stream>>>(p, k);       \
        h->launches++;                                                                \
        TK_CUDA(cudaGetLastError());                                                  \
        return 0;                                                                     \
    } while (0)
    // blocked sweep (4 columns per reduction round) where the registers allow it; TK_MGS_BLOCK=0 keeps strict MGS
#define TK_BGS_LAUNCH(E, TH, BB)                                                      \
    do {                                                                              \
        arnoldi_bgs_kernel<E, TH, BB><<<h->dk, TH, hsm, h->stream>>>(p, k);           \
        h->launches++;                                                                \
        TK_CUDA(cudaGetLastError());                                                  \
        return 0;                                                                     \
    } while (0)
    if (reg_ok && hsm <= 40 * 1024 && env_int("TK_MGS_BLOCK", 1)) {
        if (h->n <= 128 * 4) TK_BGS_LAUNCH(4, 128, 4);
        if (h->n <= 256 * 4) TK_BGS_LAUNCH(4, 256, 4);
        // one CTA per mode leaves the SM with few warps: more, thinner threads hide the issue latency of the sweep
        const int bth = env_int("TK_BGS_THREADS", 256);
        if (h->n <= 1024 * 2 && bth == 1024) TK_BGS_LAUNCH(2, 1024, 4);
        if (h->n <= 512 * 4 && bth >= 512) TK_BGS_LAUNCH(4, 512, 4);
        if (h->n <= 256 * 8) TK_BGS_LAUNCH(8, 256, 4);
        if (h->n <= 512 * 8) TK_BGS_LAUNCH(8, 512, 2);     // 128 registers per thread at 512 threads: two columns per round
    }
#undef TK_BGS_LAUNCH
    if (reg_ok && hsm <= 40 * 1024) {
        if (h->n <= 128 * 4) TK_MGS_LAUNCH(4, 128, 3);
        if (h->n <= 256 * 4) TK_MGS_LAUNCH(4, 256, 3);
        if (h->n <= 256 * 8) TK_MGS_LAUNCH(8, 256, 3);
        if (h->n <= 512 * 8) TK_MGS_LAUNCH(8, 512, 3);
        if (h->n <= 512 * 12) TK_MGS_LAUNCH(12, 512, 1);
        if (h->n <= 1024 * 10) TK_MGS_LAUNCH(10, 1024, 0);
    }
#undef TK_MGS_LAUNCH
    size_t smem; double* vscr;
    TK_TRY(mgs_smem(h, &smem, &vscr));
    TK_TRY(allow_smem(arnoldi_mgs_kernel, smem));
    arnoldi_mgs_kernel<<<h->dk, 512, smem, h->stream>>>(p, k, vscr);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return 0;
}

// orthonormalize!(decomp, k) for every local mode + update_rhs!   (orthogonal_bases.jl:162-180, utils.jl:466-476)
static int enqueue_step_bases(tk_handle* h, int k) {
    const int mode0 = (h->first == 0 && h->dl > 0) ? 1 : 0;
    if (h->variant == TK_ARNOLDI) {
        TK_TRY(launch_arnoldi(h, k));
        TK_TRY(launch_gram(h, k + 1, 0, mode0, 0));
    } else {
        TK_TRY(launch_ttr(h, k));
        if (h->variant == TK_LANCZOS_REORTH) {
            TK_TRY(launch_gram(h, k + 1, 0, h->dk, 1, nullptr, h->gram_does_bt));
        } else {
            TK_TRY(launch_gram(h, k + 1, 0, mode0, 0));
        }
    }
    return 0;
}

// Kernel (2): bisection + twisted factorisation first; the QL kernel runs right behind it on the problems the first
// one flags (tight clusters, degenerate spectra).  TK_EIG_MODE=1 forces QL for everything.
static int launch_eig(const double* T, long long tstride, int ncol, int k, int nprob, double* theta,
                      int thstride, double* Q, long long qstride, int ldq, const int* status, int* fail, double* scratch,
                      int* need, cudaStream_t st, long long* launches) {
    if (k > 2048) return set_error(TK_EUNSUPPORTED, "eigensolver supports k <= 2048");
    const bool bisect = env_int("TK_EIG_MODE", 0) == 0 && k <= 1024 && scratch && need;
    if (bisect) {
        // threads per eigenvalue during the multisection: as many as a 1024-thread CTA allows, at most 8
        int tpe = 8;
        while (tpe > 1 && k * tpe > 1024) tpe >>= 1;
        const int threads = ((k * tpe + 31) / 32) * 32;
        const size_t smem = ((size_t)5 * k + 2) * 8;
#define TK_BISECT_LAUNCH(MT)                                                                                          \
        do {                                                                                                          \
            TK_TRY(allow_smem(tridiag_eig_bisect_kernel<MT>, smem));                                                  \
            tridiag_eig_bisect_kernel<MT><<<nprob, threads, smem, st>>>(T, tstride, ncol, k, tpe, theta, thstride, Q, \
                                                                      scratch, qstride, ldq, status, need);           \
        } while (0)
        if (threads <= 256) TK_BISECT_LAUNCH(256);
        else if (threads <= 512) TK_BISECT_LAUNCH(512);
        else TK_BISECT_LAUNCH(1024);
#undef TK_BISECT_LAUNCH
        TK_CUDA(cudaGetLastError());
        if (launches) ++*launches;
    }
    // rows of Q per CTA: as many as fit in shared memory next to the per-warp (d, e) copies
    const size_t budget = 200 * 1024;
    int rows = std::min(256, ((k + 31) / 32) * 32);
    while (rows > 8) {
        const int nwarp = (rows + 31) / 32;
        if ((size_t)nwarp * 2 * k * 8 + (size_t)k * rows * 8 <= budget) break;
        rows = rows > 32 ? rows - 32 : rows / 2;
    }
    const int nwarp = (rows + 31) / 32;
    const int ldz = rows;
    const size_t smem = (size_t)nwarp * 2 * k * 8 + (size_t)k * ldz * 8;
    if (smem > budget) return set_error(TK_EUNSUPPORTED, "eigensolver: k = %d does not fit", k);
    TK_TRY(allow_smem(tridiag_eig_kernel, smem));
    const int nrb = (k + rows - 1) / rows;
    tridiag_eig_kernel<<<dim3(nprob, nrb), nwarp * 32, smem, st>>>(T, tstride, ncol, k, rows, ldz, theta, thstride, Q, qstride,
                                                                 ldq, status, fail, bisect ? need : nullptr);
    TK_CUDA(cudaGetLastError());
    if (launches) ++*launches;
    return 0;
}

static CompressParams make_cp(tk_handle* h, int k) {
    const SchedEntry& se = h->sched[k];
    CompressParams c;
    c.k = k; c.t = se.t; c.tld = (se.t + 3) & ~3; c.ncol = h->ncol;
    c.per_mode = h->per_mode;
    const size_t eslot = h->use_expm ? 0 : (size_t)(k % h->ring_depth);
    c.theta = h->theta.p + eslot * h->ncls * h->ncol; c.thstride = h->ncol;
    c.Q = h->Q.p + eslot * h->ncls * h->ldq * h->ldq; c.qstride = (long long)h->ldq * h->ldq; c.ldq = h->ldq;
    c.bt = h->bt.p;
    c.alpha = h->alpha_d.p + se.off; c.omega = h->omega_d.p + se.off;
    c.lam_inv = 1.0 / se.lambda_min;
    c.Y = h->Y.p; c.Z = h->Z.p; c.ystride = h->ystride;
    c.T = h->T.p; c.Hd = h->Hd.p;
    c.E = h->E.p; c.estride = h->estride;
    c.bb = h->bbm.p;
    c.status = h->status_d.p;
    return c;
}

static int enqueue_expm(tk_handle* h, int k, cudaStream_t st);

// solve_compressed_system (tensor_krylov_method.jl:10-34), first half: the eigendecomposition(s), or for
// NonSymInstance the dense exponentials exp(gamma_j H)
static int enqueue_eig(tk_handle* h, int k) {
    if (h->use_expm) return enqueue_expm(h, k, h->stream3[k % tk_handle::NEIG]);
    // under TK_FLAG_REFERENCE_H1 the one problem is mode 1's H (this rank's own copy or its shadow copy)
    const double* Tsrc = h->per_mode ? h->T.p : h->T.p + (size_t)h->eig_slot * 3 * h->ncol;
    CompressParams c = make_cp(h, k);
    cudaStream_t st = h->stream3[k % tk_handle::NEIG];
    TimedScope ts(h, TM_EIG, 0.0, st);
    const size_t ring = (size_t)(k % h->ring_depth);
    TK_TRY(launch_eig(Tsrc, 3LL * h->ncol, h->ncol, k, h->ncls, const_cast<double*>(c.theta), h->ncol, const_cast<double*>(c.Q),
                      c.qstride, h->ldq, h->status_d.p, h->eigfail_d.p, h->eig_scratch.p + ring * h->ncls * h->ldq * h->ldq,
                      h->eig_need.p + ring * h->ncls, st, &h->launches));
    return 0;
}

// NonSymInstance: Y_s[:,j] = exp(gamma_j H) b~_s through the batched Taylor scaling-and-squaring exponential
static ExpmParams make_ex(tk_handle* h, int k) {
    const SchedEntry& se = h->sched[k];
    const size_t ring = (size_t)(k % h->ring_depth), per = (size_t)h->ncls * std::max(h->tmax, 1);
    ExpmParams p;
    p.k = k; p.ld = h->ex_ld; p.t = se.t; p.ncls = h->ncls; p.ncol = h->ncol;
    p.mslot = (long long)h->ex_ld * h->ex_ld;
    p.W = h->exW.p + ring * per * EX_SLOTS * (size_t)p.mslot;
    p.nsq = h->ex_nsq.p + ring * per; p.where = h->ex_where.p + ring * per;
    p.Hd = h->Hd.p; p.T = h->T.p; p.cls_mode = h->cls_mode_d.p;
    p.alpha = h->alpha_d.p + se.off; p.lam_inv = 1.0 / se.lambda_min;
    p.status = h->status_d.p;
    return p;
}

// Y_s[:, j] = E_j b~_s on the assembly stream, once the exponentials of iteration k are there
static int enqueue_expm_apply(tk_handle* h, int k) {
    const SchedEntry& se = h->sched[k];
    ExpmParams p = make_ex(h, k);
    const int tld = (se.t + 3) & ~3;
    if (h->dl > 0) {
        TimedScope ts(h, TM_ASM, 0.0, h->stream2);
        expm_apply_kernel<<<dim3(h->dl, se.t), 256, (size_t)k * 8, h->stream2>>>(p, h->per_mode, h->bt.p, h->Y.p, h->ystride, tld);
        h->launches++;
        TK_CUDA(cudaGetLastError());
    }
    h->last_k = k; h->last_t = se.t; h->last_tld = tld; h->last_lam_inv = p.lam_inv;
    return 0;
}

static int enqueue_expm(tk_handle* h, int k, cudaStream_t st) {
    const SchedEntry& se = h->sched[k];
    ExpmParams p = make_ex(h, k);
    const int nmat = h->ncls * se.t;
    TimedScope ts(h, TM_EIG, 0.0, st);
    double f[17];
    f[0] = 1.0;
    for (int i = 1; i <= 16; ++i) f[i] = f[i - 1] / (double)i;    // 1/i!
    const int tiles_f = (k + 63) / 64;
    if (tiles_f <= 4 && !env_int("TK_EXPM_UNFUSED", 0)) {
        // one launch: a cluster of tiles^2 CTAs per matrix walks through all products (cluster barriers in between)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)nmat * tiles_f * tiles_f);
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = tiles_f * tiles_f; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = tiles_f > 1 ? 1 : 0;
#define TK_EXPM_FUSED(TL)                                                                                             \
        do {                                                                                                          \
            if (TL * TL > 8) TK_CUDA(cudaFuncSetAttribute(expm_fused_kernel<TL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); \
            TK_CUDA(cudaLaunchKernelEx(&cfg, expm_fused_kernel<TL>, p, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8],      \
                                       f[9], f[10], f[11], f[12], f[13], f[14], f[15], f[16]));                      \
        } while (0)
        if (tiles_f == 1) TK_EXPM_FUSED(1); else if (tiles_f == 2) TK_EXPM_FUSED(2); else if (tiles_f == 3) TK_EXPM_FUSED(3); else TK_EXPM_FUSED(4);
#undef TK_EXPM_FUSED
        h->launches++;
        return 0;
    }
    expm_setup_kernel<<<nmat, 256, 0, st>>>(p);
    h->launches++;
    // upper bound on the squarings: ||gamma H||_1 <= |gamma| sqrt(k) ||H||_2 <= |gamma| sqrt(k) ||A||_2
    double amax = 0.0, nb = 0.0;
    for (int j = 0; j < se.t; ++j) amax = std::max(amax, std::fabs(h->alpha_pool[se.off + j]));
    for (int s = 0; s < h->dk; ++s) nb = std::max(nb, h->ops[h->mode_op[s]]->norm_bound);
    const double bound = amax * p.lam_inv * std::sqrt((double)k) * nb;
    int smax = bound > EX_THETA ? (int)std::ceil(std::log2(bound / EX_THETA)) + 1 : 1;
    smax = std::min(std::max(smax, 1), 60);
    const int tiles = (k + 63) / 64;
    dim3 grid(tiles, tiles, nmat);
    auto gemm = [&](int a, int b, int c, int comb, double c0, double c1, double c2, double c3, double c4, int sq) {
        GemmJob job; job.a = a; job.b = b; job.c = c; job.comb = comb; job.c0 = c0; job.c1 = c1; job.c2 = c2; job.c3 = c3;
        job.c4 = c4; job.sq_step = sq;
        expm_gemm_kernel<<<grid, 256, 0, st>>>(p, job);
        h->launches++;
    };
    gemm(0, 0, 1, 0, 0, 0, 0, 0, 0, -1);     // A2 = A A
    gemm(1, 0, 2, 0, 0, 0, 0, 0, 0, -1);     // A3 = A2 A
    gemm(1, 1, 3, 0, 0, 0, 0, 0, 0, -1);     // A4 = A2 A2
    expm_top_kernel<<<dim3(std::max(1, (h->ex_ld * h->ex_ld + 255) / 256 / 4), nmat), 256, 0, st>>>(p, f[12], f[13], f[14], f[15], f[16]);
    h->launches++;
    gemm(3, 5, 4, 1, f[8], f[9], f[10], f[11], 0.0, -1);
    gemm(3, 4, 5, 1, f[4], f[5], f[6], f[7], 0.0, -1);
    gemm(3, 5, 4, 1, f[0], f[1], f[2], f[3], 0.0, -1);
    for (int sq = 0; sq < smax; ++sq) gemm(0, 0, 0, 0, 0, 0, 0, 0, 0, sq);
    TK_CUDA(cudaGetLastError());
    return 0;
}

// second half: CP assembly of Y_s for every local mode
static int enqueue_assemble(tk_handle* h, int k) {
    if (h->use_expm) return enqueue_expm_apply(h, k);
    CompressParams c = make_cp(h, k);
    const size_t smem = ((size_t)2 * k + (size_t)k * ASM_TJ) * 8;
    TK_TRY(allow_smem(assemble_cp_kernel, smem));
    if (h->dl > 0) {
        TimedScope ts(h, TM_ASM, 0.0, h->stream2);
        assemble_cp_kernel<<<h->dl, 256, smem, h->stream2>>>(c);     // also forms Z and the Gram blocks of the mode
        h->launches++;
        TK_CUDA(cudaGetLastError());
    }
    h->last_k = k; h->last_t = c.t; h->last_tld = c.tld; h->last_lam_inv = c.lam_inv;
    return 0;
}

// residualnorm! (utils.jl:402-443) + exits of the loop body (tensor_krylov_method.jl:85-118)
static int enqueue_residual(tk_handle* h, int k) {
    CompressParams c = make_cp(h, k);
    if (h->dl > 0 && h->use_expm) {      // the symmetric path did this inside assemble_cp_kernel
        TimedScope ts(h, TM_ASM, 0.0, h->stream2);
        if ((long long)h->dl * 4 < h->sm_count && (long long)c.t * c.t > 1024 && env_int("TK_GRAM_SPLIT", 1)) {
            // few modes, many terms: spread Z and the Gram blocks of a mode over many CTAs (two launches)
            gram_z_kernel<<<dim3(h->dl, (k * c.t + 255) / 256), 256, 0, h->stream2>>>(c);
            gram_e_kernel<<<dim3(h->dl, (c.t * c.t + 255) / 256), 256, 0, h->stream2>>>(c);
            h->launches += 2;
        } else {
            gram_blocks_kernel<<<h->dl, 256, 0, h->stream2>>>(c);
            h->launches++;
        }
        TK_CUDA(cudaGetLastError());
    }
    const long long pst = 5LL * c.t * c.t + 2LL * c.t + 8;
    FinalizeParams f;
    f.k = k; f.t = c.t; f.nmax = h->nmax; f.nparts = 1;
    f.fixed_iterations = (h->flags & TK_FLAG_FIXED_ITERATIONS) ? 1 : 0;
    f.pstride = pst; f.partials = h->merged.p; f.omega = c.omega;
    f.lam_inv = c.lam_inv; f.lambda_min = h->sched[k].lambda_min;
    f.ctl = h->ctl_d.p; f.hctl = h->hctl_dev;
    f.relres = h->relres_d.p; f.projres = h->projres_d.p; f.orth = h->orth_d.p; f.detail = h->detail_d.p;
    f.status = h->status_d.p; f.niter = h->niter_d.p; f.term_k = h->term_k_d.p;
    // one GPU, or several with the peer exchange: the last CTA of the combine kernel also does the final merge
    const int fin_here = (h->world == 1 || h->px_ready) ? 1 : 0;
    PeerExchange px = h->px;
    if (!h->px_ready) px.world = 1;
    {
        TimedScope ts(h, TM_COMBINE, 0.0, h->stream2);
        combine_chunk_kernel<<<h->nchunks, 256, 0, h->stream2>>>(c, h->dl, h->chunk_modes, h->chunk_base, h->partials.p, pst,
                                                                h->orthS.p, (h->first == 0 && h->dl > 0) ? 0 : -1,
                                                                h->bnorm2.p, h->ticket_d.p, h->merged.p, fin_here, f, px);
        h->launches++;
        TK_CUDA(cudaGetLastError());
    }
    if (fin_here) return 0;
    // fallback without peer-mapped memory: NCCL all-gather of the merged partials, then the final merge
    TK_NCCL(g_nccl.AllGather(h->merged.p, h->gathered.p, (size_t)pst, ncclDouble, h->comm, h->stream2));
    f.partials = h->gathered.p;
    f.nparts = h->world;
    finalize_kernel<<<1, 256, 0, h->stream2>>>(f);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    return 0;
}

// Peer exchange (world > 1): every rank exports its receive buffer and its flag array through CUDA IPC, the
// 2 x 64-byte handles travel once over the existing NCCL communicator, and every rank maps everybody else's.  All
// ranks must agree on the path, so the outcome is voted on (a second all-gather); if any rank could not map a
// peer the solve keeps NCCL's all-gather.  Collective: called by all ranks from the first tk_solve / tk_compress.
static int setup_peer_exchange(tk_handle* h) {
    NvtxRange range("tk peer exchange setup");
    h->px_ready = false;
    h->xc = nullptr;
    if (h->world > PX_MAX || !env_int("TK_PEER", 1)) return 0;
    const long long slot = (h->pstride_max + 1) & ~1LL;
    const auto key = std::make_pair(h->comm_key, slot);
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = g_exchange.find(key);
        if (it != g_exchange.end()) {        // set up by an earlier handle (on every rank alike): nothing collective left
            h->xc = it->second.get();
            h->px = h->xc->px;
            h->px_ready = h->xc->ready;
            if (!h->px_ready && env_int("TK_PEER", 1) == 2)
                return set_error(TK_ECUDA, "peer mapping of the exchange buffers failed (TK_PEER=2 forbids the NCCL fallback)");
            return 0;
        }
    }
    std::unique_ptr<ExchangeCtx> xc(new ExchangeCtx());
    TK_TRY(xc->recv.alloc((size_t)2 * h->world * slot));
    TK_TRY(xc->flags.alloc((size_t)h->world));
    struct Rec { cudaIpcMemHandle_t recv, flag; int ok; int pad[15]; };
    static_assert(sizeof(Rec) % 8 == 0, "record is shipped as doubles");
    Rec mine;
    std::memset(&mine, 0, sizeof(mine));
    mine.ok = cudaIpcGetMemHandle(&mine.recv, xc->recv.p) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.flag, xc->flags.p) == cudaSuccess;
    cudaGetLastError();
    DevBuf<double> ex;
    const size_t rd = sizeof(Rec) / 8;
    TK_TRY(ex.alloc(rd * (h->world + 1)));
    std::vector<Rec> all(h->world);
    auto gather = [&]() -> int {
        TK_CUDA(cudaMemcpyAsync(ex.p + rd * h->world, &mine, sizeof(Rec), cudaMemcpyHostToDevice, h->stream));
        TK_NCCL(g_nccl.AllGather(ex.p + rd * h->world, ex.p, rd, ncclDouble, h->comm, h->stream));
        TK_CUDA(cudaMemcpyAsync(all.data(), ex.p, sizeof(Rec) * h->world, cudaMemcpyDeviceToHost, h->stream));
        TK_CUDA(cudaStreamSynchronize(h->stream));
        return 0;
    };
    TK_TRY(gather());
    PeerExchange px;
    std::memset(&px, 0, sizeof(px));
    px.world = h->world; px.rank = h->rank; px.slot_stride = slot;
    bool ok = true;
    for (int r = 0; r < h->world; ++r) ok = ok && all[r].ok;
    // Mappings are kept for the life of the process, keyed by the exported handle (opening one costs ~0.1 ms).
    auto open_cached = [&](const cudaIpcMemHandle_t& mh, void** out) -> bool {
        const std::string k2(reinterpret_cast<const char*>(&mh), sizeof(mh));
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = g_ipc_open.find(k2);
        if (it != g_ipc_open.end()) { *out = it->second; return true; }
        if (cudaIpcOpenMemHandle(out, mh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return false;
        g_ipc_open[k2] = *out;
        return true;
    };
    for (int r = 0; r < h->world && ok; ++r) {
        if (r == h->rank) { px.recv[r] = xc->recv.p; px.flag[r] = xc->flags.p; continue; }
        void *pr = nullptr, *pf = nullptr;
        if (!open_cached(all[r].recv, &pr) || !open_cached(all[r].flag, &pf)) { ok = false; break; }
        px.recv[r] = static_cast<double*>(pr);
        px.flag[r] = static_cast<unsigned long long*>(pf);
    }
    cudaGetLastError();
    mine.ok = ok ? 1 : 0;
    TK_TRY(gather());                       // vote: also orders every rank's mapping before anybody's first store
    for (int r = 0; r < h->world; ++r) ok = ok && all[r].ok;
    cudaGetLastError();
    xc->px = px;
    xc->ready = ok;
    h->px = px;
    h->px_ready = ok;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        h->xc = xc.get();
        g_exchange[key] = std::move(xc);
    }
    if (!ok && env_int("TK_PEER", 1) == 2)
        return set_error(TK_ECUDA, "peer mapping of the exchange buffers failed (TK_PEER=2 forbids the NCCL fallback)");
    return 0;
}

// Everything a recorded segment carries as a launch argument or used to choose a kernel: device pointers, workspace
// strides, operator structure, the schedule (term counts, offsets, lambda_min, coefficients) and the TK_* knobs.  A
// recorded graph is replayed only under the signature it was recorded with.
static unsigned long long config_signature(const tk_handle* h) {
    unsigned long long x = 1469598103934665603ULL;
    auto mix = [&](const void* p, size_t nbytes) {
        const unsigned char* c = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < nbytes; ++i) { x ^= c[i]; x *= 1099511628211ULL; }
    };
    auto ptr = [&](const void* p) { mix(&p, sizeof(p)); };
    ptr(h->V.p); ptr(h->b.p); ptr(h->T.p); ptr(h->Hd.p); ptr(h->bt.p); ptr(h->g.p); ptr(h->S.p); ptr(h->orthS.p);
    ptr(h->bnorm2.p); ptr(h->vscratch.p); ptr(h->fallbacks.p); ptr(h->mode_op_d.p); ptr(h->status_d.p);
    ptr(h->term_k_d.p); ptr(h->eigfail_d.p); ptr(h->niter_d.p); ptr(h->ops_d.p); ptr(h->alpha_d.p); ptr(h->omega_d.p);
    ptr(h->theta.p); ptr(h->Q.p); ptr(h->Y.p); ptr(h->Z.p); ptr(h->E.p); ptr(h->bbm.p); ptr(h->partials.p);
    ptr(h->gathered.p); ptr(h->relres_d.p); ptr(h->projres_d.p); ptr(h->orth_d.p); ptr(h->detail_d.p);
    ptr(h->eig_scratch.p); ptr(h->eig_need.p); ptr(h->ticket_d.p); ptr(h->tickets.p); ptr(h->merged.p);
    ptr(h->exW.p); ptr(h->ex_nsq.p); ptr(h->ex_where.p); ptr(h->cls_mode_d.p); ptr(h->ctl_d.p); ptr(h->hctl_dev);
    mix(&h->px, sizeof(h->px));
    const long long ints[] = {h->px_ready, h->ring_depth, h->ldq, h->ystride, h->estride, h->pstride_max, h->ex_ld, h->tmax};
    mix(ints, sizeof(ints));
    for (const auto& op : h->ops) {
        const long long m[] = {op->type, op->ndiag, op->ld, op->nnz, op->constd};
        mix(m, sizeof(m)); mix(op->offs, sizeof(op->offs)); mix(op->cval, sizeof(op->cval)); mix(&op->norm_bound, 8);
    }
    mix(h->mode_op.data(), h->mode_op.size() * sizeof(int));
    for (size_t k = 2; k < h->sched.size(); ++k) {
        const SchedEntry& se = h->sched[k];
        const long long m[] = {se.set, se.t, (long long)se.off};
        mix(m, sizeof(m)); mix(&se.lambda_min, 8);
    }
    mix(h->alpha_pool.data(), h->alpha_pool.size() * 8);
    mix(h->omega_pool.data(), h->omega_pool.size() * 8);
    for (char** e = environ; e && *e; ++e)
        if (std::strncmp(*e, "TK_", 3) == 0) mix(*e, std::strlen(*e));
    return x ? x : 1;
}

static int prepare(tk_handle* h, bool with_schedule) {
    TK_TRY(upload_ops(h));
    for (int s = 0; s < h->dk; ++s)
        if (!h->rhs_set[s]) return set_error(TK_ESTATE, "right-hand side of mode %d not set", s < h->dl ? h->first + s : 0);
    if (with_schedule) {
        TK_TRY(upload_schedule(h));
        TK_TRY(alloc_work(h));
    }
    // working vector of the MGS step when it does not fit in shared memory
    const size_t need_gram = ((size_t)32 * GRAM_PSTRIDE + ((h->ncol + 1) & ~1) + (size_t)h->n) * 8;
    const size_t need_mgs = ((size_t)h->ncol + (size_t)h->n) * 8;
    // ... or when the Gram-row kernel keeps the new vector in L1 instead of shared memory (launch_gram)
    const int wsm_env = env_int("TK_GRAM_WSMEM", -1);
    const bool gram_l1 = wsm_env >= 0 ? wsm_env == 0 : h->dk >= 64;
    if ((need_gram > smem_limit(h) || need_mgs > smem_limit(h) || gram_l1) && !h->vscratch.p) {
        TK_TRY(h->vscratch.alloc((size_t)h->dk * h->ldv));
    }
    h->sig = config_signature(h);
    return 0;
}

// Values the recorded launches read at run time instead of carrying them as arguments
static void arm_solve(tk_handle* h, double tol) {
    h->solve_count++;
    h->hctl->tol = tol;
    if (h->xc && h->px_ready) {            // flags of the shared exchange buffers grow across handles
        h->xc->epoch += h->nmax + 2;
        h->hctl->epoch = h->xc->epoch;
    } else {
        h->hctl->epoch = h->solve_count * (long long)(h->nmax + 2);
    }
    h->hctl->term_k = 0;
    h->hctl->niter = h->nmax;
    *reinterpret_cast<volatile int*>(&h->hctl->status) = ST_RUNNING;
    for (int i = 0; i < 8; ++i) { h->step_slot[i].k = -1; h->eigdone_slot[i].k = -1; h->asm_slot[i].k = -1; }
    h->ev_used = 0;
    h->timed.clear();
    h->launches = 0;
    h->graph_build_ms = 0.0;
    h->graphs_launched = 0;
    for (int i = 0; i < TM_KINDS; ++i) { h->tm_ms[i] = 0; h->tm_bytes[i] = 0; h->tm_launches[i] = 0; }
}

// orthonormalize!(decomp, b), initialize_compressed_rhs, kronprodnorm   (tensor_krylov_method.jl:48-55).
// No host round trip: kronprodnorm(b) = sqrt(prod_s b_s.b_s) travels with the cross-mode partials (bnorm2 -> combine).
static int enqueue_begin(tk_handle* h) {
    ResetParams r;
    r.status4 = h->status_d.p; r.term_k = h->term_k_d.p; r.eigfail = h->eigfail_d.p; r.niter = h->niter_d.p;
    r.relres = h->relres_d.p; r.projres = h->projres_d.p; r.orth = h->orth_d.p; r.nmax = h->nmax;
    r.tickets = h->tickets.p; r.ntickets = (int)h->tickets.count; r.ticket1 = h->ticket_d.p;
    r.ctl = h->ctl_d.p; r.hctl = h->hctl_dev;
    reset_kernel<<<1, 256, 0, h->stream>>>(r);
    h->launches++;
    TK_CUDA(cudaGetLastError());
    TK_CUDA(cudaMemsetAsync(h->T.p, 0, 8 * h->T.count, h->stream));
    TK_CUDA(cudaMemsetAsync(h->bt.p, 0, 8 * h->bt.count, h->stream));
    TK_CUDA(cudaMemsetAsync(h->detail_d.p, 0, 8 * h->detail_d.count, h->stream));
    if (h->Hd.p) TK_CUDA(cudaMemsetAsync(h->Hd.p, 0, 8 * h->Hd.count, h->stream));
    if (h->dk > 0) {
        init_basis_kernel<<<h->dk, 512, 0, h->stream>>>(h->kp(), h->bnorm2.p);
        h->launches++;
        TK_CUDA(cudaGetLastError());
    }
    // Gram "row" of column 1 starts the orthogonality bookkeeping, then step k = 1
    const int mode0 = (h->first == 0 && h->dl > 0) ? 1 : 0;
    const int nmon = h->variant == TK_LANCZOS_REORTH ? h->dk : mode0;
    TK_TRY(launch_gram(h, 1, 0, nmon, 0));
    TK_TRY(enqueue_step_bases(h, 1));
    return 0;
}

static int begin_solve(tk_handle* h, double tol) {
    TK_TRY(prepare(h, false));
    arm_solve(h, tol);
    TK_TRY(enqueue_begin(h));
    h->begun = true;
    return 0;
}

// ---- dependencies inside a segment -------------------------------------------------------
static int slot_record(tk_handle* h, tk_handle::EvSlot* ring, int k, cudaStream_t st) {
    tk_handle::EvSlot& sl = ring[k & 7];
    TK_CUDA(cudaEventRecord(sl.ev, st));
    sl.k = k; sl.seg = h->cur_seg;
    return 0;
}
static int slot_wait(tk_handle* h, tk_handle::EvSlot* ring, int k, cudaStream_t st) {
    const tk_handle::EvSlot& sl = ring[k & 7];
    if (sl.k == k && sl.seg == h->cur_seg) TK_CUDA(cudaStreamWaitEvent(st, sl.ev, 0));
    return 0;   // recorded in an earlier segment: complete before this one started
}

// eigensolve -> CP assembly -> residual estimate of iteration k on the side streams
static int enqueue_chain(tk_handle* h, int k) {
    cudaStream_t es = h->stream3[k % tk_handle::NEIG];

    TK_TRY(slot_wait(h, h->step_slot, k, es));                             // needs the Krylov step k
    if (k - h->ring_depth >= 2) TK_TRY(slot_wait(h, h->asm_slot, k - h->ring_depth, es));   // its ring buffer is free
    TK_TRY(enqueue_eig(h, k));
    TK_TRY(slot_record(h, h->eigdone_slot, k, es));
    TK_TRY(slot_wait(h, h->eigdone_slot, k, h->stream2));
    TK_TRY(enqueue_assemble(h, k));
    TK_TRY(slot_record(h, h->asm_slot, k, h->stream2));
    TK_TRY(enqueue_residual(h, k));
    return 0;
}

static std::vector<cudaStream_t> side_streams(const tk_handle* h) {
    std::vector<cudaStream_t> out;
    auto add = [&](cudaStream_t st) {
        if (st == h->stream) return;
        for (auto q : out) if (q == st) return;
        out.push_back(st);
    };
    add(h->stream2);
    for (auto st : h->stream3) add(st);
    return out;
}

static int enqueue_segment(tk_handle* h, int idx) {
    const tk_handle::Segment& sg = h->segs[idx];
    h->cur_seg = idx;
    if (sg.first) TK_TRY(enqueue_begin(h));
    const std::vector<cudaStream_t> side = side_streams(h);
    TK_CUDA(cudaEventRecord(h->res->ev_fork, h->stream));
    for (auto st : side) TK_CUDA(cudaStreamWaitEvent(st, h->res->ev_fork, 0));
    for (int k = sg.c0; k <= sg.c1 && k < sg.k0; ++k) TK_TRY(enqueue_chain(h, k));      // deferred by the last segment
    for (int k = sg.k0; k <= sg.k1; ++k) {
        TK_TRY(enqueue_step_bases(h, k));
        TK_TRY(slot_record(h, h->step_slot, k, h->stream));
        if (k >= sg.c0 && k <= sg.c1) TK_TRY(enqueue_chain(h, k));
    }
    for (size_t i = 0; i < side.size(); ++i) {
        TK_CUDA(cudaEventRecord(h->res->join_ev[i], side[i]));
        TK_CUDA(cudaStreamWaitEvent(h->stream, h->res->join_ev[i], 0));
    }
    return 0;
}

static void plan_segments(tk_handle* h) {
    if (!h->segs.empty()) return;
    // short segments first (a solve that ends after a few iterations has little enqueued behind its exit), then 16
    // chains deferred across a boundary: 3 in fixed-iteration runs (no exit to detect; the Krylov stream never waits
    // for a chain), 1 otherwise (an exit is seen at most one chain later; the bubble at a boundary stays under one step)
    const bool fixed = (h->flags & TK_FLAG_FIXED_ITERATIONS) != 0;
    const int cap = std::max(1, env_int("TK_SEG", 16)), lag = std::max(0, env_int("TK_SEG_LAG", fixed ? 3 : 1));
    int k = 2, size = fixed ? cap : std::min(cap, 4), count = 0, prev_c1 = 1;
    do {
        tk_handle::Segment sg;
        sg.first = h->segs.empty();
        sg.k0 = k;
        sg.k1 = std::min(h->nmax, k + size - 1);
        const bool last = sg.k1 >= h->nmax;
        sg.c0 = prev_c1 + 1;
        sg.c1 = last ? h->nmax : std::max(sg.c0 - 1, sg.k1 - lag);
        prev_c1 = sg.c1;
        h->segs.push_back(sg);
        k = sg.k1 + 1;
        if (++count >= 2 && size < cap) { size = std::min(cap, size * 2); count = 0; }
    } while (k <= h->nmax);
}

static int launch_segment(tk_handle* h, int idx, bool graph) {
    tk_handle::Segment& sg = h->segs[idx];
    NvtxRange range("tk segment");
    if (!graph) return enqueue_segment(h, idx);
    if (!sg.exec || sg.epoch != h->sig) {
        NvtxRange rec("tk record graph");
        const auto t0 = std::chrono::steady_clock::now();
        if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
        const long long before = h->launches;
        TK_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
        const int rc = enqueue_segment(h, idx);
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        sg.launches = h->launches - before;
        h->launches = before;
        if (rc != 0) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return rc; }
        if (e != cudaSuccess) { cudaGetLastError(); return set_error(TK_ECUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e)); }
        const cudaError_t ei = cudaGraphInstantiate(&sg.exec, g, 0);
        cudaGraphDestroy(g);
        if (ei != cudaSuccess) { sg.exec = nullptr; return set_error(TK_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei)); }
        sg.epoch = h->sig;
        h->graph_build_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    TK_CUDA(cudaGraphLaunch(sg.exec, h->stream));
    h->launches += sg.launches;
    h->graphs_launched++;
    return 0;
}

static int collect_timing(tk_handle* h) {
    for (auto& t : h->timed) {
        float ms = 0.f;
        TK_CUDA(cudaEventElapsedTime(&ms, t.a, t.b));
        h->tm_ms[t.kind] += ms;
        h->tm_bytes[t.kind] += t.bytes;
        h->tm_launches[t.kind] += 1;
    }
    h->timed.clear();
    return 0;
}

}  // namespace tk

// =============================================================================================
extern "C" {

const char* tk_last_error(void) { return tk::g_err; }
int tk_version(void) { return 100; }

int tk_device_count(int* count) {
    if (!count) return set_error(TK_EINVAL, "null count");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return set_error(TK_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    return 0;
}

int tk_comm_unique_id(void* out128) {
    if (!out128) return set_error(TK_EINVAL, "null buffer");
    TK_TRY(nccl_bind());
    ncclUniqueId id;
    TK_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out128, &id, 128);
    return 0;
}

int tk_create(tk_handle** out, int32_t d, const int64_t* n, int32_t nmax, int32_t instance, int32_t matrixclass,
              int32_t variant, int32_t flags, int32_t device, int32_t rank, int32_t world, const void* unique_id) {
    if (!out || !n) return set_error(TK_EINVAL, "null argument");
    *out = nullptr;
    if (d < 1 || nmax < 1) return set_error(TK_EINVAL, "need d >= 1 and nmax >= 1");
    for (int s = 1; s < d; ++s)
        if (n[s] != n[0]) return set_error(TK_EUNSUPPORTED, "all modes must have the same order (n[%d]=%lld, n[0]=%lld)", s, (long long)n[s], (long long)n[0]);
    if (n[0] < 1 || n[0] > (1 << 30)) return set_error(TK_EINVAL, "bad order n = %lld", (long long)n[0]);
    if (nmax > n[0]) return set_error(TK_EINVAL, "nmax = %d exceeds n = %lld", nmax, (long long)n[0]);
    if (instance != TK_SYM && instance != TK_NONSYM) return set_error(TK_EINVAL, "bad instance %d", instance);
    if (variant < TK_LANCZOS || variant > TK_ARNOLDI) return set_error(TK_EINVAL, "bad variant %d", variant);
    if (matrixclass < 0 || matrixclass > TK_GENERIC) return set_error(TK_EINVAL, "bad matrix class %d", matrixclass);
    if (world < 1 || rank < 0 || rank >= world) return set_error(TK_EINVAL, "bad rank/world %d/%d", rank, world);
    if (world > 1 && !unique_id) return set_error(TK_EINVAL, "world > 1 needs the NCCL unique id");
    int ndev = 0;
    TK_TRY(tk_device_count(&ndev));
    if (device < 0 || device >= ndev) return set_error(TK_ECUDA, "CUDA device %d not available (%d visible)", device, ndev);
    TK_CUDA(cudaSetDevice(device));

    char keybuf[256];
    snprintf(keybuf, sizeof(keybuf), "%d|%lld|%d|%d|%d|%d|%d|%d|%d|%d|", d, (long long)n[0], nmax, instance, matrixclass, variant,
             flags, device, rank, world);
    std::string shape_key(keybuf);
    if (world > 1) shape_key.append(static_cast<const char*>(unique_id), 128);
    for (char** e = environ; e && *e; ++e)       // developer knobs are read when a handle and its workspaces are built
        if (std::strncmp(*e, "TK_", 3) == 0) { shape_key.push_back('|'); shape_key.append(*e); }
    {
        tk_handle* found = nullptr;
        {
            std::lock_guard<std::mutex> lock(g_mutex);
            for (size_t i = g_parked.size(); i-- > 0;)
                if (g_parked[i]->shape_key == shape_key) { found = g_parked[i]; g_parked.erase(g_parked.begin() + i); break; }
        }
        if (found) {       // revive: same buffers and recorded graphs, no inputs
            found->ops.clear();
            found->mode_op.assign(found->mode_op.size(), -1);
            found->rhs_set.assign(found->rhs_set.size(), 0);
            found->ops_dirty = true;
            found->sched.assign(nmax + 1, SchedEntry());
            found->alpha_pool.clear(); found->omega_pool.clear();
            found->sched_dirty = true;
            found->a1_lead.clear();
            found->begun = false;
            found->last_k = 0; found->last_t = 0; found->last_tld = 0;
            found->timed.clear(); found->ev_used = 0; found->launches = 0;
            *out = found;
            return 0;
        }
    }

    std::unique_ptr<tk_handle> h(new tk_handle());
    h->shape_key = shape_key;
    h->d = d; h->n = (int)n[0]; h->nmax = nmax; h->ncol = nmax + 1;
    h->instance = instance; h->matrixclass = matrixclass; h->variant = variant; h->flags = flags;
    h->device = device; h->rank = rank; h->world = world;
    h->ldv = ((long long)h->n + 15) & ~15LL;
    // block partition of the modes, aligned to combine chunks so the product order does not depend on world
    const int mc = h->chunk_modes;
    int per = (d + world - 1) / world;
    if (d >= world * mc) per = ((per + mc - 1) / mc) * mc;
    h->first = std::min(d, rank * per);
    h->dl = std::max(0, std::min(per, d - h->first));
    h->chunk_base = h->first % mc;
    h->nchunks = std::max(1, (per + mc - 1) / mc + ((per % mc) && world > 1 ? 1 : 0));
    h->per_mode = (flags & TK_FLAG_REFERENCE_H1) ? 0 : 1;
    // EigValMat has its own method in the reference (utils.jl:525-546): every mode exponentiates its OWN H_s, and on
    // the raw view -- which stops being symmetric after the first MGS fallback -- so it takes the dense exponential
    if (matrixclass == TK_EIGVALMAT) h->per_mode = 1;
    h->use_expm = (instance == TK_NONSYM) || (matrixclass == TK_EIGVALMAT);
    h->ncls = h->per_mode ? std::max(h->dl, 1) : 1;
    const bool shadow = !h->per_mode && h->first != 0;   // every rank advances its own copy of mode 1: no broadcast needed
    h->dk = h->dl + (shadow ? 1 : 0);
    h->eig_slot = shadow ? h->dl : 0;

    TK_CUDA(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
    TK_TRY(acquire_resources(device, &h->res));
    tk_resources* r = h->res;
    h->stream = r->s_main;
    h->stream2 = env_int("TK_SINGLE_STREAM", 0) ? h->stream : r->s_asm;
    // eigensolver streams: 8 for the bisection kernel (one CTA per problem), 4 for the dense exponentials (clusters of
    // CTAs per term: more of them in flight only take SMs from the Arnoldi step -- measured at C4)
    const int neig = env_int("TK_EIG_STREAMS", h->use_expm ? 4 : tk_handle::NEIG);
    for (int i = 0; i < tk_handle::NEIG; ++i) {
        if (env_int("TK_SINGLE_STREAM", 0) || env_int("TK_TWO_STREAMS", 0)) h->stream3[i] = h->stream2;
        else h->stream3[i] = r->s_eig[i % std::max(1, std::min(neig, (int)tk_handle::NEIG))];
    }
    for (int i = 0; i < 8; ++i) {
        h->step_slot[i].ev = r->step_ev[i]; h->eigdone_slot[i].ev = r->eig_ev[i]; h->asm_slot[i].ev = r->asm_ev[i];
    }
    h->ev_solve[0] = r->ev_solve[0]; h->ev_solve[1] = r->ev_solve[1];
    h->ev_region = r->ev_region;
    h->hctl = r->hctl; h->hctl_dev = r->hctl_dev;
    h->ev_pool.swap(r->ev_pool);
    const size_t dl = std::max(h->dk, 1);
    TK_TRY(h->V.alloc(dl * (size_t)h->ncol * h->ldv, false));
    TK_TRY(h->b.alloc(dl * (size_t)h->ldv));
    TK_TRY(h->T.alloc(dl * 3 * (size_t)h->ncol));
    if (variant == TK_ARNOLDI) TK_TRY(h->Hd.alloc(dl * (size_t)h->ncol * h->ncol));
    TK_TRY(h->bt.alloc(dl * (size_t)h->ncol));
    TK_TRY(h->g.alloc(dl * (size_t)h->ncol));
    TK_TRY(h->S.alloc(dl));
    TK_TRY(h->orthS.alloc(h->ncol));
    TK_TRY(h->bnorm2.alloc(dl));
    TK_TRY(h->fallbacks.alloc(dl));
    TK_TRY(h->tickets.alloc(dl));
    TK_TRY(h->mode_op_d.alloc(dl));
    TK_TRY(h->status_d.alloc(4));        // [0] live status word, [1..2] snapshots read by the cluster kernels
    TK_TRY(h->term_k_d.alloc(1));
    TK_TRY(h->eigfail_d.alloc(1));
    TK_TRY(h->niter_d.alloc(1));
    TK_TRY(h->ctl_d.alloc(1));
    TK_TRY(h->relres_d.alloc(nmax));
    TK_TRY(h->projres_d.alloc(nmax));
    TK_TRY(h->orth_d.alloc(nmax));
    TK_TRY(h->detail_d.alloc((size_t)(nmax + 1) * 8));
    h->mode_op.assign(dl, -1);
    h->rhs_set.assign(dl, 0);
    h->sched.assign(nmax + 1, SchedEntry());
    if (world > 1) {
        TK_TRY(nccl_bind());
        const std::string key(static_cast<const char*>(unique_id), 128);
        h->comm_key = key;
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = g_comms.find(key);
        if (it == g_comms.end()) {
            ncclUniqueId id;
            std::memcpy(&id, unique_id, 128);
            ncclComm_t comm;
            TK_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
            it = g_comms.emplace(key, CommEntry{comm, rank, world, 0}).first;
        } else if (it->second.rank != rank || it->second.world != world) {
            return set_error(TK_EINVAL, "unique id already bound to rank %d of %d in this process", it->second.rank, it->second.world);
        }
        it->second.refs++;
        h->comm = it->second.comm;
    }
    *out = h.release();
    return 0;
}

void tk_destroy(tk_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->stream2) cudaStreamSynchronize(h->stream2);
    for (auto st : h->stream3) if (st) cudaStreamSynchronize(st);
    const int cap = env_int("TK_HANDLE_CACHE", 2);
    tk_handle* evicted = nullptr;
    if (cap > 0 && cudaGetLastError() == cudaSuccess) {
        std::lock_guard<std::mutex> lock(g_mutex);
        if ((int)g_parked.size() >= cap) { evicted = g_parked.front(); g_parked.erase(g_parked.begin()); }
        g_parked.push_back(h);
        h = nullptr;
    }
    delete evicted;
    delete h;                           // parks streams, events and the pinned control block for the next handle
}

int tk_release_cache(void) {
    evict_parked_handles();
    {   // the exchange contexts hand their blocks back to the block cache (which takes the lock itself)
        decltype(g_exchange) gone;
        { std::lock_guard<std::mutex> lock(g_mutex); gone.swap(g_exchange); }
    }
    std::lock_guard<std::mutex> lock(g_mutex);
    for (tk_resources* r : g_res_free) {
        cudaSetDevice(r->device);
        cudaStreamDestroy(r->s_main); cudaStreamDestroy(r->s_asm);
        for (auto st : r->s_eig) cudaStreamDestroy(st);
        cudaStreamDestroy(r->s_copy);
        for (int i = 0; i < 8; ++i) {
            cudaEventDestroy(r->step_ev[i]); cudaEventDestroy(r->eig_ev[i]); cudaEventDestroy(r->asm_ev[i]);
        }
        for (auto e : r->seg_ev) cudaEventDestroy(e);
        for (auto e : r->join_ev) cudaEventDestroy(e);
        for (auto e : r->copy_ev) cudaEventDestroy(e);
        for (auto e : r->fill_ev) cudaEventDestroy(e);
        cudaEventDestroy(r->ev_fork);
        cudaEventDestroy(r->ev_solve[0]); cudaEventDestroy(r->ev_solve[1]); cudaEventDestroy(r->ev_region);
        for (auto e : r->ev_pool) cudaEventDestroy(e);
        cudaFreeHost(r->hctl);
        for (auto q : r->stage) if (q) cudaFreeHost(q);
        delete r;
    }
    g_res_free.clear();
    for (auto& kv : g_ipc_open) cudaIpcCloseMemHandle(kv.second);
    g_ipc_open.clear();
    for (auto& kv : g_host_pool)
        for (void* q : kv.second) cudaFreeHost(q);
    g_host_pool.clear();
    pool_trim();
    return 0;
}

int tk_local_modes(const tk_handle* h, int32_t* first, int32_t* count) {
    if (!h) return set_error(TK_EINVAL, "null handle");
    if (first) *first = h->first;
    if (count) *count = h->dl;
    return 0;
}

int tk_set_operator_csc(tk_handle* h, int32_t s, int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval) {
    bool local = false;
    TK_TRY(check_mode(h, s, &local));
    if (n != h->n) return set_error(TK_EINVAL, "operator order %lld != n = %d (system.jl:27-28)", (long long)n, h->n);
    if (!colptr || !rowval || !nzval) return set_error(TK_EINVAL, "null CSC array");
    if (colptr[0] != 1) return set_error(TK_EINVAL, "colptr must be 1-based (Julia SparseMatrixCSC)");
    if (s == 0) {        // every rank keeps the leading block of A_1: the exp-sum schedule is derived from its minors
        const int m = h->nmax;
        h->a1_lead.assign((size_t)m * m, 0.0);
        for (int64_t j = 0; j < m; ++j)
            for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) {
                const int64_t i = rowval[p] - 1;
                if (i >= 0 && i < m) h->a1_lead[(size_t)j * m + i] += nzval[p];
            }
    }
    int slots[2];
    const int nslots = slots_of(h, s, slots);
    if (nslots == 0) return 0;
    TK_CUDA(cudaSetDevice(h->device));
    const int64_t nnz = colptr[n] - 1;
    std::set<long long> offs;
    for (int64_t j = 0; j < n; ++j) {
        if (colptr[j + 1] < colptr[j]) return set_error(TK_EINVAL, "colptr not monotone");
        for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) {
            const int64_t i = rowval[p] - 1;
            if (i < 0 || i >= n) return set_error(TK_EINVAL, "rowval out of range");
            if (offs.size() <= (size_t)MAX_DIAG) offs.insert(j - i);
        }
    }
    std::unique_ptr<HostOp> op(new HostOp());
    op->ld = n; op->nnz = nnz;
    {
        std::vector<double> rows(n, 0.0);
        double n1 = 0.0, ninf = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            double cs = 0.0;
            for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) { cs += std::fabs(nzval[p]); rows[rowval[p] - 1] += std::fabs(nzval[p]); }
            n1 = std::max(n1, cs);
        }
        for (double r : rows) ninf = std::max(ninf, r);
        op->norm_bound = std::sqrt(n1 * ninf);
    }
    if (offs.size() <= (size_t)MAX_DIAG && !offs.empty()) {
        op->type = OP_DIA;
        op->ndiag = (int)offs.size();
        std::map<long long, int> idx;
        int c = 0;
        for (long long o : offs) { op->offs[c] = (int)o; idx[o] = c++; }
        std::vector<double> diag((size_t)op->ndiag * n, 0.0);
        for (int64_t j = 0; j < n; ++j)
            for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) {
                const int64_t i = rowval[p] - 1;
                diag[(size_t)idx[j - i] * n + i] += nzval[p];
            }
        op->constd = true;
        for (int j = 0; j < op->ndiag && op->constd; ++j) {
            const long long o = op->offs[j], i0 = std::max(0LL, -o), i1 = std::min<long long>(n, n - o);
            op->cval[j] = i0 < i1 ? diag[(size_t)j * n + i0] : 0.0;
            for (long long i = i0; i < i1; ++i)
                if (std::memcmp(&diag[(size_t)j * n + i], &op->cval[j], 8) != 0) { op->constd = false; break; }
        }
        TK_TRY(op->vals.alloc(diag.size(), false));
        TK_CUDA(cudaMemcpy(op->vals.p, diag.data(), 8 * diag.size(), cudaMemcpyHostToDevice));
    } else {
        op->type = OP_CSR;
        std::vector<int> rowptr(n + 1, 0);
        for (int64_t p = 0; p < nnz; ++p) rowptr[rowval[p]]++;   // rowval is 1-based: counts land at row+1
        for (int64_t i = 0; i < n; ++i) rowptr[i + 1] += rowptr[i];
        std::vector<int> fill(rowptr.begin(), rowptr.end() - 1), colidx(std::max<int64_t>(nnz, 1));
        std::vector<double> val(std::max<int64_t>(nnz, 1));
        for (int64_t j = 0; j < n; ++j)
            for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) {
                const int64_t i = rowval[p] - 1;
                const int q = fill[i]++;
                colidx[q] = (int)j;
                val[q] = nzval[p];
            }
        TK_TRY(op->vals.alloc(val.size(), false));
        TK_TRY(op->rowptr.alloc(rowptr.size(), false));
        TK_TRY(op->colidx.alloc(colidx.size(), false));
        TK_CUDA(cudaMemcpy(op->vals.p, val.data(), 8 * val.size(), cudaMemcpyHostToDevice));
        TK_CUDA(cudaMemcpy(op->rowptr.p, rowptr.data(), 4 * rowptr.size(), cudaMemcpyHostToDevice));
        TK_CUDA(cudaMemcpy(op->colidx.p, colidx.data(), 4 * colidx.size(), cudaMemcpyHostToDevice));
    }
    h->ops.push_back(std::move(op));
    for (int i = 0; i < nslots; ++i) h->mode_op[slots[i]] = (int)h->ops.size() - 1;
    h->ops_dirty = true;
    return 0;
}

int tk_set_operator_dense(tk_handle* h, int32_t s, int64_t n, const double* a, char uplo) {
    bool local = false;
    TK_TRY(check_mode(h, s, &local));
    if (n != h->n) return set_error(TK_EINVAL, "operator order %lld != n = %d", (long long)n, h->n);
    if (!a) return set_error(TK_EINVAL, "null matrix");
    if (uplo != 'L' && uplo != 'F') return set_error(TK_EINVAL, "uplo must be 'L' or 'F'");
    if (s == 0) {
        const int m = h->nmax;
        h->a1_lead.assign((size_t)m * m, 0.0);
        for (int j = 0; j < m; ++j)
            for (int i = 0; i < m; ++i)
                h->a1_lead[(size_t)j * m + i] = (uplo == 'L' && i < j) ? a[(size_t)i * n + j] : a[(size_t)j * n + i];
    }
    int slots[2];
    const int nslots = slots_of(h, s, slots);
    if (nslots == 0) return 0;
    TK_CUDA(cudaSetDevice(h->device));
    std::unique_ptr<HostOp> op(new HostOp());
    op->type = OP_DENSE; op->ld = n; op->nnz = n * n;
    {
        std::vector<double> rows(n, 0.0);
        double n1 = 0.0, ninf = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            double cs = 0.0;
            for (int64_t i = (uplo == 'L' ? j : 0); i < n; ++i) {
                const double v = std::fabs(a[(size_t)j * n + i]);
                cs += v; rows[i] += v;
                if (uplo == 'L' && i != j) { rows[j] += v; }
            }
            n1 = std::max(n1, cs);
        }
        for (double r : rows) ninf = std::max(ninf, r);
        op->norm_bound = uplo == 'L' ? ninf : std::sqrt(n1 * ninf);
    }
    TK_TRY(op->vals.alloc((size_t)n * n, false));
    if (uplo == 'F') {
        TK_CUDA(cudaMemcpy(op->vals.p, a, 8 * (size_t)n * n, cudaMemcpyHostToDevice));
    } else {
        std::vector<double> full((size_t)n * n);
        for (int64_t j = 0; j < n; ++j)
            for (int64_t i = j; i < n; ++i) {
                full[(size_t)j * n + i] = a[(size_t)j * n + i];
                full[(size_t)i * n + j] = a[(size_t)j * n + i];
            }
        TK_CUDA(cudaMemcpy(op->vals.p, full.data(), 8 * full.size(), cudaMemcpyHostToDevice));
    }
    h->ops.push_back(std::move(op));
    for (int i = 0; i < nslots; ++i) h->mode_op[slots[i]] = (int)h->ops.size() - 1;
    h->ops_dirty = true;
    return 0;
}

int tk_share_operator(tk_handle* h, int32_t s_dst, int32_t s_src) {
    bool ld = false, ls = false;
    TK_TRY(check_mode(h, s_dst, &ld));
    TK_TRY(check_mode(h, s_src, &ls));
    int dst[2], src[2];
    const int nd = slots_of(h, s_dst, dst), ns = slots_of(h, s_src, src);
    if (nd == 0) return 0;
    if (ns == 0) return set_error(TK_EINVAL, "mode %d is not held by this rank; set its operator here first", s_src);
    if (h->mode_op[src[0]] < 0) return set_error(TK_ESTATE, "operator of mode %d not set", s_src);
    for (int i = 0; i < nd; ++i) h->mode_op[dst[i]] = h->mode_op[src[0]];
    h->ops_dirty = true;
    return 0;
}

int tk_share_operator_all(tk_handle* h, int32_t s_src) {
    bool ls = false;
    TK_TRY(check_mode(h, s_src, &ls));
    int src[2];
    if (slots_of(h, s_src, src) == 0)
        return set_error(TK_EINVAL, "mode %d is not held by this rank; set its operator here first", s_src);
    if (h->mode_op[src[0]] < 0) return set_error(TK_ESTATE, "operator of mode %d not set", s_src);
    std::fill(h->mode_op.begin(), h->mode_op.begin() + h->dk, h->mode_op[src[0]]);
    h->ops_dirty = true;
    return 0;
}

int tk_needs_mode(const tk_handle* h, int32_t s, int32_t* needed) {
    bool local = false;
    TK_TRY(check_mode(h, s, &local));
    if (!needed) return set_error(TK_EINVAL, "null output");
    int slots[2];
    *needed = slots_of(h, s, slots) > 0 ? 1 : 0;
    return 0;
}

int tk_set_rhs(tk_handle* h, int32_t s, const double* b, int64_t n) {
    bool local = false;
    TK_TRY(check_mode(h, s, &local));
    if (n != h->n) return set_error(TK_EINVAL, "rhs length %lld != n = %d (system.jl:28)", (long long)n, h->n);
    if (!b) return set_error(TK_EINVAL, "null rhs");
    int slots[2];
    const int nslots = slots_of(h, s, slots);
    if (nslots == 0) return 0;
    TK_CUDA(cudaSetDevice(h->device));
    for (int i = 0; i < nslots; ++i) {
        TK_CUDA(cudaMemcpyAsync(h->b.p + (size_t)slots[i] * h->ldv, b, 8 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
        h->rhs_set[slots[i]] = 1;
    }
    TK_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int tk_set_rhs_all(tk_handle* h, const double* b, int64_t n) {
    if (!h || !b) return set_error(TK_EINVAL, "null argument");
    if (n != h->n) return set_error(TK_EINVAL, "rhs length %lld != n = %d", (long long)n, h->n);
    TK_CUDA(cudaSetDevice(h->device));
    if (h->dk > 0) {
        // one host->device copy, then replicate on the device (a 2D copy with source pitch 0 is not allowed, so
        // double the filled prefix until every slot is written)
        TK_CUDA(cudaMemcpyAsync(h->b.p, b, 8 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
        for (int filled = 1; filled < h->dk; filled *= 2) {
            const int cnt = std::min(filled, h->dk - filled);
            TK_CUDA(cudaMemcpyAsync(h->b.p + (size_t)filled * h->ldv, h->b.p, 8 * (size_t)cnt * h->ldv, cudaMemcpyDeviceToDevice, h->stream));
        }
    }
    TK_CUDA(cudaStreamSynchronize(h->stream));
    std::fill(h->rhs_set.begin(), h->rhs_set.end(), 1);
    return 0;
}

int tk_set_schedule(tk_handle* h, int32_t k, double lambda_min, int32_t t, const double* alpha, const double* omega) {
    if (!h || !alpha || !omega) return set_error(TK_EINVAL, "null argument");
    if (k < 2 || k > h->nmax) return set_error(TK_EINVAL, "k = %d outside 2..nmax", k);
    if (t < 1 || t > 4096) return set_error(TK_EINVAL, "bad term count t = %d", t);
    if (!(lambda_min > 0.0)) return set_error(TK_EINVAL, "lambda_min must be positive");
    SchedEntry& se = h->sched[k];
    if (!se.set || se.t != t) {
        se.off = h->alpha_pool.size();
        h->alpha_pool.resize(se.off + t);
        h->omega_pool.resize(se.off + t);
    }
    std::memcpy(h->alpha_pool.data() + se.off, alpha, 8 * (size_t)t);
    std::memcpy(h->omega_pool.data() + se.off, omega, 8 * (size_t)t);
    se.set = true; se.lambda_min = lambda_min; se.t = t;
    h->sched_dirty = true;
    return 0;
}

int tk_schedule_laplace(tk_handle* h, double tol) {
    if (!h) return set_error(TK_EINVAL, "null handle");
    for (int k = 2; k <= h->nmax; ++k) {
        double lmin, lmax;
        laplace_extremes(h->d, h->n, k, &lmin, &lmax);
        const double kappa = lmax * (1.0 / lmin);   // eigenvalues.jl:360
        int t, dg, od;
        const double *om, *al;
        TK_TRY(tables_sym_lookup(kappa, tol, &t, &dg, &od, &om, &al));
        TK_TRY(tk_set_schedule(h, k, lmin, t, al, om));
    }
    return 0;
}

int tk_schedule(tk_handle* h, double tol) {
    if (!h) return set_error(TK_EINVAL, "null handle");
    if (h->instance == TK_SYM && h->matrixclass == TK_LAPLACE) return tk_schedule_laplace(h, tol);
    if (h->a1_lead.empty())
        return set_error(TK_ESTATE, "operator of mode 0 not set on this rank: its leading minors define the spectral data (eigenvalues.jl:276-282); feed mode 0 to every rank");
    const int m = h->nmax;
    std::vector<double> ext(2 * (size_t)(m + 1), 0.0);
    if (h->instance == TK_NONSYM) {
        TK_TRY(minor_extremes(h->a1_lead, m, m, 1, ext));                   // eigenvalues.jl:344-350
    } else if (h->matrixclass == TK_RANDSPD) {
        TK_TRY(minor_extremes(h->a1_lead, m, m, 0, ext));                   // eigenvalues.jl:337
    } else if (h->matrixclass == TK_EIGVALMAT) {
        double lo = INFINITY, hi = -INFINITY;                               // eigenvalues.jl:339
        for (int k = 1; k <= m; ++k) {
            const double dg = h->a1_lead[(size_t)(k - 1) * m + (k - 1)];
            lo = std::min(lo, dg); hi = std::max(hi, dg);
            ext[2 * k] = lo; ext[2 * k + 1] = hi;
        }
    } else {
        return set_error(TK_EUNSUPPORTED, "the reference has no extreme_eigvals method for SymInstance with matrix class %d (eigenvalues.jl:335-350)", h->matrixclass);
    }
    std::vector<double> om, al;
    for (int k = 2; k <= m; ++k) {
        const double lmin = ext[2 * k] * (double)h->d;
        if (h->instance == TK_NONSYM) {
            int rank = 0;
            TK_TRY(nonsym_coefficients(lmin, tol, om, al, &rank));
            TK_TRY(tk_set_schedule(h, k, lmin, (int)om.size(), al.data(), om.data()));
        } else {
            const double lmax = ext[2 * k + 1] * (double)h->d;
            const double kappa = lmax * (1.0 / lmin);                       // eigenvalues.jl:360
            int t, dg, od;
            const double *o, *a;
            TK_TRY(tables_sym_lookup(kappa, tol, &t, &dg, &od, &o, &a));
            TK_TRY(tk_set_schedule(h, k, lmin, t, a, o));
        }
    }
    return 0;
}

int tk_begin(tk_handle* h) {
    if (!h) return set_error(TK_EINVAL, "null handle");
    TK_CUDA(cudaSetDevice(h->device));
    TK_TRY(begin_solve(h, 0.0));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int tk_step_bases(tk_handle* h, int32_t k) {
    if (!h || !h->begun) return set_error(TK_ESTATE, "call tk_begin first");
    if (k < 2 || k > h->nmax) return set_error(TK_EINVAL, "k = %d outside 2..nmax", k);
    TK_CUDA(cudaSetDevice(h->device));
    TK_TRY(enqueue_step_bases(h, k));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int tk_compress(tk_handle* h, int32_t k) {
    if (!h || !h->begun) return set_error(TK_ESTATE, "call tk_begin first");
    if (k < 2 || k > h->nmax) return set_error(TK_EINVAL, "k = %d outside 2..nmax", k);
    TK_CUDA(cudaSetDevice(h->device));
    TK_TRY(prepare(h, true));
    TK_TRY(enqueue_eig(h, k));
    TK_CUDA(cudaStreamSynchronize(h->stream3[k % tk_handle::NEIG]));
    TK_TRY(enqueue_assemble(h, k));
    TK_CUDA(cudaStreamSynchronize(h->stream2));
    return 0;
}

int tk_residual(tk_handle* h, int32_t k, double tol, double* out8) {
    if (!h || !h->begun || h->last_k != k) return set_error(TK_ESTATE, "call tk_compress(k) first");
    TK_CUDA(cudaSetDevice(h->device));
    TK_CUDA(cudaMemcpy(&h->ctl_d.p->tol, &tol, sizeof(double), cudaMemcpyHostToDevice));
    TK_TRY(enqueue_residual(h, k));
    TK_CUDA(cudaStreamSynchronize(h->stream2));
    if (out8) TK_CUDA(cudaMemcpy(out8, h->detail_d.p + (size_t)k * 8, 64, cudaMemcpyDeviceToHost));
    return 0;
}

int tk_solve(tk_handle* h, double tol, int32_t* status, int64_t* niter, int32_t* term_k, double* relres, double* projres,
             double* orth) {
    if (!h) return set_error(TK_EINVAL, "null handle");
    NvtxRange range("tk_solve");
    TK_CUDA(cudaSetDevice(h->device));
    TK_TRY(prepare(h, true));
    plan_segments(h);
    // Enqueue mode.  Kernel timing needs events between the launches, so it takes the direct path.  Otherwise the
    // segments are recorded as CUDA graphs the second time a handle solves with an unchanged configuration (a
    // one-shot solve would pay for the recording without using it twice); TK_GRAPH = 0 never, 2 from the first solve.
    // "Handle" includes a parked one that an identical tk_create revived: a caller that builds a solver per solve
    // (the reference's calling convention) replays graphs from its third solve on.
    const int gmode = env_int("TK_GRAPH", 1);
    const bool timed = (h->flags & (TK_FLAG_TIME_KERNELS | TK_FLAG_TIME_ALL)) != 0;
    const bool repeat = h->last_solve_sig == h->sig;
    const bool graph = !timed && (gmode >= 2 || (gmode == 1 && repeat));
    h->last_solve_sig = h->sig;
    arm_solve(h, tol);
    const bool fixed = (h->flags & TK_FLAG_FIXED_ITERATIONS) != 0;
    TK_CUDA(cudaEventRecord(h->ev_solve[0], h->stream));
    // The device decides: once finalize has set the status word every later kernel returns at once.  The host only
    // stops enqueueing: it looks at the pinned copy of the status two segments behind the launch front, so the GPU
    // queue never drains.
    const int nseg = (int)h->segs.size();
    for (int i = 0; i < nseg; ++i) {
        if (!fixed && i >= 2) {
            TK_CUDA(cudaEventSynchronize(h->res->seg_ev[(i - 2) & 3]));
            if (*reinterpret_cast<volatile int*>(&h->hctl->status) != ST_RUNNING) break;
        }
        TK_TRY(launch_segment(h, i, graph));
        TK_CUDA(cudaEventRecord(h->res->seg_ev[i & 3], h->stream));
    }
    TK_CUDA(cudaEventRecord(h->ev_solve[1], h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));      // every segment joins its side streams into this one
    h->begun = true;
    int st = ST_RUNNING, tk_ = 0, eigfail = 0;
    long long nit = 0;
    TK_CUDA(cudaMemcpy(&st, h->status_d.p, sizeof(int), cudaMemcpyDeviceToHost));
    TK_CUDA(cudaMemcpy(&tk_, h->term_k_d.p, sizeof(int), cudaMemcpyDeviceToHost));
    TK_CUDA(cudaMemcpy(&nit, h->niter_d.p, sizeof(long long), cudaMemcpyDeviceToHost));
    TK_CUDA(cudaMemcpy(&eigfail, h->eigfail_d.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (h->nmax < 2) { st = ST_NMAX; tk_ = 1; }
    if (relres) TK_CUDA(cudaMemcpy(relres, h->relres_d.p, 8 * (size_t)h->nmax, cudaMemcpyDeviceToHost));
    if (projres) TK_CUDA(cudaMemcpy(projres, h->projres_d.p, 8 * (size_t)h->nmax, cudaMemcpyDeviceToHost));
    if (orth) TK_CUDA(cudaMemcpy(orth, h->orth_d.p, 8 * (size_t)h->nmax, cudaMemcpyDeviceToHost));
    TK_TRY(collect_timing(h));
    {
        float ms = 0.f;
        TK_CUDA(cudaEventElapsedTime(&ms, h->ev_solve[0], h->ev_solve[1]));
        h->tm_ms[TM_SOLVE] = ms; h->tm_launches[TM_SOLVE] = 1;
    }
    // The compressed solution resident in Y is the one of the iteration the loop left at: the assembly kernels of
    // the iterations enqueued behind it returned at once.
    if (tk_ >= 2) {
        const SchedEntry& se = h->sched[tk_];
        h->last_k = tk_; h->last_t = se.t; h->last_tld = (se.t + 3) & ~3; h->last_lam_inv = 1.0 / se.lambda_min;
    } else {
        h->last_k = 0; h->last_t = 0; h->last_tld = 0;
    }
    if (status) *status = st;
    if (niter) *niter = nit;
    if (term_k) *term_k = tk_;
    if (eigfail) return set_error(TK_ESTATE, "tridiagonal eigensolver did not converge");
    if (st == ST_RUNNING) return set_error(TK_ESTATE, "solve left the loop while still running");
    return 0;
}

int tk_solution_rank(tk_handle* h, int32_t* t) {
    if (!h || !t) return set_error(TK_EINVAL, "null argument");
    *t = h->last_t;       // rank of the iterate the getters below return (after tk_solve: the iteration it left at)
    return 0;
}

}  // extern "C"

namespace tk {

// which iterate the solution getters return, and whether they may
static int solution_state(tk_handle* h, int32_t force, int* k, int* t, int* tld) {
    if (h->last_k < 2) return set_error(TK_ESTATE, "no compressed solution available");
    int st = ST_RUNNING;
    TK_CUDA(cudaMemcpy(&st, h->status_d.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (st != ST_CONVERGED && !force)
        return set_error(TK_ESTATE, "solve did not converge (status %d); pass force to read the last iterate", st);
    *k = h->last_k; *t = h->last_t; *tld = h->last_tld;
    return 0;
}

static void solution_lambda(const tk_handle* h, int k, int t, double* lambda) {
    const SchedEntry& se = h->sched[k];
    const double lam_inv = 1.0 / se.lambda_min;
    for (int j = 0; j < t; ++j) lambda[j] = lam_inv * h->omega_pool[se.off + j];   // y.lambda, tensor_krylov_method.jl:23
}

// basis_tensor_mul! for local modes [m0, m0 + nm) into X ([mode][t][n]) on `st`
static int launch_basis_mul(tk_handle* h, int k, int t, int tld, int m0, int nm, double* X, cudaStream_t st) {
    if (nm <= 0) return 0;
    const int TJ = t <= 8 ? 8 : t <= 16 ? 16 : 32;
    dim3 grid((h->n + 255) / 256, nm, (t + TJ - 1) / TJ);
    const size_t smem = (size_t)k * TJ * 8;
#define TK_BM_LAUNCH(T_)                                                                                         \
    do {                                                                                                         \
        TK_TRY(allow_smem(basis_mul_all_kernel<T_>, smem));                                                      \
        basis_mul_all_kernel<T_><<<grid, 256, smem, st>>>(h->V.p, (long long)h->ncol * h->ldv, h->ldv, h->n, k,  \
                                                          h->Y.p, h->ystride, tld, t, X, m0);                    \
    } while (0)
    if (TJ == 8) TK_BM_LAUNCH(8); else if (TJ == 16) TK_BM_LAUNCH(16); else TK_BM_LAUNCH(32);
#undef TK_BM_LAUNCH
    TK_CUDA(cudaGetLastError());
    return 0;
}

static bool host_pointer_is_pinned(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}

// Factor matrices of local modes [m0, m0 + nm) to host memory: the modes are cut into chunks, chunk c+1 is computed
// while chunk c crosses PCIe.  Pinned destinations (tk_alloc_host) are written by DMA directly; pageable ones go
// through two pinned staging buffers and a host copy.
static int solution_to_host(tk_handle* h, int k, int t, int tld, int m0, int nm, double* fmat) {
    if (nm <= 0) return 0;
    NvtxRange range("tk solution to host");
    tk_resources* r = h->res;
    const size_t per_mode = (size_t)h->n * t;                      // doubles
    const size_t chunk_bytes_target = (size_t)std::max(1, env_int("TK_SOL_CHUNK_MB", 32)) << 20;
    const int cm = (int)std::max<size_t>(1, std::min<size_t>((size_t)nm, chunk_bytes_target / (per_mode * 8)));
    const bool pinned = host_pointer_is_pinned(fmat);
    DevBuf<double> X[2];
    TK_TRY(X[0].alloc((size_t)cm * per_mode, false));
    TK_TRY(X[1].alloc((size_t)cm * per_mode, false));
    if (!pinned && r->stage_bytes < (size_t)cm * per_mode * 8) {
        for (auto& q : r->stage) { if (q) cudaFreeHost(q); q = nullptr; }
        r->stage_bytes = 0;
        for (auto& q : r->stage) TK_CUDA(cudaMallocHost(reinterpret_cast<void**>(&q), (size_t)cm * per_mode * 8));
        r->stage_bytes = (size_t)cm * per_mode * 8;
    }
    const int nch = (nm + cm - 1) / cm;
    auto drain = [&](int c) -> int {           // pageable destination: staging buffer of chunk c -> caller's memory
        const int c0 = c * cm, cn = std::min(cm, nm - c0);
        TK_CUDA(cudaEventSynchronize(r->copy_ev[c & 1]));
        // a fresh destination is first touched here (page faults): several host threads share the copy
        const size_t bytes = (size_t)cn * per_mode * 8;
        char* dst = reinterpret_cast<char*>(fmat + (size_t)c0 * per_mode);
        const char* src = reinterpret_cast<const char*>(r->stage[c & 1]);
        const int nthr = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(8, std::thread::hardware_concurrency()), bytes >> 20));
        std::vector<std::thread> pool;
        const size_t piece = ((bytes / nthr) + 4095) & ~(size_t)4095;
        for (int i = 1; i < nthr; ++i) {
            const size_t o = (size_t)i * piece;
            if (o < bytes) pool.emplace_back([=]() { std::memcpy(dst + o, src + o, std::min(piece, bytes - o)); });
        }
        std::memcpy(dst, src, std::min(piece, bytes));
        for (auto& th : pool) th.join();
        return 0;
    };
    for (int c = 0; c < nch; ++c) {
        const int c0 = c * cm, cn = std::min(cm, nm - c0), bi = c & 1;
        if (c >= 2) {
            TK_CUDA(cudaStreamWaitEvent(h->stream, r->copy_ev[bi], 0));     // X[bi] has left the device
            if (!pinned) TK_TRY(drain(c - 2));                              // ... and its staging buffer is free again
        }
        TK_TRY(launch_basis_mul(h, k, t, tld, m0 + c0, cn, X[bi].p, h->stream));
        TK_CUDA(cudaEventRecord(r->fill_ev[bi], h->stream));
        TK_CUDA(cudaStreamWaitEvent(r->s_copy, r->fill_ev[bi], 0));
        double* dst = pinned ? fmat + (size_t)c0 * per_mode : r->stage[bi];
        TK_CUDA(cudaMemcpyAsync(dst, X[bi].p, (size_t)cn * per_mode * 8, cudaMemcpyDeviceToHost, r->s_copy));
        TK_CUDA(cudaEventRecord(r->copy_ev[bi], r->s_copy));
    }
    if (!pinned)
        for (int c = std::max(0, nch - 2); c < nch; ++c) TK_TRY(drain(c));
    TK_CUDA(cudaStreamSynchronize(r->s_copy));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

}  // namespace tk

extern "C" {

int tk_get_solution(tk_handle* h, int32_t s, double* lambda, int32_t lambda_cap, double* fmat, int64_t fmat_cap, int32_t force) {
    bool local = false;
    TK_TRY(check_mode(h, s, &local));
    TK_CUDA(cudaSetDevice(h->device));
    int k, t, tld;
    TK_TRY(solution_state(h, force, &k, &t, &tld));
    if (lambda) {
        if (lambda_cap < t) return set_error(TK_EINVAL, "lambda holds %d entries, the solution has rank %d (tk_solution_rank)", lambda_cap, t);
        solution_lambda(h, k, t, lambda);
    }
    if (!local || !fmat) return 0;
    if (fmat_cap < (int64_t)h->n * t) return set_error(TK_EINVAL, "fmat holds %lld doubles, need n*t = %lld", (long long)fmat_cap, (long long)h->n * t);
    return solution_to_host(h, k, t, tld, s - h->first, 1, fmat);
}

int tk_get_solution_all(tk_handle* h, double* lambda, int32_t lambda_cap, double* fmat, int64_t fmat_cap, int32_t force) {
    if (!h) return set_error(TK_EINVAL, "null handle");
    TK_CUDA(cudaSetDevice(h->device));
    int k, t, tld;
    TK_TRY(solution_state(h, force, &k, &t, &tld));
    if (lambda) {
        if (lambda_cap < t) return set_error(TK_EINVAL, "lambda holds %d entries, the solution has rank %d (tk_solution_rank)", lambda_cap, t);
        solution_lambda(h, k, t, lambda);
    }
    if (!fmat || h->dl == 0) return 0;
    if (fmat_cap < (int64_t)h->dl * h->n * t)
        return set_error(TK_EINVAL, "fmat holds %lld doubles, need count*n*t = %lld", (long long)fmat_cap, (long long)h->dl * h->n * t);
    return solution_to_host(h, k, t, tld, 0, h->dl, fmat);
}

int tk_get_solution_device(tk_handle* h, double* lambda, int32_t lambda_cap, double* fmat_dev, int64_t fmat_cap, int32_t force) {
    if (!h) return set_error(TK_EINVAL, "null handle");
    TK_CUDA(cudaSetDevice(h->device));
    int k, t, tld;
    TK_TRY(solution_state(h, force, &k, &t, &tld));
    if (lambda) {
        if (lambda_cap < t) return set_error(TK_EINVAL, "lambda holds %d entries, the solution has rank %d (tk_solution_rank)", lambda_cap, t);
        solution_lambda(h, k, t, lambda);
    }
    if (!fmat_dev || h->dl == 0) return 0;
    if (fmat_cap < (int64_t)h->dl * h->n * t)
        return set_error(TK_EINVAL, "fmat_dev holds %lld doubles, need count*n*t = %lld", (long long)fmat_cap, (long long)h->dl * h->n * t);
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, fmat_dev) != cudaSuccess || attr.type != cudaMemoryTypeDevice || attr.device != h->device) {
        cudaGetLastError();
        return set_error(TK_EINVAL, "fmat_dev is not device memory of CUDA device %d", h->device);
    }
    TK_TRY(launch_basis_mul(h, k, t, tld, 0, h->dl, fmat_dev, h->stream));
    TK_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// Page-locking memory costs ~0.3 ms per MB, far more than moving it: freed blocks are parked (by size) for the next
// request of the same size; tk_release_cache returns them to the system.
int tk_alloc_host(void** out, int64_t bytes) {
    if (!out || bytes < 0) return set_error(TK_EINVAL, "bad arguments");
    const size_t sz = (size_t)std::max<int64_t>(bytes, 1);
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = g_host_pool.find(sz);
        if (it != g_host_pool.end() && !it->second.empty()) {
            *out = it->second.back();
            it->second.pop_back();
            g_host_live[*out] = sz;
            return 0;
        }
    }
    TK_CUDA(cudaMallocHost(out, sz));
    std::lock_guard<std::mutex> lock(g_mutex);
    g_host_live[*out] = sz;
    return 0;
}

int tk_free_host(void* p) {
    if (!p) return 0;
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_host_live.find(p);
    if (it == g_host_live.end()) return set_error(TK_EINVAL, "pointer was not returned by tk_alloc_host");
    g_host_pool[it->second].push_back(p);
    g_host_live.erase(it);
    return 0;
}

int tk_get_detail(tk_handle* h, int32_t k0, int32_t k1, double* out) {
    if (!h || !out) return set_error(TK_EINVAL, "null argument");
    if (k0 < 2 || k1 > h->nmax || k1 < k0) return set_error(TK_EINVAL, "iterations %d..%d outside 2..nmax", k0, k1);
    TK_CUDA(cudaSetDevice(h->device));
    TK_CUDA(cudaMemcpy(out, h->detail_d.p + (size_t)k0 * 8, 64 * (size_t)(k1 - k0 + 1), cudaMemcpyDeviceToHost));
    return 0;
}

int tk_get_solve_info(tk_handle* h, int32_t* graphs_launched, double* graph_build_ms, int32_t* peer_exchange, int32_t* segments) {
    if (!h) return set_error(TK_EINVAL, "null handle");
    if (graphs_launched) *graphs_launched = h->graphs_launched;
    if (graph_build_ms) *graph_build_ms = h->graph_build_ms;
    if (peer_exchange) *peer_exchange = h->px_ready ? 1 : 0;
    if (segments) *segments = (int)h->segs.size();
    return 0;
}

int tk_get_H(tk_handle* h, int32_t s, double* H) {
    bool local = false;
    TK_TRY(check_mode(h, s, &local));
    if (!H) return set_error(TK_EINVAL, "null output");
    if (!local) return set_error(TK_EINVAL, "mode %d is owned by another rank", s);
    TK_CUDA(cudaSetDevice(h->device));
    const int nc = h->ncol, sl = s - h->first;
    if (h->Hd.p) {
        TK_CUDA(cudaMemcpy(H, h->Hd.p + (size_t)sl * nc * nc, 8 * (size_t)nc * nc, cudaMemcpyDeviceToHost));
        return 0;
    }
    std::vector<double> T(3 * (size_t)nc);
    TK_CUDA(cudaMemcpy(T.data(), h->T.p + (size_t)sl * 3 * nc, 8 * T.size(), cudaMemcpyDeviceToHost));
    std::fill(H, H + (size_t)nc * nc, 0.0);
    for (int j = 0; j < nc; ++j) {
        H[(size_t)j * nc + j] = T[j];
        if (j + 1 < nc) {
            H[(size_t)j * nc + (j + 1)] = T[nc + j];         // H[j+2, j+1] (1-based): sub-diagonal
            H[(size_t)(j + 1) * nc + j] = T[2 * nc + j];     // H[j+1, j+2]: super-diagonal
        }
    }
    return 0;
}

int tk_get_V(tk_handle* h, int32_t s, int32_t col, double* v) {
    bool local = false;
    TK_TRY(check_mode(h, s, &local));
    if (!v || col < 1 || col > h->ncol) return set_error(TK_EINVAL, "bad column %d", col);
    if (!local) return set_error(TK_EINVAL, "mode %d is owned by another rank", s);
    TK_CUDA(cudaSetDevice(h->device));
    TK_CUDA(cudaMemcpy(v, h->V.p + ((size_t)(s - h->first) * h->ncol + (col - 1)) * h->ldv, 8 * (size_t)h->n, cudaMemcpyDeviceToHost));
    return 0;
}

int tk_get_bt(tk_handle* h, int32_t s, double* bt) {
    bool local = false;
    TK_TRY(check_mode(h, s, &local));
    if (!bt) return set_error(TK_EINVAL, "null output");
    if (!local) return set_error(TK_EINVAL, "mode %d is owned by another rank", s);
    TK_CUDA(cudaSetDevice(h->device));
    TK_CUDA(cudaMemcpy(bt, h->bt.p + (size_t)(s - h->first) * h->ncol, 8 * (size_t)h->ncol, cudaMemcpyDeviceToHost));
    return 0;
}

int tk_get_Y(tk_handle* h, int32_t s, int32_t k, double* Y, int32_t* t) {
    bool local = false;
    TK_TRY(check_mode(h, s, &local));
    if (h->last_k != k) return set_error(TK_ESTATE, "Y of iteration %d is not resident (last compress was k=%d)", k, h->last_k);
    if (!local) return set_error(TK_EINVAL, "mode %d is owned by another rank", s);
    TK_CUDA(cudaSetDevice(h->device));
    const int tt = h->last_t, tld = h->last_tld;
    if (t) *t = tt;
    if (!Y) return 0;
    std::vector<double> rm((size_t)k * tld);
    TK_CUDA(cudaMemcpy(rm.data(), h->Y.p + (size_t)(s - h->first) * h->ystride, 8 * rm.size(), cudaMemcpyDeviceToHost));
    for (int j = 0; j < tt; ++j)
        for (int r = 0; r < k; ++r) Y[(size_t)j * k + r] = rm[(size_t)r * tld + j];
    return 0;
}

int tk_get_eig(tk_handle* h, int32_t s, int32_t k, double* theta, double* Q) {
    bool local = false;
    TK_TRY(check_mode(h, s, &local));
    if (h->last_k != k) return set_error(TK_ESTATE, "eigendecomposition of iteration %d is not resident", k);
    if (!local) return set_error(TK_EINVAL, "mode %d is owned by another rank", s);
    TK_CUDA(cudaSetDevice(h->device));
    const int cls = h->per_mode ? s - h->first : 0;
    const size_t par = h->use_expm ? 0 : (size_t)(k % h->ring_depth);
    if (theta) TK_CUDA(cudaMemcpy(theta, h->theta.p + (par * h->ncls + cls) * h->ncol, 8 * (size_t)k, cudaMemcpyDeviceToHost));
    if (Q) {
        std::vector<double> q((size_t)h->ldq * k);
        TK_CUDA(cudaMemcpy(q.data(), h->Q.p + (par * h->ncls + cls) * h->ldq * h->ldq, 8 * q.size(), cudaMemcpyDeviceToHost));
        for (int i = 0; i < k; ++i)
            for (int r = 0; r < k; ++r) Q[(size_t)i * k + r] = q[(size_t)i * h->ldq + r];
    }
    return 0;
}

int tk_get_orth_state(tk_handle* h, int32_t s, double* S, int32_t* fallbacks) {
    bool local = false;
    TK_TRY(check_mode(h, s, &local));
    if (!local) return set_error(TK_EINVAL, "mode %d is owned by another rank", s);
    TK_CUDA(cudaSetDevice(h->device));
    if (S) TK_CUDA(cudaMemcpy(S, h->S.p + (s - h->first), 8, cudaMemcpyDeviceToHost));
    if (fallbacks) TK_CUDA(cudaMemcpy(fallbacks, h->fallbacks.p + (s - h->first), 4, cudaMemcpyDeviceToHost));
    return 0;
}

int tk_tridiag_eig_batched(int32_t device, int32_t nb, int32_t k, const double* diag, const double* sub, double* theta, double* Q,
                           int32_t* fallbacks) {
    if (nb < 1 || k < 1 || !diag || !theta || (k > 1 && !sub)) return set_error(TK_EINVAL, "bad arguments");
    TK_CUDA(cudaSetDevice(device));
    const int ncol = k;
    std::vector<double> T((size_t)nb * 2 * ncol, 0.0);
    for (int p = 0; p < nb; ++p) {
        std::memcpy(&T[(size_t)p * 2 * ncol], diag + (size_t)p * k, 8 * (size_t)k);
        if (k > 1) std::memcpy(&T[(size_t)p * 2 * ncol + ncol], sub + (size_t)p * (k - 1), 8 * (size_t)(k - 1));
    }
    DevBuf<double> Td, thd, Qd, Sd;
    DevBuf<int> fail, need;
    TK_TRY(Td.alloc(T.size(), false));
    TK_TRY(thd.alloc((size_t)nb * k));
    TK_TRY(Qd.alloc((size_t)nb * k * k));
    TK_TRY(Sd.alloc((size_t)nb * k * k, false));
    TK_TRY(fail.alloc(1));
    TK_TRY(need.alloc(nb));
    TK_CUDA(cudaMemcpy(Td.p, T.data(), 8 * T.size(), cudaMemcpyHostToDevice));
    TK_TRY(launch_eig(Td.p, 2LL * ncol, ncol, k, nb, thd.p, k, Qd.p, (long long)k * k, k, nullptr, fail.p, Sd.p, need.p, 0, nullptr));
    if (fallbacks) {
        std::vector<int> nd(nb);
        TK_CUDA(cudaMemcpy(nd.data(), need.p, 4 * (size_t)nb, cudaMemcpyDeviceToHost));
        int c = 0;
        for (int v : nd) c += v;
        *fallbacks = c;
    }
    TK_CUDA(cudaDeviceSynchronize());
    int f = 0;
    TK_CUDA(cudaMemcpy(&f, fail.p, 4, cudaMemcpyDeviceToHost));
    TK_CUDA(cudaMemcpy(theta, thd.p, 8 * (size_t)nb * k, cudaMemcpyDeviceToHost));
    if (Q) TK_CUDA(cudaMemcpy(Q, Qd.p, 8 * (size_t)nb * k * k, cudaMemcpyDeviceToHost));
    if (f) return set_error(TK_ESTATE, "tridiagonal eigensolver did not converge");
    return 0;
}

int tk_timing_mark(tk_handle* h) {
    if (!h) return set_error(TK_EINVAL, "null handle");
    TK_CUDA(cudaSetDevice(h->device));
    TK_CUDA(cudaEventRecord(h->ev_region, h->stream));
    return 0;
}

int tk_get_timing(tk_handle* h, int32_t which, double* ms_total, int64_t* launches, double* algorithmic_bytes) {
    if (!h || which < 0 || which >= TM_KINDS) return set_error(TK_EINVAL, "bad timing kind %d", which);
    if (which == TM_REGION) {
        if (!h->ev_region || !h->ev_solve[1]) return set_error(TK_ESTATE, "no timed region (tk_timing_mark, then tk_solve)");
        float ms = 0.f;
        TK_CUDA(cudaEventElapsedTime(&ms, h->ev_region, h->ev_solve[1]));
        if (ms_total) *ms_total = ms;
        if (launches) *launches = 0;
        if (algorithmic_bytes) *algorithmic_bytes = 0.0;
        return 0;
    }
    if (ms_total) *ms_total = h->tm_ms[which];
    if (launches) *launches = h->tm_launches[which];
    if (algorithmic_bytes) *algorithmic_bytes = h->tm_bytes[which];
    return 0;
}

int tk_launch_count(tk_handle* h, int64_t* launches) {
    if (!h || !launches) return set_error(TK_EINVAL, "null argument");
    *launches = h->launches;
    return 0;
}

}  // extern "C"
