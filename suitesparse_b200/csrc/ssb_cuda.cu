// ssb_cuda.cu — device plan, launch orchestration and the C ABI (plain layer + CHOLMOD drop-in layer) declared in
// include/suitesparse_b200.h.  No CPU fallback: every numeric step below is a kernel from ssb_kernels.cuh.
#include "ssb_kernels.cuh"
#include "../../include/suitesparse_b200.h"
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <link.h>
#include <sys/mman.h>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <condition_variable>
#include <deque>
#include <atomic>
#include <chrono>
#include <string>
#include <vector>
#include <algorithm>

using namespace ssb;

static thread_local std::string g_last_error;
static void set_error(const std::string &s) { g_last_error = s; }

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + std::to_string(__LINE__)); \
            return SSB_CHOLMOD_GPU_PROBLEM;                                                            \
        }                                                                                              \
    } while (0)

// device-side job arrays of one schedule (the plan's own, or a temporary one for the not-posdef repeat)
struct DevJobs {
    GemmJob *gemm_jobs = nullptr; int *gemm_tiles = nullptr;
    PanelJob *potrf_jobs = nullptr, *trsm_jobs = nullptr; int *trsm_tiles = nullptr;
};

// Device-to-host streaming runs on a helper thread: cudaMemcpyAsync into cudaHostRegister'ed (4 KiB-paged) memory costs
// the CALLING thread ~100 us per 50 MB (DMA descriptors), which would starve the latency-bound potrf/trsm launch chain.
//
// Two kinds of destination.  Page-locked L->x: the copy goes straight there.  Pageable L->x (the first factorization into a
// fresh factor: page-locking 29 GB of untouched memory costs 9.4 s, scripts/pin_probe.cu): the copy is STAGED - the DMA
// lands in a small ring of driver-allocated pinned slots and a pool of host threads moves each slot into L->x while the
// next slots are in flight (16 threads move 31 GB into untouched pageable memory in 1.07 s, 0.25 s once the pages exist).
struct CopyJob { cudaEvent_t gate; const double *src; double *dst; size_t bytes; bool staged; };

// K threads that split one memcpy; the caller blocks until it is done
struct HostMover {
    std::vector<std::thread> th; std::mutex mu; std::condition_variable cv, done_cv;
    const char *src = nullptr; char *dst = nullptr; size_t bytes = 0; unsigned long long gen = 0; int remaining = 0, K = 0; bool stop = false;
    void start(int k) { K = k; if (K > 1) for (int t = 0; t < K; t++) th.emplace_back([this, t] { run(t); }); }
    void run(int t)
    {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen; const char *s = src; char *d = dst; const size_t n = bytes;
            lk.unlock();
            const size_t per = ((n / K) + 4096) & ~(size_t) 4095, lo = std::min(n, per * t), hi = std::min(n, lo + per);
            if (hi > lo) memcpy(d + lo, s + lo, hi - lo);
            lk.lock();
            if (--remaining == 0) done_cv.notify_one();
        }
    }
    void copy(char *d, const char *s, size_t n)
    {
        if (n < ((size_t) 1 << 20) || K <= 1) { memcpy(d, s, n); return; }
        std::unique_lock<std::mutex> lk(mu);
        src = s; dst = d; bytes = n; remaining = K; gen++;
        cv.notify_all();
        done_cv.wait(lk, [&] { return remaining == 0; });
    }
    void shutdown() { { std::lock_guard<std::mutex> lk(mu); stop = true; } cv.notify_all(); for (auto &t : th) t.join(); th.clear(); }
};

struct Copier {
    std::thread th; std::mutex mu; std::condition_variable cv; std::deque<CopyJob> q;
    bool stop = false; std::atomic<int> pending{0}; std::atomic<int> failed{0};
    int device = 0; cudaStream_t stream = nullptr;
    // staging ring (allocated on first use)
    char *ring = nullptr; size_t slot_bytes = 0; int nslots = 0; std::vector<cudaEvent_t> slot_ev; HostMover mover;
    struct Piece { int slot; char *dst; size_t n; bool last; };
    void start(int dev, cudaStream_t s) { device = dev; stream = s; th = std::thread([this] { run(); }); }
    int ensure_ring(size_t total_bytes)                    // called by the factorizing thread before staged jobs are pushed
    {
        if (ring) return 0;
        size_t mb = 64; int ns = 4, k = (int) std::min(8u, std::max(2u, std::thread::hardware_concurrency() / 2));
        if (const char *v = getenv("SSB200_STAGE_SLOT_MB")) mb = (size_t) std::max(1, atoi(v));
        if (const char *v = getenv("SSB200_STAGE_SLOTS")) ns = std::max(2, atoi(v));
        if (const char *v = getenv("SSB200_STAGE_THREADS")) k = std::max(1, atoi(v));
        mb = std::max<size_t>(1, std::min(mb, (total_bytes / ns >> 20) + 1));      // a small factor gets a small ring
        if (total_bytes < ((size_t) 64 << 20)) k = 1;
        if (cudaHostAlloc((void **) &ring, mb * ns << 20, cudaHostAllocDefault) != cudaSuccess) { (void) cudaGetLastError(); ring = nullptr; return 1; }
        slot_bytes = mb << 20; nslots = ns; slot_ev.resize(ns);
        for (auto &e : slot_ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return 1;
        mover.start(k);
        return 0;
    }
    void run()
    {
        cudaSetDevice(device);
        std::deque<Piece> outstanding;                     // staged pieces whose DMA was issued, oldest first
        CopyJob cur{}; size_t cur_off = 0; bool have = false; int next_slot = 0;
        for (;;) {
            if (!have) {
                std::unique_lock<std::mutex> lk(mu);
                if (outstanding.empty()) cv.wait(lk, [this] { return stop || !q.empty(); });
                if (!q.empty()) { cur = q.front(); q.pop_front(); have = true; cur_off = 0; }
                else if (outstanding.empty()) return;      // stop
            }
            if (have && !cur.staged) {
                if (cur.gate && cudaStreamWaitEvent(stream, cur.gate, 0) != cudaSuccess) failed++;
                if (cudaMemcpyAsync(cur.dst, cur.src, cur.bytes, cudaMemcpyDeviceToHost, stream) != cudaSuccess) failed++;
                pending--; have = false;
                continue;
            }
            if (have && (int) outstanding.size() < nslots) {
                // the slot is free: pieces complete in the order they were issued
                const int slot = next_slot; next_slot = (next_slot + 1) % nslots;
                const size_t n = std::min(slot_bytes, cur.bytes - cur_off);
                if (cur_off == 0 && cur.gate && cudaStreamWaitEvent(stream, cur.gate, 0) != cudaSuccess) failed++;
                if (cudaMemcpyAsync(ring + (size_t) slot * slot_bytes, (const char *) cur.src + cur_off, n, cudaMemcpyDeviceToHost, stream) != cudaSuccess) failed++;
                if (cudaEventRecord(slot_ev[slot], stream) != cudaSuccess) failed++;
                outstanding.push_back(Piece{slot, (char *) cur.dst + cur_off, n, cur_off + n == cur.bytes});
                cur_off += n;
                if (cur_off == cur.bytes) have = false;
                continue;
            }
            if (!outstanding.empty()) {
                const Piece pc = outstanding.front(); outstanding.pop_front();
                if (cudaEventSynchronize(slot_ev[pc.slot]) != cudaSuccess) failed++;
                else mover.copy(pc.dst, ring + (size_t) pc.slot * slot_bytes, pc.n);
                if (pc.last) pending--;
            }
        }
    }
    void push(const CopyJob &j) { pending++; { std::lock_guard<std::mutex> lk(mu); q.push_back(j); } cv.notify_one(); }
    // direct jobs: every copy has been enqueued on the copy stream; staged jobs: every byte is in its destination
    void drain() { while (pending.load() > 0) std::this_thread::yield(); }
    void shutdown()
    {
        if (th.joinable()) { { std::lock_guard<std::mutex> lk(mu); stop = true; } cv.notify_one(); th.join(); }
        if (ring) { mover.shutdown(); cudaSetDevice(device); for (auto &e : slot_ev) cudaEventDestroy(e); cudaFreeHost(ring); ring = nullptr; }
    }
};

// parameters baked into the captured factorization graph
struct FactorGraphKey { double beta0; double *host; int stype; int two_streams; const void *a[4]; const void *f[4]; };

struct CscBuf { long long *p = nullptr, *i = nullptr, *nz = nullptr; double *x = nullptr; size_t capP = 0, capI = 0, capNz = 0, capX = 0; bool haveNz = false; };
static void free_cscbuf(CscBuf *b) { if (!b) return; if (b->p) cudaFree(b->p); if (b->i) cudaFree(b->i); if (b->nz) cudaFree(b->nz); if (b->x) cudaFree(b->x); delete b; }

struct ssb200_plan {
    HostPlan hp;
    int device = 0;
    cudaStream_t stream = nullptr;           // launches go here (the plan's own stream, or the caller's: ssb200_set_stream)
    cudaStream_t own_stream = nullptr;
    int *d_owner = nullptr;                  // sharded plans: owner rank per supernode (-1 = panel-cyclic)
    cudaStream_t panel_stream = nullptr;     // look-ahead: the potrf/trsm/small-K chain of the next outer panel (high priority)
    std::vector<cudaEvent_t> la_events;      // cross-stream events of the look-ahead schedule (HostPlan::n_events)
    int panel_prio = 0;
    bool capturing = false;                  // a factorization graph is being captured
    std::vector<cudaGraphNode_t> cap_high_nodes;   // captured kernel nodes of the panel stream (their priority is set on the graph)
    bool lookahead = true;                   // run the schedule on two streams; false: one stream, per-launch timing events
    cudaStream_t copy_stream = nullptr;      // device-to-host streaming of finished supernodes
    cudaEvent_t copy_gate = nullptr, copy_done = nullptr;
    std::vector<cudaEvent_t> copy_gates;     // one per copy-task group
    cudaGraphExec_t fgraph = nullptr; FactorGraphKey fg_key{}; size_t fg_ev = 0; ssb_long fg_launches = 0;
    std::vector<std::pair<int, size_t>> fg_marks;   // (launch index, event index) of the last enqueue / capture
    Copier *copier = nullptr;
    int *d_super = nullptr, *d_ls = nullptr, *d_supermap = nullptr, *d_relmap = nullptr, *d_info = nullptr;
    long long *d_pi = nullptr, *d_px = nullptr;
    double *d_Lx = nullptr;
    long long lx_alloc = 0;                  // doubles behind d_Lx: xsize, or the local share under distributed storage
    double *d_winv = nullptr; long long winv_slots = 0;   // inverses of the wide 64x64 diagonal blocks (trsm_tc + solves)
    double *d_probe = nullptr;               // 8-byte scratch of the host-registration probe
    DevJobs jobs;
    SolveJob *d_solve_jobs = nullptr; int *d_solve_tiles = nullptr;
    // second solve schedule (256-column blocks of the wide supernodes; needs the inverses of the diagonal blocks)
    SolveJob *d_solve2_jobs = nullptr; int *d_solve2_tiles = nullptr; SolveBlk *d_solve_blks = nullptr; int *d_solve_blk_ctas = nullptr;
    double *d_sol_scratch = nullptr; int *d_sol_flags = nullptr; long long sol_nrhs_cap = 0;
    int *h_info = nullptr;                 // pinned
    // matrix on the device
    struct CscBuf *bufA = nullptr, *bufF = nullptr;
    bool haveA = false;
    int stype = -1;
    double *d_X = nullptr; size_t capX = 0;
    bool factor_on_device = false;
    double last_beta0 = 0.0;               // beta of the running sharded factorization (not-positive-definite repeat)
    bool winv_valid = false;               // d_winv matches d_Lx (false after ssb200_upload_L or a sharded factorization)
    // the whole solve sequence is replayed as one CUDA graph (thousands of tiny dependent kernels)
    // one cached graph per direction (L, L', both): cholmod_solve2 calls lsolve and ltsolve alternately, each with its own graph
    struct SolveGraph { cudaGraphExec_t exec = nullptr; double *X = nullptr; long long nrhs = 0, ldx = 0, launches = 0; bool winv = false, blk = false; } sg[3];
    std::vector<cudaEvent_t> events;
    ssb200_stats stats{};
    std::vector<float> launch_ms;          // device time of every launch of the last factorize (debug / tuning)
    size_t device_bytes = 0;
};

template <typename T> static int dev_alloc_copy(ssb200_plan *p, T **dst, const std::vector<T> &src)
{
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    CU_TRY(cudaMalloc((void **) dst, bytes));
    p->device_bytes += bytes;
    if (!src.empty()) CU_TRY(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, p->stream));
    return 0;
}

static DevSym dev_sym(const ssb200_plan *p)
{
    DevSym s; s.super = p->d_super; s.pi = p->d_pi; s.px = p->d_px; s.ls = p->d_ls; s.supermap = p->d_supermap;
    s.n = p->hp.n; s.nsuper = p->hp.nsuper; return s;
}

static int configure_kernels_once()
{
    // attributes are per device context; set them every time a plan is created (cheap)
    cudaError_t e1 = cudaFuncSetAttribute(gemm_nt_sub_kernel<128, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gemm_smem_bytes<128, 16>());
    cudaError_t e2 = cudaFuncSetAttribute(gemm_nt_sub_kernel<64, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gemm_smem_bytes<64, 16>());
    cudaError_t e3 = cudaFuncSetAttribute(gemm_nt_sub_kernel<128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gemm_smem_bytes<128, 32>());
    if (e3 != cudaSuccess) e1 = e3;
    cudaError_t e5 = cudaFuncSetAttribute(potrf_block_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) potrf2_smem_bytes());
    if (e5 != cudaSuccess) e1 = e5;
    cudaError_t e4 = cudaFuncSetAttribute(trsm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) trsm_tc_smem_bytes());
    if (e4 != cudaSuccess) e1 = e4;
    cudaError_t e6 = cudaFuncSetAttribute(solve_blk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) solve_blk_smem_bytes());
    cudaError_t e7 = cudaFuncSetAttribute(solve_blk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) solve_blk_smem_bytes());
    if (e6 != cudaSuccess) e1 = e6;
    if (e7 != cudaSuccess) e1 = e7;
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2)); return SSB_CHOLMOD_GPU_PROBLEM; }
    return 0;
}

extern "C" const char *ssb200_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *ssb200_version(void) { return "suitesparse_b200 0.1 (sm_100a; CHOLMOD 3.0.14 ABI)"; }
extern "C" int ssb200_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }

static void plan_free(ssb200_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    void *ptrs[] = {p->d_probe, p->d_owner, p->d_super, p->d_ls, p->d_supermap, p->d_relmap, p->d_info, p->d_pi, p->d_px, p->d_Lx, p->d_winv, p->jobs.gemm_jobs,
                    p->jobs.gemm_tiles, p->jobs.potrf_jobs, p->jobs.trsm_jobs, p->jobs.trsm_tiles, p->d_solve_jobs, p->d_solve_tiles,
                    p->d_solve2_jobs, p->d_solve2_tiles, p->d_solve_blks, p->d_solve_blk_ctas, p->d_sol_scratch, p->d_sol_flags,
                    p->d_X};
    for (void *q : ptrs) if (q) cudaFree(q);
    free_cscbuf(p->bufA); free_cscbuf(p->bufF);
    for (auto &g : p->sg) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (p->fgraph) cudaGraphExecDestroy(p->fgraph);
    if (p->h_info) cudaFreeHost(p->h_info);
    for (auto e : p->events) cudaEventDestroy(e);
    if (p->copier) { p->copier->shutdown(); delete p->copier; }
    for (auto e : p->copy_gates) cudaEventDestroy(e);
    if (p->copy_gate) cudaEventDestroy(p->copy_gate);
    if (p->copy_done) cudaEventDestroy(p->copy_done);
    for (auto e : p->la_events) cudaEventDestroy(e);
    if (p->panel_stream) cudaStreamDestroy(p->panel_stream);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    delete p;
}

static int upload_jobs(ssb200_plan *p, const HostPlan &hp, DevJobs &dj)
{
    if (dev_alloc_copy(p, &dj.gemm_jobs, hp.gemm_jobs)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &dj.gemm_tiles, hp.gemm_tiles)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &dj.potrf_jobs, hp.potrf_jobs)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &dj.trsm_jobs, hp.trsm_jobs)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &dj.trsm_tiles, hp.trsm_tiles)) return SSB_CHOLMOD_GPU_PROBLEM;
    return 0;
}
static void free_jobs(DevJobs &dj)
{
    void *ptrs[] = {dj.gemm_jobs, dj.gemm_tiles, dj.potrf_jobs, dj.trsm_jobs, dj.trsm_tiles};
    for (void *q : ptrs) if (q) cudaFree(q);
    dj = DevJobs();
}

static int plan_build_device(ssb200_plan *p)
{
    HostPlan &hp = p->hp;
    if (configure_kernels_once()) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &p->d_super, hp.super)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &p->d_pi, hp.pi)) return SSB_CHOLMOD_GPU_PROBLEM;
    {
        // distributed storage: local offsets; a transient root is marked -2 - (base of its packed own panels) for scatter_A_kernel
        std::vector<long long> dpx = hp.compact ? hp.lpx : hp.px;
        if (hp.compact) for (long long t = 0; t < hp.nsuper; t++) if (hp.transient[t]) dpx[t] = -2 - hp.tr_own_base[t];
        if (dev_alloc_copy(p, &p->d_px, dpx)) return SSB_CHOLMOD_GPU_PROBLEM;
        CU_TRY(cudaStreamSynchronize(p->stream));           // dpx is a temporary
    }
    if (dev_alloc_copy(p, &p->d_ls, hp.ls)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &p->d_supermap, hp.supermap)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (hp.nranks > 1) if (dev_alloc_copy(p, &p->d_owner, hp.owner)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (const char *fa = getenv("SSB200_FORCE_ATOMIC")) { if (atoi(fa)) for (auto &g : hp.gemm_jobs) g.atomic = 1; }
    if (upload_jobs(p, hp, p->jobs)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &p->d_solve_jobs, hp.solve_jobs)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &p->d_solve_tiles, hp.solve_tiles)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &p->d_solve2_jobs, hp.solve2_jobs)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &p->d_solve2_tiles, hp.solve2_tiles)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &p->d_solve_blks, hp.solve_blks)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (dev_alloc_copy(p, &p->d_solve_blk_ctas, hp.solve_blk_ctas)) return SSB_CHOLMOD_GPU_PROBLEM;
    { const size_t fb = std::max<size_t>(hp.solve_blks.size(), 1) * sizeof(int); CU_TRY(cudaMalloc((void **) &p->d_sol_flags, fb)); p->device_bytes += fb; }
    p->lx_alloc = hp.compact ? hp.lxsize : hp.xsize;
    const size_t xbytes = std::max<long long>(p->lx_alloc, 1) * sizeof(double);
    CU_TRY(cudaMalloc((void **) &p->d_Lx, xbytes)); p->device_bytes += xbytes;
    p->winv_slots = std::max(1, hp.max_winv_slots);
    const size_t wbytes = (size_t) p->winv_slots * NB_INNER * NB_INNER * sizeof(double);
    CU_TRY(cudaMalloc((void **) &p->d_winv, wbytes)); p->device_bytes += wbytes;
    CU_TRY(cudaMalloc((void **) &p->d_probe, 64)); p->device_bytes += 64;
    const size_t ibytes = std::max<long long>(hp.nsuper, 1) * sizeof(int);
    CU_TRY(cudaMalloc((void **) &p->d_info, ibytes)); p->device_bytes += ibytes;
    CU_TRY(cudaMallocHost((void **) &p->h_info, ibytes));
    const size_t rbytes = std::max<long long>(hp.relmap_size, 1) * sizeof(int);
    CU_TRY(cudaMalloc((void **) &p->d_relmap, rbytes)); p->device_bytes += rbytes;
    // relative maps of all updates, computed on the device
    if (!hp.updates.empty()) {
        std::vector<DevUpdate> du(hp.updates.size());
        for (size_t t = 0; t < du.size(); t++) {
            const Update &u = hp.updates[t];
            du[t].ls_d = hp.pi[u.d] + u.p0; du[t].ls_s = hp.pi[u.s]; du[t].map_off = u.map_off;
            du[t].nsrow_s = (int) (hp.pi[u.s + 1] - hp.pi[u.s]); du[t].nd2 = u.nd2;
        }
        DevUpdate *d_du = nullptr;
        CU_TRY(cudaMalloc((void **) &d_du, du.size() * sizeof(DevUpdate)));
        CU_TRY(cudaMemcpyAsync(d_du, du.data(), du.size() * sizeof(DevUpdate), cudaMemcpyHostToDevice, p->stream));
        const size_t chunk = 1u << 30;
        for (size_t b = 0; b < du.size(); b += chunk) {
            const unsigned g = (unsigned) std::min(chunk, du.size() - b);
            relmap_kernel<<<g, 128, 0, p->stream>>>(d_du + b, p->d_ls, p->d_relmap);
        }
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaStreamSynchronize(p->stream));
        cudaFree(d_du);
    }
    CU_TRY(cudaStreamSynchronize(p->stream));
    p->stats.nsuper = hp.nsuper; p->stats.nlevels = hp.nlevels; p->stats.nupdates = (ssb_long) hp.updates.size();
    p->stats.flops_update = hp.flops_update; p->stats.flops_potrf = hp.flops_potrf; p->stats.flops_trsm = hp.flops_trsm;
    p->stats.bytes_update_panel = hp.bytes_update_panel; p->stats.bytes_update_scatter = hp.bytes_update_scatter;
    p->stats.device_bytes = (ssb_long) p->device_bytes;
    return 0;
}

static ssb200_plan *plan_create_impl(ssb_long n, ssb_long nsuper, const ssb_long *super, const ssb_long *pi, const ssb_long *px,
                                     const ssb_long *s, int device, int nranks, int rank, bool compact = false)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: the hot path has no CPU fallback"); return nullptr; }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= ndev) { set_error("device ordinal out of range"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
    ssb200_plan *p = new ssb200_plan();
    p->device = device; p->bufA = new CscBuf(); p->bufF = new CscBuf();
    if (!build_host_plan(n, nsuper, (const long long *) super, (const long long *) pi, (const long long *) px, (const long long *) s,
                         nranks, rank, p->hp, compact)) {
        set_error("invalid symbolic factor: " + p->hp.error); delete p; return nullptr;
    }
    if (cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->copy_gate, cudaEventDisableTiming) != cudaSuccess || cudaEventCreate(&p->copy_done) != cudaSuccess) { set_error("cudaStreamCreate failed"); delete p; return nullptr; }
    p->stream = p->own_stream;
    {
        int lo = 0, hi = 0;                  // numerically lowest value = highest priority
        if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) { lo = 0; hi = 0; }
        if (cudaStreamCreateWithPriority(&p->panel_stream, cudaStreamNonBlocking, hi) != cudaSuccess) { set_error("cudaStreamCreateWithPriority failed"); plan_free(p); return nullptr; }
        p->panel_prio = hi;
        if (const char *e = getenv("SSB200_LOOKAHEAD")) p->lookahead = atoi(e) != 0;
        if (nranks > 1 && !compact) p->lookahead = false;   // the multi-process sharded path drives the launches from outside, on one stream
    }
    if (plan_build_device(p) != 0) { plan_free(p); return nullptr; }
    return p;
}

extern "C" ssb200_plan *ssb200_plan_create(ssb_long n, ssb_long nsuper, const ssb_long *super, const ssb_long *pi, const ssb_long *px,
                                           const ssb_long *s, int device)
{
    return plan_create_impl(n, nsuper, super, pi, px, s, device, 1, 0);
}

extern "C" void ssb200_plan_destroy(ssb200_plan *plan) { plan_free(plan); }

// ---- sharded factorization: one process per GPU, the caller (suitesparse_b200/dist.py) moves the finished Lx ranges
// between the ranks with torch.distributed / NCCL broadcasts on the same stream ---------------------------------------
extern "C" ssb200_plan *ssb200_plan_create_dist(ssb_long n, ssb_long nsuper, const ssb_long *super, const ssb_long *pi, const ssb_long *px,
                                                const ssb_long *s, int device, int nranks, int rank)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad rank / nranks"); return nullptr; }
    return plan_create_impl(n, nsuper, super, pi, px, s, device, nranks, rank);
}

extern "C" int ssb200_set_stream(ssb200_plan *p, void *stream)
{
    if (!p) return SSB_CHOLMOD_INVALID;
    p->stream = stream ? (cudaStream_t) stream : p->own_stream;
    return 0;
}

extern "C" ssb_long ssb200_dist_num_steps(const ssb200_plan *p) { return p ? (ssb_long) p->hp.steps.size() : 0; }

// step k: *src = broadcasting rank (-1: none), [*off, *off+*cnt) = the Lx range that is final on src after part 0 of the
// step's launches; *wait_remote != 0: every broadcast started earlier must have landed before the step's launches run
extern "C" int ssb200_dist_step_info(const ssb200_plan *p, ssb_long k, int *src, ssb_long *off, ssb_long *cnt, int *wait_remote)
{
    if (!p || k < 0 || k >= (ssb_long) p->hp.steps.size()) return SSB_CHOLMOD_INVALID;
    const DistStep &st = p->hp.steps[k];
    *src = st.bcast_src; *off = st.off; *cnt = st.cnt; *wait_remote = st.wait_remote;
    return 0;
}

static int scatter_A(ssb200_plan *p, double beta0, long long kfirst, long long kcount, bool ignore_owner);
static int handle_not_posdef(ssb200_plan *p, int sfail, int info, double beta0, int quick_return, ssb_long *minor_out);
static int run_launch(ssb200_plan *p, const Launch &L, const DevJobs &dj, bool two_streams = false);

// zero Lx, assemble the columns this rank computes; everything is enqueued on the plan's stream, nothing is synchronized
extern "C" int ssb200_dist_begin(ssb200_plan *p, const double beta[2])
{
    if (!p || !p->haveA) { set_error("no plan / no matrix uploaded"); return SSB_CHOLMOD_INVALID; }
    HostPlan &hp = p->hp;
    CU_TRY(cudaSetDevice(p->device));
    p->stats.kernel_launches = 0; p->factor_on_device = false;
    if (hp.nsuper == 0) return 0;
    CU_TRY(cudaMemsetAsync(p->d_Lx, 0, (size_t) hp.xsize * sizeof(double), p->stream));
    fill_int_kernel<<<(unsigned) ((hp.nsuper + 255) / 256), 256, 0, p->stream>>>(p->d_info, hp.nsuper, INT_MAX);
    p->stats.kernel_launches++;
    p->last_beta0 = beta ? beta[0] : 0.0;
    return scatter_A(p, p->last_beta0, 0, hp.n, false);
}

// part 0: the launches before the step's broadcast starts; part 1: the look-ahead launches that overlap it
extern "C" int ssb200_dist_run_step(ssb200_plan *p, ssb_long k, int part)
{
    if (!p || k < 0 || k >= (ssb_long) p->hp.steps.size()) return SSB_CHOLMOD_INVALID;
    const DistStep &st = p->hp.steps[k];
    const int lo = part == 0 ? st.launch_begin : st.launch_mid, hi = part == 0 ? st.launch_mid : st.launch_end;
    for (int t = lo; t < hi; t++)
        if (run_launch(p, p->hp.launches[t], p->jobs)) return SSB_CHOLMOD_GPU_PROBLEM;
    CU_TRY(cudaGetLastError());
    return 0;
}

// Synchronizes; *first_bad_column = smallest column at which one of THIS rank's diagonal blocks failed (n if none).  The
// caller takes the minimum over the ranks.  The reference's partial refactorization of the failing supernode
// (t_cholmod_super_numeric.c:944-967) is not reproduced in the sharded path: ssb200_dist_zero_from() zeroes the failing
// supernode and everything after it on every rank.
extern "C" int ssb200_dist_end(ssb200_plan *p, ssb_long *first_bad_column)
{
    if (!p) return SSB_CHOLMOD_INVALID;
    HostPlan &hp = p->hp;
    if (first_bad_column) *first_bad_column = hp.n;
    if (hp.nsuper == 0) { p->factor_on_device = true; return 0; }
    CU_TRY(cudaMemcpyAsync(p->h_info, p->d_info, hp.nsuper * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    for (long long s = 0; s < hp.nsuper; s++)
        if (p->h_info[s] != INT_MAX) { if (first_bad_column) *first_bad_column = hp.super[s] + p->h_info[s] - 1; break; }
    p->stats.kernel_launches_total += p->stats.kernel_launches;
    p->factor_on_device = true;
    p->winv_valid = (hp.nranks == 1);      // the inverses exist only on the rank that factorized the block
    return 0;
}

extern "C" int ssb200_dist_zero_from(ssb200_plan *p, ssb_long column)
{
    if (!p || column < 0 || column >= p->hp.n) return SSB_CHOLMOD_INVALID;
    const int s = p->hp.supermap[column];
    const long long off = p->hp.px[s];
    CU_TRY(cudaMemsetAsync(p->d_Lx + off, 0, (size_t) (p->hp.xsize - off) * sizeof(double), p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

// Not positive definite at global column `minor` (the minimum over the ranks): every rank zeroes the failing supernode
// and everything after it; unless quick_return or the failure is in the supernode's first column, ONE rank - the owner of
// the failing supernode, rank 0 for a panel-cyclic one; it holds all the descendants - repeats the supernode on its first
// info-1 columns (t_cholmod_super_numeric.c:944-967).  *redo_rank >= 0: the caller broadcasts Lx[*off, *off+*cnt) from it.
extern "C" int ssb200_dist_not_posdef(ssb200_plan *p, ssb_long minor, int quick_return, int *redo_rank, ssb_long *off, ssb_long *cnt)
{
    if (!p || minor < 0 || minor >= p->hp.n || !redo_rank || !off || !cnt) return SSB_CHOLMOD_INVALID;
    HostPlan &hp = p->hp;
    const int sfail = hp.supermap[minor];
    const int info = (int) (minor - hp.super[sfail]) + 1;
    *redo_rank = -1; *off = hp.px[sfail]; *cnt = hp.px[sfail + 1] - hp.px[sfail];
    const bool redo = !(info == 1 || quick_return);
    const int who = hp.owner[sfail] >= 0 ? hp.owner[sfail] : 0;
    if (redo) *redo_rank = who;
    ssb_long m = 0;
    // the non-repeating ranks only zero (handle_not_posdef with quick_return does exactly that)
    const int rc = handle_not_posdef(p, sfail, info, p->last_beta0, (redo && who == hp.rank) ? 0 : 1, &m);
    p->winv_valid = false;
    return rc;
}

// dense flops this rank executes in one factorization, and the global total (load balance of the shard)
extern "C" int ssb200_dist_flops(const ssb200_plan *p, double *mine, double *total)
{
    if (!p) return SSB_CHOLMOD_INVALID;
    *mine = p->hp.my_flops; *total = p->hp.flops_update + p->hp.flops_potrf + p->hp.flops_trsm;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// matrix upload
// ---------------------------------------------------------------------------------------------------------------
template <typename T> static int ensure_cap(ssb200_plan *p, T **ptr, size_t *cap, size_t need)
{
    if (*cap >= need && *ptr) return 0;
    if (*ptr) { cudaFree(*ptr); p->device_bytes -= *cap * sizeof(T); }
    *cap = std::max<size_t>(need, 1);
    CU_TRY(cudaMalloc((void **) ptr, *cap * sizeof(T)));
    p->device_bytes += *cap * sizeof(T);
    return 0;
}

static int upload_csc(ssb200_plan *p, const ssb_long *Hp, const ssb_long *Hi, const ssb_long *Hnz, const double *Hx, ssb_long ncol, CscBuf &b)
{
    const long long nz = Hp[ncol];     // also bounds the entries an unpacked matrix can reference
    if (ensure_cap(p, &b.p, &b.capP, (size_t) ncol + 1)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (ensure_cap(p, &b.i, &b.capI, (size_t) std::max<long long>(nz, 1))) return SSB_CHOLMOD_GPU_PROBLEM;
    if (ensure_cap(p, &b.x, &b.capX, (size_t) std::max<long long>(nz, 1))) return SSB_CHOLMOD_GPU_PROBLEM;
    CU_TRY(cudaMemcpyAsync(b.p, Hp, (ncol + 1) * sizeof(long long), cudaMemcpyHostToDevice, p->stream));
    if (nz > 0) {
        CU_TRY(cudaMemcpyAsync(b.i, Hi, nz * sizeof(long long), cudaMemcpyHostToDevice, p->stream));
        CU_TRY(cudaMemcpyAsync(b.x, Hx, nz * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    }
    b.haveNz = (Hnz != nullptr);
    if (Hnz) {
        if (ensure_cap(p, &b.nz, &b.capNz, (size_t) std::max<long long>(ncol, 1))) return SSB_CHOLMOD_GPU_PROBLEM;
        CU_TRY(cudaMemcpyAsync(b.nz, Hnz, ncol * sizeof(long long), cudaMemcpyHostToDevice, p->stream));
    }
    return 0;
}

extern "C" int ssb200_upload_A(ssb200_plan *p, int stype, const ssb_long *Ap, const ssb_long *Ai, const ssb_long *Anz, const double *Ax,
                               ssb_long ncolA, const ssb_long *Fp, const ssb_long *Fi, const ssb_long *Fnz, const double *Fx)
{
    if (!p) { set_error("null plan"); return SSB_CHOLMOD_INVALID; }
    if (stype > 0) { set_error("symmetric upper case not supported"); return SSB_CHOLMOD_INVALID; }
    if (!Ap || (Ap[ncolA] > 0 && (!Ai || !Ax))) { set_error("null matrix arrays"); return SSB_CHOLMOD_INVALID; }
    if (stype < 0 && ncolA != p->hp.n) { set_error("invalid dimensions"); return SSB_CHOLMOD_INVALID; }
    if (stype == 0 && (!Fp || (Fp[p->hp.n] > 0 && (!Fi || !Fx)))) { set_error("F invalid"); return SSB_CHOLMOD_INVALID; }
    CU_TRY(cudaSetDevice(p->device));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, p->stream);
    if (upload_csc(p, Ap, Ai, Anz, Ax, ncolA, *p->bufA)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (stype == 0)
        if (upload_csc(p, Fp, Fi, Fnz, Fx, p->hp.n, *p->bufF)) return SSB_CHOLMOD_GPU_PROBLEM;
    cudaEventRecord(e1, p->stream);
    CU_TRY(cudaStreamSynchronize(p->stream));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); p->stats.ms_h2d = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    p->stype = stype; p->haveA = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------------------------
template <typename... KArgs, typename... Args>
static cudaError_t launch_ex(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool high, int prio, Args... args)
{
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributePriority; at[0].val.priority = prio;     // explicit: also recorded in captured graph nodes
    cfg.attrs = at; cfg.numAttrs = high ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// two_streams: the launch goes to the stream its schedule entry names (panel chain: high priority), after that stream has
// waited for L.wait_ev; L.rec_ev is recorded behind it.  Otherwise everything runs on the plan's stream in list order.
static int run_launch(ssb200_plan *p, const Launch &L, const DevJobs &dj, bool two_streams)
{
    const bool high = two_streams && L.stream == 1;
    cudaStream_t st = high ? p->panel_stream : p->stream;
    const int pr = p->panel_prio;
    if (two_streams && L.wait_ev >= 0) CU_TRY(cudaStreamWaitEvent(st, p->la_events[L.wait_ev], 0));
    cudaError_t ce = cudaSuccess;
    switch (L.kind) {
    case L_GEMM_BIG: {
        static int kb32 = -1;
        if (kb32 < 0) { const char *v = getenv("SSB200_KB32"); kb32 = (v && atoi(v)) ? 1 : 0; }
        if (kb32)
            ce = launch_ex(gemm_nt_sub_kernel<128, 32>, L.ntiles, gemm_threads<128>(), gemm_smem_bytes<128, 32>(), st, high, pr, (const GemmJob *) (dj.gemm_jobs + L.job0), (const int *) (dj.gemm_tiles + L.tile0), p->d_Lx, (const int *) p->d_relmap);
        else
            ce = launch_ex(gemm_nt_sub_kernel<128, 16>, L.ntiles, gemm_threads<128>(), gemm_smem_bytes<128, 16>(), st, high, pr, (const GemmJob *) (dj.gemm_jobs + L.job0), (const int *) (dj.gemm_tiles + L.tile0), p->d_Lx, (const int *) p->d_relmap);
        break;
    }
    case L_GEMM_SMALL:
        ce = launch_ex(gemm_nt_sub_kernel<64, 16>, L.ntiles, gemm_threads<64>(), gemm_smem_bytes<64, 16>(), st, high, pr, (const GemmJob *) (dj.gemm_jobs + L.job0), (const int *) (dj.gemm_tiles + L.tile0), p->d_Lx, (const int *) p->d_relmap);
        break;
    case L_POTRF:
    {
        // version 4 (four columns per step, rank-4 updates on the fp64 tensor cores) is the default; SSB200_POTRF=3: the same
        // scheme with fp64 FMAs, =1: one barrier per column, =2: 16-column sub-panels with a warp-shuffle diagonal block
        static int ver = -1;
        if (ver < 0) { const char *v = getenv("SSB200_POTRF"); ver = v ? atoi(v) : 4; if (ver < 1 || ver > 4) ver = 4; }
        if (ver == 4) ce = launch_ex(potrf_block_kernel4, L.njobs, POTRF_THREADS, 0, st, high, pr, (const PanelJob *) (dj.potrf_jobs + L.job0), p->d_Lx, p->d_info, p->d_winv);
        else if (ver == 3) ce = launch_ex(potrf_block_kernel3, L.njobs, POTRF_THREADS, 0, st, high, pr, (const PanelJob *) (dj.potrf_jobs + L.job0), p->d_Lx, p->d_info, p->d_winv);
        else if (ver == 1) ce = launch_ex(potrf_block_kernel, L.njobs, POTRF_THREADS, 0, st, high, pr, (const PanelJob *) (dj.potrf_jobs + L.job0), p->d_Lx, p->d_info, p->d_winv);
        else ce = launch_ex(potrf_block_kernel2, L.njobs, POTRF_THREADS, potrf2_smem_bytes(), st, high, pr, (const PanelJob *) (dj.potrf_jobs + L.job0), p->d_Lx, p->d_info, p->d_winv);
    }
        break;
    case L_TRSM:
        ce = launch_ex(trsm_rows_kernel, L.ntiles, TRSM_ROWS, 0, st, high, pr, (const PanelJob *) (dj.trsm_jobs + L.job0), (const int *) (dj.trsm_tiles + L.tile0), p->d_Lx);
        break;
    case L_TRSM_TC:
        ce = launch_ex(trsm_tc_kernel, L.ntiles, TRSM_ROWS, trsm_tc_smem_bytes(), st, high, pr, (const PanelJob *) (dj.trsm_jobs + L.job0), (const int *) (dj.trsm_tiles + L.tile0), p->d_Lx, (const double *) p->d_winv);
        break;
    case L_SYNC:
        break;                                  // only the wait above
    default: set_error("bad launch kind"); return SSB_CHOLMOD_GPU_PROBLEM;
    }
    if (ce != cudaSuccess) { set_error(std::string("kernel launch: ") + cudaGetErrorString(ce)); return SSB_CHOLMOD_GPU_PROBLEM; }
    if (L.kind != L_SYNC) p->stats.kernel_launches++;
    if (high && p->capturing && L.kind != L_SYNC) {
        // remember the node: stream priorities do not carry over into a captured graph by themselves on every driver
        cudaStreamCaptureStatus cs; const cudaGraphNode_t *deps = nullptr; size_t nd = 0;
        if (cudaStreamGetCaptureInfo(st, &cs, nullptr, nullptr, &deps, &nd) == cudaSuccess && cs == cudaStreamCaptureStatusActive && nd == 1)
            p->cap_high_nodes.push_back(deps[0]);
    }
    if (two_streams && L.rec_ev >= 0) CU_TRY(cudaEventRecord(p->la_events[L.rec_ev], st));
    return 0;
}

static cudaEvent_t get_event(ssb200_plan *p, size_t idx)
{
    while (p->events.size() <= idx) { cudaEvent_t e; cudaEventCreate(&e); p->events.push_back(e); }
    return p->events[idx];
}

static int scatter_A(ssb200_plan *p, double beta0, long long kfirst, long long kcount, bool ignore_owner = false)
{
    if (kcount <= 0) return 0;
    DevCsc A{p->bufA->p, p->bufA->i, p->bufA->haveNz ? p->bufA->nz : nullptr, p->bufA->x};
    DevCsc F{p->bufF->p, p->bufF->i, p->bufF->haveNz ? p->bufF->nz : nullptr, p->bufF->x};
    const int T = 128;
    const long long g = (kcount + T - 1) / T;
    scatter_A_kernel<<<(unsigned) g, T, 0, p->stream>>>(dev_sym(p), p->stype, A, F, beta0, p->d_Lx, kfirst, kcount, ignore_owner ? nullptr : p->d_owner, p->hp.nranks, p->hp.rank);
    p->stats.kernel_launches++;
    CU_TRY(cudaGetLastError());
    return 0;
}

// Not positive definite: reproduce the reference's protocol (t_cholmod_super_numeric.c:905-968, 1052-1064).
static int handle_not_posdef(ssb200_plan *p, int sfail, int info, double beta0, int quick_return, ssb_long *minor_out)
{
    HostPlan &hp = p->hp;
    *minor_out = hp.super[sfail] + info - 1;
    const long long psx = hp.px[sfail];
    CU_TRY(cudaMemsetAsync(p->d_Lx + psx, 0, (size_t) (hp.xsize - psx) * sizeof(double), p->stream));
    if (info == 1 || quick_return) { CU_TRY(cudaStreamSynchronize(p->stream)); return 0; }
    // repeat supernode sfail, factorizing only its first info-1 columns
    const int nscol = hp.super[sfail + 1] - hp.super[sfail];
    const int nsrow = (int) (hp.pi[sfail + 1] - hp.pi[sfail]);
    const int ncol_new = info - 1;
    if (scatter_A(p, beta0, hp.super[sfail], nscol, /*ignore_owner=*/true)) return SSB_CHOLMOD_GPU_PROBLEM;
    HostPlan tmp;
    {   // descendant updates of sfail
        std::vector<GemmJob> all;
        for (const Update &u : hp.updates) {
            if (u.s != sfail) continue;
            GemmJob g{};
            g.a_off = hp.px[u.d] + u.p0; g.c_off = psx; g.map_off = u.map_off;
            g.lda = (int) (hp.pi[u.d + 1] - hp.pi[u.d]); g.ldc = nsrow; g.K = hp.super[u.d + 1] - hp.super[u.d];
            g.nd1 = u.nd1; g.nd2 = u.nd2; g.atomic = 1;
            all.push_back(g);
        }
        // reuse the plan builder's tiling through a one-level fake: route everything through the 128-tile kernel
        std::vector<int> &tiles = tmp.gemm_tiles;
        Launch L{}; L.kind = L_GEMM_BIG; L.phase = 0; L.job0 = 0; L.tile0 = 0;
        long long nt = 0; int nj = 0;
        for (GemmJob g : all) {
            g.nti = (g.nd2 + 127) / 128; g.ntj = (g.nd1 + 127) / 128; g.tile_start = (int) nt;
            long long t = 0; for (int tj = 0; tj < g.ntj; tj++) t += g.nti - tj;
            for (long long q = 0; q < t; q++) tiles.push_back(nj);
            tmp.gemm_jobs.push_back(g); nt += t; nj++;
        }
        L.njobs = nj; L.ntiles = (int) nt;
        if (nj) tmp.launches.push_back(L);
    }
    std::vector<int> one{sfail};
    append_factor_jobs(hp, one, ncol_new, tmp);
    DevJobs dj;
    size_t saved = p->device_bytes;
    if (upload_jobs(p, tmp, dj)) return SSB_CHOLMOD_GPU_PROBLEM;
    for (const Launch &L : tmp.launches) if (run_launch(p, L, dj)) return SSB_CHOLMOD_GPU_PROBLEM;
    CU_TRY(cudaGetLastError());
    // zero the columns that were not factorized
    CU_TRY(cudaMemsetAsync(p->d_Lx + psx + (long long) nsrow * ncol_new, 0, (size_t) nsrow * (nscol - ncol_new) * sizeof(double), p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    free_jobs(dj); p->device_bytes = saved;
    return 0;
}

// Is this host pointer page-locked (cudaHostAlloc / cudaHostRegister)?  Only then can finished supernodes stream to
// the host asynchronously while the factorization continues.
static bool host_is_pinned(const void *ptr)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { (void) cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// Enqueue one whole factorization on the plan's stream: zero, assemble, every launch bracketed by timing events, the
// device-to-host copies of finished ranges on the copy stream, and the read-back of the potrf info array.
// capture == true: the calls are being recorded into a CUDA graph (copies are issued here, the copy stream joins the
// capture through the gate events and is joined back at the end); otherwise the helper thread issues the copies.
static int enqueue_factorization(ssb200_plan *p, double beta0, double *Lx_host, bool streaming, bool capture, int stop_level,
                                 bool two_streams, std::vector<std::pair<int, size_t>> &marks, size_t &ev, bool staged = false)
{
    HostPlan &hp = p->hp;
    // per-launch events only make sense on one stream; events recorded by graph nodes cannot be used with cudaEventElapsedTime
    const bool timing = !capture && !two_streams;
    marks.clear(); ev = 0;
    if (!capture) cudaEventRecord(get_event(p, ev), p->stream);                   // 0: start
    ev++;
    // zero all supernodes (:305-317); in pieces, a captured memset node does not take 2^32 bytes or more
    for (size_t off = 0, tot = (size_t) hp.xsize * sizeof(double); off < tot; off += (size_t) 1 << 30)
        CU_TRY(cudaMemsetAsync((char *) p->d_Lx + off, 0, std::min<size_t>((size_t) 1 << 30, tot - off), p->stream));
    {
        const long long g = (hp.nsuper + 255) / 256;
        fill_int_kernel<<<(unsigned) g, 256, 0, p->stream>>>(p->d_info, hp.nsuper, INT_MAX);
        p->stats.kernel_launches++;
    }
    if (scatter_A(p, beta0, 0, hp.n)) return SSB_CHOLMOD_GPU_PROBLEM;
    if (!capture) cudaEventRecord(get_event(p, ev), p->stream);                   // 1: assembled
    ev++;
    if (two_streams) {
        // the panel stream never runs ahead of the assembly (its first launch may have no other event to wait for)
        CU_TRY(cudaEventRecord(p->copy_gate, p->stream));
        CU_TRY(cudaStreamWaitEvent(p->panel_stream, p->copy_gate, 0));
    }
    size_t ctask = 0, cgroup = 0;
    bool copies_captured = false;
    for (int l = 0; l < hp.nlevels && l < stop_level; l++) {
        for (int t = hp.level_launch_begin[l]; t < hp.level_launch_begin[l + 1]; t++) {
            const Launch &L = hp.launches[t];
            if (timing) cudaEventRecord(get_event(p, ev), p->stream);
            marks.push_back({t, ev}); ev++;
            if (run_launch(p, L, p->jobs, two_streams)) return SSB_CHOLMOD_GPU_PROBLEM;
            if (streaming && ctask < hp.copy_tasks.size() && hp.copy_tasks[ctask].after_launch == t) {
                // these ranges of Lx are final: copy them out behind this point of the stream that finished them
                cudaEvent_t gate = p->copy_gates[cgroup++];
                CU_TRY(cudaEventRecord(gate, (two_streams && L.stream == 1) ? p->panel_stream : p->stream));
                if (capture) { CU_TRY(cudaStreamWaitEvent(p->copy_stream, gate, 0)); copies_captured = true; }
                for (bool first = true; ctask < hp.copy_tasks.size() && hp.copy_tasks[ctask].after_launch == t; ctask++, first = false) {
                    const CopyTask &ct = hp.copy_tasks[ctask];
                    if (capture)
                        CU_TRY(cudaMemcpyAsync(Lx_host + ct.off, p->d_Lx + ct.off, (size_t) ct.cnt * sizeof(double), cudaMemcpyDeviceToHost, p->copy_stream));
                    else
                        p->copier->push(CopyJob{first ? gate : nullptr, p->d_Lx + ct.off, Lx_host + ct.off, (size_t) ct.cnt * sizeof(double), staged});
                }
            }
        }
    }
    if (two_streams && stop_level < hp.nlevels) {
        // debugging stop in the middle of the schedule: join the panel stream by hand
        cudaEvent_t e = get_event(p, hp.launches.size() + 7);
        CU_TRY(cudaEventRecord(e, p->panel_stream)); CU_TRY(cudaStreamWaitEvent(p->stream, e, 0));
    }
    if (!capture) cudaEventRecord(get_event(p, ev), p->stream);
    marks.push_back({-1, ev}); ev++;                                               // end of the compute chain
    CU_TRY(cudaMemcpyAsync(p->h_info, p->d_info, hp.nsuper * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    if (capture && copies_captured) {                                              // join the copy stream back into the capture
        CU_TRY(cudaEventRecord(p->copy_gate, p->copy_stream));
        CU_TRY(cudaStreamWaitEvent(p->stream, p->copy_gate, 0));
    }
    if (!capture) cudaEventRecord(get_event(p, ev), p->stream);
    ev++;                                                                          // everything, copies included
    CU_TRY(cudaGetLastError());
    return 0;
}

static int factorize_impl(ssb200_plan *p, const double beta[2], int quick_return_if_not_posdef, ssb_long *minor_out, double *Lx_host)
{
    if (!p) { set_error("null plan"); return SSB_CHOLMOD_INVALID; }
    if (!p->haveA) { set_error("no matrix uploaded"); return SSB_CHOLMOD_INVALID; }
    HostPlan &hp = p->hp;
    CU_TRY(cudaSetDevice(p->device));
    p->stats.kernel_launches = 0;
    if (minor_out) *minor_out = hp.n;
    p->factor_on_device = false;
    if (hp.nsuper == 0) { p->factor_on_device = true; return 0; }
    const double beta0 = beta ? beta[0] : 0.0;
    const char *stop = getenv("SSB200_DEBUG_STOP_LEVEL");                          // debugging aid: stop after this many levels
    const int stop_level = stop ? atoi(stop) : INT_MAX;
    (void) cudaGetLastError();                                                     // drop stale errors of earlier calls
    static int stream_d2h = -1, use_graph = -2;
    if (stream_d2h < 0) { const char *v = getenv("SSB200_STREAM_D2H"); stream_d2h = (v && atoi(v) == 0) ? 0 : 1; }
    if (use_graph == -2) { const char *v = getenv("SSB200_FACTOR_GRAPH"); use_graph = v ? (atoi(v) ? 1 : 0) : -1; }
    // pageable destination: staged through a pinned ring by the helper thread (SSB200_STAGE_D2H=0: one plain copy at the end)
    static int stage_d2h = -1;
    if (stage_d2h < 0) { const char *v = getenv("SSB200_STAGE_D2H"); stage_d2h = (v && atoi(v) == 0) ? 0 : 1; }
    const bool pinned_dst = Lx_host && host_is_pinned(Lx_host);
    const bool staged = Lx_host && !pinned_dst && stage_d2h && (size_t) hp.xsize >= ((size_t) 1 << 16);
    const bool streaming = Lx_host && stream_d2h && stop_level == INT_MAX && (pinned_dst || staged);
    // Graph replay (SSB200_FACTOR_GRAPH=1, or by default when the look-ahead is off and the factor streams to the host): with
    // ONE stream the ~2700 dependent launches otherwise wait behind the PCIe link that the streaming saturates (+24 us per
    // launch, round 1).  With the two-stream look-ahead schedule stream launches are faster than the replayed graph (1 125 vs
    // 1 134 ms end to end, 1 069 vs 1 083 ms resident), so that is the default.
    const bool graph = stop_level == INT_MAX && !(streaming && staged) && (use_graph == 1 || (use_graph == -1 && streaming && !(p->lookahead && hp.n_events > 0)));
    const bool two_streams = p->lookahead && hp.n_events > 0 && hp.nranks == 1;
    while ((int) p->la_events.size() < hp.n_events) { cudaEvent_t e; CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->la_events.push_back(e); }
    // everything the enqueue needs exists before a capture starts
    while (p->events.size() < hp.launches.size() + 8) { cudaEvent_t e; CU_TRY(cudaEventCreate(&e)); p->events.push_back(e); }
    {
        size_t groups = 0; int last = -1;
        for (const CopyTask &ct : hp.copy_tasks) if (ct.after_launch != last) { groups++; last = ct.after_launch; }
        while (p->copy_gates.size() < groups) { cudaEvent_t e; CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->copy_gates.push_back(e); }
    }
    if (streaming && !graph && !p->copier) { p->copier = new Copier(); p->copier->start(p->device, p->copy_stream); }
    if (streaming && staged && p->copier->ensure_ring((size_t) hp.xsize * sizeof(double))) { set_error("cannot allocate the pinned staging ring"); return SSB_CHOLMOD_GPU_PROBLEM; }
    p->stats.d2h_staged = (streaming && staged) ? 1 : 0;

    std::vector<std::pair<int, size_t>> &marks = p->fg_marks;
    size_t ev = 0;
    if (graph) {
        // The whole factorization is one CUDA graph (thousands of small dependent launches; replaying it also keeps the
        // launch chain off the PCIe link that the host streaming saturates).  Rebuilt when a baked-in parameter changes.
        FactorGraphKey key; memset(&key, 0, sizeof(key));
        key.beta0 = beta0; key.host = streaming ? Lx_host : nullptr; key.stype = p->stype; key.two_streams = two_streams ? 1 : 0;
        key.a[0] = p->bufA->p; key.a[1] = p->bufA->i; key.a[2] = p->bufA->x; key.a[3] = p->bufA->haveNz ? p->bufA->nz : nullptr;
        key.f[0] = p->bufF->p; key.f[1] = p->bufF->i; key.f[2] = p->bufF->x; key.f[3] = p->bufF->haveNz ? p->bufF->nz : nullptr;
        if (!p->fgraph || memcmp(&key, &p->fg_key, sizeof(key)) != 0) {
            if (p->fgraph) { cudaGraphExecDestroy(p->fgraph); p->fgraph = nullptr; }
            CU_TRY(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
            p->capturing = true; p->cap_high_nodes.clear();
            const int rc = enqueue_factorization(p, beta0, Lx_host, streaming, true, stop_level, two_streams, marks, ev);
            p->capturing = false;
            cudaGraph_t g = nullptr;
            cudaError_t ce = cudaStreamEndCapture(p->stream, &g);
            if (!rc && ce == cudaSuccess) {
                int nset = 0, nalready = 0;
                for (cudaGraphNode_t nd : p->cap_high_nodes) {
                    cudaLaunchAttributeValue v; memset(&v, 0, sizeof(v));
                    if (cudaGraphKernelNodeGetAttribute(nd, cudaLaunchAttributePriority, &v) == cudaSuccess && v.priority == p->panel_prio) nalready++;
                    memset(&v, 0, sizeof(v)); v.priority = p->panel_prio;
                    if (cudaGraphKernelNodeSetAttribute(nd, cudaLaunchAttributePriority, &v) == cudaSuccess) nset++;
                }
                (void) cudaGetLastError();
                if (getenv("SSB200_VERBOSE") && atoi(getenv("SSB200_VERBOSE")))
                    fprintf(stderr, "[suitesparse_b200] factorization graph: %zu panel-stream kernel nodes, priority %d set on %d (captured with it: %d)\n",
                            p->cap_high_nodes.size(), p->panel_prio, nset, nalready);
            }
            if (rc || ce != cudaSuccess) { if (g) cudaGraphDestroy(g); (void) cudaGetLastError(); if (!rc) set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce)); return SSB_CHOLMOD_GPU_PROBLEM; }
            ce = cudaGraphInstantiate(&p->fgraph, g, 0);
            cudaGraphDestroy(g);
            if (ce != cudaSuccess) { p->fgraph = nullptr; set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); return SSB_CHOLMOD_GPU_PROBLEM; }
            memcpy(&p->fg_key, &key, sizeof(key)); p->fg_ev = ev; p->fg_launches = p->stats.kernel_launches;
        }
        ev = p->fg_ev; p->stats.kernel_launches = p->fg_launches;
        CU_TRY(cudaEventRecord(p->events[0], p->stream));
        CU_TRY(cudaGraphLaunch(p->fgraph, p->stream));
        CU_TRY(cudaEventRecord(p->events[1], p->stream));
    } else {
        if (enqueue_factorization(p, beta0, Lx_host, streaming, false, stop_level, two_streams, marks, ev, staged)) return SSB_CHOLMOD_GPU_PROBLEM;
    }
    CU_TRY(cudaStreamSynchronize(p->stream));
    // timings: events[0] start, [1] assembled, one per launch, end of compute (marks.back()), [ev-1] end of everything
    float ms = 0;
    p->stats.ms_update = p->stats.ms_factor = p->stats.ms_assemble = 0;
    for (int k = 0; k < 6; k++) { p->stats.ms_kind[k] = 0; p->stats.flops_kind[k] = 0; p->stats.launches_kind[k] = 0; }
    p->launch_ms.assign(hp.launches.size(), 0.f);
    const size_t ev_compute_end = marks.back().second;
    if (graph) {
        // one replayed graph: only the total is known (copies to the host included)
        cudaEventElapsedTime(&ms, p->events[0], p->events[1]); p->stats.ms_total = ms;
        for (const Launch &L : hp.launches) if (L.kind != L_SYNC) { p->stats.flops_kind[L.kind] += L.flops; p->stats.launches_kind[L.kind]++; }
    } else if (two_streams) {
        // two concurrent streams: only the assembly and the total are meaningful
        cudaEventElapsedTime(&ms, p->events[0], p->events[1]); p->stats.ms_assemble = ms;
        cudaEventElapsedTime(&ms, p->events[0], p->events[ev_compute_end]); p->stats.ms_total = ms;
        for (const Launch &L : hp.launches) if (L.kind != L_SYNC) { p->stats.flops_kind[L.kind] += L.flops; p->stats.launches_kind[L.kind]++; }
    } else {
        cudaEventElapsedTime(&ms, p->events[0], p->events[1]); p->stats.ms_assemble = ms;
        for (size_t t = 0; t + 1 < marks.size(); t++) {
            cudaEventElapsedTime(&ms, p->events[marks[t].second], p->events[marks[t + 1].second]);
            const Launch &L = hp.launches[marks[t].first];
            p->launch_ms[marks[t].first] = ms;
            if (L.phase == 0) p->stats.ms_update += ms; else p->stats.ms_factor += ms;
            p->stats.ms_kind[L.kind] += ms; p->stats.flops_kind[L.kind] += L.flops; p->stats.launches_kind[L.kind]++;
        }
        cudaEventElapsedTime(&ms, p->events[0], p->events[ev_compute_end]); p->stats.ms_total = ms;
    }
    int status = 0;
    int sfail = -1;
    for (long long s = 0; s < hp.nsuper; s++) if (p->h_info[s] != INT_MAX) { sfail = (int) s; break; }
    double staged_tail_ms = 0;
    if (streaming && !graph) {
        const auto t0 = std::chrono::steady_clock::now();
        p->copier->drain();                             // every copy has been enqueued on the copy stream (staged: has arrived in L->x)
        staged_tail_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (p->copier->failed.load()) { p->copier->failed = 0; set_error("device-to-host streaming failed"); return SSB_CHOLMOD_GPU_PROBLEM; }
    }
    if (sfail >= 0) {
        status = SSB_CHOLMOD_NOT_POSDEF;
        ssb_long minor = hp.n;
        CU_TRY(cudaStreamSynchronize(p->copy_stream));
        if (handle_not_posdef(p, sfail, p->h_info[sfail], beta0, quick_return_if_not_posdef, &minor)) return SSB_CHOLMOD_GPU_PROBLEM;
        if (minor_out) *minor_out = minor;
    }
    p->stats.kernel_launches_total += p->stats.kernel_launches;
    p->factor_on_device = true;
    p->winv_valid = (status == 0);
    if (Lx_host) {
        if (streaming && status == 0) {
            // only the tail of the copy stream is still exposed
            if (graph) {
                p->stats.ms_d2h = 0;                     // inside the graph's total
            } else {
                CU_TRY(cudaEventRecord(p->copy_done, p->copy_stream));
                CU_TRY(cudaStreamSynchronize(p->copy_stream));
                cudaEventElapsedTime(&ms, p->events[ev_compute_end], p->copy_done); p->stats.ms_d2h = ms > 0 ? ms : 0;
                if (staged) p->stats.ms_d2h = staged_tail_ms;   // host time after the compute stream finished (DMA tail + the last moves)
            }
        } else {
            CU_TRY(cudaStreamSynchronize(p->copy_stream));
            int r2 = ssb200_download_L(p, Lx_host);        // not positive definite (rare) or pageable host memory: one plain copy
            if (r2) return r2;
        }
    }
    return status;
}

extern "C" int ssb200_factorize_resident(ssb200_plan *p, const double beta[2], int quick_return_if_not_posdef, ssb_long *minor_out)
{
    return factorize_impl(p, beta, quick_return_if_not_posdef, minor_out, nullptr);
}

extern "C" int ssb200_download_L(ssb200_plan *p, double *Lx_host)
{
    if (!p || !Lx_host) { set_error("null argument"); return SSB_CHOLMOD_INVALID; }
    CU_TRY(cudaSetDevice(p->device));
    if (p->hp.xsize == 0) return 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, p->stream);
    CU_TRY(cudaMemcpyAsync(Lx_host, p->d_Lx, (size_t) p->hp.xsize * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    cudaEventRecord(e1, p->stream);
    CU_TRY(cudaStreamSynchronize(p->stream));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); p->stats.ms_d2h = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

extern "C" int ssb200_upload_L(ssb200_plan *p, const double *Lx_host)
{
    if (!p || !Lx_host) { set_error("null argument"); return SSB_CHOLMOD_INVALID; }
    CU_TRY(cudaSetDevice(p->device));
    if (p->hp.xsize > 0) CU_TRY(cudaMemcpyAsync(p->d_Lx, Lx_host, (size_t) p->hp.xsize * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    p->factor_on_device = true;
    p->winv_valid = false;                 // the inverses of the diagonal blocks belong to the previous factor
    return 0;
}

extern "C" int ssb200_factorize(ssb200_plan *p, int stype, const ssb_long *Ap, const ssb_long *Ai, const ssb_long *Anz, const double *Ax,
                                ssb_long ncolA, const ssb_long *Fp, const ssb_long *Fi, const ssb_long *Fnz, const double *Fx,
                                const double beta[2], int quick_return_if_not_posdef, double *Lx_host, ssb_long *minor_out)
{
    int rc = ssb200_upload_A(p, stype, Ap, Ai, Anz, Ax, ncolA, Fp, Fi, Fnz, Fx);
    if (rc) return rc;
    return factorize_impl(p, beta, quick_return_if_not_posdef, minor_out, Lx_host);
}

// ---------------------------------------------------------------------------------------------------------------
// solves
// ---------------------------------------------------------------------------------------------------------------
// Which solve schedule.  Default: 64-column steps (diag + update kernel per step).  SSB200_SOLVE_BLK=1 selects the second
// schedule, where the big supernodes are solved in fused 256-column block steps (solve_blk_kernel; needs the inverses of
// the diagonal blocks, i.e. a factor computed by this library).  Measured on B200 at lap7 128^3: 32.1-35.3 ms against
// 31.6 ms for the default - the diagonal CTA's dependent load batches cost more than the launches they save - so it stays
// an experiment.
static bool use_blk_schedule(const ssb200_plan *p)
{
    const char *v = getenv("SSB200_SOLVE_BLK");
    return v && atoi(v) != 0 && p->winv_valid;
}
static const std::vector<SolveStep> &solve_schedule(const ssb200_plan *p, bool blk) { return blk ? p->hp.solve2_steps : p->hp.solve_steps; }

// scratch of the block schedule (256 doubles per job and right-hand side) and its reset before a pass
static int solve_blk_prepare(ssb200_plan *p, long long nrhs)
{
    const size_t nb = p->hp.solve_blks.size();
    if (nb == 0) return 0;
    if (nrhs > p->sol_nrhs_cap) {
        if (p->d_sol_scratch) { cudaFree(p->d_sol_scratch); p->d_sol_scratch = nullptr; }
        CU_TRY(cudaMalloc((void **) &p->d_sol_scratch, nb * SB_W * nrhs * sizeof(double)));
        p->sol_nrhs_cap = nrhs;
    }
    return 0;
}
static int solve_blk_reset(ssb200_plan *p, long long nrhs)
{
    const size_t nb = p->hp.solve_blks.size();
    if (nb == 0) return 0;
    CU_TRY(cudaMemsetAsync(p->d_sol_flags, 0, nb * sizeof(int), p->stream));
    CU_TRY(cudaMemsetAsync(p->d_sol_scratch, 0, nb * SB_W * nrhs * sizeof(double), p->stream));
    return 0;
}

// the kernels of one step of a pass (dir > 0: forward, L; dir < 0: backward, L'); returns the number of launches or < 0
static int enqueue_solve_step(ssb200_plan *p, const SolveStep &st, bool blk, int dir, double *dX, int nrhs, long long ldx)
{
    const SolveJob *jobs = (blk ? p->d_solve2_jobs : p->d_solve_jobs) + st.job0;
    const int *tiles = (blk ? p->d_solve2_tiles : p->d_solve_tiles) + st.tile0;
    const double *winv = p->winv_valid ? p->d_winv : nullptr;
    int n = 0;
    if (dir > 0) {
        if (st.njobs > 0) {
            lsolve_diag_kernel<<<st.njobs, SOLVE_THREADS, 0, p->stream>>>(jobs, p->d_Lx, winv, dX, nrhs, ldx); n++;
            if (st.ntiles > 0) { lsolve_update_kernel<<<st.ntiles, SOLVE_THREADS, 0, p->stream>>>(jobs, tiles, p->d_Lx, p->d_ls, dX, nrhs, ldx); n++; }
        }
        if (blk && st.nblk > 0) {
            solve_blk_kernel<true><<<st.nctas + st.nblk, SB_THREADS, solve_blk_smem_bytes(), p->stream>>>(p->d_solve_blks + st.blk0, p->d_solve_blk_ctas + st.cta0, p->d_Lx, p->d_ls,
                                                                                                p->d_winv, dX, nrhs, ldx, p->d_sol_scratch, p->d_sol_flags); n++;
        }
    } else {
        if (blk && st.nblk > 0) {
            solve_blk_kernel<false><<<st.nctas + st.nblk, SB_THREADS, solve_blk_smem_bytes(), p->stream>>>(p->d_solve_blks + st.blk0, p->d_solve_blk_ctas + st.cta0, p->d_Lx, p->d_ls,
                                                                                                 p->d_winv, dX, nrhs, ldx, p->d_sol_scratch, p->d_sol_flags); n++;
        }
        if (st.njobs > 0) {
            if (st.ntiles > 0) { ltsolve_update_kernel<<<st.ntiles, SOLVE_THREADS, 0, p->stream>>>(jobs, tiles, p->d_Lx, p->d_ls, dX, nrhs, ldx); n++; }
            ltsolve_diag_kernel<<<st.njobs, SOLVE_THREADS, 0, p->stream>>>(jobs, p->d_Lx, winv, dX, nrhs, ldx); n++;
        }
    }
    return n;
}

extern "C" int ssb200_solve_resident(ssb200_plan *p, int which, double *dX, ssb_long nrhs, ssb_long ldx)
{
    if (!p || (!dX && p->hp.n > 0 && nrhs > 0)) { set_error("null argument"); return SSB_CHOLMOD_INVALID; }
    if (!p->factor_on_device) { set_error("no numeric factor on the device"); return SSB_CHOLMOD_INVALID; }
    if (which < 0 || which > 2) { set_error("which must be 0 (L), 1 (L') or 2 (both)"); return SSB_CHOLMOD_INVALID; }
    CU_TRY(cudaSetDevice(p->device));
    p->stats.kernel_launches = 0;
    if (p->hp.n == 0 || nrhs == 0) return 0;
    cudaEvent_t e0 = get_event(p, 0), e1 = get_event(p, 1);
    const bool blk = use_blk_schedule(p);
    const std::vector<SolveStep> &steps = solve_schedule(p, blk);
    const int nsteps = (int) steps.size();
    if (blk && solve_blk_prepare(p, nrhs)) return SSB_CHOLMOD_GPU_PROBLEM;
    static int use_graph = -1;
    if (use_graph < 0) { const char *v = getenv("SSB200_SOLVE_GRAPH"); use_graph = (v && atoi(v) == 0) ? 0 : 1; }
    ssb200_plan::SolveGraph &sg = p->sg[which];
    const bool cached = use_graph && sg.exec && sg.X == dX && sg.nrhs == nrhs && sg.ldx == ldx && sg.winv == p->winv_valid && sg.blk == blk;
    long long launches = 0;
    if (!cached) {
        if (use_graph) {
            if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
            CU_TRY(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
        } else cudaEventRecord(e0, p->stream);
        int rc = 0;
        if (which == 0 || which == 2) {
            if (blk) rc |= solve_blk_reset(p, nrhs);
            for (int t = 0; t < nsteps && !rc; t++) { const int n = enqueue_solve_step(p, steps[t], blk, +1, dX, (int) nrhs, ldx); if (n < 0) rc = 1; else launches += n; }
        }
        if (which == 1 || which == 2) {
            if (blk) rc |= solve_blk_reset(p, nrhs);
            for (int t = nsteps - 1; t >= 0 && !rc; t--) { const int n = enqueue_solve_step(p, steps[t], blk, -1, dX, (int) nrhs, ldx); if (n < 0) rc = 1; else launches += n; }
        }
        if (use_graph) {
            cudaGraph_t g = nullptr;
            cudaError_t ce = cudaStreamEndCapture(p->stream, &g);
            if (rc || ce != cudaSuccess) { if (g) cudaGraphDestroy(g); (void) cudaGetLastError(); set_error("solve graph capture failed"); return SSB_CHOLMOD_GPU_PROBLEM; }
            cudaError_t ie = cudaGraphInstantiate(&sg.exec, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) { sg.exec = nullptr; set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); return SSB_CHOLMOD_GPU_PROBLEM; }
            sg.X = dX; sg.nrhs = nrhs; sg.ldx = ldx; sg.winv = p->winv_valid; sg.blk = blk; sg.launches = launches;
        } else if (rc) return SSB_CHOLMOD_GPU_PROBLEM;
    }
    p->stats.kernel_launches = use_graph ? sg.launches : launches;
    if (use_graph) {
        cudaEventRecord(e0, p->stream);
        CU_TRY(cudaGraphLaunch(sg.exec, p->stream));
    }
    cudaEventRecord(e1, p->stream);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(p->stream));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); p->stats.ms_total = ms;
    p->stats.kernel_launches_total += p->stats.kernel_launches;
    return 0;
}

extern "C" int ssb200_solve(ssb200_plan *p, int which, double *X, ssb_long nrhs, ssb_long ldx)
{
    if (!p || (!X && p->hp.n > 0 && nrhs > 0)) { set_error("null argument"); return SSB_CHOLMOD_INVALID; }
    const long long n = p->hp.n;
    if (n == 0 || nrhs == 0) return 0;
    if (ldx < n) { set_error("X and L dimensions must match"); return SSB_CHOLMOD_INVALID; }
    CU_TRY(cudaSetDevice(p->device));
    if (ensure_cap(p, &p->d_X, &p->capX, (size_t) n * nrhs)) return SSB_CHOLMOD_GPU_PROBLEM;
    CU_TRY(cudaMemcpy2DAsync(p->d_X, n * sizeof(double), X, ldx * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyHostToDevice, p->stream));
    int rc = ssb200_solve_resident(p, which, p->d_X, nrhs, n);
    if (rc) return rc;
    CU_TRY(cudaMemcpy2DAsync(X, ldx * sizeof(double), p->d_X, n * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

// diag_host[k] = L(k,k) of the device-resident factor (n doubles): enough for cholmod_rcond's min/max ratio and for
// log det A = 2 sum log L(k,k), without moving the factor
extern "C" int ssb200_factor_diag(ssb200_plan *p, double *diag_host)
{
    if (!p || !diag_host) { set_error("null argument"); return SSB_CHOLMOD_INVALID; }
    if (!p->factor_on_device) { set_error("no numeric factor on the device"); return SSB_CHOLMOD_INVALID; }
    const long long n = p->hp.n;
    if (n == 0) return 0;
    CU_TRY(cudaSetDevice(p->device));
    if (ensure_cap(p, &p->d_X, &p->capX, (size_t) n)) return SSB_CHOLMOD_GPU_PROBLEM;
    factor_diag_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, p->stream>>>(dev_sym(p), p->d_Lx, p->d_X);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(diag_host, p->d_X, n * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

// on (default): two-stream look-ahead schedule, only total device times are reported; off: one stream in list order with
// timing events around every launch (per-kernel statistics in ssb200_stats.ms_kind)
extern "C" int ssb200_set_lookahead(ssb200_plan *p, int on)
{
    if (!p) return SSB_CHOLMOD_INVALID;
    p->lookahead = on != 0 && (p->hp.nranks == 1 || p->hp.compact);
    return 0;
}

extern "C" double *ssb200_device_Lx(ssb200_plan *p) { return p ? p->d_Lx : nullptr; }
extern "C" ssb_long ssb200_xsize(const ssb200_plan *p) { return p ? p->hp.xsize : 0; }
extern "C" void *ssb200_stream(ssb200_plan *p) { return p ? (void *) p->stream : nullptr; }
extern "C" int ssb200_get_stats(const ssb200_plan *p, ssb200_stats *out)
{
    if (!p || !out) return SSB_CHOLMOD_INVALID;
    *out = p->stats; out->device_bytes = (ssb_long) p->device_bytes;
    return 0;
}

// tuning aid: per launch of the last factorize: kind, phase, ntiles, njobs, flops, ms, and K of its first (heaviest) job
extern "C" ssb_long ssb200_debug_launches(ssb200_plan *p, double *out, ssb_long cap)
{
    if (!p) return -1;
    const ssb_long nl = (ssb_long) p->hp.launches.size();
    if (!out || cap < nl * 7) return nl;
    for (ssb_long t = 0; t < nl; t++) {
        const Launch &L = p->hp.launches[t];
        double K = 0;
        if (L.kind == L_GEMM_BIG || L.kind == L_GEMM_SMALL) K = p->hp.gemm_jobs[L.job0].K;
        out[7 * t + 0] = L.kind; out[7 * t + 1] = L.phase; out[7 * t + 2] = L.ntiles; out[7 * t + 3] = L.njobs;
        out[7 * t + 4] = L.flops; out[7 * t + 5] = t < (ssb_long) p->launch_ms.size() ? p->launch_ms[t] : 0; out[7 * t + 6] = K;
    }
    return nl;
}

// debugging aid for tests: copy the relative maps back (size = sum of ndrow2 over updates)
extern "C" ssb_long ssb200_debug_relmap(ssb200_plan *p, int32_t *out, ssb_long cap)
{
    if (!p) return -1;
    if (out && cap >= p->hp.relmap_size && p->hp.relmap_size > 0)
        cudaMemcpy(out, p->d_relmap, p->hp.relmap_size * sizeof(int), cudaMemcpyDeviceToHost);
    return p->hp.relmap_size;
}

// ===============================================================================================================
// Multi-GPU inside one process: the elimination-tree shard (ssb_plan.cpp) with DISTRIBUTED STORAGE.  One host thread per
// device walks the rank's step list; a finished range (subtree, whole-owned top supernode, 256-column panel of a cyclic
// supernode) is PULLED over NVLink by exactly the ranks whose updates read it, with a copy kernel on peer-mapped
// pointers (mg_pull_kernel) - no NCCL, no host staging.  Cross-device ordering = CUDA events; the host threads only
// hand-shake on "this event has been recorded" flags.  The reference has no counterpart (GPU/cholmod_gpu.c:160-164:
// "can only utilize a single GPU").
// ===============================================================================================================
struct MgPiece { long long src_off, dst_off, cnt; };

// chunks of <= MG_CHUNK doubles, one WARP per chunk (grid-stride over warps): the pieces range from whole panels to the
// trailing rows of one 16-column supernode column (a few hundred doubles), so the unit of work is small; 8 loads in flight
// per lane (16-byte when both sides are aligned) keep ~2 MB in flight per device with 64 CTAs
constexpr int MG_CHUNK = 4096, MG_THREADS = 256;
__global__ void __launch_bounds__(MG_THREADS) mg_pull_kernel(const MgPiece *__restrict__ chunks, int nchunks,
                                                            const double *__restrict__ src, double *__restrict__ dst)
{
    const int lane = threadIdx.x & 31;
    const int wid = (int) blockIdx.x * (MG_THREADS / 32) + (threadIdx.x >> 5), nw = (int) gridDim.x * (MG_THREADS / 32);
    for (int c = wid; c < nchunks; c += nw) {
        const MgPiece pc = chunks[c];
        const double *__restrict__ sp = src + pc.src_off;
        double *__restrict__ dp = dst + pc.dst_off;
        if (((pc.src_off | pc.dst_off | pc.cnt) & 1) == 0) {
            const double2 *__restrict__ s2 = reinterpret_cast<const double2 *>(sp);
            double2 *__restrict__ d2 = reinterpret_cast<double2 *>(dp);
            const long long n2 = pc.cnt >> 1;
            for (long long i = lane; i < n2; i += 32 * 8) {
                double2 v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) { const long long t = i + 32 * q; v[q] = t < n2 ? s2[t] : make_double2(0.0, 0.0); }
#pragma unroll
                for (int q = 0; q < 8; q++) { const long long t = i + 32 * q; if (t < n2) d2[t] = v[q]; }
            }
        } else {
            for (long long i = lane; i < pc.cnt; i += 32 * 8) {
                double v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) { const long long t = i + 32 * q; v[q] = t < pc.cnt ? sp[t] : 0.0; }
#pragma unroll
                for (int q = 0; q < 8; q++) { const long long t = i + 32 * q; if (t < pc.cnt) dp[t] = v[q]; }
            }
        }
    }
}

struct MgDev {
    ssb200_plan *plan = nullptr;
    int device = 0;
    cudaStream_t comm = nullptr, d2h = nullptr;          // pulls (high priority), copies to the host factor
    cudaEvent_t ev_begin = nullptr, ev_done = nullptr;
    std::vector<cudaEvent_t> ev_arrived;                  // per step: this device's pull of the step's range has landed
    MgPiece *d_chunks = nullptr;                          // all chunks of all steps, device array
    std::vector<long long> chunk0; std::vector<int> nchunk;   // per step
    std::vector<cudaEvent_t> ev_solve;                    // per solve step (sync steps): this device's part of the step is done
    int rc = 0; std::string err; ssb_long bad = 0;
    double ms = 0;                                        // device time of the last factorization (assembly .. last kernel / pull)
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;         // timing
    std::vector<cudaEvent_t> ring_ev;                     // per ring slot of a transient root: its panel has been applied
    std::vector<cudaEvent_t> ev_trace;                    // SSB200_MG_TRACE=1: one timing event per step on the compute stream
    std::vector<float> trace_ms;                          // per step: device time from the start of the factorization
};

struct ssb200_mg {
    int N = 0;
    long long n = 0, xsize = 0;
    std::vector<MgDev> d;
    std::vector<cudaEvent_t> ev_ready;                    // per step: the step's range is final on its source device
    std::atomic<long long> A_epoch{0};                    // host hand-shake: device 0 holds the matrix of this factorization
    std::vector<std::atomic<long long>> ready_epoch;      // host hand-shake: ev_ready[k] has been recorded in this epoch
    std::vector<std::atomic<long long>> arrived_epoch;    // [r * nsteps + k]
    std::vector<std::atomic<long long>> solve_epoch;      // [r * solve_stride + t]
    long long solve_stride = 0;
    long long epoch = 0;
    double *d_X = nullptr; size_t capX = 0;               // right-hand sides, on device 0 (peers reach it over NVLink)
    cudaEvent_t ev_x = nullptr;
    void *pinned_ptr = nullptr; size_t pinned_bytes = 0;  // host L->x registered for all devices (ssb200_mg_pin_host)
    bool factor_on_devices = false;
    double last_ms = 0, last_solve_ms = 0;
    long long pulled_bytes = 0;                           // NVLink bytes per factorization (sum over ranks)
};

static void mg_free(ssb200_mg *m)
{
    if (!m) return;
    for (auto &dv : m->d) {
        if (!dv.plan) continue;
        cudaSetDevice(dv.device);
        for (auto e : dv.ev_arrived) if (e) cudaEventDestroy(e);
        for (auto e : dv.ev_solve) if (e) cudaEventDestroy(e);
        if (dv.ev_begin) cudaEventDestroy(dv.ev_begin);
        if (dv.ev_done) cudaEventDestroy(dv.ev_done);
        if (dv.ev_t0) cudaEventDestroy(dv.ev_t0);
        if (dv.ev_t1) cudaEventDestroy(dv.ev_t1);
        for (auto e : dv.ev_trace) if (e) cudaEventDestroy(e);
        for (auto e : dv.ring_ev) if (e) cudaEventDestroy(e);
        if (dv.d_chunks) cudaFree(dv.d_chunks);
        if (dv.comm) cudaStreamDestroy(dv.comm);
        if (dv.d2h) cudaStreamDestroy(dv.d2h);
    }
    for (size_t k = 0; k < m->ev_ready.size(); k++) if (m->ev_ready[k]) { cudaEventDestroy(m->ev_ready[k]); }
    if (m->d_X) { cudaSetDevice(m->d[0].device); cudaFree(m->d_X); }
    if (m->ev_x) cudaEventDestroy(m->ev_x);
    if (m->pinned_ptr) { if (cudaHostUnregister(m->pinned_ptr) != cudaSuccess) (void) cudaGetLastError(); }
    for (auto &dv : m->d) plan_free(dv.plan);
    delete m;
}

extern "C" void ssb200_mg_destroy(ssb200_mg *m) { mg_free(m); }

extern "C" ssb200_mg *ssb200_mg_create(ssb_long n, ssb_long nsuper, const ssb_long *super, const ssb_long *pi, const ssb_long *px,
                                       const ssb_long *s, int ndev, const int *devices)
{
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) { set_error("no CUDA device: the hot path has no CPU fallback"); return nullptr; }
    if (ndev < 2 || ndev > have) { set_error("ssb200_mg_create: need 2 <= ndev <= number of devices"); return nullptr; }
    std::vector<int> devs(ndev);
    for (int r = 0; r < ndev; r++) { devs[r] = devices ? devices[r] : r; if (devs[r] < 0 || devs[r] >= have) { set_error("device ordinal out of range"); return nullptr; } }
    // every device maps every other device's memory (NVLink / NVSwitch peer access)
    for (int a = 0; a < ndev; a++) {
        cudaSetDevice(devs[a]);
        for (int b = 0; b < ndev; b++) {
            if (a == b) continue;
            int can = 0; cudaDeviceCanAccessPeer(&can, devs[a], devs[b]);
            if (!can) { set_error("devices cannot access each other's memory (no NVLink/PCIe peer access)"); return nullptr; }
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return nullptr; }
            (void) cudaGetLastError();
        }
    }
    ssb200_mg *m = new ssb200_mg();
    m->N = ndev; m->n = n; m->xsize = nsuper > 0 ? px[nsuper] : 0;
    m->d.resize(ndev);
    {   // one plan per device, built in parallel (the host plan builder is single-threaded per rank)
        std::vector<std::thread> th;
        std::vector<std::string> errs(ndev);
        for (int r = 0; r < ndev; r++)
            th.emplace_back([&, r] {
                m->d[r].device = devs[r];
                m->d[r].plan = plan_create_impl(n, nsuper, super, pi, px, s, devs[r], ndev, r, /*compact=*/true);
                if (!m->d[r].plan) errs[r] = g_last_error;
            });
        for (auto &t : th) t.join();
        for (int r = 0; r < ndev; r++) if (!m->d[r].plan) { set_error("rank " + std::to_string(r) + ": " + errs[r]); mg_free(m); return nullptr; }
    }
    const size_t ns = m->d[0].plan->hp.steps.size();
    for (int r = 1; r < ndev; r++) {
        const auto &a = m->d[0].plan->hp.steps, &b = m->d[r].plan->hp.steps;
        bool same = a.size() == b.size();
        for (size_t k = 0; same && k < ns; k++) same = a[k].bcast_src == b[k].bcast_src && a[k].off == b[k].off && a[k].cnt == b[k].cnt && a[k].wait_remote == b[k].wait_remote;
        if (!same) { set_error("internal: the ranks disagree on the step list"); mg_free(m); return nullptr; }
    }
    m->ev_ready.assign(ns, nullptr);
    m->ready_epoch = std::vector<std::atomic<long long>>(ns);
    m->arrived_epoch = std::vector<std::atomic<long long>>(ns * ndev);
    for (auto &a : m->ready_epoch) a.store(0);
    for (auto &a : m->arrived_epoch) a.store(0);
    const size_t nsolve = std::max(m->d[0].plan->hp.solve_steps.size(), m->d[0].plan->hp.solve2_steps.size());
    m->solve_stride = (long long) nsolve;
    m->solve_epoch = std::vector<std::atomic<long long>>(nsolve * ndev);
    for (auto &a : m->solve_epoch) a.store(0);
    for (size_t k = 0; k < ns; k++) {
        const DistStep &st = m->d[0].plan->hp.steps[k];
        if (st.bcast_src < 0) continue;
        cudaSetDevice(devs[st.bcast_src]);
        if (cudaEventCreateWithFlags(&m->ev_ready[k], cudaEventDisableTiming) != cudaSuccess) { set_error("cudaEventCreate failed"); mg_free(m); return nullptr; }
    }
    for (int r = 0; r < ndev; r++) {
        MgDev &dv = m->d[r];
        HostPlan &hp = dv.plan->hp;
        cudaSetDevice(dv.device);
        int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&dv.comm, cudaStreamNonBlocking, hi) != cudaSuccess || cudaStreamCreateWithFlags(&dv.d2h, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&dv.ev_begin, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&dv.ev_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreate(&dv.ev_t0) != cudaSuccess || cudaEventCreate(&dv.ev_t1) != cudaSuccess) {
            set_error("stream/event creation failed"); mg_free(m); return nullptr; }
        dv.ev_arrived.assign(ns, nullptr); dv.chunk0.assign(ns, 0); dv.nchunk.assign(ns, 0);
        dv.ev_solve.assign(nsolve, nullptr);
        for (size_t t = 0; t < nsolve; t++) cudaEventCreateWithFlags(&dv.ev_solve[t], cudaEventDisableTiming);
        std::vector<MgPiece> chunks;
        for (size_t k = 0; k < ns; k++) {
            if (hp.step_recv[k].empty()) continue;
            const int src = hp.steps[k].bcast_src;
            const HostPlan &sp = m->d[src].plan->hp;
            cudaEventCreateWithFlags(&dv.ev_arrived[k], cudaEventDisableTiming);
            dv.chunk0[k] = (long long) chunks.size();
            for (const HostPlan::Piece &pc : hp.step_recv[k]) {
                // 1-D piece: a run of supernodes (or a panel) that both ranks store contiguously with the same inner offsets;
                // 2-D piece: the trailing rows of every column of one supernode (strided at the source, packed here)
                for (int c = 0; c < pc.ncols; c++) {
                    const long long home = pc.home_off + (long long) c * pc.src_ld;
                    const long long so = sp.local_of(home), dn = hp.local_of(home);
                    if (so < 0 || dn < 0) { set_error("internal: transfer of a supernode that one side does not store"); mg_free(m); return nullptr; }
                    for (long long o = 0; o < pc.cnt; o += MG_CHUNK) chunks.push_back(MgPiece{so + o, dn + o, std::min<long long>(MG_CHUNK, pc.cnt - o)});
                }
                m->pulled_bytes += pc.cnt * 8 * pc.ncols;
            }
            dv.nchunk[k] = (int) (chunks.size() - dv.chunk0[k]);
        }
        if (!chunks.empty()) {
            if (cudaMalloc((void **) &dv.d_chunks, chunks.size() * sizeof(MgPiece)) != cudaSuccess ||
                cudaMemcpy(dv.d_chunks, chunks.data(), chunks.size() * sizeof(MgPiece), cudaMemcpyHostToDevice) != cudaSuccess) { set_error("chunk list upload failed"); mg_free(m); return nullptr; }
        }
    }
    cudaSetDevice(devs[0]);
    cudaEventCreateWithFlags(&m->ev_x, cudaEventDisableTiming);
    return m;
}

static void touch_pages_parallel(void *ptr, size_t bytes);
// Page-lock the host factor for ALL devices (every device copies its own share out over its own PCIe link).
extern "C" int ssb200_mg_pin_host(ssb200_mg *m, double *Lx_host)
{
    if (!m) return SSB_CHOLMOD_INVALID;
    const size_t bytes = (size_t) m->xsize * sizeof(double);
    if (m->pinned_ptr == Lx_host && m->pinned_bytes == bytes) return 0;
    if (m->pinned_ptr) { if (cudaHostUnregister(m->pinned_ptr) != cudaSuccess) (void) cudaGetLastError(); m->pinned_ptr = nullptr; }
    if (!Lx_host || bytes == 0) return 0;
    touch_pages_parallel(Lx_host, bytes);
    if (cudaHostRegister(Lx_host, bytes, cudaHostRegisterPortable) != cudaSuccess) { (void) cudaGetLastError(); return 1; }   // best effort
    m->pinned_ptr = Lx_host; m->pinned_bytes = bytes;
    return 0;
}

static void mg_spin_until(const std::atomic<long long> &a, long long epoch, const std::atomic<int> &abort_flag)
{
    while (a.load(std::memory_order_acquire) != epoch) { if (abort_flag.load(std::memory_order_relaxed)) return; std::this_thread::yield(); }
}

// One factorization over all devices.  Host arrays as in ssb200_factorize.  Lx_host (may be NULL) receives the factor: every
// device copies the ranges it finished.  Returns 0 ok, 1 not positive definite (*minor_out = failing column; the factor on
// the devices is then NOT usable - see ssb200_mg_factorize's comment in the header), < 0 error.
extern "C" int ssb200_mg_factorize(ssb200_mg *m, int stype, const ssb_long *Ap, const ssb_long *Ai, const ssb_long *Anz, const double *Ax,
                                   ssb_long ncolA, const ssb_long *Fp, const ssb_long *Fi, const ssb_long *Fnz, const double *Fx,
                                   const double beta[2], double *Lx_host, ssb_long *minor_out)
{
    if (!m) { set_error("null plan"); return SSB_CHOLMOD_INVALID; }
    const int N = m->N;
    const size_t ns = m->d[0].plan->hp.steps.size();
    const long long epoch = ++m->epoch;
    const double beta0 = beta ? beta[0] : 0.0;
    const bool host_pinned = Lx_host && host_is_pinned(Lx_host);
    std::atomic<int> abort_flag{0};
    m->factor_on_devices = false;
    if (minor_out) *minor_out = m->n;
    const auto t_begin = std::chrono::steady_clock::now();
    auto worker = [&](int r) {
        MgDev &dv = m->d[r];
        ssb200_plan *p = dv.plan;
        HostPlan &hp = p->hp;
        dv.rc = 0; dv.err.clear(); dv.bad = hp.n;
        auto fail = [&](const std::string &msg) { dv.rc = SSB_CHOLMOD_GPU_PROBLEM; dv.err = msg; abort_flag.store(1); };
#define MG_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { fail(std::string(#expr) + ": " + cudaGetErrorString(_e)); return; } } while (0)
        MG_TRY(cudaSetDevice(dv.device));
        if (Ap && r == 0) {
            // the matrix crosses PCIe once (device 0); the other devices take it from there over NVLink
            if (ssb200_upload_A(p, stype, Ap, Ai, Anz, Ax, ncolA, Fp, Fi, Fnz, Fx)) { fail(g_last_error); return; }
            m->A_epoch.store(epoch, std::memory_order_release);
        } else if (Ap) {
            if (stype > 0 || (stype < 0 && ncolA != hp.n)) { fail("invalid matrix"); return; }
            auto take = [&](CscBuf &dst, const CscBuf &src, const ssb_long *Hp, const ssb_long *Hnz, ssb_long ncol) -> bool {
                const long long nz = Hp[ncol];
                if (ensure_cap(p, &dst.p, &dst.capP, (size_t) ncol + 1) || ensure_cap(p, &dst.i, &dst.capI, (size_t) std::max<long long>(nz, 1)) ||
                    ensure_cap(p, &dst.x, &dst.capX, (size_t) std::max<long long>(nz, 1))) return false;
                if (Hnz && ensure_cap(p, &dst.nz, &dst.capNz, (size_t) std::max<long long>(ncol, 1))) return false;
                dst.haveNz = Hnz != nullptr;
                mg_spin_until(m->A_epoch, epoch, abort_flag);
                if (abort_flag.load()) return false;
                const int d0 = m->d[0].device;
                bool ok = cudaMemcpyPeerAsync(dst.p, dv.device, src.p, d0, (ncol + 1) * sizeof(long long), p->stream) == cudaSuccess;
                if (nz > 0) {
                    ok = ok && cudaMemcpyPeerAsync(dst.i, dv.device, src.i, d0, nz * sizeof(long long), p->stream) == cudaSuccess;
                    ok = ok && cudaMemcpyPeerAsync(dst.x, dv.device, src.x, d0, nz * sizeof(double), p->stream) == cudaSuccess;
                }
                if (Hnz) ok = ok && cudaMemcpyPeerAsync(dst.nz, dv.device, src.nz, d0, ncol * sizeof(long long), p->stream) == cudaSuccess;
                return ok;
            };
            if (!take(*p->bufA, *m->d[0].plan->bufA, Ap, Anz, ncolA)) { if (!abort_flag.load()) fail("matrix transfer from device 0 failed"); return; }
            if (stype == 0 && !take(*p->bufF, *m->d[0].plan->bufF, Fp, Fnz, hp.n)) { if (!abort_flag.load()) fail("matrix transfer from device 0 failed"); return; }
            p->stype = stype; p->haveA = true;
        }
        else if (!p->haveA) { fail("no matrix on the devices yet"); return; }      // Ap == NULL: the matrix of the previous call
        p->stats.kernel_launches = 0; p->factor_on_device = false;
        if (hp.nsuper == 0) return;
        static int trace = -1;
        if (trace < 0) { const char *v = getenv("SSB200_MG_TRACE"); trace = (v && atoi(v)) ? 1 : 0; }
        if (trace) while (dv.ev_trace.size() < ns + 1) { cudaEvent_t e; MG_TRY(cudaEventCreate(&e)); dv.ev_trace.push_back(e); }
        const bool two = p->lookahead && hp.n_events > 0;
        while ((int) p->la_events.size() < hp.n_events) { cudaEvent_t e; MG_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->la_events.push_back(e); }
        MG_TRY(cudaEventRecord(dv.ev_t0, p->stream));
        // zero the local storage, assemble the columns this rank computes
        for (size_t off = 0, tot = (size_t) p->lx_alloc * sizeof(double); off < tot; off += (size_t) 1 << 30)
            MG_TRY(cudaMemsetAsync((char *) p->d_Lx + off, 0, std::min<size_t>((size_t) 1 << 30, tot - off), p->stream));
        fill_int_kernel<<<(unsigned) ((hp.nsuper + 255) / 256), 256, 0, p->stream>>>(p->d_info, hp.nsuper, INT_MAX);
        p->stats.kernel_launches++;
        p->last_beta0 = beta0;
        if (scatter_A(p, beta0, 0, hp.n, false)) { fail(g_last_error); return; }
        MG_TRY(cudaEventRecord(dv.ev_begin, p->stream));
        MG_TRY(cudaStreamWaitEvent(dv.comm, dv.ev_begin, 0));
        if (two) MG_TRY(cudaStreamWaitEvent(p->panel_stream, dv.ev_begin, 0));
        std::vector<cudaEvent_t> outstanding;
        std::vector<char> waited(ns, 0);
        static int wait_all_env = -1;
        if (wait_all_env < 0) { const char *v = getenv("SSB200_MG_WAIT_ALL"); wait_all_env = (v && atoi(v)) ? 1 : 0; }
        bool any_transient = false;
        for (char c : hp.transient) any_transient = any_transient || c;
        const bool wait_all = wait_all_env && !any_transient;      // the ring of a transient root needs the per-step dependencies
        std::vector<int> slot_of_step(ns, -1), consumed;
        std::vector<char> ring_busy(hp.ring_depth, 0);
        while ((int) dv.ring_ev.size() < hp.ring_depth) { cudaEvent_t e; MG_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); dv.ring_ev.push_back(e); }
        std::vector<cudaEvent_t> &ring_ev = dv.ring_ev;
        for (size_t k = 0; k < ns; k++) {
            if (abort_flag.load()) return;
            const DistStep &st = hp.steps[k];
            if (trace) MG_TRY(cudaEventRecord(dv.ev_trace[k], p->stream));
            if (wait_all) {
                if (st.wait_remote && !outstanding.empty()) {
                    for (cudaEvent_t e : outstanding) MG_TRY(cudaStreamWaitEvent(p->stream, e, 0));
                    outstanding.clear();
                }
            } else {
                // only the pulls whose data this step's launches read (the others keep streaming in behind the compute)
                for (int dep : hp.step_deps[k])
                    if (!waited[dep] && dv.ev_arrived[dep]) {
                        MG_TRY(cudaStreamWaitEvent(p->stream, dv.ev_arrived[dep], 0)); waited[dep] = 1;
                        if (slot_of_step[dep] >= 0) consumed.push_back(slot_of_step[dep]);    // this step applies a panel that sits in a ring slot
                    }
            }
            for (int t = st.launch_begin; t < st.launch_mid; t++) if (run_launch(p, hp.launches[t], p->jobs, two)) { fail(g_last_error); return; }
            if (st.bcast_src == r) {
                MG_TRY(cudaEventRecord(m->ev_ready[k], p->stream));
                m->ready_epoch[k].store(epoch, std::memory_order_release);
                if (Lx_host && host_pinned && st.cnt > 0) {
                    // final here: this device's share of the factor goes to the host while the factorization continues
                    MG_TRY(cudaStreamWaitEvent(dv.d2h, m->ev_ready[k], 0));
                    MG_TRY(cudaMemcpyAsync(Lx_host + st.off, p->d_Lx + hp.local_of(st.off), (size_t) st.cnt * sizeof(double), cudaMemcpyDeviceToHost, dv.d2h));
                }
            } else if (st.bcast_src >= 0 && dv.nchunk[k] > 0) {
                mg_spin_until(m->ready_epoch[k], epoch, abort_flag);
                if (abort_flag.load()) return;
                MG_TRY(cudaStreamWaitEvent(dv.comm, m->ev_ready[k], 0));
                // A finished panel goes to everybody: binomial relay instead of N-1 pulls from one source.  The receivers are
                // ordered starting with the owner of the NEXT panel (the critical path); receiver q (0-based) pulls from holder
                // (q+1) - 2^floor(log2(q+1)), where holder 0 is the source and holder i the receiver i-1: after round j, 2^j devices
                // hold the panel and every NVLink port carries one copy at a time.
                const int nxt = hp.step_next[k];
                int from = st.bcast_src;
                const int t_step = (int) (std::upper_bound(hp.px.begin(), hp.px.end(), st.off) - hp.px.begin()) - 1;
                const bool tr_step = !hp.transient.empty() && hp.transient[t_step];
                int ring_slot = -1;
                if (tr_step) {
                    // a panel of the transient root lands in a ring slot: its previous occupant must have been applied (the ring
                    // slots of other ranks are not stable sources either, so everybody pulls from the owner)
                    const long long nsrow_t = hp.pi[t_step + 1] - hp.pi[t_step];
                    ring_slot = (int) (((st.off - hp.px[t_step]) / ((long long) NB_MID * nsrow_t)) % hp.ring_depth);
                    if (ring_busy[ring_slot]) { MG_TRY(cudaStreamWaitEvent(dv.comm, ring_ev[ring_slot], 0)); ring_busy[ring_slot] = 0; }
                    slot_of_step[k] = ring_slot;
                }
                if (nxt >= 0 && N > 2 && !tr_step) {
                    const int q = ((r - nxt) % N + N) % N - (((st.bcast_src - nxt) % N + N) % N < ((r - nxt) % N + N) % N ? 1 : 0);   // position among the receivers
                    int pw = 1; while (2 * pw <= q + 1) pw *= 2;
                    const int holder = (q + 1) - pw;                            // 0 = source, i = receiver at position i-1
                    if (holder > 0) {
                        int pos = holder - 1, cand = nxt;                       // walk to the receiver at position `pos`
                        for (int seen = 0;; cand = (cand + 1) % N) { if (cand == st.bcast_src) continue; if (seen == pos) break; seen++; }
                        from = cand;
                        mg_spin_until(m->arrived_epoch[(size_t) from * ns + k], epoch, abort_flag);
                        if (abort_flag.load()) return;
                        MG_TRY(cudaStreamWaitEvent(dv.comm, m->d[from].ev_arrived[k], 0));
                    }
                }
                const double *src_Lx = m->d[from].plan->d_Lx;
                bool few_1d = hp.step_recv[k].size() <= 4;
                for (const HostPlan::Piece &pc : hp.step_recv[k]) few_1d = few_1d && pc.ncols == 1;
                if (few_1d) {
                    // a few large contiguous pieces (a panel, a whole supernode): the copy engines move them, no SM is taken
                    const HostPlan &sp = m->d[from].plan->hp;
                    for (const HostPlan::Piece &pc : hp.step_recv[k])
                        MG_TRY(cudaMemcpyPeerAsync(p->d_Lx + hp.local_of(pc.home_off), dv.device, src_Lx + sp.local_of(pc.home_off), m->d[from].device,
                                                   (size_t) pc.cnt * sizeof(double), dv.comm));
                } else {
                    // thousands of scattered supernodes of a finished subtree: one gather kernel on peer pointers
                    static int pull_ctas = -1;
                    if (pull_ctas < 0) { const char *v = getenv("SSB200_MG_PULL_CTAS"); pull_ctas = v ? std::max(1, atoi(v)) : 64; }
                    const int grid = std::min((dv.nchunk[k] + MG_THREADS / 32 - 1) / (MG_THREADS / 32), pull_ctas);
                    mg_pull_kernel<<<grid, MG_THREADS, 0, dv.comm>>>(dv.d_chunks + dv.chunk0[k], dv.nchunk[k], src_Lx, p->d_Lx);
                    p->stats.kernel_launches++;
                }
                MG_TRY(cudaEventRecord(dv.ev_arrived[k], dv.comm));
                m->arrived_epoch[(size_t) r * ns + k].store(epoch, std::memory_order_release);
                outstanding.push_back(dv.ev_arrived[k]);
            }
            for (int t = st.launch_mid; t < st.launch_end; t++) if (run_launch(p, hp.launches[t], p->jobs, two)) { fail(g_last_error); return; }
            for (int sl : consumed) { MG_TRY(cudaEventRecord(ring_ev[sl], p->stream)); ring_busy[sl] = 1; }     // the slot may be overwritten after this point
            consumed.clear();
        }
        for (cudaEvent_t e : outstanding) MG_TRY(cudaStreamWaitEvent(p->stream, e, 0));
        if (trace) MG_TRY(cudaEventRecord(dv.ev_trace[ns], p->stream));
        MG_TRY(cudaEventRecord(dv.ev_t1, p->stream));
        MG_TRY(cudaMemcpyAsync(p->h_info, p->d_info, hp.nsuper * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        MG_TRY(cudaStreamSynchronize(p->stream));
        MG_TRY(cudaStreamSynchronize(dv.comm));
        MG_TRY(cudaStreamSynchronize(dv.d2h));
        MG_TRY(cudaGetLastError());
        for (long long sn = 0; sn < hp.nsuper; sn++) if (p->h_info[sn] != INT_MAX) { dv.bad = hp.super[sn] + p->h_info[sn] - 1; break; }
        { float f = 0; cudaEventElapsedTime(&f, dv.ev_t0, dv.ev_t1); dv.ms = f; }
        if (trace) { dv.trace_ms.assign(ns + 1, 0.f); for (size_t k = 0; k <= ns; k++) cudaEventElapsedTime(&dv.trace_ms[k], dv.ev_t0, dv.ev_trace[k]); }
        p->stats.kernel_launches_total += p->stats.kernel_launches;
        p->factor_on_device = true; p->winv_valid = true;      // the inverses of the blocks THIS rank factorized (what its solve jobs use)
#undef MG_TRY
    };
    {
        std::vector<std::thread> th;
        for (int r = 0; r < N; r++) th.emplace_back(worker, r);
        for (auto &t : th) t.join();
    }
    m->last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    ssb_long minor = m->n;
    for (int r = 0; r < N; r++) {
        if (m->d[r].rc) { set_error("rank " + std::to_string(r) + ": " + m->d[r].err); return m->d[r].rc; }
        minor = std::min(minor, m->d[r].bad);
    }
    if (minor < m->n) { if (minor_out) *minor_out = minor; return SSB_CHOLMOD_NOT_POSDEF; }
    if (Lx_host && !host_pinned) {
        // pageable host memory: plain copies of every device's share at the end
        for (int r = 0; r < N; r++) {
            MgDev &dv = m->d[r]; HostPlan &hp = dv.plan->hp;
            cudaSetDevice(dv.device);
            for (size_t k = 0; k < ns; k++) {
                const DistStep &st = hp.steps[k];
                if (st.bcast_src != r || st.cnt <= 0) continue;
                CU_TRY(cudaMemcpy(Lx_host + st.off, dv.plan->d_Lx + hp.local_of(st.off), (size_t) st.cnt * sizeof(double), cudaMemcpyDeviceToHost));
            }
        }
    }
    m->factor_on_devices = true;
    return 0;
}

// Triangular solves on the distributed factor: every device solves with the blocks it factorized (same kernels as the
// single-GPU path), the right-hand sides live on device 0 and are read / atomically updated by the peers over NVLink
// (n doubles per right-hand side against 16*xsize bytes of local factor reads).  Steps below the subtree cut need no
// cross-device ordering; above it, before a step a device waits for the latest step of every other device (streams are
// in order, so that is a full barrier; a wait is skipped when nothing new happened on that device since the last one).
extern "C" int ssb200_mg_solve(ssb200_mg *m, int which, double *X, ssb_long nrhs, ssb_long ldx)
{
    if (!m || (!X && m->n > 0 && nrhs > 0)) { set_error("null argument"); return SSB_CHOLMOD_INVALID; }
    if (!m->factor_on_devices) { set_error("no numeric factor on the devices"); return SSB_CHOLMOD_INVALID; }
    if (which < 0 || which > 2) { set_error("which must be 0 (L), 1 (L') or 2 (both)"); return SSB_CHOLMOD_INVALID; }
    const long long n = m->n;
    if (n == 0 || nrhs == 0) return 0;
    if (ldx < n) { set_error("X and L dimensions must match"); return SSB_CHOLMOD_INVALID; }
    const int N = m->N;
    MgDev &d0 = m->d[0];
    CU_TRY(cudaSetDevice(d0.device));
    if ((size_t) n * nrhs > m->capX) { if (m->d_X) cudaFree(m->d_X); m->d_X = nullptr; m->capX = (size_t) n * nrhs; CU_TRY(cudaMalloc((void **) &m->d_X, m->capX * sizeof(double))); }
    CU_TRY(cudaMemcpy2DAsync(m->d_X, n * sizeof(double), X, ldx * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyHostToDevice, d0.plan->stream));
    CU_TRY(cudaStreamSynchronize(d0.plan->stream));
    bool blk = true;                                     // the block schedule only if every device has its inverses
    for (int r = 0; r < N; r++) blk = blk && use_blk_schedule(m->d[r].plan);
    const long long T = (long long) solve_schedule(d0.plan, blk).size();
    std::atomic<int> abort_flag{0};
    const auto t_begin = std::chrono::steady_clock::now();
    auto run_pass = [&](int r, int dir, long long ep) {
        MgDev &dv = m->d[r];
        ssb200_plan *p = dv.plan;
        const std::vector<SolveStep> &steps = solve_schedule(p, blk);
        auto fail = [&](const std::string &msg) { dv.rc = SSB_CHOLMOD_GPU_PROBLEM; dv.err = msg; abort_flag.store(1); };
        if (cudaSetDevice(dv.device) != cudaSuccess) { fail("cudaSetDevice failed"); return; }
        std::vector<long long> last_waited(N, -1);
        bool was_sync = false;          // backward pass: the first step below the cut still waits for the top
        long long launches = 0;
        if (blk && (solve_blk_prepare(p, nrhs) || solve_blk_reset(p, nrhs))) { fail(g_last_error); return; }
        for (long long q = 0; q < T; q++) {
            if (abort_flag.load()) return;
            const long long t = dir > 0 ? q : T - 1 - q;
            const SolveStep &st = steps[t];
            const bool barrier = st.sync || (dir < 0 && was_sync);
            if (st.sync) was_sync = true;
            if (st.njobs == 0 && st.nblk == 0) continue;
            if (barrier) {
                for (int o = 0; o < N; o++) {
                    if (o == r) continue;
                    long long tw = t - dir;                                     // the other device's latest non-empty step before t
                    const auto &os = solve_schedule(m->d[o].plan, blk);
                    while (tw >= 0 && tw < T && os[tw].njobs == 0 && os[tw].nblk == 0) tw -= dir;
                    if (tw < 0 || tw >= T || tw == last_waited[o]) continue;
                    mg_spin_until(m->solve_epoch[(size_t) o * m->solve_stride + tw], ep, abort_flag);
                    if (abort_flag.load()) return;
                    if (cudaStreamWaitEvent(p->stream, m->d[o].ev_solve[tw], 0) != cudaSuccess) { fail("cudaStreamWaitEvent failed"); return; }
                    last_waited[o] = tw;
                }
                if (!st.sync) was_sync = false;                                  // below the cut from here on: no more waits
            }
            const int nl = enqueue_solve_step(p, st, blk, dir, m->d_X, (int) nrhs, n);
            if (nl < 0) { fail("solve launch failed"); return; }
            launches += nl;
            if (cudaEventRecord(dv.ev_solve[t], p->stream) != cudaSuccess) { fail("cudaEventRecord failed"); return; }
            m->solve_epoch[(size_t) r * m->solve_stride + t].store(ep, std::memory_order_release);
        }
        if (cudaStreamSynchronize(p->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) { fail("solve kernels failed"); return; }
        p->stats.kernel_launches = launches;
    };
    const int dirs[2] = {which == 1 ? -1 : +1, -1};
    const int npass = which == 2 ? 2 : 1;
    for (int ps = 0; ps < npass; ps++) {
        const long long ep = ++m->epoch;                 // a fresh hand-shake epoch per pass; the join below is the barrier between passes
        for (int r = 0; r < N; r++) { m->d[r].rc = 0; m->d[r].err.clear(); }
        std::vector<std::thread> th;
        for (int r = 0; r < N; r++) th.emplace_back(run_pass, r, dirs[ps], ep);
        for (auto &t : th) t.join();
        for (int r = 0; r < N; r++) if (m->d[r].rc) { set_error("rank " + std::to_string(r) + ": " + m->d[r].err); return m->d[r].rc; }
    }
    m->last_solve_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    CU_TRY(cudaSetDevice(d0.device));
    CU_TRY(cudaMemcpy2DAsync(X, ldx * sizeof(double), m->d_X, n * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyDeviceToHost, d0.plan->stream));
    CU_TRY(cudaStreamSynchronize(d0.plan->stream));
    return 0;
}

// Host factor (CHOLMOD layout) -> every device's local storage: each device takes the supernodes it stores.  The inverses
// of the diagonal blocks are not rebuilt: the solves then substitute inside the 64-column blocks.
extern "C" int ssb200_mg_upload_L(ssb200_mg *m, const double *Lx_host)
{
    if (!m || !Lx_host) { set_error("null argument"); return SSB_CHOLMOD_INVALID; }
    for (int r = 0; r < m->N; r++) {
        MgDev &dv = m->d[r]; const HostPlan &hp = dv.plan->hp;
        CU_TRY(cudaSetDevice(dv.device));
        long long run_home = -1, run_loc = -1, run_cnt = 0;
        for (long long t = 0; t <= hp.nsuper; t++) {
            const bool tr = t < hp.nsuper && !hp.transient.empty() && hp.transient[t];
            const bool red = t < hp.nsuper && hp.lpx[t] >= 0 && !hp.rmin.empty() && hp.rmin[t] > 0;
            const bool present = t < hp.nsuper && hp.lpx[t] >= 0 && !tr && !red;
            if (present && run_home >= 0 && hp.px[t] == run_home + run_cnt && hp.lpx[t] == run_loc + run_cnt) { run_cnt += hp.px[t + 1] - hp.px[t]; continue; }
            if (run_home >= 0) CU_TRY(cudaMemcpyAsync(dv.plan->d_Lx + run_loc, Lx_host + run_home, (size_t) run_cnt * sizeof(double), cudaMemcpyHostToDevice, dv.plan->stream));
            run_home = -1;
            if (present) { run_home = hp.px[t]; run_loc = hp.lpx[t]; run_cnt = hp.px[t + 1] - hp.px[t]; }
            if (red) {
                // a remote supernode stored from row rmin on: trailing rows of every column
                const long long nsrow = hp.pi[t + 1] - hp.pi[t], ld = nsrow - hp.rmin[t];
                CU_TRY(cudaMemcpy2DAsync(dv.plan->d_Lx + hp.lpx[t], ld * sizeof(double), Lx_host + hp.px[t] + hp.rmin[t], nsrow * sizeof(double),
                                         ld * sizeof(double), (size_t) (hp.super[t + 1] - hp.super[t]), cudaMemcpyHostToDevice, dv.plan->stream));
            }
            if (tr) {
                // transient root: only the panels this rank owns are stored
                const long long nsrow = hp.pi[t + 1] - hp.pi[t];
                const int nscol = hp.super[t + 1] - hp.super[t];
                for (int J = 0; J * NB_MID < nscol; J++) {
                    if (panel_owner(hp, (int) t, J) != hp.rank) continue;
                    const long long home = hp.px[t] + (long long) J * NB_MID * nsrow, cnt = (long long) std::min(NB_MID, nscol - J * NB_MID) * nsrow;
                    CU_TRY(cudaMemcpyAsync(dv.plan->d_Lx + hp.local_of(home), Lx_host + home, (size_t) cnt * sizeof(double), cudaMemcpyHostToDevice, dv.plan->stream));
                }
            }
        }
        CU_TRY(cudaStreamSynchronize(dv.plan->stream));
        dv.plan->factor_on_device = true; dv.plan->winv_valid = false;
    }
    m->factor_on_devices = true;
    return 0;
}

// out[0] = wall ms of the last factorization (threads started -> all devices idle), out[1] = last solve ms,
// out[2] = bytes pulled over NVLink per factorization (all ranks), out[3 + r] = HBM bytes held by rank r,
// out[3 + N + r] = dense flops executed by rank r
// SSB200_MG_TRACE=1: per step k and rank r, device time (ms since the start) at which the rank's compute stream reached the
// step; out[r * (nsteps + 1) + k], last entry = end.  Also the step table: steps[k*4] = src, level-ish wait flag, cnt, launches.
extern "C" ssb_long ssb200_mg_trace(const ssb200_mg *m, float *out, ssb_long cap, ssb_long *steps, ssb_long steps_cap)
{
    if (!m) return -1;
    const ssb_long ns = (ssb_long) m->d[0].plan->hp.steps.size();
    if (out && cap >= (ns + 1) * m->N)
        for (int r = 0; r < m->N; r++)
            for (ssb_long k = 0; k <= ns; k++) out[r * (ns + 1) + k] = k < (ssb_long) m->d[r].trace_ms.size() ? m->d[r].trace_ms[k] : 0.f;
    if (steps && steps_cap >= ns * 4)
        for (ssb_long k = 0; k < ns; k++) {
            const DistStep &st = m->d[0].plan->hp.steps[k];
            steps[4 * k] = st.bcast_src; steps[4 * k + 1] = st.wait_remote; steps[4 * k + 2] = st.cnt; steps[4 * k + 3] = m->d[0].plan->hp.step_next.empty() ? -1 : m->d[0].plan->hp.step_next[k];
        }
    return ns;
}

// kernels launched by the last call, all devices
extern "C" ssb_long ssb200_mg_launches(const ssb200_mg *m)
{
    if (!m) return 0;
    ssb_long t = 0;
    for (const auto &dv : m->d) t += dv.plan->stats.kernel_launches;
    return t;
}

extern "C" int ssb200_mg_info(const ssb200_mg *m, double *out, int cap)
{
    if (!m || !out || cap < 3 + 2 * m->N) return SSB_CHOLMOD_INVALID;
    out[0] = m->last_ms; out[1] = m->last_solve_ms; out[2] = (double) m->pulled_bytes;
    if (cap >= 4 + 2 * m->N) { double mx = 0; for (const auto &dv : m->d) mx = std::max(mx, dv.ms); out[3 + 2 * m->N] = mx; }
    for (int r = 0; r < m->N; r++) { out[3 + r] = (double) m->d[r].plan->device_bytes; out[3 + m->N + r] = m->d[r].plan->hp.my_flops; }
    return m->N;
}

// ===============================================================================================================
// CHOLMOD drop-in layer
// ===============================================================================================================
typedef int (*change_factor_fn)(int, int, int, int, int, ssb_cholmod_factor *, ssb_cholmod_common *);
typedef int (*error_fn)(int, const char *, int, const char *, ssb_cholmod_common *);

static void *host_sym(const char *name) { return dlsym(RTLD_DEFAULT, name); }

// The host library's own definition of a symbol this library interposes.  RTLD_NEXT finds it when we sit in front of it in the
// main search order (LD_PRELOAD, link order).  When both libraries were dlopen'ed (ctypes), RTLD_NEXT only looks at our own
// dependencies, so walk the loaded objects and take the first other definition.
struct NextLookup { const char *name; void *self; void *found; };
static int next_lookup_cb(struct dl_phdr_info *info, size_t, void *data)
{
    NextLookup *q = (NextLookup *) data;
    if (q->found || !info->dlpi_name || !info->dlpi_name[0]) return 0;
    void *h = dlopen(info->dlpi_name, RTLD_NOLOAD | RTLD_LAZY);
    if (!h) return 0;
    void *f = dlsym(h, q->name);
    if (f && f != q->self) q->found = f;
    dlclose(h);
    return 0;
}
static void *next_definition(const char *name, void *self)
{
    void *f = dlsym(RTLD_NEXT, name);
    if (f && f != self) return f;
    NextLookup q{name, self, nullptr};
    dl_iterate_phdr(next_lookup_cb, &q);
    return q.found;
}

static int raise_error(ssb_cholmod_common *cm, int status, int line, const char *msg)
{
    static error_fn f = (error_fn) host_sym("cholmod_l_error");
    if (!f) f = (error_fn) host_sym("cholmod_l_error");
    if (f) return f(status, "suitesparse_b200/csrc/ssb_cuda.cu", line, msg, cm);
    cm->status = status;                      // host libcholmod not loaded: keep the status protocol at least
    return 1;
}
#define RAISE(cm, st, msg) raise_error(cm, st, __LINE__, msg)

// plan cache keyed by the factor object, validated by a fingerprint of its symbolic structure
struct CacheEntry {
    const ssb_cholmod_factor *L = nullptr;
    ssb200_plan *plan = nullptr;           // single-GPU plan, or
    ssb200_mg *mg = nullptr;               // the multi-GPU plan (SSB200_DEVICES lists two or more devices)
    size_t n = 0, nsuper = 0, ssize = 0, xsize = 0;
    unsigned long long sym_hash = 0;
    std::vector<long long> sample_idx; std::vector<double> sample_val;   // fingerprint of the numeric values last written to L->x
    const void *xptr = nullptr;
    void *pinned_ptr = nullptr;            // L->x range registered with cudaHostRegister (fast D2H into the caller's buffer)
    void *seen_x = nullptr; long seen_count = 0;   // L->x of the previous factorizations and how many went there (page-lock policy 1)
    size_t pinned_bytes = 0;
    struct CplxAux *cx = nullptr;          // complex factor: the plan works on the real matrix of twice the order
    long long *d_perm = nullptr;           // L->Perm on the device (cholmod_l_solve fast path)
    double *d_B = nullptr; size_t capB = 0;
};
// complex / zomplex input (cholmod_super_numeric.c:81-86, t_cholmod_super_numeric.c:41-83): the blockified real problem
struct CplxAux {
    std::vector<long long> super2, pi2, px2, s2;     // symbolic structure of the real matrix of order 2n
    long long *d_pxc = nullptr, *d_px2 = nullptr, *d_pi2 = nullptr;
    double *d_Lc = nullptr;                           // complex factor on the device (2 * xsize doubles), staging for L->x
    std::vector<long long> Ap2, Ai2, Fp2, Fi2; std::vector<double> Ax2, Fx2;
};
static std::mutex g_cache_mu;
static std::vector<CacheEntry> g_cache;

static unsigned long long hash_symbolic(const ssb_cholmod_factor *L)
{
    unsigned long long h = 1469598103934665603ULL;
    auto mix = [&](unsigned long long v) { h ^= v; h *= 1099511628211ULL; };
    const long long *super = (const long long *) L->super, *pi = (const long long *) L->pi, *px = (const long long *) L->px, *s = (const long long *) L->s;
    for (size_t t = 0; t <= L->nsuper; t++) { mix(super[t]); mix(pi[t]); mix(px[t]); }
    const size_t ss = L->ssize ? (size_t) pi[L->nsuper] : 0;
    for (size_t t = 0; t < ss; t++) mix(s[t]);          // every row index: a factor with another pattern at the same address gets a new plan
    return h;
}

static size_t cache_capacity()
{
    const char *e = getenv("SSB200_PLAN_CACHE");
    int c = e ? atoi(e) : 2;
    return (size_t) std::max(1, c);
}

static CacheEntry *cache_find(const ssb_cholmod_factor *L)
{
    for (auto &e : g_cache) if (e.L == L) return &e;
    return nullptr;
}

static void unpin(CacheEntry *e) { if (e->plan && e->plan->fgraph) { cudaGraphExecDestroy(e->plan->fgraph); e->plan->fgraph = nullptr; }   // its copy nodes point into this registration
    if (e->pinned_ptr) { if (cudaHostUnregister(e->pinned_ptr) != cudaSuccess) (void) cudaGetLastError(); e->pinned_ptr = nullptr; e->pinned_bytes = 0; } }
static void cplx_free(CplxAux *cx)
{
    if (!cx) return;
    for (void *q : {(void *) cx->d_pxc, (void *) cx->d_px2, (void *) cx->d_pi2, (void *) cx->d_Lc}) if (q) cudaFree(q);
    delete cx;
}
static void cache_drop(CacheEntry *e) { unpin(e); if (e->d_perm) cudaFree(e->d_perm); if (e->d_B) cudaFree(e->d_B); cplx_free(e->cx); plan_free(e->plan); mg_free(e->mg); g_cache.erase(g_cache.begin() + (e - g_cache.data())); }

// SSB200_DEVICES = "0,1,2,3" | "all": the devices one factorization fans out over (two or more -> ssb200_mg_*)
static std::vector<int> devices_from_env()
{
    std::vector<int> v;
    const char *e = getenv("SSB200_DEVICES");
    if (!e || !*e) return v;
    if (!strcmp(e, "all")) { int n = ssb200_device_count(); for (int i = 0; i < n; i++) v.push_back(i); return v; }
    for (const char *q = e; *q;) { char *end = nullptr; long d = strtol(q, &end, 10); if (end == q) break; v.push_back((int) d); q = (*end == ',') ? end + 1 : end; if (*end && *end != ',') break; }
    return v;
}

// Page-lock the caller's L->x so the factor streams back at PCIe speed.  Best effort: a failure only costs bandwidth.
// SSB200_PIN_HOST=0 disables it.  The registration is dropped when the plan is evicted, when L->x moves, or by
// cholmod_l_gpu_deallocate().
// A registration can go STALE: if the application frees the factor and the allocator later hands out the same
// virtual range again, CUDA would still DMA into the old physical pages.  Every call therefore probes the mapping with
// three 8-byte device-to-host copies (first, middle, last element) and re-registers when the CPU does not see them.
static bool pin_probe(CacheEntry *e, const ssb_cholmod_factor *L)
{
    ssb200_plan *p = e->plan;
    if (!p) return true;
    volatile double *x = (volatile double *) L->x;
    const size_t idx[3] = {0, L->xsize / 2, L->xsize - 1};
    static const double magic = 0x1.b200b200b200bp+77;
    double saved[3];
    for (int t = 0; t < 3; t++) saved[t] = x[idx[t]];       // a refactorization that fails later must leave L->x as it was
    bool ok = cudaMemcpyAsync(p->d_probe, &magic, sizeof(double), cudaMemcpyHostToDevice, p->stream) == cudaSuccess;
    if (ok) {
        for (size_t t : idx) x[t] = 0.0;
        for (size_t t : idx)
            if (cudaMemcpyAsync((void *) &x[t], p->d_probe, sizeof(double), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess) ok = false;
        if (cudaStreamSynchronize(p->stream) != cudaSuccess) ok = false;
        if (ok) for (size_t t : idx) { const double v = x[t]; if (memcmp(&v, &magic, sizeof(double)) != 0) ok = false; }
    }
    if (!ok) (void) cudaGetLastError();
    for (int t = 0; t < 3; t++) x[idx[t]] = saved[t];
    return ok;
}

// SSB200_PRETOUCH=1 (off by default; only matters with SSB200_PIN_HOST=2, page-locking at the first call): touch the pages
// of a fresh L->x from 16 threads before cudaHostRegister (a write of the byte that is already there: existing values
// survive).  scripts/pin_probe.cu on a B200 box, 29 GB: register untouched 9.4 s; touch 0.9 s + register 3.2 s; register
// resident pages 2.7 s.  Inside the first cholmod_l_super_numeric the gain did not show (16.2 s against 14.4 s).
static void touch_pages_parallel(void *ptr, size_t bytes)
{
    if (bytes < ((size_t) 256 << 20)) return;
    static int on = -1;
    if (on < 0) { const char *v = getenv("SSB200_PRETOUCH"); on = (v && atoi(v) != 0) ? 1 : 0; }
    if (!on) return;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nt = std::min(16u, hw);
    const size_t page = 4096, slice = ((bytes / nt) / page + 1) * page;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([=] {
            volatile char *q = (volatile char *) ptr;
            const size_t lo = (size_t) t * slice, hi = std::min(bytes, lo + slice);
            for (size_t o = lo; o < hi; o += page) q[o] = q[o];
            if (hi == bytes && bytes > 0) q[bytes - 1] = q[bytes - 1];
        });
    for (auto &x : th) x.join();
}

// First factorization into a fresh (never touched) pageable L->x: the page faults of 29 GB are the largest part of that call
// once the page-lock is gone (1.07 s even from 16 threads).  While the device works, a few threads walk L->x ahead of the
// staged copies and fault the pages in with an atomic `or 0` on one word per page - a write fault that cannot lose a value
// the moving threads store concurrently - after asking for transparent huge pages on the aligned interior (the boxes run
// THP in `madvise` mode).  Joined before cholmod_l_super_numeric returns: nothing touches L->x behind the caller's back.
struct FirstTouch {
    std::vector<std::thread> th; std::atomic<bool> stop{false};
    static inline void poke(volatile unsigned long long *w)
    {
#if defined(__x86_64__)
        asm volatile("lock; orq $0, %0" : "+m"(*w) : : "memory");
#else
        __atomic_fetch_or((unsigned long long *) w, 0ULL, __ATOMIC_RELAXED);
#endif
    }
    void start(void *ptr, size_t bytes)
    {
        size_t min_mb = 256; int on = 1;
        if (const char *v = getenv("SSB200_FIRST_TOUCH")) on = atoi(v);
        if (const char *v = getenv("SSB200_FIRST_TOUCH_MIN_MB")) min_mb = (size_t) std::max(0, atoi(v));
        if (!on || bytes < (min_mb << 20) || bytes < 8192) return;
        char *base = (char *) ptr;
        {
            const size_t huge = (size_t) 2 << 20;
            char *a = (char *) (((size_t) base + huge - 1) & ~(huge - 1));
            if (a < base + bytes) { const size_t len = (size_t) (base + bytes - a) & ~(huge - 1); if (len) (void) madvise(a, len, MADV_HUGEPAGE); }
        }
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        int nt = (int) std::min(8u, hw > 12 ? hw - 10 : 2u);
        if (const char *v = getenv("SSB200_FIRST_TOUCH_THREADS")) nt = std::max(1, atoi(v));
        const size_t chunk = (size_t) 32 << 20, nchunk = (bytes + chunk - 1) / chunk;
        for (int t = 0; t < nt; t++)
            th.emplace_back([=] {
                for (size_t c = t; c < nchunk && !stop.load(std::memory_order_relaxed); c += nt) {
                    const size_t lo = c * chunk, hi = std::min(bytes, lo + chunk);
                    // one aligned word per 4 KiB page of [lo, hi)
                    size_t o = lo == 0 ? 0 : ((((size_t) base + lo + 4095) & ~(size_t) 4095) - (size_t) base);
                    if (lo == 0) { poke((volatile unsigned long long *) base); o = (((size_t) base + 4096) & ~(size_t) 4095) - (size_t) base; }
                    for (; o + 8 <= hi; o += 4096) poke((volatile unsigned long long *) (base + o));
                }
            });
    }
    void finish() { stop = true; for (auto &t : th) t.join(); th.clear(); }
    ~FirstTouch() { finish(); }
};

// Host-only exerciser of the two helpers above (tests, no GPU needed): the page-touching threads run against a threaded
// copy of src into dst, slot by slot as the staged copies do.
extern "C" int ssb200_debug_first_touch_copy(void *dst, const void *src, size_t bytes, int mover_threads, size_t slot_bytes)
{
    if (!dst || !src || !slot_bytes) return 1;
    HostMover mv; mv.start(std::max(1, mover_threads));
    {
        FirstTouch ft; ft.start(dst, bytes);
        for (size_t o = 0; o < bytes; o += slot_bytes) mv.copy((char *) dst + o, (const char *) src + o, std::min(slot_bytes, bytes - o));
    }
    mv.shutdown();
    return 0;
}

static int g_pin_policy = -1;               // SSB200_PIN_HOST, or ssb200_set_pin_policy
extern "C" int ssb200_set_pin_policy(int policy)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    const int old = g_pin_policy;
    g_pin_policy = policy < 0 ? -1 : std::min(2, policy);   // -1: back to the environment's
    return old;
}

static void pin_host_x(CacheEntry *e, const ssb_cholmod_factor *L)
{
    // SSB200_PIN_HOST: 0 = never page-lock L->x (every factorization leaves through the staging ring: 1.145 s per call at lap7
    // 128^3, against 1.119 s into page-locked memory); 1 (default) = page-lock once the SAME L->x has been factorized into
    // SSB200_PIN_AFTER (32) times - page-locking 29 GB costs 3.1 s even when the pages exist, which 26 ms per call only pays
    // back after ~120 refactorizations, so a long loop is locked after it has spent a quarter of that; 2 = page-lock at
    // the first call (9.4 s for a fresh 29 GB factor, the behaviour of round 1).
    if (g_pin_policy < 0) { const char *v = getenv("SSB200_PIN_HOST"); g_pin_policy = v ? std::max(0, std::min(2, atoi(v))) : 1; }
    const int enabled = g_pin_policy;
    const size_t bytes = L->xsize * sizeof(double);
    // A page-lock this layer took for ANOTHER factor that was freed behind its back may still cover part of this L->x: the
    // range would look page-locked to the factorization (cudaPointerGetAttributes) while its physical pages are new ones, and
    // direct copies would land in the old pages.  Drop every such registration before L->x is used unregistered.
    auto drop_stale_overlaps = [&]() {
        bool dropped = false;
        for (auto &o : g_cache) {
            if (&o == e || !o.pinned_ptr) continue;
            const char *a0 = (const char *) o.pinned_ptr, *a1 = a0 + o.pinned_bytes, *b0 = (const char *) L->x, *b1 = b0 + bytes;
            if (a0 < b1 && b0 < a1) { unpin(&o); dropped = true; }
        }
        return dropped;
    };
    if (!enabled) { if (!e->mg) { unpin(e); drop_stale_overlaps(); } return; }
    if (e->mg) { ssb200_mg_pin_host(e->mg, (double *) L->x); return; }
    if (e->pinned_ptr == L->x && e->pinned_bytes == bytes) {
        if (pin_probe(e, L)) return;
        unpin(e);                                       // stale: the pages behind this range were replaced
    } else unpin(e);
    if (L->xsize < (1u << 16)) { drop_stale_overlaps(); return; }   // not worth it for small factors
    if (enabled == 1) {
        static long after = -1;
        if (after < 0) { const char *v = getenv("SSB200_PIN_AFTER"); after = v ? std::max(1, atoi(v)) : 32; }
        if (e->seen_x != L->x) { e->seen_x = L->x; e->seen_count = 0; }
        if (++e->seen_count <= after) { drop_stale_overlaps(); return; }   // staged copies until the loop has shown itself
    }
    touch_pages_parallel(L->x, bytes);
    cudaError_t err = cudaHostRegister(L->x, bytes, cudaHostRegisterDefault);
    if (err != cudaSuccess) {
        (void) cudaGetLastError();
        // a stale registration of another (freed) factor may overlap this range: drop those and retry once
        const bool dropped = drop_stale_overlaps();
        if (dropped) err = cudaHostRegister(L->x, bytes, cudaHostRegisterDefault);
        if (err != cudaSuccess) { (void) cudaGetLastError(); return; }
    }
    e->pinned_ptr = L->x; e->pinned_bytes = bytes;
    if (!pin_probe(e, L)) unpin(e);
}

// returns the (possibly new) entry for L; nullptr on failure
static CacheEntry *cache_get_plan(ssb_cholmod_factor *L, bool single_gpu_only = false, bool cplx = false)
{
    const unsigned long long h = hash_symbolic(L);
    CacheEntry *e = cache_find(L);
    if (e && single_gpu_only && e->mg) { cache_drop(e); e = nullptr; }
    if (e && (e->cx != nullptr) != cplx) { cache_drop(e); e = nullptr; }
    if (e && (e->n != L->n || e->nsuper != L->nsuper || e->ssize != L->ssize || e->xsize != L->xsize || e->sym_hash != h)) { cache_drop(e); e = nullptr; }
    if (e) return e;
    while (g_cache.size() >= cache_capacity()) cache_drop(&g_cache.front());
    int dev = -1;
    if (const char *d = getenv("SSB200_DEVICE")) dev = atoi(d);
    ssb200_plan *plan = nullptr; ssb200_mg *mg = nullptr;
    const std::vector<int> devs = devices_from_env();
    CplxAux *cx = nullptr;
    if (cplx) {
        // the real matrix of order 2n: supernode s has columns 2k, 2k+1 and rows 2r, 2r+1 of the complex one (one GPU)
        cx = new CplxAux();
        const long long ns = (long long) L->nsuper;
        const long long *sup = (const long long *) L->super, *pi = (const long long *) L->pi, *rows = (const long long *) L->s;
        cx->super2.resize(ns + 1); cx->pi2.resize(ns + 1); cx->px2.resize(ns + 1);
        cx->px2[0] = 0;
        for (long long t = 0; t <= ns; t++) { cx->super2[t] = 2 * sup[t]; cx->pi2[t] = 2 * pi[t]; }
        for (long long t = 0; t < ns; t++) cx->px2[t + 1] = cx->px2[t] + 4 * (pi[t + 1] - pi[t]) * (sup[t + 1] - sup[t]);
        const long long ss = ns ? pi[ns] : 0;
        cx->s2.resize(2 * ss);
        for (long long q = 0; q < ss; q++) { cx->s2[2 * q] = 2 * rows[q]; cx->s2[2 * q + 1] = 2 * rows[q] + 1; }
        if (devs.size() >= 1 && dev < 0) dev = devs[0];
        plan = ssb200_plan_create((ssb_long) (2 * L->n), (ssb_long) ns, (const ssb_long *) cx->super2.data(), (const ssb_long *) cx->pi2.data(),
                                  (const ssb_long *) cx->px2.data(), (const ssb_long *) cx->s2.data(), dev);
        bool ok = plan != nullptr;
        const size_t nb = (size_t) (ns + 1) * sizeof(long long);
        ok = ok && cudaMalloc((void **) &cx->d_pxc, nb) == cudaSuccess && cudaMalloc((void **) &cx->d_px2, nb) == cudaSuccess && cudaMalloc((void **) &cx->d_pi2, nb) == cudaSuccess &&
             cudaMalloc((void **) &cx->d_Lc, std::max<size_t>(L->xsize, 1) * 2 * sizeof(double)) == cudaSuccess;
        ok = ok && cudaMemcpy(cx->d_pxc, L->px, nb, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(cx->d_px2, cx->px2.data(), nb, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(cx->d_pi2, cx->pi2.data(), nb, cudaMemcpyHostToDevice) == cudaSuccess;
        if (!ok) { if (plan) set_error("device allocation for the complex factor failed"); plan_free(plan); cplx_free(cx); return nullptr; }
    } else
    if (devs.size() >= 2 && !single_gpu_only) {
        mg = ssb200_mg_create((ssb_long) L->n, (ssb_long) L->nsuper, (const ssb_long *) L->super, (const ssb_long *) L->pi,
                              (const ssb_long *) L->px, (const ssb_long *) L->s, (int) devs.size(), devs.data());
        if (!mg) return nullptr;
    } else {
        if (devs.size() >= 1 && dev < 0) dev = devs[0];
        plan = ssb200_plan_create((ssb_long) L->n, (ssb_long) L->nsuper, (const ssb_long *) L->super, (const ssb_long *) L->pi,
                                  (const ssb_long *) L->px, (const ssb_long *) L->s, dev);
        if (!plan) return nullptr;
    }
    CacheEntry ne; ne.L = L; ne.plan = plan; ne.mg = mg; ne.cx = cx; ne.n = L->n; ne.nsuper = L->nsuper; ne.ssize = L->ssize; ne.xsize = L->xsize; ne.sym_hash = h;
    g_cache.push_back(ne);
    return &g_cache.back();
}

static void take_value_fingerprint(CacheEntry *e, const ssb_cholmod_factor *L)
{
    const double *x = (const double *) L->x;
    e->sample_idx.clear(); e->sample_val.clear(); e->xptr = L->x;
    const size_t nd = L->xsize * (e->cx ? 2 : 1);       // doubles behind L->x
    const size_t cnt = std::min<size_t>(nd, 4096);
    if (!cnt) return;
    const size_t step = nd / cnt;
    for (size_t t = 0; t < cnt; t++) { const long long idx = (long long) (t * step); e->sample_idx.push_back(idx); e->sample_val.push_back(x[idx]); }
}
static bool value_fingerprint_ok(const CacheEntry *e, const ssb_cholmod_factor *L)
{
    if (!(e->mg ? e->mg->factor_on_devices : e->plan->factor_on_device) || e->xptr != L->x) return false;
    const double *x = (const double *) L->x;
    for (size_t t = 0; t < e->sample_idx.size(); t++)
        if (memcmp(&x[e->sample_idx[t]], &e->sample_val[t], sizeof(double)) != 0) return false;
    return true;
}

static bool gpu_enabled_by_env()
{
    const char *e = getenv("CHOLMOD_USE_GPU");      // same switch the reference reads (cholmod_super_symbolic.c:257-296)
    return !(e && atoi(e) == 0);
}

// Blockified copy of a complex (interleaved x) or zomplex (x, z) CHOLMOD matrix: entry a+ib at (i,j) becomes [a -b; b a] at rows
// 2i,2i+1 / columns 2j,2j+1.  lower: symmetric-lower input (entries above the diagonal are ignored, as the reference does; the
// imaginary part of the diagonal is dropped, zpotrf never reads it).
static void blockify_csc(const ssb_cholmod_sparse *A, bool lower, std::vector<long long> &p2, std::vector<long long> &i2, std::vector<double> &x2)
{
    const long long ncol = (long long) A->ncol;
    const long long *Ap = (const long long *) A->p, *Ai = (const long long *) A->i, *Anz = A->packed ? nullptr : (const long long *) A->nz;
    const double *Ax = (const double *) A->x, *Az = (const double *) A->z;
    const bool zomplex = A->xtype == SSB_CHOLMOD_ZOMPLEX;
    p2.assign(2 * ncol + 1, 0); i2.clear(); x2.clear();
    for (long long j = 0; j < ncol; j++) {
        const long long pb = Ap[j], pe = Anz ? pb + Anz[j] : Ap[j + 1];
        for (int half = 0; half < 2; half++) {             // column 2j, then column 2j+1
            for (long long q = pb; q < pe; q++) {
                const long long i = Ai[q];
                if (lower && i < j) continue;
                const double re = zomplex ? Ax[q] : Ax[2 * q];
                const double im = (lower && i == j) ? 0.0 : (zomplex ? Az[q] : Ax[2 * q + 1]);
                if (half == 0) { i2.push_back(2 * i); x2.push_back(re); i2.push_back(2 * i + 1); x2.push_back(im); }
                else {
                    if (!(lower && i == j)) { i2.push_back(2 * i); x2.push_back(-im); }      // (2j, 2j+1) of a diagonal block is above the diagonal
                    i2.push_back(2 * i + 1); x2.push_back(re);
                }
            }
            p2[2 * j + half + 1] = (long long) i2.size();
        }
    }
}

// test hook (no GPU needed): the blockified copy of a complex / zomplex matrix; returns its number of entries, fills the arrays
// when cap suffices (p2 has 2*ncol+1 entries)
extern "C" ssb_long ssb200_debug_blockify(const ssb_cholmod_sparse *A, int lower, ssb_long *p2, ssb_long *i2, double *x2, ssb_long cap)
{
    if (!A || A->xtype < SSB_CHOLMOD_COMPLEX) return -1;
    std::vector<long long> p, i; std::vector<double> x;
    blockify_csc(A, lower != 0, p, i, x);
    if (p2 && i2 && x2 && cap >= (ssb_long) i.size()) {
        memcpy(p2, p.data(), p.size() * sizeof(long long)); memcpy(i2, i.data(), i.size() * sizeof(long long)); memcpy(x2, x.data(), x.size() * sizeof(double));
    }
    return (ssb_long) i.size();
}

// complex / zomplex A (and F): factorize the blockified real matrix with the real kernels, then write CHOLMOD's complex L->x
static int factorize_complex(CacheEntry *e, const ssb_cholmod_sparse *A, const ssb_cholmod_sparse *F, const double beta[2], int quick,
                             ssb_cholmod_factor *L, ssb_long *minor)
{
    CplxAux *cx = e->cx;
    ssb200_plan *p = e->plan;
    const int stype = A->stype;
    blockify_csc(A, stype < 0, cx->Ap2, cx->Ai2, cx->Ax2);
    if (stype == 0) blockify_csc(F, false, cx->Fp2, cx->Fi2, cx->Fx2);
    ssb_long minor2 = 0;
    const int rc = ssb200_factorize(p, stype, (const ssb_long *) cx->Ap2.data(), (const ssb_long *) cx->Ai2.data(), nullptr, cx->Ax2.data(), (ssb_long) (2 * A->ncol),
                                    stype == 0 ? (const ssb_long *) cx->Fp2.data() : nullptr, stype == 0 ? (const ssb_long *) cx->Fi2.data() : nullptr, nullptr,
                                    stype == 0 ? cx->Fx2.data() : nullptr, beta, quick, nullptr, &minor2);
    if (rc < 0) return rc;
    *minor = minor2 / 2;                                   // a non-positive pivot shows up at the even column of the pair first
    if (L->xsize > 0) {
        const long long xc = (long long) L->xsize;
        cplx_compress_kernel<<<(unsigned) ((xc + 255) / 256), 256, 0, p->stream>>>(cx->d_pxc, cx->d_px2, cx->d_pi2, (long long) L->nsuper, xc, p->d_Lx, cx->d_Lc);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(L->x, cx->d_Lc, (size_t) xc * 2 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        CU_TRY(cudaStreamSynchronize(p->stream));
        p->stats.kernel_launches++;
    }
    return rc;
}

// host complex factor -> blockified real factor on the device (a factor computed elsewhere, or changed by the caller)
static int upload_complex_L(CacheEntry *e, const ssb_cholmod_factor *L)
{
    CplxAux *cx = e->cx;
    ssb200_plan *p = e->plan;
    CU_TRY(cudaSetDevice(p->device));
    const long long xc = (long long) L->xsize;
    if (xc > 0) {
        CU_TRY(cudaMemcpyAsync(cx->d_Lc, L->x, (size_t) xc * 2 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        CU_TRY(cudaMemsetAsync(p->d_Lx, 0, (size_t) p->lx_alloc * sizeof(double), p->stream));
        cplx_expand_kernel<<<(unsigned) ((xc + 255) / 256), 256, 0, p->stream>>>(cx->d_pxc, cx->d_px2, cx->d_pi2, (long long) L->nsuper, xc, cx->d_Lc, p->d_Lx);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaStreamSynchronize(p->stream));
    }
    p->factor_on_device = true; p->winv_valid = false;
    return 0;
}

extern "C" int cholmod_l_super_numeric(ssb_cholmod_sparse *A, ssb_cholmod_sparse *F, double beta[2], ssb_cholmod_factor *L,
                                       ssb_cholmod_common *Common)
{
    // ---- checks: same order and codes as cholmod_super_numeric.c:120-175 ----
    if (!Common) return 0;
    if (Common->itype != SSB_CHOLMOD_LONG || Common->dtype != SSB_CHOLMOD_DOUBLE) { Common->status = SSB_CHOLMOD_INVALID; return 0; }
    if (!L || !A) { if (Common->status != SSB_CHOLMOD_OUT_OF_MEMORY) RAISE(Common, SSB_CHOLMOD_INVALID, "argument missing"); return 0; }
    if (A->xtype < SSB_CHOLMOD_REAL || A->xtype > SSB_CHOLMOD_ZOMPLEX || !A->x || (A->xtype == SSB_CHOLMOD_ZOMPLEX && !A->z)) {
        if (Common->status != SSB_CHOLMOD_OUT_OF_MEMORY) RAISE(Common, SSB_CHOLMOD_INVALID, "invalid xtype"); return 0; }
    if (L->xtype < SSB_CHOLMOD_PATTERN || L->xtype > SSB_CHOLMOD_COMPLEX || (L->xtype != SSB_CHOLMOD_PATTERN && !L->x)) {
        if (Common->status != SSB_CHOLMOD_OUT_OF_MEMORY) RAISE(Common, SSB_CHOLMOD_INVALID, "invalid xtype"); return 0; }
    const int stype = A->stype;
    if (stype < 0) {
        if (A->nrow != A->ncol || A->nrow != L->n) { RAISE(Common, SSB_CHOLMOD_INVALID, "invalid dimensions"); return 0; }
    } else if (stype == 0) {
        if (A->nrow != L->n) { RAISE(Common, SSB_CHOLMOD_INVALID, "invalid dimensions"); return 0; }
        if (!F) { if (Common->status != SSB_CHOLMOD_OUT_OF_MEMORY) RAISE(Common, SSB_CHOLMOD_INVALID, "argument missing"); return 0; }
        if (F->xtype < SSB_CHOLMOD_REAL || F->xtype > SSB_CHOLMOD_ZOMPLEX || !F->x) { RAISE(Common, SSB_CHOLMOD_INVALID, "invalid xtype"); return 0; }
        if (A->nrow != F->ncol || A->ncol != F->nrow || F->stype != 0) { RAISE(Common, SSB_CHOLMOD_INVALID, "F invalid"); return 0; }
        if (A->xtype != F->xtype) { RAISE(Common, SSB_CHOLMOD_INVALID, "A and F must have same xtype"); return 0; }
    } else { RAISE(Common, SSB_CHOLMOD_INVALID, "symmetric upper case not supported"); return 0; }
    if (!L->is_super) { RAISE(Common, SSB_CHOLMOD_INVALID, "L not supernodal"); return 0; }
    if (L->xtype != SSB_CHOLMOD_PATTERN) {
        const bool ok = (A->xtype == SSB_CHOLMOD_REAL && L->xtype == SSB_CHOLMOD_REAL) || (A->xtype == SSB_CHOLMOD_COMPLEX && L->xtype == SSB_CHOLMOD_COMPLEX) ||
                        (A->xtype == SSB_CHOLMOD_ZOMPLEX && L->xtype == SSB_CHOLMOD_COMPLEX);
        if (!ok) { RAISE(Common, SSB_CHOLMOD_INVALID, "complex type mismatch"); return 0; }
    }
    Common->status = SSB_CHOLMOD_OK;
    // ---- scope of the B200 path: real, 64-bit indices.  No CPU fallback. ----
    const bool cplx = A->xtype != SSB_CHOLMOD_REAL;     // complex / zomplex: through the blockified real problem (one GPU)
    if (A->itype != SSB_CHOLMOD_LONG || L->itype != SSB_CHOLMOD_LONG) { RAISE(Common, SSB_CHOLMOD_INVALID, "suitesparse_b200: only the cholmod_l_ (64-bit index) interface is accelerated"); return 0; }
    if (!gpu_enabled_by_env()) { RAISE(Common, SSB_CHOLMOD_GPU_PROBLEM, "suitesparse_b200: CHOLMOD_USE_GPU=0 but this library has no CPU path; unload it to use the CPU"); return 0; }
    // ---- numeric part of L (cholmod_super_numeric.c:206-228) ----
    const bool symbolic = (L->xtype == SSB_CHOLMOD_PATTERN);
    if (symbolic) {
        static change_factor_fn cf = nullptr;
        if (!cf) cf = (change_factor_fn) host_sym("cholmod_l_change_factor");
        if (!cf) { RAISE(Common, SSB_CHOLMOD_INVALID, "suitesparse_b200: host libcholmod (cholmod_l_change_factor) not found in the process"); return 0; }
        cf(cplx ? SSB_CHOLMOD_COMPLEX : SSB_CHOLMOD_REAL, 1, 1, 1, 1, L, Common);      // a zomplex A gives a complex L (cholmod_super_numeric.c:211-223)
        if (Common->status < SSB_CHOLMOD_OK) return 0;               // L stays symbolic
    }
    L->is_ll = 1;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    {
        static int verbose = -1;
        if (verbose < 0) { const char *v = getenv("SSB200_VERBOSE"); verbose = (v && atoi(v)) ? 1 : 0; }
        if (verbose) fprintf(stderr, "[suitesparse_b200] cholmod_l_super_numeric: n=%zu nsuper=%zu xsize=%zu (CUDA path)\n", L->n, L->nsuper, L->xsize);
    }
    // On failure L is given back in the form it had on input (cholmod_super_numeric.c:235-248): a factor that was symbolic
    // goes back to CHOLMOD_PATTERN (its freshly allocated L->x holds garbage), the plan and its page-lock are dropped.
    FirstTouch first_touch;                              // joined before L->x can go away and before this function returns
    auto fail = [&](int status, const std::string &msg) {
        first_touch.finish();
        if (CacheEntry *ce = cache_find(L)) {
            if (ce->plan) { ce->plan->factor_on_device = false; ce->plan->winv_valid = false; }
            if (ce->mg) ce->mg->factor_on_devices = false;
            if (symbolic) cache_drop(ce);
        }
        if (symbolic) {
            static change_factor_fn cf2 = (change_factor_fn) host_sym("cholmod_l_change_factor");
            if (cf2) cf2(SSB_CHOLMOD_PATTERN, 1, 1, 1, 1, L, Common);
        }
        raise_error(Common, status, __LINE__, msg.c_str());
        return 0;
    };
    static int vtime = -1;
    if (vtime < 0) { const char *v = getenv("SSB200_VERBOSE"); vtime = (v && atoi(v)) ? 1 : 0; }
    const auto tv0 = std::chrono::steady_clock::now();
    CacheEntry *e = cache_get_plan(L, cplx, cplx);
    if (!e) return fail(SSB_CHOLMOD_GPU_PROBLEM, ssb200_last_error());
    const auto tv1 = std::chrono::steady_clock::now();
    if (!cplx) pin_host_x(e, L);
    const auto tv2 = std::chrono::steady_clock::now();
    if (symbolic && !cplx && !e->mg && L->xsize >= (1u << 16) && !host_is_pinned(L->x)) first_touch.start(L->x, L->xsize * sizeof(double));
    ssb_long minor = (ssb_long) L->n;
    const ssb_long *Anz_ = A->packed ? nullptr : (const ssb_long *) A->nz;
    const ssb_long *Fp_ = F ? (const ssb_long *) F->p : nullptr, *Fi_ = F ? (const ssb_long *) F->i : nullptr, *Fnz_ = (F && !F->packed) ? (const ssb_long *) F->nz : nullptr;
    const double *Fx_ = F ? (const double *) F->x : nullptr;
    int rc;
    if (cplx) rc = factorize_complex(e, A, F, beta, Common->quick_return_if_not_posdef, L, &minor);
    else if (e->mg) {
        rc = ssb200_mg_factorize(e->mg, stype, (const ssb_long *) A->p, (const ssb_long *) A->i, Anz_, (const double *) A->x, (ssb_long) A->ncol,
                                 Fp_, Fi_, Fnz_, Fx_, beta, (double *) L->x, &minor);
        if (rc == SSB_CHOLMOD_NOT_POSDEF) {
            // the reference's protocol for a matrix that is not positive definite (zero the tail, repeat the failing supernode
            // on its leading columns, t_cholmod_super_numeric.c:905-968) is implemented by the single-GPU path: run it there
            e = cache_get_plan(L, /*single_gpu_only=*/true);
            if (!e) return fail(SSB_CHOLMOD_GPU_PROBLEM, ssb200_last_error());
            pin_host_x(e, L);
        }
    }
    if (!cplx && !e->mg)
        rc = ssb200_factorize(e->plan, stype, (const ssb_long *) A->p, (const ssb_long *) A->i, Anz_, (const double *) A->x, (ssb_long) A->ncol,
                              Fp_, Fi_, Fnz_, Fx_, beta, Common->quick_return_if_not_posdef, (double *) L->x, &minor);
    if (rc < 0) return fail(rc == SSB_CHOLMOD_INVALID ? SSB_CHOLMOD_INVALID : SSB_CHOLMOD_GPU_PROBLEM, ssb200_last_error());
    if (vtime) {
        const auto tv3 = std::chrono::steady_clock::now();
        auto sec = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
        fprintf(stderr, "[suitesparse_b200] cholmod_l_super_numeric: plan %.3f s, page-lock %.3f s, factorize + copies %.3f s%s\n", sec(tv0, tv1), sec(tv1, tv2), sec(tv2, tv3),
                (!e->mg && e->plan && e->plan->stats.d2h_staged) ? " (staged)" : "");
    }
    take_value_fingerprint(e, L);
    // statistics the reference keeps in Common (cholmod_core.h:1002-1048)
    ssb200_stats st = e->mg ? e->mg->d[0].plan->stats : e->plan->stats;
    if (e->mg) {
        st.kernel_launches = 0; st.ms_total = e->mg->last_ms;
        for (auto &dv : e->mg->d) st.kernel_launches += dv.plan->stats.kernel_launches;
    }
    Common->gpu_syrk_calls = (size_t) st.nupdates; Common->gpu_gemm_calls = (size_t) st.nupdates;
    Common->gpu_potrf_calls = (size_t) st.nsuper; Common->gpu_trsm_calls = (size_t) st.nsuper;
    Common->cpu_syrk_calls = Common->cpu_gemm_calls = Common->cpu_potrf_calls = Common->cpu_trsm_calls = 0;
    Common->gpu_syrk_time = st.ms_update * 1e-3; Common->gpu_gemm_time = 0; Common->gpu_potrf_time = st.ms_factor * 1e-3; Common->gpu_trsm_time = 0;
    Common->assemble_time = st.ms_assemble * 1e-3; Common->assemble_time2 = st.ms_d2h * 1e-3;
    Common->gpuKernelTime = st.ms_total * 1e-3; Common->gpuNumKernelLaunches = (int) st.kernel_launches;
    Common->gpuFlops = (ssb_long) (st.flops_update + st.flops_potrf + st.flops_trsm);
    if (rc == SSB_CHOLMOD_NOT_POSDEF) {
        RAISE(Common, SSB_CHOLMOD_NOT_POSDEF, "matrix not positive definite");
        L->minor = (size_t) minor;
    } else {
        L->minor = L->n;
    }
    return Common->status >= SSB_CHOLMOD_OK;
}

static int super_solve_common(ssb_cholmod_factor *L, ssb_cholmod_dense *X, ssb_cholmod_dense *E, ssb_cholmod_common *Common, int which)
{
    // checks: cholmod_super_solve.c:59-97
    if (!Common) return 0;
    if (Common->itype != SSB_CHOLMOD_LONG || Common->dtype != SSB_CHOLMOD_DOUBLE) { Common->status = SSB_CHOLMOD_INVALID; return 0; }
    if (!L || !X || !E) { if (Common->status != SSB_CHOLMOD_OUT_OF_MEMORY) RAISE(Common, SSB_CHOLMOD_INVALID, "argument missing"); return 0; }
    auto bad_xtype = [](int xt, const void *x) { return xt < SSB_CHOLMOD_REAL || xt > SSB_CHOLMOD_COMPLEX || !x; };
    if (bad_xtype(L->xtype, L->x) || bad_xtype(X->xtype, X->x) || bad_xtype(E->xtype, E->x)) {
        if (Common->status != SSB_CHOLMOD_OUT_OF_MEMORY) RAISE(Common, SSB_CHOLMOD_INVALID, "invalid xtype"); return 0; }
    if (L->xtype != X->xtype) { RAISE(Common, SSB_CHOLMOD_INVALID, "L and X must have the same xtype"); return 0; }
    if (L->xtype != E->xtype) { RAISE(Common, SSB_CHOLMOD_INVALID, "L and E must have the same xtype"); return 0; }
    if (X->d < X->nrow || L->n != X->nrow) { RAISE(Common, SSB_CHOLMOD_INVALID, "X and L dimensions must match"); return 0; }
    if (E->nzmax < X->ncol * L->maxesize) { RAISE(Common, SSB_CHOLMOD_INVALID, "workspace E not large enough"); return 0; }
    if (!L->is_ll || !L->is_super) { RAISE(Common, SSB_CHOLMOD_INVALID, "L not supernodal"); return 0; }
    Common->status = SSB_CHOLMOD_OK;
    if (L->n == 0 || X->ncol == 0) return 1;
    const bool cplx = L->xtype == SSB_CHOLMOD_COMPLEX;  // X is complex too (checked above): interleaved (re,im) = the real vector of order 2n
    std::lock_guard<std::mutex> lk(g_cache_mu);
    {
        static int verbose = -1;
        if (verbose < 0) { const char *v = getenv("SSB200_VERBOSE"); verbose = (v && atoi(v)) ? 1 : 0; }
        if (verbose) fprintf(stderr, "[suitesparse_b200] cholmod_l_super_%ssolve: n=%zu nrhs=%zu (CUDA path)\n", which ? "lt" : "l", L->n, X->ncol);
    }
    CacheEntry *e = cache_get_plan(L, cplx, cplx);
    if (!e) { RAISE(Common, SSB_CHOLMOD_GPU_PROBLEM, ssb200_last_error()); return 0; }
    if (cplx) {
        if (!value_fingerprint_ok(e, L)) {
            if (upload_complex_L(e, L)) { RAISE(Common, SSB_CHOLMOD_GPU_PROBLEM, ssb200_last_error()); return 0; }
            take_value_fingerprint(e, L);
        }
        const int rc = ssb200_solve(e->plan, which, (double *) X->x, (ssb_long) X->ncol, (ssb_long) (2 * X->d));
        if (rc) { RAISE(Common, rc == SSB_CHOLMOD_INVALID ? SSB_CHOLMOD_INVALID : SSB_CHOLMOD_GPU_PROBLEM, ssb200_last_error()); return 0; }
        return 1;
    }
    if (!value_fingerprint_ok(e, L)) {
        // L->x was produced or modified elsewhere: bring it to the device (still the GPU path, just slower)
        if (e->mg ? ssb200_mg_upload_L(e->mg, (const double *) L->x) : ssb200_upload_L(e->plan, (const double *) L->x)) { RAISE(Common, SSB_CHOLMOD_GPU_PROBLEM, ssb200_last_error()); return 0; }
        take_value_fingerprint(e, L);
    }
    const int rc = e->mg ? ssb200_mg_solve(e->mg, which, (double *) X->x, (ssb_long) X->ncol, (ssb_long) X->d)
                         : ssb200_solve(e->plan, which, (double *) X->x, (ssb_long) X->ncol, (ssb_long) X->d);
    if (rc) { RAISE(Common, rc == SSB_CHOLMOD_INVALID ? SSB_CHOLMOD_INVALID : SSB_CHOLMOD_GPU_PROBLEM, ssb200_last_error()); return 0; }
    return 1;
}

extern "C" int cholmod_l_super_lsolve(ssb_cholmod_factor *L, ssb_cholmod_dense *X, ssb_cholmod_dense *E, ssb_cholmod_common *Common)
{
    return super_solve_common(L, X, E, Common, 0);
}
extern "C" int cholmod_l_super_ltsolve(ssb_cholmod_factor *L, ssb_cholmod_dense *X, ssb_cholmod_dense *E, ssb_cholmod_common *Common)
{
    return super_solve_common(L, X, E, Common, 1);
}

// ---- cholmod_l_solve (Cholesky/cholmod_solve.c:1020-1060 -> cholmod_solve2 :1097-1680) interposed for its main case: Ax = b with
// a real supernodal LL' factor that lives on the device and a real dense B.  The reference permutes on the host (Y = P B,
// :1552-1556), calls super_lsolve / super_ltsolve (two host<->device round trips of Y through our symbols) and permutes back
// (:1578-1580).  Here B goes to the device once, P, L, L', P' are applied there, X comes back once.  Everything else (other
// systems, simplicial or complex factors, sparse right-hand sides, the multi-GPU plan) goes to the host library's own
// cholmod_l_solve, which still reaches the GPU through the interposed lsolve / ltsolve.
typedef ssb_cholmod_dense *(*solve_fn)(int, ssb_cholmod_factor *, ssb_cholmod_dense *, ssb_cholmod_common *);
typedef ssb_cholmod_dense *(*alloc_dense_fn)(size_t, size_t, size_t, int, ssb_cholmod_common *);
extern "C" ssb_cholmod_dense *cholmod_l_solve(int sys, ssb_cholmod_factor *L, ssb_cholmod_dense *B, ssb_cholmod_common *Common)
{
    static solve_fn next = nullptr;
    static alloc_dense_fn alloc_dense = (alloc_dense_fn) host_sym("cholmod_l_allocate_dense");
    static int fast_on = -1;
    if (fast_on < 0) { const char *v = getenv("SSB200_SOLVE_FAST"); fast_on = (v && atoi(v) == 0) ? 0 : 1; }
    auto fallback = [&]() -> ssb_cholmod_dense * {
        if (!next) next = (solve_fn) next_definition("cholmod_l_solve", (void *) &cholmod_l_solve);
        if (!next) { if (Common) Common->status = SSB_CHOLMOD_INVALID; return nullptr; }
        return next(sys, L, B, Common);
    };
    const bool fast = fast_on && alloc_dense && Common && L && B && sys == 0 /* CHOLMOD_A */ && Common->itype == SSB_CHOLMOD_LONG && Common->dtype == SSB_CHOLMOD_DOUBLE &&
                      L->is_super && L->is_ll && L->xtype == SSB_CHOLMOD_REAL && L->x && L->itype == SSB_CHOLMOD_LONG && L->minor == L->n &&
                      B->xtype == SSB_CHOLMOD_REAL && B->x && B->nrow == L->n && B->d >= B->nrow && L->n > 0 && B->ncol > 0 && gpu_enabled_by_env();
    if (!fast) return fallback();
    ssb_cholmod_dense *X = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        CacheEntry *e = cache_find(L);
        if (!e || e->mg || e->cx || !e->plan || !value_fingerprint_ok(e, L)) e = nullptr;       // only a factor this library holds on one device
        if (e) {
            ssb200_plan *p = e->plan;
            const long long n = (long long) L->n; const int nrhs = (int) B->ncol;
            bool ok = cudaSetDevice(p->device) == cudaSuccess;
            if (ok && !e->d_perm && L->Perm) {
                ok = cudaMalloc((void **) &e->d_perm, n * sizeof(long long)) == cudaSuccess &&
                     cudaMemcpyAsync(e->d_perm, L->Perm, n * sizeof(long long), cudaMemcpyHostToDevice, p->stream) == cudaSuccess;
            }
            const size_t need = (size_t) n * nrhs;
            if (ok && e->capB < need) { if (e->d_B) cudaFree(e->d_B); e->d_B = nullptr; e->capB = 0; ok = cudaMalloc((void **) &e->d_B, need * sizeof(double)) == cudaSuccess; if (ok) e->capB = need; }
            if (ok) ok = ensure_cap(p, &p->d_X, &p->capX, need) == 0;
            if (ok) X = alloc_dense((size_t) n, (size_t) nrhs, (size_t) n, SSB_CHOLMOD_REAL, Common);
            ok = ok && X != nullptr;
            if (ok) {
                const unsigned g = (unsigned) ((n + 255) / 256);
                ok = cudaMemcpy2DAsync(e->d_B, n * sizeof(double), B->x, B->d * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyHostToDevice, p->stream) == cudaSuccess;
                if (ok) perm_gather_kernel<<<g, 256, 0, p->stream>>>(e->d_perm, e->d_B, n, p->d_X, n, nrhs);
                ok = ok && ssb200_solve_resident(p, 2, p->d_X, nrhs, n) == 0;
                if (ok) perm_scatter_kernel<<<g, 256, 0, p->stream>>>(e->d_perm, p->d_X, e->d_B, n, n, nrhs);
                ok = ok && cudaMemcpyAsync(X->x, e->d_B, need * sizeof(double), cudaMemcpyDeviceToHost, p->stream) == cudaSuccess &&
                     cudaStreamSynchronize(p->stream) == cudaSuccess && cudaGetLastError() == cudaSuccess;
                if (ok) { Common->status = SSB_CHOLMOD_OK; return X; }
            }
            (void) cudaGetLastError();
        }
    }
    if (X) {            // the device path failed half way: give the workspace back and let the host library do it
        typedef int (*free_dense_fn)(ssb_cholmod_dense **, ssb_cholmod_common *);
        static free_dense_fn free_dense = (free_dense_fn) host_sym("cholmod_l_free_dense");
        if (free_dense) free_dense(&X, Common);
    }
    return fallback();
}

// ---- cholmod_l_gpu_* (GPU/cholmod_gpu.c:71,170,208,255,364): resource queries and teardown -------------------------
extern "C" int cholmod_l_gpu_memorysize(size_t *total_mem, size_t *available_mem, ssb_cholmod_common *Common)
{
    if (total_mem) *total_mem = 0;
    if (available_mem) *available_mem = 0;
    if (!Common) return 1;
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { Common->status = SSB_CHOLMOD_GPU_PROBLEM; return 1; }   // nonzero = problem, as in the reference
    if (total_mem) *total_mem = tot;
    if (available_mem) *available_mem = fr;
    Common->gpuMemorySize = fr;
    return 0;
}
extern "C" int cholmod_l_gpu_probe(ssb_cholmod_common *Common)
{
    (void) Common;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return 0;
    int dev = 0;                    // the device the drop-in layer will use: SSB200_DEVICE, else the current one
    if (const char *d = getenv("SSB200_DEVICE")) dev = atoi(d); else if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    if (dev < 0 || dev >= ndev) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    return prop.major == 10;        // the kernels are sm_100a only
}
extern "C" int cholmod_l_gpu_allocate(ssb_cholmod_common *Common) { (void) Common; return 0; }      // plans own their memory; nothing to pre-allocate
extern "C" int cholmod_l_gpu_deallocate(ssb_cholmod_common *Common)
{
    (void) Common;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    while (!g_cache.empty()) cache_drop(&g_cache.back());
    return 0;
}
extern "C" void cholmod_l_gpu_end(ssb_cholmod_common *Common) { cholmod_l_gpu_deallocate(Common); }

// The device copy of a factor is trusted by the solves when L->x still is the buffer the last factorization wrote and a
// sampled fingerprint matches.  A caller that edits L->x in place tells the library so with this call (the next solve
// uploads the host values again); freeing the factor through cholmod_l_free_factor below drops everything.
extern "C" int ssb200_invalidate_factor(const ssb_cholmod_factor *L)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    CacheEntry *e = cache_find(L);
    if (!e) return 0;
    if (e->plan) { e->plan->factor_on_device = false; e->plan->winv_valid = false; }
    if (e->mg) e->mg->factor_on_devices = false;
    e->xptr = nullptr;
    return 1;
}

// cholmod_l_free_factor (Core/cholmod_factor.c:152) interposed: the plan cache is keyed by the factor's address and holds
// xsize doubles of HBM plus a page-lock on L->x, so the entry must die with the factor.  Then the host library's own
// definition (next in the symbol search order) does the real work.
typedef int (*free_factor_fn)(ssb_cholmod_factor **, ssb_cholmod_common *);
extern "C" int cholmod_l_free_factor(ssb_cholmod_factor **LHandle, ssb_cholmod_common *Common)
{
    if (LHandle && *LHandle) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        if (CacheEntry *e = cache_find(*LHandle)) cache_drop(e);
    }
    static free_factor_fn next = nullptr;
    if (!next) next = (free_factor_fn) next_definition("cholmod_l_free_factor", (void *) &cholmod_l_free_factor);
    if (!next) { if (Common) Common->status = SSB_CHOLMOD_INVALID; return 0; }
    return next(LHandle, Common);
}

// plan of a cached factor (tests / bench: statistics of the drop-in path)
extern "C" ssb200_plan *ssb200_plan_of_factor(const ssb_cholmod_factor *L)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    CacheEntry *e = cache_find(L);
    return e ? e->plan : nullptr;
}
extern "C" ssb200_mg *ssb200_mg_of_factor(const ssb_cholmod_factor *L)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    CacheEntry *e = cache_find(L);
    return e ? e->mg : nullptr;
}
