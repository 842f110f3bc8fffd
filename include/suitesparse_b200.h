/* suitesparse_b200.h — C ABI of the B200-native supernodal Cholesky hot path.
 *
 * Two layers, both `extern "C"`, plain pointers and sizes only:
 *
 *  (1) DROP-IN layer: the symbols CHOLMOD's own driver binds for this path.  They keep the
 *      reference's names, argument meaning, return values and Common->status protocol, so that
 *      this library placed ahead of libcholmod in link order (or LD_PRELOAD / RTLD_GLOBAL) takes
 *      over the calls made by cholmod_l_factorize_p (CHOLMOD/Cholesky/cholmod_factorize.c:265)
 *      and cholmod_l_solve2 (CHOLMOD/Cholesky/cholmod_solve.c:1568-1577):
 *          cholmod_l_super_numeric   replaces CHOLMOD/Supernodal/cholmod_super_numeric.c:97
 *                                    (+ t_cholmod_super_numeric.c:93, GPU/t_cholmod_gpu.c)
 *          cholmod_l_super_lsolve    replaces CHOLMOD/Supernodal/cholmod_super_solve.c:43
 *          cholmod_l_super_ltsolve   replaces CHOLMOD/Supernodal/cholmod_super_solve.c:136
 *          cholmod_l_gpu_*           replaces CHOLMOD/GPU/cholmod_gpu.c:71,170,208,255,364
 *      The object layouts below restate CHOLMOD 3.0.14's public structs
 *      (CHOLMOD/Include/cholmod_core.h:416-1026, 1243-1270, 1673-1798, 1976-1990) field for field;
 *      tests/test_abi.py checks sizeof/offsetof against the reference headers.
 *
 *  (2) PLAIN layer (ssb200_*): the same computation on raw arrays — the supernodal symbolic
 *      structure (super/pi/px/s of cholmod_factor, cholmod_core.h:1723-1732), a CSC matrix and
 *      the dense block array Lx.  This is what another language's FFI (ctypes / cgo / JNI) binds.
 *
 * Only the `long` (64-bit index, `cholmod_l_`) real double-precision case is accelerated: the
 * reference enables its GPU path for exactly that case (CHOLMOD/Include/cholmod_internal.h:249-252).
 * There is no CPU fallback: if no sm_100 device is usable the calls fail with CHOLMOD_GPU_PROBLEM.
 */
#ifndef SUITESPARSE_B200_H
#define SUITESPARSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t ssb_long;           /* SuiteSparse_long on LP64 (SuiteSparse_config.h) */

/* status codes, cholmod_core.h:386-394 */
#define SSB_CHOLMOD_OK              0
#define SSB_CHOLMOD_NOT_INSTALLED (-1)
#define SSB_CHOLMOD_OUT_OF_MEMORY (-2)
#define SSB_CHOLMOD_TOO_LARGE     (-3)
#define SSB_CHOLMOD_INVALID       (-4)
#define SSB_CHOLMOD_GPU_PROBLEM   (-5)
#define SSB_CHOLMOD_NOT_POSDEF      1
#define SSB_CHOLMOD_DSMALL          2
/* xtype / itype / dtype, cholmod_core.h:310-333 */
#define SSB_CHOLMOD_PATTERN 0
#define SSB_CHOLMOD_REAL    1
#define SSB_CHOLMOD_COMPLEX 2
#define SSB_CHOLMOD_ZOMPLEX 3
#define SSB_CHOLMOD_LONG    2
#define SSB_CHOLMOD_DOUBLE  0
#define SSB_CHOLMOD_MAXMETHODS 9
#define SSB_CHOLMOD_HOST_SUPERNODE_BUFFERS 8

#ifndef SSB200_NO_CHOLMOD_TYPES
/* ---- cholmod_sparse, cholmod_core.h:1243 ------------------------------------------------- */
typedef struct ssb_cholmod_sparse {
    size_t nrow, ncol, nzmax;
    void *p, *i, *nz, *x, *z;
    int stype, itype, xtype, dtype, sorted, packed;
} ssb_cholmod_sparse;

/* ---- cholmod_dense, cholmod_core.h:1976 -------------------------------------------------- */
typedef struct ssb_cholmod_dense {
    size_t nrow, ncol, nzmax, d;
    void *x, *z;
    int xtype, dtype;
} ssb_cholmod_dense;

/* ---- cholmod_factor, cholmod_core.h:1673 ------------------------------------------------- */
typedef struct ssb_cholmod_factor {
    size_t n, minor;
    void *Perm, *ColCount, *IPerm;
    size_t nzmax;                                   /* simplicial part */
    void *p, *i, *x, *z, *nz, *next, *prev;
    size_t nsuper, ssize, xsize, maxcsize, maxesize;/* supernodal part */
    void *super, *pi, *px, *s;
    int ordering, is_ll, is_super, is_monotonic, itype, xtype, dtype, useGPU;
} ssb_cholmod_factor;

/* ---- cholmod_common, cholmod_core.h:416 -------------------------------------------------- */
typedef struct ssb_cholmod_method {
    double lnz, fl, prune_dense, prune_dense2, nd_oksep, other_1[4];
    size_t nd_small, other_2[4];
    int aggressive, order_for_lu, nd_compress, nd_camd, nd_components, ordering;
    size_t other_3[4];
} ssb_cholmod_method;

typedef struct ssb_cholmod_common {
    double dbound, grow0, grow1;
    size_t grow2, maxrank;
    double supernodal_switch;
    int supernodal, final_asis, final_super, final_ll, final_pack, final_monotonic, final_resymbol;
    double zrelax[3];
    size_t nrelax[3];
    int prefer_zomplex, prefer_upper, quick_return_if_not_posdef, prefer_binary, print, precise, try_catch;
    void (*error_handler)(int status, const char *file, int line, const char *message);
    int nmethods, current, selected;
    ssb_cholmod_method method[SSB_CHOLMOD_MAXMETHODS + 1];
    int postorder, default_nesdis;
    double metis_memory, metis_dswitch;
    size_t metis_nswitch;
    size_t nrow;
    ssb_long mark;
    size_t iworksize, xworksize;
    void *Flag, *Head, *Xwork, *Iwork;
    int itype, dtype, no_workspace_reallocate, status;
    double fl, lnz, anz, modfl;
    size_t malloc_count, memory_usage, memory_inuse;
    double nrealloc_col, nrealloc_factor, ndbounds_hit, rowfacfl, aatfl;
    int called_nd, blas_ok;
    double SPQR_grain, SPQR_small;
    int SPQR_shrink, SPQR_nthreads;
    double SPQR_flopcount, SPQR_analyze_time, SPQR_factorize_time, SPQR_solve_time,
           SPQR_flopcount_bound, SPQR_tol_used, SPQR_norm_E_fro;
    ssb_long SPQR_istat[10];
    /* GPU configuration and statistics (cholmod_core.h:954-1024); handles are void* in a non-GPU_BLAS build */
    int useGPU;
    size_t maxGpuMemBytes;
    double maxGpuMemFraction;
    size_t gpuMemorySize;
    double gpuKernelTime;
    ssb_long gpuFlops;
    int gpuNumKernelLaunches;
    void *cublasHandle;
    void *gpuStream[SSB_CHOLMOD_HOST_SUPERNODE_BUFFERS];
    void *cublasEventPotrf[3];
    void *updateCKernelsComplete;
    void *updateCBuffersFree[SSB_CHOLMOD_HOST_SUPERNODE_BUFFERS];
    void *dev_mempool;
    size_t dev_mempool_size;
    void *host_pinned_mempool;
    size_t host_pinned_mempool_size;
    size_t devBuffSize;
    int ibuffer;
    double syrkStart;
    double cpu_gemm_time, cpu_syrk_time, cpu_trsm_time, cpu_potrf_time;
    double gpu_gemm_time, gpu_syrk_time, gpu_trsm_time, gpu_potrf_time;
    double assemble_time, assemble_time2;
    size_t cpu_gemm_calls, cpu_syrk_calls, cpu_trsm_calls, cpu_potrf_calls;
    size_t gpu_gemm_calls, gpu_syrk_calls, gpu_trsm_calls, gpu_potrf_calls;
} ssb_cholmod_common;

/* =========================== (1) DROP-IN layer ============================================== */
/* Prototypes are those of CHOLMOD/Include/cholmod_supernodal.h:112-127,137-152,162-177 and
 * CHOLMOD/Include/cholmod_gpu.h:59-93 (struct tags differ in name only). */

#ifndef SSB200_NO_DROPIN_PROTOTYPES   /* define when cholmod.h is included in the same unit */
/* Numeric supernodal LL' of A (stype<0: lower triangle of the permuted symmetric matrix) or of
 * A*F (stype==0, F=A').  Returns 1 (TRUE) on success AND when the matrix is not positive definite
 * (Common->status = CHOLMOD_NOT_POSDEF, L->minor = failing column, t_cholmod_super_numeric.c:905-968);
 * returns 0 with Common->status<0 on invalid input / out of memory / GPU failure. */
int cholmod_l_super_numeric(ssb_cholmod_sparse *A, ssb_cholmod_sparse *F, double beta[2],
                            ssb_cholmod_factor *L, ssb_cholmod_common *Common);
/* X <- L \ X (forward) and X <- L' \ X (backward), X is n-by-nrhs with leading dimension X->d;
 * E is the caller's workspace of >= nrhs*L->maxesize entries (checked, unused on the device). */
int cholmod_l_super_lsolve (ssb_cholmod_factor *L, ssb_cholmod_dense *X, ssb_cholmod_dense *E,
                            ssb_cholmod_common *Common);
int cholmod_l_super_ltsolve(ssb_cholmod_factor *L, ssb_cholmod_dense *X, ssb_cholmod_dense *E,
                            ssb_cholmod_common *Common);

int  cholmod_l_gpu_memorysize(size_t *total_mem, size_t *available_mem, ssb_cholmod_common *Common);
int  cholmod_l_gpu_probe     (ssb_cholmod_common *Common);
int  cholmod_l_gpu_allocate  (ssb_cholmod_common *Common);
int  cholmod_l_gpu_deallocate(ssb_cholmod_common *Common);
void cholmod_l_gpu_end       (ssb_cholmod_common *Common);
/* Cholesky/cholmod_solve.c:1020 interposed for its main case (sys = CHOLMOD_A, real supernodal LL' factor held on one device,
 * real dense B): B goes to the device once, P, L, L', P' are applied there, X (allocated by the host library) comes back once.
 * Every other case is handed to the host library's own cholmod_l_solve. */
ssb_cholmod_dense *cholmod_l_solve(int sys, ssb_cholmod_factor *L, ssb_cholmod_dense *B, ssb_cholmod_common *Common);
/* Core/cholmod_factor.c:152 interposed: drops the cached device plan and the page-lock of L->x, then calls the host
 * library's own cholmod_l_free_factor (next definition in the symbol search order). */
int  cholmod_l_free_factor   (ssb_cholmod_factor **L, ssb_cholmod_common *Common);
/* The drop-in entry points fan one factorization out over the devices listed in $SSB200_DEVICES ("0,1,2,3" or "all";
 * two or more -> ssb200_mg_*); this returns the multi-GPU plan cached for L, or NULL. */
struct ssb200_mg *ssb200_mg_of_factor(const ssb_cholmod_factor *L);
/* When the drop-in layer page-locks L->x (same values as the SSB200_PIN_HOST environment variable): 0 never - every
   factorization leaves through the pinned staging ring; 1 (default) once the same L->x has been factorized into
   SSB200_PIN_AFTER (32) times; 2 at the first call; -1 back to the environment's setting.  Returns the previous policy. */
int ssb200_set_pin_policy(int policy);
/* Test hook, host only: copy src to dst with the staging ring's mover threads while the first-touch threads fault dst in. */
int ssb200_debug_first_touch_copy(void *dst, const void *src, size_t bytes, int mover_threads, size_t slot_bytes);
/* Complex and zomplex matrices (cholmod_super_numeric.c:81-86) go through the real kernels: every entry a+ib becomes the block
 * [a -b; b a]; the Cholesky factor of that real SPD matrix of order 2n is the blockified complex factor, which is written to
 * L->x in CHOLMOD's complex layout.  Test hook: the blockified copy of A (lower != 0: symmetric-lower input). */
ssb_long ssb200_debug_blockify(const ssb_cholmod_sparse *A, int lower, ssb_long *p2, ssb_long *i2, double *x2, ssb_long cap);
/* the caller changed L->x in place: the next solve uploads the host values again.  Returns 1 if L had a cached plan. */
int  ssb200_invalidate_factor(const ssb_cholmod_factor *L);
#endif /* SSB200_NO_DROPIN_PROTOTYPES */
#endif /* SSB200_NO_CHOLMOD_TYPES */

/* =========================== (2) PLAIN layer ================================================ */
typedef struct ssb200_plan ssb200_plan;       /* opaque: device-resident symbolic plan + factor storage */

/* Build the device plan from the symbolic supernodal structure (host arrays, read once).
 *   n, nsuper           cholmod_factor.n / .nsuper
 *   super[nsuper+1]     first column of each supernode         (cholmod_factor.super)
 *   pi[nsuper+1]        offsets into s                          (cholmod_factor.pi)
 *   px[nsuper+1]        offsets into Lx; px[nsuper] = xsize     (cholmod_factor.px)
 *   s[pi[nsuper]]       row indices, first nscol of a supernode are its own columns (cholmod_factor.s)
 *   device              CUDA device ordinal; -1 = current device
 * Returns NULL on failure (ssb200_last_error() says why).  The descendant-update lists that the
 * reference discovers with Head/Next/Lpos at run time (t_cholmod_super_numeric.c:442-460,583-611,
 * 787-808) are a pure function of this structure and are precomputed here. */
ssb200_plan *ssb200_plan_create(ssb_long n, ssb_long nsuper, const ssb_long *super, const ssb_long *pi,
                                const ssb_long *px, const ssb_long *s, int device);
void ssb200_plan_destroy(ssb200_plan *plan);

/* Elimination-tree shard over nranks GPUs, one process per GPU.  The plan builder cuts the supernodal
 * etree into independent subtrees (LPT-balanced by dense flops) owned by one rank each; the wide supernodes above the cut
 * are shared panel-cyclically (256-column panels).  Every rank keeps the whole Lx array; a finished subtree / panel is
 * replicated by ONE broadcast of a contiguous Lx range, issued by the caller (torch.distributed / NCCL) between steps:
 *     ssb200_upload_A(); ssb200_dist_begin();
 *     for k < ssb200_dist_num_steps():  ssb200_dist_step_info(k,&src,&off,&cnt,&wait);
 *         if (wait) wait for every broadcast started so far;   ssb200_dist_run_step(k, 0);
 *         if (src >= 0) start broadcast(Lx+off, cnt, src) asynchronously, ordered after the launches above;
 *         ssb200_dist_run_step(k, 1);          // look-ahead work that overlaps the broadcast
 *     ssb200_dist_end(&bad);  minor = min over ranks of bad;
 *     if (minor < n) { ssb200_dist_not_posdef(minor, quick, &redo_rank, &off, &cnt); if (redo_rank >= 0) broadcast(Lx+off, cnt, redo_rank); }
 * Kernels go to the stream given by ssb200_set_stream(plan, cudaStream_t); the caller orders its broadcasts against it. */
ssb200_plan *ssb200_plan_create_dist(ssb_long n, ssb_long nsuper, const ssb_long *super, const ssb_long *pi,
                                     const ssb_long *px, const ssb_long *s, int device, int nranks, int rank);
int      ssb200_set_stream(ssb200_plan *plan, void *cuda_stream);      /* NULL: back to the plan's own stream */
ssb_long ssb200_dist_num_steps(const ssb200_plan *plan);
int      ssb200_dist_step_info(const ssb200_plan *plan, ssb_long k, int *src, ssb_long *off, ssb_long *cnt, int *wait_remote);
int      ssb200_dist_begin(ssb200_plan *plan, const double beta[2]);
int      ssb200_dist_run_step(ssb200_plan *plan, ssb_long k, int part);
int      ssb200_dist_end(ssb200_plan *plan, ssb_long *first_bad_column);
int      ssb200_dist_zero_from(ssb200_plan *plan, ssb_long column);
int      ssb200_dist_not_posdef(ssb200_plan *plan, ssb_long minor, int quick_return, int *redo_rank, ssb_long *off, ssb_long *cnt);
int      ssb200_dist_flops(const ssb200_plan *plan, double *mine, double *total);

/* ---- Multi-GPU inside ONE process (the reference "can only utilize a single GPU", GPU/cholmod_gpu.c:160-164).
 * The elimination tree is cut into subtrees owned by one device each, the wide top supernodes are shared panel-cyclically;
 * the factor is DISTRIBUTED: a device stores its own supernodes, the cyclic ones, and the remote supernodes its updates read,
 * which it pulls over NVLink (peer-mapped pointers, copy kernel) as soon as their owner has finished them.  One host thread
 * per device, CUDA events across devices, no NCCL.  devices == NULL: ordinals 0..ndev-1.  Lx_host (optional, xsize doubles,
 * CHOLMOD's layout) receives every device's share over that device's own PCIe link while the factorization runs
 * (page-lock it once with ssb200_mg_pin_host).  A matrix that is not positive definite returns 1 with *minor_out set; the
 * partial refactorization of the failing supernode (t_cholmod_super_numeric.c:944-967) is then left to the single-GPU path. */
/* Ap == NULL in ssb200_mg_factorize: factorize the matrix uploaded by the previous call again (resident refactorization). */
typedef struct ssb200_mg ssb200_mg;
ssb200_mg *ssb200_mg_create(ssb_long n, ssb_long nsuper, const ssb_long *super, const ssb_long *pi, const ssb_long *px,
                            const ssb_long *s, int ndev, const int *devices);
void ssb200_mg_destroy(ssb200_mg *mg);
int  ssb200_mg_pin_host(ssb200_mg *mg, double *Lx_host);
int  ssb200_mg_factorize(ssb200_mg *mg, int stype,
                         const ssb_long *Ap, const ssb_long *Ai, const ssb_long *Anz, const double *Ax, ssb_long ncolA,
                         const ssb_long *Fp, const ssb_long *Fi, const ssb_long *Fnz, const double *Fx,
                         const double beta[2], double *Lx_host, ssb_long *minor_out);
/* X (host, n-by-nrhs, leading dimension ldx) <- L\X, L'\X or both, on the distributed factor: right-hand sides on device
 * 0, every device solves with its own blocks and reaches X over NVLink. */
int  ssb200_mg_solve(ssb200_mg *mg, int which, double *X, ssb_long nrhs, ssb_long ldx);
/* host factor (CHOLMOD layout) -> the devices' local storage (a factor computed elsewhere; the solves then work without
 * the inverses of the diagonal blocks) */
int  ssb200_mg_upload_L(ssb200_mg *mg, const double *Lx_host);
/* out[0] wall ms of the last factorization, out[1] of the last solve, out[2] NVLink bytes pulled per factorization,
 * out[3+r] HBM bytes held by rank r, out[3+N+r] dense flops of rank r, out[3+2N] device ms of the last factorization (CUDA
 * events, max over the devices); returns the number of devices. */
int  ssb200_mg_info(const ssb200_mg *mg, double *out, int cap);
ssb_long ssb200_mg_launches(const ssb200_mg *mg);      /* kernels launched by the last call, all devices */
/* tuning aid ($SSB200_MG_TRACE=1): device time at which every rank reached every step of the last factorization */
ssb_long ssb200_mg_trace(const ssb200_mg *mg, float *out, ssb_long cap, ssb_long *steps, ssb_long steps_cap);

/* Numeric factorization.  A (and F for stype==0) are HOST CSC arrays with 64-bit indices:
 * Ap[n+1], Ai, Ax and optional Anz (unpacked).  stype<0 symmetric-lower input, stype==0 A*F.
 * Lx_host (xsize doubles) receives the factor if not NULL.  *minor_out = n on success, else the
 * failing column.  Returns 0 ok, 1 not positive definite, <0 error (SSB_CHOLMOD_* codes). */
int ssb200_factorize(ssb200_plan *plan, int stype,
                     const ssb_long *Ap, const ssb_long *Ai, const ssb_long *Anz, const double *Ax, ssb_long ncolA,
                     const ssb_long *Fp, const ssb_long *Fi, const ssb_long *Fnz, const double *Fx,
                     const double beta[2], int quick_return_if_not_posdef,
                     double *Lx_host, ssb_long *minor_out);

/* Device-resident variants: upload A once, factorize on the device only (no host traffic), read back. */
int ssb200_upload_A(ssb200_plan *plan, int stype,
                    const ssb_long *Ap, const ssb_long *Ai, const ssb_long *Anz, const double *Ax, ssb_long ncolA,
                    const ssb_long *Fp, const ssb_long *Fi, const ssb_long *Fnz, const double *Fx);
int ssb200_factorize_resident(ssb200_plan *plan, const double beta[2], int quick_return_if_not_posdef,
                              ssb_long *minor_out);
int ssb200_download_L(ssb200_plan *plan, double *Lx_host);
int ssb200_upload_L(ssb200_plan *plan, const double *Lx_host);

/* Triangular solves with the device-resident factor.  X is a HOST n-by-nrhs column-major array,
 * leading dimension ldx.  which: 0 = L\X, 1 = L'\X, 2 = both (L then L'). */
int ssb200_solve(ssb200_plan *plan, int which, double *X, ssb_long nrhs, ssb_long ldx);
/* same with X already on the device (device pointer) */
int ssb200_solve_resident(ssb200_plan *plan, int which, double *dX, ssb_long nrhs, ssb_long ldx);

/* Look-ahead schedule (default on): the latency-bound potrf/trsm chain of the next 1024-column outer panel runs on a
 * high-priority stream while the main stream applies the previous outer panel to the rest of the supernode.  With it on
 * only total device times are reported; off = one stream, timing events around every launch (ssb200_stats.ms_kind). */
int ssb200_set_lookahead(ssb200_plan *plan, int on);

/* diag_host[k] = L(k,k), k < n, read from the device-resident factor (what Cholesky/cholmod_rcond.c:102-125 scans). */
int ssb200_factor_diag(ssb200_plan *plan, double *diag_host);

/* Raw device pointers / sizes for callers that manage their own streams or collectives. */
double  *ssb200_device_Lx(ssb200_plan *plan);
ssb_long ssb200_xsize(const ssb200_plan *plan);
void    *ssb200_stream(ssb200_plan *plan);            /* cudaStream_t the plan launches on */

/* Statistics of the last factorize / solve on this plan. */
typedef struct ssb200_stats {
    ssb_long nsuper, nlevels, nupdates;       /* symbolic sizes: supernodes, etree levels, (d,s) update pairs */
    ssb_long kernel_launches;                 /* kernels launched by the last call */
    ssb_long kernel_launches_total;           /* since plan creation */
    double   flops_update, flops_potrf, flops_trsm; /* dense flops executed (incl. amalgamation zeros) */
    double   ms_total, ms_assemble, ms_update, ms_factor, ms_d2h, ms_h2d; /* device time (CUDA events), last call */
    double   bytes_update_panel, bytes_update_scatter;  /* algorithmic bytes of the update kernel */
    ssb_long device_bytes;                    /* HBM held by the plan */
    /* per kernel of the last factorize: [0] gemm_nt_sub<128>, [1] gemm_nt_sub<64>, [2] potrf_block, [3] trsm_rows,
       [4] trsm_tc, [5] unused */
    double   ms_kind[6];                      /* device time (CUDA events on the plan's stream) */
    double   flops_kind[6];                   /* algorithmic flops executed by that kernel */
    ssb_long launches_kind[6];
    ssb_long d2h_staged;                      /* last factorize: 1 = the factor went to pageable host memory through the pinned staging ring */
} ssb200_stats;
int ssb200_get_stats(const ssb200_plan *plan, ssb200_stats *out);

const char *ssb200_last_error(void);
const char *ssb200_version(void);
int ssb200_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
