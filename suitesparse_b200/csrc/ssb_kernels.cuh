// ssb_kernels.cuh — hand-written sm_100a kernels of the supernodal Cholesky hot path.
//
//   gemm_nt_sub_kernel   C -= A*B^T on fp64 tensor cores (mma.sync m8n8k4 = DMMA.8x8x4; tcgen05 has no f64 kind),
//                        operands staged by a 3-stage cp.async ring, extend-add scatter through the relative map fused
//                        into the epilogue.  Replaces dsyrk+dgemm+assembly (t_cholmod_super_numeric.c:676-772) and the
//                        reference GPU path's cublasDsyrk/cublasDgemm/kernelAddUpdate (GPU/t_cholmod_gpu.c:514-620,
//                        GPU/cholmod_gpu_kernels.cu:39-50); also the trailing updates inside a supernode's own potrf/trsm.
//   potrf_block_kernel   unblocked Cholesky of one <=64x64 diagonal block in shared memory, LAPACK info contract
//                        (t_cholmod_super_numeric.c:864-867).
//   trsm_rows_kernel     rows-below substitution  B <- B * L11^{-T}  (dtrsm R,L,C,N, t_cholmod_super_numeric.c:997-1002).
//   scatter_A_kernel     supernode assembly of A or A*F (+beta) (t_cholmod_super_numeric.c:353-431)
//   relmap_kernel        RelativeMap of every update (t_cholmod_super_numeric.c:743-750, cholmod_gpu_kernels.cu:17-37)
//   lsolve/ltsolve       level-scheduled supernodal triangular solves (t_cholmod_super_solve.c:60-130, 268-332)
#pragma once
#include "ssb_internal.h"
#include <cuda_runtime.h>
#include <type_traits>

namespace ssb {

// ------------------------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc, int src_bytes)
{
    unsigned s = (unsigned) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// D(8x8) += A(8x4, row) * B(4x8, col), fp64 tensor core
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    // not volatile: the scheduler may interleave the staging code of the next chunk between the DMMAs
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void red_add_f64(double *addr, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(addr), "d"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// C -= A * B^T   (one CTA = one BT x BT tile of one job)
// ------------------------------------------------------------------------------------------------------------------
template <int BT> struct GemmCfg;
template <> struct GemmCfg<128> { static constexpr int WARPS_M = 2, WARPS_N = 4, STAGES = 3; };
template <> struct GemmCfg<64>  { static constexpr int WARPS_M = 2, WARPS_N = 2, STAGES = 3; };

template <int BT> constexpr int gemm_threads() { return GemmCfg<BT>::WARPS_M * GemmCfg<BT>::WARPS_N * 32; }
template <int BT, int KB> constexpr size_t gemm_smem_bytes()
{
    return (size_t) GemmCfg<BT>::STAGES * 2 * KB * (BT + 4) * sizeof(double) + 2 * BT * sizeof(int);
}

template <int BT, int KB>
__global__ void __launch_bounds__(GemmCfg<BT>::WARPS_M *GemmCfg<BT>::WARPS_N * 32)
gemm_nt_sub_kernel(const GemmJob *__restrict__ jobs, const int *__restrict__ tile_job, double *__restrict__ Lx,
                   const int *__restrict__ relmap)
{
    using Cfg = GemmCfg<BT>;
    constexpr int NT = Cfg::WARPS_M * Cfg::WARPS_N * 32;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int WM = BT / Cfg::WARPS_M, WN = BT / Cfg::WARPS_N;
    constexpr int MT = WM / 8, NTL = WN / 8;
    constexpr int LD = BT + 4;                      // LD % 16 == 4: conflict-free 8-byte fragment loads
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    int *rowmap = reinterpret_cast<int *>(smem + (size_t) STAGES * 2 * KB * LD);
    int *colmap = rowmap + BT;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const GemmJob job = jobs[tile_job[blockIdx.x]];
    // decode the tile.  The job's lower-trapezoid of tiles (tile column tj holds rows ti = tj .. nti-1) is enumerated in BANDS of
    // GB tile columns, row by row inside a band, so that the CTAs in flight at one time (148 consecutive indices) share GB
    // column operands and ~148/GB row operands - a working set of tens of MB that stays in L2 - instead of one column operand
    // and 148 different row operands (ncu, K = 1024 update of the root: 19.5 GB of DRAM reads for a 0.2 GB operand panel).
    constexpr int GB = 8;
    int rem = (int) blockIdx.x - job.tile_start, tj, ti;
    {
        int c0 = 0, wb;
        for (;;) {
            wb = min(GB, job.ntj - c0);
            const int cnt = wb * job.nti - (wb * c0 + wb * (wb - 1) / 2);          // tiles of the band [c0, c0+wb)
            if (rem < cnt || c0 + wb >= job.ntj) break;
            rem -= cnt; c0 += wb;
        }
        const int head = wb * (wb + 1) / 2;                                          // its first wb rows form a triangle
        if (rem < head) {
            int r = (int) ((sqrtf(8.0f * (float) rem + 1.0f) - 1.0f) * 0.5f);
            while (r > 0 && r * (r + 1) / 2 > rem) r--;
            while ((r + 1) * (r + 2) / 2 <= rem) r++;
            ti = c0 + r; tj = c0 + rem - r * (r + 1) / 2;
        } else {
            const int r2 = rem - head;
            ti = c0 + wb + r2 / wb; tj = c0 + r2 % wb;
        }
    }
    const int rowA0 = ti * BT, rowB0 = tj * BT;
    const bool diag = (ti == tj);
    const int K = job.K, nd1 = job.nd1, nd2 = job.nd2;
    const long long lda = job.lda;
    const double *__restrict__ P = Lx + job.a_off;

    if (job.map_off >= 0) {
        const int *rm = relmap + job.map_off;
        for (int t = tid; t < BT; t += NT) {
            rowmap[t] = (rowA0 + t < nd2) ? rm[rowA0 + t] : 0;
            colmap[t] = (rowB0 + t < nd1) ? rm[rowB0 + t] : 0;
        }
    }

    const int wm0 = (warp / Cfg::WARPS_N) * WM, wn0 = (warp % Cfg::WARPS_N) * WN;
    // on a diagonal tile, a warp tile entirely above the diagonal contributes nothing
    const bool warp_active = !(diag && (wm0 + WM <= wn0));

    double acc[MT][NTL][2];
#pragma unroll
    for (int a = 0; a < MT; a++)
#pragma unroll
        for (int b = 0; b < NTL; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    // Every thread stages the same (row, k mod KPT) pattern in every chunk: row lr = tid % BT of the A tile and of the
    // B tile, columns k = kq + KPT*i.  The source pointers simply advance by KB columns per chunk; only the last chunk
    // needs the k < K test.  The staging of chunk c+STAGES-1 is cut into KB/4 slices issued between the k-steps of
    // chunk c, so its address arithmetic fills the issue slots between DMMAs instead of stalling the tensor pipe.
    constexpr int KPT = NT / BT;                    // k-columns covered by one pass of the CTA
    constexpr int NLD = KB / KPT;                   // cp.async per operand per thread per chunk
    constexpr int KSTEPS = KB / 4;
    constexpr int LPS = (NLD + KSTEPS - 1) / KSTEPS; // cp.async per operand per thread per k-step slice
    const int lr = tid % BT, kq = tid / BT;
    const bool okA = rowA0 + lr < nd2, okB = (!diag) && (rowB0 + lr < nd1);
    const int szA = okA ? 8 : 0, szB = okB ? 8 : 0;
    const long long kstep = (long long) KPT * lda;
    const double *gA = P + (okA ? rowA0 + lr : 0) + (long long) kq * lda;     // advance as chunks are staged
    const double *gB = P + (okB ? rowB0 + lr : 0) + (long long) kq * lda;
    const int soff = kq * LD + lr;
    const int nk = (K + KB - 1) / KB;
    int ld_kcol = kq, ld_stage = 0;                 // first k-column this thread stages in the next chunk; its ring slot
    // slice `sl` (0..KSTEPS-1) of the staging of the next chunk.  Branch-free: columns at or beyond K (the tail of the
    // last chunk, or chunks past the end) are zero-filled with a zero-size copy from a valid dummy address.
    auto stage_slice = [&](int sl) {
        double *As = smem + (size_t) ld_stage * 2 * KB * LD + soff;
        double *Bs = As + KB * LD;
#pragma unroll
        for (int q = 0; q < LPS; q++) {
            const int i = sl * LPS + q;
            if (i < NLD) {                          // compile-time
                const bool kin = ld_kcol + KPT * i < K;
                cp_async8(As + i * KPT * LD, kin ? gA : P, kin ? szA : 0);
                cp_async8(Bs + i * KPT * LD, kin ? gB : P, kin ? szB : 0);
                gA += kstep; gB += kstep;
            }
        }
    };
    auto stage_advance = [&]() { ld_kcol += KB; ld_stage = (ld_stage + 1 == STAGES) ? 0 : ld_stage + 1; };

#pragma unroll
    for (int st = 0; st < STAGES - 1; st++) {
#pragma unroll
        for (int sl = 0; sl < KSTEPS; sl++) stage_slice(sl);
        stage_advance();
        cp_async_commit();
    }
    // The loop exists twice (warps that own output, and the idle warps of a diagonal tile that only stage and sync) so
    // that the k-step body is one straight-line block.
    auto mainloop = [&](auto ACTIVE) {
        constexpr bool act = decltype(ACTIVE)::value;
        int cs = 0;                                 // ring slot of the chunk being consumed
        for (int kc = 0; kc < nk; kc++) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();                        // chunk kc landed for everyone; slot (kc-1)%STAGES is free again
            const double *As = smem + (size_t) cs * 2 * KB * LD;
            const double *Bs = diag ? As : As + KB * LD;
            const double *ap = As + (lane & 3) * LD + wm0 + (lane >> 2);
            const double *bp = Bs + (lane & 3) * LD + wn0 + (lane >> 2);
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ks++) {
                if constexpr (act) {
                    double a[MT], b[NTL];
#pragma unroll
                    for (int m = 0; m < MT; m++) a[m] = ap[ks * 4 * LD + m * 8];
#pragma unroll
                    for (int nn = 0; nn < NTL; nn++) b[nn] = bp[ks * 4 * LD + nn * 8];
#pragma unroll
                    for (int m = 0; m < MT; m++)
#pragma unroll
                        for (int nn = 0; nn < NTL; nn++) dmma884(acc[m][nn][0], acc[m][nn][1], a[m], b[nn]);
                }
                stage_slice(ks);                    // part of chunk kc+STAGES-1 goes into the slot freed by chunk kc-1
            }
            stage_advance();
            cp_async_commit();
            cs = (cs + 1 == STAGES) ? 0 : cs + 1;
        }
    };
    if (warp_active) mainloop(std::true_type{}); else mainloop(std::false_type{});
    cp_async_wait<0>();
    __syncthreads();      // rowmap/colmap visible (also when nk == 0)

    if (!warp_active) return;
    // epilogue: extend-add.  acc[m][nn][e] = C(rowA0+wm0+m*8+lane/4, rowB0+wn0+nn*8+2*(lane%4)+e)
    double *__restrict__ Cb = Lx + job.c_off;
    const long long ldc = job.ldc;
    const bool mapped = job.map_off >= 0;
    const bool atomic = job.atomic != 0;
    // row offsets of this thread's MT rows (identity or relative map), -1 = outside the job
    int roff[MT];
#pragma unroll
    for (int m = 0; m < MT; m++) {
        const int il = wm0 + m * 8 + (lane >> 2);
        roff[m] = (rowA0 + il < nd2) ? (mapped ? rowmap[il] : rowA0 + il) : -1;
    }
#pragma unroll
    for (int nn = 0; nn < NTL; nn++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int jl = wn0 + nn * 8 + 2 * (lane & 3) + e;
            const int j = rowB0 + jl;
            if (j >= nd1) continue;
            double *__restrict__ colp = Cb + (mapped ? (long long) (colmap[jl] - job.c_col0) * ldc : (long long) j * ldc);
            if (atomic) {
#pragma unroll
                for (int m = 0; m < MT; m++) {
                    const int i = rowA0 + wm0 + m * 8 + (lane >> 2);
                    if (roff[m] >= 0 && i >= j) red_add_f64(colp + roff[m], -acc[m][nn][e]);
                }
            } else {
                // exclusive owner of the target entries: batch the loads so their latencies overlap
                double old[MT];
#pragma unroll
                for (int m = 0; m < MT; m++) {
                    const int i = rowA0 + wm0 + m * 8 + (lane >> 2);
                    old[m] = (roff[m] >= 0 && i >= j) ? __ldcg(colp + roff[m]) : 0.0;
                }
#pragma unroll
                for (int m = 0; m < MT; m++) {
                    const int i = rowA0 + wm0 + m * 8 + (lane >> 2);
                    if (roff[m] >= 0 && i >= j) __stcg(colp + roff[m], old[m] - acc[m][nn][e]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// diagonal block Cholesky: one CTA per job, block w x w (w <= 64) in shared memory.
// info[snode] = min(info, col0 + j + 1) at the first non-positive (or NaN) pivot (LAPACK dpotrf contract).
// ------------------------------------------------------------------------------------------------------------------
constexpr int POTRF_THREADS = 256;
// 256 threads = 16 x 16; thread (ti,tk) keeps the 4x4 strided sub-block T(ti+16a, tk+16b) in registers for the whole
// factorization.  Per column j: the owners of column j publish it to shared memory (double-buffered), one barrier,
// then every thread applies the rank-1 update to its registers.  L's column j goes straight to global memory.
__global__ void __launch_bounds__(POTRF_THREADS) potrf_block_kernel(const PanelJob *__restrict__ jobs, double *__restrict__ Lx,
                                                                   int *__restrict__ info, double *__restrict__ winv)
{
    __shared__ double colbuf[2][NB_INNER];
    __shared__ double rowbuf[2][NB_INNER];
    const PanelJob job = jobs[blockIdx.x];
    const int w = job.w, tid = threadIdx.x;
    const int ti = tid & 15, tk = tid >> 4;
    const long long lda = job.lda;
    const bool want_inv = job.winv_slot >= 0;       // also build W = L^{-1} (forward substitution on the identity)
    double *__restrict__ A = Lx + job.x_off;
    double t[4][4], y[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int i = ti + 16 * a, k = tk + 16 * b;
            t[a][b] = (i < w && k < w && i >= k) ? A[i + k * lda] : 0.0;
            y[a][b] = (i == k) ? 1.0 : 0.0;
        }
    for (int j = 0; j < w; j++) {
        double *cb = colbuf[j & 1], *rb = rowbuf[j & 1];
        // owners of column j of T (tk == j%16, b == j/16) publish their rows; owners of row j of Y publish it
        if (tk == (j & 15)) {
            const int b = j >> 4;
#pragma unroll
            for (int a = 0; a < 4; a++) {
                double v = 0.0;
#pragma unroll
                for (int bb = 0; bb < 4; bb++) if (bb == b) v = t[a][bb];
                cb[ti + 16 * a] = v;
            }
        }
        if (want_inv && ti == (j & 15)) {
            const int a = j >> 4;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                double v = 0.0;
#pragma unroll
                for (int aa = 0; aa < 4; aa++) if (aa == a) v = y[aa][b];
                rb[tk + 16 * b] = v;
            }
        }
        __syncthreads();
        const double d = cb[j];
        if (!(d > 0.0)) {                           // uniform: every thread reads the same d
            if (tid == 0) atomicMin(&info[job.snode], job.col0 + j + 1);
            break;
        }
        const double rinv = rsqrt(d), r = d * rinv, dinv = rinv * rinv;
        if (tid < w - j) { const int i = j + tid; A[i + j * lda] = (tid == 0) ? r : cb[i] * rinv; }   // w <= 64 < 256 threads
        double ci[4], ck[4];
#pragma unroll
        for (int a = 0; a < 4; a++) ci[a] = cb[ti + 16 * a];
#pragma unroll
        for (int b = 0; b < 4; b++) ck[b] = cb[tk + 16 * b] * dinv;
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int i = ti + 16 * a, k = tk + 16 * b;
                if (k > j && i >= k) t[a][b] -= ci[a] * ck[b];
            }
        if (want_inv) {
            // Y(j,:) /= L(j,j);  Y(i,:) -= L(i,j) * Y(j,:) for i > j      (L(i,j) = T(i,j) / r)
            double yr[4];
#pragma unroll
            for (int b = 0; b < 4; b++) yr[b] = rb[tk + 16 * b] * rinv;
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int i = ti + 16 * a;
                    if (i == j) y[a][b] = yr[b];
                    else if (i > j) y[a][b] -= (ci[a] * rinv) * yr[b];
                }
        }
    }
    if (want_inv) {
        double *__restrict__ W = winv + (long long) job.winv_slot * (NB_INNER * NB_INNER);
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) W[(ti + 16 * a) + NB_INNER * (tk + 16 * b)] = y[a][b];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// diagonal block Cholesky, version 3 (default): FOUR columns per barrier.  Same register layout as version 1 (thread
// (ti,tk) keeps T(ti+16a, tk+16b) and Y(ti+16a, tk+16b)); per step the owners publish the four panel columns (and the four
// panel rows of Y) as they are BEFORE the panel is factorized, one barrier, then EVERY thread factorizes the 4x4 diagonal
// block redundantly (four dependent rsqrt instead of four barrier round trips), forward-substitutes the panel entries of
// the 8 rows it needs, and applies a rank-4 update to its registers.  16 barriers per 64 columns instead of 64.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(POTRF_THREADS) potrf_block_kernel3(const PanelJob *__restrict__ jobs, double *__restrict__ Lx,
                                                                    int *__restrict__ info, double *__restrict__ winv)
{
    __shared__ double cbuf[2][4][NB_INNER];
    __shared__ double rbuf[2][4][NB_INNER];
    const PanelJob job = jobs[blockIdx.x];
    const int w = job.w, tid = threadIdx.x;
    const int ti = tid & 15, tk = tid >> 4;
    const long long lda = job.lda;
    const bool want_inv = job.winv_slot >= 0;
    double *__restrict__ A = Lx + job.x_off;
    double t[4][4], y[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int i = ti + 16 * a, k = tk + 16 * b;
            t[a][b] = (i < w && k < w && i >= k) ? A[i + k * lda] : 0.0;
            y[a][b] = (i == k) ? 1.0 : 0.0;
        }
    for (int j = 0, it = 0; j < w; j += 4, it++) {
        double (*cb)[NB_INNER] = cbuf[it & 1];
        double (*rb)[NB_INNER] = rbuf[it & 1];
        const int jb = j >> 4, jq = j & 15;            // the four columns share the 16-block jb; their owners have tk = jq..jq+3
        if (tk >= jq && tk < jq + 4) {
            const int c = tk - jq;
#pragma unroll
            for (int a = 0; a < 4; a++) {
                double v = 0.0;
#pragma unroll
                for (int bb = 0; bb < 4; bb++) if (bb == jb) v = t[a][bb];
                cb[c][ti + 16 * a] = v;
            }
        }
        if (want_inv && ti >= jq && ti < jq + 4) {
            const int c = ti - jq;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                double v = 0.0;
#pragma unroll
                for (int aa = 0; aa < 4; aa++) if (aa == jb) v = y[aa][b];
                rb[c][tk + 16 * b] = v;
            }
        }
        __syncthreads();
        // 4x4 diagonal block, factorized by every thread (identical arithmetic everywhere: uniform control flow)
        double l[4][4], rinv[4];
        int failc = -1;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            double dcc = cb[c][j + c];
#pragma unroll
            for (int q = 0; q < c; q++) dcc -= l[c][q] * l[c][q];
            if (j + c >= w) dcc = 1.0;                 // padding beyond the block: identity
            if (!(dcc > 0.0)) { if (failc < 0) failc = c; dcc = 1.0; }
            const double ri = rsqrt(dcc);
            rinv[c] = ri; l[c][c] = dcc * ri;
#pragma unroll
            for (int r = c + 1; r < 4; r++) {
                double v = cb[c][j + r];
#pragma unroll
                for (int q = 0; q < c; q++) v -= l[r][q] * l[c][q];
                l[r][c] = v * ri;
            }
        }
        if (failc >= 0) {                               // first non-positive (or NaN) pivot: LAPACK's info
            if (tid == 0) atomicMin(&info[job.snode], job.col0 + j + failc + 1);
            break;
        }
        // panel entries L(i, j..j+3) of the rows this thread needs: i = ti+16a (Li) and i = tk+16b (Lk).  Rows at or above
        // the panel give values that are never used (updates touch k > j+3 only, stores i >= column only).
        double Li[4][4], Lk[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int i = ti + 16 * a;
            const double x0 = cb[0][i] * rinv[0];
            const double x1 = (cb[1][i] - x0 * l[1][0]) * rinv[1];
            const double x2 = (cb[2][i] - x0 * l[2][0] - x1 * l[2][1]) * rinv[2];
            const double x3 = (cb[3][i] - x0 * l[3][0] - x1 * l[3][1] - x2 * l[3][2]) * rinv[3];
            Li[a][0] = x0; Li[a][1] = x1; Li[a][2] = x2; Li[a][3] = x3;
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int k = tk + 16 * b;
            const double x0 = cb[0][k] * rinv[0];
            const double x1 = (cb[1][k] - x0 * l[1][0]) * rinv[1];
            const double x2 = (cb[2][k] - x0 * l[2][0] - x1 * l[2][1]) * rinv[2];
            const double x3 = (cb[3][k] - x0 * l[3][0] - x1 * l[3][1] - x2 * l[3][2]) * rinv[3];
            Lk[b][0] = x0; Lk[b][1] = x1; Lk[b][2] = x2; Lk[b][3] = x3;
        }
        // the finished panel columns go to global memory: threads tk < 4 write column j+tk, rows ti+16a
        if (tk < 4 && j + tk < w) {
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const int i = ti + 16 * a;
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < 4; c++) if (c == tk) v = Li[a][c];
                if (i >= j + tk && i < w) A[i + (long long) (j + tk) * lda] = v;
            }
        }
        // rank-4 update of the trailing part
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int i = ti + 16 * a, k = tk + 16 * b;
                if (k > j + 3 && i >= k)
                    t[a][b] -= Li[a][0] * Lk[b][0] + Li[a][1] * Lk[b][1] + Li[a][2] * Lk[b][2] + Li[a][3] * Lk[b][3];
            }
        if (want_inv) {
            // rows j..j+3 of Y = L^{-1} become final (4-step forward substitution), then Y(i,:) -= L(i,panel) * Y(panel,:)
            double Yf[4][4];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int k = tk + 16 * b;
                const double y0 = rb[0][k] * rinv[0];
                const double y1 = (rb[1][k] - l[1][0] * y0) * rinv[1];
                const double y2 = (rb[2][k] - l[2][0] * y0 - l[2][1] * y1) * rinv[2];
                const double y3 = (rb[3][k] - l[3][0] * y0 - l[3][1] * y1 - l[3][2] * y2) * rinv[3];
                Yf[0][b] = y0; Yf[1][b] = y1; Yf[2][b] = y2; Yf[3][b] = y3;
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int i = ti + 16 * a;
                    if (i > j + 3) y[a][b] -= Li[a][0] * Yf[0][b] + Li[a][1] * Yf[1][b] + Li[a][2] * Yf[2][b] + Li[a][3] * Yf[3][b];
                    else if (i >= j) {
#pragma unroll
                        for (int c = 0; c < 4; c++) if (i == j + c) y[a][b] = Yf[c][b];
                    }
                }
        }
    }
    if (want_inv) {
        double *__restrict__ W = winv + (long long) job.winv_slot * (NB_INNER * NB_INNER);
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) W[(ti + 16 * a) + NB_INNER * (tk + 16 * b)] = y[a][b];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// diagonal block Cholesky, version 4 (default): version 3's four-columns-per-step scheme with the rank-4 updates on the
// fp64 tensor cores.  ncu on version 3 (profiles/r2_ncu_full_potrf3.txt): 35 us, one SM, stall reason "wait" - the ~290
// dependent fp64 instructions per thread and step, not the barriers.  Here the 64x64 tile T and its inverse Y (both lower
// triangular, 36 8x8 tiles each) live in DMMA accumulator fragments: warp I owns tile row I of T and tile row 7-I of Y
// (9 tiles per warp).  Per step: the owners publish the four panel columns of T and the four panel rows of Y; barrier;
// every thread factorizes the 4x4 diagonal block (four dependent rsqrt); thread t then computes ONE final panel entry
// L(t%64, j+t/64) and ONE final Y(j+t/64, t%64), stores them to shared memory (and L to global); barrier; each warp
// applies T -= Lp Lp' and Y -= Lp Yp with one m8n8k4 DMMA per tile (K = 4 is exactly the panel width).
// Entries at or above the panel are masked to zero in the fragments, so no garbage ever enters the accumulators.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(POTRF_THREADS) potrf_block_kernel4(const PanelJob *__restrict__ jobs, double *__restrict__ Lx,
                                                                    int *__restrict__ info, double *__restrict__ winv)
{
    __shared__ double cb[4][NB_INNER];      // panel columns of T before the panel is factorized
    __shared__ double rb[4][NB_INNER];      // panel rows of Y before
    __shared__ double Lp[4][NB_INNER];      // final L(row, j+c), 0 for rows at or above the panel
    __shared__ double Yp[4][NB_INNER];      // final Y(j+c, col)
    const PanelJob job = jobs[blockIdx.x];
    const int w = job.w, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gr = lane >> 2, gc = lane & 3;            // fragment coordinates: C(gr, 2gc+e), A(gr, gc), B(gc, gr)
    const int I = warp, IY = 7 - warp;
    const long long lda = job.lda;
    const bool want_inv = job.winv_slot >= 0;
    double *__restrict__ A = Lx + job.x_off;
    double tT[8][2], tY[8][2];
#pragma unroll
    for (int K = 0; K < 8; K++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int i = 8 * I + gr, k = 8 * K + 2 * gc + e;
            tT[K][e] = (K <= I && i < w && k < w && i >= k) ? A[i + k * lda] : 0.0;
            tY[K][e] = (8 * IY + gr == k) ? 1.0 : 0.0;
        }
    for (int e = tid; e < 4 * NB_INNER; e += POTRF_THREADS) (&rb[0][0])[e] = 0.0;
    __syncthreads();
    for (int j = 0; j < w; j += 4) {
        const int Jt = j >> 3, jj = j & 7;
        // publish columns j..j+3 of T (rows of this warp's tile row) and rows j..j+3 of Y
        if (I >= Jt && (gc >> 1) == (jj >> 2)) {
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int K = 0; K < 8; K++) if (K == Jt) { v0 = tT[K][0]; v1 = tT[K][1]; }
            const int c = 2 * (gc & 1);
            cb[c][8 * I + gr] = v0; cb[c + 1][8 * I + gr] = v1;
        }
        if (want_inv && IY == Jt && gr >= jj && gr < jj + 4) {
            const int c = gr - jj;
#pragma unroll
            for (int K = 0; K < 8; K++) if (K <= Jt) { rb[c][8 * K + 2 * gc] = tY[K][0]; rb[c][8 * K + 2 * gc + 1] = tY[K][1]; }
        }
        __syncthreads();
        // 4x4 diagonal block, factorized by every thread
        double l[4][4], rinv[4];
        int failc = -1;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            double dcc = cb[c][j + c];
#pragma unroll
            for (int q = 0; q < c; q++) dcc -= l[c][q] * l[c][q];
            if (j + c >= w) dcc = 1.0;
            if (!(dcc > 0.0)) { if (failc < 0) failc = c; dcc = 1.0; }
            const double ri = rsqrt(dcc);
            rinv[c] = ri; l[c][c] = dcc * ri;
#pragma unroll
            for (int r = c + 1; r < 4; r++) {
                double v = cb[c][j + r];
#pragma unroll
                for (int q = 0; q < c; q++) v -= l[r][q] * l[c][q];
                l[r][c] = v * ri;
            }
        }
        if (failc >= 0) {
            if (tid == 0) atomicMin(&info[job.snode], job.col0 + j + failc + 1);
            break;
        }
        // one final panel entry of L and one of Y per thread
        {
            const int row = tid & 63, c = tid >> 6;
            const double x0 = cb[0][row] * rinv[0];
            const double x1 = (cb[1][row] - x0 * l[1][0]) * rinv[1];
            const double x2 = (cb[2][row] - x0 * l[2][0] - x1 * l[2][1]) * rinv[2];
            const double x3 = (cb[3][row] - x0 * l[3][0] - x1 * l[3][1] - x2 * l[3][2]) * rinv[3];
            const double xc = c == 0 ? x0 : c == 1 ? x1 : c == 2 ? x2 : x3;
            Lp[c][row] = row > j + 3 ? xc : 0.0;
            if (row >= j + c && row < w && j + c < w) A[row + (long long) (j + c) * lda] = xc;
            if (want_inv) {
                const double y0 = rb[0][row] * rinv[0];
                const double y1 = (rb[1][row] - l[1][0] * y0) * rinv[1];
                const double y2 = (rb[2][row] - l[2][0] * y0 - l[2][1] * y1) * rinv[2];
                const double y3 = (rb[3][row] - l[3][0] * y0 - l[3][1] * y1 - l[3][2] * y2) * rinv[3];
                Yp[c][row] = c == 0 ? y0 : c == 1 ? y1 : c == 2 ? y2 : y3;
            }
        }
        __syncthreads();
        // T(i,k) -= sum_c Lp[c][i] Lp[c][k] for the tiles right of the panel; rows/columns at or above the panel are zero in Lp
        {
            const double a = -Lp[gc][8 * I + gr];
#pragma unroll
            for (int K = 0; K < 8; K++)
                if (K >= Jt && K <= I) dmma884(tT[K][0], tT[K][1], a, Lp[gc][8 * K + gr]);
        }
        if (want_inv) {
            // Y(i,:) -= sum_c Lp[c][i] Yp[c][:] for the rows below the panel; the panel rows themselves become Yp
            const double a = -Lp[gc][8 * IY + gr];
#pragma unroll
            for (int K = 0; K < 8; K++)
                if (K <= IY && K <= Jt) dmma884(tY[K][0], tY[K][1], a, Yp[gc][8 * K + gr]);
            if (IY == Jt && gr >= jj && gr < jj + 4) {
                const int c = gr - jj;
#pragma unroll
                for (int K = 0; K < 8; K++) if (K <= Jt) { tY[K][0] = Yp[c][8 * K + 2 * gc]; tY[K][1] = Yp[c][8 * K + 2 * gc + 1]; }
            }
        }
    }
    if (want_inv) {
        double *__restrict__ W = winv + (long long) job.winv_slot * (NB_INNER * NB_INNER);
#pragma unroll
        for (int K = 0; K < 8; K++)
#pragma unroll
            for (int e = 0; e < 2; e++) W[(8 * IY + gr) + NB_INNER * (8 * K + 2 * gc + e)] = K <= IY ? tY[K][e] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// diagonal block Cholesky, version 2: blocked by 16-column sub-panels so that the column-by-column part has NO block
// barriers.  Per sub-panel: (1) ONE warp factorizes the 16x16 diagonal sub-block in registers with shuffles - lanes 0..15
// hold its rows, lanes 16..31 hold the rows of its inverse, built by the same elimination; (2) all threads bring the rows
// below up to date (R <- R * D^{-T} with the 16x16 inverse); (3) rank-16 update of the rest of the tile.  Then W = L^{-1}
// is assembled block by block from the four 16x16 inverses.  ~4x fewer serialized steps than version 1.
// ------------------------------------------------------------------------------------------------------------------
constexpr int POTRF2_LDT = NB_INNER + 1;
constexpr size_t potrf2_smem_bytes() { return (size_t) (2 * NB_INNER * POTRF2_LDT + 4 * 16 * 17 + 16 * 17) * sizeof(double); }

__global__ void __launch_bounds__(POTRF_THREADS) potrf_block_kernel2(const PanelJob *__restrict__ jobs, double *__restrict__ Lx,
                                                                    int *__restrict__ info, double *__restrict__ winv)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *T = reinterpret_cast<double *>(smem_raw);               // T[i*LDT + j], lower part of the tile
    double *Wf = T + NB_INNER * POTRF2_LDT;                         // W[i*LDT + j]
    double *Dinv = Wf + NB_INNER * POTRF2_LDT;                      // Dinv[p][a*17 + b]
    double *Sb = Dinv + 4 * 16 * 17;                                // scratch block
    __shared__ int failcol;
    constexpr int LDT = POTRF2_LDT;
    const PanelJob job = jobs[blockIdx.x];
    const int w = job.w, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long lda = job.lda;
    const bool want_inv = job.winv_slot >= 0;
    double *__restrict__ A = Lx + job.x_off;
    // tile, padded with the identity beyond w (the padded part factorizes trivially)
    for (int e = tid; e < NB_INNER * NB_INNER; e += POTRF_THREADS) {
        const int i = e & 63, j = e >> 6;
        T[i * LDT + j] = (i < w && j < w && i >= j) ? A[i + j * lda] : ((i == j) ? 1.0 : 0.0);
        Wf[i * LDT + j] = 0.0;
    }
    if (tid == 0) failcol = NB_INNER;
    __syncthreads();
    for (int p = 0; p < 4; p++) {
        const int c0 = 16 * p;
        if (c0 >= w) break;                                         // uniform
        if (warp == 0) {
            const int l = lane & 15;
            const bool isY = lane >= 16;
            double a[16];
#pragma unroll
            for (int c = 0; c < 16; c++) a[c] = isY ? ((c == l) ? 1.0 : 0.0) : ((c <= l) ? T[(c0 + l) * LDT + c0 + c] : 0.0);
            int bad = NB_INNER;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                double d = __shfl_sync(0xffffffffu, a[j], j);       // pivot, from the lane that holds row j of the block
                if (!(d > 0.0)) { bad = min(bad, c0 + j); d = 1.0; }
                const double rinv = rsqrt(d);
                if (!isY) a[j] = (l == j) ? d * rinv : a[j] * rinv; // column j of L (zero above the diagonal)
                const double lj = __shfl_sync(0xffffffffu, a[j], l);    // L(l,j) for this lane's row, also for the inverse lanes
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k > j) {                                    // compile-time after unrolling
                        const double lkj = __shfl_sync(0xffffffffu, a[j], k);
                        if (!isY && l >= k) a[k] -= lj * lkj;
                    }
                }
                if (isY && l == j) {
#pragma unroll
                    for (int c = 0; c < 16; c++) a[c] *= rinv;      // Y(j,:) /= L(j,j)
                }
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    if (c <= j) {                                   // compile-time after unrolling
                        const double yjc = __shfl_sync(0xffffffffu, a[c], 16 + j);
                        if (isY && l > j) a[c] -= lj * yjc;         // Y(l,:) -= L(l,j) Y(j,:)
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 16; c++) {
                if (!isY) { if (c <= l) T[(c0 + l) * LDT + c0 + c] = a[c]; }
                else Dinv[p * 272 + l * 17 + c] = (c <= l) ? a[c] : 0.0;
            }
            if (lane == 0 && bad < NB_INNER) failcol = min(failcol, bad);
        }
        __syncthreads();
        const int r0 = c0 + 16, nrows = NB_INNER - r0;
        if (nrows > 0) {
            // (2) R <- R * D^{-T}: new(r,c) = sum_{k<=c} R(r,k) Dinv(c,k)
            double nv[3];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int e = tid + q * POTRF_THREADS;
                nv[q] = 0.0;
                if (e < nrows * 16) {
                    const int r = r0 + (e >> 4), c = e & 15;
                    double acc = 0.0;
                    for (int k = 0; k <= c; k++) acc += T[r * LDT + c0 + k] * Dinv[p * 272 + c * 17 + k];
                    nv[q] = acc;
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int e = tid + q * POTRF_THREADS;
                if (e < nrows * 16) T[(r0 + (e >> 4)) * LDT + c0 + (e & 15)] = nv[q];
            }
            __syncthreads();
            // (3) rank-16 update of the rest of the tile (lower part)
            for (int e = tid; e < nrows * nrows; e += POTRF_THREADS) {
                const int r = r0 + e % nrows, c = r0 + e / nrows;
                if (r >= c) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 16; k++) acc += T[r * LDT + c0 + k] * T[c * LDT + c0 + k];
                    T[r * LDT + c] -= acc;
                }
            }
            __syncthreads();
        }
    }
    // L goes back to Lx
    for (int e = tid; e < NB_INNER * NB_INNER; e += POTRF_THREADS) {
        const int i = e & 63, j = e >> 6;
        if (i < w && j < w && i >= j) A[i + j * lda] = T[i * LDT + j];
    }
    if (tid == 0 && failcol < w) atomicMin(&info[job.snode], job.col0 + failcol + 1);
    if (!want_inv) return;
    // W = L^{-1}: diagonal 16x16 blocks are Dinv; W(i,q) = -Dinv_i * sum_{k=q}^{i-1} L(i,k) W(k,q), block column by block column
    {
        const int a = tid >> 4, b = tid & 15;                       // one thread per entry of a 16x16 block
        for (int pb = 0; pb < 4; pb++) Wf[(16 * pb + a) * LDT + 16 * pb + b] = Dinv[pb * 272 + a * 17 + b];
        __syncthreads();
        for (int q = 0; q < 3; q++) {
            for (int i = q + 1; i < 4; i++) {
                double acc = 0.0;
                for (int k = q; k < i; k++)
#pragma unroll
                    for (int t = 0; t < 16; t++) acc += T[(16 * i + a) * LDT + 16 * k + t] * Wf[(16 * k + t) * LDT + 16 * q + b];
                Sb[a * 17 + b] = acc;
                __syncthreads();
                double v = 0.0;
                for (int t = 0; t <= a; t++) v += Dinv[i * 272 + a * 17 + t] * Sb[t * 17 + b];
                Wf[(16 * i + a) * LDT + 16 * q + b] = -v;
                __syncthreads();
            }
        }
    }
    double *__restrict__ W = winv + (long long) job.winv_slot * (NB_INNER * NB_INNER);
    for (int e = tid; e < NB_INNER * NB_INNER; e += POTRF_THREADS) W[e] = Wf[(e & 63) * LDT + (e >> 6)];
}

// ------------------------------------------------------------------------------------------------------------------
// rows below the diagonal block:  B <- B * L11^{-T}.  One CTA = TRSM_ROWS rows, one thread per row, the row kept in
// registers (NC = compile-time bound on the block width), L11 broadcast from shared memory.
// ------------------------------------------------------------------------------------------------------------------
template <int NC>
__device__ __forceinline__ void trsm_rows_body(const PanelJob &job, int tile, double *__restrict__ Lx, double *Lsm)
{
    constexpr int LDS = NB_INNER + 1;
    const int w = job.w, tid = threadIdx.x;
    const long long lda = job.lda;
    const double *__restrict__ L11 = Lx + job.x_off;
    for (int e = tid; e < NC * NC; e += TRSM_ROWS) {
        const int i = e % NC, j = e / NC;              // NC is a power of two: shifts
        Lsm[i * LDS + j] = (i < w && j < w && i >= j) ? L11[i + j * lda] : 0.0;
    }
    __syncthreads();
    const int r = tile * TRSM_ROWS + tid;
    if (r >= job.rows_below) return;
    double *__restrict__ B = Lx + job.x_off + w + r;      // row r below the block, column c at B[c*lda]
    double x[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) x[c] = (c < w) ? B[c * lda] : 0.0;
    // right-looking substitution: after x[j] is final, it is eliminated from every later column (independent FMAs)
#pragma unroll
    for (int j = 0; j < NC; j++) {
        if (j < w) {
            const double xj = x[j] / Lsm[j * LDS + j];
            x[j] = xj;
#pragma unroll
            for (int k = j + 1; k < NC; k++) x[k] -= xj * Lsm[k * LDS + j];       // rows k >= w of Lsm are zero
        }
    }
#pragma unroll
    for (int c = 0; c < NC; c++) if (c < w) B[c * lda] = x[c];
}

__global__ void __launch_bounds__(TRSM_ROWS) trsm_rows_kernel(const PanelJob *__restrict__ jobs, const int *__restrict__ tile_job,
                                                              double *__restrict__ Lx)
{
    __shared__ double Lsm[NB_INNER * (NB_INNER + 1)];
    const PanelJob job = jobs[tile_job[blockIdx.x]];
    const int tile = (int) blockIdx.x - job.tile_start;
    if (job.w <= 8) trsm_rows_body<8>(job, tile, Lx, Lsm);
    else if (job.w <= 16) trsm_rows_body<16>(job, tile, Lx, Lsm);
    else if (job.w <= 32) trsm_rows_body<32>(job, tile, Lx, Lsm);
    else trsm_rows_body<64>(job, tile, Lx, Lsm);
}

// ------------------------------------------------------------------------------------------------------------------
// rows below a WIDE diagonal block on the tensor cores:  B <- B * W^T with W = L11^{-1} from potrf_block_kernel.
// One CTA = 128 rows x (<= 64) columns, all of K (= w <= 64) resident in shared memory, 4 warps x (32 x 64) DMMA tiles.
// In place: a CTA reads only its own rows (all columns) before it writes them.
// ------------------------------------------------------------------------------------------------------------------
constexpr int TRSM_TC_LDA = TRSM_ROWS + 4, TRSM_TC_LDW = NB_INNER + 4;
constexpr size_t trsm_tc_smem_bytes() { return (size_t) NB_INNER * (TRSM_TC_LDA + TRSM_TC_LDW) * sizeof(double); }

__global__ void __launch_bounds__(TRSM_ROWS) trsm_tc_kernel(const PanelJob *__restrict__ jobs, const int *__restrict__ tile_job,
                                                            double *__restrict__ Lx, const double *__restrict__ winv)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *As = reinterpret_cast<double *>(smem_raw);          // As[k][r], k-major
    double *Ws = As + NB_INNER * TRSM_TC_LDA;                   // Ws[k][j] = W(j,k)
    const PanelJob job = jobs[tile_job[blockIdx.x]];
    const int tile = (int) blockIdx.x - job.tile_start;
    const int w = job.w, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long lda = job.lda;
    const int r0 = tile * TRSM_ROWS;
    const int rows = min(TRSM_ROWS, job.rows_below - r0);
    double *__restrict__ B = Lx + job.x_off + w + r0;           // row r, column k at B[r + k*lda]
    const double *__restrict__ W = winv + (long long) job.winv_slot * (NB_INNER * NB_INNER);
    {
        const bool okr = tid < rows;
        const double *src = B + (okr ? tid : 0);
        for (int k = 0; k < NB_INNER; k++) {
            const bool ok = okr && k < w;
            cp_async8(As + k * TRSM_TC_LDA + tid, ok ? src + (long long) k * lda : B, ok ? 8 : 0);
        }
        for (int e = tid; e < NB_INNER * NB_INNER; e += TRSM_ROWS) {
            const int j = e % NB_INNER, k = e / NB_INNER;
            cp_async8(Ws + k * TRSM_TC_LDW + j, W + e, 8);
        }
        cp_async_commit();
    }
    double acc[4][8][2];
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
        for (int nn = 0; nn < 8; nn++) acc[m][nn][0] = acc[m][nn][1] = 0.0;
    cp_async_wait<0>();
    __syncthreads();
    const double *ap = As + (lane & 3) * TRSM_TC_LDA + warp * 32 + (lane >> 2);
    const double *bp = Ws + (lane & 3) * TRSM_TC_LDW + (lane >> 2);
    const int nks = (w + 3) >> 2;
    for (int ks = 0; ks < nks; ks++) {
        double a[4], b[8];
#pragma unroll
        for (int m = 0; m < 4; m++) a[m] = ap[ks * 4 * TRSM_TC_LDA + m * 8];
#pragma unroll
        for (int nn = 0; nn < 8; nn++) b[nn] = bp[ks * 4 * TRSM_TC_LDW + nn * 8];
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
            for (int nn = 0; nn < 8; nn++) dmma884(acc[m][nn][0], acc[m][nn][1], a[m], b[nn]);
    }
#pragma unroll
    for (int nn = 0; nn < 8; nn++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int j = nn * 8 + 2 * (lane & 3) + e;
            if (j >= w) continue;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int r = warp * 32 + m * 8 + (lane >> 2);
                if (r < rows) B[r + (long long) j * lda] = acc[m][nn][e];
            }
        }
}

// ------------------------------------------------------------------------------------------------------------------
// assembly
// ------------------------------------------------------------------------------------------------------------------
struct DevSym {                 // device view of the symbolic factor
    const int *super;           // nsuper+1
    const long long *pi, *px;   // nsuper+1
    const int *ls;              // ssize
    const int *supermap;        // n
    long long n, nsuper;
};

struct DevCsc { const long long *p, *i, *nz; const double *x; };

__device__ __forceinline__ int find_row(const int *__restrict__ rows, int nrows, int target)
{
    int lo = 0, hi = nrows - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int v = rows[mid];
        if (v == target) return mid;
        if (v < target) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

// one thread per column k of L: copy A(k:n,k) (stype<0) or (A*F)(k:n,k) (stype==0) into its supernode, add beta.
// Entries outside the pattern of L are dropped (the reference only avoids the segfault, :366-378).
// [kfirst, kfirst+kcount) restricts the kernel to a column range (not-posdef repeat of one supernode).
__global__ void scatter_A_kernel(DevSym sym, int stype, DevCsc A, DevCsc F, double beta, double *__restrict__ Lx,
                                 long long kfirst, long long kcount, const int *__restrict__ owner, int nranks, int rank)
{
    const long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (t >= kcount) return;
    const long long k = kfirst + t;
    const int s = sym.supermap[k];
    const int k1 = sym.super[s];
    if (owner) {        // sharded factorization: only the rank that computes this column assembles it
        const int o = owner[s];
        if (o >= 0 ? (o != rank) : (((int) ((k - k1) / NB_MID) + (-1 - o)) % nranks != rank)) return;      // cyclic: owner = -1 - offset
    }
    const long long psi = sym.pi[s];
    const int nsrow = (int) (sym.pi[s + 1] - psi);
    const int *__restrict__ rows = sym.ls + psi;
    long long base = sym.px[s] + (k - k1) * (long long) nsrow;
    if (sym.px[s] < -1) {
        // transient root of the multi-GPU path (px = -2 - base of the packed own panels): this rank stores only its own
        // 256-column panels, back to back
        const int off = -1 - owner[s];
        const long long J = (k - k1) / NB_MID;
        const long long Jf = ((rank - off) % nranks + nranks) % nranks;
        base = (-2 - sym.px[s]) + ((J - Jf) / nranks) * (long long) NB_MID * nsrow + ((k - k1) - J * NB_MID) * (long long) nsrow;
    }
    double *__restrict__ col = Lx + base;
    if (stype != 0) {
        long long p = A.p[k];
        const long long pend = A.nz ? p + A.nz[k] : A.p[k + 1];
        for (; p < pend; p++) {
            const long long i = A.i[p];
            if (i >= k) { const int im = find_row(rows, nsrow, (int) i); if (im >= 0) col[im] = A.x[p]; }
        }
    } else {
        long long pf = F.p[k];
        const long long pfend = F.nz ? pf + F.nz[k] : F.p[k + 1];
        for (; pf < pfend; pf++) {
            const long long j = F.i[pf];
            const double fjk = F.x[pf];
            long long p = A.p[j];
            const long long pend = A.nz ? p + A.nz[j] : A.p[j + 1];
            for (; p < pend; p++) {
                const long long i = A.i[p];
                if (i >= k) { const int im = find_row(rows, nsrow, (int) i); if (im >= 0) col[im] += A.x[p] * fjk; }
            }
        }
    }
    if (beta != 0.0) col[k - k1] += beta;
}

struct DevUpdate { long long ls_d, ls_s, map_off; int nsrow_s, nd2; };

// RelativeMap: position of each remaining row of d inside s's row list.  One CTA per update.
__global__ void relmap_kernel(const DevUpdate *__restrict__ ups, const int *__restrict__ ls, int *__restrict__ relmap)
{
    const DevUpdate u = ups[blockIdx.x];
    const int *__restrict__ rows_s = ls + u.ls_s;
    for (int i = threadIdx.x; i < u.nd2; i += blockDim.x)
        relmap[u.map_off + i] = find_row(rows_s, u.nsrow_s, ls[u.ls_d + i]);
}

// diagonal of L (what cholmod_rcond reads, Cholesky/cholmod_rcond.c:102-125): one thread per column
__global__ void factor_diag_kernel(DevSym sym, const double *__restrict__ Lx, double *__restrict__ diag)
{
    const long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (k >= sym.n) return;
    const int s = sym.supermap[k];
    const long long j = k - sym.super[s];
    const long long nsrow = sym.pi[s + 1] - sym.pi[s];
    diag[k] = Lx[sym.px[s] + j + j * nsrow];
}

// Y = P B and X = P' Y of cholmod_solve2 (Cholesky/cholmod_solve.c:1552-1580) on the device: Y(k,:) = B(Perm[k],:), X(Perm[k],:) = Y(k,:)
__global__ void perm_gather_kernel(const long long *__restrict__ perm, const double *__restrict__ B, long long ldb, double *__restrict__ Y,
                                   long long n, int nrhs)
{
    const long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (k >= n) return;
    const long long src = perm ? perm[k] : k;
    for (int r = 0; r < nrhs; r++) Y[k + r * n] = B[src + r * ldb];
}
__global__ void perm_scatter_kernel(const long long *__restrict__ perm, const double *__restrict__ Y, double *__restrict__ X, long long ldx,
                                    long long n, int nrhs)
{
    const long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (k >= n) return;
    const long long dst = perm ? perm[k] : k;
    for (int r = 0; r < nrhs; r++) X[dst + r * ldx] = Y[k + r * n];
}

// Complex Hermitian factors through the real kernels (t_cholmod_super_numeric.c:41-83 are the reference's zherk/zgemm/zpotrf/
// ztrsm instantiations).  With every complex entry a+ib replaced by the 2x2 block [a -b; b a] (rows/columns 2i, 2i+1), a
// Hermitian positive definite A becomes a real SPD matrix of twice the order whose Cholesky factor is exactly the
// blockified complex factor (its diagonal blocks are l*I with l real and positive, so it is lower triangular, and the
// Cholesky factor is unique).  These two kernels convert between the blockified real factor (the plan's d_Lx, supernode s =
// 2nsrow x 2nscol at px2[s]) and CHOLMOD's complex layout (nsrow x nscol interleaved (re,im) pairs at pxc[s]).
__global__ void cplx_compress_kernel(const long long *__restrict__ pxc, const long long *__restrict__ px2, const long long *__restrict__ pi2,
                                     long long nsuper, long long xsize_c, const double *__restrict__ Lt, double *__restrict__ Lc)
{
    const long long e = blockIdx.x * (long long) blockDim.x + threadIdx.x;       // complex entry
    if (e >= xsize_c) return;
    long long lo = 0, hi = nsuper;                                                // supernode with pxc[s] <= e < pxc[s+1]
    while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (pxc[mid] <= e) lo = mid; else hi = mid; }
    const long long s = lo, nsrow2 = pi2[s + 1] - pi2[s], nsrow = nsrow2 >> 1;
    const long long q = e - pxc[s];
    if (q >= nsrow * ((px2[s + 1] - px2[s]) / (2 * nsrow2))) { Lc[2 * e] = 0.0; Lc[2 * e + 1] = 0.0; return; }   // padding behind the block
    const long long i = q % nsrow, j = q / nsrow;
    const double *__restrict__ col = Lt + px2[s] + (2 * j) * nsrow2;
    Lc[2 * e] = col[2 * i]; Lc[2 * e + 1] = col[2 * i + 1];
}

__global__ void cplx_expand_kernel(const long long *__restrict__ pxc, const long long *__restrict__ px2, const long long *__restrict__ pi2,
                                   long long nsuper, long long xsize_c, const double *__restrict__ Lc, double *__restrict__ Lt)
{
    const long long e = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (e >= xsize_c) return;
    long long lo = 0, hi = nsuper;
    while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (pxc[mid] <= e) lo = mid; else hi = mid; }
    const long long s = lo, nsrow2 = pi2[s + 1] - pi2[s], nsrow = nsrow2 >> 1;
    const long long q = e - pxc[s];
    if (q >= nsrow * ((px2[s + 1] - px2[s]) / (2 * nsrow2))) return;
    const long long i = q % nsrow, j = q / nsrow;
    const double re = Lc[2 * e], im = (i == j) ? 0.0 : Lc[2 * e + 1];             // the diagonal of L is real
    double *__restrict__ c0 = Lt + px2[s] + (2 * j) * nsrow2, *__restrict__ c1 = c0 + nsrow2;
    c0[2 * i] = re; c0[2 * i + 1] = im;
    c1[2 * i] = (i == j) ? 0.0 : -im; c1[2 * i + 1] = re;
}

__global__ void fill_int_kernel(int *p, long long n, int v)
{
    const long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (t < n) p[t] = v;
}

// ------------------------------------------------------------------------------------------------------------------
// triangular solves.  A "solve block" is a <=64-column slice of a supernode: its diagonal block, and every row of the
// supernode below it (rows of the diagonal part are addressed through Ls too: Ls[psi+r] = k1+r for r < nscol).
// ------------------------------------------------------------------------------------------------------------------
constexpr int SOLVE_THREADS = 128;

// forward, step (a): x1 <- L11^{-1} x1 for every right-hand side.  Wide blocks have their inverse W = L11^{-1} from the
// factorization: the solve is a 64x64 mat-vec (no column-by-column barriers); narrow blocks substitute in shared memory.
__global__ void __launch_bounds__(SOLVE_THREADS) lsolve_diag_kernel(const SolveJob *__restrict__ jobs, const double *__restrict__ Lx,
                                                                   const double *__restrict__ winv, double *__restrict__ X, int nrhs, long long ldx)
{
    constexpr int LDS = NB_INNER + 1;
    __shared__ double T[NB_INNER * LDS];
    __shared__ double xs[NB_INNER];
    const SolveJob job = jobs[blockIdx.x];
    const int w = job.w, tid = threadIdx.x;
    if (winv && job.winv_slot >= 0) {
        const double *__restrict__ W = winv + (long long) job.winv_slot * (NB_INNER * NB_INNER);   // W(i,k) at W[i + 64k]
        for (int r = 0; r < nrhs; r++) {
            double *__restrict__ x = X + r * ldx + job.xcol0;
            __syncthreads();
            if (tid < w) xs[tid] = x[tid];
            __syncthreads();
            // two threads per row: even/odd k, 32 independent loads each, coalesced over the rows
            const int i = tid & 63, half = tid >> 6;
            double acc = 0.0;
            if (i < w) {
#pragma unroll 8
                for (int k = half; k <= i; k += 2) acc += W[i + NB_INNER * k] * xs[k];
            }
            T[tid] = acc;
            __syncthreads();
            if (tid < w) x[tid] = T[tid] + T[tid + 64];
        }
        return;
    }
    const double *__restrict__ A = Lx + job.x_off;
    for (int e = tid; e < w * w; e += SOLVE_THREADS) {
        const int i = e % w, j = e / w;
        if (i >= j) T[i * LDS + j] = A[i + (long long) j * job.lda];
    }
    for (int r = 0; r < nrhs; r++) {
        double *__restrict__ x = X + r * ldx + job.xcol0;
        __syncthreads();
        if (tid < w) xs[tid] = x[tid];
        __syncthreads();
        for (int j = 0; j < w; j++) {
            if (tid == j) xs[j] = xs[j] / T[j * LDS + j];
            __syncthreads();
            if (tid > j && tid < w) xs[tid] -= T[tid * LDS + j] * xs[j];
            __syncthreads();
        }
        if (tid < w) x[tid] = xs[tid];
    }
}

// forward, step (b): X[rows below] -= L2 * x1.  One CTA = SOLVE_ROWS rows of one job, one thread per row (coalesced
// along every column), 8 loads in flight per thread; atomics because several supernodes of a level hit the same rows.
__global__ void __launch_bounds__(SOLVE_THREADS) lsolve_update_kernel(const SolveJob *__restrict__ jobs, const int *__restrict__ tile_job,
                                                                     const double *__restrict__ Lx, const int *__restrict__ ls,
                                                                     double *__restrict__ X, int nrhs, long long ldx)
{
    __shared__ double xs[NB_INNER];
    const SolveJob job = jobs[tile_job[blockIdx.x]];
    const int tile = (int) blockIdx.x - job.tile_start;
    const int w = job.w, tid = threadIdx.x;
    const int r = tile * SOLVE_ROWS + tid;
    const bool live = r < job.rows_below;
    const long long lda = job.lda;
    const double *__restrict__ row = Lx + job.x_off + w + (live ? r : 0);     // row r below the block at row[c*lda]
    const int xrow = live ? ls[job.ls_off + r] : 0;
    for (int rh = 0; rh < nrhs; rh++) {
        double *__restrict__ x = X + rh * ldx;
        __syncthreads();
        if (tid < w) xs[tid] = x[job.xcol0 + tid];
        __syncthreads();
        double acc = 0.0;
        int c = 0;
        for (; c + 32 <= w; c += 32) {                  // 32 independent loads in flight per thread
            double v[32];
#pragma unroll
            for (int q = 0; q < 32; q++) v[q] = live ? __ldcs(row + (long long) (c + q) * lda) : 0.0;
            __syncwarp();                               // scheduling fence: all 32 loads are issued before the first FMA
#pragma unroll
            for (int q = 0; q < 32; q++) acc += v[q] * xs[c + q];
        }
        for (; c + 8 <= w; c += 8) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = live ? __ldcs(row + (long long) (c + q) * lda) : 0.0;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; q++) acc += v[q] * xs[c + q];
        }
        for (; c < w; c++) acc += (live ? __ldcs(row + (long long) c * lda) : 0.0) * xs[c];
        if (live) red_add_f64(x + xrow, -acc);
    }
}

// backward, step (a): x1 -= L2^T * X[rows below].  One CTA = SOLVE_ROWS rows of one job; a warp takes 4 columns at a
// time (16 loads in flight per lane), reduces with shuffles and issues one atomic per column.
__global__ void __launch_bounds__(SOLVE_THREADS) ltsolve_update_kernel(const SolveJob *__restrict__ jobs, const int *__restrict__ tile_job,
                                                                      const double *__restrict__ Lx, const int *__restrict__ ls,
                                                                      double *__restrict__ X, int nrhs, long long ldx)
{
    __shared__ double xr[SOLVE_ROWS];
    const SolveJob job = jobs[tile_job[blockIdx.x]];
    const int tile = (int) blockIdx.x - job.tile_start;
    const int w = job.w, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = tile * SOLVE_ROWS;
    const int nr = min(job.rows_below, r0 + SOLVE_ROWS) - r0;
    const long long lda = job.lda;
    const double *__restrict__ L2 = Lx + job.x_off + w + r0;
    constexpr int NW = SOLVE_THREADS / 32, RPL = SOLVE_ROWS / 32;
    for (int rh = 0; rh < nrhs; rh++) {
        double *__restrict__ x = X + rh * ldx;
        __syncthreads();
        for (int r = tid; r < SOLVE_ROWS; r += SOLVE_THREADS) xr[r] = (r < nr) ? x[ls[job.ls_off + r0 + r]] : 0.0;
        __syncthreads();
        constexpr int CPW = 8;                          // columns per warp per round: CPW*RPL loads in flight per lane
        for (int c0 = warp * CPW; c0 < w; c0 += NW * CPW) {
            double acc[CPW], v[CPW][RPL];
#pragma unroll
            for (int q = 0; q < CPW; q++) {
                const double *__restrict__ col = L2 + (long long) (c0 + q) * lda;
#pragma unroll
                for (int t = 0; t < RPL; t++) {
                    const int r = lane + 32 * t;
                    v[q][t] = (c0 + q < w && r < nr) ? __ldcs(col + r) : 0.0;
                }
            }
            asm volatile("" ::: "memory");              // CPW*RPL loads in flight per lane before the first FMA
#pragma unroll
            for (int q = 0; q < CPW; q++) {
                acc[q] = 0.0;
#pragma unroll
                for (int t = 0; t < RPL; t++) acc[q] += v[q][t] * xr[lane + 32 * t];
            }
#pragma unroll
            for (int q = 0; q < CPW; q++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
                if (lane == 0 && c0 + q < w) red_add_f64(x + job.xcol0 + c0 + q, -acc[q]);
            }
        }
    }
}

// backward, step (b): x1 <- L11^{-T} x1  (W^T mat-vec for wide blocks: thread i sums W(k,i) x(k), k >= i; W is staged
// through shared memory so that the global reads stay coalesced)
__global__ void __launch_bounds__(SOLVE_THREADS) ltsolve_diag_kernel(const SolveJob *__restrict__ jobs, const double *__restrict__ Lx,
                                                                    const double *__restrict__ winv, double *__restrict__ X, int nrhs, long long ldx)
{
    constexpr int LDS = NB_INNER + 1;
    __shared__ double T[NB_INNER * LDS];
    __shared__ double xs[NB_INNER];
    __shared__ double part[SOLVE_THREADS];
    const SolveJob job = jobs[blockIdx.x];
    const int w = job.w, tid = threadIdx.x;
    if (winv && job.winv_slot >= 0) {
        const double *__restrict__ W = winv + (long long) job.winv_slot * (NB_INNER * NB_INNER);
        for (int e = tid; e < NB_INNER * NB_INNER; e += SOLVE_THREADS) T[(e & 63) * LDS + (e >> 6)] = W[e];   // T[k][i] = W(k,i)
        for (int r = 0; r < nrhs; r++) {
            double *__restrict__ x = X + r * ldx + job.xcol0;
            __syncthreads();
            if (tid < w) xs[tid] = x[tid];
            __syncthreads();
            const int i = tid & 63, half = tid >> 6;
            double acc = 0.0;
            if (i < w) {
#pragma unroll 8
                for (int k = i + half; k < w; k += 2) acc += T[k * LDS + i] * xs[k];
            }
            part[tid] = acc;
            __syncthreads();
            if (tid < w) x[tid] = part[tid] + part[tid + 64];
        }
        return;
    }
    const double *__restrict__ A = Lx + job.x_off;
    for (int e = tid; e < w * w; e += SOLVE_THREADS) {
        const int i = e % w, j = e / w;
        if (i >= j) T[i * LDS + j] = A[i + (long long) j * job.lda];
    }
    for (int r = 0; r < nrhs; r++) {
        double *__restrict__ x = X + r * ldx + job.xcol0;
        __syncthreads();
        if (tid < w) xs[tid] = x[tid];
        __syncthreads();
        for (int j = w - 1; j >= 0; j--) {
            if (tid == j) xs[j] = xs[j] / T[j * LDS + j];
            __syncthreads();
            if (tid < j) xs[tid] -= T[j * LDS + tid] * xs[j];     // L^T(tid,j) = L(j,tid)
            __syncthreads();
        }
        if (tid < w) x[tid] = xs[tid];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Wide supernodes: one launch per 256-column block step (SolveBlk).  FWD: CTA cta0 of a job is its diagonal CTA, the next
// ntiles CTAs own 32-row tiles of the rows below the block; they stage their tile with cp.async while the diagonal CTA
// solves, wait for its flag, then subtract L2 * x1 from X.  BWD: the row CTAs come first, add their part of L2' * X[rows]
// to the job's scratch and bump its counter; the diagonal CTA (last) waits for the counter, then solves with the
// transposed blocks.  Waiting is safe: CTAs are dispatched in blockIdx order and a job's producers always precede its
// consumers.  flags/counters/scratch are zeroed by the caller before every pass.
// Inside the 256x256 triangle: the four 64x64 inverses W (mat-vec), the off-diagonal panels staged through shared memory.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SB_THREADS = 256;
constexpr int SB_TLD = SB_ROWS + 1;                       // row-tile leading dimension (odd: conflict-free transposed reads)
constexpr int SB_PLD = 3 * NB_INNER + 1;                  // off-diagonal panel: up to 192 rows below a 64-column sub-block
constexpr size_t solve_blk_smem_bytes()
{
    const size_t tile = (size_t) SB_W * SB_TLD, panel = (size_t) NB_INNER * SB_PLD;
    return ((tile > panel ? tile : panel) + 2 * SB_W + SB_THREADS + 64) * sizeof(double);
}

__device__ __forceinline__ int ld_acquire_int(const int *p)
{
    int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

template <bool FWD>
__global__ void __launch_bounds__(SB_THREADS) solve_blk_kernel(const SolveBlk *__restrict__ blks, const int *__restrict__ cta_blk,
                                                              const double *__restrict__ Lx, const int *__restrict__ ls,
                                                              const double *__restrict__ winv, double *__restrict__ X, int nrhs, long long ldx,
                                                              double *__restrict__ scratch, int *__restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *buf = reinterpret_cast<double *>(smem_raw);                          // row tile [c][SB_TLD] or panel [k][SB_PLD]
    constexpr size_t BUF = ((size_t) SB_W * SB_TLD > (size_t) NB_INNER * SB_PLD) ? (size_t) SB_W * SB_TLD : (size_t) NB_INNER * SB_PLD;
    double *xs = buf + BUF;                                                      // [SB_W]
    double *xn = xs + SB_W;                                                      // [SB_W] second vector
    double *part = xn + SB_W;                                                    // [SB_THREADS]
    int *rowidx = reinterpret_cast<int *>(part + SB_THREADS);                    // [SB_ROWS]
    // CTA layout of the launch: forward [diagonal CTAs of all jobs][row CTAs], backward [row CTAs][diagonal CTAs]: every
    // diagonal solve of a step runs side by side, and a waiting CTA always has its producers in front of it
    const int nj = blks[0].njobs_step, nrow_ctas = (int) gridDim.x - nj;
    const bool is_diag = FWD ? ((int) blockIdx.x < nj) : ((int) blockIdx.x >= nrow_ctas);
    const int rowcta = FWD ? (int) blockIdx.x - nj : (int) blockIdx.x;
    const SolveBlk job = blks[is_diag ? (FWD ? (int) blockIdx.x : (int) blockIdx.x - nrow_ctas) : cta_blk[rowcta]];
    const int local = rowcta - job.cta0;                                         // row tile of the job (row CTAs only)
    const int tid = threadIdx.x, w = job.w;
    const long long lda = job.lda;
    const double *__restrict__ A = Lx + job.x_off;                              // A(i,k) = A[i + k*lda], block-local indices
    double *__restrict__ scr = scratch + (long long) job.scratch * SB_W * nrhs; // [nrhs][SB_W]
    int *flag = flags + job.scratch;
    const int nsb = (w + NB_INNER - 1) / NB_INNER;

    if (!is_diag) {
        // ---------------------------------------------------------------- row tile: rows [r0, r0+nr) below the block
        // thread = (row r, group g of 32 columns): its 32 entries are loaded into registers right away - 32 independent,
        // coalesced loads in flight per thread while (forward) the diagonal CTA is still solving
        const int tile = local;
        const int r0 = tile * SB_ROWS;
        const int nr = min(SB_ROWS, job.rows_below - r0);
        const int r = tid % SB_ROWS, g = tid / SB_ROWS;
        const double *__restrict__ L2 = A + w + r0 + (r < nr ? r : 0) + (long long) (g * 32) * lda;   // column 32g+j at L2[j*lda]
        double v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = (r < nr && g * 32 + j < w) ? __ldcs(L2 + (long long) j * lda) : 0.0;
        if (tid < SB_ROWS) rowidx[tid] = tid < nr ? ls[job.ls_off + r0 + tid] : 0;
        if (FWD) {
            if (tid == 0) { while (ld_acquire_int(flag) == 0) { } }              // the diagonal CTA has published x1
        } else {
            // transposed use: through shared memory, [c][SB_TLD]
#pragma unroll
            for (int j = 0; j < 32; j++) buf[(g * 32 + j) * SB_TLD + r] = v[j];
        }
        __syncthreads();
        for (int rh = 0; rh < nrhs; rh++) {
            double *__restrict__ x = X + rh * ldx;
            if (FWD) {
                for (int c = tid; c < w; c += SB_THREADS) xs[c] = __ldcg(scr + rh * SB_W + c);
                for (int c = w + tid; c < SB_W; c += SB_THREADS) xs[c] = 0.0;
                __syncthreads();
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < 32; j++) acc += v[j] * xs[g * 32 + j];
                part[tid] = acc;
                __syncthreads();
                if (tid < nr) {
                    double sum = 0.0;
#pragma unroll
                    for (int q = 0; q < SB_THREADS / SB_ROWS; q++) sum += part[q * SB_ROWS + tid];
                    red_add_f64(x + rowidx[tid], -sum);
                }
                __syncthreads();
            } else {
                if (tid < SB_ROWS) xs[tid] = tid < nr ? x[rowidx[tid]] : 0.0;
                __syncthreads();
                // thread = column c: dot product over the tile's rows
                if (tid < w) {
                    double acc = 0.0;
#pragma unroll 8
                    for (int rr = 0; rr < SB_ROWS; rr++) acc += buf[tid * SB_TLD + rr] * xs[rr];
                    red_add_f64(scr + rh * SB_W + tid, acc);
                }
                __syncthreads();
            }
        }
        if (!FWD) {
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicAdd(flag, 1);                                     // this tile's contribution is in the scratch
        }
        return;
    }

    // -------------------------------------------------------------------- diagonal CTA
    {
        // the triangle (and inverses) of the block that follows in this pass go to L2 now: the next step's diagonal CTA then
        // works out of L2 instead of HBM (its dependent load batches are the critical path of the whole solve)
        const int nw = FWD ? job.nxt_w : job.prv_w;
        if (nw > 0) {
            const double *__restrict__ An = FWD ? A + w + (long long) w * lda : A - SB_W - (long long) SB_W * lda;
            for (int idx = tid; idx < nw * (nw / 16 + 1); idx += SB_THREADS) {   // one 128-byte line per 16 rows of each column's lower part
                const int c = idx / (nw / 16 + 1), rb = idx % (nw / 16 + 1);
                const int i = c + 16 * rb;
                if (i < nw) asm volatile("prefetch.global.L2 [%0];" ::"l"(An + i + (long long) c * lda));
            }
            const double *__restrict__ Wn = winv + (long long) (FWD ? job.slot[0] + 4 : job.slot[0] - 4) * (NB_INNER * NB_INNER);
            const int nwb = (nw + NB_INNER - 1) / NB_INNER;
            for (int idx = tid; idx < nwb * (NB_INNER * NB_INNER / 16); idx += SB_THREADS) asm volatile("prefetch.global.L2 [%0];" ::"l"(Wn + 16 * idx));
        }
    }
    if (!FWD) {
        if (tid == 0) { while (ld_acquire_int(flag) < job.ntiles) { } }          // every row tile has contributed
        __syncthreads();
    }
    // stage the off-diagonal panel below sub-block sb: rows [64(sb+1), w) x 64 columns -> buf[k][i]
    auto stage_panel = [&](int sb) {
        const int i0 = NB_INNER * (sb + 1), ni = w - i0, k0 = NB_INNER * sb;
        const int wk = min(NB_INNER, w - k0);
        const int tot = wk * ni;
        // batches of 16 coalesced loads per thread (consecutive threads = consecutive rows), then the stores
        for (int base = 0; base < tot; base += SB_THREADS * 16) {
            double q[16];
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int idx = base + u * SB_THREADS + tid;
                q[u] = idx < tot ? A[(i0 + idx % ni) + (long long) (k0 + idx / ni) * lda] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int idx = base + u * SB_THREADS + tid;
                if (idx < tot) buf[(idx / ni) * SB_PLD + idx % ni] = q[u];
            }
        }
    };
    for (int rh = 0; rh < nrhs; rh++) {
        double *__restrict__ x = X + rh * ldx + job.xcol0;
        __syncthreads();
        for (int c = tid; c < SB_W; c += SB_THREADS) xs[c] = c < w ? (FWD ? x[c] : x[c] - __ldcg(scr + rh * SB_W + c)) : 0.0;
        __syncthreads();
        if (FWD) {
            for (int sb = 0; sb < nsb; sb++) {
                const int k0 = NB_INNER * sb, wk = min(NB_INNER, w - k0);
                const bool below = sb + 1 < nsb;
                if (below) stage_panel(sb);                                      // its loads are in flight while the inverse is applied
                // x_sb <- W_sb x_sb: thread = (row i, quarter h): W(i,k) at W[i + 64k], k <= i
                const double *__restrict__ W = winv + (long long) job.slot[sb] * (NB_INNER * NB_INNER);
                {
                    const int i = tid & 63, h = tid >> 6;
                    double acc = 0.0, wv[16];
#pragma unroll
                    for (int u = 0; u < 16; u++) { const int k = h + 4 * u; wv[u] = (i < wk && k <= i) ? W[i + NB_INNER * k] : 0.0; }   // one batch of loads
#pragma unroll
                    for (int u = 0; u < 16; u++) acc += wv[u] * xs[k0 + h + 4 * u];
                    part[tid] = acc;
                }
                __syncthreads();
                if (tid < wk) xs[k0 + tid] = part[tid] + part[tid + 64] + part[tid + 128] + part[tid + 192];
                __syncthreads();
                if (below) {
                    // x_rest -= P x_sb: thread = row i of the rest
                    const int i0 = k0 + NB_INNER, ni = w - i0;
                    if (tid < ni) {
                        double acc = 0.0;
#pragma unroll 8
                        for (int k = 0; k < wk; k++) acc += buf[k * SB_PLD + tid] * xs[k0 + k];
                        xs[i0 + tid] -= acc;
                    }
                    __syncthreads();
                }
            }
            for (int c = tid; c < w; c += SB_THREADS) { x[c] = xs[c]; scr[rh * SB_W + c] = xs[c]; }
        } else {
            for (int sb = nsb - 1; sb >= 0; sb--) {
                const int k0 = NB_INNER * sb, wk = min(NB_INNER, w - k0);
                if (sb + 1 < nsb) {
                    stage_panel(sb);
                    __syncthreads();
                    // x_sb -= P' x_rest: thread = (column k, quarter h) over the rows of the rest
                    const int i0 = k0 + NB_INNER, ni = w - i0;
                    const int k = tid & 63, h = tid >> 6;
                    double acc = 0.0;
                    if (k < wk) {
#pragma unroll 4
                        for (int i = h; i < ni; i += 4) acc += buf[k * SB_PLD + i] * xs[i0 + i];
                    }
                    part[tid] = acc;
                    __syncthreads();
                    if (tid < wk) xs[k0 + tid] -= part[tid] + part[tid + 64] + part[tid + 128] + part[tid + 192];
                    __syncthreads();
                }
                // x_sb <- W_sb' x_sb: sum over k >= i of W(k,i) x(k).  W(k,i) sits at W[k + 64 i]: staged through shared memory
                // (coalesced reads, then conflict-free transposed use: buf[i*65 + k])
                const double *__restrict__ W = winv + (long long) job.slot[sb] * (NB_INNER * NB_INNER);
                {
                    double q[16];
#pragma unroll
                    for (int u = 0; u < 16; u++) q[u] = W[u * SB_THREADS + tid];
#pragma unroll
                    for (int u = 0; u < 16; u++) { const int idx = u * SB_THREADS + tid; buf[(idx >> 6) * (NB_INNER + 1) + (idx & 63)] = q[u]; }
                }
                __syncthreads();
                {
                    const int i = tid & 63, h = tid >> 6;
                    double acc = 0.0;
                    if (i < wk) {
#pragma unroll 4
                        for (int k = i + h; k < wk; k += 4) acc += buf[i * (NB_INNER + 1) + k] * xs[k0 + k];
                    }
                    part[tid] = acc;
                }
                __syncthreads();
                if (tid < wk) xn[tid] = part[tid] + part[tid + 64] + part[tid + 128] + part[tid + 192];
                __syncthreads();
                if (tid < wk) xs[k0 + tid] = xn[tid];
                __syncthreads();
            }
            for (int c = tid; c < w; c += SB_THREADS) x[c] = xs[c];
        }
    }
    if (FWD) {
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicExch(flag, 1);                                        // x1 of every right-hand side is in the scratch
    }
}

}  // namespace ssb
