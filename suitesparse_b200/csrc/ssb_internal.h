// ssb_internal.h — shared between the host plan builder (ssb_plan.cpp) and the CUDA side (ssb_cuda.cu).
// Vocabulary follows CHOLMOD: supernode s, descendant d, panel, update, Ls (row indices), Lx (values).
#pragma once
#include <cstdint>
#include <vector>
#include <string>

namespace ssb {

// Blocking of one supernode's dense factorization (potrf of the diagonal block + trsm of the rows below,
// t_cholmod_super_numeric.c:864,997), done as ONE right-looking blocked Cholesky of the tall nsrow x nscol block:
//   three nested panel widths 64 / 256 / 1024: the K = 64 update stays inside the 256-wide panel, the K = 256 update
//   inside the 1024-wide panel, and only the K = 1024 update touches the whole trailing block, so that the bulk of the
//   flops runs in GEMMs whose C tile is re-read once per 1024 columns
//   inner panel: potrf of the 64 x 64 block (+ its inverse), trsm of the rows below, K = 64 update
constexpr int NB_OUTER = 1024;     // outermost panel: its trailing update runs with K = 1024 (C re-read once per 1024 columns)
constexpr int NB_MID = 256;        // middle panel: K = 256 update, confined to the outer panel; also the panel width of the
                                   // panel-cyclic distribution over GPUs and the granularity of the device-to-host streaming
constexpr int NB_INNER = 64;
constexpr int TRSM_ROWS = 128;      // rows per trsm tile

// One C -= A * B^T job: A,B are row ranges of the SAME column-major panel (descendant d or a factor panel).
//   C(i,j) = sum_k P(i,k) P(j,k),  i < nd2, j < nd1, i >= j (lower part only)
//   target  Lx[c_off + map(i) + map(j)*ldc] -= C(i,j), map = relmap[map_off + .] or identity when map_off < 0
struct GemmJob {
    long long a_off;     // offset in Lx of P(0,0)
    long long c_off;     // offset in Lx of the target block's (0,0)
    long long map_off;   // offset in relmap, or -1 for the identity map
    int lda, ldc;
    int K, nd1, nd2;
    int tile_start;      // first tile of this job inside its launch
    int nti, ntj;        // tile rows / tile columns (tile column tj holds tiles ti = tj .. nti-1)
    int atomic;          // 1: several jobs of the launch may hit the same target entries -> red.add
    int c_col0;          // mapped jobs: column index (inside the target supernode) that c_off stands for; the relative map's
                         // column positions are taken relative to it (a job cut to one 256-column panel of a cyclic supernode
                         // addresses that panel only, wherever the panel is stored)
};

// One panel job of the in-supernode factorization.
struct PanelJob {
    long long x_off;     // offset in Lx of the nb x nb diagonal block of this step
    int lda;
    int w;               // columns of this step (<= NB_INNER)
    int rows_below;      // rows under the diagonal block (trsm rows)
    int col0;            // first column of this step inside its supernode (for info)
    int snode;           // supernode index (for info)
    int tile_start;      // trsm: first tile of this job inside its launch
    int winv_slot;       // >= 0: potrf also writes the inverse of the diagonal block to winv[slot]; trsm_tc and the solves read it
    int pad;
};

// One <=64-column slice of a supernode for the triangular solves: its diagonal block and every row of the supernode
// below it (rows inside the diagonal part are addressed through Ls as well: Ls[psi+r] = k1+r for r < nscol).
struct SolveJob {
    long long x_off;     // offset in Lx of the diagonal block
    long long ls_off;    // offset in Ls of the first row below the block
    int lda, w, rows_below;
    int xcol0;           // first column (= row of X) of the block
    int tile_start;
    int winv_slot;       // >= 0: the inverse of the diagonal block is in winv[slot] (diagonal solve = 64x64 mat-vec)
};
constexpr int SOLVE_ROWS = 128;     // rows per solve-update tile
// Wide supernodes (more than 64 columns) are solved in 256-column blocks by ONE launch per step: the block's diagonal CTA
// solves with the four 64x64 inverses and the off-diagonal blocks inside the 256x256 triangle, the row-tile CTAs of the
// same launch prefetch their 32-row tiles of the rows below meanwhile and wait for its flag (forward), or run first and
// are counted by it (backward).  4x fewer dependent steps than the 64-column schedule, and no second launch per step.
constexpr int SB_W = 256, SB_ROWS = 32;
struct SolveBlk {
    long long x_off;     // offset in Lx of the block's diagonal entry (0,0)
    long long ls_off;    // offset in Ls of the first row below the block
    int lda, w, rows_below, xcol0;
    int cta0, ntiles;    // row CTAs of this job inside its launch are [cta0, cta0+ntiles) counted after (forward) / before (backward) the
                         // launch's diagonal CTAs; the job's diagonal CTA is its index in the step
    int slot[4];         // inverse slot of each 64-column sub-block
    int scratch;         // job index over the whole schedule (flag / counter / 256*nrhs doubles of scratch)
    int nxt_w;           // columns of the following 256-block of the same supernode (0: none) - its triangle is prefetched into L2
    int prv_w;           // columns of the preceding block (0: none), for the backward pass
    int njobs_step;      // jobs in this launch: the diagonal CTAs of ALL jobs come first (forward) / last (backward)
};
struct SolveStep { long long job0; int njobs; long long tile0; int ntiles; int level = 0; int sync = 0;
                   long long blk0 = 0; int nblk = 0; long long cta0 = 0; int nctas = 0; };   // + the wide (256-column) jobs of the step   // sync: multi-GPU solve - a step above the subtree cut (all devices pass it in lock step)

enum LaunchKind : int { L_GEMM_BIG = 0, L_GEMM_SMALL = 1, L_POTRF = 2, L_TRSM = 3, L_TRSM_TC = 4, L_NKINDS = 5,
                        L_SYNC = 5 };   // L_SYNC: no kernel, only the launch's wait (joins the panel stream into the main stream)
constexpr int TRSM_TC_MIN_W = 33;   // panels wider than 32 columns use the tensor-core trsm (inverse of the diagonal block)

struct Launch {
    int kind;
    int phase;           // 0 = descendant update, 1 = factor (stats only)
    long long job0;      // first job in the flat job array of that kind
    int njobs;
    long long tile0;     // first entry in the flat tile->job array of that kind (gemm, trsm)
    int ntiles;
    double flops;        // algorithmic flops of the launch
    // Look-ahead schedule: stream 0 = main (descendant updates, big trailing updates), stream 1 = panel stream (the
    // latency-bound potrf / trsm / small-K chain of the NEXT outer panel, high priority).  wait_ev: event the launch's
    // stream waits for before the launch (-1 none); rec_ev: event recorded on its stream after the launch (-1 none).
    // The launch list stays a valid serial order: with look-ahead off everything runs on one stream in list order.
    int stream = 0, wait_ev = -1, rec_ev = -1;
};

// Device-to-host streaming of the factor: the Lx range [off, off+cnt) is final once launch `after_launch` has run.
struct CopyTask { int after_launch; long long off, cnt; };
constexpr int COPY_MAX_PANEL_SNODES = 40;  // at most this many supernodes (the ones nearest the root) stream panel by panel

// One step of the distributed schedule.  Every rank walks the same step list (same broadcasts); only the launches differ.
//   [wait for every outstanding broadcast, if wait_remote]  launches [begin, mid)  [start the broadcast, asynchronously]
//   launches [mid, end)
// The post-range is the look-ahead: work that does not depend on the broadcast just started (trailing updates with the
// previous panel) runs while the next panel travels.
struct DistStep { int launch_begin, launch_mid, launch_end; int bcast_src; long long off, cnt; int wait_remote; };

struct Update { int d, s; int p0, nd1, nd2; long long map_off; };

struct HostPlan {
    long long n = 0, nsuper = 0, ssize = 0, xsize = 0;
    std::vector<int> super;          // nsuper+1
    std::vector<long long> pi, px;   // nsuper+1
    std::vector<int> ls;             // ssize (row indices, < 2^31)
    std::vector<int> supermap;       // n
    std::vector<int> winv_base;      // per supernode: first inverse slot of its blocks, or -1
    std::vector<int> level;          // etree level of every supernode
    std::vector<int> parent;
    int nlevels = 0;
    std::vector<Update> updates;     // all (d,s) pairs, grouped by level of s
    long long relmap_size = 0;
    // flat device-bound arrays
    std::vector<GemmJob> gemm_jobs;
    std::vector<int> gemm_tiles;     // tile -> job index relative to the launch's job0
    std::vector<PanelJob> potrf_jobs;
    std::vector<PanelJob> trsm_jobs;
    std::vector<int> trsm_tiles;
    int max_winv_slots = 0;          // number of inverse-diagonal-block slots (one per wide 64-column block, kept for the solves)
    std::vector<Launch> launches;    // in execution order
    std::vector<int> level_launch_begin; // nlevels+1
    std::vector<CopyTask> copy_tasks;    // sorted by after_launch
    int n_events = 0;                    // cross-stream events of the look-ahead schedule
    // ---- elimination-tree shard over several GPUs (one process per GPU) -------------------------------------
    int nranks = 1, rank = 0;
    std::vector<int> owner;              // per supernode: rank that computes it, or < 0: its 256-column panels are cyclic over the
                                         // ranks, panel J belongs to rank (J + offset) mod nranks with offset = -1 - owner
    std::vector<DistStep> steps;         // launches [begin,end) of this rank, then an optional broadcast of a finished Lx range
    double my_flops = 0;                 // dense flops this rank executes
    // ---- distributed storage (multi-GPU inside one process): a rank stores only the supernodes it owns, the panel-cyclic
    // supernodes (all of them: it computes some of their panels and reads the others) and the remote supernodes its
    // updates read.  Present supernodes get consecutive local offsets in index order; every job offset is relocated.
    bool compact = false;
    std::vector<long long> lpx;          // local offset of supernode t in this rank's storage, -1 = not stored here
    long long lxsize = 0;                // doubles of local factor storage
    struct Piece { long long home_off, cnt; int ncols = 1; long long src_ld = 0; };   // ncols > 1: ncols segments of cnt doubles, src_ld apart at the
                                                                                  // source (home offsets), back to back at the destination
    std::vector<std::vector<Piece>> step_recv;   // per step: the parts of the step's finished range this rank reads (home offsets)
    std::vector<std::vector<int>> step_deps;     // per step: the earlier steps whose (remote) finished ranges its launches read
    std::vector<int> step_next;          // per step: for a finished panel of a cyclic supernode, the owner of the next panel (-1 none)
    int top_min_level = 0;               // lowest etree level that holds a supernode above the subtree cut
    // A panel-cyclic ROOT supernode is read by nobody but its own trailing updates: a rank keeps only the panels it owns
    // (packed) and receives the others into a ring of `ring_depth` panel slots - a received panel is dead after the one step
    // that applies it.  This is what lets a factor larger than one GPU's HBM be factorized (the root is the largest block).
    // A remote supernode that this rank's updates read is stored from its first needed row on (rmin): its trailing rows only,
    // with leading dimension nsrow - rmin (the updates of an ancestor read rows p0.. of ALL columns of a descendant).
    std::vector<int> rmin;               // per supernode: first stored row (0: the whole block)
    std::vector<char> transient;         // per supernode
    std::vector<long long> tr_own_base, tr_ring_base;   // local offsets of the packed own panels / of the ring (transient supernodes)
    int ring_depth = 4;
    long long local_of(long long home) const;    // home offset in Lx -> offset in this rank's storage (-1: not stored here)
    // solve schedule: supernodes ordered by level
    std::vector<int> level_ptr;      // nlevels+1
    std::vector<int> level_nodes;    // supernodes sorted by level
    std::vector<SolveJob> solve_jobs;
    std::vector<int> solve_tiles;
    std::vector<SolveStep> solve_steps;   // forward order; the backward solve walks them in reverse
    // second schedule, used when the inverses of the diagonal blocks are valid: supernodes up to 64 columns as above,
    // wider ones in 256-column blocks (SolveBlk)
    std::vector<SolveJob> solve2_jobs; std::vector<int> solve2_tiles;
    std::vector<SolveBlk> solve_blks; std::vector<int> solve_blk_ctas;   // CTA -> job index relative to the step's blk0
    std::vector<SolveStep> solve2_steps;
    double flops_update = 0, flops_potrf = 0, flops_trsm = 0;
    double bytes_update_panel = 0, bytes_update_scatter = 0;
    std::string error;
};

// Builds everything above from the symbolic factor.  Returns false (plan.error set) on invalid structure.
bool build_host_plan(long long n, long long nsuper, const long long *super, const long long *pi, const long long *px,
                     const long long *s, int nranks, int rank, HostPlan &plan, bool compact = false);

// Job lists for factorizing ONE supernode restricted to its first ncol_limit columns (not-positive-definite repeat,
// t_cholmod_super_numeric.c:944-967).  Appends launches to `out`.
// slot of the inverse of the 64-column diagonal block starting at column j0 of supernode s (-1: none kept)
int winv_slot_of(const HostPlan &hp, int s, int j0);

// lookahead (whole supernodes only): the factorization of outer panel O+1 runs on the panel stream while the main stream
// applies outer panel O to everything behind panel O+1; after_ev = event the first panel has to wait for (-1 none)
void append_factor_jobs(const HostPlan &hp, const std::vector<int> &snodes, int ncol_limit, HostPlan &out, bool panel_copies = false,
                        int only_panel_J0 = -1, bool lookahead = false, int after_ev = -1);

int cyc_offset(const HostPlan &hp, int sn);
int panel_owner(const HostPlan &hp, int sn, int J);
int gemm_tile_size(int kind);       // 128 for L_GEMM_BIG, 64 for L_GEMM_SMALL

}  // namespace ssb
