// ssb_plan.cpp — host-side plan builder: everything the reference discovers serially at run time with its
// Head/Next/Lpos linked lists (t_cholmod_super_numeric.c:442-460,583-611,787-808,1021-1034) is a pure function of
// the symbolic factor (super/pi/px/s), so it is computed once here and turned into flat, batched GPU work lists:
//   * every (descendant d, ancestor s) update = a maximal run of d's sub-diagonal rows inside one ancestor
//     (same walk as cholmod_super_symbolic.c:916-941), with its relative-map slot (t_cholmod_super_numeric.c:746-750)
//   * the supernodal elimination tree parent(s) = SuperMap[Ls[psi+nscol]] (t_cholmod_super_numeric.c:1025) and its levels:
//     supernodes of one level are independent, so a level is factorized by a handful of batched launches
//   * per level: C -= A*B^T tile lists for the descendant updates, then the blocked potrf/trsm/update steps.
#include "ssb_internal.h"
#include <algorithm>
#include <numeric>
#include <cstring>
#include <cstdlib>
#include <map>
#include <climits>

namespace ssb {

int gemm_tile_size(int kind) { return kind == L_GEMM_BIG ? 128 : 64; }

// A supernode keeps the inverses of its 64x64 diagonal blocks if it is wider than 32 columns and has rows below; block b
// (columns [64b, 64b+w)) qualifies if w >= TRSM_TC_MIN_W and rows remain below it.
int winv_slot_of(const HostPlan &hp, int s, int j0)
{
    if (hp.winv_base.empty() || hp.winv_base[s] < 0) return -1;
    const int nscol = hp.super[s + 1] - hp.super[s];
    const int nsrow = (int) (hp.pi[s + 1] - hp.pi[s]);
    const int w = std::min(NB_INNER, nscol - j0);
    if (nscol > NB_INNER) return hp.winv_base[s] + j0 / NB_INNER;       // wide supernode: every block (the 256-column solve needs them all)
    if (w < TRSM_TC_MIN_W || nsrow - j0 - w <= 0) return -1;
    return hp.winv_base[s] + j0 / NB_INNER;
}

long long HostPlan::local_of(long long home) const
{
    const int t = (int) (std::upper_bound(px.begin(), px.end(), home) - px.begin()) - 1;
    if (t < 0 || t >= (int) nsuper || lpx[t] < 0) return -1;
    if (!rmin.empty() && rmin[t] > 0) {
        const long long nsrow = pi[t + 1] - pi[t];
        const long long col = (home - px[t]) / nsrow, row = (home - px[t]) % nsrow;
        if (row < rmin[t]) return -1;
        return lpx[t] + col * (nsrow - rmin[t]) + (row - rmin[t]);
    }
    if (transient.empty() || !transient[t]) return lpx[t] + (home - px[t]);
    const long long nsrow = pi[t + 1] - pi[t];
    const long long col = (home - px[t]) / nsrow, row = (home - px[t]) % nsrow;
    const int J = (int) (col / NB_MID);
    const long long within = (col - (long long) J * NB_MID) * nsrow + row;
    const long long P = (long long) NB_MID * nsrow;
    const int off = -1 - owner[t];
    if ((J + off) % nranks == rank) {
        const int Jf = ((rank - off) % nranks + nranks) % nranks;       // this rank's first panel; its panels are Jf, Jf+nranks, ...
        return tr_own_base[t] + (long long) ((J - Jf) / nranks) * P + within;
    }
    return tr_ring_base[t] + (long long) (J % ring_depth) * P + within;
}

// Panel-cyclic supernodes: owner[sn] = -1 - offset, panel J belongs to rank (J + offset) mod nranks
int cyc_offset(const HostPlan &hp, int sn) { return -1 - hp.owner[sn]; }
int panel_owner(const HostPlan &hp, int sn, int J) { return (J + cyc_offset(hp, sn)) % hp.nranks; }

namespace {

// Append one launch of gemm jobs (already final except tile_start) of a given kind; builds the tile->job array.
void emit_gemm_launch(HostPlan &hp, std::vector<GemmJob> &jobs, int kind, int phase)
{
    if (jobs.empty()) return;
    const int T = gemm_tile_size(kind);
    // heaviest jobs first: CTAs are dispatched in blockIdx order, so this is longest-processing-time-first
    std::stable_sort(jobs.begin(), jobs.end(), [](const GemmJob &a, const GemmJob &b) {
        return (double) a.K * a.nd1 * a.nd2 > (double) b.K * b.nd1 * b.nd2;
    });
    size_t pos = 0;
    while (pos < jobs.size()) {
        // keep every launch's tile count inside int range
        Launch L{};
        L.kind = kind; L.phase = phase;
        L.job0 = (long long) hp.gemm_jobs.size();
        L.tile0 = (long long) hp.gemm_tiles.size();
        long long ntiles = 0;
        int nj = 0;
        while (pos < jobs.size()) {
            GemmJob j = jobs[pos];
            j.nti = (j.nd2 + T - 1) / T;
            j.ntj = (j.nd1 + T - 1) / T;
            long long t = 0;
            for (int tj = 0; tj < j.ntj; tj++) t += j.nti - tj;
            if (ntiles + t > 1500000000LL && nj > 0) break;
            j.tile_start = (int) ntiles;
            L.flops += 2.0 * j.K * ((double) j.nd1 * j.nd2 - 0.5 * (double) j.nd1 * (j.nd1 - 1));
            hp.gemm_jobs.push_back(j);
            for (long long q = 0; q < t; q++) hp.gemm_tiles.push_back(nj);
            ntiles += t; nj++; pos++;
        }
        L.njobs = nj; L.ntiles = (int) ntiles;
        hp.launches.push_back(L);
    }
    jobs.clear();
}

void emit_update_launches(HostPlan &hp, std::vector<GemmJob> &small, std::vector<GemmJob> &big, int phase)
{
    emit_gemm_launch(hp, big, L_GEMM_BIG, phase);
    emit_gemm_launch(hp, small, L_GEMM_SMALL, phase);
}

inline void route_gemm(const GemmJob &j, std::vector<GemmJob> &small, std::vector<GemmJob> &big)
{
    if (j.nd2 <= 64) small.push_back(j); else big.push_back(j);
}

}  // namespace

// Blocked factorization steps for a set of mutually independent supernodes (one etree level, or one repeated
// supernode with ncol_limit >= 0).  Appends to out.{potrf_jobs,trsm_jobs,trsm_tiles,gemm_jobs,gemm_tiles,launches}.
void append_factor_jobs(const HostPlan &hp, const std::vector<int> &snodes, int ncol_limit, HostPlan &out, bool panel_copies,
                        int only_panel_J0, bool lookahead, int after_ev)
{
    int maxcol = 0;
    for (int s : snodes) {
        int nscol = hp.super[s + 1] - hp.super[s];
        if (ncol_limit >= 0) nscol = std::min(nscol, ncol_limit);
        maxcol = std::max(maxcol, nscol);
    }
    int NB_OUTER = ssb::NB_OUTER;           // tests shrink the outer panel so that small meshes exercise the look-ahead schedule
    if (const char *e = getenv("SSB200_NB_OUTER")) { const int v = atoi(e); if (v >= NB_MID && v % NB_MID == 0) NB_OUTER = v; }
    if (ncol_limit >= 0 || only_panel_J0 >= 0 || maxcol <= NB_OUTER) lookahead = false;   // nothing to overlap
    std::vector<GemmJob> gs, gb;
    // trailing update of one supernode with the finished panel [p0, p0+W) (K = W): columns [c0, c1) of the tall block, rows
    // from c0 down (c0 >= p0+W)
    auto trailing_cols = [&](int s, int p0, int W, int c0, int c1) {
        const int nsrow = (int) (hp.pi[s + 1] - hp.pi[s]);
        if (c1 - c0 <= 0) return;
        GemmJob g{};
        g.a_off = hp.px[s] + c0 + (long long) p0 * nsrow;
        g.c_off = hp.px[s] + c0 + (long long) c0 * nsrow;
        g.map_off = -1; g.lda = nsrow; g.ldc = nsrow; g.K = W; g.nd1 = c1 - c0; g.nd2 = nsrow - c0; g.atomic = 1;
        route_gemm(g, gs, gb);
    };
    auto trailing = [&](int s, int p0, int W, int c1) { trailing_cols(s, p0, W, p0 + W, c1); };
    auto ncols = [&](int s) { const int c = hp.super[s + 1] - hp.super[s]; return ncol_limit >= 0 ? std::min(c, ncol_limit) : c; };
    const int O_first = only_panel_J0 >= 0 ? (only_panel_J0 / NB_OUTER) * NB_OUTER : 0;
    int ev_before_panel = after_ev;         // look-ahead: what the panel stream waits for before it touches the next outer panel
    int ev_last_panel = -1;
    for (int O0 = O_first; O0 < maxcol; O0 += NB_OUTER) {
        const int O1 = std::min(O0 + NB_OUTER, maxcol);
        const size_t panel_first = out.launches.size();
        for (int M0 = (only_panel_J0 >= 0 ? only_panel_J0 : O0); M0 < O1; M0 += NB_MID) {
            const int M1 = std::min(M0 + NB_MID, maxcol);
            for (int j0 = M0; j0 < M1; j0 += NB_INNER) {
                Launch LP{}; LP.kind = L_POTRF; LP.phase = 1; LP.job0 = (long long) out.potrf_jobs.size();
                // rows-below jobs: narrow panels go to the substitution kernel, wide ones to the tensor-core kernel.  Both
                // job lists live in trsm_jobs (substitution jobs of this step first), each with its own tile->job array.
                std::vector<PanelJob> sub_jobs, tc_jobs;
                for (int s : snodes) {
                    const int nscol = ncols(s);
                    if (nscol <= j0) continue;
                    const int nsrow = (int) (hp.pi[s + 1] - hp.pi[s]);
                    const int w = std::min(NB_INNER, nscol - j0);
                    PanelJob pj{};
                    pj.x_off = hp.px[s] + j0 + (long long) j0 * nsrow;
                    pj.lda = nsrow; pj.w = w; pj.rows_below = nsrow - j0 - w; pj.col0 = j0; pj.snode = s; pj.tile_start = 0;
                    pj.winv_slot = winv_slot_of(hp, s, j0);
                    out.potrf_jobs.push_back(pj);
                    LP.njobs++; LP.flops += (double) w * w * w / 3.0;
                    if (pj.rows_below > 0) {
                        (pj.winv_slot >= 0 ? tc_jobs : sub_jobs).push_back(pj);
                        trailing(s, j0, w, std::min(M0 + NB_MID, nscol));      // K = 64 update inside the middle panel
                    }
                }
                auto emit_trsm = [&](std::vector<PanelJob> &jobs, int kind) {
                    Launch LT{}; LT.kind = kind; LT.phase = 1; LT.job0 = (long long) out.trsm_jobs.size();
                    LT.tile0 = (long long) out.trsm_tiles.size();
                    long long ttiles = 0;
                    for (PanelJob &pj : jobs) {
                        pj.tile_start = (int) ttiles;
                        const int nt = (pj.rows_below + TRSM_ROWS - 1) / TRSM_ROWS;
                        for (int q = 0; q < nt; q++) out.trsm_tiles.push_back(LT.njobs);
                        ttiles += nt;
                        out.trsm_jobs.push_back(pj);
                        LT.njobs++; LT.flops += (double) pj.w * pj.w * pj.rows_below;
                    }
                    LT.ntiles = (int) ttiles;
                    return LT;
                };
                Launch LT = emit_trsm(sub_jobs, L_TRSM);
                Launch LT2 = emit_trsm(tc_jobs, L_TRSM_TC);
                if (LP.njobs) out.launches.push_back(LP);
                if (LT.njobs) out.launches.push_back(LT);
                if (LT2.njobs) out.launches.push_back(LT2);
                emit_update_launches(out, gs, gb, 1);
            }
            // columns [M0, M0+Wm) of every active supernode are final now: they can stream to the host while the
            // trailing updates run
            if (panel_copies && !out.launches.empty()) {
                const int after = (int) out.launches.size() - 1;
                for (int s : snodes) {
                    const int nscol = ncols(s);
                    if (nscol <= M0) continue;
                    const long long nsrow = hp.pi[s + 1] - hp.pi[s];
                    const int Wm = std::min(NB_MID, nscol - M0);
                    out.copy_tasks.push_back(CopyTask{after, hp.px[s] + (long long) M0 * nsrow, (long long) Wm * nsrow});
                }
            }
            if (only_panel_J0 >= 0) return;         // distributed supernode: the caller splits the trailing update over the ranks
            // K = 256 update inside the outer panel
            for (int s : snodes) {
                const int nscol = ncols(s);
                if (nscol <= M0) continue;
                trailing(s, M0, std::min(NB_MID, nscol - M0), std::min(O0 + NB_OUTER, nscol));
            }
            emit_update_launches(out, gs, gb, 1);
        }
        if (!lookahead) {
            // K = 1024 update of everything behind the outer panel
            for (int s : snodes) {
                const int nscol = ncols(s);
                if (nscol <= O0) continue;
                trailing(s, O0, std::min(NB_OUTER, nscol - O0), nscol);
            }
            emit_update_launches(out, gs, gb, 1);
            continue;
        }
        // ---- look-ahead: this outer panel was factorized on the panel stream ------------------------------------------
        const size_t panel_end = out.launches.size();
        for (size_t t = panel_first; t < panel_end; t++) out.launches[t].stream = 1;
        if (panel_end > panel_first) {
            out.launches[panel_first].wait_ev = ev_before_panel;
            ev_last_panel = out.n_events++;
            out.launches[panel_end - 1].rec_ev = ev_last_panel;
        }
        // main stream, part 1: bring the NEXT outer panel's columns up to date (K = 1024) - the panel stream is waiting for it
        const size_t next_first = out.launches.size();
        for (int s : snodes) {
            const int nscol = ncols(s);
            if (nscol <= O0 + NB_OUTER) continue;
            trailing_cols(s, O0, NB_OUTER, O0 + NB_OUTER, std::min(O0 + 2 * NB_OUTER, nscol));
        }
        emit_update_launches(out, gs, gb, 1);
        ev_before_panel = -1;
        if (out.launches.size() > next_first) {
            out.launches[next_first].wait_ev = ev_last_panel;
            ev_before_panel = out.n_events++;
            out.launches.back().rec_ev = ev_before_panel;
        }
        // main stream, part 2: everything behind the next outer panel; overlaps the factorization of the next outer panel
        for (int s : snodes) {
            const int nscol = ncols(s);
            if (nscol <= O0 + 2 * NB_OUTER) continue;
            trailing_cols(s, O0, NB_OUTER, O0 + 2 * NB_OUTER, nscol);
        }
        emit_update_launches(out, gs, gb, 1);
    }
    if (lookahead && ev_last_panel >= 0) {
        // the main stream joins the panel stream before anything else reads these supernodes
        Launch J{}; J.kind = L_SYNC; J.phase = 1; J.stream = 0; J.wait_ev = ev_last_panel;
        out.launches.push_back(J);
    }
}

bool build_host_plan(long long n, long long nsuper, const long long *super, const long long *pi, const long long *px,
                     const long long *s, int nranks, int rank, HostPlan &hp, bool compact)
{
    hp = HostPlan();
    hp.compact = compact && nranks > 1;
    if (n < 0 || nsuper < 0 || (nsuper > 0 && (!super || !pi || !px || !s))) { hp.error = "null symbolic arrays"; return false; }
    if (n >= (1LL << 31) - 1 || nsuper >= (1LL << 31) - 1) { hp.error = "n too large for 32-bit device row indices"; return false; }
    hp.n = n; hp.nsuper = nsuper;
    hp.super.resize(nsuper + 1); hp.pi.resize(nsuper + 1); hp.px.resize(nsuper + 1);
    for (long long t = 0; t <= nsuper; t++) { hp.super[t] = (int) super[t]; hp.pi[t] = pi[t]; hp.px[t] = px[t]; }
    if (nsuper == 0) { hp.level_ptr.assign(1, 0); hp.level_launch_begin.assign(1, 0); return true; }
    if (hp.super[0] != 0 || super[nsuper] != n || pi[0] != 0) { hp.error = "super/pi do not start at 0 or end at n"; return false; }
    hp.ssize = pi[nsuper]; hp.xsize = px[nsuper];
    hp.ls.resize(hp.ssize);
    hp.supermap.assign(n, -1);
    for (long long t = 0; t < nsuper; t++) {
        const long long nscol = super[t + 1] - super[t], nsrow = pi[t + 1] - pi[t];
        if (nscol <= 0 || nsrow < nscol || nsrow >= (1LL << 31) - 1 || px[t + 1] - px[t] < nscol * nsrow) { hp.error = "inconsistent supernode sizes"; return false; }
        for (long long k = super[t]; k < super[t + 1]; k++) hp.supermap[k] = (int) t;
        long long prev = -1;
        for (long long p = pi[t]; p < pi[t + 1]; p++) {
            const long long r = s[p];
            if (r < 0 || r >= n || r <= prev) { hp.error = "row indices of a supernode must be sorted, unique and in range"; return false; }
            if (p - pi[t] < nscol && r != super[t] + (p - pi[t])) { hp.error = "leading rows of a supernode must be its own columns"; return false; }
            prev = r; hp.ls[p] = (int) r;
        }
    }
    // ---- updates, etree parent, levels --------------------------------------------------------------------
    hp.parent.assign(nsuper, -1);
    hp.level.assign(nsuper, 0);
    std::vector<Update> ups;
    for (int d = 0; d < (int) nsuper; d++) {
        const int ndcol = hp.super[d + 1] - hp.super[d];
        const long long pdi = hp.pi[d], pdend = hp.pi[d + 1];
        long long p = pdi + ndcol;
        while (p < pdend) {
            const int sa = hp.supermap[hp.ls[p]];
            if (sa <= d) { hp.error = "row index below a supernode's own columns"; return false; }
            long long q = p;
            while (q < pdend && hp.supermap[hp.ls[q]] == sa) q++;
            if (hp.parent[d] < 0) hp.parent[d] = sa;
            Update u{d, sa, (int) (p - pdi), (int) (q - p), (int) (pdend - p), 0};
            ups.push_back(u);
            p = q;
        }
    }
    for (int t = 0; t < (int) nsuper; t++)
        if (hp.parent[t] >= 0) hp.level[hp.parent[t]] = std::max(hp.level[hp.parent[t]], hp.level[t] + 1);
    hp.nlevels = 1 + *std::max_element(hp.level.begin(), hp.level.end());
    // supernodes by level (counting sort keeps index order inside a level)
    hp.level_ptr.assign(hp.nlevels + 1, 0);
    for (int t = 0; t < (int) nsuper; t++) hp.level_ptr[hp.level[t] + 1]++;
    for (int l = 0; l < hp.nlevels; l++) hp.level_ptr[l + 1] += hp.level_ptr[l];
    hp.level_nodes.resize(nsuper);
    { std::vector<int> fill(hp.level_ptr.begin(), hp.level_ptr.end() - 1);
      for (int t = 0; t < (int) nsuper; t++) hp.level_nodes[fill[hp.level[t]]++] = t; }
    // updates grouped by the level of their target
    std::stable_sort(ups.begin(), ups.end(), [&](const Update &a, const Update &b) { return hp.level[a.s] < hp.level[b.s]; });
    long long moff = 0;
    for (auto &u : ups) { u.map_off = moff; moff += u.nd2; }
    hp.relmap_size = moff;
    hp.updates.swap(ups);
    // ---- inverse-diagonal-block slots (tensor-core trsm during the factorization, mat-vec diagonal solves afterwards) --
    hp.winv_base.assign(nsuper, -1);
    {
        int slots = 0;
        for (int t = 0; t < (int) nsuper; t++) {
            const int nscol = hp.super[t + 1] - hp.super[t];
            const int nsrow = (int) (hp.pi[t + 1] - hp.pi[t]);
            if ((nscol >= TRSM_TC_MIN_W && nsrow > NB_INNER / 2) || nscol > NB_INNER) { hp.winv_base[t] = slots; slots += (nscol + NB_INNER - 1) / NB_INNER; }
        }
        hp.max_winv_slots = slots;
    }
    // ---- shard: which rank computes which supernode ---------------------------------------------------------------
    hp.nranks = std::max(1, nranks); hp.rank = rank;
    std::vector<double> sn_flops(nsuper, 0.0);          // dense flops with this supernode as the target
    for (int t = 0; t < (int) nsuper; t++) {
        const double nscol = hp.super[t + 1] - hp.super[t], nsrow = (double) (hp.pi[t + 1] - hp.pi[t]);
        sn_flops[t] = nscol * nscol * nscol / 3.0 + nscol * nscol * (nsrow - nscol);
    }
    for (const Update &u : hp.updates) {
        const double ndcol = hp.super[u.d + 1] - hp.super[u.d];
        sn_flops[u.s] += 2.0 * ndcol * ((double) u.nd1 * u.nd2 - 0.5 * (double) u.nd1 * (u.nd1 - 1));
    }
    const double dist_rate = 25e12;                     // flop/s of one GPU in the shard's cost model
    double dist_tau = 0.6e-3;                           // seconds per panel step (serial panel factorization + broadcast)
    if (const char *e = getenv("SSB200_DIST_TAU")) dist_tau = atof(e);   // tests set 0 to force sharing
    hp.owner.assign(nsuper, 0);
    std::vector<int> first_desc(nsuper);                // subtree of t = supernodes [first_desc[t], t] (postordered etree)
    if (hp.nranks > 1) {
        std::vector<double> sub(sn_flops);
        for (int t = 0; t < (int) nsuper; t++) first_desc[t] = t;
        for (int t = 0; t < (int) nsuper; t++)
            if (hp.parent[t] >= 0) { sub[hp.parent[t]] += sub[t]; first_desc[hp.parent[t]] = std::min(first_desc[hp.parent[t]], first_desc[t]); }
        // The shard treats the subtree of r as the index range [first_desc[r], r] and its Lx as one contiguous range: true
        // only for a postordered supernodal etree (cholmod_analyze.c:855 postorders when Common->postorder is TRUE).
        {
            std::vector<int> cnt(nsuper, 1);
            for (int t = 0; t < (int) nsuper; t++) {
                if (hp.parent[t] >= 0) {
                    if (hp.parent[t] <= t) { hp.error = "supernodal elimination tree is not topologically ordered"; return false; }
                    cnt[hp.parent[t]] += cnt[t];
                }
            }
            for (int t = 0; t < (int) nsuper; t++)
                if (cnt[t] != t - first_desc[t] + 1) {
                    hp.error = "supernodal elimination tree is not postordered (analyze with Common->postorder = TRUE for the multi-GPU factorization)";
                    return false;
                }
        }
        std::vector<std::vector<int>> kids(nsuper);
        std::vector<int> cand;                          // roots of the candidate subtrees
        double total = 0;
        for (int t = 0; t < (int) nsuper; t++) { if (hp.parent[t] >= 0) kids[hp.parent[t]].push_back(t); else { cand.push_back(t); total += sub[t]; } }
        std::vector<char> in_top(nsuper, 0);
        // split the heaviest subtree until there are enough, small enough subtrees to balance the ranks
        for (;;) {
            int best = -1;
            for (int c = 0; c < (int) cand.size(); c++) if (best < 0 || sub[cand[c]] > sub[cand[best]]) best = c;
            if (best < 0) break;
            const int r = cand[best];
            const bool enough = (int) cand.size() >= 4 * hp.nranks && sub[r] <= total / (4.0 * hp.nranks);
            if (enough || kids[r].empty() || sub[r] < 1e-4 * total) break;
            in_top[r] = 1;
            cand.erase(cand.begin() + best);
            for (int k : kids[r]) cand.push_back(k);
        }
        // longest-processing-time-first assignment of the subtrees
        std::sort(cand.begin(), cand.end(), [&](int a, int b) { return sub[a] > sub[b] || (sub[a] == sub[b] && a < b); });
        std::vector<double> load(hp.nranks, 0.0);
        for (int r : cand) {
            const int q = (int) (std::min_element(load.begin(), load.end()) - load.begin());
            load[q] += sub[r];
            for (int t = first_desc[r]; t <= r; t++) hp.owner[t] = q;
        }
        // Supernodes above the cut, level by level (a level is a barrier: its supernodes read the lower levels).  A top
        // supernode is either whole-owned - the top supernodes of one level then run side by side on different ranks - or
        // shared panel-cyclically by all ranks.  A panel step costs a serial panel factorization plus a broadcast
        // (tau ~ 0.6 ms), so sharing only pays for supernodes that dominate their level.  Per level, the heaviest whole
        // supernodes are turned cyclic one by one while the modelled level time
        //     sum_cyclic max(npanels * tau, flops / (N * rate))  +  max_rank(whole load) / rate
        // keeps going down.
        const double rate = dist_rate;
        const double tau = dist_tau;
        bool balance_global = false;
        if (const char *e = getenv("SSB200_DIST_BALANCE")) balance_global = atoi(e) != 0;
        int cyc_min_cols = 2 * NB_MID;                  // narrower supernodes are never shared (tests lower it)
        if (const char *e = getenv("SSB200_DIST_CYC_MIN")) cyc_min_cols = std::max(2, atoi(e));
        std::vector<std::vector<int>> top_by_level(hp.nlevels);
        for (int t = 0; t < (int) nsuper; t++) if (in_top[t]) top_by_level[hp.level[t]].push_back(t);
        for (int l = 0; l < hp.nlevels; l++) {
            auto &v = top_by_level[l];
            if (v.empty()) continue;
            std::sort(v.begin(), v.end(), [&](int a, int b) { return sn_flops[a] > sn_flops[b] || (sn_flops[a] == sn_flops[b] && a < b); });
            auto model = [&](int ncyc, std::vector<int> *assign) {
                double tc = 0;
                for (int i = 0; i < ncyc; i++) {
                    const int nscol = hp.super[v[i] + 1] - hp.super[v[i]];
                    tc += std::max(((nscol + NB_MID - 1) / NB_MID) * tau, sn_flops[v[i]] / (hp.nranks * rate));
                }
                std::vector<double> lvl(hp.nranks, 0.0);
                if (assign) assign->assign(v.size(), -1);
                for (int i = ncyc; i < (int) v.size(); i++) {
                    int q = 0;
                    if (balance_global) {
                        // the ranks are not in lock step (a step only waits for the data it reads): balance the work a rank has
                        // been given so far, subtrees included, not just this level's
                        for (int r = 1; r < hp.nranks; r++) if (load[r] + lvl[r] < load[q] + lvl[q]) q = r;
                    } else
                    for (int r = 1; r < hp.nranks; r++) if (lvl[r] < lvl[q] || (lvl[r] == lvl[q] && load[r] < load[q])) q = r;
                    lvl[q] += sn_flops[v[i]];
                    if (assign) (*assign)[i] = q;
                }
                return tc + *std::max_element(lvl.begin(), lvl.end()) / rate;
            };
            // every prefix of the flops-sorted list is tried: turning only ONE of several equally heavy supernodes cyclic
            // does not shorten the level, turning all of them does
            int ncyc = 0;
            double best = model(0, nullptr);
            for (int c = 1; c <= (int) v.size(); c++) {
                const int nscol = hp.super[v[c - 1] + 1] - hp.super[v[c - 1]];
                if (nscol < cyc_min_cols) break;
                const double t1 = model(c, nullptr);
                if (t1 < best) { best = t1; ncyc = c; }
            }
            std::vector<int> assign;
            model(ncyc, &assign);
            int ncyc_level = 0;
            for (int i = 0; i < (int) v.size(); i++) if (assign[i] < 0) ncyc_level++;
            const int stride = std::max(1, hp.nranks / std::max(1, ncyc_level));
            for (int i = 0, c = 0; i < (int) v.size(); i++) {
                // cyclic: owner = -1 - offset; the level's cyclic supernodes start their panel chains on different ranks
                hp.owner[v[i]] = assign[i] >= 0 ? assign[i] : -1 - ((c++ * stride) % hp.nranks);
                if (assign[i] >= 0) load[assign[i]] += sn_flops[v[i]];
            }
        }
        // subtree roots / narrow top supernodes: their finished Lx ranges are broadcast after their level
        for (int r : cand) first_desc[r] = -1 - first_desc[r];     // mark: negative = broadcast root, range starts at -1-value
        for (int t = 0; t < (int) nsuper; t++) if (in_top[t] && hp.owner[t] >= 0) first_desc[t] = -1 - t;
    }
    auto mine_whole = [&](int sn) { return hp.owner[sn] == hp.rank; };
    bool lookahead_enabled = true;
    if (const char *e = getenv("SSB200_LOOKAHEAD")) lookahead_enabled = atoi(e) != 0;
    // lowest etree level that holds a supernode above the subtree cut: from there on a step may read remote data
    int top_min_level = hp.nlevels;
    if (hp.nranks > 1)
        for (int t = 0; t < (int) nsuper; t++) if (hp.owner[t] < 0 || first_desc[t] == -1 - t) top_min_level = std::min(top_min_level, hp.level[t]);
    hp.top_min_level = top_min_level;
    int step_mid = -1;                                  // >= 0: launch index where the current step's post-range starts
    auto close_step = [&](int &step_begin, int src, long long off, long long cnt, int level) {
        const int end = (int) hp.launches.size();
        hp.steps.push_back(DistStep{step_begin, step_mid >= 0 ? step_mid : end, end, src, off, cnt, level >= top_min_level ? 1 : 0});
        step_begin = end; step_mid = -1;
    };
    // ---- per level launches ----------------------------------------------------------------------------------
    hp.level_launch_begin.assign(hp.nlevels + 1, 0);
    size_t upos = 0;
    std::vector<GemmJob> gs, gb;
    std::map<std::pair<int, int>, std::vector<GemmJob>> cyc_jobs;   // (panel-cyclic supernode, panel) -> descendant updates of this rank
    std::vector<int> nodes;
    std::vector<char> need(nsuper, 0);      // supernodes whose values this rank's updates read
    std::vector<int> need_row(nsuper, INT_MAX);   // ... and the first row of them that is read
    int step_begin = 0;
    // Host streaming: every copy is a host-side call between kernel launches, so only the few supernodes near the root are
    // streamed panel by panel; everything up to the `flush` level goes out in merged contiguous ranges right after that
    // level (the lowest level above which at most COPY_MAX_PANEL_SNODES supernodes remain).
    int flush = hp.nlevels - 1;
    {
        int above = 0;
        for (int l = hp.nlevels - 1; l >= 0; l--) {
            above += hp.level_ptr[l + 1] - hp.level_ptr[l];
            if (above > COPY_MAX_PANEL_SNODES) { flush = l; break; }
            flush = l - 1;
        }
        flush = std::max(0, std::min(flush, hp.nlevels - 1));
    }
    for (int l = 0; l < hp.nlevels; l++) {
        hp.level_launch_begin[l] = (int) hp.launches.size();
        size_t ubeg = upos;
        while (upos < hp.updates.size() && hp.level[hp.updates[upos].s] == l) upos++;
        for (size_t t = ubeg; t < upos; t++) {
            const Update &u = hp.updates[t];
            const int ndcol = hp.super[u.d + 1] - hp.super[u.d];
            const int ndrow = (int) (hp.pi[u.d + 1] - hp.pi[u.d]);
            const int nsrow = (int) (hp.pi[u.s + 1] - hp.pi[u.s]);
            const double tri = (double) u.nd1 * u.nd2 - 0.5 * (double) u.nd1 * (u.nd1 - 1);
            hp.flops_update += 2.0 * ndcol * tri;
            hp.bytes_update_panel += 8.0 * (double) u.nd2 * ndcol;
            hp.bytes_update_scatter += 16.0 * tri;
            GemmJob g{};
            g.a_off = hp.px[u.d] + u.p0;
            g.c_off = hp.px[u.s];
            g.map_off = u.map_off; g.lda = ndrow; g.ldc = nsrow; g.K = ndcol; g.nd1 = u.nd1; g.nd2 = u.nd2; g.atomic = 1;
            if (hp.owner[u.s] >= 0) {
                if (!mine_whole(u.s)) continue;
                hp.my_flops += 2.0 * ndcol * tri;
                need[u.d] = 1; need_row[u.d] = std::min(need_row[u.d], u.p0);
                route_gemm(g, gs, gb);
            } else {
                // panel-cyclic target: cut the update where its target column crosses a 256-column panel boundary; a cut
                // at local row jlo is itself a valid update (rows jlo.. of the descendant, columns [jlo, jhi))
                const int k1 = hp.super[u.s];
                const int *rows = hp.ls.data() + hp.pi[u.d] + u.p0;
                int jlo = 0;
                while (jlo < u.nd1) {
                    const int blk = (rows[jlo] - k1) / NB_MID;
                    int jhi = jlo + 1;
                    while (jhi < u.nd1 && (rows[jhi] - k1) / NB_MID == blk) jhi++;
                    if (panel_owner(hp, u.s, blk) == hp.rank) {
                        GemmJob h = g;
                        h.a_off += jlo; h.map_off += jlo; h.nd1 = jhi - jlo; h.nd2 = u.nd2 - jlo;
                        h.c_off = g.c_off + (long long) blk * NB_MID * nsrow; h.c_col0 = blk * NB_MID;   // the target is panel `blk`
                        hp.my_flops += 2.0 * ndcol * ((double) h.nd1 * h.nd2 - 0.5 * (double) h.nd1 * (h.nd1 - 1));
                        need[u.d] = 1; need_row[u.d] = std::min(need_row[u.d], u.p0 + jlo);
                        // not launched with the level's other updates: the descendant updates of a panel are scheduled just
                        // in time inside the panel loop, where they fill the ranks' idle time behind the serial panel chain
                        cyc_jobs[{u.s, blk}].push_back(h);
                    }
                    jlo = jhi;
                }
            }
        }
        emit_update_launches(hp, gs, gb, 0);
        nodes.clear();
        for (int t = hp.level_ptr[l]; t < hp.level_ptr[l + 1]; t++) {
            const int sn = hp.level_nodes[t];
            const double nscol = hp.super[sn + 1] - hp.super[sn], nsrow = (double) (hp.pi[sn + 1] - hp.pi[sn]);
            hp.flops_potrf += nscol * nscol * nscol / 3.0;
            hp.flops_trsm += nscol * nscol * (nsrow - nscol);
            if (hp.owner[sn] < 0 || !mine_whole(sn)) continue;
            nodes.push_back(sn);
            hp.my_flops += nscol * nscol * nscol / 3.0 + nscol * nscol * (nsrow - nscol);
        }
        if (!nodes.empty()) {
            // single GPU: look-ahead inside the level's supernodes; the panel stream first waits for the level's descendant
            // updates (everything launched so far on the main stream)
            int after_ev = -1;
            const bool la = (hp.nranks == 1 || hp.compact) && lookahead_enabled;      // also inside a rank of the in-process multi-GPU path
            if (la && !hp.launches.empty()) {
                int maxcol = 0;
                for (int sn : nodes) maxcol = std::max(maxcol, hp.super[sn + 1] - hp.super[sn]);
                {
                    Launch &last = hp.launches.back();
                    if (last.rec_ev < 0) last.rec_ev = hp.n_events++;
                    after_ev = last.rec_ev;
                }
            }
            append_factor_jobs(hp, nodes, -1, hp, /*panel_copies=*/hp.nranks == 1 && l > flush, -1, la, after_ev);
        }
        if (hp.nranks == 1 && l == flush && !hp.launches.empty()) {
            // every supernode of level <= flush is final: merge consecutive indices into contiguous Lx ranges
            const int after = (int) hp.launches.size() - 1;
            long long run_off = -1, run_end = -1;
            for (int t = 0; t < (int) nsuper; t++) {
                const bool low = hp.level[t] <= flush;
                if (low) {
                    if (run_off < 0) run_off = hp.px[t];
                    run_end = hp.px[t + 1];
                }
                if ((!low || t == (int) nsuper - 1) && run_off >= 0) {
                    hp.copy_tasks.push_back(CopyTask{after, run_off, run_end - run_off});
                    run_off = -1;
                }
            }
        }
        if (hp.nranks > 1) {
            // panel-cyclic supernodes of this level, with a look-ahead of one panel: the owner of panel J+1 first brings that
            // panel up to date with panel J and factorizes it, the broadcast of panel J+1 starts, and only then everybody
            // runs the rest of the trailing update with panel J (which overlaps the broadcast).  The level's cyclic supernodes
            // are independent: their panel chains are INTERLEAVED (panel J of every supernode, then panel J+1 ...), and their
            // panel owners are staggered (owner = (J + offset) mod nranks), so that different ranks work on different chains
            // at the same time and one chain's latency (factor + transfer) is hidden behind the others.
            struct Cyc { int sn, nscol, npan, off; long long nsrow; bool jit; };
            std::vector<Cyc> cyc;
            for (int t = hp.level_ptr[l]; t < hp.level_ptr[l + 1]; t++) {
                const int sn = hp.level_nodes[t];
                if (hp.owner[sn] >= 0) continue;
                Cyc c; c.sn = sn; c.nscol = hp.super[sn + 1] - hp.super[sn]; c.nsrow = hp.pi[sn + 1] - hp.pi[sn];
                c.npan = (c.nscol + NB_MID - 1) / NB_MID; c.off = cyc_offset(hp, sn);
                // Just-in-time descendant updates only pay when the panel chain is the bottleneck, i.e. when a rank's share of
                // the trailing update per panel is shorter than a panel step; otherwise one big up-front launch is more
                // efficient (measured: 8 GPUs 324 -> 296 ms with, 2 GPUs 670 -> 697 ms with).
                const double own_flops = (double) c.nscol * c.nscol * c.nscol / 3.0 + (double) c.nscol * c.nscol * (double) (c.nsrow - c.nscol);
                c.jit = own_flops / (hp.nranks * dist_rate * c.npan) < dist_tau;
                if (const char *e = getenv("SSB200_DIST_JIT")) c.jit = atoi(e) != 0;      // tests force either schedule
                cyc.push_back(c);
            }
            auto panel_w = [&](const Cyc &c, int J) { return std::min(NB_MID, c.nscol - J * NB_MID); };
            auto factor_panel = [&](const Cyc &c, int J) {
                const int J0 = J * NB_MID, W = panel_w(c, J);
                std::vector<int> one{c.sn};
                append_factor_jobs(hp, one, -1, hp, false, J0);
                hp.my_flops += (double) W * W * W / 3.0 + (double) W * W * (c.nsrow - J0 - W);
            };
            auto update_block = [&](const Cyc &c, int J, int J1) {       // block J1 -= panel J contribution
                const int J0 = J * NB_MID, C0 = J1 * NB_MID, W = panel_w(c, J), W1 = panel_w(c, J1);
                GemmJob g{};
                g.a_off = hp.px[c.sn] + C0 + (long long) J0 * c.nsrow;
                g.c_off = hp.px[c.sn] + C0 + (long long) C0 * c.nsrow;
                g.map_off = -1; g.lda = (int) c.nsrow; g.ldc = (int) c.nsrow; g.K = W; g.nd1 = W1; g.nd2 = (int) (c.nsrow - C0); g.atomic = 1;
                hp.my_flops += 2.0 * W * ((double) g.nd1 * g.nd2 - 0.5 * (double) g.nd1 * (g.nd1 - 1));
                route_gemm(g, gs, gb);
            };
            auto descendant_updates = [&](const Cyc &c, int J) {         // this rank's descendant updates into panel J (it owns J)
                auto it = cyc_jobs.find({c.sn, J});
                if (it == cyc_jobs.end()) return;
                for (const GemmJob &h : it->second) route_gemm(h, gs, gb);
                cyc_jobs.erase(it);
                emit_update_launches(hp, gs, gb, 0);
            };
            auto owner_of = [&](const Cyc &c, int J) { return (J + c.off) % hp.nranks; };
            // prologues: every rank brings its FIRST panel (or, without jit, all its panels) up to date with the descendants,
            // the owner of panel 0 factorizes it
            for (const Cyc &c : cyc) {
                if (c.jit) {
                    const int Jfirst = ((hp.rank - c.off) % hp.nranks + hp.nranks) % hp.nranks;
                    if (Jfirst < c.npan) descendant_updates(c, Jfirst);
                } else {
                    for (int J = 0; J < c.npan; J++) {
                        if (owner_of(c, J) != hp.rank) continue;
                        auto it = cyc_jobs.find({c.sn, J});
                        if (it == cyc_jobs.end()) continue;
                        for (const GemmJob &h : it->second) route_gemm(h, gs, gb);
                        cyc_jobs.erase(it);
                    }
                    emit_update_launches(hp, gs, gb, 0);
                }
                if (owner_of(c, 0) == hp.rank) factor_panel(c, 0);
                close_step(step_begin, owner_of(c, 0), hp.px[c.sn], (long long) panel_w(c, 0) * c.nsrow, l);
            }
            int maxpan = 0;
            for (const Cyc &c : cyc) maxpan = std::max(maxpan, c.npan);
            for (int J = 0; J < maxpan; J++)
                for (const Cyc &c : cyc) {
                    if (J >= c.npan) continue;
                    const bool own_next = (J + 1 < c.npan) && (owner_of(c, J + 1) == hp.rank);
                    if (own_next) { update_block(c, J, J + 1); emit_update_launches(hp, gs, gb, 1); factor_panel(c, J + 1); }
                    step_mid = (int) hp.launches.size();
                    for (int J1 = J + 2; J1 < c.npan; J1++) if (owner_of(c, J1) == hp.rank) update_block(c, J, J1);
                    emit_update_launches(hp, gs, gb, 1);
                    // the owner of panel J has just finished its turn in the chain: its NEXT panel (J + nranks) gets its
                    // descendant updates now, nranks-1 steps before it is needed
                    if (c.jit && owner_of(c, J) == hp.rank && J + hp.nranks < c.npan) descendant_updates(c, J + hp.nranks);
                    if (J + 1 < c.npan) close_step(step_begin, owner_of(c, J + 1), hp.px[c.sn] + (long long) (J + 1) * NB_MID * c.nsrow, (long long) panel_w(c, J + 1) * c.nsrow, l);
                    else close_step(step_begin, -1, 0, 0, l);
                }
            // finished subtrees / narrow top supernodes of this level: replicate them
            for (int t = hp.level_ptr[l]; t < hp.level_ptr[l + 1]; t++) {
                const int sn = hp.level_nodes[t];
                if (first_desc[sn] >= 0 || hp.owner[sn] < 0) continue;
                const int lo = -1 - first_desc[sn];
                close_step(step_begin, hp.owner[sn], hp.px[lo], hp.px[sn + 1] - hp.px[lo], l);
            }
        }
    }
    hp.level_launch_begin[hp.nlevels] = (int) hp.launches.size();
    close_step(step_begin, -1, 0, 0, hp.nlevels);
    if (hp.nranks == 1) hp.my_flops = hp.flops_update + hp.flops_potrf + hp.flops_trsm;
    // ---- solve schedule: per level, per 64-column block index ------------------------------------------------
    // Distributed storage: a rank solves with the blocks it factorized (its supernodes, its panels of the cyclic ones);
    // every rank walks the same step list, steps above the subtree cut are passed in lock step.
    for (int l = 0; l < hp.nlevels; l++) {
        int maxcol = 0;
        for (int t = hp.level_ptr[l]; t < hp.level_ptr[l + 1]; t++) {
            const int sn = hp.level_nodes[t];
            maxcol = std::max(maxcol, hp.super[sn + 1] - hp.super[sn]);
        }
        for (int j0 = 0; j0 < maxcol; j0 += NB_INNER) {
            SolveStep st{(long long) hp.solve_jobs.size(), 0, (long long) hp.solve_tiles.size(), 0};
            st.level = l; st.sync = (hp.compact && l >= top_min_level) ? 1 : 0;
            for (int t = hp.level_ptr[l]; t < hp.level_ptr[l + 1]; t++) {
                const int sn = hp.level_nodes[t];
                const int nscol = hp.super[sn + 1] - hp.super[sn];
                if (nscol <= j0) continue;
                if (hp.compact) {
                    const int o = hp.owner[sn];
                    if (o >= 0 ? (o != hp.rank) : (panel_owner(hp, sn, j0 / NB_MID) != hp.rank)) continue;
                }
                const int nsrow = (int) (hp.pi[sn + 1] - hp.pi[sn]);
                SolveJob sj{};
                sj.w = std::min(NB_INNER, nscol - j0);
                sj.x_off = hp.px[sn] + j0 + (long long) j0 * nsrow;
                sj.ls_off = hp.pi[sn] + j0 + sj.w;
                sj.lda = nsrow; sj.rows_below = nsrow - j0 - sj.w; sj.xcol0 = hp.super[sn] + j0;
                sj.tile_start = st.ntiles;
                sj.winv_slot = winv_slot_of(hp, sn, j0);
                const int nt = (sj.rows_below + SOLVE_ROWS - 1) / SOLVE_ROWS;
                for (int q = 0; q < nt; q++) hp.solve_tiles.push_back(st.njobs);
                st.ntiles += nt;
                hp.solve_jobs.push_back(sj);
                st.njobs++;
            }
            if (st.njobs || hp.compact) hp.solve_steps.push_back(st);
        }
    }
    // ---- second solve schedule: the big supernodes (>= blk_min columns) in fused 256-column block steps, everything else
    // in 64-column steps as in the first schedule.  Supernodes of one level are independent, so a level is: the 64-column
    // steps of its small supernodes, then the block steps of its big ones.  (Measured: the fused kernel wins where a step
    // has few jobs with many rows - the top of the tree; with thousands of jobs per step its diagonal CTA's latency does not.)
    int blk_min = 1024;
    if (const char *e = getenv("SSB200_SOLVE_BLK_MIN")) blk_min = std::max(NB_INNER + 1, atoi(e));
    for (int l = 0; l < hp.nlevels; l++) {
        int maxsmall = 0, maxbig = 0;
        for (int t = hp.level_ptr[l]; t < hp.level_ptr[l + 1]; t++) {
            const int sn = hp.level_nodes[t];
            const int nscol = hp.super[sn + 1] - hp.super[sn];
            if (nscol >= blk_min) maxbig = std::max(maxbig, nscol); else maxsmall = std::max(maxsmall, nscol);
        }
        auto mine = [&](int sn, int j0) {
            if (!hp.compact) return true;
            const int o = hp.owner[sn];
            return o >= 0 ? (o == hp.rank) : (panel_owner(hp, sn, j0 / NB_MID) == hp.rank);
        };
        for (int j0 = 0; j0 < maxsmall; j0 += NB_INNER) {
            SolveStep st{(long long) hp.solve2_jobs.size(), 0, (long long) hp.solve2_tiles.size(), 0};
            st.level = l; st.sync = (hp.compact && l >= top_min_level) ? 1 : 0;
            st.blk0 = (long long) hp.solve_blks.size(); st.cta0 = (long long) hp.solve_blk_ctas.size();
            for (int t = hp.level_ptr[l]; t < hp.level_ptr[l + 1]; t++) {
                const int sn = hp.level_nodes[t];
                const int nscol = hp.super[sn + 1] - hp.super[sn];
                if (nscol >= blk_min || nscol <= j0 || !mine(sn, j0)) continue;
                const int nsrow = (int) (hp.pi[sn + 1] - hp.pi[sn]);
                SolveJob sj{};
                sj.w = std::min(NB_INNER, nscol - j0);
                sj.x_off = hp.px[sn] + j0 + (long long) j0 * nsrow;
                sj.ls_off = hp.pi[sn] + j0 + sj.w;
                sj.lda = nsrow; sj.rows_below = nsrow - j0 - sj.w; sj.xcol0 = hp.super[sn] + j0;
                sj.tile_start = st.ntiles; sj.winv_slot = winv_slot_of(hp, sn, j0);
                const int nt = (sj.rows_below + SOLVE_ROWS - 1) / SOLVE_ROWS;
                for (int z = 0; z < nt; z++) hp.solve2_tiles.push_back(st.njobs);
                st.ntiles += nt; hp.solve2_jobs.push_back(sj); st.njobs++;
            }
            if (st.njobs || hp.compact) hp.solve2_steps.push_back(st);
        }
        for (int j0 = 0; j0 < maxbig; j0 += SB_W) {
            SolveStep st{(long long) hp.solve2_jobs.size(), 0, (long long) hp.solve2_tiles.size(), 0};
            st.level = l; st.sync = (hp.compact && l >= top_min_level) ? 1 : 0;
            st.blk0 = (long long) hp.solve_blks.size(); st.cta0 = (long long) hp.solve_blk_ctas.size();
            for (int t = hp.level_ptr[l]; t < hp.level_ptr[l + 1]; t++) {
                const int sn = hp.level_nodes[t];
                const int nscol = hp.super[sn + 1] - hp.super[sn];
                if (nscol < blk_min || nscol <= j0 || !mine(sn, j0)) continue;
                const int nsrow = (int) (hp.pi[sn + 1] - hp.pi[sn]);
                SolveBlk b{};
                b.w = std::min(SB_W, nscol - j0);
                b.x_off = hp.px[sn] + j0 + (long long) j0 * nsrow;
                b.ls_off = hp.pi[sn] + j0 + b.w;
                b.lda = nsrow; b.rows_below = nsrow - j0 - b.w; b.xcol0 = hp.super[sn] + j0;
                b.cta0 = st.nctas; b.ntiles = (b.rows_below + SB_ROWS - 1) / SB_ROWS;
                for (int z = 0; z < 4; z++) b.slot[z] = (64 * z < b.w) ? winv_slot_of(hp, sn, j0 + 64 * z) : -1;
                b.scratch = (int) hp.solve_blks.size();
                b.nxt_w = (j0 + SB_W < nscol) ? std::min(SB_W, nscol - j0 - SB_W) : 0;
                b.prv_w = j0 > 0 ? SB_W : 0;
                if (hp.compact && hp.owner[sn] < 0) b.nxt_w = b.prv_w = 0;         // the neighbouring block lives on another rank
                for (int z = 0; z < b.ntiles; z++) hp.solve_blk_ctas.push_back(st.nblk);
                st.nctas += b.ntiles; hp.solve_blks.push_back(b); st.nblk++;
            }
            // the launch = [nblk diagonal CTAs][row CTAs] forward, [row CTAs][nblk diagonal CTAs] backward
            for (int z = 0; z < st.nblk; z++) hp.solve_blks[st.blk0 + z].njobs_step = st.nblk;
            if (st.nblk || hp.compact) hp.solve2_steps.push_back(st);
        }
    }
    // ---- distributed storage: local layout, relocation of every offset, receive lists ---------------------------------
    if (hp.compact) {
        // ---- which finished ranges does every step read?  (before the offsets are relocated: home offsets identify supernodes)
        {
            std::vector<int> final_step(nsuper, -1);                       // whole supernodes: the step that finished them
            std::map<std::pair<int, int>, int> panel_step;                  // (cyclic supernode, panel) -> step
            for (size_t k = 0; k < hp.steps.size(); k++) {
                const DistStep &st = hp.steps[k];
                if (st.bcast_src < 0 || st.cnt <= 0) continue;
                const int t0 = (int) (std::upper_bound(hp.px.begin(), hp.px.end(), st.off) - hp.px.begin()) - 1;
                if (hp.owner[t0] < 0) {
                    const long long nsrow = hp.pi[t0 + 1] - hp.pi[t0];
                    panel_step[{t0, (int) ((st.off - hp.px[t0]) / ((long long) NB_MID * nsrow))}] = (int) k;
                } else
                    for (int t = t0; t < (int) nsuper && hp.px[t] < st.off + st.cnt; t++) final_step[t] = (int) k;
            }
            hp.step_deps.assign(hp.steps.size(), {});
            for (size_t k = 0; k < hp.steps.size(); k++) {
                const DistStep &st = hp.steps[k];
                std::vector<int> &deps = hp.step_deps[k];
                for (int t = st.launch_begin; t < st.launch_end; t++) {
                    const Launch &L = hp.launches[t];
                    if (L.kind != L_GEMM_BIG && L.kind != L_GEMM_SMALL) continue;
                    for (long long q = L.job0; q < L.job0 + L.njobs; q++) {
                        const GemmJob &g = hp.gemm_jobs[q];
                        const int d = (int) (std::upper_bound(hp.px.begin(), hp.px.end(), g.a_off) - hp.px.begin()) - 1;
                        if (hp.owner[d] == hp.rank) continue;
                        if (hp.owner[d] >= 0) { if (final_step[d] >= 0) deps.push_back(final_step[d]); continue; }
                        const long long nsrow = hp.pi[d + 1] - hp.pi[d];
                        const int c0 = (int) ((g.a_off - hp.px[d]) / nsrow);
                        for (int J = c0 / NB_MID; J <= (c0 + g.K - 1) / NB_MID; J++) {
                            if (panel_owner(hp, d, J) == hp.rank) continue;
                            auto it = panel_step.find({d, J});
                            if (it != panel_step.end()) deps.push_back(it->second);
                        }
                    }
                }
                std::sort(deps.begin(), deps.end());
                deps.erase(std::unique(deps.begin(), deps.end()), deps.end());
                for (int dep : deps) if (dep >= (int) k) { hp.error = "internal: a step reads a range that is finished later"; return false; }
            }
        }
        hp.lpx.assign(nsuper + 1, -1);
        hp.transient.assign(nsuper, 0); hp.tr_own_base.assign(nsuper, -1); hp.tr_ring_base.assign(nsuper, -1);
        {
            int tr_on = 1, tr_min = 2 * hp.nranks;
            if (const char *e = getenv("SSB200_MG_TRANSIENT")) tr_on = atoi(e);
            if (const char *e = getenv("SSB200_MG_TRANSIENT_MIN")) tr_min = std::max(1, atoi(e));     // panels; tests lower it
            if (const char *e = getenv("SSB200_MG_RING")) hp.ring_depth = std::max(2, atoi(e));
            for (int t = 0; t < (int) nsuper && tr_on; t++) {
                const int nscol = hp.super[t + 1] - hp.super[t];
                if (hp.owner[t] < 0 && hp.parent[t] < 0 && (nscol + NB_MID - 1) / NB_MID >= std::max(tr_min, hp.ring_depth + 1)) hp.transient[t] = 1;
            }
        }
        hp.rmin.assign(nsuper, 0);
        {
            int on = 1;
            if (const char *e = getenv("SSB200_MG_TRAILING_ROWS")) on = atoi(e);
            for (int t = 0; t < (int) nsuper && on; t++)
                if (need[t] && hp.owner[t] >= 0 && hp.owner[t] != hp.rank && need_row[t] != INT_MAX) hp.rmin[t] = need_row[t];
        }
        long long pos = 0;
        for (int t = 0; t < (int) nsuper; t++) {
            const bool present = hp.owner[t] == hp.rank || hp.owner[t] < 0 || need[t];
            if (!present) continue;
            hp.lpx[t] = pos;
            if (hp.rmin[t] > 0) { pos += (hp.pi[t + 1] - hp.pi[t] - hp.rmin[t]) * (long long) (hp.super[t + 1] - hp.super[t]); continue; }
            if (!hp.transient[t]) { pos += hp.px[t + 1] - hp.px[t]; continue; }   // no padding: a run of consecutive supernodes keeps its internal offsets on every rank
            const long long nsrow = hp.pi[t + 1] - hp.pi[t];
            const int nscol = hp.super[t + 1] - hp.super[t];
            const int npan = (nscol + NB_MID - 1) / NB_MID;
            hp.tr_own_base[t] = pos;
            for (int J = 0; J < npan; J++) if (panel_owner(hp, t, J) == hp.rank) pos += (long long) std::min(NB_MID, nscol - J * NB_MID) * nsrow;
            hp.tr_ring_base[t] = pos;
            pos += (long long) hp.ring_depth * NB_MID * nsrow;
        }
        hp.lxsize = pos; hp.lpx[nsuper] = pos;
        auto loc = [&](long long home) -> long long {
            const long long v = hp.local_of(home);
            if (v < 0) hp.error = "internal: job reads a supernode that is not stored on this rank";
            return v;
        };
        for (auto &g : hp.gemm_jobs) {
            const int d = (int) (std::upper_bound(hp.px.begin(), hp.px.end(), g.a_off) - hp.px.begin()) - 1;
            if (hp.rmin[d] > 0) g.lda = (int) (hp.pi[d + 1] - hp.pi[d] - hp.rmin[d]);       // trailing rows only: shorter columns
            g.a_off = loc(g.a_off); g.c_off = loc(g.c_off);
        }
        for (auto &j : hp.potrf_jobs) j.x_off = loc(j.x_off);
        for (auto &j : hp.trsm_jobs) j.x_off = loc(j.x_off);
        for (auto &j : hp.solve_jobs) j.x_off = loc(j.x_off);
        for (auto &j : hp.solve2_jobs) j.x_off = loc(j.x_off);
        for (auto &j : hp.solve_blks) j.x_off = loc(j.x_off);
        if (!hp.error.empty()) return false;
        // what this rank pulls out of every finished range: maximal runs of consecutive needed supernodes
        hp.step_recv.assign(hp.steps.size(), {});
        hp.step_next.assign(hp.steps.size(), -1);
        for (size_t k = 0; k < hp.steps.size(); k++) {
            const DistStep &st = hp.steps[k];
            if (st.bcast_src < 0 || st.bcast_src == hp.rank || st.cnt <= 0) continue;
            const int t0 = (int) (std::upper_bound(hp.px.begin(), hp.px.end(), st.off) - hp.px.begin()) - 1;
            if (hp.owner[t0] < 0) {                      // a panel of a cyclic supernode: always read by everybody
                hp.step_recv[k].push_back(HostPlan::Piece{st.off, st.cnt});
                // the owner of the NEXT panel is on the critical path: it pulls first
                const long long nsrow = hp.pi[t0 + 1] - hp.pi[t0];
                const int nscol = hp.super[t0 + 1] - hp.super[t0];
                const int J = (int) ((st.off - hp.px[t0]) / ((long long) NB_MID * nsrow));
                if ((J + 1) * NB_MID < nscol) hp.step_next[k] = panel_owner(hp, t0, J + 1);
                continue;
            }
            long long run_off = -1, run_end = -1;
            for (int t = t0; t < (int) nsuper && hp.px[t] < st.off + st.cnt; t++) {
                if (need[t] && hp.rmin[t] > 0) {
                    // trailing rows of every column: a strided (2-D) piece
                    if (run_off >= 0) { hp.step_recv[k].push_back(HostPlan::Piece{run_off, run_end - run_off}); run_off = -1; }
                    const long long nsrow = hp.pi[t + 1] - hp.pi[t];
                    HostPlan::Piece pc{hp.px[t] + hp.rmin[t], nsrow - hp.rmin[t]};
                    pc.ncols = hp.super[t + 1] - hp.super[t]; pc.src_ld = nsrow;
                    hp.step_recv[k].push_back(pc);
                } else if (need[t]) {
                    if (run_off < 0) run_off = hp.px[t];
                    run_end = hp.px[t + 1];
                } else if (run_off >= 0) { hp.step_recv[k].push_back(HostPlan::Piece{run_off, run_end - run_off}); run_off = -1; }
            }
            if (run_off >= 0) hp.step_recv[k].push_back(HostPlan::Piece{run_off, run_end - run_off});
        }
    }
    return true;
}

}  // namespace ssb

// Host-only view of the plan (no GPU needed): sizes of the schedule, for tests and for DESIGN.md's tables.
//  out[0]=nlevels out[1]=nupdates out[2]=relmap_size out[3]=nlaunches out[4]=gemm jobs out[5]=gemm tiles
//  out[6]=potrf jobs out[7]=trsm jobs out[8]=trsm tiles out[9]=solve steps out[10]=solve jobs
//  out[11]=flops_update out[12]=flops_potrf out[13]=flops_trsm out[14]=bytes_update_panel out[15]=bytes_update_scatter
extern "C" int ssb200_plan_summary(long long n, long long nsuper, const long long *super, const long long *pi, const long long *px,
                                   const long long *s, double *out, int *level_of_supernode)
{
    ssb::HostPlan hp;
    if (!ssb::build_host_plan(n, nsuper, super, pi, px, s, 1, 0, hp)) return -4;
    out[0] = hp.nlevels; out[1] = (double) hp.updates.size(); out[2] = (double) hp.relmap_size; out[3] = (double) hp.launches.size();
    out[4] = (double) hp.gemm_jobs.size(); out[5] = (double) hp.gemm_tiles.size(); out[6] = (double) hp.potrf_jobs.size();
    out[7] = (double) hp.trsm_jobs.size(); out[8] = (double) hp.trsm_tiles.size(); out[9] = (double) hp.solve_steps.size();
    out[10] = (double) hp.solve_jobs.size(); out[11] = hp.flops_update; out[12] = hp.flops_potrf; out[13] = hp.flops_trsm;
    out[14] = hp.bytes_update_panel; out[15] = hp.bytes_update_scatter;
    { double c = 0; for (const auto &ct : hp.copy_tasks) c += (double) ct.cnt; out[16] = c; out[17] = (double) hp.copy_tasks.size(); }
    out[18] = hp.max_winv_slots;
    { int k = 0; for (const auto &g : hp.gemm_jobs) if (g.map_off < 0) k = std::max(k, g.K); out[19] = k; }
    if (level_of_supernode) for (long long t = 0; t < nsuper; t++) level_of_supernode[t] = hp.level[t];
    return 0;
}

// ---- host-only export of one rank's schedule (no GPU): tests/emulate_plan.py replays it in numpy, with gloo broadcasts
// between ranks, to check the shard + schedule logic on a CPU-only box -------------------------------------------------
static ssb::HostPlan *g_export = nullptr;

extern "C" int ssb200_export_begin(long long n, long long nsuper, const long long *super, const long long *pi, const long long *px,
                                   const long long *s, int nranks, int rank, long long *sizes /*[8]*/)
{
    delete g_export; g_export = new ssb::HostPlan();
    if (!ssb::build_host_plan(n, nsuper, super, pi, px, s, nranks, rank, *g_export)) { delete g_export; g_export = nullptr; return -4; }
    sizes[0] = (long long) g_export->launches.size(); sizes[1] = (long long) g_export->gemm_jobs.size();
    sizes[2] = (long long) g_export->potrf_jobs.size(); sizes[3] = (long long) g_export->trsm_jobs.size();
    sizes[4] = (long long) g_export->steps.size(); sizes[5] = (long long) g_export->updates.size();
    sizes[6] = g_export->relmap_size; sizes[7] = g_export->nlevels;
    return 0;
}

// distributed-storage variant of ssb200_export_begin: same, with the compact per-rank layout (local offsets in every job)
extern "C" int ssb200_export_begin_compact(long long n, long long nsuper, const long long *super, const long long *pi, const long long *px,
                                           const long long *s, int nranks, int rank, long long *sizes /*[8]*/, long long *csizes /*[2]: lxsize, pieces*/)
{
    delete g_export; g_export = new ssb::HostPlan();
    if (!ssb::build_host_plan(n, nsuper, super, pi, px, s, nranks, rank, *g_export, true)) { delete g_export; g_export = nullptr; return -4; }
    sizes[0] = (long long) g_export->launches.size(); sizes[1] = (long long) g_export->gemm_jobs.size();
    sizes[2] = (long long) g_export->potrf_jobs.size(); sizes[3] = (long long) g_export->trsm_jobs.size();
    sizes[4] = (long long) g_export->steps.size(); sizes[5] = (long long) g_export->updates.size();
    sizes[6] = g_export->relmap_size; sizes[7] = g_export->nlevels;
    long long np = 0;
    for (const auto &v : g_export->step_recv) np += (long long) v.size();
    csizes[0] = g_export->lxsize; csizes[1] = np;
    return 0;
}

// lpx[nsuper+1]; pieces[np*5] = step, home_off, cnt, ncols, src_ld; step_next[nsteps]; solve[nsolvejobs*4] = x_off (local), w, xcol0, rows_below.
// Call before ssb200_export_fetch (which releases the plan).
extern "C" long long ssb200_export_compact_fetch(long long *lpx, long long *pieces, int *step_next, long long *solve, long long solve_cap)
{
    if (!g_export || !g_export->compact) return -4;
    const ssb::HostPlan &hp = *g_export;
    for (size_t t = 0; t < hp.lpx.size(); t++) lpx[t] = hp.lpx[t];
    long long q = 0;
    for (size_t k = 0; k < hp.step_recv.size(); k++)
        for (const auto &pc : hp.step_recv[k]) { pieces[5 * q] = (long long) k; pieces[5 * q + 1] = pc.home_off; pieces[5 * q + 2] = pc.cnt; pieces[5 * q + 3] = pc.ncols; pieces[5 * q + 4] = pc.src_ld; q++; }
    for (size_t k = 0; k < hp.step_next.size(); k++) step_next[k] = hp.step_next[k];
    const long long nsj = (long long) hp.solve_jobs.size();
    if (!lpx && !pieces) return nsj;
    if (solve && solve_cap >= nsj)
        for (long long t = 0; t < nsj; t++) { solve[4 * t] = hp.solve_jobs[t].x_off; solve[4 * t + 1] = hp.solve_jobs[t].w; solve[4 * t + 2] = hp.solve_jobs[t].xcol0; solve[4 * t + 3] = hp.solve_jobs[t].rows_below; }
    return nsj;
}

// transient (ring-stored) and trailing-rows-only supernodes of the compact plan: tr[4*t] = transient flag, own base, ring
// base, first stored row; returns the ring depth
extern "C" int ssb200_export_compact_tr(long long *tr)
{
    if (!g_export || !g_export->compact) return -4;
    for (long long t = 0; t < g_export->nsuper; t++) { tr[4 * t] = g_export->transient[t]; tr[4 * t + 1] = g_export->tr_own_base[t]; tr[4 * t + 2] = g_export->tr_ring_base[t]; tr[4 * t + 3] = g_export->rmin[t]; }
    return g_export->ring_depth;
}

// (step, dependency) pairs of the compact plan: returns their number; fills out[2*i], out[2*i+1] when cap suffices
extern "C" long long ssb200_export_compact_deps(long long *out, long long cap)
{
    if (!g_export || !g_export->compact) return -4;
    long long q = 0;
    for (size_t k = 0; k < g_export->step_deps.size(); k++)
        for (int d : g_export->step_deps[k]) { if (out && q < cap) { out[2 * q] = (long long) k; out[2 * q + 1] = d; } q++; }
    return q;
}

// launches[nl*7] = kind, job0, njobs, phase, stream, wait_ev, rec_ev; gemm[ng*9] = a_off,c_off,map_off,lda,ldc,K,nd1,nd2,c_col0; panel arrays [..*6] = x_off,lda,w,
// rows_below,col0,snode; steps[ns*7] = launch_begin,launch_mid,launch_end,src,off,cnt,wait_remote; updates[nu*6] = d,s,p0,nd1,nd2,map_off; owner[nsuper]
extern "C" int ssb200_export_fetch(long long *launches, long long *gemm, long long *potrf, long long *trsm, long long *steps,
                                   long long *updates, int *owner)
{
    if (!g_export) return -4;
    const ssb::HostPlan &hp = *g_export;
    for (size_t t = 0; t < hp.launches.size(); t++) {
        const auto &L = hp.launches[t]; long long *o = launches + 7 * t;
        o[0] = L.kind; o[1] = L.job0; o[2] = L.njobs; o[3] = L.phase; o[4] = L.stream; o[5] = L.wait_ev; o[6] = L.rec_ev;
    }
    for (size_t t = 0; t < hp.gemm_jobs.size(); t++) { const auto &g = hp.gemm_jobs[t]; long long *o = gemm + 9 * t; o[0] = g.a_off; o[1] = g.c_off; o[2] = g.map_off; o[3] = g.lda; o[4] = g.ldc; o[5] = g.K; o[6] = g.nd1; o[7] = g.nd2; o[8] = g.c_col0; }
    auto panel = [](const std::vector<ssb::PanelJob> &v, long long *out) { for (size_t t = 0; t < v.size(); t++) { long long *o = out + 6 * t; o[0] = v[t].x_off; o[1] = v[t].lda; o[2] = v[t].w; o[3] = v[t].rows_below; o[4] = v[t].col0; o[5] = v[t].snode; } };
    panel(hp.potrf_jobs, potrf); panel(hp.trsm_jobs, trsm);
    for (size_t t = 0; t < hp.steps.size(); t++) { const auto &st = hp.steps[t]; long long *o = steps + 7 * t; o[0] = st.launch_begin; o[1] = st.launch_mid; o[2] = st.launch_end; o[3] = st.bcast_src; o[4] = st.off; o[5] = st.cnt; o[6] = st.wait_remote; }
    for (size_t t = 0; t < hp.updates.size(); t++) { const auto &u = hp.updates[t]; long long *o = updates + 6 * t; o[0] = u.d; o[1] = u.s; o[2] = u.p0; o[3] = u.nd1; o[4] = u.nd2; o[5] = u.map_off; }
    for (size_t t = 0; t < hp.owner.size(); t++) owner[t] = hp.owner[t];
    delete g_export; g_export = nullptr;
    return 0;
}
