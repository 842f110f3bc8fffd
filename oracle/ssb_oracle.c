/* ssb_oracle.c — CPU restatement of CHOLMOD's supernodal numeric factorization and solves.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may load it.  The product
 * (suitesparse_b200/csrc) never calls it and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against the reference's own
 * build (oracle/_ref/libcholmod_ref.so, compiled from /root/reference by oracle/Makefile) on the
 * reference's bundled matrices (CHOLMOD/Demo/Matrix, CHOLMOD/Tcov/Matrix) and on the mesh problems
 * whose nnz(L)/flops the reference pins in MATLAB_Tools/MESHND/meshnd_quality_out.txt, and against
 * the golden fixtures under tests/golden/ generated from that build (tests/golden/make_golden.py).
 * The arithmetic itself lives in BLAS/LAPACK (un-vendored, unpinned: CMakeLists.txt:122,124), so
 * L's values agree with the reference build to rounding (tests state the tolerance), the integer
 * side (update enumeration, minor, status) exactly.
 *
 * Plain C, naive loops, no BLAS.  Real double precision, 64-bit indices.  Follows
 *   CHOLMOD/Supernodal/t_cholmod_super_numeric.c:279-1068   (factorize)
 *   CHOLMOD/Supernodal/t_cholmod_super_solve.c:60-130,268-332 (+ nrhs>1 branches)  (solves)
 *   CHOLMOD/Supernodal/cholmod_super_symbolic.c:907-948     (update enumeration / maxcsize)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef int64_t Long;

/* Column k -> supernode (cholmod_super_numeric.c:266-275). */
static Long *build_supermap(Long n, Long nsuper, const Long *Super)
{
    Long *SuperMap = (Long *) malloc((size_t)(n > 0 ? n : 1) * sizeof(Long));
    if (!SuperMap) return NULL;
    for (Long s = 0; s < nsuper; s++)
        for (Long k = Super[s]; k < Super[s + 1]; k++) SuperMap[k] = s;
    return SuperMap;
}

/* Enumerate every (descendant d, ancestor s) update.  The reference finds them at run time with the
 * Head/Next/Lpos lists (t_cholmod_super_numeric.c:442-460,583-611,787-808); they are the maximal runs
 * of d's sub-diagonal rows falling into one ancestor, the same walk cholmod_super_symbolic.c:916-941
 * uses for maxcsize.  Output arrays (caller-sized to the count returned by a first call with NULLs):
 * ud, us, up0 (offset of the run in d's row list), und1 (rows inside s), und2 (rows from p0 to the end).
 * Returns the number of updates; *maxcsize_out = max(ndrow1*ndrow2), at least 1. */
Long ssbo_enumerate_updates(Long n, Long nsuper, const Long *Super, const Long *Lpi, const Long *Ls,
                            Long *ud, Long *us, Long *up0, Long *und1, Long *und2, Long *maxcsize_out)
{
    Long *SuperMap = build_supermap(n, nsuper, Super);
    Long cnt = 0, maxc = 1;
    for (Long d = 0; d < nsuper; d++) {
        Long ndcol = Super[d + 1] - Super[d];
        Long pdi = Lpi[d], pdend = Lpi[d + 1];
        Long p = pdi + ndcol;
        while (p < pdend) {
            Long s = SuperMap[Ls[p]];
            Long q = p;
            while (q < pdend && SuperMap[Ls[q]] == s) q++;
            Long nd1 = q - p, nd2 = pdend - p;
            if (ud) { ud[cnt] = d; us[cnt] = s; up0[cnt] = p - pdi; und1[cnt] = nd1; und2[cnt] = nd2; }
            if (nd1 * nd2 > maxc) maxc = nd1 * nd2;
            cnt++;
            p = q;
        }
    }
    if (maxcsize_out) *maxcsize_out = maxc;
    free(SuperMap);
    return cnt;
}

/* Unblocked lower Cholesky of the leading ncol x ncol block (LAPACK dpotrf 'L' contract:
 * returns 0, or j (1-based) if the leading minor of order j is not positive definite). */
static Long potrf_lower(double *A, Long ncol, Long lda)
{
    for (Long j = 0; j < ncol; j++) {
        double ajj = A[j + j * lda];
        for (Long k = 0; k < j; k++) ajj -= A[j + k * lda] * A[j + k * lda];
        if (!(ajj > 0.0)) { A[j + j * lda] = ajj; return j + 1; }
        ajj = sqrt(ajj);
        A[j + j * lda] = ajj;
        for (Long i = j + 1; i < ncol; i++) {
            double v = A[i + j * lda];
            for (Long k = 0; k < j; k++) v -= A[i + k * lda] * A[j + k * lda];
            A[i + j * lda] = v / ajj;
        }
    }
    return 0;
}

/* B (m x ncol, ldb) <- B * L^{-T}, L ncol x ncol lower (dtrsm R,L,C,N; t_cholmod_super_numeric.c:997). */
static void trsm_rltn(const double *L, Long ncol, Long lda, double *B, Long m, Long ldb)
{
    for (Long j = 0; j < ncol; j++) {
        double ljj = L[j + j * lda];
        for (Long i = 0; i < m; i++) {
            double v = B[i + j * ldb];
            for (Long k = 0; k < j; k++) v -= B[i + k * ldb] * L[j + k * lda];
            B[i + j * ldb] = v / ljj;
        }
    }
}

/* Supernodal LL' factorization.  A: CSC (Ap/Ai/Ax, Anz!=NULL if unpacked), stype<0 symmetric lower
 * (only i>=k used) or stype==0 with F (computes A*F).  Lx has px[nsuper] entries.
 * Returns 0 ok, 1 not positive definite (then *minor = failing column), -4 invalid. */
int ssbo_factorize(Long n, Long nsuper, const Long *Super, const Long *Lpi, const Long *Lpx, const Long *Ls,
                   int stype, const Long *Ap, const Long *Ai, const Long *Anz, const double *Ax,
                   const Long *Fp, const Long *Fi, const Long *Fnz, const double *Fx,
                   const double *beta, int quick_return_if_not_posdef, double *Lx, Long *minor)
{
    if (stype > 0 || (stype == 0 && !Fp)) return -4;
    Long xsize = Lpx[nsuper];
    Long nup = ssbo_enumerate_updates(n, nsuper, Super, Lpi, Ls, NULL, NULL, NULL, NULL, NULL, NULL);
    Long *ud = malloc((nup + 1) * sizeof(Long)), *us = malloc((nup + 1) * sizeof(Long));
    Long *up0 = malloc((nup + 1) * sizeof(Long)), *und1 = malloc((nup + 1) * sizeof(Long));
    Long *und2 = malloc((nup + 1) * sizeof(Long)), maxc;
    ssbo_enumerate_updates(n, nsuper, Super, Lpi, Ls, ud, us, up0, und1, und2, &maxc);
    /* bucket the updates by target supernode */
    Long *tp = calloc(nsuper + 2, sizeof(Long)), *tl = malloc((nup + 1) * sizeof(Long));
    for (Long u = 0; u < nup; u++) tp[us[u] + 1]++;
    for (Long s = 0; s < nsuper; s++) tp[s + 1] += tp[s];
    Long *fill = malloc((nsuper + 1) * sizeof(Long));
    memcpy(fill, tp, (nsuper + 1) * sizeof(Long));
    for (Long u = 0; u < nup; u++) tl[fill[us[u]]++] = u;
    Long *Map = malloc((n > 0 ? n : 1) * sizeof(Long));
    for (Long i = 0; i < n; i++) Map[i] = -1;
    double *Cbuf = malloc((size_t) maxc * sizeof(double));
    int status = 0;
    *minor = n;
    memset(Lx, 0, (size_t) xsize * sizeof(double));                                   /* :305-317 */

    for (Long s = 0; s < nsuper; s++) {
        Long k1 = Super[s], k2 = Super[s + 1], nscol = k2 - k1;
        Long psi = Lpi[s], nsrow = Lpi[s + 1] - psi, psx = Lpx[s];
        Long ncol_fact = nscol;
        int repeat = 0;
    again:
        for (Long p = psx; p < psx + nsrow * nscol; p++) Lx[p] = 0;
        for (Long r = 0; r < nsrow; r++) Map[Ls[psi + r]] = r;                        /* :329-333 */
        for (Long k = k1; k < k2; k++) {                                             /* :358-417 */
            if (stype != 0) {
                Long p = Ap[k], pend = Anz ? p + Anz[k] : Ap[k + 1];
                for (; p < pend; p++) {
                    Long i = Ai[p];
                    if (i >= k) { Long im = Map[i]; if (im >= 0 && im < nsrow && Ls[psi + im] == i) Lx[psx + im + (k - k1) * nsrow] = Ax[p]; }
                }
            } else {
                Long pf = Fp[k], pfend = Fnz ? pf + Fnz[k] : Fp[k + 1];
                for (; pf < pfend; pf++) {
                    Long j = Fi[pf]; double fjk = Fx[pf];
                    Long p = Ap[j], pend = Anz ? p + Anz[j] : Ap[j + 1];
                    for (; p < pend; p++) {
                        Long i = Ai[p];
                        if (i >= k) { Long im = Map[i]; if (im >= 0 && im < nsrow && Ls[psi + im] == i) Lx[psx + im + (k - k1) * nsrow] += Ax[p] * fjk; }
                    }
                }
            }
        }
        if (beta && beta[0] != 0.0)                                                   /* :421-431 */
            for (Long c = 0; c < nscol; c++) Lx[psx + c * (nsrow + 1)] += beta[0];
        /* descendant updates: C = Lall * L1' then extend-add (:676-772) */
        for (Long t = tp[s]; t < tp[s + 1]; t++) {
            Long u = tl[t], d = ud[u];
            Long ndcol = Super[d + 1] - Super[d], ndrow = Lpi[d + 1] - Lpi[d];
            Long p0 = up0[u], nd1 = und1[u], nd2 = und2[u];
            const double *Ld = Lx + Lpx[d] + p0;
            const Long *rows = Ls + Lpi[d] + p0;
            for (Long j = 0; j < nd1; j++)
                for (Long i = j; i < nd2; i++) {
                    double c = 0;
                    for (Long k = 0; k < ndcol; k++) c += Ld[i + k * ndrow] * Ld[j + k * ndrow];
                    Cbuf[i + j * nd2] = c;
                }
            for (Long j = 0; j < nd1; j++) {
                Long pj = psx + Map[rows[j]] * nsrow;
                for (Long i = j; i < nd2; i++) Lx[pj + Map[rows[i]]] -= Cbuf[i + j * nd2];
            }
        }
        Long info = potrf_lower(Lx + psx, ncol_fact, nsrow);                          /* :864 */
        if (repeat) {
            info = 0;
            for (Long p = psx + nsrow * ncol_fact; p < psx + nsrow * nscol; p++) Lx[p] = 0;
        }
        if (info != 0) {                                                              /* :905-968 */
            status = 1;
            *minor = k1 + info - 1;
            for (Long p = psx; p < xsize; p++) Lx[p] = 0;
            if (info == 1 || quick_return_if_not_posdef) break;
            repeat = 1; ncol_fact = info - 1;
            goto again;
        }
        if (nsrow - ncol_fact > 0)                                                    /* :974-1002 */
            trsm_rltn(Lx + psx, ncol_fact, nsrow, Lx + psx + ncol_fact, nsrow - ncol_fact, nsrow);
        if (repeat) break;
    }
    free(ud); free(us); free(up0); free(und1); free(und2); free(tp); free(tl); free(fill); free(Map); free(Cbuf);
    return status;
}

/* X <- L \ X, X n-by-nrhs, leading dimension ldx (t_cholmod_super_solve.c:60-219). */
void ssbo_lsolve(Long nsuper, const Long *Super, const Long *Lpi, const Long *Lpx, const Long *Ls,
                 const double *Lx, double *X, Long nrhs, Long ldx)
{
    for (Long s = 0; s < nsuper; s++) {
        Long k1 = Super[s], nscol = Super[s + 1] - k1, psi = Lpi[s], nsrow = Lpi[s + 1] - psi;
        const double *L1 = Lx + Lpx[s];
        for (Long r = 0; r < nrhs; r++) {
            double *x = X + r * ldx;
            for (Long j = 0; j < nscol; j++) {                       /* dtrsv L,N,N */
                double v = x[k1 + j] / L1[j + j * nsrow];
                x[k1 + j] = v;
                for (Long i = j + 1; i < nscol; i++) x[k1 + i] -= L1[i + j * nsrow] * v;
            }
            for (Long i = nscol; i < nsrow; i++) {                   /* gather, dgemv, scatter */
                double acc = 0;
                for (Long j = 0; j < nscol; j++) acc += L1[i + j * nsrow] * x[k1 + j];
                x[Ls[psi + i]] -= acc;
            }
        }
    }
}

/* X <- L' \ X (t_cholmod_super_solve.c:268-411). */
void ssbo_ltsolve(Long nsuper, const Long *Super, const Long *Lpi, const Long *Lpx, const Long *Ls,
                  const double *Lx, double *X, Long nrhs, Long ldx)
{
    for (Long s = nsuper - 1; s >= 0; s--) {
        Long k1 = Super[s], nscol = Super[s + 1] - k1, psi = Lpi[s], nsrow = Lpi[s + 1] - psi;
        const double *L1 = Lx + Lpx[s];
        for (Long r = 0; r < nrhs; r++) {
            double *x = X + r * ldx;
            for (Long j = nscol - 1; j >= 0; j--) {
                double v = x[k1 + j];
                for (Long i = nscol; i < nsrow; i++) v -= L1[i + j * nsrow] * x[Ls[psi + i]];   /* dgemv 'C' */
                for (Long i = j + 1; i < nscol; i++) v -= L1[i + j * nsrow] * x[k1 + i];          /* dtrsv L,C,N */
                x[k1 + j] = v / L1[j + j * nsrow];
            }
        }
    }
}
