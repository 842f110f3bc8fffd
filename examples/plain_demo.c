/* plain_demo.c — the PLAIN layer of include/suitesparse_b200.h from C: factorize and solve a tiny SPD system whose
 * supernodal structure is written by hand (one supernode = a dense 3x3 block).  Build:
 *     gcc -std=c99 -I include examples/plain_demo.c -L suitesparse_b200/csrc -lsuitesparse_b200 -Wl,-rpath,$PWD/suitesparse_b200/csrc -o plain_demo
 * Needs a B200 to run (there is no CPU fallback); tests/test_abi.py only compiles and links it. */
#include <stdio.h>
#include "suitesparse_b200.h"

int main(void)
{
    /* A = [4 1 0; 1 5 2; 0 2 6], lower triangle in CSC; one supernode holding all three columns */
    ssb_long Ap[] = {0, 2, 4, 5}, Ai[] = {0, 1, 1, 2, 2};
    double Ax[] = {4, 1, 5, 2, 6};
    ssb_long super[] = {0, 3}, pi[] = {0, 3}, px[] = {0, 9}, s[] = {0, 1, 2};
    double Lx[9], x[3] = {1, 2, 3}, beta[2] = {0, 0};
    ssb_long minor = -1;
    ssb200_plan *P = ssb200_plan_create(3, 1, super, pi, px, s, -1);
    if (!P) { fprintf(stderr, "plan: %s\n", ssb200_last_error()); return 2; }
    int st = ssb200_factorize(P, -1, Ap, Ai, NULL, Ax, 3, NULL, NULL, NULL, NULL, beta, 0, Lx, &minor);
    if (st != 0) { fprintf(stderr, "factorize: status %d minor %lld %s\n", st, (long long) minor, ssb200_last_error()); return 3; }
    if (ssb200_solve(P, 2, x, 1, 3) != 0) { fprintf(stderr, "solve: %s\n", ssb200_last_error()); return 4; }
    ssb200_stats stats;
    ssb200_get_stats(P, &stats);
    printf("%s\nL(0,0)=%.6f x = %.6f %.6f %.6f  (kernel launches since plan creation: %lld)\n", ssb200_version(), Lx[0], x[0], x[1], x[2],
           (long long) stats.kernel_launches_total);
    ssb200_plan_destroy(P);
    return 0;
}
