/*
 * kincar_batch.c -- plain C host program on the batched C ABI (include/ntg_b200.h).
 *
 * Evaluates cost, gradient, constraints and the band Jacobian of the kinematic-car problem
 * (2 flat outputs, 2 intervals, order 5, multiplicity 3, 64 breakpoints; callbacks from the
 * kincar pack) for a batch of coefficient vectors with ONE call of ntgb_eval_host(), and prints
 * a few numbers a test can compare with the CPU oracle.
 *
 *   gcc -O2 -Iinclude examples/kincar_batch.c -Lntg_b200/lib -lntgpack_kincar -lntg_b200 \
 *       -Wl,-rpath,$PWD/ntg_b200/lib -o kincar_batch
 *   ./kincar_batch 1000
 *
 * The callbacks are referenced by their ordinary C names: the pack shared object exports the
 * host versions, and the evaluator finds the device versions by those host addresses.  (Build
 * the program position-independent -- gcc's default -- so the addresses it takes are the
 * definitions' own, not PLT stubs.)
 */
#include <stdio.h>
#include <stdlib.h>

#include "ntg.h" /* linspace, AV; pulls in ntg_b200.h */

void kc_ucf(int *mode, int *nstate, int *i, double *f, double *df, double **zp);
void kc_nltcf(int *mode, int *nstate, int *i, double *f, double **df, double **zp);

int main(int argc, char **argv)
{
    const int P = argc > 1 ? atoi(argv[1]) : 256;
    enum { NOUT = 2, NBPS = 64, NC = 14, NCNLN = 2 * NBPS, S = 10 };
    int ninterv[NOUT] = {2, 2}, order[NOUT] = {5, 5}, mult[NOUT] = {3, 3}, maxderiv[NOUT] = {3, 3};
    double k0[3], k1[3], bps[NBPS];
    const double *knots[NOUT] = {k0, k1};
    AV costav[2] = {{0, 2}, {1, 2}};
    AV conav[4] = {{0, 1}, {0, 2}, {1, 1}, {1, 2}};
    double lower[2] = {0.0, -50.0}, upper[2] = {400.0, 50.0};
    ntgb_setup s = {0};
    ntgb_problem *pb = NULL;
    ntgb_dims d;
    ntgb_eval_args a = {0};
    double *C, *f, *g, *c, *J, *res;
    unsigned long long lcg = 12345;
    int p, e;

    linspace(k0, 0.0, 5.0, 3);
    linspace(k1, 0.0, 5.0, 3);
    linspace(bps, 0.0, 5.0, NBPS);
    s.nout = NOUT; s.bps = bps; s.nbps = NBPS; s.kninterv = ninterv; s.knots = knots;
    s.order = order; s.mult = mult; s.maxderiv = maxderiv;
    s.nucf = 1; s.ucf = kc_ucf; s.ntrajectorycostav = 2; s.trajectorycostav = costav;
    s.nnltc = 2; s.nltcf = kc_nltcf; s.ntrajectoryconstrav = 4; s.trajectoryconstrav = conav;
    s.lowerb = lower; s.upperb = upper;
    if (ntgb_create(&pb, &s, 0) != 0) {
        fprintf(stderr, "ntgb_create: %s\n", ntgb_last_error());
        return 2;
    }
    ntgb_get_dims(pb, &d);
    if (d.nC != NC || d.ncnln != NCNLN || d.sorder != S) {
        fprintf(stderr, "unexpected sizes %d %d %d\n", d.nC, d.ncnln, d.sorder);
        return 3;
    }
    /* page-locked buffers: ntgb_eval_host then overlaps its copies with the kernels */
    C = ntgb_host_alloc(sizeof(double) * (size_t)P * NC);
    f = ntgb_host_alloc(sizeof(double) * (size_t)P);
    g = ntgb_host_alloc(sizeof(double) * (size_t)P * NC);
    c = ntgb_host_alloc(sizeof(double) * (size_t)P * NCNLN);
    J = ntgb_host_alloc(sizeof(double) * (size_t)P * NCNLN * S);
    res = ntgb_host_alloc(sizeof(double) * (size_t)P * 2);
    if (!C || !f || !g || !c || !J || !res) {
        fprintf(stderr, "host allocation: %s\n", ntgb_last_error());
        return 5;
    }
    for (p = 0; p < P; p++)
        for (e = 0; e < NC; e++) { /* a reproducible batch: x coefficients in [0,40), y in [-2,2) */
            double u;
            lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
            u = (double)(lcg >> 11) / 9007199254740992.0;
            C[(size_t)p * NC + e] = e < 7 ? 40.0 * u : 4.0 * u - 2.0;
        }
    a.P = P; a.C = C; a.mode_obj = 2; a.mode_con = 2; a.nstate = 1;
    a.f = f; a.g = g; a.c = c; a.J = J; a.jac_layout = NTGB_JAC_BAND; a.result = res;
    if (ntgb_eval_host(pb, &a) != 0) {
        fprintf(stderr, "ntgb_eval_host: %s\n", ntgb_last_error());
        return 4;
    }
    for (p = 0; p < P; p += (P > 4 ? P / 4 : 1))
        printf("p %d f %.17g g0 %.17g c0 %.17g J0 %.17g viol %.17g\n", p, f[p], g[(size_t)p * NC],
               c[(size_t)p * NCNLN], J[(size_t)p * NCNLN * S], res[2 * p + 1]);
    ntgb_destroy(pb);
    ntgb_host_free(C); ntgb_host_free(f); ntgb_host_free(g); ntgb_host_free(c); ntgb_host_free(J); ntgb_host_free(res);
    return 0;
}
