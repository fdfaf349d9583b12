/* tests/tools/sqp_host.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Compiles the algebra of ntg_b200/csrc/ntg_sqp.cuh (Goldfarb-Idnani QP, one SQP iteration) with g++
 * for a "CTA" of one thread, so that tests/test_sqp_host.py can check it against the numpy
 * restatement (tests/tools/sqp_reference.py) without a GPU.  The library never runs this on the host. */
#include <vector>
#include <cstring>
#include "ntg_sqp.cuh"

using namespace ntgb::sqp;

extern "C" {

int sqp_host_qp(int n, int m, const double *G, const double *g0, const double *A, const double *bl, const double *bu,
                double *x, double *lam, int *state)
{
    const int nt = 1;
    std::vector<double> dbl(sqp_smem_doubles(n, m, nt) + 16, 0.0);
    std::vector<int> ints(sqp_smem_ints(n, m, nt) + 16, 0);
    Qp w;
    double *Bm, *Lm, *vec, *gr, *hrow;
    sqp_carve(dbl.data(), ints.data(), n, m, nt, w, Bm, Lm, vec, gr, hrow);
    const Coop cg{0, 1, 0, 1};
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) Lm[i * w.ld + j] = G[i * n + j];
    for (int i = 0; i < m; i++) {
        for (int k = 0; k < n; k++) w.A[i * w.ld + k] = A[i * n + k];
        w.bl[i] = bl[i];
        w.bu[i] = bu[i];
    }
    if (chol_lower(cg, Lm, n, w.ld, w.shi)) return -1;
    const int st = gi_solve(cg, w, Lm, g0);
    memcpy(x, w.x, sizeof(double) * n);
    memcpy(lam, w.lam, sizeof(double) * m);
    memcpy(state, w.state, sizeof(int) * m);
    return st;
}

/* one SQP iteration: state arrays as in StepState (y unused here) */
void sqp_host_step(int n, int m, double f, const double *gr_in, const double *h, const double *hl, const double *hu,
                   const double *A, double *B, double *lam, double *sprev, double *grLold, double *grold, double *d, double *scal,
                   int *flag, int *istate, double gtol, double ctol, double rho_pen)
{
    const int nt = 1;
    std::vector<double> dbl(sqp_smem_doubles(n, m, nt) + 16, 0.0);
    std::vector<int> ints(sqp_smem_ints(n, m, nt) + 16, 0);
    Qp w;
    double *Bm, *Lm, *vec, *gr, *hrow;
    sqp_carve(dbl.data(), ints.data(), n, m, nt, w, Bm, Lm, vec, gr, hrow);
    const Coop cg{0, 1, 0, 1};
    for (int i = 0; i < m; i++)
        for (int k = 0; k < n; k++) w.A[i * w.ld + k] = A[i * n + k];
    memcpy(gr, gr_in, sizeof(double) * n);
    StepState S{nullptr, B, lam, sprev, grLold, grold, d, scal, flag, istate};
    StepOpts o{gtol, ctol, rho_pen};
    sqp_step(cg, w, S, o, f, gr, h, hl, hu, Bm, Lm, vec);
}
}
