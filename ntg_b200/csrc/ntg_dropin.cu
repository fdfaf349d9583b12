/*
 * ntg_dropin.cu -- the reference's own entry points on top of the batched
 * evaluator, so a program written against NTG links unchanged:
 *
 *   ntg()            reference src/ntg.c:54-267
 *   npsoloption()    src/ntg.c:269-272
 *   linspace()       src/ntg.c:374-389
 *   printNTGBanner() src/ntg.c:391-405
 *   SplineInterp()   src/colloc.c:449-484
 *   Matrix helpers   src/matrix.c (the subset user programs call)
 *
 * ntg() builds the NPSOL problem exactly as the reference does (n, nclin,
 * ncnln, A, bl, bu, workspaces, "nolist", "derivative level = 3") and hands
 * NPSOL two trampolines with NPSOL's Fortran callback ABI; each trampoline is
 * one P=1 evaluation on the GPU through ntgb_eval_host().  NPSOL itself is
 * separately licensed: npsol_ / npoptn_ are resolved with dlsym() at run time
 * and ntg() reports NTG_INFORM_NO_NPSOL when they are absent.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ntg.h"
#include "pgs_device.cuh"

namespace {

typedef void (*funcon_t)(int *, int *, int *, int *, int *, double *, double *, double *, int *);
typedef void (*funobj_t)(int *, int *, double *, double *, double *, int *);
typedef void (*npsol_t)(int *, int *, int *, int *, int *, int *, double *, double *, double *, funcon_t,
                        funobj_t, int *, int *, int *, double *, double *, double *, double *, double *,
                        double *, double *, int *, int *, double *, int *);
typedef void (*npoptn_t)(const char *, long);

/* NPSOL's callbacks carry no user pointer, so -- like the reference's
 * file-static globals (src/ntg.c:17-41) -- one ntg() runs at a time */
ntgb_problem *g_pb = nullptr;
ntgb_dims g_dims;
int g_eval_error = 0;

void np_funobj(int *mode, int *n, double *x, double *y, double *yprime, int *nstate)
{
    (void)n;
    if (*mode < 0 || *mode > 2) { *nstate = -1; return; } /* reference src/ntg.c:332-333 */
    ntgb_eval_args a;
    memset(&a, 0, sizeof a);
    a.P = 1; a.C = x; a.mode_obj = *mode; a.mode_con = -1; a.nstate = *nstate;
    a.f = y; a.g = yprime; a.jac_layout = NTGB_JAC_NONE;
    const int rc = ntgb_eval_host(g_pb, &a);
    if (rc == NTGB_EABORT) {
        *mode = -1; /* the user's callback asked NPSOL to stop, as in the reference */
    } else if (rc != 0) {
        fprintf(stderr, "ntg: objective evaluation failed: %s\n", ntgb_last_error());
        g_eval_error = 1;
        *mode = -1;
    }
}

void np_funcon(int *mode, int *ncnln, int *n, int *nrowj, int *needc, double *x, double *c, double *cjac,
               int *nstate)
{
    (void)ncnln; (void)n; (void)nrowj; (void)needc;
    if (*mode < 0 || *mode > 2) { *mode = -1; return; } /* reference src/ntg.c:368-369 */
    ntgb_eval_args a;
    memset(&a, 0, sizeof a);
    a.P = 1; a.C = x; a.mode_obj = -1; a.mode_con = *mode; a.nstate = *nstate;
    a.c = c; a.J = cjac; a.jac_layout = NTGB_JAC_DENSE;
    const int rc = ntgb_eval_host(g_pb, &a);
    if (rc == NTGB_EABORT) {
        *mode = -1;
    } else if (rc != 0) {
        fprintf(stderr, "ntg: constraint evaluation failed: %s\n", ntgb_last_error());
        g_eval_error = 1;
        *mode = -1;
    }
}

__global__ void k_spline_single(const double *knots, int ninterv, const double *coefs, int order, int mult,
                                int md, double x, double *aug, double *f)
{
    const int n = ninterv * (order - mult) + mult, naug = n + order;
    for (int i = 0; i < naug; i++) aug[i] = ntgb::pgs_augknot(knots, ninterv, order, mult, i);
    double a[PGS_MAXK * PGS_MAXK], db[PGS_MAXK * PGS_MAXK];
    for (int e = 0; e < order * md; e++) db[e] = 0.0;
    const int left1 = ntgb::pgs_interv(aug, naug, x);
    ntgb::pgs_bsplvd(aug, order, x, left1, a, db, md);
    const int left2 = ntgb::pgs_interv(knots, ninterv + 1, x);
    const int offset = (left2 - 1) * (order - mult);
    for (int d = 0; d < md; d++) {
        double s = 0.0;
        for (int k = 0; k < order; k++) s = s + db[d * order + k] * coefs[offset + k];
        f[d] = s;
    }
}

FILE *open_out(const char *filename, bool *close_it)
{
    *close_it = false;
    if (!strcmp(filename, "stdout")) return stdout;
    if (!strcmp(filename, "stderr")) return stderr;
    FILE *fp = fopen(filename, "w");
    *close_it = fp != nullptr;
    return fp;
}

} /* namespace */

extern "C" {

void npsoloption(const char *option)
{
    static npoptn_t fn = (npoptn_t)dlsym(RTLD_DEFAULT, "npoptn_");
    if (fn) fn(option, (long)strlen(option)); /* hidden Fortran length argument, src/ntg.c:271 */
}

/* accumulating recurrence, NOT numpy's linspace: the last point may land past
 * d1, and interval indices depend on it */
void linspace(double *v, double d0, double d1, int n)
{
    if (d0 == d1) {
        for (int i = 0; i < n; i++) v[i] = d0;
        return;
    }
    const double step = (d1 - d0) / (n - 1);
    v[0] = d0;
    for (int i = 1; i < n; i++) v[i] = v[i - 1] + step;
}

void printNTGBanner(void)
{
    printf("\n  ntg_b200 -- NTG-compatible trajectory generation, collocation evaluated on NVIDIA B200\n\n");
}

Matrix *MakeMatrix(int rows, int cols)
{
    Matrix *m = (Matrix *)malloc(sizeof(Matrix));
    m->elements = DoubleMatrix(rows, cols);
    m->rows = rows;
    m->cols = cols;
    return m;
}

void FreeMatrix(Matrix *m)
{
    FreeDoubleMatrix(m->elements);
    free(m);
}

/* row pointers over one zeroed block, so FreeDoubleMatrix(m->elements) works
 * the way user code expects (examples/kincar.c:300-303,410-412) */
double **DoubleMatrix(int rows, int cols)
{
    double **r = (double **)malloc(sizeof(double *) * (size_t)rows);
    r[0] = (double *)calloc((size_t)rows * cols, sizeof(double));
    for (int i = 1; i < rows; i++) r[i] = r[0] + (size_t)i * cols;
    return r;
}

void FreeDoubleMatrix(double **d)
{
    free(d[0]);
    free(d);
}

void PrintMatrix(const char *filename, Matrix *m)
{
    bool cl;
    FILE *fp = open_out(filename, &cl);
    if (!fp) return;
    for (int i = 0; i < m->rows; i++) {
        for (int j = 0; j < m->cols; j++) fprintf(fp, "%f ", m->elements[i][j]);
        fprintf(fp, "\n");
    }
    fprintf(fp, "\n\n\n");
    if (cl) fclose(fp);
}

void PrintVector(const char *filename, double *f, int nf)
{
    bool cl;
    FILE *fp = open_out(filename, &cl);
    if (!fp) return;
    for (int i = 0; i < nf; i++) fprintf(fp, "%g ", f[i]);
    fprintf(fp, "\n");
    if (cl) fclose(fp);
}

void PrintiVector(const char *filename, int *f, int nf)
{
    bool cl;
    FILE *fp = open_out(filename, &cl);
    if (!fp) return;
    for (int i = 0; i < nf; i++) fprintf(fp, "%d\n", f[i]);
    if (cl) fclose(fp);
}

/* One page-locked, device-mapped staging block kept for the life of the process: a call copies the
 * knots and coefficients into it, launches one thread and reads the derivatives back out of it --
 * no allocation, no cudaMemcpy (examples/kincar.c:396-406 calls this in a loop over its output
 * times).  There is no CPU path: every CUDA failure is fatal and says so. */
static void spline_fail(const char *what, cudaError_t e)
{
    fprintf(stderr, "SplineInterp: %s: %s (there is no CPU path)\n", what, cudaGetErrorString(e));
    abort();
}

void SplineInterp(double *f, double x, double *knots, int ninterv, double *coefs, int ncoefs, int order,
                  int mult, int maxderiv)
{
    static double *h_blk = nullptr, *d_blk = nullptr;
    static size_t cap = 0;
    const int n = ninterv * (order - mult) + mult;
    if (n != ncoefs || ninterv < 1 || order < 1 || order > PGS_MAXK || maxderiv < 1 || maxderiv > order ||
        mult < 0 || mult >= order) {
        fprintf(stderr, "SplineInterp: bad spline description (ninterv %d order %d mult %d maxderiv %d ncoefs %d; "
                        "need ncoefs = ninterv*(order-mult)+mult, maxderiv <= order <= %d)\n",
                ninterv, order, mult, maxderiv, ncoefs, PGS_MAXK);
        abort();
    }
    const size_t nk = (size_t)ninterv + 1, total = nk + n + (n + order) + maxderiv;
    if (total > cap) {
        if (h_blk) cudaFreeHost(h_blk);
        h_blk = nullptr;
        cap = 0;
        const size_t want = total < 4096 ? 4096 : 2 * total;
        cudaError_t e = cudaHostAlloc((void **)&h_blk, want * sizeof(double), cudaHostAllocMapped);
        if (e != cudaSuccess) spline_fail("no CUDA device / cannot allocate the staging block", e);
        e = cudaHostGetDevicePointer((void **)&d_blk, h_blk, 0);
        if (e != cudaSuccess) spline_fail("cudaHostGetDevicePointer", e);
        cap = want;
    }
    double *hk = h_blk, *hc = hk + nk, *hf = hc + n + (n + order);
    memcpy(hk, knots, nk * sizeof(double));
    memcpy(hc, coefs, (size_t)n * sizeof(double));
    double *dk = d_blk, *dc = dk + nk, *daug = dc + n, *df = daug + (n + order);
    k_spline_single<<<1, 1>>>(dk, ninterv, dc, order, mult, maxderiv, x, daug, df);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) spline_fail("kernel launch", e);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) spline_fail("kernel execution", e);
    memcpy(f, hf, (size_t)maxderiv * sizeof(double));
}

void ntg(int nout, double *bps, int nbps, int *kninterv, double **knots, int *order, int *mult, int *maxderiv,
         double *initialguess, int nlic, double **lic, int nltc, double **ltc, int nlfc, double **lfc,
         int nnlic, void (*nlicf)(int *, int *, double *, double **, double **),
         int nnltc, void (*nltcf)(int *, int *, int *, double *, double **, double **),
         int nnlfc, void (*nlfcf)(int *, int *, double *, double **, double **),
         int ninitialconstrav, AV *initialconstrav, int ntrajectoryconstrav, AV *trajectoryconstrav,
         int nfinalconstrav, AV *finalconstrav, double *lowerb, double *upperb,
         int nicf, void (*icf)(int *, int *, double *, double *, double **),
         int nucf, void (*ucf)(int *, int *, int *, double *, double *, double **),
         int nfcf, void (*fcf)(int *, int *, double *, double *, double **),
         int ninitialcostav, AV *initialcostav, int ntrajectorycostav, AV *trajectorycostav,
         int nfinalcostav, AV *finalcostav, int *istate, double *clambda, double *R, int *inform,
         double *objective)
{
    ntgb_setup s;
    memset(&s, 0, sizeof s);
    s.nout = nout; s.bps = bps; s.nbps = nbps; s.kninterv = kninterv; s.knots = knots;
    s.order = order; s.mult = mult; s.maxderiv = maxderiv;
    s.nlic = nlic; s.lic = lic; s.nltc = nltc; s.ltc = ltc; s.nlfc = nlfc; s.lfc = lfc;
    s.nnlic = nnlic; s.nlicf = nlicf; s.nnltc = nnltc; s.nltcf = nltcf; s.nnlfc = nnlfc; s.nlfcf = nlfcf;
    s.ninitialconstrav = ninitialconstrav; s.initialconstrav = initialconstrav;
    s.ntrajectoryconstrav = ntrajectoryconstrav; s.trajectoryconstrav = trajectoryconstrav;
    s.nfinalconstrav = nfinalconstrav; s.finalconstrav = finalconstrav;
    s.lowerb = lowerb; s.upperb = upperb;
    s.nicf = nicf; s.icf = icf; s.nucf = nucf; s.ucf = ucf; s.nfcf = nfcf; s.fcf = fcf;
    s.ninitialcostav = ninitialcostav; s.initialcostav = initialcostav;
    s.ntrajectorycostav = ntrajectorycostav; s.trajectorycostav = trajectorycostav;
    s.nfinalcostav = nfinalcostav; s.finalcostav = finalcostav;

    printNTGBanner(); /* the reference prints its banner on every call, src/ntg.c:161 */

    int device = 0;
    if (const char *e = getenv("NTG_B200_DEVICE")) device = atoi(e);
    ntgb_problem *pb = nullptr;
    if (ntgb_create(&pb, &s, device) != 0) {
        fprintf(stderr, "ntg: setup failed: %s\n", ntgb_last_error());
        if (inform) *inform = NTG_INFORM_SETUP_FAILED;
        return;
    }
    ntgb_get_dims(pb, &g_dims);
    g_pb = pb;
    g_eval_error = 0;

    int NPn = g_dims.nC, NPnclin = g_dims.nclin, NPncnln = g_dims.ncnln;
    int NPldA = NPnclin == 0 ? 1 : NPnclin; /* src/ntg.c:162-170 */
    int NPldJ = NPncnln == 0 ? 1 : NPncnln; /* src/ntg.c:210-219 */
    int NPldR = NPn, NPiter = 0;
    std::vector<double> A((size_t)NPldA * (NPnclin == 0 ? 1 : NPn), 0.0);
    std::vector<double> cJac((size_t)NPldJ * (NPncnln == 0 ? 1 : NPn), 0.0);
    std::vector<double> bl((size_t)NPn + NPnclin + NPncnln), bu(bl.size());
    std::vector<double> c((size_t)(NPncnln > 0 ? NPncnln : 1), 0.0), g((size_t)NPn, 0.0);
    if (NPnclin > 0) ntgb_get_linear(pb, A.data());
    ntgb_get_bounds(pb, bl.data(), bu.data());
    int NPleniw = 3 * NPn + NPnclin + 2 * NPncnln; /* src/ntg.c:237-246 */
    int NPlenw;
    if (NPnclin == 0 && NPncnln == 0) NPlenw = 20 * NPn;
    else if (NPncnln == 0) NPlenw = 2 * NPn * NPn + 20 * NPn + 11 * NPnclin;
    else NPlenw = 2 * NPn * NPn + NPn * NPnclin + 2 * NPn * NPncnln + 20 * NPn + 11 * NPnclin + 21 * NPncnln;
    std::vector<int> iw((size_t)NPleniw, 0);
    std::vector<double> w((size_t)NPlenw, 0.0);

    npsol_t npsol = (npsol_t)dlsym(RTLD_DEFAULT, "npsol_");
    if (!npsol) {
        /* No NPSOL in the process: the library's own batched solvers stand in (P = 1).  Problems
         * of the shipped examples' class (no nonlinear constraints, linear equalities only) go to the
         * reduced-space BFGS (ntgb_solve_eq), everything else to the SQP solver (ntgb_solve_sqp) with
         * the augmented-Lagrangian driver (ntgb_solve_nlp) behind it.  NTG_B200_NO_BUILTIN_SOLVER=1
         * switches this off. */
        const bool builtin = getenv("NTG_B200_NO_BUILTIN_SOLVER") == nullptr;
        bool eq_only = NPncnln == 0;
        for (int i = 0; eq_only && i < NPnclin; i++) eq_only = bl[NPn + i] == bu[NPn + i];
        int rc = NTGB_EINVAL;
        if (builtin) {
            double *dC = nullptr;
            int *dst = nullptr;
            int st = 0;
            double fv = 0.0;
            cudaSetDevice(device);
            const int ngen = NPnclin + NPncnln;
            double *dlam = nullptr;
            int *dist = nullptr;
            const char *used = "reduced-space BFGS";
            bool have_mult = false;
            if (cudaMalloc((void **)&dC, sizeof(double) * ((size_t)NPn + 2)) == cudaSuccess &&
                cudaMalloc((void **)&dst, sizeof(int) * 2) == cudaSuccess &&
                cudaMalloc((void **)&dlam, sizeof(double) * (size_t)(ngen > 0 ? ngen : 1)) == cudaSuccess &&
                cudaMalloc((void **)&dist, sizeof(int) * (size_t)(ngen > 0 ? ngen : 1)) == cudaSuccess &&
                cudaMemcpy(dC, initialguess, sizeof(double) * NPn, cudaMemcpyHostToDevice) == cudaSuccess) {
                /* equalities only: reduced-space BFGS.  Anything else: the SQP solver (NPSOL's method:
                 * active-set QP subproblems, quasi-Newton Hessian; returns multipliers and the active set);
                 * if the dense reduced QP does not fit in shared memory or it does not converge, the
                 * augmented-Lagrangian driver from the same guess. */
                if (eq_only) {
                    rc = ntgb_solve_eq(pb, 1, dC, dC + NPn, dst, dst + 1, nullptr, nullptr);
                } else {
                    used = "SQP solver (dual active-set QP, BFGS)";
                    rc = ntgb_solve_sqp(pb, 1, dC, dC + NPn, dC + NPn + 1, dst, dst + 1, dlam, dist, nullptr, nullptr);
                    int s1 = 0;
                    if (rc == 0 && cudaMemcpy(&s1, dst + 1, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) rc = NTGB_ECUDA;
                    have_mult = rc == 0 && s1 == 1;
                    if (rc == NTGB_ELIMIT || (rc == 0 && s1 != 1)) {
                        used = "augmented-Lagrangian / reduced-space BFGS";
                        rc = cudaMemcpy(dC, initialguess, sizeof(double) * NPn, cudaMemcpyHostToDevice) == cudaSuccess
                                 ? ntgb_solve_nlp(pb, 1, dC, dC + NPn, dC + NPn + 1, dst, dst + 1, nullptr, nullptr)
                                 : NTGB_ECUDA;
                    }
                }
                if (rc == 0 && cudaMemcpy(&st, dst + 1, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess &&
                    cudaMemcpy(&fv, dC + NPn, sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess &&
                    cudaMemcpy(initialguess, dC, sizeof(double) * NPn, cudaMemcpyDeviceToHost) == cudaSuccess) {
                    if (objective) *objective = fv;
                    /* NPSOL's codes: 0 optimal, 1 no further improvement possible, 4 iteration limit
                     * (6 = not converged, the nonlinear constraints may be violated) */
                    if (inform) *inform = st == 1 ? 0 : (st == 2 ? 1 : (eq_only ? 4 : 6));
                    if (have_mult) {
                        /* NPSOL's layout (src/ntg.h:64-68): n variables (no bounds in NTG: free, multiplier 0),
                         * then the nclin linear and the ncnln nonlinear rows */
                        if (istate) {
                            for (int i = 0; i < NPn; i++) istate[i] = 0;
                            if (ngen > 0 && cudaMemcpy(istate + NPn, dist, sizeof(int) * ngen, cudaMemcpyDeviceToHost) != cudaSuccess)
                                have_mult = false;
                        }
                        if (clambda) {
                            for (int i = 0; i < NPn; i++) clambda[i] = 0.0;
                            if (ngen > 0 && cudaMemcpy(clambda + NPn, dlam, sizeof(double) * ngen, cudaMemcpyDeviceToHost) != cudaSuccess)
                                have_mult = false;
                        }
                    }
                    fprintf(stderr,
                            "ntg: NPSOL (npsol_) is not linked into this process; solved with ntg_b200's built-in\n"
                            "     %s instead (n=%d, nclin=%d, ncnln=%d).\n"
                            "     %s\n",
                            used, NPn, NPnclin, NPncnln,
                            have_mult ? "istate and clambda are set (NPSOL's layout); R is not."
                                      : "istate, clambda and R are not set on this path.");
                } else if (rc == 0) {
                    rc = NTGB_ECUDA;
                }
            }
            if (dlam) cudaFree(dlam);
            if (dist) cudaFree(dist);
            if (dC) cudaFree(dC);
            if (dst) cudaFree(dst);
            if (rc != 0) fprintf(stderr, "ntg: built-in solver failed: %s\n", ntgb_last_error());
        }
        if (rc != 0) {
            fprintf(stderr,
                    "ntg: NPSOL (npsol_) is not linked into this process -- it is separately licensed and not part of\n"
                    "     ntg_b200.  The problem was set up on the GPU (n=%d, nclin=%d, ncnln=%d) but cannot be solved;\n"
                    "     use the batched evaluation API (ntg_b200.h) or link NPSOL.\n",
                    NPn, NPnclin, NPncnln);
            if (inform) *inform = NTG_INFORM_NO_NPSOL;
        }
    } else {
        npsoloption("nolist");                 /* src/ntg.c:248 */
        npsoloption("derivative level = 3");   /* src/ntg.c:249 */
        npsol(&NPn, &NPnclin, &NPncnln, &NPldA, &NPldJ, &NPldR, A.data(), bl.data(), bu.data(), np_funcon,
              np_funobj, inform, &NPiter, istate, c.data(), cJac.data(), clambda, objective, g.data(), R,
              initialguess, iw.data(), &NPleniw, w.data(), &NPlenw);
        if (g_eval_error && inform) *inform = NTG_INFORM_SETUP_FAILED;
    }
    g_pb = nullptr;
    ntgb_destroy(pb);
}

} /* extern "C" */
