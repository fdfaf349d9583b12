/*
 * oracle/npsol_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A stand-in for the separately licensed NPSOL: npsol_() with NPSOL's
 * 25-argument Fortran interface (reference src/ntg.c:250-253) and npoptn_().
 * Instead of solving, it calls the funcon / funobj callbacks it was handed on
 * every coefficient vector of a batch supplied through shim_set_request() and
 * copies the results out.  It is linked into oracle/_ref/libntg_ref.so (to
 * drive the unmodified reference) and built stand-alone as
 * oracle/libnpsol_shim.so (to drive THIS repo's ntg() in the drop-in tests --
 * the same shim on both sides, so the two ntg() implementations see identical
 * calls).
 *
 * The Jacobian buffer handed to funcon can be pre-filled with NaN so entries
 * the callee never writes are distinguishable from written zeros.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "npsol_shim.h"


static shim_request *g_req;

void shim_set_request(shim_request *r) { g_req = r; }
shim_request *shim_get_request(void) { return g_req; }

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void npoptn_(char *option, long len)
{
    (void)option;
    (void)len;
}

typedef void (*funcon_t)(int *, int *, int *, int *, int *, double *, double *, double *, int *);
typedef void (*funobj_t)(int *, int *, double *, double *, double *, int *);

static long check_pattern(const shim_request *r, const double *cJac, int ldJ)
{
    long bad = 0;
    int row, col, j, k;
    char *inband = malloc((size_t)r->n);
    for (row = 0; row < r->ncnln; row++) {
        memset(inband, 0, (size_t)r->n);
        for (j = 0; j < r->nout; j++)
            for (k = 0; k < r->order[j]; k++)
                inband[r->col0[row * r->nout + j] + k] = 1;
        for (col = 0; col < r->n; col++) {
            int written = !isnan(cJac[(size_t)col * ldJ + row]);
            if (written != inband[col])
                bad++;
        }
    }
    free(inband);
    return bad;
}

void npsol_(int *n_, int *nclin_, int *ncnln_, int *ldA, int *ldJ_, int *ldR, double *A,
            double *bl, double *bu, funcon_t funcon, funobj_t funobj, int *inform,
            int *iter, int *istate, double *c, double *cJac, double *clambda, double *f,
            double *g, double *R, double *x, int *iw, int *leniw, double *w, int *lenw)
{
    shim_request *r = g_req;
    int n = *n_, nclin = *nclin_, ncnln = *ncnln_, ldJ = *ldJ_;
    int rep, p, i, j, k;
    int *needc;
    (void)ldR; (void)istate; (void)clambda; (void)R; (void)iw; (void)leniw; (void)w; (void)lenw;
    (void)ldA;
    *inform = 0;
    *iter = 0;
    if (r == NULL)
        return;
    r->calls++;
    r->n = n; r->nclin = nclin; r->ncnln = ncnln;
    if (r->A && nclin > 0)
        memcpy(r->A, A, sizeof(double) * (size_t)nclin * n);
    if (r->bl) memcpy(r->bl, bl, sizeof(double) * (size_t)(n + nclin + ncnln));
    if (r->bu) memcpy(r->bu, bu, sizeof(double) * (size_t)(n + nclin + ncnln));

    needc = malloc(sizeof(int) * (size_t)(ncnln > 0 ? ncnln : 1));
    for (i = 0; i < ncnln; i++) needc[i] = 1;

    if (ncnln > 0 && r->nan_fill && (r->Jdense || r->Jband))
        for (i = 0; i < ldJ * n; i++) cJac[i] = NAN;

    r->best_seconds = 1e300;
    r->pattern_bad = 0;
    for (rep = 0; rep < (r->reps > 0 ? r->reps : 1); rep++) {
        double acc = 0.0;
        for (p = 0; p < r->P; p++) {
            int nstate = (rep == 0 && p == 0) ? 1 : 0;
            int mode;
            double t0;
            memcpy(x, r->X + (size_t)p * n, sizeof(double) * (size_t)n);
            t0 = now_s();
            if (ncnln > 0 && r->mode_con >= 0) {
                mode = r->mode_con;
                funcon(&mode, &ncnln, &n, &ldJ, needc, x, c, cJac, &nstate);
            }
            if (r->mode_obj >= 0) {
                mode = r->mode_obj;
                funobj(&mode, &n, x, f, g, &nstate);
            }
            acc += now_s() - t0;
            if (rep > 0) continue;
            if (r->mode_obj >= 0) {
                if (r->f && r->mode_obj != 1) r->f[p] = *f;
                if (r->g && r->mode_obj != 0) memcpy(r->g + (size_t)p * n, g, sizeof(double) * (size_t)n);
            }
            if (ncnln > 0 && r->mode_con >= 0) {
                if (r->c && r->mode_con != 1)
                    memcpy(r->c + (size_t)p * ncnln, c, sizeof(double) * (size_t)ncnln);
                if (r->mode_con != 0) {
                    if (r->Jdense)
                        memcpy(r->Jdense + (size_t)p * ncnln * n, cJac,
                               sizeof(double) * (size_t)ncnln * n);
                    if (r->Jband) {
                        double *dst = r->Jband + (size_t)p * ncnln * r->S;
                        for (i = 0; i < ncnln; i++) {
                            int s = 0;
                            for (j = 0; j < r->nout; j++)
                                for (k = 0; k < r->order[j]; k++, s++)
                                    dst[(size_t)i * r->S + s] =
                                        cJac[(size_t)(r->col0[i * r->nout + j] + k) * ldJ + i];
                        }
                    }
                    if (r->nan_fill && (r->Jdense || r->Jband) && (p == 0 || p == r->P - 1))
                        r->pattern_bad += check_pattern(r, cJac, ldJ);
                }
            }
        }
        if (acc < r->best_seconds) r->best_seconds = acc;
    }
    free(needc);
}


/*
 * Run a whole user program (an NTG example's main(), renamed at compile time)
 * with the batch request installed: its ntg() call ends up in npsol_() above.
 * stdout is silenced for the duration (banner + the example's own printing).
 * Returns the number of times npsol_ was entered (1 for the shipped examples).
 */
int shim_run_main(int (*mainfn)(int, char **), int P, const double *X, int mode_obj, int mode_con,
                  double *f, double *g, double *c, double *Jdense, double *A, double *bl, double *bu,
                  int *dims_out /* n, nclin, ncnln */)
{
    shim_request req;
    char arg0[] = "ntg_example";
    char *argv[2] = {arg0, NULL};
    FILE *devnull = fopen("/dev/null", "w"), *keep = stdout;
    memset(&req, 0, sizeof req);
    req.P = P; req.X = X; req.mode_obj = mode_obj; req.mode_con = mode_con;
    req.f = f; req.g = g; req.c = c; req.Jdense = Jdense; req.A = A; req.bl = bl; req.bu = bu;
    req.reps = 1; req.nan_fill = 0;
    shim_set_request(&req);
    if (devnull) stdout = devnull;
    mainfn(1, argv);
    fflush(stdout);
    stdout = keep;
    if (devnull) fclose(devnull);
    shim_set_request(NULL);
    if (dims_out) { dims_out[0] = req.n; dims_out[1] = req.nclin; dims_out[2] = req.ncnln; }
    return req.calls;
}
