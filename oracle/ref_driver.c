/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Drives the UNMODIFIED reference C code (compiled from where it lies under
 * /root/reference/src by oracle/Makefile into oracle/_ref/libntg_ref.so) on a
 * batch of coefficient vectors.
 *
 * How: ref_eval() calls the reference's real ntg() (src/ntg.c:54).  ntg()
 * builds its tables and globals and then calls npsol_() -- which is licensed,
 * absent, and replaced here by a fake that, instead of solving, calls the
 * captured NPfuncon / NPfunobj pointers (src/ntg.c:250-253, :274-371) on every
 * vector of the batch and copies the results out.  It has to happen inside
 * npsol_ because ntg() frees all its state on return (src/ntg.c:255-266).
 *
 * The Jacobian buffer NPSOL hands to funcon is pre-filled with NaN so that the
 * entries the reference never writes are distinguishable from written zeros:
 * that is how the sparsity pattern is extracted bit-exactly.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ntg.h"      /* the reference's own header (-I/root/reference/src) */
#include "ntg_b200.h" /* ntgb_setup: same fields as ntg()'s arguments        */
#include "npsol_shim.h"

/* Jacobian row pattern from the reference's OWN tables (public symbols of
 * src/colloc.c): rows [0,nnlic) use column iC[j] (CollocConcatMultI,
 * src/colloc.c:254), trajectory row m*nbps+bp uses iC[j]+block[bp].offset
 * (:274-276), final rows use iC[j]+block[nbps-1].offset (:298). */
static void ref_pattern(const ntgb_setup *s, ConcatColloc *cc, int *col0)
{
    int row = 0, r, m, bp, j;
    for (r = 0; r < s->nnlic; r++, row++)
        for (j = 0; j < s->nout; j++) col0[row * s->nout + j] = cc->iC[j];
    for (m = 0; m < s->nnltc; m++)
        for (bp = 0; bp < s->nbps; bp++, row++)
            for (j = 0; j < s->nout; j++)
                col0[row * s->nout + j] = cc->iC[j] + cc->colloc[j]->block[bp].offset;
    for (r = 0; r < s->nnlfc; r++, row++)
        for (j = 0; j < s->nout; j++)
            col0[row * s->nout + j] = cc->iC[j] + cc->colloc[j]->block[s->nbps - 1].offset;
}

static ConcatColloc *make_cc(const ntgb_setup *s)
{
    return ConcatCollocMatrix(s->nout, (double **)s->knots, (int *)s->kninterv, (double *)s->bps,
                              s->nbps, (int *)s->maxderiv, (int *)s->order, (int *)s->mult);
}

/* sizes the reference derives (src/colloc.c:34-52, src/ntg.c:155-157) */
int ref_dims(const ntgb_setup *s, ntgb_dims *d)
{
    ConcatColloc *cc = make_cc(s);
    int j;
    memset(d, 0, sizeof *d);
    d->nout = s->nout; d->nbps = s->nbps;
    d->nC = cc->nC; d->nz = cc->nz; d->nZ = cc->nZ;
    d->nclin = s->nlic + s->nltc * s->nbps + s->nlfc;
    d->ncnln = s->nnlic + s->nnltc * s->nbps + s->nnlfc;
    for (j = 0; j < s->nout; j++) d->sorder += s->order[j];
    d->device = -1;
    FreeConcatColloc(cc);
    return 0;
}

/* tables exactly as the reference holds them: B index (bp*order+k)*maxderiv+d
 * = block[bp].matrix->elements[k][d] (src/colloc.c:99-101), offset, pattern */
int ref_tables(const ntgb_setup *s, double *B, int *offset, int *col0)
{
    ConcatColloc *cc = make_cc(s);
    int j, bp, k, d;
    size_t pos = 0;
    for (j = 0; j < s->nout; j++) {
        Colloc *co = cc->colloc[j];
        for (bp = 0; bp < s->nbps; bp++) {
            if (offset) offset[j * s->nbps + bp] = co->block[bp].offset;
            for (k = 0; k < co->order; k++)
                for (d = 0; d < co->maxderiv; d++, pos++)
                    if (B) B[pos] = co->block[bp].matrix->elements[k][d];
        }
    }
    if (col0) ref_pattern(s, cc, col0);
    FreeConcatColloc(cc);
    return 0;
}

/* updateZ (src/colloc.c:344-367) for one coefficient vector and one AV list */
int ref_updateZ(const ntgb_setup *s, const double *C, const AV *av, int nav, int type, double *Z)
{
    ConcatColloc *cc = make_cc(s);
    updateZ(Z, cc, (double *)C, (AV *)av, nav, type);
    FreeConcatColloc(cc);
    return 0;
}

/* SplineInterp (src/colloc.c:449-484) */
int ref_spline_interp(double *f, double x, const double *knots, int ninterv, const double *coefs,
                      int ncoefs, int order, int mult, int maxderiv)
{
    SplineInterp(f, x, (double *)knots, ninterv, (double *)coefs, ncoefs, order, mult, maxderiv);
    return 0;
}

void ref_linspace(double *v, double d0, double d1, int n) { linspace(v, d0, d1, n); }

/*
 * Batched evaluation through the reference's ntg() + NPfuncon/NPfunobj.
 * Any output pointer may be NULL.  Returns 0, fills *seconds with the best
 * wall time of one pass over the batch (funcon+funobj calls only).
 */
int ref_eval(const ntgb_setup *s, int P, const double *X, int mode_obj, int mode_con, double *f,
             double *g, double *c, double *Jdense, double *Jband, long *pattern_bad, double *A,
             double *bl, double *bu, int reps, double *seconds)
{
    shim_request req;
    ntgb_dims d;
    int *col0 = NULL;
    double *x;
    int *istate;
    double *clambda, *R;
    int inform = 0;
    double objective = 0.0;
    int saved_stdout_quiet = 1;
    (void)saved_stdout_quiet;

    ref_dims(s, &d);
    memset(&req, 0, sizeof req);
    req.P = P; req.X = X; req.mode_obj = mode_obj; req.mode_con = mode_con;
    req.f = f; req.g = g; req.c = c; req.Jdense = Jdense; req.Jband = Jband;
    req.A = A; req.bl = bl; req.bu = bu; req.reps = reps;
    req.nout = s->nout; req.order = s->order; req.S = d.sorder;
    if (d.ncnln > 0) {
        col0 = malloc(sizeof(int) * (size_t)d.ncnln * s->nout);
        ref_tables(s, NULL, NULL, col0);
        req.col0 = col0;
    }
    x = calloc((size_t)d.nC, sizeof(double));
    istate = calloc((size_t)(d.nC + d.nclin + d.ncnln), sizeof(int));
    clambda = calloc((size_t)(d.nC + d.nclin + d.ncnln), sizeof(double));
    R = calloc((size_t)(d.nC + 1) * (d.nC + 1), sizeof(double));

    req.nan_fill = 1;
    shim_set_request(&req);
    /* the reference prints its banner on every ntg() call (src/ntg.c:161):
     * silence stdout for the duration */
    {
        FILE *devnull = fopen("/dev/null", "w");
        FILE *keep = stdout;
        if (devnull) stdout = devnull;
        ntg(s->nout, (double *)s->bps, s->nbps, (int *)s->kninterv, (double **)s->knots,
            (int *)s->order, (int *)s->mult, (int *)s->maxderiv, x,
            s->nlic, (double **)s->lic, s->nltc, (double **)s->ltc, s->nlfc, (double **)s->lfc,
            s->nnlic, s->nlicf, s->nnltc, s->nltcf, s->nnlfc, s->nlfcf,
            s->ninitialconstrav, (AV *)s->initialconstrav,
            s->ntrajectoryconstrav, (AV *)s->trajectoryconstrav,
            s->nfinalconstrav, (AV *)s->finalconstrav,
            (double *)s->lowerb, (double *)s->upperb,
            s->nicf, s->icf, s->nucf, s->ucf, s->nfcf, s->fcf,
            s->ninitialcostav, (AV *)s->initialcostav,
            s->ntrajectorycostav, (AV *)s->trajectorycostav,
            s->nfinalcostav, (AV *)s->finalcostav,
            istate, clambda, R, &inform, &objective);
        stdout = keep;
        if (devnull) fclose(devnull);
    }
    shim_set_request(NULL);
    if (pattern_bad) *pattern_bad = req.pattern_bad;
    if (seconds) *seconds = req.best_seconds;
    free(col0); free(x); free(istate); free(clambda); free(R);
    return 0;
}
