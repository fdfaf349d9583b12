/*
 * oracle/ntg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A CPU restatement, in plain C, of the algorithm behind the reference's
 * NPfunobj / NPfuncon (src/ntg.c:274-371).  It is written band-only (no dense
 * nbps x nC or (nbps*nnltc) x nZ scratch matrices) but performs every floating
 * point operation the reference performs, in the reference's order, so its
 * results are bit-identical to the reference's -- tests/test_oracle_golden.py
 * checks exactly that against oracle/_ref (the unmodified reference sources)
 * and against the committed fixtures under tests/golden/ generated from it.
 *
 * PGS (knots/interv/bsplvd) comes from oracle/pgs_restated.c; at that boundary
 * parity is unpinned by the reference (no vectors shipped) and is pinned by
 * exact-rational B-spline evaluation instead (tests/test_oracle_pgs.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load this library.  Each function cites the reference lines it restates.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ntg_b200.h"

extern struct side { int m; int iside; double xside[10]; } side_;
void knots_(double *, int *, int *, double *, int *);
void interv_(double *, int *, double *, int *, int *);
void bsplvd_(double *, int *, double *, int *, double *, double *, int *);

typedef struct {
    int nout, nbps, nC, nz, nZ, S;
    int *order, *mult, *maxderiv, *ncoef;
    int *iC, *iz, *iZ, *jk0;
    double **B;  /* per output: [(bp*order + k)*maxderiv + d]   (colloc.c:99-101) */
    int **off;   /* per output: [bp]                            (colloc.c:104-111) */
    int **left;  /* per output: [bp] interv on augmented knots, 1-based            */
    const double *bps;
} tables_t;

/* ConcatCollocMatrix + CollocMatrix, src/colloc.c:15-117 */
static tables_t *tables_build(const ntgb_setup *s)
{
    tables_t *t = calloc(1, sizeof *t);
    int j, bp, k, d;
    t->nout = s->nout; t->nbps = s->nbps; t->bps = s->bps;
    t->order = malloc(sizeof(int) * s->nout); t->mult = malloc(sizeof(int) * s->nout);
    t->maxderiv = malloc(sizeof(int) * s->nout); t->ncoef = malloc(sizeof(int) * s->nout);
    t->iC = malloc(sizeof(int) * s->nout); t->iz = malloc(sizeof(int) * s->nout);
    t->iZ = malloc(sizeof(int) * s->nout); t->jk0 = malloc(sizeof(int) * s->nout);
    t->B = malloc(sizeof(double *) * s->nout); t->off = malloc(sizeof(int *) * s->nout);
    t->left = malloc(sizeof(int *) * s->nout);
    for (j = 0; j < s->nout; j++) {
        int order = s->order[j], mult = s->mult[j], md = s->maxderiv[j], ninterv = s->kninterv[j];
        int n = ninterv * (order - mult) + mult; /* colloc.c:67 */
        int nknots = ninterv + 1, naug = n + order, nn = 0, lft, mflag;
        double *aug = malloc(sizeof(double) * naug);
        double *a = malloc(sizeof(double) * order * order);
        double *dbiatx = calloc((size_t)order * md, sizeof(double));
        t->order[j] = order; t->mult[j] = mult; t->maxderiv[j] = md; t->ncoef[j] = n;
        t->iC[j] = t->nC; t->iz[j] = t->nz; t->iZ[j] = t->nz * s->nbps; t->jk0[j] = t->S;
        t->nC += n; t->nz += md; t->S += order;
        t->B[j] = malloc(sizeof(double) * (size_t)s->nbps * order * md);
        t->off[j] = malloc(sizeof(int) * s->nbps);
        t->left[j] = malloc(sizeof(int) * s->nbps);
        side_.m = mult;                                     /* colloc.c:92 */
        knots_((double *)s->knots[j], &ninterv, &order, aug, &nn);
        for (bp = 0; bp < s->nbps; bp++) {
            double x = s->bps[bp];
            interv_(aug, &naug, &x, &lft, &mflag);          /* colloc.c:98 */
            bsplvd_(aug, &order, &x, &lft, a, dbiatx, &md); /* colloc.c:99 */
            t->left[j][bp] = lft;
            for (k = 0; k < order; k++)                     /* FTranspose, colloc.c:101 */
                for (d = 0; d < md; d++)
                    t->B[j][((size_t)bp * order + k) * md + d] = dbiatx[d * order + k];
            interv_((double *)s->knots[j], &nknots, &x, &lft, &mflag); /* colloc.c:107 */
            t->off[j][bp] = (lft - 1) * (order - mult);
        }
        free(aug); free(a); free(dbiatx);
    }
    t->nZ = t->nz * s->nbps;
    return t;
}

static void tables_free(tables_t *t)
{
    int j;
    for (j = 0; j < t->nout; j++) { free(t->B[j]); free(t->off[j]); free(t->left[j]); }
    free(t->order); free(t->mult); free(t->maxderiv); free(t->ncoef); free(t->iC); free(t->iz);
    free(t->iZ); free(t->jk0); free(t->B); free(t->off); free(t->left); free(t);
}

#define TB(t, j, bp, k, d) ((t)->B[j][((size_t)(bp) * (t)->order[j] + (k)) * (t)->maxderiv[j] + (d)])

/* Zvalue + odb2lin + updateZ, src/colloc.c:318-367 */
static void update_Z(double *Z, const tables_t *t, const double *C, const AV *av, int nav, int type)
{
    int i, bp, k, b0, b1;
    if (type == AVINITIAL) { b0 = 0; b1 = 1; }
    else if (type == AVTRAJECTORY) { b0 = 0; b1 = t->nbps; }
    else { b0 = t->nbps - 1; b1 = t->nbps; }
    for (i = 0; i < nav; i++) {
        int j = av[i].output, d = av[i].deriv;
        for (bp = b0; bp < b1; bp++) {
            double acc = 0.0;
            for (k = 0; k < t->order[j]; k++)
                acc += TB(t, j, bp, k, d) * C[t->iC[j] + t->off[j][bp] + k];
            Z[t->iZ[j] + t->maxderiv[j] * bp + d] = acc;
        }
    }
}

/* Z2zpI / Z2zpT / Z2zpF, src/colloc.c:425-447: all three are &Z[iZ[j] + bp*maxderiv_j] */
static void make_zp(double **zp, double *Z, const tables_t *t, int bp)
{
    int j;
    for (j = 0; j < t->nout; j++) zp[j] = &Z[t->iZ[j] + bp * t->maxderiv[j]];
}

/* one band row: out[jk0[j]+k] = sum_l dz[iz[j]+l] * B_j[bp][k][l], l ascending from 0.0
 * (CollocConcatMultI/T/F, src/colloc.c:243-316; cost.c:124-129) */
static void band_row(double *out, const double *dz, const tables_t *t, int bp)
{
    int j, k, l;
    for (j = 0; j < t->nout; j++)
        for (k = 0; k < t->order[j]; k++) {
            double acc = 0.0;
            for (l = 0; l < t->maxderiv[j]; l++) acc += dz[t->iz[j] + l] * TB(t, j, bp, k, l);
            out[t->jk0[j] + k] = acc;
        }
}

/* first column of output j's band for a row evaluated at breakpoint bp;
 * the "initial" variant ignores block[0].offset (src/colloc.c:254) */
static int band_col0(const tables_t *t, int j, int bp, int initial)
{
    return t->iC[j] + (initial ? 0 : t->off[j][bp]);
}

typedef struct {
    const ntgb_setup *s;
    tables_t *t;
    double *Z;        /* persists across calls like GZ (src/ntg.c:18,119) */
    double **zp;
    double *fbp;      /* [nbps] */
    double *dz;       /* [nbps][nz] cost derivatives per breakpoint */
    double *D;        /* [nbps][S] band of dIdC */
    double *dzc;      /* [maxcon][nz] constraint derivative scratch */
    double **dzcp;
    double *tmp;
    double *dI, *dIn, *dF;
} work_t;

static work_t *work_new(const ntgb_setup *s)
{
    work_t *w = calloc(1, sizeof *w);
    int maxcon = s->nnlic, i;
    tables_t *t = tables_build(s);
    if (s->nnltc > maxcon) maxcon = s->nnltc;
    if (s->nnlfc > maxcon) maxcon = s->nnlfc;
    if (maxcon < 1) maxcon = 1;
    w->s = s; w->t = t;
    w->Z = calloc((size_t)t->nZ, sizeof(double));
    w->zp = malloc(sizeof(double *) * t->nout);
    w->fbp = calloc((size_t)t->nbps, sizeof(double));
    w->dz = calloc((size_t)t->nbps * t->nz, sizeof(double));
    w->D = calloc((size_t)t->nbps * t->S, sizeof(double));
    w->dzc = calloc((size_t)maxcon * t->nz, sizeof(double));
    w->dzcp = malloc(sizeof(double *) * maxcon);
    for (i = 0; i < maxcon; i++) w->dzcp[i] = w->dzc + (size_t)i * t->nz;
    w->tmp = calloc((size_t)maxcon, sizeof(double));
    w->dI = calloc((size_t)t->nC, sizeof(double));
    w->dIn = calloc((size_t)t->nC, sizeof(double));
    w->dF = calloc((size_t)t->nC, sizeof(double));
    return w;
}

static void work_free(work_t *w)
{
    tables_free(w->t);
    free(w->Z); free(w->zp); free(w->fbp); free(w->dz); free(w->D); free(w->dzc); free(w->dzcp);
    free(w->tmp); free(w->dI); free(w->dIn); free(w->dF); free(w);
}

/* NPfunobj, src/ntg.c:274-335 with InitialCost / IntegratedCost / FinalCost (src/cost.c) */
static void funobj(work_t *w, int mode, int nstate, const double *x, double *y, double *yprime)
{
    const ntgb_setup *s = w->s;
    tables_t *t = w->t;
    double I = 0.0, In = 0.0, F = 0.0;
    int wantd = (mode == 1 || mode == 2);
    int bp, j, k, c, q;
    /* mode 0 gates on count==1, modes 1/2 on count!=0 (src/ntg.c:297-302 vs :309-314) */
    int doI = mode == 0 ? s->nicf == 1 : s->nicf != 0;
    int doU = mode == 0 ? s->nucf == 1 : s->nucf != 0;
    int doF = mode == 0 ? s->nfcf == 1 : s->nfcf != 0;
    if (mode < 0 || mode > 2) return;
    if (s->nicf != 0) update_Z(w->Z, t, x, s->initialcostav, s->ninitialcostav, AVINITIAL);
    if (s->nucf != 0) update_Z(w->Z, t, x, s->trajectorycostav, s->ntrajectorycostav, AVTRAJECTORY);
    if (s->nfcf != 0) update_Z(w->Z, t, x, s->finalcostav, s->nfinalcostav, AVFINAL);
    if (wantd) {
        memset(w->dI, 0, sizeof(double) * t->nC);
        memset(w->dIn, 0, sizeof(double) * t->nC);
        memset(w->dF, 0, sizeof(double) * t->nC);
    }
    if (doI) { /* InitialCost, src/cost.c:4-36 */
        double *dz = w->dz, band[64 * 8];
        make_zp(w->zp, w->Z, t, 0);
        memset(dz, 0, sizeof(double) * t->nz);
        s->icf(&mode, &nstate, &I, wantd ? dz : NULL, w->zp);
        if (wantd) {
            band_row(band, dz, t, 0);
            for (j = 0; j < t->nout; j++)
                for (k = 0; k < t->order[j]; k++)
                    w->dI[band_col0(t, j, 0, 1) + k] = band[t->jk0[j] + k];
        }
    }
    if (doU) { /* IntegratedCost, src/cost.c:38-139 */
        for (bp = 0; bp < t->nbps; bp++) {
            int i = bp;
            make_zp(w->zp, w->Z, t, bp);
            s->ucf(&mode, &nstate, &i, mode == 1 ? NULL : &w->fbp[bp],
                   wantd ? &w->dz[(size_t)bp * t->nz] : NULL, w->zp);
            if (wantd) band_row(&w->D[(size_t)bp * t->S], &w->dz[(size_t)bp * t->nz], t, bp);
        }
        if (mode != 1) { /* IntegrateVector TRAPEZOID, src/integrator.c:21-24 */
            for (bp = 0; bp < t->nbps - 1; bp++)
                In += (t->bps[bp + 1] - t->bps[bp]) * (w->fbp[bp + 1] + w->fbp[bp]) / 2;
        }
        if (wantd) { /* IntegrateFMatrixCols TRAPEZOID over the dense nbps x nC matrix,
                      * src/integrator.c:44-48; out-of-band entries are exact zeros */
            for (j = 0; j < t->nout; j++)
                for (c = 0; c < t->ncoef[j]; c++) {
                    double acc = 0.0;
                    for (bp = 0; bp < t->nbps - 1; bp++) {
                        int k1 = c - t->off[j][bp + 1], k0 = c - t->off[j][bp];
                        double d1 = (k1 >= 0 && k1 < t->order[j]) ? w->D[(size_t)(bp + 1) * t->S + t->jk0[j] + k1] : 0.0;
                        double d0 = (k0 >= 0 && k0 < t->order[j]) ? w->D[(size_t)bp * t->S + t->jk0[j] + k0] : 0.0;
                        acc += (t->bps[bp + 1] - t->bps[bp]) * (d1 + d0) / 2;
                    }
                    w->dIn[t->iC[j] + c] = acc;
                }
        }
    }
    if (doF) { /* FinalCost, src/cost.c:141-174 */
        double *dz = w->dz, band[64 * 8];
        int last = t->nbps - 1;
        make_zp(w->zp, w->Z, t, last);
        memset(dz, 0, sizeof(double) * t->nz);
        s->fcf(&mode, &nstate, &F, wantd ? dz : NULL, w->zp);
        if (wantd) {
            band_row(band, dz, t, last);
            for (j = 0; j < t->nout; j++)
                for (k = 0; k < t->order[j]; k++)
                    w->dF[band_col0(t, j, last, 0) + k] = band[t->jk0[j] + k];
        }
    }
    if (mode != 1) *y = I + In + F;                                   /* src/ntg.c:303,328 */
    if (wantd)
        for (q = 0; q < t->nC; q++) yprime[q] = w->dI[q] + w->dIn[q] + w->dF[q]; /* Vector3Add */
}

/* NPfuncon, src/ntg.c:337-371 with NonLinearConstraints (src/constraints.c:36-195).
 * Jb is the band [ncnln][S] (row-major), or NULL. */
static void funcon(work_t *w, int mode, int nstate, const double *x, double *cvec, double *Jb)
{
    const ntgb_setup *s = w->s;
    tables_t *t = w->t;
    int wantd = (mode == 1 || mode == 2), wantc = (mode != 1);
    int row = 0, r, m, bp;
    if (mode < 0 || mode > 2) return;
    if (s->nnlic != 0) update_Z(w->Z, t, x, s->initialconstrav, s->ninitialconstrav, AVINITIAL);
    if (s->nnltc != 0) update_Z(w->Z, t, x, s->trajectoryconstrav, s->ntrajectoryconstrav, AVTRAJECTORY);
    if (s->nnlfc != 0) update_Z(w->Z, t, x, s->finalconstrav, s->nfinalconstrav, AVFINAL);
    if (s->nnlic != 0) { /* NonLinearInitialConstraints, src/constraints.c:88-117 */
        memset(w->dzc, 0, sizeof(double) * s->nnlic * t->nz);
        make_zp(w->zp, w->Z, t, 0);
        s->nlicf(&mode, &nstate, wantc ? cvec : w->tmp, wantd ? w->dzcp : NULL, w->zp);
        if (wantd && Jb)
            for (r = 0; r < s->nnlic; r++) band_row(Jb + (size_t)r * t->S, w->dzcp[r], t, 0);
        row = s->nnlic;
    }
    if (s->nnltc != 0) { /* NonLinearTrajectoryConstraints, src/constraints.c:120-162 */
        /* dIdz is allocated (zeroed) once per call and NOT re-zeroed per breakpoint (:146-155) */
        memset(w->dzc, 0, sizeof(double) * s->nnltc * t->nz);
        for (bp = 0; bp < t->nbps; bp++) {
            int i = bp;
            make_zp(w->zp, w->Z, t, bp);
            s->nltcf(&mode, &nstate, &i, w->tmp, wantd ? w->dzcp : NULL, w->zp);
            for (m = 0; m < s->nnltc; m++) {
                if (wantc) cvec[row + m * t->nbps + bp] = w->tmp[m];
                if (wantd && Jb)
                    band_row(Jb + (size_t)(row + m * t->nbps + bp) * t->S, w->dzcp[m], t, bp);
            }
        }
        row += s->nnltc * t->nbps;
    }
    if (s->nnlfc != 0) { /* NonLinearFinalConstraints, src/constraints.c:165-195 */
        int last = t->nbps - 1;
        memset(w->dzc, 0, sizeof(double) * s->nnlfc * t->nz);
        make_zp(w->zp, w->Z, t, last);
        s->nlfcf(&mode, &nstate, wantc ? cvec + row : w->tmp, wantd ? w->dzcp : NULL, w->zp);
        if (wantd && Jb)
            for (r = 0; r < s->nnlfc; r++)
                band_row(Jb + (size_t)(row + r) * t->S, w->dzcp[r], t, last);
    }
}

/* row -> (breakpoint, initial?) and pattern, src/colloc.c:243-316 */
static void pattern(const ntgb_setup *s, const tables_t *t, int *col0)
{
    int row = 0, r, m, bp, j;
    for (r = 0; r < s->nnlic; r++, row++)
        for (j = 0; j < t->nout; j++) col0[row * t->nout + j] = band_col0(t, j, 0, 1);
    for (m = 0; m < s->nnltc; m++)
        for (bp = 0; bp < t->nbps; bp++, row++)
            for (j = 0; j < t->nout; j++) col0[row * t->nout + j] = band_col0(t, j, bp, 0);
    for (r = 0; r < s->nnlfc; r++, row++)
        for (j = 0; j < t->nout; j++) col0[row * t->nout + j] = band_col0(t, j, t->nbps - 1, 0);
}

/* ---- exported drivers (same signatures as oracle/ref_driver.c's ref_*) ---- */

int port_dims(const ntgb_setup *s, ntgb_dims *d)
{
    tables_t *t = tables_build(s);
    memset(d, 0, sizeof *d);
    d->nout = s->nout; d->nbps = s->nbps; d->nC = t->nC; d->nz = t->nz; d->nZ = t->nZ;
    d->nclin = s->nlic + s->nltc * s->nbps + s->nlfc;
    d->ncnln = s->nnlic + s->nnltc * s->nbps + s->nnlfc;
    d->sorder = t->S; d->device = -1;
    tables_free(t);
    return 0;
}

int port_tables(const ntgb_setup *s, double *B, int *offset, int *col0)
{
    tables_t *t = tables_build(s);
    int j;
    size_t pos = 0;
    for (j = 0; j < t->nout; j++) {
        size_t n = (size_t)t->nbps * t->order[j] * t->maxderiv[j];
        if (B) memcpy(B + pos, t->B[j], sizeof(double) * n);
        if (offset) memcpy(offset + (size_t)j * t->nbps, t->off[j], sizeof(int) * t->nbps);
        pos += n;
    }
    if (col0) pattern(s, t, col0);
    tables_free(t);
    return 0;
}

int port_left(const ntgb_setup *s, int *left)
{
    tables_t *t = tables_build(s);
    int j;
    for (j = 0; j < t->nout; j++) memcpy(left + (size_t)j * t->nbps, t->left[j], sizeof(int) * t->nbps);
    tables_free(t);
    return 0;
}

int port_updateZ(const ntgb_setup *s, const double *C, const AV *av, int nav, int type, double *Z)
{
    tables_t *t = tables_build(s);
    update_Z(Z, t, C, av, nav, type);
    tables_free(t);
    return 0;
}

/* SplineInterp, src/colloc.c:449-484 */
int port_spline_interp(double *f, double x, const double *knots, int ninterv, const double *coefs,
                       int ncoefs, int order, int mult, int maxderiv)
{
    int nknots = ninterv + 1, n = ninterv * (order - mult) + mult, naug = n + order, nn = 0;
    int left1, left2, mflag, i, j, offset;
    double *aug = malloc(sizeof(double) * naug), *a = malloc(sizeof(double) * order * order);
    double *dbiatx = calloc((size_t)order * maxderiv, sizeof(double));
    if (n != ncoefs) { free(aug); free(a); free(dbiatx); return -1; }
    side_.m = mult;
    knots_((double *)knots, &ninterv, &order, aug, &nn);
    interv_(aug, &naug, &x, &left1, &mflag);
    bsplvd_(aug, &order, &x, &left1, a, dbiatx, &maxderiv);
    interv_((double *)knots, &nknots, &x, &left2, &mflag);
    offset = (left2 - 1) * (order - mult);
    for (i = 0; i < maxderiv; i++) {
        f[i] = 0.0;
        for (j = 0; j < order; j++) f[i] += dbiatx[i * order + j] * coefs[offset + j];
    }
    free(aug); free(a); free(dbiatx);
    return 0;
}

/* linspace, src/ntg.c:374-389 */
void port_linspace(double *v, double d0, double d1, int n)
{
    int i;
    double step;
    if (d0 == d1) { for (i = 0; i < n; i++) v[i] = d0; return; }
    step = (d1 - d0) / (n - 1);
    v[0] = d0;
    for (i = 1; i < n; i++) v[i] = v[i - 1] + step;
}

/* bounds, src/constraints.c:5-33 */
static void expand_bounds(double *bbar, const double *b, const ntgb_setup *s, int nC, double big)
{
    int i, j, pos = 0, src = 0, nb = s->nbps;
    for (i = 0; i < nC; i++) bbar[pos++] = big;
    for (i = 0; i < s->nlic; i++) bbar[pos++] = b[src++];
    for (i = 0; i < s->nltc; i++, src++) for (j = 0; j < nb; j++) bbar[pos++] = b[src];
    for (i = 0; i < s->nlfc; i++) bbar[pos++] = b[src++];
    for (i = 0; i < s->nnlic; i++) bbar[pos++] = b[src++];
    for (i = 0; i < s->nnltc; i++, src++) for (j = 0; j < nb; j++) bbar[pos++] = b[src];
    for (i = 0; i < s->nnlfc; i++) bbar[pos++] = b[src++];
}

/* LinearConstraintsMatrix, src/constraints.c:198-261: the nonlinear machinery
 * with constant derivatives lic/ltc/lfc.  A is column-major nclin x nC. */
static void linear_matrix(const ntgb_setup *s, const tables_t *t, double *A)
{
    int nclin = s->nlic + s->nltc * s->nbps + s->nlfc;
    int row = 0, r, bp, j, k;
    double band[64 * 8];
    memset(A, 0, sizeof(double) * (size_t)nclin * t->nC);
    for (r = 0; r < s->nlic; r++, row++) {
        band_row(band, s->lic[r], t, 0);
        for (j = 0; j < t->nout; j++)
            for (k = 0; k < t->order[j]; k++)
                A[(size_t)(band_col0(t, j, 0, 1) + k) * nclin + row] = band[t->jk0[j] + k];
    }
    for (r = 0; r < s->nltc; r++)
        for (bp = 0; bp < s->nbps; bp++, row++) {
            band_row(band, s->ltc[r], t, bp);
            for (j = 0; j < t->nout; j++)
                for (k = 0; k < t->order[j]; k++)
                    A[(size_t)(band_col0(t, j, bp, 0) + k) * nclin + row] = band[t->jk0[j] + k];
        }
    for (r = 0; r < s->nlfc; r++, row++) {
        band_row(band, s->lfc[r], t, s->nbps - 1);
        for (j = 0; j < t->nout; j++)
            for (k = 0; k < t->order[j]; k++)
                A[(size_t)(band_col0(t, j, s->nbps - 1, 0) + k) * nclin + row] = band[t->jk0[j] + k];
    }
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int port_eval(const ntgb_setup *s, int P, const double *X, int mode_obj, int mode_con, double *f,
              double *g, double *c, double *Jdense, double *Jband, long *pattern_bad, double *A,
              double *bl, double *bu, int reps, double *seconds)
{
    work_t *w = work_new(s);
    tables_t *t = w->t;
    int nC = t->nC, S = t->S;
    int nclin = s->nlic + s->nltc * s->nbps + s->nlfc;
    int ncnln = s->nnlic + s->nnltc * s->nbps + s->nnlfc;
    int *col0 = malloc(sizeof(int) * (size_t)(ncnln > 0 ? ncnln : 1) * t->nout);
    double *cbuf = calloc((size_t)(ncnln > 0 ? ncnln : 1), sizeof(double));
    double *gbuf = calloc((size_t)nC, sizeof(double));
    double *jbuf = calloc((size_t)(ncnln > 0 ? ncnln : 1) * S, sizeof(double));
    double best = 1e300, fval = 0.0;
    int rep, p, r, j, k;
    pattern(s, t, col0);
    if (A && nclin > 0) linear_matrix(s, t, A);
    if (bl) expand_bounds(bl, s->lowerb, s, nC, -1.7976931348623157e308);
    if (bu) expand_bounds(bu, s->upperb, s, nC, 1.7976931348623157e308);
    for (rep = 0; rep < (reps > 0 ? reps : 1); rep++) {
        double acc = 0.0;
        for (p = 0; p < P; p++) {
            int nstate = (rep == 0 && p == 0) ? 1 : 0;
            const double *x = X + (size_t)p * nC;
            double t0 = now_s();
            if (ncnln > 0 && mode_con >= 0) funcon(w, mode_con, nstate, x, cbuf, jbuf);
            if (mode_obj >= 0) funobj(w, mode_obj, nstate, x, &fval, gbuf);
            acc += now_s() - t0;
            if (rep > 0) continue;
            if (mode_obj >= 0) {
                if (f && mode_obj != 1) f[p] = fval;
                if (g && mode_obj != 0) memcpy(g + (size_t)p * nC, gbuf, sizeof(double) * nC);
            }
            if (ncnln > 0 && mode_con >= 0) {
                if (c && mode_con != 1) memcpy(c + (size_t)p * ncnln, cbuf, sizeof(double) * ncnln);
                if (mode_con != 0) {
                    if (Jband) memcpy(Jband + (size_t)p * ncnln * S, jbuf, sizeof(double) * ncnln * S);
                    if (Jdense) { /* caller pre-fills with NaN; band entries only, like the reference */
                        double *dst = Jdense + (size_t)p * ncnln * nC;
                        for (r = 0; r < ncnln; r++)
                            for (j = 0; j < t->nout; j++)
                                for (k = 0; k < t->order[j]; k++)
                                    dst[(size_t)(col0[r * t->nout + j] + k) * ncnln + r] =
                                        jbuf[(size_t)r * S + t->jk0[j] + k];
                    }
                }
            }
        }
        if (acc < best) best = acc;
    }
    if (pattern_bad) *pattern_bad = 0;
    if (seconds) *seconds = best;
    free(col0); free(cbuf); free(gbuf); free(jbuf);
    work_free(w);
    return 0;
}
