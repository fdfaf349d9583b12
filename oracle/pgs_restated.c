/*
 * oracle/pgs_restated.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C restatement of the three routines of C. de Boor's PGS / netlib "pppack"
 * that the reference calls, with the Fortran calling convention the reference
 * declares in /root/reference/src/colloc.h:31-40 (everything by pointer,
 * trailing underscore, common block /side/ as `struct side side_`).
 *
 * PGS is a third-party dependency that is NOT under /root/reference (it is
 * fetched by wget at build time, pgs/Makefile:6,75-77, not version-pinned) and
 * there is no Fortran compiler in this image.  The routines below restate the
 * published algorithms ("A Practical Guide to Splines", de Boor; pppack
 * knots.f / interv.f / bsplvb.f / bsplvd.f) with every REAL a double
 * (reference Makefile:19 builds PGS with -fdefault-real-8).
 *
 * PARITY UNPINNED at this boundary: the reference ships no golden vectors for
 * PGS output.  What pins this file instead (tests/test_oracle_pgs.py):
 *   - exact-rational Cox-de Boor recursion on the same double inputs,
 *   - partition-of-unity / derivative-sum identities,
 *   - off == left_aug - order for every breakpoint,
 *   - the known answers of SURVEY.md section 8(c).
 *
 * Call sites in the reference that this file serves:
 *   src/colloc.c:92-93  side_.m=mult; knots_(knots,&ninterv,&order,augknots,&n)
 *   src/colloc.c:98     interv_(augknots,&naugknots,&x,&left,&mflag)
 *   src/colloc.c:99     bsplvd_(augknots,&order,&x,&left,a,dbiatx,&maxderiv)
 *   src/colloc.c:107    interv_(knots,&nknots,&x,&left,&mflag)
 *   src/colloc.c:466-472 (SplineInterp, same four calls)
 */
#include <stddef.h>

struct side {
    int m;
    int iside;
    double xside[10];
} side_;

/* knots: break(1) kpm times, break(2..l) k=kpm-m times each, break(l+1) kpm
 * times; n = l*k + m.  (pppack knots.f) */
void knots_(double *brk, int *l_, int *kpm_, double *t, int *n_)
{
    int l = *l_, kpm = *kpm_;
    int m = side_.m;
    int k = kpm - m;
    int n = l * k + m;
    int jj = n + kpm + 1; /* 1-based cursor walking downwards */
    int jjj = l + 1;
    int ll, j;
    *n_ = n;
    for (ll = 1; ll <= kpm; ll++) {
        jj--;
        t[jj - 1] = brk[jjj - 1];
    }
    for (j = 1; j <= l; j++) {
        jjj--;
        for (ll = 1; ll <= k; ll++) {
            jj--;
            t[jj - 1] = brk[jjj - 1];
        }
    }
    for (ll = 1; ll <= kpm; ll++)
        t[ll - 1] = brk[0];
}

/* interv, de Boor-site version:
 *   left = max{ i : xt(i) < xt(lxt) and xt(i) <= x }  (1-based)
 *   x <  xt(1)               -> left = 1, mflag = -1
 *   xt(1) <= x < xt(lxt)     -> mflag = 0
 *   x == xt(lxt)             -> mflag = 0
 *   x >  xt(lxt)             -> mflag = 1
 * The saved `ilo` of the Fortran only affects search speed, never the result,
 * so a plain bisection is an exact restatement of the outputs. */
void interv_(double *xt, int *lxt_, double *x_, int *left, int *mflag)
{
    int lxt = *lxt_;
    double x = *x_;
    int lo, hi, mid;
    if (x < xt[0]) {
        *left = 1;
        *mflag = -1;
        return;
    }
    if (x >= xt[lxt - 1]) {
        int i = lxt;
        *mflag = (x == xt[lxt - 1]) ? 0 : 1;
        while (i > 1 && !(xt[i - 1] < xt[lxt - 1]))
            i--;
        *left = i;
        return;
    }
    /* xt(1) <= x < xt(lxt): largest i with xt(i) <= x */
    lo = 1;
    hi = lxt; /* invariant xt(lo) <= x < xt(hi) */
    while (hi - lo > 1) {
        mid = (lo + hi) / 2;
        if (x >= xt[mid - 1])
            lo = mid;
        else
            hi = mid;
    }
    *left = lo;
    *mflag = 0;
}

/* bsplvb with its SAVEd state (j, deltal, deltar) made file-static, exactly as
 * the Fortran keeps it between the index=1 and index=2 calls of bsplvd. */
#define PGS_JMAX 20
static int sv_j = 1;
static double sv_deltal[PGS_JMAX], sv_deltar[PGS_JMAX];

static void bsplvb(const double *t, int jhigh, int index, double x, int left,
                   double *biatx)
{
    int i, jp1;
    double saved, term;
    if (index == 1) {
        sv_j = 1;
        biatx[0] = 1.0;
        if (sv_j >= jhigh)
            return;
    }
    do {
        jp1 = sv_j + 1;
        sv_deltar[sv_j - 1] = t[left + sv_j - 1] - x;
        sv_deltal[sv_j - 1] = x - t[left + 1 - sv_j - 1];
        saved = 0.0;
        for (i = 1; i <= sv_j; i++) {
            term = biatx[i - 1] / (sv_deltar[i - 1] + sv_deltal[jp1 - i - 1]);
            biatx[i - 1] = saved + sv_deltar[i - 1] * term;
            saved = sv_deltal[jp1 - i - 1] * term;
        }
        biatx[jp1 - 1] = saved;
        sv_j = jp1;
    } while (sv_j < jhigh);
}

/* bsplvd: values and derivatives of the k B-splines that are non-zero at x.
 * a(k,k), dbiatx(k,nderiv) are column-major (Fortran). */
void bsplvd_(double *t, int *k_, double *x_, int *left_, double *a,
             double *dbiatx, int *nderiv_)
{
    int k = *k_, left = *left_, nderiv = *nderiv_;
    double x = *x_;
    int mhigh, kp1, ideriv, m, j, jp1mid, jlow, i, il, kp1mm, ldummy;
    double factor, fkp1mm, sum;
#define A(i, j) a[((j)-1) * k + ((i)-1)]
#define DB(i, j) dbiatx[((j)-1) * k + ((i)-1)]
    mhigh = nderiv < k ? nderiv : k;
    if (mhigh < 1)
        mhigh = 1;
    kp1 = k + 1;
    bsplvb(t, kp1 - mhigh, 1, x, left, dbiatx);
    if (mhigh == 1)
        return;
    ideriv = mhigh;
    for (m = 2; m <= mhigh; m++) {
        jp1mid = 1;
        for (j = ideriv; j <= k; j++) {
            DB(j, ideriv) = DB(jp1mid, 1);
            jp1mid++;
        }
        ideriv--;
        bsplvb(t, kp1 - ideriv, 2, x, left, dbiatx);
    }
    jlow = 1;
    for (i = 1; i <= k; i++) {
        for (j = jlow; j <= k; j++)
            A(j, i) = 0.0;
        jlow = i;
        A(i, i) = 1.0;
    }
    for (m = 2; m <= mhigh; m++) {
        kp1mm = kp1 - m;
        fkp1mm = (double)kp1mm;
        il = left;
        i = k;
        for (ldummy = 1; ldummy <= kp1mm; ldummy++) {
            factor = fkp1mm / (t[il + kp1mm - 1] - t[il - 1]);
            for (j = 1; j <= i; j++)
                A(i, j) = (A(i, j) - A(i - 1, j)) * factor;
            il--;
            i--;
        }
        for (i = 1; i <= k; i++) {
            sum = 0.0;
            jlow = i > m ? i : m;
            for (j = jlow; j <= k; j++)
                sum = A(j, i) * DB(j, m) + sum;
            DB(i, m) = sum;
        }
    }
#undef A
#undef DB
}
