/*
 * pgs_device.cuh -- the three spline routines NTG takes from de Boor's PGS,
 * as stateless __host__ __device__ functions (no SAVEd state, no common
 * block), used by K0 (one-time table build) and by the batched SplineInterp.
 *
 * What they replace: the Fortran calls at reference src/colloc.c:92-99,107
 * and :466-472 (knots_, interv_, bsplvd_ with bsplvb_ underneath).  The
 * arithmetic is de Boor's published recurrences evaluated in the published
 * order, every quantity a double (the reference builds PGS with
 * -fdefault-real-8, reference Makefile:19), and the translation unit is
 * compiled with -fmad=false, so tables are bit-identical to the CPU oracle's.
 */
#ifndef NTG_PGS_DEVICE_CUH_
#define NTG_PGS_DEVICE_CUH_

#include <cuda_runtime.h>

#define PGS_MAXK 20 /* bsplvb: jmax = 20 */

namespace ntgb {

/* augmented knot i (0-based) of the sequence `knots` builds: first break
 * `order` times, interior breaks (order-mult) times, last break `order` times */
__host__ __device__ inline double pgs_augknot(const double *brk, int ninterv, int order, int mult, int i)
{
    const int k = order - mult;
    const int n = ninterv * k + mult;
    if (i < order) return brk[0];
    if (i >= n) return brk[ninterv];
    return brk[1 + (i - order) / k];
}

/* interv, de Boor-site rule: 1-based left = max{ i : xt(i) < xt(lxt), xt(i) <= x };
 * x < xt(1) -> 1.  (mflag is ignored by NTG.) */
__host__ __device__ inline int pgs_interv(const double *xt, int lxt, double x)
{
    if (x < xt[0]) return 1;
    if (x >= xt[lxt - 1]) {
        int i = lxt;
        while (i > 1 && !(xt[i - 1] < xt[lxt - 1])) i--;
        return i;
    }
    int lo = 1, hi = lxt;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= xt[mid - 1]) lo = mid; else hi = mid;
    }
    return lo;
}

/* one Cox-de Boor order-raising sweep of bsplvb: from order j to j+1.
 * t is 0-based here: t[left + j - 1] is Fortran t(left+j). */
__host__ __device__ inline void pgs_raise(const double *t, int left, double x, int j, double *biatx,
                                          double *deltal, double *deltar)
{
    deltar[j - 1] = t[left + j - 1] - x;
    deltal[j - 1] = x - t[left - j];
    double saved = 0.0;
    for (int i = 1; i <= j; i++) {
        const double term = biatx[i - 1] / (deltar[i - 1] + deltal[j - i]);
        biatx[i - 1] = saved + deltar[i - 1] * term;
        saved = deltal[j - i] * term;
    }
    biatx[j] = saved;
}

/*
 * bsplvd: dbiatx(i,m) (column-major k x nderiv, index (m-1)*k + (i-1)) = the
 * (m-1)st derivative at x of the i-th of the k B-splines of order k that are
 * non-zero on [t(left), t(left+1)).  a is k*k scratch.
 */
__host__ __device__ inline void pgs_bsplvd(const double *t, int k, double x, int left, double *a,
                                           double *dbiatx, int nderiv)
{
    double deltal[PGS_MAXK], deltar[PGS_MAXK];
    int mhigh = nderiv < k ? nderiv : k;
    if (mhigh < 1) mhigh = 1;
    const int kp1 = k + 1;
    /* values of order kp1-mhigh in column 1 */
    int j = 1;
    dbiatx[0] = 1.0;
    while (j < kp1 - mhigh) { pgs_raise(t, left, x, j, dbiatx, deltal, deltar); j++; }
    if (mhigh == 1) return;
    /* save each lower order in a later column, then raise column 1 by one */
    int ideriv = mhigh;
    for (int m = 2; m <= mhigh; m++) {
        int src = 0;
        for (int r = ideriv; r <= k; r++) dbiatx[(ideriv - 1) * k + (r - 1)] = dbiatx[src++];
        ideriv--;
        pgs_raise(t, left, x, j, dbiatx, deltal, deltar);
        j++;
    }
    /* a = identity on the part that is read */
    for (int c = 0; c < k; c++)
        for (int r = 0; r < k; r++) a[c * k + r] = (r == c) ? 1.0 : 0.0;
    /* difference the B-coefficients, combine with the saved lower-order values */
    for (int m = 2; m <= mhigh; m++) {
        const int kp1mm = kp1 - m;
        const double fkp1mm = (double)kp1mm;
        int il = left, i = k;
        for (int ld = 1; ld <= kp1mm; ld++) {
            const double factor = fkp1mm / (t[il + kp1mm - 1] - t[il - 1]);
            for (int c = 1; c <= i; c++)
                a[(c - 1) * k + (i - 1)] = (a[(c - 1) * k + (i - 1)] - a[(c - 1) * k + (i - 2)]) * factor;
            il--;
            i--;
        }
        for (int c = 1; c <= k; c++) {
            double sum = 0.0;
            const int jlow = c > m ? c : m;
            for (int r = jlow; r <= k; r++) sum = a[(c - 1) * k + (r - 1)] * dbiatx[(m - 1) * k + (r - 1)] + sum;
            dbiatx[(m - 1) * k + (c - 1)] = sum;
        }
    }
}

} /* namespace ntgb */
#endif
