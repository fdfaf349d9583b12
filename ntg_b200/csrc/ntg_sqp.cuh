/*
 * ntg_sqp.cuh -- the per-problem part of the batched SQP solver (ntgb_solve_sqp):
 * the role NPSOL plays for the reference (/root/reference/src/ntg.c:250-253 hands funobj / funcon to
 * npsol_(), an SQP method with a dense active-set QP subproblem and a quasi-Newton Hessian).
 *
 * One CTA works on one problem, everything in shared memory:
 *   reduced problem    C = Cpart + N y (linear equalities eliminated, ntg_core.cu reduce_linear),
 *                      rows h = [A_in C ; c(C)] with hl <= h <= hu;
 *   QP subproblem      min gr'd + 1/2 d'B d  s.t.  hl - h <= Ar d <= hu - h   (Ar = dh/dy), solved by the
 *                      dual active-set method of Goldfarb and Idnani (1983): Cholesky factor of B,
 *                      J = L^-T, active normals kept as J'N = [R; 0] with Givens rotations -- exact
 *                      multipliers and the active set (NPSOL's clambda / istate) come with it;
 *   inconsistent rows  when the linearised rows admit no point, the rows violated NOW leave the
 *                      constraint set and enter the objective as rho/2 (a_i d - r_i)^2: a regularised
 *                      Gauss-Newton step on the violation (B + rho Av'Av) d = -(gr - rho Av'r);
 *   merit              L1: f + nu * sum(violation), nu >= the largest multiplier seen (SLSQP's rule);
 *   Hessian            damped BFGS (Powell) on the reduced Lagrangian, skipped after a Gauss-Newton step.
 *
 * The code between NTG_SQP_CORE_BEGIN / END is plain C++ over a small "cooperative group" (tid, nt,
 * sync): nvcc compiles it for a CTA, and tests/tools/sqp_host.cpp compiles THE SAME TEXT with g++ for
 * a group of one thread so that the algebra can be checked against tests/tools/sqp_reference.py
 * without a GPU (test infrastructure only: the library never runs it on the host).
 */
#ifndef NTG_SQP_CUH_
#define NTG_SQP_CUH_

#include <math.h>

#ifdef __CUDACC__
#define NTG_HD __host__ __device__
#else
#define NTG_HD
#endif

namespace ntgb {
namespace sqp {

constexpr double kInf = 1e300;
constexpr double kBig = 1e19; /* |bound| >= kBig: no bound (the reference passes 1e20 for "infinite") */

/* the threads working on one problem */
struct Coop {
    int tid, nt; /* thread index and count */
    int lane, wl; /* index and size of the sub-group that runs the short sequential recurrences */
    NTG_HD void sync() const
    {
#ifdef __CUDA_ARCH__
        __syncthreads();
#endif
    }
    NTG_HD void syncw() const
    {
#ifdef __CUDA_ARCH__
        __syncwarp();
#endif
    }
    /* sum / maximum over the group, result in every thread; EVERY thread must call it.  red: [nt / 32]
     * doubles of shared scratch (only touched by groups of more than one warp).  Fixed tree: deterministic. */
    template <bool MAX>
    NTG_HD double reduce(double v, double *red) const
    {
#ifdef __CUDA_ARCH__
        for (int d = 16; d > 0; d >>= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, v, d);
            v = MAX ? fmax(v, o) : v + o;
        }
        if (nt > 32) {
            __syncthreads();
            if (lane == 0) red[tid >> 5] = v;
            __syncthreads();
            v = red[0];
            for (int k = 1; k < (nt >> 5); k++) v = MAX ? fmax(v, red[k]) : v + red[k];
        }
#else
        (void)red;
#endif
        return v;
    }
    /* (largest key, smallest index among equals) over the group; idx < 0 = no candidate */
    NTG_HD void argmax(double &key, int &idx, double *red, int *redi) const
    {
#ifdef __CUDA_ARCH__
        for (int d = 16; d > 0; d >>= 1) {
            const double ok = __shfl_xor_sync(0xffffffffu, key, d);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, d);
            if (oi >= 0 && (idx < 0 || ok > key || (ok == key && oi < idx))) {
                key = ok;
                idx = oi;
            }
        }
        if (nt > 32) {
            __syncthreads();
            if (lane == 0) {
                red[tid >> 5] = key;
                redi[tid >> 5] = idx;
            }
            __syncthreads();
            key = red[0];
            idx = redi[0];
            for (int k = 1; k < (nt >> 5); k++)
                if (redi[k] >= 0 && (idx < 0 || red[k] > key || (red[k] == key && redi[k] < idx))) {
                    key = red[k];
                    idx = redi[k];
                }
        }
#else
        (void)red;
        (void)redi;
#endif
    }
};

/* work space of one QP (all in shared memory on the GPU) */
struct Qp {
    int n, m, ld;     /* variables, rows, leading dimension of Jm / Rm / A (odd: conflict-free columns) */
    double *Jm, *Rm;  /* [n][ld] */
    double *A;        /* [m][ld] row i = gradient of row i */
    double *bl, *bu;  /* [m] bounds on A d (-1e20 / 1e20 = none) */
    double *rs;       /* [m] row scale max(1, |A_i|_inf) */
    double *x, *z, *dv, *rv, *npv, *cs, *sn; /* [n] */
    double *u;        /* [n + 1] multipliers of the active rows */
    double *lam;      /* [m] out: signed multipliers (> 0 at the lower bound, < 0 at the upper) */
    int *act;         /* [n + 1] +-(row + 1) of the active rows, sign = side */
    int *state;       /* [m] 0 inactive, 1 lower, 2 upper, 3 equality */
    double *red;      /* [nt] */
    int *redi;        /* [nt] */
    double *sh;       /* [4] broadcast */
    int *shi;         /* [4] */
};

/* NTG_SQP_CORE_BEGIN */

/* in-place Cholesky factor (lower) of the n x n matrix M (ld); returns 0, or 1 if a pivot is not positive */
template <class C>
NTG_HD int chol_lower(const C &cg, double *M, int n, int ld, int *shi)
{
    if (cg.tid == 0) shi[3] = 0;
    cg.sync();
    for (int j = 0; j < n; j++) {
        if (cg.tid == 0) {
            const double p = M[j * ld + j];
            if (!(p > 0.0) || !(p < kInf)) {
                shi[3] = 1;
                M[j * ld + j] = 1.0;
            } else
                M[j * ld + j] = sqrt(p);
        }
        cg.sync();
        const double djj = M[j * ld + j];
        for (int i = j + 1 + cg.tid; i < n; i += cg.nt) M[i * ld + j] /= djj;
        cg.sync();
        for (int i = j + 1 + cg.tid; i < n; i += cg.nt) {
            const double lij = M[i * ld + j];
            for (int k = j + 1; k <= i; k++) M[i * ld + k] -= lij * M[k * ld + j];
        }
        cg.sync();
    }
    return shi[3];
}

/* Jm = L^-T (upper triangular) from the Cholesky factor L (lower) */
template <class C>
NTG_HD void inv_transpose(const C &cg, const double *L, double *Jm, int n, int ld)
{
    for (int j = cg.tid; j < n; j += cg.nt) {
        for (int i = j + 1; i < n; i++) Jm[i * ld + j] = 0.0;
        Jm[j * ld + j] = 1.0 / L[j * ld + j];
        for (int i = j - 1; i >= 0; i--) {
            double acc = 0.0;
            for (int k = i + 1; k <= j; k++) acc += L[k * ld + i] * Jm[k * ld + j];
            Jm[i * ld + j] = -acc / L[i * ld + i];
        }
    }
    cg.sync();
}

/* x = -(Jm Jm') g */
template <class C>
NTG_HD void newton_point(const C &cg, const Qp &w, const double *g)
{
    const int n = w.n, ld = w.ld;
    for (int j = cg.tid; j < n; j += cg.nt) {
        double a = 0.0;
        for (int i = 0; i <= j; i++) a += w.Jm[i * ld + j] * g[i];
        w.dv[j] = a;
    }
    cg.sync();
    for (int i = cg.tid; i < n; i += cg.nt) {
        double a = 0.0;
        for (int j = i; j < n; j++) a += w.Jm[i * ld + j] * w.dv[j];
        w.x[i] = -a;
    }
    cg.sync();
}

/* remove the active row at position l (Goldfarb-Idnani step 2(c)); the multiplier of the row being
 * added, u[q], moves down with the others.  Returns the new q. */
template <class C>
NTG_HD int gi_drop(const C &cg, const Qp &w, int l, int q)
{
    const int n = w.n, ld = w.ld;
    if (cg.tid == 0) {
        const int a = w.act[l];
        w.state[(a > 0 ? a : -a) - 1] = 0;
        for (int k = l; k < q - 1; k++) {
            w.u[k] = w.u[k + 1];
            w.act[k] = w.act[k + 1];
        }
        w.u[q - 1] = w.u[q];
        w.u[q] = 0.0;
    }
    for (int i = cg.tid; i < n; i += cg.nt) {
        for (int k = l; k < q - 1; k++) w.Rm[i * ld + k] = w.Rm[i * ld + k + 1];
        w.Rm[i * ld + q - 1] = 0.0;
    }
    cg.sync();
    q -= 1;
    if (cg.tid < cg.wl) { /* R back to upper triangular: rotations of rows (j, j+1), sequential in j */
        for (int j = l; j < q; j++) {
            cg.syncw();
            const double a = w.Rm[j * ld + j], b = w.Rm[(j + 1) * ld + j];
            const double h = hypot(a, b);
            const double c = h == 0.0 ? 1.0 : a / h, s = h == 0.0 ? 0.0 : b / h;
            cg.syncw();
            for (int col = j + cg.lane; col < q; col += cg.wl) {
                const double p = w.Rm[j * ld + col], r = w.Rm[(j + 1) * ld + col];
                w.Rm[j * ld + col] = c * p + s * r;
                w.Rm[(j + 1) * ld + col] = col == j ? 0.0 : -s * p + c * r;
            }
            if (cg.lane == 0) {
                w.cs[j] = c;
                w.sn[j] = s;
            }
        }
    }
    cg.sync();
    for (int i = cg.tid; i < n; i += cg.nt)
        for (int j = l; j < q; j++) {
            const double p = w.Jm[i * ld + j], r = w.Jm[i * ld + j + 1];
            w.Jm[i * ld + j] = w.cs[j] * p + w.sn[j] * r;
            w.Jm[i * ld + j + 1] = -w.sn[j] * p + w.cs[j] * r;
        }
    cg.sync();
    return q;
}

/* min g0'x + 1/2 x'Gx  s.t.  bl <= A x <= bu, G = L L' given by its factor.  Returns 0 (solved),
 * 1 (the rows are inconsistent) or 2 (iteration limit); w.x, w.lam, w.state hold the result. */
template <class C>
NTG_HD int gi_solve(const C &cg, const Qp &w, const double *L, const double *g0)
{
    const int n = w.n, m = w.m, ld = w.ld;
    inv_transpose(cg, L, w.Jm, n, ld);
    for (int e = cg.tid; e < n * ld; e += cg.nt) w.Rm[e] = 0.0;
    for (int i = cg.tid; i < m; i += cg.nt) {
        w.state[i] = 0;
        w.lam[i] = 0.0;
        double a = 1.0;
        for (int k = 0; k < n; k++) a = fmax(a, fabs(w.A[i * ld + k]));
        w.rs[i] = a;
    }
    for (int k = cg.tid; k <= n; k += cg.nt) w.u[k] = 0.0;
    newton_point(cg, w, g0);
    int q = 0, status = 0;
    const int max_outer = 10 * (n + m) + 20, max_inner = 4 * (n + m) + 10;
    const double tol = 1e-10;
    for (int it = 0;; it++) {
        if (it >= max_outer) {
            status = 2;
            break;
        }
        /* the most violated inactive row; violated equalities first */
        double bkey = tol;
        int bi = -1;
        for (int i = cg.tid; i < m; i += cg.nt) {
            if (w.state[i] != 0) continue;
            double ax = 0.0;
            for (int k = 0; k < n; k++) ax += w.A[i * ld + k] * w.x[k];
            const double lo = w.bl[i], hi = w.bu[i];
            const double vlo = lo > -kBig ? lo - ax : -kInf, vhi = hi < kBig ? ax - hi : -kInf;
            double v = fmax(vlo, vhi) / w.rs[i];
            if (lo == hi && v > tol) v += 1e100;
            if (v > bkey) {
                bkey = v;
                bi = i;
            }
        }
        cg.argmax(bkey, bi, w.red, w.redi);
        if (cg.tid == 0) {
            const int i0 = bi;
            w.shi[0] = i0;
            if (i0 >= 0) {
                double ax = 0.0;
                for (int k = 0; k < n; k++) ax += w.A[i0 * ld + k] * w.x[k];
                const double lo = w.bl[i0], hi = w.bu[i0];
                const double vlo = lo > -kBig ? lo - ax : -kInf, vhi = hi < kBig ? ax - hi : -kInf;
                const double sg = vlo >= vhi ? 1.0 : -1.0;
                w.sh[1] = sg;
                w.sh[2] = sg > 0 ? lo : -hi;       /* b: the row reads sg*a'x >= b */
                w.sh[3] = sg * ax - w.sh[2];       /* s < 0 */
                w.u[q] = 0.0;
            }
        }
        cg.sync();
        const int ip = w.shi[0];
        if (ip < 0) break;
        const double sg = w.sh[1];
        for (int k = cg.tid; k < n; k += cg.nt) w.npv[k] = sg * w.A[ip * ld + k];
        cg.sync();
        for (int inner = 0;; inner++) {
            if (inner >= max_inner) {
                status = 2;
                break;
            }
            for (int j = cg.tid; j < n; j += cg.nt) {
                double a = 0.0;
                for (int i = 0; i < n; i++) a += w.Jm[i * ld + j] * w.npv[i];
                w.dv[j] = a;
            }
            cg.sync();
            for (int i = cg.tid; i < n; i += cg.nt) {
                double a = 0.0;
                for (int j = q; j < n; j++) a += w.Jm[i * ld + j] * w.dv[j];
                w.z[i] = a;
                if (i < q) w.rv[i] = w.dv[i];
            }
            cg.sync();
            if (cg.tid < cg.wl) { /* r = R^-1 d[0:q], column-oriented back substitution */
                for (int k = q - 1; k >= 0; k--) {
                    cg.syncw();
                    const double rk = w.rv[k] / w.Rm[k * ld + k];
                    cg.syncw();
                    if (cg.lane == 0) w.rv[k] = rk;
                    for (int i = cg.lane; i < k; i += cg.wl) w.rv[i] -= w.Rm[i * ld + k] * rk;
                }
            }
            cg.sync();
            if (cg.tid == 0) {
                double zn = 0.0, nn = 0.0;
                for (int i = 0; i < n; i++) {
                    zn += w.z[i] * w.npv[i];
                    nn += w.npv[i] * w.npv[i];
                }
                const double s = w.sh[3];
                const double t2 = zn > 1e-13 * fmax(1.0, nn) ? -s / zn : kInf;
                double t1 = kInf;
                int l = -1;
                for (int k = 0; k < q; k++) {
                    const int a = w.act[k], row = (a > 0 ? a : -a) - 1;
                    if (w.bl[row] != w.bu[row] && w.rv[k] > 0.0) {
                        const double tk = w.u[k] / w.rv[k];
                        if (tk < t1) {
                            t1 = tk;
                            l = k;
                        }
                    }
                }
                const double t = fmin(t1, t2);
                int code;
                if (t >= kInf) code = 1;
                else if (t2 >= kInf) code = 2;
                else if (t2 <= t1) code = 3;
                else code = 4;
                w.sh[0] = t;
                w.shi[1] = l;
                w.shi[2] = code;
            }
            cg.sync();
            const int code = w.shi[2], l = w.shi[1];
            const double t = w.sh[0];
            if (code == 1) {
                status = 1;
                break;
            }
            if (code != 2)
                for (int i = cg.tid; i < n; i += cg.nt) w.x[i] += t * w.z[i];
            for (int k = cg.tid; k < q; k += cg.nt) w.u[k] -= t * w.rv[k];
            if (cg.tid == 0) w.u[q] += t;
            cg.sync();
            if (code == 3) { /* the full step: the row joins the active set */
                /* Givens rotations that fold d[q+1:] into d[q], back to front: the value carried into the
                 * rotation of (j-1, j) is the norm of d[j:], so every pair (c, s) follows from the suffix sums
                 * of squares -- one short chain of additions, then one sqrt and two divisions PER THREAD
                 * instead of a chain of hypot() calls on one thread */
                if (cg.tid == 0) {
                    double ss = 0.0;
                    for (int j = n - 1; j >= q; j--) {
                        ss += w.dv[j] * w.dv[j];
                        w.z[j] = ss; /* z is free again: x has been updated */
                    }
                }
                cg.sync();
                for (int j = q + 1 + cg.tid; j < n; j += cg.nt) {
                    const double h = sqrt(w.z[j - 1]);
                    w.cs[j] = h == 0.0 ? 1.0 : w.dv[j - 1] / h;
                    w.sn[j] = h == 0.0 ? 0.0 : (j == n - 1 ? w.dv[j] : sqrt(w.z[j])) / h; /* the last entry keeps its sign */
                }
                if (cg.tid == 0) {
                    for (int k = 0; k < q; k++) w.Rm[k * ld + q] = w.dv[k];
                    w.Rm[q * ld + q] = n - 1 > q ? sqrt(w.z[q]) : w.dv[q];
                    w.act[q] = sg > 0 ? ip + 1 : -(ip + 1);
                    w.state[ip] = w.bl[ip] == w.bu[ip] ? 3 : (sg > 0 ? 1 : 2);
                }
                cg.sync();
                for (int i = cg.tid; i < n; i += cg.nt)
                    for (int j = n - 1; j > q; j--) {
                        const double p = w.Jm[i * ld + j - 1], r = w.Jm[i * ld + j];
                        w.Jm[i * ld + j - 1] = w.cs[j] * p + w.sn[j] * r;
                        w.Jm[i * ld + j] = -w.sn[j] * p + w.cs[j] * r;
                    }
                q += 1;
                cg.sync();
                break;
            }
            q = gi_drop(cg, w, l, q);
            if (code == 4) {
                if (cg.tid == 0) {
                    double a = 0.0;
                    for (int k = 0; k < n; k++) a += w.npv[k] * w.x[k];
                    w.sh[3] = a - w.sh[2];
                }
                cg.sync();
            }
        }
        if (status) break;
    }
    cg.sync();
    if (cg.tid == 0)
        for (int k = 0; k < q; k++) {
            const int a = w.act[k];
            if (a > 0) w.lam[a - 1] = w.u[k];
            else w.lam[-a - 1] = -w.u[k];
        }
    cg.sync();
    return status;
}

/* per-problem solver state that lives in global memory between iterations */
struct StepState {
    double *y;      /* [nr] reduced variables */
    double *B;      /* [nr][nr] quasi-Newton Hessian of the reduced Lagrangian */
    double *lam;    /* [m] multipliers of the last QP */
    double *sprev;  /* [nr] the last accepted step alpha*d */
    double *grLold; /* [nr] gr - Ar'lam at the previous point, with the multipliers of ITS QP */
    double *grold;  /* [nr] gr at the previous point */
    double *d;      /* [nr] out: search direction */
    double *scal;   /* [0] nu  [1] phi0 (out)  [2] dphi0 (out)  [3] violation, scaled max (out)  [4] kkt (out) */
    int *flag;      /* [0] a previous step exists  [1] that step was a Gauss-Newton (restoration) step
                       [2] reset B requested  [3] B is the identity (no update since the last reset)
                       [4] out: this direction is a restoration step  [5] out: status (0 go on, 1 converged,
                       4 stationary point of the violation) [6] out: rows active in the QP */
    int *istate;    /* [m] out */
};

struct StepOpts {
    double gtol, ctol, rho_pen;
};

/* One SQP iteration of one problem up to the search direction.  In: f, gr (reduced gradient), h (row
 * values), hl / hu, w.A (reduced row gradients).  smem: Bm, Lm [nr][ld]; vec [8*nr + m] scratch. */
template <class C>
NTG_HD void sqp_step(const C &cg, const Qp &w, const StepState &S, const StepOpts &o, double f, const double *gr,
                     const double *h, const double *hl, const double *hu, double *Bm, double *Lm, double *vec)
{
    const int n = w.n, m = w.m, ld = w.ld;
    double *grL = vec, *Bs = vec + n, *uu = vec + 2 * n, *g2 = vec + 3 * n, *Ad = vec + 4 * n; /* Ad [m] */
    for (int i = cg.tid; i < m; i += cg.nt) {
        w.bl[i] = hl[i] > -kBig ? hl[i] - h[i] : -1e20;
        w.bu[i] = hu[i] < kBig ? hu[i] - h[i] : 1e20;
    }
    for (int e = cg.tid; e < n * n; e += cg.nt) {
        const int i = e / n, j = e - i * n;
        Bm[i * ld + j] = S.flag[2] ? (i == j ? 1.0 : 0.0) : S.B[e];
    }
    cg.sync();
    const bool reset = S.flag[2] != 0;
    bool fresh = reset || S.flag[3] != 0;
    /* ---- damped BFGS update with the step that led here ---- */
    if (!reset && S.flag[0] && S.flag[1]) {
        /* it was a Gauss-Newton step, whose multipliers say nothing about the Lagrangian: B only takes
         * the SCALE of the cost's curvature along the step (never down), so that an identity that is
         * orders of magnitude too small does not send the next steps far outside the model */
        for (int k = cg.tid; k < n; k += cg.nt) {
            double b = 0.0;
            for (int j = 0; j < n; j++) b += Bm[k * ld + j] * S.sprev[j];
            Bs[k] = b;
        }
        cg.sync();
        if (cg.tid == 0) {
            double sBs = 0.0, suf = 0.0;
            for (int k = 0; k < n; k++) {
                sBs += S.sprev[k] * Bs[k];
                suf += S.sprev[k] * (gr[k] - S.grold[k]);
            }
            w.sh[0] = (sBs > 0.0 && suf > sBs) ? fmin(suf / sBs, 1e8) : 1.0;
        }
        cg.sync();
        const double fac = w.sh[0];
        if (fac != 1.0) {
            for (int e = cg.tid; e < n * n; e += cg.nt) {
                const int i = e / n, j = e - i * n;
                Bm[i * ld + j] *= fac;
            }
            fresh = false;
        }
        cg.sync();
    }
    if (!reset && S.flag[0] && !S.flag[1]) {
        for (int k = cg.tid; k < n; k += cg.nt) {
            double a = gr[k];
            for (int i = 0; i < m; i++) a -= w.A[i * ld + k] * S.lam[i];
            uu[k] = a - S.grLold[k];
            double b = 0.0;
            for (int j = 0; j < n; j++) b += Bm[k * ld + j] * S.sprev[j];
            Bs[k] = b;
        }
        cg.sync();
        if (cg.tid == 0) {
            double sBs = 0.0, su = 0.0;
            for (int k = 0; k < n; k++) {
                sBs += S.sprev[k] * Bs[k];
                su += S.sprev[k] * uu[k];
            }
            double th = 1.0;
            if (su < 0.2 * sBs) {
                th = 0.8 * sBs / (sBs - su);
                su = th * su + (1.0 - th) * sBs;
            }
            w.sh[0] = th;
            w.sh[1] = sBs;
            w.sh[2] = su;
        }
        cg.sync();
        const double th = w.sh[0], sBs = w.sh[1], su = w.sh[2];
        if (sBs > 0.0 && su > 0.0) {
            for (int e = cg.tid; e < n * n; e += cg.nt) {
                const int i = e / n, j = e - i * n;
                const double ui = th * uu[i] + (1.0 - th) * Bs[i], uj = th * uu[j] + (1.0 - th) * Bs[j];
                Bm[i * ld + j] += -Bs[i] * Bs[j] / sBs + ui * uj / su;
            }
            fresh = false;
        }
        cg.sync();
    }
    /* ---- factor (B = I if it lost definiteness to rounding) ---- */
    for (int e = cg.tid; e < n * n; e += cg.nt) {
        const int i = e / n, j = e - i * n;
        Lm[i * ld + j] = Bm[i * ld + j];
    }
    cg.sync();
    if (chol_lower(cg, Lm, n, ld, w.shi)) {
        cg.sync();
        for (int e = cg.tid; e < n * n; e += cg.nt) {
            const int i = e / n, j = e - i * n;
            Bm[i * ld + j] = Lm[i * ld + j] = i == j ? 1.0 : 0.0;
        }
        fresh = true;
        cg.sync();
    }
    for (int e = cg.tid; e < n * n; e += cg.nt) {
        const int i = e / n, j = e - i * n;
        S.B[e] = Bm[i * ld + j];
    }
    /* ---- QP ---- */
    int st = gi_solve(cg, w, Lm, gr);
    bool restor = false;
    if (st != 0) {
        restor = true;
        /* rows violated at this point: Gauss-Newton term instead of a constraint */
        if (cg.tid == 0) {
            double amax = 0.0, tr = 0.0;
            for (int i = 0; i < m; i++) {
                const bool v = (w.bl[i] > -kBig && w.bl[i] > 0.0) || (w.bu[i] < kBig && w.bu[i] < 0.0);
                if (!v) continue;
                double a = 0.0;
                for (int k = 0; k < n; k++) a += w.A[i * ld + k] * w.A[i * ld + k];
                amax = fmax(amax, a);
            }
            for (int k = 0; k < n; k++) tr += Bm[k * ld + k];
            w.sh[0] = o.rho_pen * fmax(1.0, tr / n) / fmax(1e-300, amax);
        }
        cg.sync();
        const double rho = w.sh[0];
        for (int e = cg.tid; e < n * n; e += cg.nt) {
            const int i = e / n, j = e - i * n;
            double a = 0.0;
            for (int r = 0; r < m; r++) {
                const bool v = (w.bl[r] > -kBig && w.bl[r] > 0.0) || (w.bu[r] < kBig && w.bu[r] < 0.0);
                if (v) a += w.A[r * ld + i] * w.A[r * ld + j];
            }
            Lm[i * ld + j] = Bm[i * ld + j] + rho * a;
        }
        for (int k = cg.tid; k < n; k += cg.nt) {
            double a = 0.0;
            for (int r = 0; r < m; r++) {
                const bool vl = w.bl[r] > -kBig && w.bl[r] > 0.0, vu = w.bu[r] < kBig && w.bu[r] < 0.0;
                if (vl) a += w.A[r * ld + k] * w.bl[r];
                else if (vu) a += w.A[r * ld + k] * w.bu[r];
            }
            g2[k] = gr[k] - rho * a;
        }
        cg.sync();
        chol_lower(cg, Lm, n, ld, w.shi);
        inv_transpose(cg, Lm, w.Jm, n, ld);
        newton_point(cg, w, g2);
        for (int r = cg.tid; r < m; r += cg.nt) {
            const bool vl = w.bl[r] > -kBig && w.bl[r] > 0.0, vu = w.bu[r] < kBig && w.bu[r] < 0.0;
            double ax = 0.0;
            for (int k = 0; k < n; k++) ax += w.A[r * ld + k] * w.x[k];
            w.lam[r] = vl ? -rho * (ax - w.bl[r]) : (vu ? -rho * (ax - w.bu[r]) : 0.0);
            w.state[r] = 0;
        }
        cg.sync();
    }
    /* ---- Lagrangian gradient, violation, merit and its slope along d ---- */
    for (int k = cg.tid; k < n; k += cg.nt) {
        double a = gr[k];
        for (int i = 0; i < m; i++) a -= w.A[i * ld + k] * w.lam[i];
        grL[k] = a;
        double b = 0.0;
        for (int j = 0; j < n; j++) b += Bm[k * ld + j] * w.x[j];
        Bs[k] = b; /* B d */
    }
    for (int i = cg.tid; i < m; i += cg.nt) {
        double a = 0.0;
        for (int k = 0; k < n; k++) a += w.A[i * ld + k] * w.x[k];
        Ad[i] = a;
    }
    cg.sync();
    double kkt = 0.0, dmax = 0.0, gd = 0.0, dBd = 0.0;
    for (int k = cg.tid; k < n; k += cg.nt) {
        kkt = fmax(kkt, fabs(grL[k]));
        dmax = fmax(dmax, fabs(w.x[k]));
        gd += gr[k] * w.x[k];
        dBd += w.x[k] * Bs[k];
    }
    double vmax = 0.0, vsum = 0.0, almax = 0.0, dsum = 0.0, nactd = 0.0;
    for (int i = cg.tid; i < m; i += cg.nt) {
        const bool vl = w.bl[i] > -kBig && w.bl[i] > 0.0, vu = w.bu[i] < kBig && w.bu[i] < 0.0;
        const double v = vl ? w.bl[i] : (vu ? -w.bu[i] : 0.0);
        const double sc = fmax(1.0, fmax(fabs(hl[i]) < kBig ? fabs(hl[i]) : 0.0, fabs(hu[i]) < kBig ? fabs(hu[i]) : 0.0));
        vmax = fmax(vmax, v / sc);
        vsum += v;
        almax = fmax(almax, fabs(w.lam[i]));
        if (vl) dsum -= Ad[i];
        else if (vu) dsum += Ad[i];
        nactd += w.state[i] != 0 ? 1.0 : 0.0;
    }
    kkt = cg.template reduce<true>(kkt, w.red);
    dmax = cg.template reduce<true>(dmax, w.red);
    gd = cg.template reduce<false>(gd, w.red);
    dBd = cg.template reduce<false>(dBd, w.red);
    vmax = cg.template reduce<true>(vmax, w.red);
    vsum = cg.template reduce<false>(vsum, w.red);
    almax = cg.template reduce<true>(almax, w.red);
    dsum = cg.template reduce<false>(dsum, w.red);
    nactd = cg.template reduce<false>(nactd, w.red);
    if (cg.tid == 0) {
        const int nact = (int)nactd;
        int status = 0;
        if (vmax <= o.ctol && kkt <= o.gtol * fmax(1.0, fabs(f)) && !restor) status = 1;
        else if (restor && dmax < 1e-12) status = 4;
        const double nu0 = reset ? 0.0 : S.scal[0];
        const double nu = fmax(almax, 0.5 * (nu0 + almax));
        double D = gd + nu * dsum;
        if (D > -1e-14) D = -fabs(dBd);
        S.scal[0] = nu;
        S.scal[1] = f + nu * vsum;
        S.scal[2] = D;
        S.scal[3] = vmax;
        S.scal[4] = kkt;
        S.flag[2] = 0;
        S.flag[3] = fresh ? 1 : 0;
        S.flag[4] = restor ? 1 : 0;
        S.flag[5] = status;
        S.flag[6] = nact;
    }
    for (int k = cg.tid; k < n; k += cg.nt) {
        S.d[k] = w.x[k];
        S.grLold[k] = grL[k];
        S.grold[k] = gr[k];
    }
    for (int i = cg.tid; i < m; i += cg.nt) {
        S.lam[i] = w.lam[i];
        S.istate[i] = w.state[i];
    }
    cg.sync();
}

/* NTG_SQP_CORE_END */

/* doubles / ints of shared memory one problem needs (same carve-up on the host test side) */
NTG_HD inline int sqp_ld(int n) { return n | 1; }
NTG_HD inline size_t sqp_smem_doubles(int n, int m, int nt)
{
    const size_t ld = (size_t)sqp_ld(n);
    return 4 * (size_t)n * ld + (size_t)m * ld + 5 * (size_t)m + 9 * (size_t)n + 2 + 4 * (size_t)n + (size_t)m + nt + 4;
}
NTG_HD inline size_t sqp_smem_ints(int n, int m, int nt) { return (size_t)n + 1 + (size_t)m + nt + 4; }

/* carve the work space out of one block of doubles followed by one block of ints */
NTG_HD inline void sqp_carve(double *dbl, int *ints, int n, int m, int nt, Qp &w, double *&Bm, double *&Lm, double *&vec,
                             double *&gr, double *&hrow)
{
    const int ld = sqp_ld(n);
    w.n = n;
    w.m = m;
    w.ld = ld;
    double *q = dbl;
    w.Jm = q; q += (size_t)n * ld;
    w.Rm = q; q += (size_t)n * ld;
    Bm = q;   q += (size_t)n * ld;
    Lm = q;   q += (size_t)n * ld;
    w.A = q;  q += (size_t)m * ld;
    w.bl = q; q += m;
    w.bu = q; q += m;
    w.rs = q; q += m;
    w.lam = q; q += m;
    hrow = q; q += m;
    w.x = q; q += n;
    w.z = q; q += n;
    w.dv = q; q += n;
    w.rv = q; q += n;
    w.npv = q; q += n;
    w.cs = q; q += n;
    w.sn = q; q += n;
    gr = q; q += n;
    w.u = q; q += n + 2;
    vec = q; q += 4 * (size_t)n + m;
    w.red = q; q += nt;
    w.sh = q; q += 4;
    int *r = ints;
    w.act = r; r += n + 1;
    w.state = r; r += m;
    w.redi = r; r += nt;
    w.shi = r;
}

} /* namespace sqp */
} /* namespace ntgb */
#endif
