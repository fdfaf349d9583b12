/*
 * endpt.c -- test-only pack that exercises every callback kind NTG has
 * (no shipped example uses initial/final cost or nonlinear initial/final
 * constraints; SURVEY.md section 8: "A test-only pack must also exercise
 * initial/final cost and initial/final constraints").
 *
 * Two flat outputs with different derivative depth: zp[0][0..2], zp[1][0..1]
 * (nz = 5).  Uses libm calls (sin, cos, exp, sqrt) so the transcendental path
 * of device callbacks is covered.  Signatures: reference src/ntg.c:34-41.
 */
#include <math.h>

#define P0 zp[0][0]
#define P1 zp[0][1]
#define P2 zp[0][2]
#define Q0 zp[1][0]
#define Q1 zp[1][1]

void ep_icf(int *mode, int *nstate, double *f, double *df, double **zp)
{
    (void)nstate;
    if (*mode == 0 || *mode == 2)
        *f = (P0 - 1.0) * (P0 - 1.0) + 0.5 * P1 * Q0 + exp(0.25 * Q1);
    if (*mode == 1 || *mode == 2) {
        df[0] = 2.0 * (P0 - 1.0);
        df[1] = 0.5 * Q0;
        df[2] = 0.0;
        df[3] = 0.5 * P1;
        df[4] = 0.25 * exp(0.25 * Q1);
    }
}

void ep_ucf(int *mode, int *nstate, int *i, double *f, double *df, double **zp)
{
    double s = sin(P0), c = cos(P0);
    (void)nstate;
    /* the breakpoint index enters the integrand (weight grows along the horizon) */
    double w = 1.0 + 0.125 * (double)(*i);
    if (*mode == 0 || *mode == 2)
        *f = w * (P2 * P2 + Q1 * Q1) + s * Q0 + 0.1 * P1 * P1;
    if (*mode == 1 || *mode == 2) {
        df[0] = c * Q0;
        df[1] = 0.2 * P1;
        df[2] = 2.0 * w * P2;
        df[3] = s;
        df[4] = 2.0 * w * Q1;
    }
    /* a callback may ask the solver to stop by setting *mode = -1 (reference src/ntg.c:369) */
    if (P0 > 1.0e6)
        *mode = -1;
}

void ep_fcf(int *mode, int *nstate, double *f, double *df, double **zp)
{
    (void)nstate;
    if (*mode == 0 || *mode == 2)
        *f = sqrt(1.0 + P0 * P0 + Q0 * Q0) + P2 * Q1;
    if (*mode == 1 || *mode == 2) {
        double r = sqrt(1.0 + P0 * P0 + Q0 * Q0);
        df[0] = P0 / r;
        df[1] = 0.0;
        df[2] = Q1;
        df[3] = Q0 / r;
        df[4] = P2;
    }
}

/* 2 nonlinear initial constraints */
void ep_nlicf(int *mode, int *nstate, double *f, double **df, double **zp)
{
    (void)nstate;
    if (*mode == 0 || *mode == 2) {
        f[0] = P0 * P0 + Q0 * Q0;
        f[1] = P1 * Q1 - P2;
    }
    if (*mode == 1 || *mode == 2) {
        df[0][0] = 2.0 * P0; df[0][3] = 2.0 * Q0;
        df[1][1] = Q1; df[1][2] = -1.0; df[1][4] = P1;
    }
}

/* 3 nonlinear trajectory constraints; the third one depends on the index */
void ep_nltcf(int *mode, int *nstate, int *i, double *f, double **df, double **zp)
{
    double e = exp(-0.5 * Q0);
    (void)nstate;
    if (*mode == 0 || *mode == 2) {
        f[0] = P1 * P1 + Q1 * Q1;
        f[1] = P0 * Q1 - Q0 * P1 + e;
        f[2] = P2 + 0.01 * (double)(*i) * Q0;
    }
    if (*mode == 1 || *mode == 2) {
        df[0][1] = 2.0 * P1; df[0][4] = 2.0 * Q1;
        df[1][0] = Q1; df[1][1] = -Q0; df[1][3] = -P1 - 0.5 * e; df[1][4] = P0;
        df[2][2] = 1.0; df[2][3] = 0.01 * (double)(*i);
    }
}

/* 1 nonlinear final constraint */
void ep_nlfcf(int *mode, int *nstate, double *f, double **df, double **zp)
{
    (void)nstate;
    if (*mode == 0 || *mode == 2)
        f[0] = sin(P0) + cos(Q0) + P1 * Q1;
    if (*mode == 1 || *mode == 2) {
        df[0][0] = cos(P0); df[0][1] = Q1; df[0][3] = -sin(Q0); df[0][4] = P1;
    }
}
