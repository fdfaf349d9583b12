/*
 * syn6.c -- synthetic high-order pack (CFG-5): 6 flat outputs, 4 derivatives
 * each (nz = 24), one trajectory cost and four polynomial trajectory
 * constraints that couple all 24 active variables (SURVEY.md section 8,
 * "SYN6 four polynomial couplings over all 24 z's").
 *
 * Plain C with NTG's callback signatures (reference src/ntg.c:34-41).
 */
#define SYN6_NOUT 6
#define SYN6_MAXD 4

/* cost: sum_j sum_d w_d * z[j][d]^2  +  z[j][0]*z[(j+1)%6][1] */
void syn6_ucf(int *mode, int *nstate, int *i, double *f, double *df, double **zp)
{
    int j, d;
    (void)nstate; (void)i;
    if (*mode == 0 || *mode == 2) {
        double s = 0.0;
        for (j = 0; j < SYN6_NOUT; j++) {
            for (d = 0; d < SYN6_MAXD; d++)
                s += (0.5 / (double)(1 + d)) * zp[j][d] * zp[j][d];
            s += zp[j][0] * zp[(j + 1) % SYN6_NOUT][1];
        }
        *f = s;
    }
    if (*mode == 1 || *mode == 2) {
        for (j = 0; j < SYN6_NOUT; j++) {
            int jn = (j + 1) % SYN6_NOUT, jp = (j + SYN6_NOUT - 1) % SYN6_NOUT;
            for (d = 0; d < SYN6_MAXD; d++)
                df[j * SYN6_MAXD + d] = (1.0 / (double)(1 + d)) * zp[j][d];
            df[j * SYN6_MAXD + 0] += zp[jn][1];
            df[j * SYN6_MAXD + 1] += zp[jp][0];
        }
    }
}

/*
 * constraints, m = 0..3:
 *   c_m = sum_j ( a_mj * z[j][m] * z[(j+m+1)%6][(m+1)%4] + b_mj * z[j][(m+2)%4] )
 * with a_mj = 1 + 0.25*((m+j)%3), b_mj = 0.5 - 0.125*((m+2*j)%5).
 * Every constraint touches all six outputs; together they touch all 24 z's.
 */
void syn6_nltcf(int *mode, int *nstate, int *i, double *f, double **df, double **zp)
{
    int m, j;
    (void)nstate; (void)i;
    for (m = 0; m < 4; m++) {
        int d0 = m, d1 = (m + 1) % SYN6_MAXD, d2 = (m + 2) % SYN6_MAXD;
        if (*mode == 0 || *mode == 2) {
            double s = 0.0;
            for (j = 0; j < SYN6_NOUT; j++) {
                int jn = (j + m + 1) % SYN6_NOUT;
                double a = 1.0 + 0.25 * (double)((m + j) % 3);
                double b = 0.5 - 0.125 * (double)((m + 2 * j) % 5);
                s += a * zp[j][d0] * zp[jn][d1] + b * zp[j][d2];
            }
            f[m] = s;
        }
        if (*mode == 1 || *mode == 2) {
            for (j = 0; j < SYN6_NOUT * SYN6_MAXD; j++)
                df[m][j] = 0.0;
            for (j = 0; j < SYN6_NOUT; j++) {
                int jn = (j + m + 1) % SYN6_NOUT;
                double a = 1.0 + 0.25 * (double)((m + j) % 3);
                double b = 0.5 - 0.125 * (double)((m + 2 * j) % 5);
                df[m][j * SYN6_MAXD + d0] += a * zp[jn][d1];
                df[m][jn * SYN6_MAXD + d1] += a * zp[j][d0];
                df[m][j * SYN6_MAXD + d2] += b;
            }
        }
    }
}
