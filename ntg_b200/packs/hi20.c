/*
 * hi20.c -- test-only pack for the upper limits of the spline machinery: one
 * flat output, up to order 20 (PGS bsplvb's jmax),
 * three derivatives.  Signatures: reference src/ntg.c:34-41.
 */
void hi20_ucf(int *mode, int *nstate, int *i, double *f, double *df, double **zp)
{
    (void)nstate; (void)i;
    if (*mode == 0 || *mode == 2)
        *f = zp[0][0] * zp[0][0] + 0.5 * zp[0][1] * zp[0][1] + 0.25 * zp[0][2] * zp[0][2];
    if (*mode == 1 || *mode == 2) {
        df[0] = 2.0 * zp[0][0];
        df[1] = zp[0][1];
        df[2] = 0.5 * zp[0][2];
    }
}

void hi20_nltcf(int *mode, int *nstate, int *i, double *f, double **df, double **zp)
{
    (void)nstate; (void)i;
    if (*mode == 0 || *mode == 2)
        f[0] = zp[0][0] * zp[0][2] - zp[0][1];
    if (*mode == 1 || *mode == 2) {
        df[0][0] = zp[0][2];
        df[0][1] = -1.0;
        df[0][2] = zp[0][0];
    }
}
