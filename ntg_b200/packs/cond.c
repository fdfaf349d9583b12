/*
 * cond.c -- test-only pack whose derivative SPARSITY depends on the data: some entries are
 * written only when |z| is large.  The build-time sparsity probe (ntg_b200/build.py::
 * probe_sparsity) samples z in [-3, 3] and therefore declares those entries structurally zero;
 * the kernels must notice at run time that they are not (sp_clean) and take the dense chain rule.
 * (The entries are written on every call, 0.0 when inactive: the reference does not clear its
 * derivative buffers between breakpoints -- src/cost.c:100, src/constraints.c:146-155 -- so a
 * callback that skips a write leaves the previous breakpoint's value there.)
 * Two outputs, order 4, three derivatives each.  Signatures: reference src/ntg.c:34-41.
 */
#define X0 zp[0][0]
#define X1 zp[0][1]
#define Y0 zp[1][0]
#define Y2 zp[1][2]

void cond_ucf(int *mode, int *nstate, int *i, double *f, double *df, double **zp)
{
    (void)nstate; (void)i;
    if (*mode == 0 || *mode == 2)
        *f = X1 * X1 + Y2 * Y2 + (X0 > 50.0 ? 0.001 * X0 * X0 : 0.0);
    if (*mode == 1 || *mode == 2) {
        df[1] = 2.0 * X1;
        df[5] = 2.0 * Y2;
        df[0] = X0 > 50.0 ? 0.002 * X0 : 0.0;    /* the probe only ever sees the 0.0 */
    }
}

void cond_nltcf(int *mode, int *nstate, int *i, double *f, double **df, double **zp)
{
    (void)nstate; (void)i;
    if (*mode == 0 || *mode == 2) {
        f[0] = X1 + (Y0 < -50.0 ? Y0 * Y0 : 0.0);
        f[1] = X0 * Y2;
    }
    if (*mode == 1 || *mode == 2) {
        df[0][1] = 1.0;
        df[0][3] = Y0 < -50.0 ? 2.0 * Y0 : 0.0;  /* the probe only ever sees the 0.0 */
        df[1][0] = Y2;
        df[1][5] = X0;
    }
}
