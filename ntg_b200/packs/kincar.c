/*
 * kincar.c -- callback pack for the kinematic car (CFG-3/CFG-4).
 *
 * Plain C with NTG's callback signatures (reference src/ntg.c:34-41); compiled
 * by gcc for the CPU oracle and by nvcc as __device__ code for the evaluator.
 *
 * Flat outputs: zp[0] = x, zp[1] = y (rear-axle position), 3 derivatives each
 * (reference examples/kincar.c:133-137).  Cost = integrated squared
 * acceleration, the same integrand as the reference's tcf
 * (examples/kincar.c:105-117).  KC-C is the bench's constraint pack
 * (SURVEY.md section 8): speed^2 and the curvature numerator, polynomial forms
 * of the quantities in kincar_flat_reverse (examples/kincar.c:89-91) that the
 * reference's own TODO asks for (examples/kincar.c:314-315).
 */
#define XD  zp[0][1]
#define XDD zp[0][2]
#define YD  zp[1][1]
#define YDD zp[1][2]

void kc_ucf(int *mode, int *nstate, int *i, double *f, double *df, double **zp)
{
    (void)nstate; (void)i;
    if (*mode == 0 || *mode == 2)
        *f = XDD * XDD + YDD * YDD;
    if (*mode == 1 || *mode == 2) {
        df[0] = 0.0; df[1] = 0.0; df[2] = 2.0 * XDD;
        df[3] = 0.0; df[4] = 0.0; df[5] = 2.0 * YDD;
    }
}

/* KC-C: c0 = xd^2 + yd^2 (speed^2), c1 = xd*ydd - yd*xdd (curvature * speed^3). */
void kc_nltcf(int *mode, int *nstate, int *i, double *f, double **df, double **zp)
{
    (void)nstate; (void)i;
    if (*mode == 0 || *mode == 2) {
        f[0] = XD * XD + YD * YD;
        f[1] = XD * YDD - YD * XDD;
    }
    if (*mode == 1 || *mode == 2) {
        df[0][1] = 2.0 * XD;  df[0][4] = 2.0 * YD;
        df[1][1] = YDD;       df[1][2] = -YD;
        df[1][4] = -XDD;      df[1][5] = XD;
    }
}
