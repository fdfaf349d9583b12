/*
 * vdp.c -- callback pack for the driven van der Pol oscillator (CFG-1/CFG-2).
 *
 * Plain C with NTG's callback signatures (reference src/ntg.c:34-41), so the
 * same file is compiled by gcc for the CPU oracle and by nvcc as __device__
 * code for the B200 evaluator.
 *
 * Problem statement: reference examples/vanderpol.txt:54-58 --
 *   flat output z, input u = zdd + z - (1 - z^2) zd,
 *   cost integrand 1/2 (z^2 + zd^2 + u^2).
 * The cost below is our own statement of that integrand (the reference's
 * examples/vanderpol.c:206-241 is Maple output and is exercised unmodified
 * through the drop-in pack instead).  VDP-C is the bench's constraint pack
 * ("Configs made concrete"): c0 = u.
 */
#define Z   zp[0][0]
#define ZD  zp[0][1]
#define ZDD zp[0][2]

void vdp_ucf(int *mode, int *nstate, int *i, double *f, double *df, double **zp)
{
    double w = 1.0 - Z * Z;
    double u = ZDD + Z - w * ZD;
    (void)nstate; (void)i;
    if (*mode == 0 || *mode == 2)
        *f = 0.5 * (Z * Z) + 0.5 * (ZD * ZD) + 0.5 * (u * u);
    if (*mode == 1 || *mode == 2) {
        df[0] = Z + u * (1.0 + 2.0 * Z * ZD);
        df[1] = ZD - u * w;
        df[2] = u;
    }
}

/* VDP-C: one nonlinear trajectory constraint, the input u itself. */
void vdp_nltcf(int *mode, int *nstate, int *i, double *f, double **df, double **zp)
{
    double w = 1.0 - Z * Z;
    (void)nstate; (void)i;
    if (*mode == 0 || *mode == 2)
        f[0] = ZDD + Z - w * ZD;
    if (*mode == 1 || *mode == 2) {
        df[0][0] = 1.0 + 2.0 * Z * ZD;
        df[0][1] = -w;
        df[0][2] = 1.0;
    }
}
