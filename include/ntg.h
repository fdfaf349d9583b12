/*
 * ntg.h -- drop-in replacement for the reference's public header
 * (reference src/ntg.h, which pulls in av.h, colloc.h, constraints.h, cost.h,
 * matrix.h).  A program written against NTG -- examples/vanderpol.c and
 * examples/kincar.c in the reference tree -- compiles unchanged against this
 * header and links against libntg_b200.so.
 *
 * Only the surface user programs touch is declared:
 *   ntg(), npsoloption(), linspace(), printNTGBanner()      src/ntg.h:72-104
 *   AV                                                     src/av.h:18-26
 *   Matrix, MakeMatrix, FreeMatrix, DoubleMatrix,
 *   FreeDoubleMatrix, PrintVector, PrintiVector, PrintMatrix src/matrix.h:27-45
 *   SplineInterp                                           src/colloc.h:103-105
 * The evaluation behind NPSOL's funobj/funcon callbacks runs on the GPU; see
 * ntg_b200.h for the batched entry points.
 */
#ifndef NTG_DROPIN_NTG_H_
#define NTG_DROPIN_NTG_H_

/* the reference's ntg.h drags these in; user code relies on that (kincar.c
 * uses assert() without including <assert.h>) */
#include <assert.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ntg_b200.h" /* AV and the callback typedefs */

#define MAXNOUT 5 /* reference src/ntg.h:22 (unused there too) */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct MatrixStruct {
    double **elements;
    int rows, cols;
} Matrix;

/*
 * Same 40 arguments, same order and meaning as reference src/ntg.h:72-99.
 * initialguess is overwritten with the solution (src/ntg.c:109).  Solving
 * needs NPSOL's npsol_/npoptn_ to be present in the process (they are looked
 * up at run time; NPSOL is separately licensed and never bundled).  If they
 * are absent the library's own solvers stand in: a problem with no nonlinear
 * constraints and only equality linear constraints (the class of both shipped
 * examples) goes to the reduced-space BFGS (ntgb_solve_eq), any other to the
 * augmented-Lagrangian driver (ntgb_solve_nlp); *inform = 0 / 1 / 4 / 6 with
 * NPSOL's meaning, istate / clambda / R untouched, a notice on stderr.  With
 * NTG_B200_NO_BUILTIN_SOLVER=1, or when the problem exceeds the solvers' limit
 * of 32 free directions, *inform is set to NTG_INFORM_NO_NPSOL and nothing is
 * solved.
 */
void ntg(int nout, double *bps, int nbps, int *kninterv, double **knots,
         int *order, int *mult, int *max_deriv,
         double *initialguess,
         int nlic, double **lic,
         int nltc, double **ltc,
         int nlfc, double **lfc,
         int nnlic, void (*nlicf)(int *, int *, double *, double **, double **),
         int nnltc, void (*nltcf)(int *, int *, int *, double *, double **, double **),
         int nnlfc, void (*nlfcf)(int *, int *, double *, double **, double **),
         int ninitialconstrav, AV *initialconstrav,
         int ntrajectoryconstrav, AV *trajectoryconstrav,
         int nfinalconstrav, AV *finalconstrav,
         double *lowerb, double *upperb,
         int nicf, void (*icf)(int *, int *, double *, double *, double **),
         int nucf, void (*ucf)(int *, int *, int *, double *, double *, double **),
         int nfcf, void (*fcf)(int *, int *, double *, double *, double **),
         int ninitialcostav, AV *initialcostav,
         int ntrajectorycostav, AV *trajectorycostav,
         int nfinalcostav, AV *finalcostav,
         int *istate, double *clambda, double *R,
         int *inform, double *objective);

#define NTG_INFORM_NO_NPSOL (-1000)
#define NTG_INFORM_SETUP_FAILED (-1001)

void npsoloption(const char *option);
void linspace(double *v, double d0, double d1, int n);
void printNTGBanner(void);

Matrix *MakeMatrix(int rows, int cols);
void FreeMatrix(Matrix *matrix);
double **DoubleMatrix(int rows, int cols);
void FreeDoubleMatrix(double **d);
void PrintMatrix(const char *filename, Matrix *matrix);
void PrintVector(const char *filename, double *f, int nf);
void PrintiVector(const char *filename, int *f, int nf);

void SplineInterp(double *f, double x, double *knots, int ninterv, double *coefs, int ncoefs,
                  int order, int mult, int maxderiv);

#ifdef __cplusplus
}
#endif
#endif /* NTG_DROPIN_NTG_H_ */
