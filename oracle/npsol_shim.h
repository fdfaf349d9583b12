/*
 * oracle/npsol_shim.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Request block of the fake npsol_() (oracle/npsol_shim.c).
 */
#ifndef ORACLE_NPSOL_SHIM_H_
#define ORACLE_NPSOL_SHIM_H_
typedef struct {
    int P;
    const double *X; /* [P][n] */
    int mode_obj, mode_con;
    double *f;       /* [P]            */
    double *g;       /* [P][n]         */
    double *c;       /* [P][ncnln]     */
    double *Jdense;  /* [P][ncnln*n] column-major per problem, NaN = unwritten */
    double *Jband;   /* [P][ncnln][S]  row-major band values                   */
    const int *col0; /* [ncnln][nout]  first column of each output band        */
    const int *order;
    int nout, S;
    long pattern_bad; /* written-outside-band + unwritten-inside-band entries  */
    double *A;        /* [nclin*n] column-major copy of NPSOL's A              */
    double *bl, *bu;  /* [n+nclin+ncnln]                                       */
    int reps;
    double best_seconds;
    int n, nclin, ncnln;
    int nan_fill;     /* pre-fill cJac with NaN to extract the written pattern */
    int calls;        /* number of times npsol_ was entered                    */
} shim_request;

void shim_set_request(shim_request *r);
shim_request *shim_get_request(void);
#endif
