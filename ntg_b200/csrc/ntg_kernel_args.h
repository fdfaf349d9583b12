/*
 * ntg_kernel_args.h -- the POD block handed from the core library
 * (ntg_core.cu) to a callback pack's launcher (ntg_eval_kernel.cuh).
 * Internal: not part of the public ABI, but both sides are built from this
 * header; ntgb_register_pack() refuses a pack whose NTGB_KERNEL_ABI differs
 * and every launcher checks ntgb_launch.abi again.
 */
#ifndef NTG_KERNEL_ARGS_H_
#define NTG_KERNEL_ARGS_H_

#include "ntg_b200.h"

#include <vector_types.h>

#define NTGB_KERNEL_ABI 13
#define NTGB_MAXOUT 8      /* outputs per problem the device tables can describe */
#define NTGB_MAXORDER 20   /* PGS bsplvb work arrays: jmax = 20   */
#define NTGB_MAXNLB 16     /* nonlinear bounds carried by value in the kernel params */

/* Device-resident, batch-shared problem description (built once by K0). */
typedef struct ntgb_devtab {
    int nout, nbps, nC, nz, nZ, ncnln, S;
    int nicf, nucf, nfcf, nnlic, nnltc, nnlfc;
    int order[NTGB_MAXOUT], mult[NTGB_MAXOUT], maxderiv[NTGB_MAXOUT], ncoef[NTGB_MAXOUT];
    int iC[NTGB_MAXOUT], iz[NTGB_MAXOUT], iZ[NTGB_MAXOUT], jk0[NTGB_MAXOUT];
    /* active-variable masks per breakpoint class; bit d of avmask[cls][j] set
     * <=> z[j][d] is computed there.  cls = (bp==0) | (bp==nbps-1)<<1.
     * (updateZ only fills listed variables, reference src/colloc.c:344-367) */
    unsigned avmask[4][NTGB_MAXOUT];
    int band_tile;                  /* breakpoints per tile of the band-compact Jacobian (ntg_b200.h) */
    const double *Bt[NTGB_MAXOUT];  /* [(k*maxderiv+d)*nbps + bp]  breakpoint-fastest      */
    const double *Bn[NTGB_MAXOUT];  /* [(bp*order+k)*maxderiv + d] reference block layout  */
    const int *off[NTGB_MAXOUT];    /* [bp] block offset (reference src/colloc.c:108)      */
    const double *bps;              /* [nbps]                                              */
    const double *nl_lb, *nl_ub;    /* [nnlic+nnltc+nnlfc] compact nonlinear bounds        */
    const int *col_lo, *col_hi;     /* [nC] first/last breakpoint whose band holds column  */
    /* runs of consecutive breakpoints with equal block offset, per output:
     * seg_start[j][s] .. seg_start[j][s+1]-1 share seg_off[j][s]; seg_start[j][nseg] = nbps */
    int nseg[NTGB_MAXOUT];
    const int *seg_start[NTGB_MAXOUT];
    const int *seg_off[NTGB_MAXOUT];
    const int *col_seg0;            /* [nC] run that holds breakpoint max(col_lo-1, 0)       */
    /* the same compact nonlinear bounds by value (constant bank), when they fit */
    int nl_inline;
    int one_table;                  /* every output has the same knots/order/mult/maxderiv   */
    /* quadrature plan of the cluster kernel (one_table only): for local column cl of an output,
     * entries plan[plan_ptr[cl] .. plan_ptr[cl+1]) list the breakpoints of its trapezoid chain in
     * ascending order: .x = breakpoint n, .y = (cluster rank that owns n << 24) | o24 with
     * o24 = k * bpc + (n - rank * bpc), the position of (band slot k, breakpoint n) inside one
     * output's block of that CTA's D array, or 0xffffff outside the band.  The first entry only
     * seeds the chain. */
    const int *plan_ptr;
    const int2 *plan;
    int plan_cl, plan_bpc;          /* cluster geometry the plan was built for               */
    int plan_cwin;                  /* doubles of coefficients the widest CTA stages per problem */
    int plan_n;                     /* number of plan entries                                */
    double nl_lb_v[NTGB_MAXNLB], nl_ub_v[NTGB_MAXNLB];
    /* quadrature schedule of the register-table kernel (K1s): the nC+1 trapezoid chains of a
     * problem (column nC = the scalar cost) packed into NS slots of near-equal length, longest
     * chain first, for NS = 1 .. sched_maxns.  Table NS starts at
     * sched + (NS-1)*(NTGB_SCHED_BLOCK+1 + nC+1): slot starts [NTGB_SCHED_BLOCK+1] (into the column
     * list), then the column list [nC+1]. */
    int sched_maxns;
    /* steady-state cluster kernel (K1c/H): a second plan follows the first in the same arrays --
     * plan_ptr[ncoef+1 .. 2*ncoef+2) and plan[plan_n ..) -- in which every column belongs to ONE
     * CTA (columns are dealt in contiguous ranges, rank r owns [ncoef*r/CL, ncoef*(r+1)/CL)) and
     * .y = k*(bpc + 2*plan_halo) + (n - r*bpc + plan_halo) is the position inside THAT CTA's D, whose
     * rows carry plan_halo extra breakpoints on both sides (computed redundantly by the CTA, so that
     * the quadrature never reads a neighbour's shared memory), or -1 outside the band; the plan of
     * a `fast` pack lists in-band entries only.  plan_halo < 0: no such plan. */
    int plan_halo;
    int plan_share;                 /* second plan: most entries any one rank owns */
    const int *sched;
    /* K1s steady state (funobj mode 2 + funcon mode 2): the tile-invariant shared-memory tables of
     * the kernel, laid out once.  img_w = [2][pitch+2] quadrature weights (dt, then node weights);
     * img_i = run starts [segtot], run offsets [segtot], the cost's run {0, nbps, 0, 0}, then 9 ints
     * per column 0..nC (the chain descriptions of ntg_eval_small.cuh, entry [4] WITHOUT its factor
     * G*R, the problems per tile, which only the launch knows) */
    const double *img_w;
    const int *img_i;
} ntgb_devtab;

#define NTGB_SCHED_BLOCK 256
#define NTGB_SCHED_MAXNS 32

/* cluster geometry of K1c for a horizon of nbps breakpoints: at most `per_cta` (<= 224) breakpoints per
 * CTA -- 7 warps pinned to breakpoints + 1 service warp = 256 threads.  Decided once at create time
 * (plan_cl, plan_bpc); the launchers read it from the tables. */
static inline void ntgb_cluster_geometry(int nbps, int per_cta, int *CL, int *bpc)
{
    int cl = nbps <= 2 * per_cta ? 2 : (nbps <= 4 * per_cta ? 4 : 8);
    *CL = cl;
    *bpc = (nbps + cl - 1) / cl;
}

typedef struct ntgb_launch {
    int abi;
    ntgb_devtab tab;
    ntgb_eval_args args; /* device pointers */
    int sm_count;
    int max_smem_optin;  /* bytes */
} ntgb_launch;

#endif
