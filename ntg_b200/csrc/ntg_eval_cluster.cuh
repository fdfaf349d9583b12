/*
 * ntg_eval_cluster.cuh -- K1c: the fused evaluator for LONG horizons whose
 * outputs all share one spline setup (CFG-5: 6 outputs, order 8, 401
 * breakpoints, nC = 4824, 706 KB of results per problem).
 *
 * Same math and reference citations as K1 / K1s.  The mapping is K1s's --
 * one thread pinned to one breakpoint for the whole launch, its table slice in
 * registers -- stretched over a THREAD-BLOCK CLUSTER: a problem's breakpoints
 * are split over the CL CTAs of a cluster (CL = 2/4/8, one CTA per SM), and
 * the quadrature chains of phase B read the band D[bp][slot] of neighbouring
 * CTAs through distributed shared memory where a column's support crosses a
 * CTA boundary.  Sync points are cluster barriers.
 *
 * Why (ncu, profiles/r01_v1_cfg5_general_ncu.txt): on this shape K1 needs 225
 * registers (7 warps/SM), re-reads a 616 KB table through L2 for every problem
 * and sits at 13.6 % of the HBM roofline, every FMA waiting on a load.
 *   - ONE table: all outputs have identical (knots, order, mult, maxderiv), so
 *     a thread keeps order*maxderiv doubles for its breakpoint, once.
 *   - row split: the constraint callback is inlined once per constraint row
 *     and only that row's derivatives are consumed (the compiler removes the
 *     rest), so one row of df (nz doubles) is live instead of nnltc rows.
 *   - coefficients of the next problem are staged with cp.async while the
 *     current one is evaluated.
 */
#ifndef NTG_EVAL_CLUSTER_CUH_
#define NTG_EVAL_CLUSTER_CUH_

#include <cooperative_groups.h>

#include "ntg_eval_small.cuh"

namespace ntgb {
namespace cg = cooperative_groups;

struct ClusterSmem {
    int bpc, nbps, S, cwin; /* cwin = doubles of coefficients one CTA stages per problem */
    int plan_n, plan_cols;  /* quadrature plan kept in shared memory (0 = left in global memory) */
    __host__ __device__ size_t D_off() const { return 0; }                              /* [S][bpc]          */
    __host__ __device__ size_t DI_off() const { return (size_t)S * bpc; }               /* [S]               */
    __host__ __device__ size_t DF_off() const { return DI_off() + S; }                  /* [S]               */
    __host__ __device__ size_t viol_off() const { return DF_off() + S; }                /* u64 [2]           */
    __host__ __device__ size_t sc_off() const { return viol_off() + 2; }                /* [2][2] cI,cF (rank 0) */
    __host__ __device__ size_t fall_off() const { return sc_off() + 4; }                /* [2][nbps] (rank 0) */
    __host__ __device__ size_t t_off() const { return fall_off() + 2 * (size_t)nbps; }  /* [nbps] (rank 0)   */
    __host__ __device__ size_t dt_off() const { return t_off() + nbps; }                /* [nbps]            */
    __host__ __device__ size_t wf_off() const { return dt_off() + nbps; }               /* [nbps] node weights (fast variant) */
    __host__ __device__ size_t C_off() const { return wf_off() + nbps; }                /* [2][cwin]         */
    __host__ __device__ size_t plan_off() const { return C_off() + 2 * (size_t)cwin; }   /* int2 [plan_n], int [plan_cols+1] */
    __host__ __device__ size_t bytes() const { return (plan_off() + plan_n) * 8 + (size_t)(plan_cols + 2) * 4 + 16; }
};

template <class PK>
__host__ __device__ constexpr bool pk_uniform_outputs()
{
    for (int j = 1; j < PK::kNout; j++)
        if (PK::md(j) != PK::md(0)) return false;
    return true;
}

template <class PK, bool FULL>
__global__ void __launch_bounds__(256, 1)
ntg_eval_cluster_kernel(const ntgb_devtab T, const ntgb_eval_args A, int CL, int bpc, int cwin, int plan_smem)
{
    constexpr int NOUT = PK::kNout;
    constexpr int NZ = pk_nz<PK>();
    constexpr int MD0 = PK::md(0);
    constexpr int NB = PK::kMaxOrd * MD0; /* ONE table */
    extern __shared__ double smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const ClusterSmem L{bpc, T.nbps, T.S, cwin, plan_smem ? T.plan_n : 0, plan_smem ? T.ncoef[0] : 0};
    const int nbps = T.nbps, nC = T.nC, P = A.P, S = T.S;
    double *D_s = smem + L.D_off();
    double *DI_s = smem + L.DI_off();
    double *DF_s = smem + L.DF_off();
    unsigned long long *viol_s = reinterpret_cast<unsigned long long *>(smem + L.viol_off()); /* [2], per CTA */
    double *sc_s = smem + L.sc_off();     /* rank 0 holds [buf][0] initial cost, [buf][1] final cost */
    double *fall_s = smem + L.fall_off(); /* rank 0 holds the integrand of ALL breakpoints, [buf][nbps] */
    double *t_s = smem + L.t_off();       /* rank 0: trapezoid terms of the scalar cost */
    double *dt_s = smem + L.dt_off();
    double *wf_s = smem + L.wf_off();     /* (dt[n-1] + dt[n])/2: the trapezoid rule as node weights */
    double *C_s = smem + L.C_off();
    double *fall0 = cluster.map_shared_rank(fall_s, 0);
    double *sc0 = cluster.map_shared_rank(sc_s, 0);

    const int mode_obj = A.mode_obj, mode_con = A.mode_con;
    const bool obj_on = mode_obj >= 0 && mode_obj <= 2;
    const bool con_on = mode_con >= 0 && mode_con <= 2 && T.ncnln > 0;
    const bool obj_d = obj_on && mode_obj != 0, obj_v = obj_on && mode_obj != 1;
    const bool con_d = con_on && mode_con != 0, con_v = con_on && mode_con != 1;
    /* mode 0 gates on count==1, modes 1/2 on count!=0 (reference src/ntg.c:297-302 vs :309-314) */
    const bool doI = PK::cb_icf != nullptr && obj_on && (mode_obj == 0 ? T.nicf == 1 : T.nicf != 0);
    const bool doU = PK::cb_ucf != nullptr && obj_on && (mode_obj == 0 ? T.nucf == 1 : T.nucf != 0);
    const bool doF = PK::cb_fcf != nullptr && obj_on && (mode_obj == 0 ? T.nfcf == 1 : T.nfcf != 0);
    const bool doCI = PK::cb_nlicf != nullptr && con_on && T.nnlic != 0;
    const bool doCT = PK::cb_nltcf != nullptr && con_on && T.nnltc != 0;
    const bool doCF = PK::cb_nlfcf != nullptr && con_on && T.nnlfc != 0;
    const bool wantJ = con_d && A.J != nullptr && A.jac_layout != NTGB_JAC_NONE;

    /* ---- once per CTA ---- */
    for (int i = threadIdx.x; i < nbps - 1; i += blockDim.x) dt_s[i] = __ldg(T.bps + i + 1) - __ldg(T.bps + i);
    for (int n = threadIdx.x; n < nbps; n += blockDim.x) {
        const double lo = n >= 1 ? (__ldg(T.bps + n) - __ldg(T.bps + n - 1)) * 0.5 : 0.0;
        const double hi = n + 1 < nbps ? (__ldg(T.bps + n + 1) - __ldg(T.bps + n)) * 0.5 : 0.0;
        wf_s[n] = lo + hi;
    }
    const int2 *plan = T.plan;
    const int *plan_ptr = T.plan_ptr;
    if (plan_smem) { /* the plan is re-read for every problem: keep it next to the data it indexes */
        int2 *pl_s = reinterpret_cast<int2 *>(smem + L.plan_off());
        int *pp_s = reinterpret_cast<int *>(pl_s + T.plan_n);
        for (int i = threadIdx.x; i < T.plan_n; i += blockDim.x) pl_s[i] = __ldg(T.plan + i);
        for (int i = threadIdx.x; i <= T.ncoef[0]; i += blockDim.x) pp_s[i] = __ldg(T.plan_ptr + i);
        plan = pl_s;
        plan_ptr = pp_s;
    }
    if (threadIdx.x == 0) {
        sc_s[0] = sc_s[1] = sc_s[2] = sc_s[3] = 0.0;
        viol_s[0] = 0ull;
        viol_s[1] = 0ull;
    }

    /* ---- once per thread: its breakpoint, its table slice ---- */
    const int bp0 = rank * bpc;
    const int lbp = threadIdx.x;            /* breakpoint index inside this CTA */
    const int bp = bp0 + lbp;
    const bool active = lbp < bpc && bp < nbps;
    const int cls = (bp == 0 ? 1 : 0) | (bp == nbps - 1 ? 2 : 0);
    double Bt[NB];
    int offj[NOUT];
    {
        const int order = T.order[0];
        const int o = active ? __ldg(T.off[0] + bp) : 0;
#pragma unroll
        for (int j = 0; j < NOUT; j++) offj[j] = o;
#pragma unroll
        for (int k = 0; k < PK::kMaxOrd; k++)
#pragma unroll
            for (int d = 0; d < MD0; d++)
                Bt[k * MD0 + d] = (active && k < order) ? __ldg(T.Bt[0] + (size_t)(k * MD0 + d) * nbps + bp) : 0.0;
    }

    /* this CTA only needs the coefficient windows of ITS breakpoints: per output the range
     * [w0, w0 + wl) with w0 = min offset, wl = max offset + order - w0 (same for every output) */
    __shared__ int win_s[2];
    if (threadIdx.x == 0) { win_s[0] = 0x7fffffff; win_s[1] = -1; }
    __syncthreads();
    if (active) {
        atomicMin(&win_s[0], offj[0]);
        atomicMax(&win_s[1], offj[0]);
    }
    __syncthreads();
    const int w0 = win_s[0];
    const int wl = win_s[1] + T.order[0] - w0;

    const int ncl = (int)(gridDim.x / CL);      /* clusters in the grid */
    const int clid = (int)(blockIdx.x / CL);
    auto stage_C = [&](int p, int buf) {
        const double *src = A.C + (size_t)p * nC + w0;
        for (int e = threadIdx.x; e < NOUT * wl; e += blockDim.x) {
            const int j = e / wl, q = e - j * wl;
            cp_async8(C_s + (size_t)buf * cwin + e, src + T.iC[j] + q);
        }
        cp_async_commit();
    };
    int buf = 0;
    if (clid < P) stage_C(clid, 0);

    /* The scalar cost is a sequential chain over all breakpoints (IntegrateVector TRAPEZOID,
     * src/integrator.c:21-24): the last warp of rank 0 -- an extra warp that owns no breakpoint --
     * finishes problem `pp` while everybody else is already in phase A of the next one. */
    const bool service = rank == 0 && (int)threadIdx.x >= (int)blockDim.x - 32;
    auto finish_cost = [&](int pp, int b) {
        const int lane = threadIdx.x & 31;
        const double *fp = fall_s + b * nbps;
        if (doU && obj_v) {
            for (int i = lane; i < nbps - 1; i += 32) t_s[i] = (dt_s[i] * (fp[i + 1] + fp[i])) / 2;
        }
        __syncwarp();
        if (lane == 0) {
            double In = 0.0;
            if (doU && obj_v) {
#pragma unroll 8
                for (int i = 0; i < nbps - 1; i++) In = In + t_s[i];
            }
            unsigned long long vb = 0ull;
            for (int r = 0; r < CL; r++) {
                unsigned long long *vr = cluster.map_shared_rank(viol_s, r) + b;
                const unsigned long long v = *vr;
                vb = v > vb ? v : vb;
                *vr = 0ull; /* next written two problems later, after two cluster barriers */
            }
            const double y = (sc_s[b * 2 + 0] + In) + sc_s[b * 2 + 1]; /* y = I + In + F, src/ntg.c:303,328 */
            if (obj_v && A.f != nullptr) A.f[pp] = y;
            if (want_result(A)) {
                put_result(A, (size_t)pp, 0, obj_v ? y : 0.0);
                put_result(A, (size_t)pp, 1, __longlong_as_double((long long)vb));
            }
        }
        __syncwarp();
    };

    int pprev = -1;
    for (int p = clid; p < P; p += ncl, buf ^= 1) {
        cp_async_wait_all();
        cluster.sync(); /* this problem's coefficients landed; every CTA is done with the previous phase B */
        if (p + ncl < P) stage_C(p + ncl, buf ^ 1);
        if (service && pprev >= 0) finish_cost(pprev, buf ^ 1);
        pprev = p;

        /* ---------------- phase A: this thread's breakpoint ---------------- */
        double z[NZ]; /* flat outputs of this thread's breakpoint; live across the split barrier */
        if (active) {
            const double *Cp = C_s + (size_t)buf * cwin;
            double *zp[NOUT];
            /* Zvalue, src/colloc.c:318-326 -- k ascending from 0.0 */
            static_for<0, NOUT>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                constexpr int IZ = pk_iz<PK>(j);
                const int order = FULL ? PK::kMaxOrd : T.order[j];
                const double *Cw = Cp + j * wl + (offj[j] - w0);
                const unsigned mask = T.avmask[cls][j];
                double acc[MD0];
#pragma unroll
                for (int d = 0; d < MD0; d++) acc[d] = 0.0;
#pragma unroll
                for (int k = 0; k < PK::kMaxOrd; k++) {
                    if (FULL || k < order) {
                        const double ck = Cw[k];
#pragma unroll
                        for (int d = 0; d < MD0; d++) acc[d] = acc[d] + Bt[k * MD0 + d] * ck;
                    }
                }
#pragma unroll
                for (int d = 0; d < MD0; d++) z[IZ + d] = ((mask >> d) & 1u) ? acc[d] : 0.0;
                zp[j] = &z[IZ];
                if (A.Z != nullptr) {
#pragma unroll
                    for (int d = 0; d < MD0; d++) A.Z[(size_t)p * T.nZ + T.iZ[j] + (size_t)bp * MD0 + d] = z[IZ + d];
                }
            });

            int nstate = A.nstate;

            /* unintegrated (trajectory) cost, src/cost.c:99-132 */
            if constexpr (PK::cb_ucf != nullptr) {
                if (doU) {
                    double fv = 0.0;
                    double df[NZ];
#pragma unroll
                    for (int l = 0; l < NZ; l++) df[l] = 0.0;
                    int mode = mode_obj, i = bp;
                    PK::cb_ucf(&mode, &nstate, &i, &fv, df, zp);
                    note_abort(A, mode);
                    fall0[buf * nbps + bp] = fv; /* distributed shared memory: rank 0 collects the integrand */
                    if (obj_d) {
                        double *Dp = D_s + lbp;
                        band_from_regs<PK, FULL, true>(T, Bt, df, [&](auto, auto, double v) {
                            *Dp = v;
                            Dp += bpc;
                        });
                    }
                }
            }
            /* initial cost (breakpoint 0: cluster rank 0), src/cost.c:4-36 */
            if constexpr (PK::cb_icf != nullptr) {
                if (doI && bp == 0) {
                    double fv = 0.0;
                    double df[NZ];
#pragma unroll
                    for (int l = 0; l < NZ; l++) df[l] = 0.0;
                    int mode = mode_obj;
                    PK::cb_icf(&mode, &nstate, &fv, df, zp);
                    note_abort(A, mode);
                    sc0[buf * 2 + 0] = fv;
                    if (obj_d) {
                        double *Dp = DI_s;
                        band_from_regs<PK, FULL, true>(T, Bt, df, [&](auto, auto, double v) { *Dp++ = v; });
                    }
                }
            }
            /* final cost (last breakpoint: last cluster rank), src/cost.c:141-174 */
            if constexpr (PK::cb_fcf != nullptr) {
                if (doF && bp == nbps - 1) {
                    double fv = 0.0;
                    double df[NZ];
#pragma unroll
                    for (int l = 0; l < NZ; l++) df[l] = 0.0;
                    int mode = mode_obj;
                    PK::cb_fcf(&mode, &nstate, &fv, df, zp);
                    note_abort(A, mode);
                    sc0[buf * 2 + 1] = fv;
                    if (obj_d) {
                        double *Dp = DF_s;
                        band_from_regs<PK, FULL, true>(T, Bt, df, [&](auto, auto, double v) { *Dp++ = v; });
                    }
                }
            }
                }
        /* the quadrature inputs (D, integrand, end-point terms) are written: ARRIVE at the cluster
         * barrier now and wait only after the Jacobian rows are streamed out, so the release does
         * not have to drain ~200 outstanding global stores per thread (ncu: 14 % of samples sat in
         * the fence of a plain cluster.sync() placed after them) */
        cluster.barrier_arrive();
        if (active) {
            double viol = 0.0;
            int nstate = A.nstate;
            double *zp[NOUT];
            static_for<0, NOUT>([&](auto jc) { zp[decltype(jc)::value] = &z[pk_iz<PK>(decltype(jc)::value)]; });

            /* nonlinear trajectory constraints, src/constraints.c:120-162, one ROW per inlined
             * callback: only row m's value and derivatives are consumed in iteration m */
            if constexpr (PK::cb_nltcf != nullptr && PK::kNnltc > 0) {
                if (doCT) {
                    static_for<0, PK::kNnltc>([&](auto mc) {
                        constexpr int m = decltype(mc)::value;
                        double cv[PK::kNnltc];
                        double dfc[PK::kNnltc][NZ];
                        double *dfp[PK::kNnltc];
#pragma unroll
                        for (int q = 0; q < PK::kNnltc; q++) {
                            cv[q] = 0.0;
                            dfp[q] = dfc[q];
#pragma unroll
                            for (int l = 0; l < NZ; l++) dfc[q][l] = 0.0;
                        }
                        int mode = mode_con, i = bp;
                        PK::cb_nltcf(&mode, &nstate, &i, cv, dfp, zp);
                        note_abort(A, mode);
                        if (con_v) {
                            if (A.c != nullptr)
                                st_stream(A.c + (size_t)p * T.ncnln + T.nnlic + (size_t)m * nbps + bp, cv[m]);
                            viol = fmax(viol, row_violation(cv[m], nl_bound(T, false, T.nnlic + m),
                                                            nl_bound(T, true, T.nnlic + m)));
                        }
                        if (wantJ) {
                            if (A.jac_layout == NTGB_JAC_BAND) {
                                /* tiled band layout, one tile per CTA (include/ntg_b200.h): this CTA's
                                 * rows are cnt = min(bpc, nbps - bp0) long and contiguous */
                                const int cnt = nbps - bp0 < bpc ? nbps - bp0 : bpc;
                                double *ptr = A.J + (size_t)p * T.ncnln * S + (size_t)T.nnlic * S +
                                              (size_t)rank * T.nnltc * S * bpc + (size_t)m * S * cnt + lbp;
                                band_from_regs<PK, FULL, true>(T, Bt, dfc[m], [&](auto, auto, double v) {
                                    st_stream(ptr, v);
                                    ptr += cnt;
                                });
                            } else {
                                double *Jp = A.J + (size_t)p * T.ncnln * nC;
                                const int row = T.nnlic + m * nbps + bp;
                                band_from_regs<PK, FULL, true>(T, Bt, dfc[m], [&](auto jc, auto kc, double v) {
                                    constexpr int j = decltype(jc)::value;
                                    constexpr int k = decltype(kc)::value;
                                    st_stream(Jp + (size_t)(T.iC[j] + offj[j] + k) * T.ncnln + row, v);
                                });
                            }
                        }
                    });
                }
            }
            /* nonlinear initial constraints (breakpoint 0), src/constraints.c:88-117 */
            if constexpr (PK::cb_nlicf != nullptr && PK::kNnlic > 0) {
                if (doCI && bp == 0) {
                    double cv[PK::kNnlic];
                    double dfc[PK::kNnlic][NZ];
                    double *dfp[PK::kNnlic];
#pragma unroll
                    for (int m = 0; m < PK::kNnlic; m++) {
                        cv[m] = 0.0;
                        dfp[m] = dfc[m];
#pragma unroll
                        for (int l = 0; l < NZ; l++) dfc[m][l] = 0.0;
                    }
                    int mode = mode_con;
                    PK::cb_nlicf(&mode, &nstate, cv, dfp, zp);
                    note_abort(A, mode);
                    if (con_v) {
#pragma unroll
                        for (int m = 0; m < PK::kNnlic; m++) {
                            if (A.c != nullptr) st_stream(A.c + (size_t)p * T.ncnln + m, cv[m]);
                            viol = fmax(viol, row_violation(cv[m], nl_bound(T, false, m), nl_bound(T, true, m)));
                        }
                    }
                    if (wantJ) emit_rows_regs<PK, FULL, PK::kNnlic, 0, true>(T, A, Bt, offj, p, bp, dfc, 0);
                }
            }
            /* nonlinear final constraints (last breakpoint), src/constraints.c:165-195 */
            if constexpr (PK::cb_nlfcf != nullptr && PK::kNnlfc > 0) {
                if (doCF && bp == nbps - 1) {
                    double cv[PK::kNnlfc];
                    double dfc[PK::kNnlfc][NZ];
                    double *dfp[PK::kNnlfc];
#pragma unroll
                    for (int m = 0; m < PK::kNnlfc; m++) {
                        cv[m] = 0.0;
                        dfp[m] = dfc[m];
#pragma unroll
                        for (int l = 0; l < NZ; l++) dfc[m][l] = 0.0;
                    }
                    int mode = mode_con;
                    const int rb = T.nnlic + T.nnltc * nbps;
                    PK::cb_nlfcf(&mode, &nstate, cv, dfp, zp);
                    note_abort(A, mode);
                    if (con_v) {
#pragma unroll
                        for (int m = 0; m < PK::kNnlfc; m++) {
                            if (A.c != nullptr) st_stream(A.c + (size_t)p * T.ncnln + rb + m, cv[m]);
                            viol = fmax(viol, row_violation(cv[m], nl_bound(T, false, T.nnlic + T.nnltc + m),
                                                            nl_bound(T, true, T.nnlic + T.nnltc + m)));
                        }
                    }
                    if (wantJ) emit_rows_regs<PK, FULL, PK::kNnlfc, 2, true>(T, A, Bt, offj, p, bp, dfc, rb);
                }
            }
            if (viol > 0.0) atomicMax(viol_s + buf, (unsigned long long)__double_as_longlong(viol));

        }
        cluster.barrier_wait(); /* every CTA's D, f and scalars are visible cluster-wide */

        /* ------- phase B: one trapezoid chain per gradient column (IntegrateFMatrixCols TRAPEZOID,
         * src/integrator.c:44-48, on the band of src/cost.c:118-132; ascending breakpoint), spread
         * over the whole cluster, walking the host-built plan; D of breakpoint n lives in the CTA
         * that owns n (distributed shared memory).  Then Vector3Add, src/ntg.c:329. ------- */
        if (obj_d && A.g != nullptr) {
            const int ncoef0 = T.ncoef[0];
            const int order0 = FULL ? PK::kMaxOrd : T.order[0];
            const int last_rank = (nbps - 1) / bpc;
            const int off_last = __ldg(T.off[0] + nbps - 1);
            const int jpitch = order0 * bpc; /* one output's block of D */
            /* one item = local column cl of a GROUP of outputs: the plan (which breakpoints, which
             * band position) is shared by all outputs, so it is decoded once and drives JG independent
             * chains (measured: groups of 2 outputs balance the threads better but lose more to the extra decodes) */
            constexpr int JG = NOUT;
            constexpr int NG = NOUT / JG;
            const int nitems = ncoef0 * NG;
            for (int it = rank * (int)blockDim.x + (int)threadIdx.x; it < nitems; it += CL * (int)blockDim.x) {
                const int jg = it / ncoef0;
                const int cl = it - jg * ncoef0;
                const int jbase = jg * JG * jpitch; /* first output of the group inside D */
                double gU[JG], dcur[JG];
#pragma unroll
                for (int j = 0; j < JG; j++) { gU[j] = 0.0; dcur[j] = 0.0; }
                if (doU && !PK::kExact) {
                    /* fast variant: sum_n Wf[n]*D[n] over the plan's in-band entries (the band is zero at
                     * both ends of a column's support, so the node-weight form needs no end corrections) */
                    const int eend = plan_ptr[cl + 1];
#pragma unroll 4
                    for (int e = plan_ptr[cl]; e < eend; e++) {
                        const int2 en = plan[e];
                        const int o24 = en.y & 0xffffff, r = en.y >> 24;
                        if (o24 == 0xffffff) continue;
                        const double w = wf_s[en.x];
                        if (r == rank) {
#pragma unroll
                            for (int j = 0; j < JG; j++) gU[j] = gU[j] + w * D_s[jbase + j * jpitch + o24];
                        } else {
                            const double *base = cluster.map_shared_rank(D_s, r) + jbase + o24;
#pragma unroll
                            for (int j = 0; j < JG; j++) gU[j] = gU[j] + w * base[j * jpitch];
                        }
                    }
                } else if (doU) {
                    int e = plan_ptr[cl];
                    const int eend = plan_ptr[cl + 1];
                    if (e < eend) {
                        int2 en = plan[e];
                        {
                            const int o24 = en.y & 0xffffff, r = en.y >> 24;
                            if (o24 != 0xffffff) {
                                if (r == rank) { /* own shared memory: plain LDS */
#pragma unroll
                                    for (int j = 0; j < JG; j++) dcur[j] = D_s[jbase + j * jpitch + o24];
                                } else {         /* neighbour CTA: distributed shared memory */
                                    const double *base = cluster.map_shared_rank(D_s, r) + jbase + o24;
#pragma unroll
                                    for (int j = 0; j < JG; j++) dcur[j] = base[j * jpitch];
                                }
                            }
                        }
                        for (e++; e < eend; e++) {
                            en = plan[e];
                            const int o24 = en.y & 0xffffff, r = en.y >> 24;
                            const double dt = dt_s[en.x - 1];
                            double dn[JG];
                            if (o24 == 0xffffff) {
#pragma unroll
                                for (int j = 0; j < JG; j++) dn[j] = 0.0;
                            } else if (r == rank) {
#pragma unroll
                                for (int j = 0; j < JG; j++) dn[j] = D_s[jbase + j * jpitch + o24];
                            } else {
                                const double *base = cluster.map_shared_rank(D_s, r) + jbase + o24;
#pragma unroll
                                for (int j = 0; j < JG; j++) dn[j] = base[j * jpitch];
                            }
#pragma unroll
                            for (int j = 0; j < JG; j++) {
                                gU[j] = gU[j] + (dt * (dn[j] + dcur[j])) / 2;
                                dcur[j] = dn[j];
                            }
                        }
                    }
                }
                const int kF = cl - off_last;
#pragma unroll
                for (int j = 0; j < JG; j++) {
                    const int jo = jg * JG + j;
                    const int s0 = jo * order0; /* jk0_j: every output has the same order */
                    double gI = 0.0, gF = 0.0;
                    if (doI && cl < order0) gI = cluster.map_shared_rank(DI_s, 0)[s0 + cl]; /* offset 0, src/colloc.c:254 */
                    if (doF && kF >= 0 && kF < order0) gF = cluster.map_shared_rank(DF_s, last_rank)[s0 + kF];
                    st_stream(A.g + (size_t)p * nC + (size_t)jo * ncoef0 + cl, (gI + gU[j]) + gF);
                }
            }
        }
    }
    cp_async_wait_all();
    cluster.sync(); /* last phase A complete everywhere */
    if (service && pprev >= 0) finish_cost(pprev, buf ^ 1);
    cluster.sync(); /* nobody may exit while a neighbour still reads its shared memory */
}

/* outputs all share one spline setup? (checked on the host at launch) */
inline bool devtab_one_table(const ntgb_devtab &T)
{
    return T.one_table != 0;
}

template <class PK>
int launch_eval_cluster(const ntgb_launch *L)
{
    if constexpr (!pk_uniform_outputs<PK>() || PK::kMaxOrd * PK::md(0) > 64) {
        return -1001;
    } else {
        const ntgb_devtab &T = L->tab;
        const int nbps = T.nbps, P = L->args.P;
        if (!devtab_one_table(T)) return -1001;
        /* cluster geometry and plan were decided when the tables were built (ntg_core.cu) */
        const int CL = T.plan_cl, bpc = T.plan_bpc;
        if (T.plan == nullptr || CL < 2 || T.band_tile != bpc) return -1001;
        const int block = (bpc + 31) / 32 * 32 + 32; /* + one service warp (scalar cost of the previous problem) */
        if (block > 256) return -1001;
        bool full = true;
        for (int j = 0; j < T.nout; j++) full = full && T.order[j] == PK::kMaxOrd;
        ClusterSmem lay{bpc, nbps, T.S, T.plan_cwin, T.plan_n, T.ncoef[0]};
        int plan_smem = 1;
        if (lay.bytes() > (size_t)L->max_smem_optin) { /* plan stays in global memory */
            lay.plan_n = 0;
            lay.plan_cols = 0;
            plan_smem = 0;
        }
        const size_t smem = lay.bytes();
        if (smem > (size_t)L->max_smem_optin) return -1001;
        auto kern = full ? ntg_eval_cluster_kernel<PK, true> : ntg_eval_cluster_kernel<PK, false>;
        cudaError_t e = raise_smem_limit((const void *)kern, L->max_smem_optin);
        if (e != cudaSuccess) return (int)e;
        int nclusters = L->sm_count / CL;
        if (nclusters > P) nclusters = P;
        if (nclusters < 1) return 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(nclusters * CL));
        cfg.blockDim = dim3((unsigned)block);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)L->args.stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return (int)cudaLaunchKernelEx(&cfg, kern, T, L->args, CL, bpc, T.plan_cwin, plan_smem);
    }
}

} /* namespace ntgb */
#endif
