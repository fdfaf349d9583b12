/*
 * ntg_eval_kernel.cuh -- K1, the fused batched collocation evaluator for
 * sm_100a, instantiated once per callback pack (template parameter PK carries
 * the user's cost / constraint callbacks as __device__ functions with NTG's
 * unchanged signatures, reference src/ntg.c:34-41).
 *
 * One launch computes, for P independent coefficient vectors, everything the
 * reference computes per SQP iterate in NPfuncon + NPfunobj
 * (src/ntg.c:274-371): flat outputs z (updateZ, src/colloc.c:318-367), the
 * user callbacks at every breakpoint, constraint values and the banded
 * Jacobian (src/constraints.c:36-195, src/colloc.c:243-316), the trapezoid
 * cost and its gradient (src/cost.c, src/integrator.c) -- with no dense
 * scratch matrices and no allocation.
 *
 * Decomposition (DESIGN.md "K1"):
 *   tile  = G consecutive problems of the batch, one CTA per tile (grid-stride)
 *   phase A: one thread per (problem, breakpoint) point.  Lanes are
 *            consecutive breakpoints, so table reads (Bt, breakpoint-fastest)
 *            and Jacobian / constraint stores (breakpoint-fastest band layout
 *            or NPSOL's column-major dense layout) are coalesced.
 *   phase B: one thread per (problem, coefficient column): the trapezoid
 *            quadrature of the cost gradient as a GATHER over the breakpoints
 *            whose band holds that column, in the reference's own summation
 *            order (ascending breakpoint), from cost derivatives staged in
 *            shared memory by phase A.  One more thread per problem sums the
 *            scalar cost.  No atomics, deterministic, and -- when the pack is
 *            built with -fmad=false ("exact") -- bit-identical to the
 *            reference for everything except libm calls inside callbacks.
 *
 * The path is HBM-bound FP64 (1.4-2.2 flop/byte):
 * CUDA-core DFMA/DMUL/DADD, no tensor cores (tcgen05 has no f64 kind).
 */
#ifndef NTG_EVAL_KERNEL_CUH_
#define NTG_EVAL_KERNEL_CUH_

#include <cuda_runtime.h>
#include <cstdio>
#include <type_traits>

#include "ntg_kernel_args.h"

namespace ntgb {

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

/* compile-time prefix sum of the pack's derivative depths: iz(j), src/colloc.c:44 */
template <class PK>
__host__ __device__ constexpr int pk_iz(int j)
{
    int s = 0;
    for (int q = 0; q < j; q++) s += PK::md(q);
    return s;
}
template <class PK>
__host__ __device__ constexpr int pk_nz() { return pk_iz<PK>(PK::kNout); }

__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }

/* one entry (which = 0 objective, 1 violation) of problem p's result pair: the local table and, for
 * the fused multi-GPU gather, every rank's copy of the gathered table (peer stores over NVLink) */
template <bool PEERS = true> /* PEERS = false: a specialisation that is known to run without peer tables */
__device__ __forceinline__ void put_result(const ntgb_eval_args &A, size_t p, int which, double v)
{
    if (A.result != nullptr) A.result[2 * p + which] = v;
    if constexpr (PEERS)
        for (int r = 0; r < A.npeers; r++) A.peer_result[r][2 * ((size_t)A.peer_row0 + p) + which] = v;
}
template <bool PEERS = true>
__device__ __forceinline__ bool want_result(const ntgb_eval_args &A)
{
    return A.result != nullptr || (PEERS && A.npeers > 0);
}

/* a callback that writes *mode = -1 asks the solver to stop (reference src/ntg.c:369) */
__device__ __forceinline__ void note_abort(const ntgb_eval_args &A, int mode)
{
    if (mode < 0 && A.abort_flag != nullptr) *A.abort_flag = 1;
}

/* shared-memory carve-up for one tile of G problems */
struct SmemLayout {
    int G, nbps, nz;
    __host__ __device__ size_t f_off() const { return 0; }                                   /* [G][nbps]      */
    __host__ __device__ size_t df_off() const { return (size_t)G * nbps; }                    /* [nz][G][nbps]  */
    __host__ __device__ size_t dfI_off() const { return df_off() + (size_t)nz * G * nbps; }   /* [G][nz]        */
    __host__ __device__ size_t dfF_off() const { return dfI_off() + (size_t)G * nz; }         /* [G][nz]        */
    __host__ __device__ size_t cI_off() const { return dfF_off() + (size_t)G * nz; }          /* [G]            */
    __host__ __device__ size_t cF_off() const { return cI_off() + G; }                        /* [G]            */
    __host__ __device__ size_t viol_off() const { return cF_off() + G; }                      /* [G] u64 bits   */
    __host__ __device__ size_t doubles() const { return viol_off() + G; }
};

/*
 * Emit NCON Jacobian rows evaluated at breakpoint `bp`:
 *   J[row][col0_j + k] = sum_l dfc[m][iz_j + l] * B_j[bp][k][l]   (l ascending)
 * KIND 0: initial rows (col0 = iC_j, src/colloc.c:254), 1: trajectory rows
 * (row = base + m*nbps + bp), 2: final rows.
 */
template <class PK, int NCON, int KIND>
__device__ __forceinline__ void emit_jac_rows(const ntgb_devtab &T, const ntgb_eval_args &A, int p,
                                              int bp, const double (&dfc)[NCON][pk_nz<PK>()],
                                              int row_base)
{
    if (A.J == nullptr || A.jac_layout == NTGB_JAC_NONE) return;
    const int nbps = T.nbps;
    const size_t pbase_band = (size_t)p * T.ncnln * T.S;
    const size_t pbase_dense = (size_t)p * T.ncnln * T.nC;
    static_for<0, PK::kNout>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr int MD = PK::md(j);
        constexpr int IZ = pk_iz<PK>(j);
        const int order = T.order[j];
        const double *__restrict__ Bt = T.Bt[j];
        const int off = (KIND == 0) ? 0 : __ldg(T.off[j] + bp);
        const int col0 = T.iC[j] + off;
        const int s0 = T.jk0[j];
#pragma unroll
        for (int k = 0; k < PK::kMaxOrd; k++) {
            if (k < order) {
                double b[MD];
#pragma unroll
                for (int l = 0; l < MD; l++) b[l] = __ldg(Bt + (size_t)(k * MD + l) * nbps + bp);
#pragma unroll
                for (int m = 0; m < NCON; m++) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < MD; l++) acc = acc + dfc[m][IZ + l] * b[l];
                    if (A.jac_layout == NTGB_JAC_BAND) {
                        size_t idx;
                        if (KIND == 1) /* tiled, breakpoint-fastest (include/ntg_b200.h); row_base = nnlic */
                            idx = pbase_band + ntgb_band_index(row_base, T.nnltc, T.S, nbps, T.band_tile, m, s0 + k, bp);
                        else
                            idx = pbase_band + (size_t)(row_base + m) * T.S + s0 + k;
                        st_stream(A.J + idx, acc);
                    } else {
                        const int row = (KIND == 1) ? row_base + m * nbps + bp : row_base + m;
                        st_stream(A.J + pbase_dense + (size_t)(col0 + k) * T.ncnln + row, acc);
                    }
                }
            }
        }
    });
}

__device__ __forceinline__ double row_violation(double c, double lb, double ub)
{
    double v = 0.0;
    if (lb - c > v) v = lb - c;
    if (c - ub > v) v = c - ub;
    return v;
}

template <class PK>
__global__ void __launch_bounds__(256) ntg_eval_kernel(const ntgb_devtab T, const ntgb_eval_args A, int G)
{
    constexpr int NOUT = PK::kNout;
    constexpr int NZ = pk_nz<PK>();
    extern __shared__ double smem[];
    const SmemLayout L{G, T.nbps, NZ};
    double *f_s = smem + L.f_off();
    double *df_s = smem + L.df_off();
    double *dfI_s = smem + L.dfI_off();
    double *dfF_s = smem + L.dfF_off();
    double *cI_s = smem + L.cI_off();
    double *cF_s = smem + L.cF_off();
    unsigned long long *viol_s = reinterpret_cast<unsigned long long *>(smem + L.viol_off());

    const int nbps = T.nbps, nC = T.nC, P = A.P;
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
    const double *__restrict__ bps = T.bps;

    const int ntiles = (P + G - 1) / G;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p0 = tile * G;
        for (int q = threadIdx.x; q < G; q += blockDim.x) {
            viol_s[q] = 0ull;
            cI_s[q] = 0.0;
            cF_s[q] = 0.0;
        }
        if (doI || doF)
            for (int q = threadIdx.x; q < 2 * G * NZ; q += blockDim.x) dfI_s[q] = 0.0;
        __syncthreads();

        /* ---------------- phase A: one thread per (problem, breakpoint) ---------------- */
        for (int q = threadIdx.x; q < G * nbps; q += blockDim.x) {
            const int pl = q / nbps;
            const int bp = q - pl * nbps;
            const int p = p0 + pl;
            if (p >= P) continue;
            const int cls = (bp == 0 ? 1 : 0) | (bp == nbps - 1 ? 2 : 0);
            const double *__restrict__ Cp = A.C + (size_t)p * nC;

            /* z = flat outputs and derivatives at this breakpoint:
             * Zvalue, src/colloc.c:318-326 -- k ascending from 0.0 */
            double z[NZ > 0 ? NZ : 1];
            double *zp[NOUT];
            static_for<0, NOUT>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                constexpr int MD = PK::md(j);
                constexpr int IZ = pk_iz<PK>(j);
                const int order = T.order[j];
                const double *__restrict__ Bt = T.Bt[j];
                const double *__restrict__ Cw = Cp + T.iC[j] + __ldg(T.off[j] + bp);
                const unsigned mask = T.avmask[cls][j];
                double acc[MD];
#pragma unroll
                for (int d = 0; d < MD; d++) acc[d] = 0.0;
#pragma unroll
                for (int k = 0; k < PK::kMaxOrd; k++) {
                    if (k < order) {
                        const double ck = __ldg(Cw + k);
#pragma unroll
                        for (int d = 0; d < MD; d++)
                            acc[d] = acc[d] + __ldg(Bt + (size_t)(k * MD + d) * nbps + bp) * ck;
                    }
                }
#pragma unroll
                for (int d = 0; d < MD; d++) z[IZ + d] = ((mask >> d) & 1u) ? acc[d] : 0.0;
                zp[j] = &z[IZ];
                if (A.Z != nullptr) {
#pragma unroll
                    for (int d = 0; d < MD; d++)
                        A.Z[(size_t)p * T.nZ + T.iZ[j] + (size_t)bp * MD + d] = z[IZ + d];
                }
            });

            double viol = 0.0;
            int nstate = A.nstate;

            /* nonlinear trajectory constraints, src/constraints.c:120-162 */
            if constexpr (PK::cb_nltcf != nullptr && PK::kNnltc > 0) {
                if (doCT) {
                    double cv[PK::kNnltc];
                    double dfc[PK::kNnltc][NZ];
                    double *dfp[PK::kNnltc];
#pragma unroll
                    for (int m = 0; m < PK::kNnltc; m++) {
                        cv[m] = 0.0;
                        dfp[m] = dfc[m];
#pragma unroll
                        for (int l = 0; l < NZ; l++) dfc[m][l] = 0.0;
                    }
                    int mode = mode_con, i = bp;
                    PK::cb_nltcf(&mode, &nstate, &i, cv, dfp, zp);
                    note_abort(A, mode);
                    if (con_v) {
#pragma unroll
                        for (int m = 0; m < PK::kNnltc; m++) {
                            if (A.c != nullptr)
                                st_stream(A.c + (size_t)p * T.ncnln + T.nnlic + (size_t)m * nbps + bp, cv[m]);
                            viol = fmax(viol, row_violation(cv[m], __ldg(T.nl_lb + T.nnlic + m),
                                                            __ldg(T.nl_ub + T.nnlic + m)));
                        }
                    }
                    if (con_d) emit_jac_rows<PK, PK::kNnltc, 1>(T, A, p, bp, dfc, T.nnlic);
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
                            viol = fmax(viol, row_violation(cv[m], __ldg(T.nl_lb + m), __ldg(T.nl_ub + m)));
                        }
                    }
                    if (con_d) emit_jac_rows<PK, PK::kNnlic, 0>(T, A, p, 0, dfc, 0);
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
                            viol = fmax(viol, row_violation(cv[m], __ldg(T.nl_lb + T.nnlic + T.nnltc + m),
                                                            __ldg(T.nl_ub + T.nnlic + T.nnltc + m)));
                        }
                    }
                    if (con_d) emit_jac_rows<PK, PK::kNnlfc, 2>(T, A, p, bp, dfc, rb);
                }
            }
            if (viol > 0.0) atomicMax(&viol_s[pl], (unsigned long long)__double_as_longlong(viol));

            /* unintegrated (trajectory) cost, src/cost.c:99-110 */
            if constexpr (PK::cb_ucf != nullptr) {
                if (doU) {
                    double fv = 0.0;
                    double df[NZ > 0 ? NZ : 1];
#pragma unroll
                    for (int l = 0; l < NZ; l++) df[l] = 0.0;
                    int mode = mode_obj, i = bp;
                    PK::cb_ucf(&mode, &nstate, &i, &fv, df, zp);
                    note_abort(A, mode);
                    f_s[q] = fv;
                    if (obj_d) {
#pragma unroll
                        for (int l = 0; l < NZ; l++) df_s[(size_t)l * G * nbps + q] = df[l];
                    }
                }
            }
            /* initial cost (breakpoint 0), src/cost.c:4-36 */
            if constexpr (PK::cb_icf != nullptr) {
                if (doI && bp == 0) {
                    double fv = 0.0;
                    double df[NZ > 0 ? NZ : 1];
#pragma unroll
                    for (int l = 0; l < NZ; l++) df[l] = 0.0;
                    int mode = mode_obj;
                    PK::cb_icf(&mode, &nstate, &fv, df, zp);
                    note_abort(A, mode);
                    cI_s[pl] = fv;
#pragma unroll
                    for (int l = 0; l < NZ; l++) dfI_s[pl * NZ + l] = df[l];
                }
            }
            /* final cost (last breakpoint), src/cost.c:141-174 */
            if constexpr (PK::cb_fcf != nullptr) {
                if (doF && bp == nbps - 1) {
                    double fv = 0.0;
                    double df[NZ > 0 ? NZ : 1];
#pragma unroll
                    for (int l = 0; l < NZ; l++) df[l] = 0.0;
                    int mode = mode_obj;
                    PK::cb_fcf(&mode, &nstate, &fv, df, zp);
                    note_abort(A, mode);
                    cF_s[pl] = fv;
#pragma unroll
                    for (int l = 0; l < NZ; l++) dfF_s[pl * NZ + l] = df[l];
                }
            }
        }
        __syncthreads();

        /* ------- phase B: one thread per (problem, column) + one per problem ------- */
        const int items = G * (nC + 1);
        for (int q = threadIdx.x; q < items; q += blockDim.x) {
            const int pl = q / (nC + 1);
            const int c = q - pl * (nC + 1);
            const int p = p0 + pl;
            if (p >= P) continue;
            if (c == nC) {
                /* scalar cost: IntegrateVector TRAPEZOID, src/integrator.c:21-24, then
                 * y = I + In + F, src/ntg.c:303,328 */
                if (obj_v || want_result(A)) {
                    double In = 0.0;
                    if (doU && obj_v) {
                        const double *fp = f_s + (size_t)pl * nbps;
                        double tprev = __ldg(bps), fprev = fp[0];
                        for (int i = 0; i < nbps - 1; i++) {
                            const double tn = __ldg(bps + i + 1), fn = fp[i + 1];
                            In = In + ((tn - tprev) * (fn + fprev)) / 2;
                            tprev = tn;
                            fprev = fn;
                        }
                    }
                    const double y = (cI_s[pl] + In) + cF_s[pl];
                    if (obj_v && A.f != nullptr) A.f[p] = y;
                    if (want_result(A)) {
                        put_result(A, (size_t)p, 0, obj_v ? y : 0.0);
                        put_result(A, (size_t)p, 1, __longlong_as_double((long long)viol_s[pl]));
                    }
                }
                continue;
            }
            if (!obj_d || A.g == nullptr) continue;
            /* gradient column c: IntegrateFMatrixCols TRAPEZOID over the band,
             * src/integrator.c:44-48 with src/cost.c:118-132; then Vector3Add, src/ntg.c:329 */
            double gI = 0.0, gU = 0.0, gF = 0.0;
            static_for<0, NOUT>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                constexpr int MD = PK::md(j);
                constexpr int IZ = pk_iz<PK>(j);
                const int cl = c - T.iC[j];
                if (cl < 0 || cl >= T.ncoef[j]) return;
                const int order = T.order[j];
                const double *__restrict__ Bn = T.Bn[j];
                const int *__restrict__ offj = T.off[j];
                auto Dval = [&](int i) -> double {
                    const int k = cl - __ldg(offj + i);
                    if (k < 0 || k >= order) return 0.0;
                    const double *b = Bn + ((size_t)i * order + k) * MD;
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < MD; l++)
                        acc = acc + df_s[(size_t)(IZ + l) * G * nbps + (size_t)pl * nbps + i] * __ldg(b + l);
                    return acc;
                };
                if (doU) {
                    const int lo = __ldg(T.col_lo + c), hi = __ldg(T.col_hi + c);
                    const int i0 = lo > 0 ? lo - 1 : 0;
                    const int i1 = hi < nbps - 2 ? hi : nbps - 2;
                    if (i0 <= i1) {
                        double dcur = Dval(i0), tcur = __ldg(bps + i0);
                        for (int i = i0; i <= i1; i++) {
                            const double dnext = Dval(i + 1), tn = __ldg(bps + i + 1);
                            gU = gU + ((tn - tcur) * (dnext + dcur)) / 2;
                            dcur = dnext;
                            tcur = tn;
                        }
                    }
                }
                if (doI && cl < order) { /* CollocConcatMultI: offset 0, block[0], src/colloc.c:243-260 */
                    const double *b = Bn + ((size_t)0 * order + cl) * MD;
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < MD; l++) acc = acc + dfI_s[pl * NZ + IZ + l] * __ldg(b + l);
                    gI = acc;
                }
                if (doF) { /* CollocConcatMultF, src/colloc.c:287-311 */
                    const int k = cl - __ldg(offj + nbps - 1);
                    if (k >= 0 && k < order) {
                        const double *b = Bn + ((size_t)(nbps - 1) * order + k) * MD;
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < MD; l++) acc = acc + dfF_s[pl * NZ + IZ + l] * __ldg(b + l);
                        gF = acc;
                    }
                }
            });
            st_stream(A.g + (size_t)p * nC + c, (gI + gU) + gF);
        }
        __syncthreads();
    }
}

/* ------------------------------ host launcher ------------------------------ */

/* Per-launch host work matters for the small configs (a 4096-problem batch is ~5 us of GPU time):
 * the shared-memory opt-in and the occupancy query are done once per (kernel, device, block, smem). */
struct LaunchPlanCache {
    struct Entry { const void *kern; int dev, block; size_t smem; int nb; };
    Entry e[16];
    int n = 0;
};
/* raise the kernel's dynamic shared-memory limit to everything the device offers beyond the
 * kernel's own static shared memory */
inline cudaError_t raise_smem_limit(const void *kern, int max_optin)
{
    cudaFuncAttributes fa;
    cudaError_t err = cudaFuncGetAttributes(&fa, kern);
    if (err != cudaSuccess) return err;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - (int)fa.sharedSizeBytes);
}

inline int resident_blocks(const void *kern, int block, size_t smem, int max_optin, int *nb_out)
{
    static thread_local LaunchPlanCache cache;
    int dev = 0;
    cudaGetDevice(&dev);
    for (int i = 0; i < cache.n; i++)
        if (cache.e[i].kern == kern && cache.e[i].dev == dev && cache.e[i].block == block && cache.e[i].smem == smem) {
            *nb_out = cache.e[i].nb;
            return 0;
        }
    cudaError_t err;
    if (smem > 48 * 1024) {
        /* the attribute is a LIMIT that a later, smaller value would lower again (a cached larger
         * configuration would then fail to launch): raise it once to the device maximum */
        err = raise_smem_limit(kern, max_optin);
        if (err != cudaSuccess) return (int)err;
    }
    int nb = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, block, smem);
    if (err != cudaSuccess) return (int)err;
    if (nb < 1) nb = 1;
    if (cache.n < 16) cache.e[cache.n++] = {kern, dev, block, smem, nb};
    *nb_out = nb;
    return 0;
}

template <class PK>
int launch_eval(const ntgb_launch *L)
{
    const ntgb_devtab &T = L->tab;
    const int nbps = T.nbps;
    int block, G;
    if (nbps >= 256) {
        const int iters = (nbps + 255) / 256;
        block = (((nbps + iters - 1) / iters) + 31) / 32 * 32;
        G = 1;
    } else {
        block = 256;
        G = block / nbps;
    }
    if (G > L->args.P) {
        G = L->args.P > 0 ? L->args.P : 1;
        int need = ((G * nbps) + 31) / 32 * 32;
        if (need < 64) need = 64;
        if (need < block) block = need;
    }
    SmemLayout lay{G, nbps, pk_nz<PK>()};
    const size_t smem = lay.doubles() * sizeof(double);
    if (smem > (size_t)L->max_smem_optin) return -1000; /* caller reports NTGB_ELIMIT */
    auto kern = ntg_eval_kernel<PK>;
    int nb = 0;
    if (int rc = resident_blocks((const void *)kern, block, smem, L->max_smem_optin, &nb)) return rc;
    const int ntiles = (L->args.P + G - 1) / G;
    int grid = nb * L->sm_count;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) return 0;
    /* cudaLaunchKernelEx reports THIS launch's status; cudaGetLastError() would also return a stale
     * error some other library left behind on this thread */
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)L->args.stream;
    return (int)cudaLaunchKernelEx(&cfg, kern, T, L->args, G);
}

} /* namespace ntgb */

/*
 * NTGB_DEFINE_PACK(name, traits, exact): registers the pack with the core
 * library when the shared object is loaded.
 */
#define NTGB_DEFINE_PACK(NAME, TRAITS, EXACT)                                                   \
    static int ntgb_pack_launch_##NAME(const ntgb_launch *L) { return ntgb::launch_dispatch<TRAITS>(L); } \
    namespace {                                                                                 \
    struct ntgb_pack_registrar_##NAME {                                                         \
        ntgb_pack pk;                                                                           \
        ntgb_pack_registrar_##NAME()                                                            \
        {                                                                                       \
            pk.name = #NAME;                                                                    \
            pk.icf = TRAITS::cb_icf_host(); pk.ucf = TRAITS::cb_ucf_host(); pk.fcf = TRAITS::cb_fcf_host();       \
            pk.nlicf = TRAITS::cb_nlicf_host(); pk.nltcf = TRAITS::cb_nltcf_host(); pk.nlfcf = TRAITS::cb_nlfcf_host(); \
            pk.max_nout = TRAITS::kNout; pk.max_maxderiv = 0;                                    \
            for (int j = 0; j < 8; j++) pk.maxderiv[j] = 0;                                     \
            for (int j = 0; j < TRAITS::kNout; j++) {                                            \
                pk.maxderiv[j] = TRAITS::md(j);                                                 \
                if (TRAITS::md(j) > pk.max_maxderiv) pk.max_maxderiv = TRAITS::md(j);           \
            }                                                                                   \
            pk.max_order = TRAITS::kMaxOrd;                                                      \
            pk.max_nnlic = TRAITS::kNnlic; pk.max_nnltc = TRAITS::kNnltc; pk.max_nnlfc = TRAITS::kNnlfc;  \
            pk.exact = EXACT;                                                                   \
            pk.launch = ntgb_pack_launch_##NAME;                                                \
            pk.abi = NTGB_KERNEL_ABI;                                                           \
            if (ntgb_register_pack(&pk) != 0)                                                   \
                fprintf(stderr, "ntg_b200: pack '%s' not registered: %s\n", #NAME, ntgb_last_error()); \
        }                                                                                       \
    } ntgb_pack_registrar_instance_##NAME;                                                      \
    }

#endif /* NTG_EVAL_KERNEL_CUH_ */
