/*
 * ntg_eval_small.cuh -- K1s: the fused evaluator for shapes whose per-
 * breakpoint basis table fits in registers (sum_j order_j*maxderiv_j <= 64
 * doubles and nbps <= 256: CFG-1..4, both shipped examples).
 *
 * Same math and same reference citations as ntg_eval_kernel.cuh (K1); what
 * changes is the mapping, driven by ncu (profiles/r01_v1_*, r01_v2_*):
 * K1 was limited by L1/LSU wavefronts (69 % of peak at 17 % of the HBM
 * roofline) and instruction issue -- table and index loads repeated for every
 * problem -- and a first persistent version spent half its time at the
 * barrier in front of a 2-warp quadrature phase.
 *
 *   - persistent CTAs, one thread pinned to ONE breakpoint for the whole
 *     launch: its slice of every output's table B_j[bp][.][.] and its offsets
 *     are loaded once into registers; per problem it reads only the
 *     coefficient window (staged one tile ahead into shared memory with
 *     cp.async) and writes results -- the algorithmic traffic.
 *   - a tile is R rounds of G problems (G = 256/nbps); phase A runs all R
 *     rounds back to back, so that phase B -- the trapezoid quadrature of the
 *     cost and of every gradient column, a sequential chain per column in the
 *     reference's ascending-breakpoint order -- has ~256 independent chains
 *     and fills the CTA instead of two warps.
 *   - the chain rule through B is applied in phase A with the register
 *     table; shared memory carries the band D[bp][slot], and phase B walks it
 *     run by run (breakpoints that share a knot interval share a band
 *     position): two shared loads and three flops per term.
 *   - FULL = every output's order equals the pack's bound: no per-k guards.
 */
#ifndef NTG_EVAL_SMALL_CUH_
#define NTG_EVAL_SMALL_CUH_

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ntg_eval_kernel.cuh"
#include "ntg_small_plan.h"

namespace ntgb {

template <class PK>
__host__ __device__ constexpr int pk_tab_base(int j)
{
    int s = 0;
    for (int q = 0; q < j; q++) s += PK::kMaxOrd * PK::md(q);
    return s;
}
template <class PK>
__host__ __device__ constexpr int pk_tab_doubles() { return pk_tab_base<PK>(PK::kNout); }

__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

/* Structural sparsity of a callback's derivative vector (bit iz_j + l set = the callback may write
 * a non-zero there), probed when the pack is built (ntg_b200/build.py::probe_sparsity).  Terms of
 * a structurally zero entry are skipped: the reference adds (+0.0 * B) to an accumulator that
 * starts at +0.0 and can therefore never be -0.0, so for finite B the skipped additions do not
 * change a single bit.  The kernels still CHECK that the entries are +0.0 (sp_clean) and take
 * the dense path otherwise; when the callback really never writes them the compiler has already
 * propagated the zeros and the check folds away. */
constexpr unsigned long long kDense = ~0ull;

template <int NZ>
__device__ __forceinline__ bool sp_clean(const double *df, unsigned long long mask)
{
    bool ok = true;
#pragma unroll
    for (int l = 0; l < NZ; l++)
        if (!((mask >> l) & 1ull)) ok = ok && (__double_as_longlong(df[l]) == 0ll);
    return ok;
}

/* band row from the register-resident table:
 * sink(j, k, value) gets sum_l dz[iz_j + l] * B_j[bp][k][l], l ascending from 0.0, slots in (j,k)
 * order; j and k arrive as integral constants */
template <class PK, bool FULL, bool ONE = false, unsigned long long MASK = kDense, class F>
__device__ __forceinline__ void band_from_regs(const ntgb_devtab &T, const double *Bt, const double *dz, F &&sink)
{
    static_for<0, PK::kNout>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr int MD = PK::md(j);
        constexpr int IZ = pk_iz<PK>(j);
        constexpr int TB = ONE ? 0 : pk_tab_base<PK>(j); /* ONE: every output shares table 0 */
        const int order = FULL ? PK::kMaxOrd : T.order[j];
        static_for<0, PK::kMaxOrd>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            if (FULL || k < order) {
                double acc = 0.0;
                static_for<0, MD>([&](auto lc) {
                    constexpr int l = decltype(lc)::value;
                    if constexpr (((MASK >> (IZ + l)) & 1ull) != 0ull) acc = acc + dz[IZ + l] * Bt[TB + k * MD + l];
                });
                sink(jc, kc, acc);
            }
        });
    });
}

template <class PK, int KIND>
__host__ __device__ constexpr unsigned long long sp_con(int m)
{
    return KIND == 0 ? PK::sp_nlicf(m) : (KIND == 1 ? PK::sp_nltcf(m) : PK::sp_nlfcf(m));
}

__device__ __forceinline__ double nl_bound(const ntgb_devtab &T, bool upper, int idx)
{
    if (T.nl_inline) return upper ? T.nl_ub_v[idx] : T.nl_lb_v[idx];
    return __ldg((upper ? T.nl_ub : T.nl_lb) + idx);
}

/* Trapezoid quadrature of one run of a band row (src/integrator.c:21-24, :44-48).  row[n] is the
 * band value at breakpoint n (16-byte aligned at even n), dt[n] = t[n]-t[n-1].
 *
 * trap_run_exact: terms n0..n1, g = g + (dt[n]*(row[n] + row[n-1]))/2 one after the other in
 * ascending order -- the reference's expression and operation order.  The terms do not depend on
 * the running sum, so they are computed four at a time ahead of the additions. */
__device__ __forceinline__ void trap_run_exact(const double *row, const double *wt, int n0, int n1, double &prev,
                                               double &g)
{
    int n = n0;
    if ((n & 1) && n <= n1) {
        const double d = row[n];
        g = g + (wt[n] * (d + prev)) / 2;
        prev = d;
        n++;
    }
    if (n + 3 <= n1) {
        /* software pipeline: the next block's loads and products are issued before this block's four
         * dependent additions */
        double2 d01 = *reinterpret_cast<const double2 *>(row + n), d23 = *reinterpret_cast<const double2 *>(row + n + 2);
        double2 w01 = *reinterpret_cast<const double2 *>(wt + n), w23 = *reinterpret_cast<const double2 *>(wt + n + 2);
        double t0 = (w01.x * (d01.x + prev)) / 2, t1 = (w01.y * (d01.y + d01.x)) / 2, t2 = (w23.x * (d23.x + d01.y)) / 2,
               t3 = (w23.y * (d23.y + d23.x)) / 2;
        prev = d23.y;
        n += 4;
        for (; n + 3 <= n1; n += 4) {
            d01 = *reinterpret_cast<const double2 *>(row + n);
            d23 = *reinterpret_cast<const double2 *>(row + n + 2);
            w01 = *reinterpret_cast<const double2 *>(wt + n);
            w23 = *reinterpret_cast<const double2 *>(wt + n + 2);
            const double u0 = (w01.x * (d01.x + prev)) / 2, u1 = (w01.y * (d01.y + d01.x)) / 2,
                         u2 = (w23.x * (d23.x + d01.y)) / 2, u3 = (w23.y * (d23.y + d23.x)) / 2;
            prev = d23.y;
            g = g + t0;
            g = g + t1;
            g = g + t2;
            g = g + t3;
            t0 = u0; t1 = u1; t2 = u2; t3 = u3;
        }
        g = g + t0;
        g = g + t1;
        g = g + t2;
        g = g + t3;
    }
    for (; n <= n1; n++) {
        const double d = row[n];
        g = g + (wt[n] * (d + prev)) / 2;
        prev = d;
    }
}

/* fast variant: the same integral as node weights, sum_n Wf[n]*row[n] with Wf[n] = wt[n] + wt[n+1]
 * (the band is zero at both ends of a column's support, so no end corrections), four partial sums */
__device__ __forceinline__ void trap_run_fast(const double *row, const double *Wf, int n0, int n1, double (&g)[4])
{
    int n = n0;
    if ((n & 1) && n <= n1) {
        g[0] = g[0] + Wf[n] * row[n];
        n++;
    }
    for (; n + 7 <= n1; n += 8) { /* eight loads in flight, two multiply-adds per partial sum */
        const double2 d01 = *reinterpret_cast<const double2 *>(row + n), d23 = *reinterpret_cast<const double2 *>(row + n + 2);
        const double2 d45 = *reinterpret_cast<const double2 *>(row + n + 4), d67 = *reinterpret_cast<const double2 *>(row + n + 6);
        const double2 w01 = *reinterpret_cast<const double2 *>(Wf + n), w23 = *reinterpret_cast<const double2 *>(Wf + n + 2);
        const double2 w45 = *reinterpret_cast<const double2 *>(Wf + n + 4), w67 = *reinterpret_cast<const double2 *>(Wf + n + 6);
        g[0] = g[0] + w01.x * d01.x;
        g[1] = g[1] + w01.y * d01.y;
        g[2] = g[2] + w23.x * d23.x;
        g[3] = g[3] + w23.y * d23.y;
        g[0] = g[0] + w45.x * d45.x;
        g[1] = g[1] + w45.y * d45.y;
        g[2] = g[2] + w67.x * d67.x;
        g[3] = g[3] + w67.y * d67.y;
    }
    for (; n + 3 <= n1; n += 4) {
        const double2 d01 = *reinterpret_cast<const double2 *>(row + n), d23 = *reinterpret_cast<const double2 *>(row + n + 2);
        const double2 w01 = *reinterpret_cast<const double2 *>(Wf + n), w23 = *reinterpret_cast<const double2 *>(Wf + n + 2);
        g[0] = g[0] + w01.x * d01.x;
        g[1] = g[1] + w01.y * d01.y;
        g[2] = g[2] + w23.x * d23.x;
        g[3] = g[3] + w23.y * d23.y;
    }
    for (; n <= n1; n++) g[0] = g[0] + Wf[n] * row[n];
}

/* constraint rows of one kind evaluated at this thread's breakpoint:
 * KIND 0 initial (columns from iC_j, src/colloc.c:254), 1 trajectory, 2 final.
 * BAND selects the layout at compile time; SPARSE uses the pack's probed sparsity masks. */
template <class PK, bool FULL, int NCON, int KIND, bool ONE, bool BAND, bool SPARSE, int MOFF = 0>
__device__ __forceinline__ void emit_rows_layout(const ntgb_devtab &T, const ntgb_eval_args &A, const double *Bt,
                                                 const int *offj, int p, int bp,
                                                 const double (&dfc)[NCON][pk_nz<PK>()], int row_base)
{
    const int nbps = T.nbps, S = T.S;
    if constexpr (BAND) {
        /* J band [p][row][slot][.]: rows of a trajectory constraint are breakpoint-fastest, so a
         * thread's stores are S*NCON strided ones.  With every order at the pack's bound the slot
         * of (j, k) is a compile-time constant and each address is one multiply-add from one base. */
        char *Jp = reinterpret_cast<char *>(A.J + ((size_t)p * T.ncnln + row_base) * S + (KIND == 1 ? bp : 0));
        const unsigned stride8 = (KIND == 1 ? (unsigned)nbps : 1u) * 8u;
        static_for<0, NCON>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            constexpr unsigned long long MASK = SPARSE ? sp_con<PK, KIND>(MOFF + m) : kDense;
            if constexpr (FULL) {
                constexpr int SF = PK::kNout * PK::kMaxOrd;
                band_from_regs<PK, FULL, ONE, MASK>(T, Bt, dfc[m], [&](auto jc, auto kc, double v) {
                    constexpr int slot = m * SF + decltype(jc)::value * PK::kMaxOrd + decltype(kc)::value;
                    st_stream(reinterpret_cast<double *>(Jp + (size_t)stride8 * (unsigned)slot), v);
                });
            } else {
                char *ptr = Jp + (size_t)stride8 * (unsigned)(m * S);
                band_from_regs<PK, FULL, ONE, MASK>(T, Bt, dfc[m], [&](auto, auto, double v) {
                    st_stream(reinterpret_cast<double *>(ptr), v);
                    ptr += stride8;
                });
            }
        });
    } else {
        /* NPSOL's dense column-major J[col * ncnln + row] (src/ntg.c:217-220): consecutive breakpoints
         * are consecutive rows of one column inside a knot interval, so a warp's stores coalesce.  One
         * base per output (its first column at this breakpoint, this thread's row); a band entry
         * (m, j, k) is then k columns and m row blocks away: one multiply-add per address. */
        const unsigned ld = (unsigned)T.ncnln;
        double *Jj[PK::kNout];
        static_for<0, PK::kNout>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            Jj[j] = A.J + ((size_t)p * T.nC + (unsigned)(T.iC[j] + (KIND == 0 ? 0 : offj[j]))) * ld +
                    (unsigned)(KIND == 1 ? row_base + bp : row_base);
        });
        const unsigned rstep = KIND == 1 ? (unsigned)nbps : 1u;
        static_for<0, NCON>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            constexpr unsigned long long MASK = SPARSE ? sp_con<PK, KIND>(MOFF + m) : kDense;
            band_from_regs<PK, FULL, ONE, MASK>(T, Bt, dfc[m], [&](auto jc, auto kc, double v) {
                constexpr int j = decltype(jc)::value;
                constexpr unsigned k = (unsigned)decltype(kc)::value;
                st_stream(Jj[j] + (k * ld + (unsigned)m * rstep), v);
            });
        });
    }
}

/* JL: the Jacobian layout when the kernel knows it at compile time (1 band, 2 dense), else 0 */
template <class PK, bool FULL, int NCON, int KIND, bool ONE = false, int JL = 0>
__device__ __forceinline__ void emit_rows_regs(const ntgb_devtab &T, const ntgb_eval_args &A, const double *Bt,
                                               const int *offj, int p, int bp, const double (&dfc)[NCON][pk_nz<PK>()],
                                               int row_base)
{
    bool clean = true;
    static_for<0, NCON>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        clean = clean && sp_clean<pk_nz<PK>()>(dfc[m], sp_con<PK, KIND>(m));
    });
    const bool band = JL == 1 || (JL == 0 && A.jac_layout == NTGB_JAC_BAND);
    if (band) {
        if (clean) emit_rows_layout<PK, FULL, NCON, KIND, ONE, true, true>(T, A, Bt, offj, p, bp, dfc, row_base);
        else emit_rows_layout<PK, FULL, NCON, KIND, ONE, true, false>(T, A, Bt, offj, p, bp, dfc, row_base);
    } else {
        if (clean) emit_rows_layout<PK, FULL, NCON, KIND, ONE, false, true>(T, A, Bt, offj, p, bp, dfc, row_base);
        else emit_rows_layout<PK, FULL, NCON, KIND, ONE, false, false>(T, A, Bt, offj, p, bp, dfc, row_base);
    }
}

/* chain rule of a COST derivative vector through the table, with the pack's sparsity mask */
template <class PK, bool FULL, unsigned long long MASK, class F>
__device__ __forceinline__ void cost_band(const ntgb_devtab &T, const double *Bt, const double *df, F &&sink)
{
    if (sp_clean<pk_nz<PK>()>(df, MASK)) band_from_regs<PK, FULL, false, MASK>(T, Bt, df, sink);
    else band_from_regs<PK, FULL, false, kDense>(T, Bt, df, sink);
}

/* Fused multi-GPU gather: thread q stores the (objective, violation) pair of problem p0 + q of a
 * finished tile -- collected in shared memory by phase B -- into every destination table (the local
 * one and every rank's gathered table), one 16-byte store each.  The evaluator is bound by
 * instruction issue (ncu: every instruction per warp and tile is worth ~0.07 % of the launch), so
 * this is a call that a few threads of one warp make once per tile, NOT inlined: its addresses and
 * loop stay out of the main loop's register allocation.  (Tried and dropped: pushing from the local
 * global table instead of shared memory -- the load queues behind the SM's store stream and holds
 * the tile's barrier: 129 us per step at 2 GPUs instead of 125.) */
__device__ __noinline__ void push_pairs(const double *res_s, double2 *const *dst_s, int ndst, int pq)
{
    const double2 v = *reinterpret_cast<const double2 *>(res_s + 2 * threadIdx.x);
    for (int r = 0; r < ndst; r++) dst_s[r][pq] = v;
}

/* HOT = the solver's steady state, known at compile time: funobj mode 2 + funcon mode 2, Jacobian
 * in band layout -- or, DENSE, in NPSOL's dense column-major layout (what ntg()'s funcon hands NPSOL) --
 * in band layout, f / g / c / J all requested, Z not requested. */
template <class PK, bool FULL, bool HOT = false, bool PEERS = true, int BLOCK = 256, bool DENSE = false>
__global__ void __launch_bounds__(BLOCK, 512 / BLOCK)
ntg_eval_small_kernel(const ntgb_devtab T, const ntgb_eval_args A, int G, int R, int segtot, int flags)
{
    constexpr int NOUT = PK::kNout;
    constexpr int NZ = pk_nz<PK>();
    constexpr int NB = pk_tab_doubles<PK>();
    extern __shared__ double smem[];
    /* flags: bit 0 programmatic dependent launch, bit 1 rotate the peer-store order, bits 4-9 tiles per
     * CTA of an EVEN split (0: tiles of G*R problems dealt round-robin), bits 10.. rows of a tile's
     * buffers in an even split (>= the largest tile; G*R otherwise) */
    const int GR = (flags >> 10) ? (flags >> 10) : G * R;
    const SmallSmem L{GR, T.nbps, T.S, T.nout, T.nC, segtot};
    const int nbps = T.nbps, pitch = L.pitch(), nC = T.nC, P = A.P, S = T.S;
    double *D_s = smem + L.D_off();
    double *f_s = smem + L.f_off();
    double *DI_s = smem + L.DI_off();
    double *DF_s = smem + L.DF_off();
    double *cI_s = smem + L.cI_off();
    double *cF_s = smem + L.cF_off();
    double *viol_s = smem + L.viol_off();
    double *wt_s = smem + L.dt_off();  /* exact variant: dt[n] = t[n]-t[n-1] (dt[0] = 0); the fast one only uses Wf */
    double *Wf_s = wt_s + pitch + 2;   /* node weights wt[n] + wt[n+1] */
    double *C_s = smem + L.C_off();
    int *segstart_s = reinterpret_cast<int *>(smem + L.seg_off());
    int *segoff_s = segstart_s + segtot;
    int *costseg_s = segoff_s + segtot; /* the cost as a chain: one run {0, nbps}, offset 0 */
    int *par_s = costseg_s + 4;         /* [nC+1][9] chain description per column, see below */
    int *cols_s = par_s + (nC + 1) * 9; /* [nC+1] the schedule's column list */
    /* fused multi-GPU gather (HOT && PEERS): a tile's (objective, violation) pairs are collected in
     * shared memory and pushed ONCE, as one 16-byte store per problem and table from consecutive
     * lanes, while the next tile's phase A runs -- not as scattered 8-byte stores from inside the
     * latency-bound quadrature phase */
    constexpr bool PUSH = HOT && PEERS;
    double *res_s = smem + L.res_off();
    double2 **dst_s = reinterpret_cast<double2 **>(
        reinterpret_cast<char *>(smem + L.seg_off()) + ((2 * (size_t)segtot + 4 + (size_t)(nC + 1) * 10 + 1) & ~(size_t)1) * 4);
    const int ndst = PUSH ? A.npeers + (A.result != nullptr ? 1 : 0) : 0;
    if constexpr (PUSH) { /* destination tables: [local,] rank 0, rank 1, ... offset to this rank's first row */
        const int t = (int)threadIdx.x, loc = A.result != nullptr ? 1 : 0;
        if (t == 0 && loc) dst_s[0] = reinterpret_cast<double2 *>(A.result);
        /* every CTA starts its round over the ranks at a different one (flags bit 1): 8 ranks x 296 CTAs
         * that all store to rank 0 first, then rank 1, ... would take turns on one NVLink port at a time */
        const int rot = ((flags & 2) && A.npeers > 1) ? (int)(blockIdx.x % (unsigned)A.npeers) : 0;
        if (t < A.npeers) {
            const int src = t + rot < A.npeers ? t + rot : t + rot - A.npeers;
            dst_s[loc + t] = reinterpret_cast<double2 *>(A.peer_result[src]) + A.peer_row0;
        }
    }

    const int mode_obj = HOT ? 2 : A.mode_obj, mode_con = HOT ? 2 : A.mode_con;
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
    const bool wantJ = HOT || (con_d && A.J != nullptr && A.jac_layout != NTGB_JAC_NONE);
    const bool wantZ = !HOT && A.Z != nullptr;
    const bool want_c = HOT || A.c != nullptr;

    /* tiles: GR problems each, dealt round-robin to the CTAs -- or, for a batch of a few tiles per CTA
     * (launch_eval_small), an EVEN split: CTA b owns the contiguous problems [b*P/grid, (b+1)*P/grid)
     * and walks them in tiles of GR (the launcher sizes GR so that every CTA needs the same number of
     * tiles).  Whole tiles of GR dealt round-robin leave a third of the CTA slots empty at CFG-3 and
     * put three full rounds on the others; 8192 lane changes are 683 tiles of 12, three for some CTAs
     * and two for the rest.  Either way a CTA's tiles are p_first, p_first + pstride, ... below pend:
     * the tile loop costs one add and one minimum per tile. */
    const bool even = ((flags >> 4) & 63) != 0;
    int p_first, pstride, pend;
    if (even) {
        const int q = P / (int)gridDim.x, rem = P - q * (int)gridDim.x, b = (int)blockIdx.x;
        p_first = b * q + (b < rem ? b : rem);
        pend = p_first + q + (b < rem ? 1 : 0);
        pstride = GR;
    } else {
        p_first = (int)blockIdx.x * GR;
        pend = P;
        pstride = (int)gridDim.x * GR;
    }
    const int tileC = GR * nC; /* doubles of coefficients a tile's buffer holds (contiguous in global memory) */
    auto stage_C = [&](int q0, int nq, int buf) { /* problems [q0, q0 + nq) */
        const int cnt = nq * nC;
        const double *src = A.C + (long long)q0 * nC;
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) cp_async8(C_s + (size_t)buf * tileC + e, src + e);
        cp_async_commit();
    };
    /* tiles of this CTA: p_first + it * pstride for it < ntl, GR problems each but possibly the last */
    int ntl = pend > p_first ? (pend - p_first + pstride - 1) / pstride : 0;
    /* (the same value in every lane; read back from lane 0, the loop bound sits in a register of its own.
     * This kernel's run time moves by +-5 % with the instruction schedule ptxas picks for phase A, and that
     * moves with the spelling of this loop: measured with tools/variants_perf.sh on the steady-state
     * instantiations with and without the peer stores -- 125.1 / 122.3 us with this line, 130 / 123 with
     * p0 and np carried across iterations, 124.5 / 130.5 with the tile index as the loop variable) */
    ntl = __shfl_sync(0xffffffffu, ntl, 0);
    int buf = 0;
    /* flags bit 0: launched with programmatic stream serialization -- this grid may start while the grid in
     * front of it in the stream is still running.  Everything up to griddepcontrol.wait only reads
     * the batch-shared tables (written once at create time), so the whole prologue overlaps the
     * previous launch; coefficients are read and results written after the wait.  The grid behind
     * this one may start its own prologue as soon as every CTA of this grid is resident. */
    if (flags & 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    else if (ntl > 0) stage_C(p_first, pend - p_first < GR ? pend - p_first : GR, 0); /* no PDL: requested first, ahead of the table loads */

    /* ---- once per CTA: dt, offset runs, accumulators; once per thread: its table slice ----
     * A thread's table slice is requested FIRST (it is only consumed in phase A), and the three
     * table-building loops below start at different warps of the CTA (tw / ts / tc): their global
     * loads are independent of each other, so the prologue costs about one load round trip instead
     * of one per loop (measured with device time stamps: 1.9 - 2.5 us before). */
    const int pl = threadIdx.x / nbps;
    const int bp = threadIdx.x - pl * nbps;
    const bool active = pl < G;
    const int cls = (bp == 0 ? 1 : 0) | (bp == nbps - 1 ? 2 : 0);
    double Bt[NB > 0 ? NB : 1];
    int offj[NOUT];
    static_for<0, NOUT>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr int MD = PK::md(j);
        constexpr int TB = pk_tab_base<PK>(j);
        const int order = T.order[j];
        offj[j] = active ? __ldg(T.off[j] + bp) : 0;
#pragma unroll
        for (int k = 0; k < PK::kMaxOrd; k++)
#pragma unroll
            for (int d = 0; d < MD; d++)
                Bt[TB + k * MD + d] = (active && k < order) ? __ldg(T.Bt[j] + (size_t)(k * MD + d) * nbps + bp) : 0.0;
    });
    const int bdim = (int)blockDim.x;
    auto rot = [&](int off) { const int t = (int)threadIdx.x - off; return t < 0 ? t + bdim : t; };
    const int tw = (int)threadIdx.x;                                   /* weights: from warp 0      */
    const int ts = rot(bdim >= 128 ? 96 : 0);                          /* run tables: from warp 3   */
    const int tc = rot(bdim >= 256 ? 128 : (bdim >= 128 ? 64 : 0));    /* chain table: from warp 4  */
    /* steady state: the tile-invariant tables below were laid out once at create time in exactly the
     * shape they have in shared memory (T.img_w, T.img_i; ntg_core.cu) -- the CTA copies them instead
     * of building them: the building code is ~700 of the ~1100 instructions this prologue issues per
     * warp, which a batch of one tile per CTA pays on its critical path */
    bool use_img = false;
    if constexpr (HOT)
        use_img = T.img_i != nullptr && (PK::cb_ucf != nullptr || T.nucf == 0) && (PK::cb_icf != nullptr || T.nicf == 0) &&
                  (PK::cb_fcf != nullptr || T.nfcf == 0);
    if (use_img) {
        for (int n = tw; n < 2 * (pitch + 2); n += bdim) wt_s[n] = __ldg(T.img_w + n);
        const int par0 = 2 * segtot + 4, nimg = par0 + (nC + 1) * 9;
        for (int i = ts; i < nimg; i += bdim) {
            int v = __ldg(T.img_i + i);
            const int r = i - par0;
            if (r >= 0 && r % 9 == 4) v *= GR; /* a column's base inside D_s: (slot * pitch) * GR */
            segstart_s[i] = v;
        }
    } else {
        for (int n = tw; n < pitch + 2; n += bdim) {
            const double lo = (n >= 1 && n < nbps) ? (__ldg(T.bps + n) - __ldg(T.bps + n - 1)) * 0.5 : 0.0;
            const double hi = (n + 1 < nbps) ? (__ldg(T.bps + n + 1) - __ldg(T.bps + n)) * 0.5 : 0.0;
            wt_s[n] = (n >= 1 && n < nbps) ? __ldg(T.bps + n) - __ldg(T.bps + n - 1) : 0.0;
            Wf_s[n] = lo + hi;
        }
        {
            int base = 0;
            for (int j = 0; j < T.nout; j++) {
                for (int i = ts; i <= T.nseg[j]; i += bdim) {
                    segstart_s[base + i] = __ldg(T.seg_start[j] + i);
                    segoff_s[base + i] = __ldg(T.seg_off[j] + i);
                }
                base += T.nseg[j] + 1;
            }
        }
    }
    for (int q = threadIdx.x; q < GR; q += blockDim.x) {
        cI_s[q] = 0.0;
        cF_s[q] = 0.0;
    }

    /* phase-B mapping (tile-invariant): slot and problem lane of this thread, its column list */
    const bool want_g = HOT || (obj_d && A.g != nullptr);
    const int lanesB = GR < (int)blockDim.x ? GR : (int)blockDim.x;
    const int slotB = threadIdx.x / lanesB, laneB = threadIdx.x - slotB * lanesB;
    const int NSB = (int)blockDim.x / lanesB;
    const bool use_sched = T.sched != nullptr && blockDim.x == NTGB_SCHED_BLOCK;
    const int NSS = NSB < T.sched_maxns ? NSB : T.sched_maxns; /* the schedule for this many slots */
    const int *sched_st = use_sched ? T.sched + (size_t)(NSS - 1) * (NTGB_SCHED_BLOCK + 1 + nC + 1) : nullptr;
    const int *sched_cols = use_sched ? sched_st + NTGB_SCHED_BLOCK + 1 : nullptr;
    int it0, it1, istep;
    if (use_sched) {
        const bool has = slotB < NSS;
        it0 = has ? __ldg(sched_st + slotB) : 0;
        it1 = has ? __ldg(sched_st + slotB + 1) : 0;
        istep = 1;
    } else { /* any other launch geometry: columns dealt round-robin over the slots */
        it0 = slotB < NSB ? slotB : nC + 1;
        it1 = nC + 1;
        istep = NSB;
    }
    if (!use_img) {
        if (threadIdx.x == 0) {
            costseg_s[0] = 0;
            costseg_s[1] = nbps;
            costseg_s[2] = 0;
        }
        /* chain description per column, built once (everything in it is tile-invariant):
         * [0] first breakpoint i0, [1] last term nend (no chain when i0 >= nend), [2] run holding i0,
         * [3]/[8] where the output's run starts / run offsets sit in segstart_s, [4] the column's base
         * inside D_s (doubles), [5] band width, [6] local column, [7] (slot in DI)+1 | ((slot in DF)+1)<<16 */
        for (int c = tc; c <= nC; c += bdim) {
            int *pp = par_s + c * 9;
            if (c == nC) {
                pp[0] = 0; pp[1] = (doU && obj_v) ? nbps - 1 : 0; pp[2] = 0; pp[3] = 2 * segtot; pp[8] = 2 * segtot + 2;
                pp[4] = S * GR * pitch; /* f_s follows D_s */
                pp[5] = 1; pp[6] = 0; pp[7] = 0;
            } else {
                int sb = 0;
                pp[0] = 0; pp[1] = 0; pp[2] = 0; pp[3] = 0; pp[4] = 0; pp[5] = 0; pp[6] = 0; pp[7] = 0; pp[8] = 0;
                for (int j = 0; j < T.nout; j++) {
                    const int clj = c - T.iC[j];
                    if (clj >= 0 && clj < T.ncoef[j]) {
                        const int ord = T.order[j];
                        if (doU) {
                            const int lo = __ldg(T.col_lo + c), hi = __ldg(T.col_hi + c);
                            pp[0] = lo > 0 ? lo - 1 : 0;
                            pp[1] = (hi < nbps - 2 ? hi : nbps - 2) + 1;
                            pp[2] = __ldg(T.col_seg0 + c);
                        }
                        pp[3] = sb;
                        pp[8] = segtot + sb;
                        pp[4] = T.jk0[j] * GR * pitch;
                        pp[5] = ord;
                        pp[6] = clj;
                        int iDI = 0, iDF = 0;
                        if (doI && clj < ord) iDI = T.jk0[j] + clj + 1; /* offset 0, src/colloc.c:254 */
                        if (doF) {
                            const int k = clj - __ldg(T.seg_off[j] + T.nseg[j] - 1);
                            if (k >= 0 && k < ord) iDF = T.jk0[j] + k + 1;
                        }
                        pp[7] = iDI | (iDF << 16);
                    }
                    sb += T.nseg[j] + 1;
                }
            }
        }
    }
    for (int c = tc; c <= nC; c += bdim) cols_s[c] = use_sched ? __ldg(sched_cols + c) : c;
    if (flags & 1) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (ntl > 0) stage_C(p_first, pend - p_first < GR ? pend - p_first : GR, 0);
    }

    for (int it = 0; it < ntl; it++, buf ^= 1) {
        const int p0 = p_first + it * pstride;
        const int np = pend - p0 < GR ? pend - p0 : GR; /* problems of this tile */
        cp_async_wait_all();
        __syncthreads(); /* coefficients of this tile landed; phase B of the previous tile is done */
        if (it + 1 < ntl) stage_C(p0 + pstride, pend - p0 - pstride < GR ? pend - p0 - pstride : GR, buf ^ 1);
        if constexpr (PUSH) { /* the previous tile's pairs (phase B of that tile ended at the barrier above);
                               * a tile that has a successor is a full one */
            if (it > 0 && (int)threadIdx.x < GR) push_pairs(res_s, dst_s, ndst, p0 - pstride + (int)threadIdx.x);
        }

        /* ---------------- phase A: this thread's breakpoint, R problems ---------------- */
        if (active) {
            for (int r = 0; r < R; r++) {
                const int plr = r * G + pl;
                const int p = p0 + plr;
                if (plr >= np) break;
                const double *Cp = C_s + (size_t)buf * tileC + (size_t)plr * nC;
                double z[NZ > 0 ? NZ : 1];
                double *zp[NOUT];
                /* Zvalue, src/colloc.c:318-326 -- k ascending from 0.0 */
                static_for<0, NOUT>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    constexpr int MD = PK::md(j);
                    constexpr int IZ = pk_iz<PK>(j);
                    constexpr int TB = pk_tab_base<PK>(j);
                    const int order = FULL ? PK::kMaxOrd : T.order[j];
                    const double *Cw = Cp + T.iC[j] + offj[j];
                    const unsigned mask = T.avmask[cls][j];
                    double acc[MD];
#pragma unroll
                    for (int d = 0; d < MD; d++) acc[d] = 0.0;
#pragma unroll
                    for (int k = 0; k < PK::kMaxOrd; k++) {
                        if (FULL || k < order) {
                            const double ck = Cw[k];
#pragma unroll
                            for (int d = 0; d < MD; d++) acc[d] = acc[d] + Bt[TB + k * MD + d] * ck;
                        }
                    }
#pragma unroll
                    for (int d = 0; d < MD; d++) z[IZ + d] = ((mask >> d) & 1u) ? acc[d] : 0.0;
                    zp[j] = &z[IZ];
                    if (wantZ) {
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
                            double *cp = A.c + (size_t)p * T.ncnln + T.nnlic + bp;
#pragma unroll
                            for (int m = 0; m < PK::kNnltc; m++) {
                                if (want_c) st_stream(cp + (size_t)m * nbps, cv[m]);
                                viol = fmax(viol, row_violation(cv[m], nl_bound(T, false, T.nnlic + m),
                                                                nl_bound(T, true, T.nnlic + m)));
                            }
                        }
                        if (wantJ) emit_rows_regs<PK, FULL, PK::kNnltc, 1, false, (HOT ? (DENSE ? 2 : 1) : 0)>(T, A, Bt, offj, p, bp, dfc, T.nnlic);
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
                                if (want_c) st_stream(A.c + (size_t)p * T.ncnln + m, cv[m]);
                                viol = fmax(viol, row_violation(cv[m], nl_bound(T, false, m), nl_bound(T, true, m)));
                            }
                        }
                        if (wantJ) emit_rows_regs<PK, FULL, PK::kNnlic, 0, false, (HOT ? (DENSE ? 2 : 1) : 0)>(T, A, Bt, offj, p, bp, dfc, 0);
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
                                if (want_c) st_stream(A.c + (size_t)p * T.ncnln + rb + m, cv[m]);
                                viol = fmax(viol, row_violation(cv[m], nl_bound(T, false, T.nnlic + T.nnltc + m),
                                                                nl_bound(T, true, T.nnlic + T.nnltc + m)));
                            }
                        }
                        if (wantJ) emit_rows_regs<PK, FULL, PK::kNnlfc, 2, false, (HOT ? (DENSE ? 2 : 1) : 0)>(T, A, Bt, offj, p, bp, dfc, rb);
                    }
                }
                /* per-breakpoint violation; phase B takes the maximum over the breakpoints (a maximum
                 * does not depend on the order it is taken in) */
                if (con_v) viol_s[plr * pitch + bp] = viol;

                /* unintegrated (trajectory) cost, src/cost.c:99-132: the band of dIdC (chain rule
                 * through B) goes to shared memory for the quadrature */
                if constexpr (PK::cb_ucf != nullptr) {
                    if (doU) {
                        double fv = 0.0;
                        double df[NZ > 0 ? NZ : 1];
#pragma unroll
                        for (int l = 0; l < NZ; l++) df[l] = 0.0;
                        int mode = mode_obj, i = bp;
                        PK::cb_ucf(&mode, &nstate, &i, &fv, df, zp);
                        note_abort(A, mode);
                        f_s[plr * pitch + bp] = fv;
                        if (obj_d) {
                            double *Dp = D_s + (size_t)plr * pitch + bp;
                            const int slot_pitch = GR * pitch;
                            cost_band<PK, FULL, PK::sp_ucf()>(T, Bt, df, [&](auto, auto, double v) {
                                *Dp = v;
                                Dp += slot_pitch;
                            });
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
                        cI_s[plr] = fv;
                        if (obj_d) {
                            double *Dp = DI_s + plr * S;
                            cost_band<PK, FULL, PK::sp_icf()>(T, Bt, df, [&](auto, auto, double v) { *Dp++ = v; });
                        }
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
                        cF_s[plr] = fv;
                        if (obj_d) {
                            double *Dp = DF_s + plr * S;
                            cost_band<PK, FULL, PK::sp_fcf()>(T, Bt, df, [&](auto, auto, double v) { *Dp++ = v; });
                        }
                    }
                }
            }
        }
        __syncthreads();

        /* ------- phase B: one trapezoid chain per (problem, column); the scalar cost is column nC.
         * Thread = (slot, problem lane); a slot walks the columns the schedule gave it, so that all
         * slots carry the same number of terms and a warp's lanes run the same columns. ------- */
        for (int plr = laneB; plr < np; plr += lanesB) {
            const int pb = p0 + plr;
            for (int idx = it0; idx < it1; idx += istep) {
                const int c = use_sched ? cols_s[idx] : idx;
                if (c < nC && !want_g) continue;
                const int *pp = par_s + c * 9;
                const int i0 = pp[0], nend = pp[1], s0 = pp[2], cl = pp[6], ipk = pp[7];
                const int *ss = segstart_s + pp[3], *so = segstart_s + pp[8];
                const double *Dj = D_s + pp[4] + (size_t)plr * pitch; /* row of band slot k: Dj + k*GR*pitch */
                const unsigned order = (unsigned)pp[5];
                const bool run_chain = i0 < nend;
                const double gI = (ipk & 0xffff) ? DI_s[plr * S + (ipk & 0xffff) - 1] : 0.0;
                const double gF = (ipk >> 16) ? DF_s[plr * S + (ipk >> 16) - 1] : 0.0;
                /* IntegrateVector / IntegrateFMatrixCols TRAPEZOID (src/integrator.c:21-24, :44-48 on
                 * the matrix of src/cost.c:118-132): ascending breakpoint, run by run */
                double gU = 0.0;
                if (run_chain) {
                    const int rowp = GR * pitch;
                    int s = s0;
                    int k = cl - so[s];
                    if constexpr (PK::kExact) {
                        double dcur = ((unsigned)k < order) ? Dj[(size_t)k * rowp + i0] : 0.0;
                        int n = i0 + 1;
                        while (n <= nend) {
                            int snext = ss[s + 1];
                            if (n >= snext) {
                                s++;
                                k = cl - so[s];
                                snext = ss[s + 1];
                            }
                            const int segend = snext - 1 < nend ? snext - 1 : nend;
                            if ((unsigned)k < order) {
                                trap_run_exact(Dj + (size_t)k * rowp, wt_s, n, segend, dcur, gU);
                            } else { /* leaving the band: one term against an exact zero, the rest are zeros */
                                gU = gU + (wt_s[n] * (0.0 + dcur)) / 2;
                                dcur = 0.0;
                            }
                            n = segend + 1;
                        }
                    } else {
                        double acc[4] = {0.0, 0.0, 0.0, 0.0};
                        int n = i0;
                        while (n <= nend) {
                            int snext = ss[s + 1];
                            if (n >= snext) {
                                s++;
                                k = cl - so[s];
                                snext = ss[s + 1];
                            }
                            const int segend = snext - 1 < nend ? snext - 1 : nend;
                            if ((unsigned)k < order) trap_run_fast(Dj + (size_t)k * rowp, Wf_s, n, segend, acc);
                            n = segend + 1;
                        }
                        gU = (acc[0] + acc[1]) + (acc[2] + acc[3]);
                    }
                }
                if (c < nC) {
                    st_stream(A.g + (size_t)pb * nC + c, (gI + gU) + gF); /* Vector3Add, src/ntg.c:329 */
                } else {
                    const double y = (cI_s[plr] + gU) + cF_s[plr]; /* y = I + In + F, src/ntg.c:303,328 */
                    if (HOT || (obj_v && A.f != nullptr)) A.f[pb] = y;
                    if constexpr (PUSH) {
                        res_s[2 * plr] = y;
                    } else if constexpr (!PEERS) {
                        if (A.result != nullptr) {
                            A.result[2 * (size_t)pb] = obj_v ? y : 0.0;
                            if (!con_v) A.result[2 * (size_t)pb + 1] = 0.0;
                        }
                    } else if (want_result(A)) {
                        put_result(A, (size_t)pb, 0, obj_v ? y : 0.0);
                        if (!con_v) put_result(A, (size_t)pb, 1, 0.0);
                    }
                }
            }
        }
        /* maximum constraint violation per problem: eight lanes per problem, then three shuffles.
         * Handed out from the END of the block: with fewer chains than threads these are warps that
         * have no chain to walk, so the two passes run side by side. */
        if (con_v && (PUSH || (PEERS ? want_result(A) : A.result != nullptr))) {
            /* 8 lanes per problem when they fit beside the chains, else 4 (never fewer: the loop
             * below covers any size) */
            const int chain_threads = (nC + 1) * GR;
            const int LV = (chain_threads + GR * 8 <= (int)blockDim.x) ? 8 : 4;
            const int nv = np * LV;
            const int tv = (int)blockDim.x - 1 - (int)threadIdx.x;
            for (int base = 0; base < nv; base += blockDim.x) {
                const int q = base + tv;
                const int plr = q / LV, part = q - plr * LV;
                double vm = 0.0, vn = 0.0;
                if (q < nv) {
                    const double *vp = viol_s + (size_t)plr * pitch;
                    int i = part;
                    for (; i + LV < nbps; i += 2 * LV) {
                        vm = fmax(vm, vp[i]);
                        vn = fmax(vn, vp[i + LV]);
                    }
                    if (i < nbps) vm = fmax(vm, vp[i]);
                }
                vm = fmax(vm, vn);
                vm = fmax(vm, __shfl_xor_sync(0xffffffffu, vm, 1));
                vm = fmax(vm, __shfl_xor_sync(0xffffffffu, vm, 2));
                if (LV == 8) vm = fmax(vm, __shfl_xor_sync(0xffffffffu, vm, 4));
                if (q < nv && part == 0 && plr < np) {
                    if constexpr (PUSH) res_s[2 * plr + 1] = vm;
                    else if constexpr (!PEERS) A.result[2 * (size_t)(p0 + plr) + 1] = vm;
                    else put_result(A, (size_t)(p0 + plr), 1, vm);
                }
            }
        }
        /* the barrier at the top of the next iteration separates this phase B from the next phase A */
    }
    cp_async_wait_all();
    if constexpr (PUSH) { /* the last tile's pairs */
        __syncthreads();
        const int pl0 = p_first + (ntl - 1) * pstride;
        const int npl = pend - pl0 < GR ? pend - pl0 : GR;
        if (ntl > 0 && (int)threadIdx.x < npl) push_pairs(res_s, dst_s, ndst, pl0 + (int)threadIdx.x);
    }
}

/* does this problem fit the register-table kernel? */
template <class PK>
__host__ inline bool small_shape_ok(const ntgb_devtab &T)
{
    /* up to 256 breakpoints: CTAs of 256 threads, two per SM; up to 512: one CTA of 512 threads per SM */
    return pk_tab_doubles<PK>() <= 64 && T.nbps <= 512 && T.band_tile == T.nbps; /* band rows of nbps values */
}

template <class PK>
int launch_eval_small(const ntgb_launch *L)
{
    const ntgb_devtab &T = L->tab;
    const int nbps = T.nbps, P = L->args.P;
    const bool wide = nbps > 256;
    int segtot = 0;
    bool full = true;
    for (int j = 0; j < T.nout; j++) {
        segtot += T.nseg[j] + 1;
        full = full && T.order[j] == PK::kMaxOrd;
    }
    /* the geometry (rounds per tile, tile rows, whole tiles or an even split): ntg_small_plan.h */
    static const SmallPlanKnobs knobs = {getenv("NTG_B200_ROUNDS") ? atoi(getenv("NTG_B200_ROUNDS")) : 0,
                                         getenv("NTG_B200_SMEMCAP") ? atoi(getenv("NTG_B200_SMEMCAP")) : 0,
                                         getenv("NTG_B200_NO_EVEN_SPLIT") != nullptr,
                                         getenv("NTG_B200_EVEN_MAXTILES") ? atoi(getenv("NTG_B200_EVEN_MAXTILES")) : 8};
    const SmallPlan plan = plan_small_launch(P, nbps, T.S, T.nout, T.nC, segtot, L->sm_count, knobs);
    const int block = plan.block, G = plan.G, slots = plan.slots, R_tiles = plan.R_tiles, rows_tiles = plan.rows_tiles;
    const int even_grid = plan.even_grid;
    int R = plan.R, rows = plan.rows, ktiles = plan.ktiles;
    auto smem_rows = [&](int r) { return SmallSmem{r, nbps, T.S, T.nout, T.nC, segtot}.bytes(); };
    size_t smem = smem_rows(rows);
    if (smem > (size_t)L->max_smem_optin) return -1001;
    const ntgb_eval_args &a = L->args;
    const bool steady = a.mode_obj == 2 && a.mode_con == 2 && a.J != nullptr && a.f != nullptr && a.g != nullptr &&
                        a.c != nullptr && a.Z == nullptr && T.ncnln > 0;
    const bool hot = steady && a.jac_layout == NTGB_JAC_BAND &&
                     ((uintptr_t)a.result & 15u) == 0; /* the peer-store variant writes (objective, violation) as one 16-byte pair */
    /* the same steady state with NPSOL's dense column-major Jacobian (single GPU) */
    const bool hot_dense = steady && a.jac_layout == NTGB_JAC_DENSE && a.npeers == 0;
    /* the steady-state kernel exists with and without the push epilogue of the fused multi-GPU gather
     * (a tile's result pairs collected in shared memory and stored once per tile).  The push variant
     * is the faster one on a single GPU too, where the only destination is the local table (cfg4:
     * 122.3 against 125.1 us), when a CTA walks several tiles; for one tile per CTA its last push sits
     * on the critical path (cfg2 4.6 against 4.2 us).  NTG_B200_NO_PUSH_KERNEL=1 selects the other
     * for A/B */
    static const bool push_default = getenv("NTG_B200_NO_PUSH_KERNEL") == nullptr;
    const long long tiles_tile_mode = (((long long)P + rows_tiles - 1) / rows_tiles + slots - 1) / slots;
    bool force_peers = push_default && (ktiles > 0 ? ktiles > 1 : tiles_tile_mode > 1);
    using kern_t = void (*)(const ntgb_devtab, const ntgb_eval_args, int, int, int, int);
    auto pick = [&](int tile_rows) -> kern_t {
        const bool push_ok = tile_rows <= block; /* one thread per pair of a tile */
        if (wide) return full ? ntg_eval_small_kernel<PK, true, false, true, 512> : ntg_eval_small_kernel<PK, false, false, true, 512>;
        if (!full) return ntg_eval_small_kernel<PK, false, false, true>;
        if (hot_dense) return ntg_eval_small_kernel<PK, true, true, false, 256, true>;
        if (!(hot && (a.npeers == 0 || push_ok))) return ntg_eval_small_kernel<PK, true, false, true>;
        return (a.npeers > 0 || force_peers) && push_ok ? ntg_eval_small_kernel<PK, true, true, true>
                                                        : ntg_eval_small_kernel<PK, true, true, false>;
    };
    kern_t kern = pick(rows);
    int nb = 0;
    if (int rc = resident_blocks((const void *)kern, block, smem, L->max_smem_optin, &nb)) return rc;
    if (ktiles > 0 && nb * L->sm_count < even_grid) { /* fewer resident CTAs than assumed: tiles of G*R */
        ktiles = 0;
        R = R_tiles;
        rows = rows_tiles;
        smem = smem_rows(rows);
        force_peers = push_default && tiles_tile_mode > 1;
        kern = pick(rows);
        if (int rc = resident_blocks((const void *)kern, block, smem, L->max_smem_optin, &nb)) return rc;
    }
    const int ntiles = (P + rows - 1) / rows;
    int grid = nb * L->sm_count;
    if (grid > ntiles) grid = ntiles;
    if (ktiles > 0) grid = even_grid;
    if (grid < 1) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)L->args.stream;
    /* programmatic dependent launch: the prologue (tables into registers / shared memory) runs while
     * the kernel in front of this one in the stream drains; NTG_B200_NO_PDL=1 turns it off */
    static const bool no_pdl = getenv("NTG_B200_NO_PDL") != nullptr;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    const int pdl = no_pdl ? 0 : 1;
    if (pdl) {
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    static const bool no_rot = getenv("NTG_B200_NO_PUSH_ROTATE") != nullptr; /* A/B */
    const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, T, L->args, G, R, segtot, pdl | (no_rot ? 0 : 2) | (ktiles << 4) | (ktiles > 0 ? rows << 10 : 0));
    static const bool dbg = getenv("NTG_B200_DEBUG") != nullptr;
    if (dbg)
        fprintf(stderr, "K1s launch: grid %d block %d smem %zu G %d R %d rows %d tiles/CTA (even split) %d P %d nb %d: %s\n",
                grid, block, smem, G, R, rows, ktiles, P, nb, cudaGetErrorString(le));
    return (int)le;
}

} /* namespace ntgb */
#endif
