/*
 * ntg_eval_cluster_hot.cuh -- K1c/H: the steady-state instantiation of the
 * cluster evaluator (funobj mode 2 + funcon mode 2, band Jacobian, f / g / c /
 * J all requested, Z not requested -- what a solver asks for on every iterate).
 *
 * Same math, same reference citations and the same thread <-> breakpoint
 * mapping as K1c (ntg_eval_cluster.cuh).  What changes is how the Jacobian --
 * 87 % of a problem's 706 KB at CFG-5 -- leaves the SM, driven by ncu
 * (profiles/r01_v14_cfg5_K1c_ncu.txt): with one CTA of 8 warps per SM, phase
 * A's arithmetic (436 us per 4096 problems) and its stores (443 us of traffic)
 * did not overlap -- 30 % of the stall samples were long-scoreboard waits on
 * the source registers of in-flight STGs (LSU queue full), 18 % the cluster
 * barrier and 7 % its fence draining ~200 outstanding stores per thread.
 *
 *   - ROWS THROUGH SHARED MEMORY, DRAINED BY THE COPY ENGINE.  One stage of
 *     the ring = the kMaxOrd band rows of one (constraint m, output j).  In
 *     the tiled band layout (include/ntg_b200.h: one tile per CTA of the
 *     cluster) those rows are ONE contiguous block of kMaxOrd*cnt doubles
 *     (12.9 KB at CFG-5) of the problem's Jacobian: the compute warps write
 *     their breakpoint's values with STS (lanes = consecutive breakpoints,
 *     conflict-free), a dedicated service warp waits on the stage's `full`
 *     mbarrier, issues ONE cp.async.bulk shared->global for the whole stage
 *     and hands the stage back through its `empty` mbarrier once the copy
 *     engine has READ it (wait_group.read).  No compute warp ever waits for a
 *     global store, and every CTA streams one contiguous 308 KB block per
 *     problem.  (Measured, tools/bulk_bw.cu: bulk stores of 1.6 KB rows with
 *     1.6 KB gaps between them -- a CTA's half of every row in the untiled
 *     layout -- reach 3.6 TB/s, contiguous 12.8 KB blocks 6.4 TB/s; one
 *     issuing thread sustains a copy per ~175 cycles, so small copies cap an
 *     SM near its fair share of HBM and it can never catch up after a phase
 *     without stores.)  A stage that starts on an odd element (blocks are only
 *     8-byte aligned when cnt is odd) is staged one slot to the right so that
 *     its even-aligned body is a legal 16-byte bulk copy; the service warp
 *     stores the odd head / tail element itself.
 *   - the cluster barriers no longer sit behind the store stream: the
 *     gradient band D is written AFTER the Jacobian rows, the barrier that
 *     frees D for the next problem is arrived at right after the quadrature and
 *     waited for a whole phase A later (free), and the one real rendezvous
 *     (D complete cluster-wide) has only a handful of stores in front of it.
 *   - quadrature columns are dealt to the CTA that owns their breakpoints:
 *     only columns whose support crosses the CTA boundary read distributed
 *     shared memory (215 cycles) instead of local shared memory (29).
 *   - coefficients are single-buffered: the next problem's window is fetched
 *     (cp.async) as soon as every thread has expanded its flat outputs.
 *   - the scalar cost chain and the result pair are finished by the service
 *     warp of rank 0 while the compute warps run the quadrature.
 */
#ifndef NTG_EVAL_CLUSTER_HOT_CUH_
#define NTG_EVAL_CLUSTER_HOT_CUH_

#include "ntg_eval_cluster.cuh"

namespace ntgb {


/* ---- mbarrier / bulk-copy PTX ---- */
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("{\n .reg .b64 t;\n mbarrier.arrive.shared::cta.b64 t, [%0];\n}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ unsigned mbar_try_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
/* generic-proxy writes to shared memory become visible to the async proxy (the copy engine) */
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_store(double *gdst, unsigned ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
/* one lane of a converged warp; the compiler knows a single thread runs the guarded code */
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
template <int BYTE_OFF>
__device__ __forceinline__ void sts_f64(unsigned addr, double v)
{
    asm volatile("st.shared.f64 [%0+%1], %2;\n" ::"r"(addr), "n"(BYTE_OFF), "d"(v));
}

/* mbarrier in ANOTHER CTA of the cluster */
__device__ __forceinline__ unsigned map_to_rank(unsigned local_smem_addr, int rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
/* publishes this thread's earlier writes (incl. stores into the target CTA's shared memory) cluster-wide */
__device__ __forceinline__ void mbar_arrive_remote_release(unsigned remote_bar)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(remote_bar) : "memory");
}
/* nothing to publish ("I am done reading"): no fence */
__device__ __forceinline__ void mbar_arrive_remote_relaxed(unsigned remote_bar)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads)
{
    asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

struct ClusterHotSmem {
    int bpc, nbps, S, cwin, ord, nst, halo, cl;
    int plan_n, plan_cols; /* this CTA's share of the quadrature plan in shared memory (0 = left in global memory) */
    __host__ __device__ static size_t even(size_t n) { return (n + 1) & ~(size_t)1; }
    /* one stage: [ord][cnt] values (cnt <= bpc) behind an optional alignment slot, even size */
    __host__ __device__ size_t stage_doubles() const { return even((size_t)ord * bpc + 1); }
    __host__ __device__ int dpitch() const { return bpc + 2 * halo; }                                 /* a row of D: halo | own | halo */
    __host__ __device__ size_t ring_off() const { return 0; }                                        /* [nst][stage], 16-byte aligned */
    __host__ __device__ size_t D_off() const { return (size_t)nst * stage_doubles(); }                /* [S][dpitch] */
    __host__ __device__ size_t DI_off() const { return D_off() + (size_t)S * dpitch(); }              /* [S]         */
    __host__ __device__ size_t DF_off() const { return DI_off() + S; }                                /* [S]         */
    __host__ __device__ size_t viol_off() const { return DF_off() + S; }                              /* u64 [2] this CTA's maximum */
    __host__ __device__ size_t violx_off() const { return viol_off() + 2; }                           /* u64 [2][cl] (rank 0): every rank's */
    __host__ __device__ size_t sc_off() const { return violx_off() + 2 * (size_t)cl; }                /* [2][2] cI, cF (rank 0) */
    __host__ __device__ size_t fall_off() const { return sc_off() + 4; }                              /* rank 0: [2][nbps]; others: [2][bpc] own integrand */
    __host__ __device__ size_t dt_off() const { return fall_off() + 2 * (size_t)nbps; }               /* [nbps]      */
    __host__ __device__ size_t wf_off() const { return dt_off() + nbps; }                             /* [nbps] node weights (fast variant) */
    __host__ __device__ size_t C_off() const { return wf_off() + nbps; }                              /* [cwin]      */
    __host__ __device__ size_t bar_off() const { return C_off() + cwin; }                             /* u64 [2*nst + 4] full, empty, ready[2], free[2] */
    __host__ __device__ size_t plan_off() const { return bar_off() + 2 * (size_t)nst + 4; }           /* int2 [plan_n], int [plan_cols+1] */
    __host__ __device__ size_t bytes() const { return (plan_off() + plan_n) * 8 + (size_t)(plan_cols + 2) * 4 + 16; }
};

/* outputs per ring stage: fewer, larger stages = fewer handshakes and larger bulk copies */
#ifndef HOT_JPS
#define HOT_JPS 2
#endif
template <class PK>
__host__ __device__ constexpr int hot_jps()
{
    return (PK::kNout % HOT_JPS == 0) ? HOT_JPS : 1;
}

/* band values of ONE output from the register table (every output shares table 0) */
template <class PK, int J, unsigned long long MASK>
__device__ __forceinline__ void band_one_output(const double *Bt, const double *df, double (&v)[PK::kMaxOrd])
{
    constexpr int MD = PK::md(J);
    constexpr int IZ = pk_iz<PK>(J);
    static_for<0, PK::kMaxOrd>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        double acc = 0.0;
        static_for<0, MD>([&](auto lc) {
            constexpr int l = decltype(lc)::value;
            if constexpr (((MASK >> (IZ + l)) & 1ull) != 0ull) acc = acc + df[IZ + l] * Bt[k * MD + l];
        });
        v[k] = acc;
    });
}

template <class PK, bool PEERS>
__global__ void __launch_bounds__(256, 1)
ntg_eval_cluster_hot_kernel(const ntgb_devtab T, const ntgb_eval_args A, int cwin, int plan_smem, int NST, int plan_share,
                            int pdl)
{
    /* programmatic dependent launch (see ntg_eval_small.cuh): everything up to griddepcontrol.wait reads
     * batch-shared tables only and overlaps the grid in front of this one in the stream */
    if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr int NOUT = PK::kNout;
    constexpr int NZ = pk_nz<PK>();
    constexpr int MD0 = PK::md(0);
    constexpr int ORD = PK::kMaxOrd;
    constexpr int NB = ORD * MD0; /* ONE table */
    constexpr int NCON = PK::kNnltc;
    constexpr int JPS = hot_jps<PK>();  /* outputs per stage */
    constexpr int NJG = NOUT / JPS;     /* stages per constraint row */
    extern __shared__ __align__(16) double smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int CL = T.plan_cl, bpc = T.plan_bpc, H = T.plan_halo;
    const int nbps = T.nbps, nC = T.nC, P = A.P, S = T.S, ncoef0 = T.ncoef[0];
    /* quadrature columns of this CTA: a contiguous range, so that the chains stay inside its own D */
    const int c_lo = (int)(((long long)ncoef0 * rank) / CL), c_hi = (int)(((long long)ncoef0 * (rank + 1)) / CL);
    const ClusterHotSmem L{bpc, nbps, S, cwin, ORD * JPS, NST, H, CL, plan_smem ? plan_share : 0, plan_smem ? (ncoef0 + CL - 1) / CL + 1 : 0};
    const int dpitch = L.dpitch();
    double *ring_s = smem + L.ring_off();
    double *D_s = smem + L.D_off();
    double *DI_s = smem + L.DI_off();
    double *DF_s = smem + L.DF_off();
    unsigned long long *viol_s = reinterpret_cast<unsigned long long *>(smem + L.viol_off());   /* [2] */
    unsigned long long *violx_s = reinterpret_cast<unsigned long long *>(smem + L.violx_off()); /* rank 0: [2][CL] */
    double *sc_s = smem + L.sc_off();     /* rank 0: [buf][0] initial cost, [buf][1] final cost */
    double *fall_s = smem + L.fall_off(); /* rank 0: the integrand of ALL breakpoints [buf][nbps]; else this CTA's [buf][bpc] */
    double *dt_s = smem + L.dt_off();
    double *wf_s = smem + L.wf_off();     /* (dt[n-1] + dt[n])/2: the trapezoid rule as node weights */
    double *C_s = smem + L.C_off();
    unsigned long long *bar_s = reinterpret_cast<unsigned long long *>(smem + L.bar_off());
    const unsigned ring_a = smem_u32(ring_s);
    const unsigned STAGE_BYTES = (unsigned)L.stage_doubles() * 8u;
    const unsigned full_a = smem_u32(bar_s), empty_a = full_a + 8u * (unsigned)NST;
    const unsigned ready_a = empty_a + 8u * (unsigned)NST; /* [2] rank 0: the other ranks' integrand / violation of buffer b arrived */
    const unsigned free_a = ready_a + 16u;                 /* [2] other ranks: rank 0 is done with buffer b */
    const int fpitch = rank == 0 ? nbps : bpc;

    /* mode 2 / mode 2: counts gate on != 0 (reference src/ntg.c:309-314) */
    const bool doI = PK::cb_icf != nullptr && T.nicf != 0;
    const bool doU = PK::cb_ucf != nullptr && T.nucf != 0;
    const bool doF = PK::cb_fcf != nullptr && T.nfcf != 0;
    const bool doCI = PK::cb_nlicf != nullptr && T.nnlic != 0;
    const bool doCF = PK::cb_nlfcf != nullptr && T.nnlfc != 0;

    const int NCT = (int)blockDim.x - 32; /* compute threads: warps pinned to breakpoints */
    const int NCW = NCT / 32;
    const bool service = (int)threadIdx.x >= NCT;
    const int lane = threadIdx.x & 31;
    const int last_rank = (nbps - 1) / bpc;

    /* ---- once per CTA ---- */
    for (int i = threadIdx.x; i < nbps - 1; i += blockDim.x) dt_s[i] = __ldg(T.bps + i + 1) - __ldg(T.bps + i);
    for (int n = threadIdx.x; n < nbps; n += blockDim.x) {
        const double lo = n >= 1 ? (__ldg(T.bps + n) - __ldg(T.bps + n - 1)) * 0.5 : 0.0;
        const double hi = n + 1 < nbps ? (__ldg(T.bps + n + 1) - __ldg(T.bps + n)) * 0.5 : 0.0;
        wf_s[n] = lo + hi;
    }
    /* the second plan of the tables (ntg_kernel_args.h): entries of column cl at [hptr[cl], hptr[cl+1]) */
    const int *hptr = T.plan_ptr + ncoef0 + 1;
    const int2 *plan = T.plan;
    const int *plan_ptr = hptr;
    int plan_shift = 0, col_shift = 0; /* plan[e - plan_shift], plan_ptr[cl - col_shift] */
    if (plan_smem) { /* this CTA's columns only; re-read for every problem: keep it next to the data it indexes */
        int2 *pl_s = reinterpret_cast<int2 *>(smem + L.plan_off());
        int *pp_s = reinterpret_cast<int *>(pl_s + plan_share);
        const int e0 = __ldg(hptr + c_lo), e1 = __ldg(hptr + c_hi);
        for (int i = threadIdx.x; i < e1 - e0; i += blockDim.x) pl_s[i] = __ldg(T.plan + e0 + i);
        for (int i = threadIdx.x; i <= c_hi - c_lo; i += blockDim.x) pp_s[i] = __ldg(hptr + c_lo + i);
        plan = pl_s;
        plan_ptr = pp_s;
        plan_shift = e0;
        col_shift = c_lo;
    }
    if (threadIdx.x == 0) {
        sc_s[0] = sc_s[1] = sc_s[2] = sc_s[3] = 0.0;
        viol_s[0] = 0ull;
        viol_s[1] = 0ull;
        for (int i = 0; i < 2 * CL; i++) violx_s[i] = 0ull;
        for (int i = 0; i < NST; i++) {
            mbar_init(full_a + 8u * (unsigned)i, (unsigned)NCW); /* one arrival per compute warp */
            mbar_init(empty_a + 8u * (unsigned)i, 1u);           /* the service warp */
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(ready_a + 8u * (unsigned)b, (unsigned)(CL - 1)); /* the service warps of the other ranks */
            mbar_init(free_a + 8u * (unsigned)b, 1u);                  /* the service warp of rank 0 */
        }
    }

    /* ---- once per thread: its breakpoint (own or halo), its table slice ---- */
    const int bp0 = rank * bpc;
    const int lbp = threadIdx.x;
    const int cnt = nbps - bp0 < bpc ? (nbps - bp0 > 0 ? nbps - bp0 : 0) : bpc; /* breakpoints of this CTA */
    const bool own = !service && lbp < cnt;
    /* lanes [cnt, cnt + 2H): H breakpoints in front of and H behind this CTA's range -- their cost
     * derivatives are computed here a second time, so that a quadrature chain never leaves the CTA */
    const int hidx = lbp - cnt;
    const int bp = own ? bp0 + lbp : (hidx < H ? bp0 - H + hidx : bp0 + cnt + hidx - H);
    const bool halo = !service && !own && hidx < 2 * H && bp >= 0 && bp < nbps;
    const bool live = own || halo;
    const int dpos = own ? H + lbp : (hidx < H ? hidx : cnt + hidx); /* position in a row of D */
    double Bt[NB];
    int off0 = 0;
    /* active-variable mask of this thread's breakpoint class, bit iz_j + d (updateZ fills listed variables only) */
    unsigned long long zmask = 0ull;
    static_for<0, NOUT>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const unsigned mk = T.avmask[(bp == 0 ? 1 : 0) | (bp == nbps - 1 ? 2 : 0)][j];
        zmask |= (unsigned long long)(mk & ((1u << MD0) - 1u)) << pk_iz<PK>(j);
    });
    {
        off0 = live ? __ldg(T.off[0] + bp) : 0;
#pragma unroll
        for (int k = 0; k < ORD; k++)
#pragma unroll
            for (int d = 0; d < MD0; d++) Bt[k * MD0 + d] = live ? __ldg(T.Bt[0] + (size_t)(k * MD0 + d) * nbps + bp) : 0.0;
    }

    /* this CTA only needs the coefficient windows of ITS breakpoints: per output the range
     * [w0, w0 + wl) with w0 = min offset, wl = max offset + order - w0 (same for every output) */
    __shared__ int win_s[2];
    if (threadIdx.x == 0) { win_s[0] = 0x7fffffff; win_s[1] = -1; }
    __syncthreads();
    if (live) {
        atomicMin(&win_s[0], off0);
        atomicMax(&win_s[1], off0);
    }
    cluster.sync(); /* publishes the mbarrier initialisation, cluster-wide (remote arrivals below) */
    const int w0 = win_s[0];
    const int wl = win_s[1] + ORD - w0;

    const int ncl = (int)(gridDim.x / CL); /* clusters in the grid */
    const int clid = (int)(blockIdx.x / CL);
    /* element index (from A.J) of this CTA's tile of problem p's trajectory rows (tiled band layout,
     * include/ntg_b200.h): [m][slot][breakpoint of the tile], n_st = JPS*ORD*cnt elements per stage */
    auto tile_base = [&](int p) { return ((size_t)p * T.ncnln + T.nnlic) * S + (size_t)rank * T.nnltc * S * bpc; };
    const unsigned n_st = (unsigned)(JPS * ORD * cnt);

    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory"); /* coefficients are read and results written from here on */

    if (service) {
        /* =================== the service warp: drain stages, finish the scalar cost =================== */
        unsigned st = 0, ph = 0;
        int prev_st = -1;
        int buf = 0;
        double res_y = 0.0, res_v = 0.0;
        double2 *my_peer = nullptr;
        if constexpr (PEERS) {
            if (rank == 0 && lane < A.npeers) my_peer = reinterpret_cast<double2 *>(A.peer_result[lane]);
        }
        unsigned use = 0; /* how often buffer `buf` has been used before: use = (iteration / 2) */
        int it = 0;
        const unsigned stage_flip = n_st & 1u; /* does the parity of a stage's first element flip from stage to stage? */
        for (int p = clid; p < P; p += ncl, buf ^= 1, it++) {
            use = (unsigned)(it >> 1);
            /* ONE bulk copy per stage, issued by an elected lane from uniform registers.  (A first version
             * copied row by row, each row from its own lane: the compiler serialises per-lane bulk copies
             * through R2UR, ~150 dependent instructions per stage on a single warp, and a thread sustains
             * only one small copy per ~175 cycles -- the drain, not HBM, was the limit.) */
            double *gp = A.J + tile_base(p);              /* stage (m = 0, j = 0) of this CTA's tile */
            unsigned par = (unsigned)(tile_base(p) & 1);  /* parity of the stage's first element */
#pragma unroll 1
            for (int s = 0; s < NCON * NJG; s++) {
                const unsigned sbase = ring_a + st * STAGE_BYTES;
                const double *stage = ring_s + (size_t)st * L.stage_doubles(); /* element i at stage[i + par] */
                const unsigned body = (n_st - par) & ~1u;
                mbar_wait(full_a + 8u * st, ph);
                const bool leader = elect_one();
                if (leader) {
                    if (body > 0) bulk_store(gp + par, sbase + 16u * par, body * 8u);
                    bulk_commit();
                }
                /* odd head / tail elements of the block */
                if (lane == 1 && par) st_stream(gp, stage[par]);
                if (lane == 2 && ((n_st - par) & 1u)) st_stream(gp + n_st - 1, stage[par + n_st - 1]);
                /* the stage issued one step earlier has been read by the copy engine: hand it back */
                if (leader) bulk_wait_read<1>();
                __syncwarp();
                if (prev_st >= 0 && leader) mbar_arrive(empty_a + 8u * (unsigned)prev_st);
                prev_st = (int)st;
                if (++st == (unsigned)NST) { st = 0; ph ^= 1u; }
                gp += n_st;
                par ^= stage_flip;
            }
            /* the compute warps have written this problem's integrand, end-point costs and violation */
            named_bar_sync(3, NCT + 32);
            if (rank != 0) {
                /* forward them to rank 0 (stores into ITS shared memory), once rank 0 is done with what this
                 * buffer held two problems ago.  The cluster-scope release behind the arrival costs a
                 * MEMBAR.GPU (~1.7k cycles): it is paid here, by a warp that has nothing else to do while
                 * the compute warps run the quadrature -- they never touch a cluster barrier. */
                mbar_wait_cluster(free_a + 8u * (unsigned)buf, (use & 1u) ^ 1u);
                double *f0 = cluster.map_shared_rank(fall_s, 0) + (size_t)buf * nbps + bp0;
                const double *fl = fall_s + (size_t)buf * bpc;
                if (doU)
                    for (int i = lane; i < cnt; i += 32) f0[i] = fl[i];
                if (lane == 0) {
                    cluster.map_shared_rank(violx_s, 0)[buf * CL + rank] = viol_s[buf];
                    viol_s[buf] = 0ull;
                    if (doF && rank == last_rank) cluster.map_shared_rank(sc_s, 0)[buf * 2 + 1] = sc_s[buf * 2 + 1];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_remote_release(map_to_rank(ready_a + 8u * (unsigned)buf, 0));
            } else {
                if (CL > 1) mbar_wait_cluster(ready_a + 8u * (unsigned)buf, use & 1u);
                /* IntegrateVector TRAPEZOID (src/integrator.c:21-24): a sequential chain over all breakpoints */
                double *fp = fall_s + (size_t)buf * nbps;
                if (doU) {
                    /* the trapezoid terms overwrite the integrand in place (term i needs f[i] and f[i+1]: 32 at a
                     * time in ascending order, every lane reads before any lane writes) -- a separate array of
                     * nbps terms was the 3 KB that kept a fourth ring stage out of shared memory */
                    for (int base = 0; base < nbps - 1; base += 32) {
                        const int i = base + lane;
                        double a = 0.0, b = 0.0;
                        if (i < nbps - 1) { a = fp[i]; b = fp[i + 1]; }
                        __syncwarp();
                        if (i < nbps - 1) fp[i] = (dt_s[i] * (b + a)) / 2;
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    double In = 0.0;
                    if (doU) {
#pragma unroll 8
                        for (int i = 0; i < nbps - 1; i++) In = In + fp[i];
                    }
                    unsigned long long vb = viol_s[buf];
                    viol_s[buf] = 0ull;
                    for (int r = 1; r < CL; r++) {
                        const unsigned long long v = violx_s[buf * CL + r];
                        vb = v > vb ? v : vb;
                    }
                    const double y = (sc_s[buf * 2 + 0] + In) + sc_s[buf * 2 + 1]; /* y = I + In + F, src/ntg.c:328 */
                    A.f[p] = y;
                    res_y = y;
                    res_v = __longlong_as_double((long long)vb);
                    if (A.result != nullptr) reinterpret_cast<double2 *>(A.result)[p] = make_double2(res_y, res_v);
                    for (int r = 1; r < CL; r++) mbar_arrive_remote_relaxed(map_to_rank(free_a + 8u * (unsigned)buf, r));
                }
                __syncwarp();
                if constexpr (PEERS) {
                    /* fused multi-GPU gather: lane r stores the pair into rank r's table -- one 16-byte
                     * peer store per table, all tables at once, pointers fetched once per launch */
                    const double yy = __shfl_sync(0xffffffffu, res_y, 0), vv = __shfl_sync(0xffffffffu, res_v, 0);
                    if (my_peer != nullptr) my_peer[(size_t)A.peer_row0 + p] = make_double2(yy, vv);
                }
            }
        }
        bulk_wait_all();
        cluster.sync(); /* nobody exits while a neighbour may still store into / arrive on its shared memory */
        return;
    }

    /* =================== compute warps =================== */
    auto stage_C = [&](int p) {
        const double *src = A.C + (size_t)p * nC + w0;
        for (int e = threadIdx.x; e < NOUT * wl; e += NCT) {
            const int j = e / wl, q = e - j * wl;
            cp_async8(C_s + e, src + T.iC[j] + q);
        }
        cp_async_commit();
    };
    if (clid < P) stage_C(clid);

    unsigned st = 0, ph = 1; /* a fresh `empty` barrier passes a wait on the opposite parity */
    const unsigned stage_flip = n_st & 1u;
    const unsigned cnt8 = (unsigned)cnt * 8u;
    int buf = 0;
    for (int p = clid; p < P; p += ncl, buf ^= 1) {
        cp_async_wait_all();
        named_bar_sync(1, NCT); /* this problem's coefficients landed; every warp is past the previous quadrature: D is free */

        double z[NZ];
        double *zp[NOUT];
        /* Zvalue, src/colloc.c:318-326 -- k ascending from 0.0 */
        static_for<0, NOUT>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr int IZ = pk_iz<PK>(j);
            const double *Cw = C_s + j * wl + (off0 - w0);
            double acc[MD0];
#pragma unroll
            for (int d = 0; d < MD0; d++) acc[d] = 0.0;
            if (live) {
#pragma unroll
                for (int k = 0; k < ORD; k++) {
                    const double ck = Cw[k];
#pragma unroll
                    for (int d = 0; d < MD0; d++) acc[d] = acc[d] + Bt[k * MD0 + d] * ck;
                }
            }
#pragma unroll
            for (int d = 0; d < MD0; d++) z[IZ + d] = ((zmask >> (IZ + d)) & 1ull) ? acc[d] : 0.0;
            zp[j] = &z[IZ];
        });
        named_bar_sync(1, NCT); /* everybody has read its window: fetch the next problem's */
        if (p + ncl < P) stage_C(p + ncl);

        int nstate = A.nstate;
        double viol = 0.0;
        /* parity of the first element of this problem's first stage; stage s starts s*n_st further on */
        const unsigned pe = (unsigned)(tile_base(p) & 1);

        /* nonlinear trajectory constraints, src/constraints.c:120-162, one ROW per inlined callback:
         * only row m's value and derivatives are consumed in iteration m */
        static_for<0, NCON>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            double cv[NCON];
            double dfc[NCON][NZ];
            double *dfp[NCON];
#pragma unroll
            for (int q = 0; q < NCON; q++) {
                cv[q] = 0.0;
                dfp[q] = dfc[q];
#pragma unroll
                for (int l = 0; l < NZ; l++) dfc[q][l] = 0.0;
            }
            /* the four inlined copies of the callback see "different" z: without this the compiler
             * merges them and keeps every row's derivatives live (144 registers at CFG-5, spills) */
#pragma unroll
            for (int l = 0; l < NZ; l++) asm volatile("" : "+d"(z[l]));
            /* lanes that own no breakpoint run along (halo lanes on their own z, the rest on zeros); only
             * their stores are predicated off: no divergence in the stage loop */
            int mode = 2, i = bp;
            PK::cb_nltcf(&mode, &nstate, &i, cv, dfp, zp);
            if (own) {
                note_abort(A, mode);
                st_stream(A.c + (size_t)p * T.ncnln + T.nnlic + (size_t)m * nbps + bp, cv[m]);
                viol = fmax(viol, row_violation(cv[m], nl_bound(T, false, T.nnlic + m), nl_bound(T, true, T.nnlic + m)));
            }
            const bool clean = sp_clean<NZ>(dfc[m], PK::sp_nltcf(m));
            static_for<0, NJG>([&](auto gc) {
                constexpr int jg = decltype(gc)::value;
                /* value (slot, this breakpoint) is element slot*cnt + lbp of the stage, staged at + parity */
                constexpr unsigned STAGE_IDX = (unsigned)(m * NJG + jg);
                const unsigned par = pe ^ (STAGE_IDX & stage_flip & 1u);
                const unsigned a0 = ring_a + st * STAGE_BYTES + ((unsigned)lbp + par) * 8u;
                unsigned ok = mbar_try_wait(empty_a + 8u * st, ph);
                static_for<0, JPS>([&](auto jjc) {
                    constexpr int jj = decltype(jjc)::value;
                    constexpr int j = jg * JPS + jj;
                    double v[ORD];
                    if (clean) band_one_output<PK, j, PK::sp_nltcf(m)>(Bt, dfc[m], v);
                    else band_one_output<PK, j, kDense>(Bt, dfc[m], v);
                    if (jj == 0 && !ok) mbar_wait(empty_a + 8u * st, ph);
                    static_for<0, ORD>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        /* (rows are contiguous: a lane that owns no breakpoint must not store) */
                        if (own) sts_f64<0>(a0 + (unsigned)(jj * ORD + k) * cnt8, v[k]);
                    });
                });
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_a + 8u * st);
                if (++st == (unsigned)NST) { st = 0; ph ^= 1u; }
            });
        });

        if (own) {
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
                    int mode = 2;
                    int offj[NOUT];
#pragma unroll
                    for (int j = 0; j < NOUT; j++) offj[j] = off0;
                    PK::cb_nlicf(&mode, &nstate, cv, dfp, zp);
                    note_abort(A, mode);
#pragma unroll
                    for (int m = 0; m < PK::kNnlic; m++) {
                        st_stream(A.c + (size_t)p * T.ncnln + m, cv[m]);
                        viol = fmax(viol, row_violation(cv[m], nl_bound(T, false, m), nl_bound(T, true, m)));
                    }
                    emit_rows_regs<PK, true, PK::kNnlic, 0, true, 1>(T, A, Bt, offj, p, bp, dfc, 0);
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
                    int mode = 2;
                    int offj[NOUT];
#pragma unroll
                    for (int j = 0; j < NOUT; j++) offj[j] = off0;
                    const int rb = T.nnlic + T.nnltc * nbps;
                    PK::cb_nlfcf(&mode, &nstate, cv, dfp, zp);
                    note_abort(A, mode);
#pragma unroll
                    for (int m = 0; m < PK::kNnlfc; m++) {
                        st_stream(A.c + (size_t)p * T.ncnln + rb + m, cv[m]);
                        viol = fmax(viol, row_violation(cv[m], nl_bound(T, false, T.nnlic + T.nnltc + m),
                                                        nl_bound(T, true, T.nnlic + T.nnltc + m)));
                    }
                    emit_rows_regs<PK, true, PK::kNnlfc, 2, true, 1>(T, A, Bt, offj, p, bp, dfc, rb);
                }
            }
            if (viol > 0.0) atomicMax(viol_s + buf, (unsigned long long)__double_as_longlong(viol));
        }

        if (live) {
            /* unintegrated (trajectory) cost, src/cost.c:99-132; halo lanes repeat a neighbour's breakpoint
             * for the band D only */
            if constexpr (PK::cb_ucf != nullptr) {
                if (doU) {
                    double fv = 0.0;
                    double df[NZ];
#pragma unroll
                    for (int l = 0; l < NZ; l++) df[l] = 0.0;
                    int mode = 2, i = bp;
                    PK::cb_ucf(&mode, &nstate, &i, &fv, df, zp);
                    if (own) {
                        note_abort(A, mode);
                        fall_s[(size_t)buf * fpitch + (rank == 0 ? bp : lbp)] = fv; /* the service warp forwards it to rank 0 */
                    }
                    double *Dp = D_s + dpos;
                    band_from_regs<PK, true, true>(T, Bt, df, [&](auto, auto, double v) {
                        *Dp = v;
                        Dp += dpitch;
                    });
                }
            }
        }
        if (own) {
            /* initial cost (breakpoint 0: cluster rank 0), src/cost.c:4-36 */
            if constexpr (PK::cb_icf != nullptr) {
                if (doI && bp == 0) {
                    double fv = 0.0;
                    double df[NZ];
#pragma unroll
                    for (int l = 0; l < NZ; l++) df[l] = 0.0;
                    int mode = 2;
                    PK::cb_icf(&mode, &nstate, &fv, df, zp);
                    note_abort(A, mode);
                    sc_s[buf * 2 + 0] = fv;
                    double *Dp = DI_s;
                    band_from_regs<PK, true, true>(T, Bt, df, [&](auto, auto, double v) { *Dp++ = v; });
                }
            }
            /* final cost (last breakpoint: last cluster rank), src/cost.c:141-174 */
            if constexpr (PK::cb_fcf != nullptr) {
                if (doF && bp == nbps - 1) {
                    double fv = 0.0;
                    double df[NZ];
#pragma unroll
                    for (int l = 0; l < NZ; l++) df[l] = 0.0;
                    int mode = 2;
                    PK::cb_fcf(&mode, &nstate, &fv, df, zp);
                    note_abort(A, mode);
                    sc_s[buf * 2 + 1] = fv;
                    double *Dp = DF_s;
                    band_from_regs<PK, true, true>(T, Bt, df, [&](auto, auto, double v) { *Dp++ = v; });
                }
            }
        }
        named_bar_arrive(3, NCT + 32); /* integrand / end-point costs / violation of this problem: over to the service warp */
        named_bar_sync(2, NCT);        /* this CTA's D (own breakpoints and halo) is complete */

        /* ------- phase B: one trapezoid chain per gradient column (IntegrateFMatrixCols TRAPEZOID,
         * src/integrator.c:44-48, on the band of src/cost.c:118-132; ascending breakpoint), walking the
         * host-built plan.  Every chain of this CTA's columns stays inside its own D (own breakpoints +
         * halo): no distributed shared memory, no cluster barrier.  Then Vector3Add, src/ntg.c:329. ------- */
        {
            const int off_last = __ldg(T.off[0] + nbps - 1);
            const int jpitch = ORD * dpitch; /* one output's block of D */
            for (int cl = c_lo + (int)threadIdx.x; cl < c_hi; cl += NCT) {
                double gU[NOUT], dcur[NOUT];
#pragma unroll
                for (int j = 0; j < NOUT; j++) { gU[j] = 0.0; dcur[j] = 0.0; }
                int e = plan_ptr[cl - col_shift] - plan_shift;
                const int eend = plan_ptr[cl + 1 - col_shift] - plan_shift;
                if (doU && !PK::kExact) {
                    /* fast variant: sum_n Wf[n]*D[n] over the in-band entries (the band is zero at both ends of
                     * a column's support, so the node-weight form needs no end corrections) */
#pragma unroll 4
                    for (; e < eend; e++) {
                        const int2 en = plan[e];
                        const double w = wf_s[en.x];
#pragma unroll
                        for (int j = 0; j < NOUT; j++) gU[j] = gU[j] + w * D_s[j * jpitch + en.y];
                    }
                } else if (doU) {
                    if (e < eend) {
                        int2 en = plan[e];
                        if (en.y >= 0) {
#pragma unroll
                            for (int j = 0; j < NOUT; j++) dcur[j] = D_s[j * jpitch + en.y];
                        }
                        for (e++; e < eend; e++) {
                            en = plan[e];
                            const double dt = dt_s[en.x - 1];
                            double dn[NOUT];
                            if (en.y < 0) {
#pragma unroll
                                for (int j = 0; j < NOUT; j++) dn[j] = 0.0;
                            } else {
#pragma unroll
                                for (int j = 0; j < NOUT; j++) dn[j] = D_s[j * jpitch + en.y];
                            }
#pragma unroll
                            for (int j = 0; j < NOUT; j++) {
                                gU[j] = gU[j] + (dt * (dn[j] + dcur[j])) / 2;
                                dcur[j] = dn[j];
                            }
                        }
                    }
                }
                const int kF = cl - off_last;
#pragma unroll
                for (int j = 0; j < NOUT; j++) {
                    const int s0 = j * ORD; /* jk0_j: every output has the same order */
                    double gI = 0.0, gF = 0.0;
                    if (doI && cl < ORD) gI = DI_s[s0 + cl];              /* offset 0, src/colloc.c:254; rank 0's columns */
                    if (doF && kF >= 0 && kF < ORD) gF = DF_s[s0 + kF];   /* the last rank's columns */
                    st_stream(A.g + (size_t)p * nC + (size_t)j * ncoef0 + cl, (gI + gU[j]) + gF);
                }
            }
        }
    }
    cp_async_wait_all();
    cluster.sync(); /* nobody exits while a neighbour may still store into / arrive on its shared memory */
}

/* does this launch qualify for the steady-state cluster kernel? */
template <class PK>
int launch_eval_cluster_hot(const ntgb_launch *L)
{
    if constexpr (!pk_uniform_outputs<PK>() || PK::kMaxOrd * PK::md(0) > 64 || PK::kNnltc < 1 || PK::cb_nltcf == nullptr) {
        return -1001;
    } else {
        const ntgb_devtab &T = L->tab;
        const ntgb_eval_args &a = L->args;
        const int nbps = T.nbps, P = a.P;
        if (!devtab_one_table(T) || T.plan == nullptr || T.plan_halo < 0) return -1001;
        const bool hot = a.mode_obj == 2 && a.mode_con == 2 && a.jac_layout == NTGB_JAC_BAND && a.J != nullptr &&
                         a.f != nullptr && a.g != nullptr && a.c != nullptr && a.Z == nullptr && T.nnltc == PK::kNnltc &&
                         ((uintptr_t)a.J & 15u) == 0 && ((uintptr_t)a.result & 15u) == 0; /* 16-byte bulk copies / pair stores */
        if (!hot) return -1001;
        for (int j = 0; j < T.nout; j++)
            if (T.order[j] != PK::kMaxOrd) return -1001;
        /* cluster geometry, halo and plans were decided when the tables were built (ntg_core.cu) */
        const int CL = T.plan_cl, bpc = T.plan_bpc, H = T.plan_halo;
        if (CL < 2 || T.band_tile != bpc || bpc + 2 * H > 224) return -1001;
        const int block = (bpc + 2 * H + 31) / 32 * 32 + 32; /* + the service warp */
        /* the widest share of the plan any CTA of the cluster copies to shared memory */
        const int plan_share = T.plan_share > 0 ? T.plan_share : 1;
        /* as many ring stages as fit, but fewer than one problem has: the compute warps must not be able
         * to finish the NEXT problem's rows before the service warp has handed this one's integrand over */
        constexpr int NS = PK::kNnltc * PK::kNout / hot_jps<PK>();
        if (NS < 3) return -1001;
        int nst = 8, plan_smem = 1;
        if (const char *e = getenv("NTG_B200_HOT_STAGES")) nst = atoi(e);
        if (nst > NS - 1) nst = NS - 1;
        if (nst < 2) nst = 2;
        if (nst > 16) nst = 16;
        ClusterHotSmem lay{bpc, nbps, T.S, T.plan_cwin, PK::kMaxOrd * hot_jps<PK>(), nst, H, CL, plan_share, (T.ncoef[0] + CL - 1) / CL + 1};
        while (lay.nst > 2 && lay.bytes() > (size_t)L->max_smem_optin) lay.nst--;
        if (lay.bytes() > (size_t)L->max_smem_optin) {
            lay.plan_n = 0;
            lay.plan_cols = 0;
            plan_smem = 0;
            lay.nst = nst;
            while (lay.nst > 2 && lay.bytes() > (size_t)L->max_smem_optin) lay.nst--;
        }
        const size_t smem = lay.bytes();
        if (smem > (size_t)L->max_smem_optin) return -1001;
        if (getenv("NTG_B200_DEBUG"))
            fprintf(stderr, "K1c/H: CL %d bpc %d halo %d stages %d x %zu B (of %d per problem and CTA) plan in smem %d (%d entries) cwin %d smem %zu\n",
                    CL, bpc, H, lay.nst, lay.stage_doubles() * 8, NS, plan_smem, plan_share, T.plan_cwin, smem);
        static const bool force_peers = getenv("NTG_B200_FORCE_PEERS_KERNEL") != nullptr; /* A/B: code shape vs NVLink */
        auto kern = (a.npeers > 0 || force_peers) ? ntg_eval_cluster_hot_kernel<PK, true> : ntg_eval_cluster_hot_kernel<PK, false>;
        cudaError_t e = raise_smem_limit((const void *)kern, L->max_smem_optin);
        if (e != cudaSuccess) return (int)e;
        int nclusters = L->sm_count / CL;
        if (nclusters > P) nclusters = P;
        if (nclusters < 1) return 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(nclusters * CL));
        cfg.blockDim = dim3((unsigned)block);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)a.stream;
        static const bool no_pdl = getenv("NTG_B200_NO_PDL") != nullptr;
        const int pdl = no_pdl ? 0 : 1;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 2 : 1;
        return (int)cudaLaunchKernelEx(&cfg, kern, T, a, T.plan_cwin, plan_smem, lay.nst, plan_share, pdl);
    }
}

/* dispatcher used by NTGB_DEFINE_PACK: K1s (small, register tables) -> K1c/H, K1c (long horizon,
 * one shared table, thread-block clusters; /H = steady state) -> K1 (general) */
template <class PK>
int launch_dispatch(const ntgb_launch *L)
{
    if (L->abi != NTGB_KERNEL_ABI) return (int)cudaErrorInvalidValue; /* core and pack built from different headers */
    /* NTG_B200_KERNEL=general forces K1 (A/B measurements, tests of both kernels) */
    const char *env = getenv("NTG_B200_KERNEL");
    const bool force_general = env != nullptr && strcmp(env, "general") == 0;
    if (!force_general) {
        if constexpr (pk_tab_doubles<PK>() <= 64) {
            if (small_shape_ok<PK>(L->tab)) {
                const int rc = launch_eval_small<PK>(L);
                if (rc != -1001) return rc;
            }
        }
        if (L->tab.nbps > 256) {
            /* NTG_B200_KERNEL=cluster keeps the steady state on the general-mode cluster kernel (A/B) */
            if (!(env != nullptr && strcmp(env, "cluster") == 0)) {
                const int rc = launch_eval_cluster_hot<PK>(L);
                if (rc != -1001) return rc;
            }
            const int rc = launch_eval_cluster<PK>(L);
            if (rc != -1001) return rc;
        }
    }
    return launch_eval<PK>(L);
}

} /* namespace ntgb */
#endif
