/*
 * ntg_small_plan.h -- the launch geometry of K1s (ntg_eval_small.cuh) as plain
 * host code: shared-memory layout of a tile, and how a batch of P problems is cut
 * into tiles and dealt to the CTAs.  No CUDA types: the launcher calls it, and the
 * CPU test suite compiles it with g++ and checks, for thousands of batch sizes and
 * shapes, that every problem lands in exactly one tile that fits its buffers
 * (tests/test_small_plan.py, tests/tools/small_plan_host.cpp).
 */
#ifndef NTG_SMALL_PLAN_H_
#define NTG_SMALL_PLAN_H_

#include <stddef.h>

#include "ntg_b200.h"

#if defined(__CUDACC__)
#define NTG_HD __host__ __device__
#else
#define NTG_HD
#endif

namespace ntgb {

struct SmallSmem {
    int GR, nbps, S, nout, nC, segtot;
    /* D, f and viol hold one ROW of `pitch` doubles per (band slot, problem): phase A lanes are
     * consecutive breakpoints of one row (stride 1), phase B lanes are consecutive problems (stride
     * pitch) and a chain walks its row with 16-byte loads.  pitch is even with pitch/2 odd, so the
     * eight lanes of a quarter-warp hit eight different 16-byte bank groups. */
    NTG_HD int pitch() const
    {
        int p = (nbps + 1) & ~1;
        if (((p >> 1) & 1) == 0) p += 2;
        return p;
    }
    NTG_HD static size_t even(size_t n) { return (n + 1) & ~(size_t)1; }
    NTG_HD size_t D_off() const { return 0; }                                         /* [S][GR][pitch] */
    NTG_HD size_t f_off() const { return (size_t)S * GR * pitch(); }                  /* [GR][pitch]    */
    NTG_HD size_t viol_off() const { return f_off() + (size_t)GR * pitch(); }         /* [GR][pitch]    */
    NTG_HD size_t DI_off() const { return viol_off() + (size_t)GR * pitch(); }        /* [GR][S]        */
    NTG_HD size_t DF_off() const { return DI_off() + even((size_t)GR * S); }          /* [GR][S]        */
    NTG_HD size_t cI_off() const { return DF_off() + even((size_t)GR * S); }          /* [GR]           */
    NTG_HD size_t cF_off() const { return cI_off() + even(GR); }                      /* [GR]           */
    NTG_HD size_t res_off() const { return cF_off() + even(GR); }                     /* [GR][2] (objective, violation) */
    NTG_HD size_t dt_off() const { return res_off() + 2 * (size_t)GR; }               /* wt, Wf: [2][pitch + 2] */
    NTG_HD size_t C_off() const { return dt_off() + 2 * (size_t)(pitch() + 2); }      /* [2][GR*nC]     */
    NTG_HD size_t seg_off() const { return C_off() + 2 * even((size_t)GR * nC); }     /* ints, see the kernel */
    NTG_HD size_t bytes() const
    {
        /* ints: run tables, cost run, chain table, column list; then (8-byte aligned) the peer table pointers */
        const size_t ints = 2 * (size_t)segtot + 4 + (size_t)(nC + 1) * 10;
        return seg_off() * 8 + ((ints + 1) & ~(size_t)1) * 4 + (NTGB_MAXPEERS + 1) * 8 + 8;
    }
};

/* what launch_eval_small decides before it knows how many CTAs are resident per SM */
struct SmallPlan {
    int block;       /* threads per CTA: 256 (two CTAs per SM), 512 for 257..512 breakpoints (one per SM), fewer for P < G */
    int G;           /* problems per round: block / nbps */
    int R;           /* rounds of phase A per tile */
    int rows;        /* problems a tile's buffers hold: G*R, or the largest tile of an even split */
    int ktiles;      /* > 0: EVEN split with this many tiles per CTA; 0: tiles of `rows` dealt round-robin */
    int even_grid;   /* CTAs of the even split */
    int R_tiles, rows_tiles; /* the round-robin geometry (the fallback when fewer CTAs are resident than assumed) */
    int slots;       /* CTAs assumed resident on the GPU */
    size_t smem;     /* dynamic shared memory of a CTA, bytes */
};

struct SmallPlanKnobs {
    int rounds;      /* > 0: NTG_B200_ROUNDS (tuning; turns the even split off) */
    int smem_cap_kb; /* > 0: NTG_B200_SMEMCAP; default 100 (two CTAs per SM) / 200 */
    bool no_even;    /* NTG_B200_NO_EVEN_SPLIT */
    int even_max;    /* NTG_B200_EVEN_MAXTILES, default 8 */
};

/* R rounds per tile: enough (problem, column) chains to fill the CTA in phase B; for small batches,
 * enough problems per tile that the whole batch is ONE wave of resident CTAs (a second, mostly empty
 * wave would double the latency); within ~100 KB of shared memory.
 *
 * Batches of a few tiles per CTA: whole tiles of G*R problems dealt round-robin leave some CTAs a
 * tile more than others (8192 lane changes: 683 tiles of 12 on 296 CTAs, three for some and two for
 * the rest; CFG-3: 228 tiles of 36 and 68 empty slots).  An EVEN split gives CTA b the contiguous
 * problems [b*P/grid, (b+1)*P/grid) and every CTA the same number of tiles, ktiles, of at most `rows`
 * problems, with buffers of exactly that many rows.  Taken when it shortens the busiest CTA's
 * critical path, counted as rounds of phase A plus passes of phase B over its (problem, column)
 * chains (measured, lane changes of 64 breakpoints: 4096 problems 12.6 -> 10.9 us, 8192 21.4 -> 18.9,
 * 16384 36.2 -> 34.6; no difference beyond 8 tiles per CTA). */
inline SmallPlan plan_small_launch(int P, int nbps, int S, int nout, int nC, int segtot, int sm_count,
                                   const SmallPlanKnobs &kn)
{
    SmallPlan pl{};
    const bool wide = nbps > 256; /* one problem per round on a CTA of 512 threads (same registers per thread, one CTA per SM) */
    int block = wide ? 512 : 256;
    int G = block / nbps;
    if (G > P) {
        G = P > 0 ? P : 1;
        int need = ((G * nbps) + 31) / 32 * 32;
        if (need < 64) need = 64;
        if (need < block) block = need;
    }
    const int slots = (wide ? 1 : 2) * sm_count; /* __launch_bounds__(256, 2) / (512, 1) */
    const int ncol = nC + 1;
    auto smem_rows = [&](int rows) { return SmallSmem{rows, nbps, S, nout, nC, segtot}.bytes(); };
    int R = (block + G * ncol - 1) / (G * ncol);
    const int r_wave = (int)(((long long)P + (long long)G * slots - 1) / ((long long)G * slots));
    if (r_wave <= 8) R = r_wave;   /* single wave */
    if (kn.rounds > 0) R = kn.rounds;
    if (R < 1) R = 1;
    if (R > 8) R = 8;
    const size_t smem_cap = (size_t)(kn.smem_cap_kb > 0 ? kn.smem_cap_kb : (wide ? 200 : 100)) * 1024;
    while (R > 1 && smem_rows(G * R) > smem_cap) R--;
    int rows = G * R;
    pl.R_tiles = R;
    pl.rows_tiles = rows;
    int ktiles = 0, even_grid = 0;
    if (!kn.no_even && kn.rounds <= 0 && P >= G) {
        const long long nt = ((long long)P + rows - 1) / rows;
        const long long grid_t = nt < slots ? nt : slots;
        const long long per_cta_t = (nt + grid_t - 1) / grid_t;
        if (per_cta_t <= (kn.even_max > 0 ? kn.even_max : 8)) {
            const long long cand = ((long long)P + G - 1) / G;
            const int ge = cand < slots ? (int)cand : slots;
            const int n = (P + ge - 1) / ge; /* problems of the busiest CTA */
            int tcap = 8 * G < block ? 8 * G : block;
            while (tcap > 1 && smem_rows(tcap) > smem_cap) tcap--;
            const int k = (n + tcap - 1) / tcap;
            if (k <= 63) {
                const long long nte = (long long)k * ge;
                const int tmax = (int)((P + nte - 1) / nte);
                const int Re = (tmax + G - 1) / G;
                auto cost = [&](int t, int r) { return r + (t * ncol + block - 1) / block; };
                /* a tie goes to the even split when some CTAs would get a tile more than others, and to whole
                 * tiles for one tile per CTA (fewer CTAs: more of the next launch's prologues overlap) */
                if ((long long)k * cost(tmax, Re) < per_cta_t * cost(rows, R) + (per_cta_t > 1 ? 1 : 0)) {
                    ktiles = k;
                    even_grid = ge;
                    R = Re;
                    rows = tmax;
                }
            }
        }
    }
    pl.block = block;
    pl.G = G;
    pl.R = R;
    pl.rows = rows;
    pl.ktiles = ktiles;
    pl.even_grid = even_grid;
    pl.slots = slots;
    pl.smem = smem_rows(rows);
    return pl;
}

/* The tiles of CTA b of `grid`, as the kernel walks them (ntg_eval_small_kernel: p_first, pstride, pend
 * and the loop `p0 = p_first + it * pstride` with min(GR, pend - p0) problems per tile).  A restatement
 * for the CPU tests; the kernel spells it inline. */
struct SmallCtaTiles {
    int p_first, pstride, pend, ntl;
};
inline SmallCtaTiles small_cta_tiles(int P, int GR, int grid, int b, bool even)
{
    SmallCtaTiles t{};
    if (even) {
        const int q = P / grid, rem = P - q * grid;
        t.p_first = b * q + (b < rem ? b : rem);
        t.pend = t.p_first + q + (b < rem ? 1 : 0);
        t.pstride = GR;
    } else {
        t.p_first = b * GR;
        t.pend = P;
        t.pstride = grid * GR;
    }
    t.ntl = t.pend > t.p_first ? (t.pend - t.p_first + t.pstride - 1) / t.pstride : 0;
    return t;
}

} /* namespace ntgb */
#endif
