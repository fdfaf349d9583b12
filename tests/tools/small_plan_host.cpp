// Test infrastructure: the launch geometry of K1s (ntg_b200/csrc/ntg_small_plan.h) compiled for the
// host.  Reads "P nbps S nout nC segtot sm_count" lines from stdin; for each prints the plan and, per
// CTA, nothing -- the coverage check is done here and reported as one line:
//   block G R rows ktiles even_grid smem grid ok max_tiles_per_cta max_tile
// ok = every problem of [0, P) lies in exactly one tile, every tile has 1..rows problems, rows <= G*R,
// the tiles of a CTA follow the kernel's walk (small_cta_tiles).
#include <cstdio>
#include <vector>

#include "ntg_small_plan.h"

int main()
{
    int P, nbps, S, nout, nC, segtot, sm;
    while (std::scanf("%d %d %d %d %d %d %d", &P, &nbps, &S, &nout, &nC, &segtot, &sm) == 7) {
        ntgb::SmallPlanKnobs kn{0, 0, false, 8};
        const ntgb::SmallPlan pl = ntgb::plan_small_launch(P, nbps, S, nout, nC, segtot, sm, kn);
        // the launcher's grid (nb resident CTAs per SM as assumed by the plan)
        long long ntiles = ((long long)P + pl.rows - 1) / pl.rows;
        int grid = pl.slots < ntiles ? pl.slots : (int)ntiles;
        if (pl.ktiles > 0) grid = pl.even_grid;
        std::vector<int> hit((size_t)P, 0);
        bool ok = pl.rows <= pl.G * pl.R && pl.rows >= 1 && grid >= 1;
        int max_tiles = 0, max_tile = 0;
        for (int b = 0; b < grid && ok; b++) {
            const ntgb::SmallCtaTiles t = ntgb::small_cta_tiles(P, pl.rows, grid, b, pl.ktiles > 0);
            if (t.ntl > max_tiles) max_tiles = t.ntl;
            for (int it = 0; it < t.ntl; it++) {
                const int p0 = t.p_first + it * t.pstride;
                const int np = t.pend - p0 < pl.rows ? t.pend - p0 : pl.rows;
                if (np < 1 || np > pl.rows || p0 < 0 || p0 + np > P) { ok = false; break; }
                if (np > max_tile) max_tile = np;
                for (int p = p0; p < p0 + np; p++) hit[(size_t)p]++;
            }
        }
        for (int p = 0; p < P && ok; p++) ok = hit[(size_t)p] == 1;
        if (pl.ktiles > 0) ok = ok && max_tiles <= pl.ktiles;
        std::printf("%d %d %d %d %d %d %zu %d %d %d %d\n", pl.block, pl.G, pl.R, pl.rows, pl.ktiles, pl.even_grid, pl.smem,
                    grid, ok ? 1 : 0, max_tiles, max_tile);
    }
    return 0;
}
