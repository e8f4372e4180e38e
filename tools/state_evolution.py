"""How the cost of a C4 frame evolves with the simulated state: per-kernel times and clump statistics at several frame counts.
usage: python tools/state_evolution.py [frame ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import coupledwateranimation_b200 as cwa  # noqa: E402

marks = [int(a) for a in sys.argv[1:]] or [10, 60, 200, 500, 1000, 2000, 3000]
with cwa.Context(0) as ctx:
    grid, sph, wave = bench.build_scene(cwa, ctx)
    done = 0
    for m in marks:
        sph.coupled_step(wave, m - done, bench.COUPLING)
        done = m
        ctx.synchronize()
        ctx.timer_begin()
        sph.coupled_step(wave, 20, bench.COUPLING)
        ms = ctx.timer_end()
        ctx.profile_begin()
        sph.coupled_step(wave, 20, bench.COUPLING)
        prof = ctx.profile_end()
        done += 40
        cnt = grid.read(cwa.GRID_COUNTER, grid.num_cells_total)
        occ = cnt[cnt > 0]
        p = sph.download()
        ok = ~np.isnan(p["pos"][:, :3]).any(1)
        top = " ".join(f"{k}={v[0] / v[1] * 1e3:.0f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:7])
        print(f"frame {done:5d}: {ms / 20 * 1e3:7.1f} us/frame | cells occ {occ.size} mean {occ.mean():.2f} max {occ.max()} sum cnt^2/N {float((occ.astype(np.float64) ** 2).sum()) / ok.sum():.1f} "
              f"| nan {int((~ok).sum())} y [{np.nanmin(p['pos'][:, 1]):.3f},{np.nanmax(p['pos'][:, 1]):.3f}] above grid {(p['pos'][ok, 1] > bench.GRID_MAX[1]).sum()} | {top}", flush=True)
