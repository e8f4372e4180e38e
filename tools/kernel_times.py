"""Per-kernel device times of the C4 coupled frame (CUDA events around every launch).

usage: python tools/kernel_times.py [warm_frames] [measured_frames] [--stats]
Used for tuning (e.g. CWA_NB_CONFIG=1 python tools/kernel_times.py 10 50); bench.py reports the same
numbers in its roofline_kernels section.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import coupledwateranimation_b200 as cwa  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
warm = int(args[0]) if len(args) > 0 else 10
meas = int(args[1]) if len(args) > 1 else 50
with cwa.Context(0) as ctx:
    grid, sph, wave = bench.build_scene(cwa, ctx)
    sph.coupled_step(wave, warm, bench.COUPLING)
    ctx.synchronize()
    ctx.timer_begin()
    sph.coupled_step(wave, meas, bench.COUPLING)
    ms = ctx.timer_end()
    ctx.profile_begin()
    sph.coupled_step(wave, meas, bench.COUPLING)
    prof = ctx.profile_end()
    tot = sum(v[0] for v in prof.values())
    print(f"cfg={os.environ.get('CWA_NB_CONFIG', '0')} warm={warm} meas={meas}: {ms / meas * 1e3:.1f} us/frame (events per kernel sum {tot / meas * 1e3:.1f} us)")
    for name, (t, c) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        print(f"   {name:18s} {t / c * 1e3:9.1f} us x{c // meas}/frame  {100 * t / tot:5.1f}%")
    if "--stats" in sys.argv:
        cnt = grid.read(cwa.GRID_COUNTER, grid.num_cells_total)
        occ = cnt[cnt > 0]
        nb = sph.neighbour_count()
        p = sph.download()
        print(f"   cells occupied {occ.size}, particles/cell mean {occ.mean():.1f} max {occ.max()}, neighbours mean {nb.mean():.1f} max {nb.max()},"
              f" nan {int(np.isnan(p['pos']).any(1).sum())}, y range [{np.nanmin(p['pos'][:, 1]):.4f}, {np.nanmax(p['pos'][:, 1]):.4f}]")
