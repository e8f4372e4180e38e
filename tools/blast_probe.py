"""Diagnostic: frame cost, crowded cells, NaN count and |v|max of the first 40 frames of the `--gpus N` scene, run on ONE GPU.
usage: python tools/blast_probe.py N"""
import math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import coupledwateranimation_b200 as cwa

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc = bench.scaled_scene(world)
with cwa.Context(0) as ctx:
    ctx.set_boundary(upper=(sc["box_x"], 1.0, sc["box_z"], 500.0), lower=bench.BOX_LOWER)
    ctx.set_sim_constants(uv_scale=sc["uv"], uv_scale_z=sc["uv_z"], torque_coeff=sc["torque"])
    n = sc["nx"] * sc["ny"] * sc["nz"]
    grid = cwa.UniformGrid(ctx, 3, sc["gmin"], sc["gmax"], sc["gn"], n, compact_index=True)
    sph = cwa.Sph(ctx, n, grid)
    sph.init_cube(sc["nx"], sc["ny"], sc["nz"])
    wave = cwa.StencilImage2DTripleBuffered(ctx, sc["wave_w"], sc["wave_h"], 1, cwa.WAVE_COUPLED)
    gn = sc["gn"]
    for fr in range(1, 41):
        ctx.timer_begin()
        sph.coupled_step(wave, 1, bench.COUPLING)
        ms = ctx.timer_end()
        if fr % 2 == 0 or ms > 1.0:
            cnt = grid.read(cwa.GRID_COUNTER, grid.num_cells_total)
            top = np.argsort(cnt)[-3:][::-1]
            desc = []
            for c in top:
                ks = gn[2]; k = c % ks; ij = c // ks; j = ij % gn[1]; i = ij // gn[1]       # compact index: k-stride = Nz
                desc.append(f"({i},{j},{k}):{int(cnt[c])}")
            p = sph.download()
            z = p["pos"][:, 2]; y = p["pos"][:, 1]
            print(f"frame {fr}: {ms*1e3:.0f} us, max/cell {int(cnt.max())}, cells>64 {int((cnt>64).sum())}, top {' '.join(desc)}, "
                  f"y[{np.nanmin(y):.3f},{np.nanmax(y):.3f}] nan {int(np.isnan(z).sum())} vmax {np.nanmax(np.abs(p['vel'][:, :3])):.0f}", flush=True)
