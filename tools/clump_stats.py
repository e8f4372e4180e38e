"""Diagnostic: how many targets does the clump path see, and how big are they?  C4 at a late frame (default 3000).
Candidates of a target = particles in the 3 x 3 x 3 cells around it (cells of ~h)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import coupledwateranimation_b200 as cwa

frame = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
with cwa.Context(0) as ctx:
    grid, sph, wave = bench.build_scene(cwa, ctx)
    sph.coupled_step(wave, frame, bench.COUPLING)
    ctx.profile_begin(); sph.coupled_step(wave, 20, bench.COUPLING); prof = ctx.profile_end()
    print({k: round(v[0] / v[1] * 1e3, 1) for k, v in prof.items()})
    nx, ny, nz = bench.GRID_N
    cnt = grid.read(cwa.GRID_COUNTER, grid.num_cells_total).reshape(nx, ny, nx)[:, :, :nz].astype(np.int64)   # index (i*Ny + j)*Nx + k
    box = cnt.copy()
    for ax in range(3):
        p = np.pad(box, [(1, 1) if a == ax else (0, 0) for a in range(3)])
        sl = lambda o: tuple(slice(o, o + box.shape[a]) if a == ax else slice(None) for a in range(3))
        box = p[sl(0)] + p[sl(1)] + p[sl(2)]
    cand = np.repeat(box.ravel(), cnt.ravel())          # candidates of every inserted particle
    nb = sph.neighbour_count()
    print(f"frame {frame + 20}: particles {cand.size}, candidates mean {cand.mean():.1f} max {cand.max()}")
    for thr in (192, 384, 768, 1536, 3072):
        m = cand > thr
        print(f"  candidates > {thr:5d}: {int(m.sum()):7d} targets, {cand[m].sum() / 1e6:8.2f} M candidate tests ({100 * cand[m].sum() / cand.sum():.1f} % of all)")
    for thr in (64, 128, 256):
        print(f"  neighbours > {thr:4d}: {int((nb > thr).sum()):7d} targets")
