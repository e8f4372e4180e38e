"""Short C4 run for ncu: build the bench scene and run a few coupled frames (no timing claims)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import coupledwateranimation_b200 as cwa  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with cwa.Context(0) as ctx:
    grid, sph, wave = bench.build_scene(cwa, ctx)
    sph.coupled_step(wave, frames, bench.COUPLING)
    ctx.synchronize()
    print("frames", frames, "launches", ctx.launch_count)
