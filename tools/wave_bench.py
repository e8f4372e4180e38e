"""Wave stencil alone (C3 and the C4 field size): us per step and fraction of the HBM peak (12 B per cell).
usage: python tools/wave_bench.py [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coupledwateranimation_b200 as cwa  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6530.0
with cwa.Context(0) as ctx:
    for n, variant in ((2048, cwa.WAVE_COUPLED), (4096, cwa.WAVE_SIMP), (8192, cwa.WAVE_SIMP)):
        w = cwa.StencilImage2DTripleBuffered(ctx, n, n, 1, variant)
        w.Compute(20)
        ctx.synchronize()
        ctx.timer_begin()
        w.Compute(steps)
        ms = ctx.timer_end()
        us = ms / steps * 1e3
        gbs = n * n * 12 / (us * 1e-6) / 1e9
        print(f"{n}^2: {us:8.2f} us/step  {gbs:7.0f} GB/s  {100 * gbs / peak:5.1f}% of {peak:.0f} GB/s  {n * n / (us * 1e-6) / 1e9:.1f} Gcell/s", flush=True)
        w.destroy()
