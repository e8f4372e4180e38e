"""Generate tests/golden/golden_v2_stencil1d.npz: the 1-D wave substrates from the CPU oracle.

    python tests/golden/make_golden_stencil1d.py

The reference holds no vectors for Shallow1D_cs / Wave1D_cs either; these fixtures freeze the oracle's restatement (initial
profiles, one frame, many frames, the Splash mode, every boundary condition) and give the GPU tests a committed target."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v2_stencil1d.npz")


def snapshot(g, key, s):
    g[key + "_images"] = np.stack(s.image)
    g[key + "_state"] = np.array(list(s.unit) + list(s.read_index) + [s.write_index], np.int32)


def main():
    g = {}
    for shader, name, width, frames in ((O.STENCIL1D_SHALLOW, "shallow", 128, 40), (O.STENCIL1D_WAVE, "wave", 1024, 5)):
        for bc in (O.BC_REFLECT, O.BC_FREE, O.BC_FIXED):
            s = O.ImageStencil(shader, width)
            s.prm.bc = bc; s.prm.boundary[0] = 0.3; s.prm.boundary[1] = -0.2
            if bc == O.BC_FREE:
                snapshot(g, f"{name}_init", s)
            s.compute(1)
            snapshot(g, f"{name}_bc{bc}_1frame", s)
            s.compute(frames - 1)
            snapshot(g, f"{name}_bc{bc}_{frames}frames", s)
    s = O.ImageStencil(O.STENCIL1D_SHALLOW, 128)
    s.compute(10)
    s.compute_func(1)                                   # Splash
    snapshot(g, "shallow_splash", s)
    np.savez_compressed(OUT, **g)
    print(f"wrote {OUT}: {len(g)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
