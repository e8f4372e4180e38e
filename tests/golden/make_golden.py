"""Generate tests/golden/golden_v1.npz from the CPU oracle (seeded, deterministic inputs).

    python tests/golden/make_golden.py

The reference holds no golden vectors for the SPH passes / wave stencil and its GLSL cannot run
here, so these fixtures pin the ORACLE (regression guard) and give the GPU tests a
committed target that does not depend on rebuilding the oracle on the GPU box.  Scan vectors are the
reference's own KATs (SphWave2D/ParallelScan.cpp:129-138, UniformGrid2D/ParallelScan.cpp:117); the
2-D grid vectors were additionally cross-checked against the reference's CPU twin (oracle/_ref).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")

SMALL_GRID3 = ((0.0, -0.02, 0.0), (0.12, 0.1, 0.12), (6, 6, 6))


def small_params():
    prm = O.default_params3()
    prm.upper[0] = prm.upper[2] = 0.11
    return prm


def small_block(prm, nx=12, ny=5, nz=12, seed=1234, vel=0.3):
    p = O.make_cube(nx, ny, nz, prm)
    rng = np.random.default_rng(seed)
    p["pos"][:, :3] += rng.uniform(-0.1, 0.1, (p.size, 3)).astype(np.float32) * np.float32(0.0085)
    p["vel"][:, :3] = rng.uniform(-vel, vel, (p.size, 3)).astype(np.float32)
    p["force"] = rng.uniform(-1e4, 1e4, (p.size, 4)).astype(np.float32)
    p["pos"][::7, 1] += np.float32(0.02)
    return p


def smooth_field(h, w, amp=0.02):
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    return np.ascontiguousarray(amp * (np.sin(x * 0.37) * np.cos(y * 0.23) + 0.5 * np.sin((x + y) * 0.11)).astype(np.float32))


def main():
    g = {}
    # ---- scan KATs (reference vectors) ------------------------------------------------------------
    g["scan_kat1_in"] = np.array([1, 0, 1, 0, 1, 2, 1, 2, 1, 2, 0, 1, 0, 2, 1, 0], np.int32)
    g["scan_kat1_out"] = np.array([0, 1, 1, 2, 2, 3, 5, 6, 8, 9, 11, 11, 12, 12, 14, 15], np.int32)
    g["scan_kat2_in"] = np.ones(16, np.int32)
    g["scan_kat2_out"] = np.arange(16, dtype=np.int32)
    rng = np.random.default_rng(2024)
    g["scan_rand_in"] = rng.integers(0, 9, 5000, dtype=np.int32)
    g["scan_rand_out"] = O.scan_exclusive(g["scan_rand_in"])

    # ---- grids ---------------------------------------------------------------------------------------
    xy = rng.uniform(-0.2, 9.8, (512, 2)).astype(np.float32)
    rows = np.zeros((512, 12), np.float32); rows[:, :2] = xy
    gg = O.grid2((0.0, 0.0), (9.6, 9.6), (32, 32))
    cell, cnt, off, idx = O.grid2_build(gg, rows)
    g["grid2_pos"], g["grid2_cell"], g["grid2_cnt"], g["grid2_off"], g["grid2_idx"] = xy, cell, cnt, off, idx
    prm = small_params()
    p = small_block(prm)
    p["pos"][3, 0] = np.nan
    g3 = O.grid3(*SMALL_GRID3)
    cell, cnt, off, idx = O.grid3_build(g3, p["pos"])
    g["grid3_pos"], g["grid3_cell"], g["grid3_cnt"], g["grid3_off"], g["grid3_idx"] = p["pos"].copy(), cell, cnt, off, idx

    # ---- wave ------------------------------------------------------------------------------------------
    for wtype in (1.0, 0.0, 0.5):
        g[f"wave_init_t{wtype}"] = O.wave_init(32, 32, 1, O.WAVE_COUPLED, wtype)
    g["wave_init_simp"] = O.wave_init(48, 32, 1, O.WAVE_SIMP, 1.0)
    u0 = (0.5 * rng.standard_normal((40, 52))).astype(np.float32)
    u1 = (0.5 * rng.standard_normal((40, 52))).astype(np.float32)
    g["wave_u0"], g["wave_u1"] = u0, u1
    g["wave_step_coupled"] = O.wave_evolve(u0, u1, O.WAVE_COUPLED, 0.01, 0.985, 0.001, 1.0)
    g["wave_step_wake"] = O.wave_evolve(np.abs(u0), 0.1 * np.abs(u1), O.WAVE_COUPLED, 0.01, 0.985, 0.001, 0.5)
    g["wave_step_simp"] = O.wave_evolve(u0, u1, O.WAVE_SIMP, 0.01, 0.9995, 0.001, 1.0)
    a = O.wave_init(32, 32, 1, O.WAVE_COUPLED, 1.0); b = a.copy()
    for _ in range(25):
        a, b = O.wave_evolve(a, b, O.WAVE_COUPLED, 0.01, 0.985, 0.001, 1.0), a
    g["wave_25_steps"] = a

    # ---- sampler ---------------------------------------------------------------------------------------
    tex = rng.standard_normal((16, 24)).astype(np.float32)
    st = rng.uniform(-0.3, 1.6, (64, 2)).astype(np.float32)
    g["tex"], g["tex_st"] = tex, st
    g["tex_val"] = np.array([O.tex_bilinear(tex, float(s), float(t)) for s, t in st], np.float32)

    # ---- 3-D SPH passes (all-pairs, as shipped) ----------------------------------------------------
    p = small_block(prm)
    field = smooth_field(32, 32)
    g["sph3_in"], g["sph3_tex"] = p.copy(), field
    q = p.copy(); O.sph3_rho_pres(q, prm, field); g["sph3_after_rho"] = q.copy()
    O.sph3_force(q, prm, field); g["sph3_after_force"] = q.copy()
    O.sph3_integrate(q, prm, field); g["sph3_after_integrate"] = q.copy()
    g["sph3_neighbours"] = O.sph3_neighbour_count(p, 0.01)

    # ---- coupled frames -----------------------------------------------------------------------------
    for name, mode in (("as_shipped", O.COUPLING_AS_SHIPPED), ("latest", O.COUPLING_LATEST)):
        oc = O.Coupled(p.size, 32, 32, 1, prm, mode)
        start = small_block(prm, vel=0.1)
        oc.particles[:] = start
        oc.step(6)
        g[f"coupled_{name}_particles"] = oc.particles.copy()
        g[f"coupled_{name}_wave"] = oc.wave(0).copy()
        oc.close()
    g["coupled_start"] = start

    # ---- 2-D Koschier ---------------------------------------------------------------------------------
    for variant in (0, 1):
        prm2 = O.default_params2(variant)
        p2 = O.sph2_init(1024, prm2)
        rng2 = np.random.default_rng(77 + variant)
        p2["pos"][:, :2] += rng2.uniform(-0.004, 0.004, (1024, 2)).astype(np.float32)
        p2["vel"][:, :2] = rng2.uniform(-0.3, 0.3, (1024, 2)).astype(np.float32)
        g[f"sph2_v{variant}_in"] = p2.copy()
        b0, b1 = p2.copy(), np.zeros_like(p2)
        r, _ = O.sph2_step(b0, b1, 0, 2, prm2, None, O.grid2((0.0, 0.0), (9.6, 9.6), (32, 32)))
        g[f"sph2_v{variant}_out"] = (b0, b1)[r].copy()

    np.savez_compressed(OUT, **g)
    print(f"wrote {OUT}: {len(g)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
