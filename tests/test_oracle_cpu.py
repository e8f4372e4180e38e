"""CPU tests: the oracle against the reference's own known-answer vectors, the reference's CPU grid
twin (oracle/_ref, compiled from the reference sources) and the committed golden fixtures."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
SMALL_GRID3 = ((0.0, -0.02, 0.0), (0.12, 0.1, 0.12), (6, 6, 6))


def small_params(O):
    prm = O.default_params3()
    prm.upper[0] = prm.upper[2] = 0.11
    return prm


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


# ---- scan: the reference's KATs ------------------------------------------------------------------
def test_scan_reference_kats(oracle):
    # ParallelScanTest (SphWave2D/ParallelScan.cpp:124-155) and the all-ones variant (UniformGrid2D/ParallelScan.cpp:117)
    for k in ("kat1", "kat2"):
        x, want = GOLD[f"scan_{k}_in"], GOLD[f"scan_{k}_out"]
        assert np.array_equal(oracle.scan_blelloch(x), want)     # restated multi-dispatch Blelloch sweep
        assert np.array_equal(oracle.scan_exclusive(x), want)    # the test's own "truth" loop


def test_scan_blelloch_requires_power_of_two_like_the_reference_assert(oracle):
    with pytest.raises(ValueError):
        oracle.scan_blelloch(np.ones(12, np.int32))               # ParallelScan.cpp:15-16


@pytest.mark.parametrize("n", [2, 64, 1024, 1 << 15])
def test_scan_blelloch_equals_sequential(oracle, n):
    x = np.random.default_rng(n).integers(0, 100, n, dtype=np.int32)
    assert np.array_equal(oracle.scan_blelloch(x), oracle.scan_exclusive(x))


# ---- grid: pinned by the reference's CPU twin -------------------------------------------------------
def test_grid2d_matches_reference_cpu_twin(oracle):
    rng = np.random.default_rng(5)
    for n, nc in ((1, (4, 4)), (777, (32, 32)), (4096, (32, 32)), (3000, (7, 19))):
        xy = rng.uniform(0.001, 9.599, (n, 2)).astype(np.float32)
        ref = oracle.ref_grid2d_build(xy, (0.0, 0.0), (9.6, 9.6), nc)
        if ref is None:
            pytest.skip("oracle/_ref not built (reference tree absent)")
        cnt, off, idx, cs = ref
        g = oracle.grid2((0.0, 0.0), (9.6, 9.6), nc)
        rows = np.zeros((n, 12), np.float32); rows[:, :2] = xy
        cell_of, ocnt, ooff, oidx = oracle.grid2_build(g, rows)
        assert np.array_equal(ocnt, cnt) and np.array_equal(ooff, off) and np.array_equal(oidx, idx)
        assert np.float32(g.cell[0]) == cs[0] and np.float32(g.cell[1]) == cs[1]


def test_grid2d_particles_on_cell_faces_match_reference_twin(oracle):
    xs = (np.arange(1, 32, dtype=np.float32) * np.float32(0.3))
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    xy = np.stack([X.ravel(), Y.ravel()], 1).astype(np.float32)
    ref = oracle.ref_grid2d_build(xy, (0.0, 0.0), (9.6, 9.6), (32, 32))
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rows = np.zeros((xy.shape[0], 12), np.float32); rows[:, :2] = xy
    _, cnt, off, idx = oracle.grid2_build(oracle.grid2((0.0, 0.0), (9.6, 9.6), (32, 32)), rows)
    assert np.array_equal(cnt, ref[0]) and np.array_equal(off, ref[1]) and np.array_equal(idx, ref[2])


def test_grid_goldens(oracle):
    rows = np.zeros((512, 12), np.float32); rows[:, :2] = GOLD["grid2_pos"]
    cell, cnt, off, idx = oracle.grid2_build(oracle.grid2((0.0, 0.0), (9.6, 9.6), (32, 32)), rows)
    for got, k in ((cell, "cell"), (cnt, "cnt"), (off, "off"), (idx, "idx")):
        assert np.array_equal(got, GOLD[f"grid2_{k}"]), k
    assert (cell == -1).sum() > 0, "fixture holds out-of-extent particles (strict point_in_aabb)"
    cell, cnt, off, idx = oracle.grid3_build(oracle.grid3(*SMALL_GRID3), GOLD["grid3_pos"])
    for got, k in ((cell, "cell"), (cnt, "cnt"), (off, "off"), (idx, "idx")):
        assert np.array_equal(got, GOLD[f"grid3_{k}"]), k
    assert cell[3] == -1, "NaN position: canonical choice = not inserted"


# ---- wave -------------------------------------------------------------------------------------------
def test_wave_goldens(oracle):
    for wtype in (1.0, 0.0, 0.5):
        assert np.array_equal(oracle.wave_init(32, 32, 1, oracle.WAVE_COUPLED, wtype), GOLD[f"wave_init_t{wtype}"])
    assert np.array_equal(oracle.wave_init(48, 32, 1, oracle.WAVE_SIMP, 1.0), GOLD["wave_init_simp"])
    u0, u1 = GOLD["wave_u0"], GOLD["wave_u1"]
    assert np.array_equal(oracle.wave_evolve(u0, u1, 0, 0.01, 0.985, 0.001, 1.0), GOLD["wave_step_coupled"])
    assert np.array_equal(oracle.wave_evolve(np.abs(u0), 0.1 * np.abs(u1), 0, 0.01, 0.985, 0.001, 0.5), GOLD["wave_step_wake"])
    assert np.array_equal(oracle.wave_evolve(u0, u1, 1, 0.01, 0.9995, 0.001, 1.0), GOLD["wave_step_simp"])


def test_wave_init_shape_matches_the_shader_text(oracle):
    # splash: peak 0.5 at ivec2(0.25*size), ivec2(0.75*size); smoothstep(5,0,d) is 0 from d >= 5 (wave_comp.glsl:91-98,138)
    f = oracle.wave_init(64, 64, 1, oracle.WAVE_COUPLED, 1.0)
    assert f[16, 16] == np.float32(0.5) and f[48, 48] == np.float32(0.5)
    assert f[16, 21] == 0.0 and f[16, 20] > 0.0 and f[0, 0] == 0.0
    rgba = oracle.wave_init(64, 64, 4, oracle.WAVE_COUPLED, 1.0)
    assert np.array_equal(rgba[..., 0], f) and not rgba[..., 1:].any()          # only .x is ever non-zero (F9)


def test_wave_rgba_equals_scalar_per_channel(oracle):
    rng = np.random.default_rng(3)
    u0 = rng.standard_normal((20, 28, 4)).astype(np.float32); u1 = rng.standard_normal((20, 28, 4)).astype(np.float32)
    out = oracle.wave_evolve(u0, u1, 0, 0.01, 0.985, 0.001, 1.0)
    for c in range(4):
        assert np.array_equal(out[..., c], oracle.wave_evolve(np.ascontiguousarray(u0[..., c]), np.ascontiguousarray(u1[..., c]), 0, 0.01, 0.985, 0.001, 1.0))


def test_wave_25_steps_golden_and_decay(oracle):
    a = oracle.wave_init(32, 32, 1, 0, 1.0); b = a.copy()
    e0 = float((a.astype(np.float64) ** 2).sum())
    for _ in range(25):
        a, b = oracle.wave_evolve(a, b, 0, 0.01, 0.985, 0.001, 1.0), a
    assert np.array_equal(a, GOLD["wave_25_steps"])
    assert float((a.astype(np.float64) ** 2).sum()) < e0          # atten = 0.985 damps the field


# ---- sampler (Appendix A.3) ----------------------------------------------------------------------
def test_bilinear_sampler_semantics(oracle):
    tex = GOLD["tex"]
    H, W = tex.shape
    got = np.array([oracle.tex_bilinear(tex, float(s), float(t)) for s, t in GOLD["tex_st"]], np.float32)
    assert np.array_equal(got, GOLD["tex_val"])
    # texel centres reproduce the texel; CLAMP_TO_EDGE outside [0,1]
    assert oracle.tex_bilinear(tex, (5 + 0.5) / W, (7 + 0.5) / H) == tex[7, 5]
    assert oracle.tex_bilinear(tex, -3.0, 0.5 / H) == tex[0, 0]
    assert oracle.tex_bilinear(tex, 2.0, 9.0) == tex[H - 1, W - 1]
    assert oracle.tex_bilinear(None, 0.3, 0.3) == 0.0             # unbound texture samples 0 (Appendix B, frame 1)
    mid = oracle.tex_bilinear(tex, 6.0 / W, 7.5 / H)               # halfway between texels 5 and 6 of row 7
    assert abs(mid - 0.5 * (float(tex[7, 5]) + float(tex[7, 6]))) < 1e-6


# ---- 3-D SPH ------------------------------------------------------------------------------------------
def test_sph3_pass_goldens(oracle):
    prm = small_params(oracle)
    q = GOLD["sph3_in"].copy(); tex = GOLD["sph3_tex"]
    oracle.sph3_rho_pres(q, prm, tex); assert np.array_equal(bits(q), bits(GOLD["sph3_after_rho"]))
    oracle.sph3_force(q, prm, tex); assert np.array_equal(bits(q), bits(GOLD["sph3_after_force"]))
    oracle.sph3_integrate(q, prm, tex); assert np.array_equal(bits(q), bits(GOLD["sph3_after_integrate"]))
    assert np.array_equal(oracle.sph3_neighbour_count(GOLD["sph3_in"].copy(), 0.01), GOLD["sph3_neighbours"])


def test_sph3_known_lattice_values(oracle):
    """self term 31 333.6, face neighbour (r = 0.0085) 669.6, interior rho ~ 35 351 (FP32)."""
    prm = oracle.default_params3()
    p = oracle.make_cube(8, 5, 8, prm)
    oracle.sph3_rho_pres(p, prm, None)
    rho = p["extras"][:, 0].reshape(8, 5, 8)
    assert abs(rho[4, 2, 4] - 35351.2) < 1.0
    single = oracle.make_cube(1, 1, 1, prm)
    oracle.sph3_rho_pres(single, prm, None)
    assert abs(single["extras"][0, 0] - 31333.6) < 0.5
    assert abs((rho[4, 2, 4] - single["extras"][0, 0]) / 6.0 - 669.6) < 0.1
    assert abs(p["extras"][:, 1].reshape(8, 5, 8)[4, 2, 4] - 4000.0 * (rho[4, 2, 4] - 1000.0)) < 64.0


def test_sph3_grid_and_all_pairs_agree(oracle):
    prm = small_params(oracle)
    p = GOLD["sph3_in"].copy(); tex = GOLD["sph3_tex"]
    g = oracle.grid3(*SMALL_GRID3)
    _, cnt, off, idx = oracle.grid3_build(g, p["pos"])
    grid = (g, cnt, off, idx)
    assert np.array_equal(oracle.sph3_neighbour_count(p, 0.01, grid), oracle.sph3_neighbour_count(p, 0.01)), "identical neighbour sets"
    a, b = p.copy(), p.copy()
    oracle.sph3_rho_pres(a, prm, tex); oracle.sph3_rho_pres(b, prm, tex, grid)
    assert np.allclose(a["extras"], b["extras"], rtol=1e-5)
    oracle.sph3_force(a, prm, tex); oracle.sph3_force(b, prm, tex, grid)
    scale = np.sqrt(np.mean(a["force"][:, :3].astype(np.float64) ** 2))
    assert np.abs(a["force"][:, :3] - b["force"][:, :3]).max() <= 1e-4 * scale


def test_integrate_rules(oracle):
    prm = oracle.default_params3()
    p = oracle.make_cube(2, 1, 1, prm)
    p["pos"][0, :3] = (0.1, 0.5, 0.1); p["vel"][0, :3] = (0.0, 600000.0, 0.0)      # foam: |v| > 25 -> v *= 0.1, rho *= 0.1, p *= 0.25
    p["pos"][1, :3] = (0.0, 0.2, 0.47999); p["vel"][1, :3] = (-10.0, 0.0, 10.0)    # walls: x < lower, z > upper
    p["extras"][:, 1] = 8.0
    p["force"][:] = 1.0
    oracle.sph3_integrate(p, prm, None)
    assert p["extras"][0, 0] == np.float32(100.0) and p["extras"][0, 1] == np.float32(2.0) and p["force"][0, 3] == np.float32(0.5)
    assert p["pos"][0, 1] == np.float32(1.0) and p["vel"][0, 1] < 0                 # clamped to upper.y, v.y *= -0.3
    assert p["pos"][1, 0] == np.float32(0.0) and p["vel"][1, 0] > 0 and p["pos"][1, 2] == np.float32(0.48) and p["vel"][1, 2] < 0
    below = oracle.make_cube(1, 1, 1, prm)
    below["pos"][0, :3] = (0.1, 0.01, 0.1)
    tex = np.full((8, 8), 0.2, np.float32)
    oracle.sph3_integrate(below, prm, tex)
    assert below["pos"][0, 1] == np.float32(np.float32(0.2) - np.float32(0.005))   # y = tex_height - PARTICLE_RADIUS


def test_make_cube_layout(oracle):
    p = oracle.make_cube(64, 5, 64)
    assert p.size == 20480                                                           # NUM_PARTICLES (Main.cpp:31)
    sp = np.float32(np.float32(2.0 * 0.85) * np.float32(0.005))
    assert p["pos"][1, 2] == sp and p["pos"][64, 1] == sp and p["pos"][320, 0] == sp  # k inner, then j, i outer
    assert (p["extras"] == np.array([1000.0, 0.0, 500.0, 50.0], np.float32)).all() and (p["pos"][:, 3] == 1).all()


# ---- coupled driver ------------------------------------------------------------------------------------
def test_as_shipped_texture_schedule(oracle):
    """Appendix B: frame 1 samples an unbound texture, then the image on unit 0 is refreshed every 3 frames."""
    prm = small_params(oracle)
    oc = oracle.Coupled(8, 16, 16, 1, prm, oracle.COUPLING_AS_SHIPPED)
    oc.particles[:] = oracle.make_cube(2, 2, 2, prm)
    sampled, contents = [], []
    levels = {}
    for frame in range(1, 9):
        img = oc.sampled_image()
        sampled.append(img)
        contents.append(None if img < 0 else oc._h and oc.wave(0).copy() * 0)   # placeholder, content checked below
        oc.step(1)
        levels[frame] = oc.wave(0).copy()
    assert sampled == [-1, 0, 0, 0, 0, 0, 0, 0]
    oc.close()


def test_coupled_goldens(oracle):
    prm = small_params(oracle)
    for name, mode in (("as_shipped", oracle.COUPLING_AS_SHIPPED), ("latest", oracle.COUPLING_LATEST)):
        oc = oracle.Coupled(GOLD["coupled_start"].size, 32, 32, 1, prm, mode)
        oc.particles[:] = GOLD["coupled_start"]
        oc.step(6)
        assert np.array_equal(bits(oc.particles), bits(GOLD[f"coupled_{name}_particles"])), name
        assert np.array_equal(oc.wave(0), GOLD[f"coupled_{name}_wave"]), name
        oc.close()
    assert not np.array_equal(bits(GOLD["coupled_as_shipped_particles"]), bits(GOLD["coupled_latest_particles"]))


def test_default_scene_frame_two_nans_are_the_reference_behaviour(oracle):
    """Config D as shipped: 7 columns/rows of the lattice overhang upper.xz = 0.48, get clamped onto the wall in
    frame 1, coincide, and normalize(0) turns them into NaN in frame 2."""
    prm = oracle.default_params3()
    oc = oracle.Coupled(20480, 64, 64, 4, prm, oracle.COUPLING_AS_SHIPPED)
    oc.particles[:] = oracle.make_cube(64, 5, 64, prm)
    oc.step(1)
    P = oc.particles
    assert not np.isnan(P["pos"]).any()
    assert int((P["pos"][:, 0] == np.float32(0.48)).sum()) == 7 * 5 * 64
    oc.step(1)
    assert int(np.isnan(oc.particles["pos"][:, :3]).any(1).sum()) > 1000
    oc.close()


# ---- 2-D Koschier ---------------------------------------------------------------------------------------
def test_sph2_goldens_and_constants(oracle):
    for variant in (0, 1):
        prm = oracle.default_params2(variant)
        p = GOLD[f"sph2_v{variant}_in"].copy()
        b0, b1 = p.copy(), np.zeros_like(p)
        r, _ = oracle.sph2_step(b0, b1, 0, 2, prm, None, oracle.grid2((0.0, 0.0), (9.6, 9.6), (32, 32)))
        assert np.array_equal(bits((b0, b1)[r]), bits(GOLD[f"sph2_v{variant}_out"])), variant
    # InitParticle of the wave variant: cols = 128, p = (9.6, 1.8) * ij/(cols, rows) + (R, 4.8 - 15R + R)
    prm = oracle.default_params2(1)
    p = oracle.sph2_init(4096, prm)
    assert p["pos"][0, 0] == np.float32(0.025) and abs(p["pos"][0, 1] - 4.45) < 1e-6
    assert abs(p["pos"][1, 0] - (0.025 + 9.6 / 128)) < 1e-6 and abs(p["pos"][128, 1] - (4.45 + 1.8 / 32)) < 1e-6
    assert (p["acc"][:, 3] == 1000.0).all()


# ---- 1-D wave substrates -----------------------------------------------------------
def test_image_stencil_bookkeeping_as_written(oracle):
    """ImageStencil::PingPong (StencilImage2D.cpp:67-83) rotates the UNITS cyclically (output -> 0 -> 1 -> output) while the
    index arrays follow a different permutation for three buffers -- reproduced as written."""
    s = oracle.ImageStencil(oracle.STENCIL1D_WAVE, 64)
    assert (s.unit, s.read_index, s.write_index) == ([2, 0, 1], [2, 0], 1)          # after Init() = Reinit(): two dispatches
    for _ in range(9):
        before = list(s.unit)
        s.pingpong()
        assert s.unit == [0 if u == 2 else u + 1 for u in before]
    d = oracle.ImageStencil(oracle.STENCIL1D_SHALLOW, 16)
    assert (d.unit, d.read_index, d.write_index) == ([1, 0], [1], 0)


def test_shallow1d_init_profile_and_conservation(oracle):
    s = oracle.ImageStencil(oracle.STENCIL1D_SHALLOW, 128)
    im = s.read_image(0)
    x = np.arange(128, dtype=np.float32) / np.float32(127)
    h = 9.6 * 0.1 * np.exp(-((x - 0.5) ** 2) / 0.005)
    assert np.allclose(im[:, 0], 4.8 + h, rtol=1e-6) and np.allclose(im[:, 1], 0.5 * h * np.sign(x - 0.5), rtol=1e-5, atol=1e-7)
    assert not im[:, 2:].any()
    m0 = im[1:-1, 0].astype(np.float64).sum()
    s.compute(500)
    r = s.read_image(0)
    assert np.isfinite(r[:, :2]).all() and abs(r[1:-1, 0].astype(np.float64).sum() - m0) < 1e-3 * m0
    # free boundary: the end texels are copies of their neighbours (FreeBC, Shallow1D_cs.glsl:176-193)
    assert np.array_equal(r[0], r[1]) and np.array_equal(r[-1], r[-2])


def test_shallow1d_half_step_store_wins_over_boundary_copies(oracle):
    """ITERATE0 ends with an unconditional imageStore of the half-step values (:160): texel 0 holds its OWN (hm, uhm), not the copy
    of texel 1 that FreeBC wrote earlier, and texel w-1 the values computed against the zero read beyond the image (uhm = NaN)."""
    prm = oracle.default_stencil1d_params(oracle.STENCIL1D_SHALLOW)
    rng = np.random.default_rng(7)
    inp = np.zeros((16, 4), np.float32)
    inp[:, 0] = rng.uniform(4.5, 5.5, 16); inp[:, 1] = rng.uniform(-0.5, 0.5, 16)
    out = np.full((16, 4), 7.0, np.float32)
    oracle.shallow1d_dispatch(inp, out, 2, prm)
    hm0 = np.float32((inp[0, 0] + inp[1, 0]) / np.float32(2)) - np.float32(np.float32(0.001) / np.float32(2) * (inp[1, 1] - inp[0, 1])) / np.float32(0.1)
    assert out[0, 2] == np.float32(hm0) and out[0, 0] == inp[0, 0]
    assert np.isnan(out[15, 3]) and np.isfinite(out[15, 2])
    out2 = np.full((16, 4), 7.0, np.float32)
    oracle.shallow1d_dispatch(out, out2, 3, prm)
    assert np.isfinite(out2).all() and np.array_equal(out2[15], out2[14]) and np.array_equal(out2[0], out2[1])


def test_wave1d_iterate_matches_the_formula_and_damps(oracle):
    prm = oracle.default_stencil1d_params(oracle.STENCIL1D_WAVE)
    rng = np.random.default_rng(9)
    u0 = np.zeros((32, 4), np.float32); u1 = np.zeros((32, 4), np.float32)
    u0[:, 0] = rng.uniform(-0.1, 0.1, 32); u1[:, 0] = rng.uniform(-0.1, 0.1, 32)
    out = np.zeros((32, 4), np.float32)
    oracle.wave1d_dispatch(u0, u1, out, 2, prm)
    lam, att, beta = 0.01, 0.9995, 0.001
    ref = att * ((2 - 2 * lam - beta) * u0[1:-1, 0].astype(np.float64) + lam * (u0[2:, 0].astype(np.float64) + u0[:-2, 0]) - (1 - beta) * u1[1:-1, 0])
    assert np.allclose(out[1:-1, 0], ref, rtol=1e-5, atol=1e-7)
    assert np.allclose(out[1:-1, 1], (out[1:-1, 0] - u1[1:-1, 0]) / 2, rtol=1e-5, atol=1e-7)
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[-1], out[-2])
    w = oracle.ImageStencil(oracle.STENCIL1D_WAVE, 1024)
    e0 = float(np.abs(w.image[w._with_unit(0)][:, 0]).max())
    w.compute(100)                                    # 1000 damped steps
    assert float(np.abs(w.image[w._with_unit(0)][:, 0]).max()) < e0


def _stencil_from_golden(oracle, gold, key, shader, width, bc):
    s = oracle.ImageStencil(shader, width)
    s.prm.bc = bc; s.prm.boundary[0] = 0.3; s.prm.boundary[1] = -0.2
    n = s.num_images
    for i in range(n):
        s.image[i][:] = gold[key + "_images"][i]
    st = gold[key + "_state"].tolist()
    s.unit, s.read_index, s.write_index = st[:n], st[n:2 * n - 1], st[-1]
    return s


def test_stencil1d_goldens(oracle):
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2_stencil1d.npz"))
    for shader, name, width, frames in ((oracle.STENCIL1D_SHALLOW, "shallow", 128, 40), (oracle.STENCIL1D_WAVE, "wave", 1024, 5)):
        fresh = oracle.ImageStencil(shader, width)
        got, ref = np.stack(fresh.image), gold[f"{name}_init_images"]
        assert np.all(np.abs(got.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64)) <= 1), "init profile (exp): 1 ulp"
        for bc in (0, 1, 2):
            s = _stencil_from_golden(oracle, gold, f"{name}_bc{bc}_1frame", shader, width, bc)
            s.compute(frames - 1)
            ref = gold[f"{name}_bc{bc}_{frames}frames_images"]
            got = np.stack(s.image)
            nan = np.isnan(ref)
            assert np.array_equal(np.isnan(got), nan) and np.array_equal(got[~nan], ref[~nan]), (name, bc)
            n = s.num_images
            assert list(s.unit) + list(s.read_index) + [s.write_index] == gold[f"{name}_bc{bc}_{frames}frames_state"].tolist()
