"""GPU parity: the three 3-D SPH passes (rho_pres_comp / force_comp / integrate_comp) and the coupled
frame, all-pairs (as shipped) and uniform-grid neighbour search, against the oracle.
Tolerance (north-star): rho, pressure, force, height within 1e-4 relative (FP32, summation order);
neighbour sets identical; branch decisions identical on the fixtures."""
import numpy as np
import pytest

from util import assert_close, jittered_block, smooth_field

pytestmark = pytest.mark.gpu

# small scene: 24 x 5 x 24 lattice inside a box that contains it
NX, NY, NZ = 24, 5, 24
BOX = dict(upper=(0.25, 1.0, 0.25, 500.0), lower=(0.0, -0.02, 0.0, 50.0))
GRID = ((0.0, -0.02, 0.0), (0.26, 0.18, 0.26), (13, 10, 13))       # 0.02-wide cells = 2h


def _params(oracle, wtype=1.0):
    prm = oracle.default_params3()
    for a in range(4):
        prm.upper[a] = BOX["upper"][a]
        prm.lower[a] = BOX["lower"][a]
    prm.attributes[3] = wtype
    return prm


def _scene(cwa, ctx, oracle, use_grid, wtype=1.0, vel=0.5, seed=1234, tex_w=64):
    prm = _params(oracle, wtype)
    ctx.set_params_from_oracle(prm)
    p = jittered_block(oracle, NX, NY, NZ, prm, seed=seed, vel=vel)
    rng = np.random.default_rng(seed + 1)
    p["force"] = rng.uniform(-1e4, 1e4, (p.size, 4)).astype(np.float32)     # previous-frame force feeds the torque term
    p["pos"][::7, 1] += np.float32(0.02)                                   # some particles above the crest threshold
    tex = smooth_field(tex_w, tex_w, 1, amp=0.02)
    grid = cwa.UniformGrid(ctx, 3, *GRID, p.size) if use_grid else None
    sph = cwa.Sph(ctx, p.size, grid, particles=p)
    wave = cwa.StencilImage2DTripleBuffered(ctx, tex_w, tex_w, 1, cwa.WAVE_COUPLED)
    wave.write_image(0, tex)
    sph.bind_wave(wave, 0)
    return prm, p, tex, sph, wave


def _oracle_grid(oracle, p):
    g = oracle.grid3(*GRID)
    _, cnt, off, idx = oracle.grid3_build(g, p["pos"])
    return (g, cnt, off, idx)


@pytest.mark.parametrize("use_grid", [False, True])
def test_neighbour_sets_identical(cwa, ctx, oracle, use_grid):
    prm, p, tex, sph, wave = _scene(cwa, ctx, oracle, use_grid)
    got = sph.neighbour_count()
    ref_all = oracle.sph3_neighbour_count(p, 0.01)
    assert np.array_equal(got, ref_all), "all-pairs and grid must see the same neighbour sets"
    assert 1 <= got.min() and got.max() > 6       # self is always counted (rho_pres_comp.glsl:60-67)


@pytest.mark.parametrize("use_grid", [False, True])
def test_rho_pres_pass(cwa, ctx, oracle, use_grid):
    prm, p, tex, sph, wave = _scene(cwa, ctx, oracle, use_grid)
    sph.rho_pres()
    got = sph.download()
    ref = p.copy()
    oracle.sph3_rho_pres(ref, prm, tex)                       # all-pairs oracle is the ground truth for both modes
    assert_close(got["extras"][:, 0], ref["extras"][:, 0], what="rho")
    assert_close(got["extras"][:, 1], ref["extras"][:, 1], what="pressure")
    for f in ("pos", "vel", "force"):
        assert np.array_equal(got[f], p[f]), f"{f} must be untouched by the density pass"
    assert np.array_equal(got["extras"][:, 2:], p["extras"][:, 2:])


@pytest.mark.parametrize("use_grid", [False, True])
def test_force_pass(cwa, ctx, oracle, use_grid):
    prm, p, tex, sph, wave = _scene(cwa, ctx, oracle, use_grid)
    sph.rho_pres()
    sph.force()
    got = sph.download()
    ref = p.copy()
    oracle.sph3_rho_pres(ref, prm, tex)
    oracle.sph3_force(ref, prm, tex)
    # per-particle scale = what the pair sums are made of (they cancel in the lattice interior)
    assert_close(got["force"][:, :3], ref["force"][:, :3], what="force.xyz")
    crest = p["pos"][:, 1] > 0.01
    assert crest.any() and (~crest).any()
    assert np.array_equal(got["force"][:, 3], ref["force"][:, 3]), "force.w: /0.25 in memory on the crest rule only"
    assert np.array_equal(got["pos"], p["pos"]) and np.array_equal(got["vel"], p["vel"])


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_allpairs_kernel_variants(cwa, ctx, oracle, mode):
    """The all-pairs passes have two CTA shapes: tiles of 64 targets (0), and one CTA per SM with an equal share of the targets for the
    density pass (1, the default) or both passes (2).  2880 particles / 148 SMs = 20 targets per CTA, 32 lanes each."""
    try:
        ctx.set_tuning(allpairs_balanced=mode)
        prm, p, tex, sph, wave = _scene(cwa, ctx, oracle, False)
        sph.rho_pres()
        sph.force()
        got = sph.download()
        ref = p.copy()
        oracle.sph3_rho_pres(ref, prm, tex)
        oracle.sph3_force(ref, prm, tex)
        assert_close(got["extras"][:, 0], ref["extras"][:, 0], what="rho")
        assert_close(got["extras"][:, 1], ref["extras"][:, 1], what="pressure")
        assert_close(got["force"][:, :3], ref["force"][:, :3], what="force.xyz")
        assert np.array_equal(got["force"][:, 3], ref["force"][:, 3])
    finally:
        ctx.set_tuning(allpairs_balanced=1)


@pytest.mark.parametrize("order", ["lattice", "shuffled"])
def test_allpairs_tile_culling_changes_nothing(cwa, ctx, oracle, order):
    """The balanced all-pairs kernels skip candidate tiles whose bounding box is farther than h from the CTA's targets (sph3_allpairs_boxes_kernel).
    A skipped tile holds no accepted pair, so every sum has the same terms in the same order: bit-identical records after several frames, with
    the particles in lattice order (most tiles skipped) and shuffled (nearly none), NaN and far-away particles included."""
    def run(cull):
        ctx.set_tuning(allpairs_balanced=2, allpairs_cull=cull)
        prm, p, tex, sph, wave = _scene(cwa, ctx, oracle, False, vel=5.0)
        if order == "shuffled":
            p = p[np.random.default_rng(5).permutation(p.size)]
        p["pos"][3, 0] = np.nan; p["pos"][700, :3] = np.nan
        p["pos"][1500, :3] = (40.0, 3.0, -25.0)                   # far outside everything: stretches its tile's box
        sph.upload(p)
        sph.step(3)
        a = sph.download()
        sph.rho_pres(); sph.force()                               # the passes dispatched alone take the boxes too
        return a, sph.download()
    try:
        a0, b0 = run(0)
        a1, b1 = run(1)
        for f in ("pos", "vel", "force", "extras"):
            assert a0[f].tobytes() == a1[f].tobytes(), f
            assert b0[f].tobytes() == b1[f].tobytes(), f
        assert np.isfinite(a1["pos"][:, :3]).sum() > 0.9 * a1["pos"][:, :3].size
    finally:
        ctx.set_tuning(allpairs_balanced=1, allpairs_cull=1)


@pytest.mark.parametrize("use_grid", [False, True])
def test_integrate_pass_and_branch_decisions(cwa, ctx, oracle, use_grid):
    prm, p, tex, sph, wave = _scene(cwa, ctx, oracle, use_grid, vel=40.0)      # some |v| > 25: foam rule fires
    sph.rho_pres(); sph.force(); sph.integrate()
    got = sph.download()
    ref = p.copy()
    oracle.sph3_rho_pres(ref, prm, tex); oracle.sph3_force(ref, prm, tex)
    pre = ref.copy()
    oracle.sph3_integrate(ref, prm, tex)
    foam = ref["extras"][:, 0] != pre["extras"][:, 0]
    assert foam.any() and (~foam).any(), "fixture must exercise both sides of the foam rule"
    wall = (ref["pos"][:, 0] == prm.lower[0]) | (ref["pos"][:, 0] == prm.upper[0]) | (ref["pos"][:, 1] == prm.lower[1])
    assert_close(got["pos"][:, :3], ref["pos"][:, :3], scale=1e-3, what="pos")
    assert_close(got["vel"][:, :3], ref["vel"][:, :3], what="vel")
    assert_close(got["extras"][:, :2], ref["extras"][:, :2], what="rho/p after foam")
    # decisions: clamped coordinates are exactly the wall values in both
    for ax in range(3):
        for bound in (prm.lower[ax], prm.upper[ax]):
            assert np.array_equal(got["pos"][:, ax] == np.float32(bound), ref["pos"][:, ax] == np.float32(bound)), (ax, bound)
    assert np.array_equal(got["pos"][:, 3], p["pos"][:, 3]) and np.array_equal(got["extras"][:, 2:], p["extras"][:, 2:])
    _ = wall


@pytest.mark.parametrize("use_grid", [False, True])
def test_fused_sph_step_equals_three_passes(cwa, ctx, oracle, use_grid):
    prm, p, tex, sph, wave = _scene(cwa, ctx, oracle, use_grid)
    sph.step(1)
    fused = sph.download()
    sph.upload(p)
    sph.rho_pres(); sph.force(); sph.integrate()
    split = sph.download()
    for f in ("pos", "vel", "force", "extras"):
        assert_close(fused[f], split[f], rtol=2e-6, what=f)     # same kernels; only contraction of the last pass may differ
    ref = p.copy()
    oracle.sph3_rho_pres(ref, prm, tex); oracle.sph3_force(ref, prm, tex); oracle.sph3_integrate(ref, prm, tex)
    assert_close(fused["pos"][:, :3], ref["pos"][:, :3], scale=1e-3, what="pos")
    assert_close(fused["vel"][:, :3], ref["vel"][:, :3], what="vel")
    assert_close(fused["force"][:, :3], ref["force"][:, :3], what="force")


def test_north_star_entry_points_sph_step_wave_step(cwa, ctx, oracle):
    prm, p, tex, sph, wave = _scene(cwa, ctx, oracle, True)
    ctx.bind_scene(sph, wave)
    ctx.sph_step(1)
    a = sph.download()
    sph.upload(p)
    sph.step(1)
    assert np.array_equal(a.view(np.uint8), sph.download().view(np.uint8)), "sph_step == cwa_sph_step, deterministic"
    before = wave.state()
    ctx.wave_step(2)
    assert wave.state() != before


def test_grid_run_is_bit_reproducible(cwa, ctx, oracle):
    prm, p, tex, sph, wave = _scene(cwa, ctx, oracle, True)
    sph.step(3)
    a = sph.download()
    sph.upload(p)
    sph.step(3)
    b = sph.download()
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), "canonical cell order makes runs bit-reproducible"


def test_compute_shader_dispatch_by_glsl_name(cwa, ctx, oracle):
    prm, p, tex, sph, wave = _scene(cwa, ctx, oracle, False)
    progs = [cwa.ComputeShader(ctx, n) for n in ("rho_pres_comp.glsl", "force_comp.glsl", "integrate_comp.glsl")]
    for pr in progs:                      # idle(): glUseProgram + glDispatchCompute(PART_WORK_GROUPS,1,1) x3
        pr.bind_object(sph)
        pr.Dispatch(20, 1, 1)
    got = sph.download()
    ref = p.copy()
    oracle.sph3_rho_pres(ref, prm, tex); oracle.sph3_force(ref, prm, tex); oracle.sph3_integrate(ref, prm, tex)
    assert_close(got["pos"][:, :3], ref["pos"][:, :3], scale=1e-3, what="pos")
    assert_close(got["force"][:, :3], ref["force"][:, :3], what="force")
    with pytest.raises(cwa.CwaError):
        cwa.ComputeShader(ctx, "no_such_shader.glsl")       # InitShader() == -1


def test_unbound_texture_samples_zero(cwa, ctx, oracle):
    prm, p, tex, sph, wave = _scene(cwa, ctx, oracle, True)
    sph.bind_wave(None)
    sph.step(1)
    got = sph.download()
    ref = p.copy()
    oracle.sph3_rho_pres(ref, prm, None); oracle.sph3_force(ref, prm, None); oracle.sph3_integrate(ref, prm, None)
    assert_close(got["force"][:, :3], ref["force"][:, :3], what="force (unbound texture)")
    assert_close(got["pos"][:, :3], ref["pos"][:, :3], scale=1e-3, what="pos")


def test_init_cube_matches_make_cube(cwa, ctx, oracle):
    sph = cwa.Sph(ctx, 64 * 5 * 64)
    sph.init_cube(64, 5, 64)
    ref = oracle.make_cube(64, 5, 64)
    assert np.array_equal(sph.download().view(np.uint8), ref.view(np.uint8))


@pytest.mark.parametrize("coupling", [0, 1])
@pytest.mark.parametrize("use_grid", [False, True])
def test_coupled_frames_small_scene(cwa, ctx, oracle, coupling, use_grid):
    """10 coupled frames: SPH sampling schedule (AS_SHIPPED / LATEST), wave bit-exact, particles 1e-4."""
    prm = _params(oracle)
    ctx.set_params_from_oracle(prm)
    p = jittered_block(oracle, NX, NY, NZ, prm, vel=0.2)
    oc = oracle.Coupled(p.size, 64, 64, 1, prm, coupling, grid=GRID if use_grid else None)
    oc.particles[:] = p
    grid = cwa.UniformGrid(ctx, 3, *GRID, p.size) if use_grid else None
    sph = cwa.Sph(ctx, p.size, grid, particles=p)
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 1, cwa.WAVE_COUPLED)
    for frame in range(1, 11):
        oc.step(1)
        sph.coupled_step(wave, 1, coupling)
        assert wave.state()["tex_unit0"] == oc.sampled_image(), frame
        assert np.array_equal(wave.read_role(0).view(np.uint32), oc.wave(0).view(np.uint32)), frame
        if frame in (1, 2, 5, 10):
            got, ref = sph.download(), oc.particles
            assert_close(got["extras"][:, 0], ref["extras"][:, 0], rtol=1e-3 if frame > 2 else 1e-4, what=f"rho f{frame}")
            assert_close(got["pos"][:, :3], ref["pos"][:, :3], rtol=1e-3 if frame > 2 else 1e-4, scale=1e-3, what=f"pos f{frame}")
    oc.close()


def test_default_scene_as_shipped_reproduces_reference_nan_behaviour(cwa, ctx, oracle):
    """Config D, frames 1-2: the shipped box (upper.xz = 0.48) is narrower than the 64x5x64 lattice, the first
    integrate clamps 7 columns/rows onto the wall, coincident particles make normalize(0) = NaN on frame 2
    (SURVEY Appendix C).  The CUDA path must reproduce exactly the same NaN set as the oracle."""
    prm = oracle.default_params3()
    ctx.set_params_from_oracle(prm)
    p = oracle.make_cube(64, 5, 64, prm)
    oc = oracle.Coupled(p.size, 64, 64, 4, prm, oracle.COUPLING_AS_SHIPPED)
    oc.particles[:] = p
    sph = cwa.Sph(ctx, p.size, None, particles=p)
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 4, cwa.WAVE_COUPLED)       # RGBA32F, as shipped
    oc.step(1); sph.coupled_step(wave, 1, cwa.COUPLING_AS_SHIPPED)
    got, ref = sph.download(), oc.particles
    assert not np.isnan(ref["pos"]).any()
    assert_close(got["extras"][:, 0], ref["extras"][:, 0], what="rho frame 1")
    assert_close(got["pos"][:, :3], ref["pos"][:, :3], scale=1e-3, what="pos frame 1")
    assert np.array_equal(got["pos"][:, 0] == np.float32(0.48), ref["pos"][:, 0] == np.float32(0.48))
    oc.step(1); sph.coupled_step(wave, 1, cwa.COUPLING_AS_SHIPPED)
    got, ref = sph.download(), oc.particles
    nan_ref = np.isnan(ref["pos"][:, :3]).any(1)
    nan_got = np.isnan(got["pos"][:, :3]).any(1)
    assert nan_ref.sum() > 1000
    assert np.array_equal(nan_got, nan_ref), "same particles blow up as in the reference arithmetic"
    ok = ~nan_ref
    assert_close(got["pos"][ok, :3], ref["pos"][ok, :3], scale=1e-3, rtol=1e-3, what="pos frame 2 (finite particles)")
    oc.close()


def test_full_size_c4_properties(cwa, ctx, oracle):
    """1M particles + 2048^2 (config C4) through size-independent properties: particle count conserved,
    no NaN in a well-posed box, neighbour counts of the lattice interior, grid/all-pairs agreement on a
    random sample of particles checked by the oracle's pair loop."""
    s = 7
    nx, ny, nz = 448, 5, 448
    n = nx * ny * nz
    ctx.set_boundary(upper=(0.55 * s, 1.0, 0.55 * s, 500.0), lower=(0.0, -0.02, 0.0, 50.0))
    ctx.set_sim_constants(uv_scale=2.0 / s)
    grid = cwa.UniformGrid(ctx, 3, (0.0, -0.02, 0.0), (0.55 * s, 1.0, 0.55 * s), (192, 51, 192), n)
    sph = cwa.Sph(ctx, n, grid)
    sph.init_cube(nx, ny, nz)
    wave = cwa.StencilImage2DTripleBuffered(ctx, 2048, 2048, 1, cwa.WAVE_COUPLED)
    p0 = sph.download()
    cnt = sph.neighbour_count()
    interior = cnt.reshape(nx, ny, nz)[2:-2, 1:-1, 2:-2]
    assert (interior == 7).all(), "lattice interior: self + 6 face neighbours inside h"
    sph.coupled_step(wave, 1, cwa.COUPLING_LATEST)
    p1 = sph.download()
    assert p1.size == n and not np.isnan(p1["pos"]).any()
    # oracle check on a sample: density of the initial lattice via the oracle's grid loop on a sub-block
    sub = p0.reshape(nx, ny, nz)[100:124, :, 200:224].reshape(-1).copy()
    prm = oracle.default_params3()
    prm.uv_scale = 2.0 / s
    for a, v in enumerate((0.55 * s, 1.0, 0.55 * s)):
        prm.upper[a] = v
    ref = sub.copy()
    tex = oracle.wave_init(2048, 2048, 1, 0, 1.0)
    oracle.sph3_rho_pres(ref, prm, tex)
    got_sub = p1.reshape(nx, ny, nz)[100:124, :, 200:224]
    core = (slice(2, -2), slice(None), slice(2, -2))
    assert_close(got_sub["extras"][..., 0][core], ref.reshape(24, ny, 24)["extras"][..., 0][core], what="rho (sample block core)")
