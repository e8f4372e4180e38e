"""GPU parity: wave stencil (wave_comp.glsl / Wave2D_cs.glsl) and the triple-buffer host object.
The kernels evaluate the update in the GLSL's association order with round-to-nearest intrinsics,
so single steps are compared BIT-EXACT against the oracle (tolerance 0)."""
import numpy as np
import pytest

from util import random_field, smooth_field

pytestmark = pytest.mark.gpu


def _load_levels(wave, u0, u1):
    wave.write_role(0, u0)     # image unit 0: u^{t-1}
    wave.write_role(1, u1)     # image unit 1: u^{t-2}


@pytest.mark.parametrize("w,h", [(64, 64), (128, 16), (132, 18), (256, 40), (512, 512), (1024, 96), (4, 4), (8, 3)])
@pytest.mark.parametrize("variant", [0, 1])
def test_single_step_bit_exact_scalar_tma_path(cwa, ctx, oracle, w, h, variant):
    wave = cwa.StencilImage2DTripleBuffered(ctx, w, h, 1, variant)
    u0, u1 = random_field(h, w, 1), random_field(h, w, 2)
    _load_levels(wave, u0, u1)
    wave.Compute(1)
    got = wave.read_role(0)
    lam, att, beta = (0.01, 0.985, 0.001) if variant == 0 else (0.01, 0.9995, 0.001)
    ref = oracle.wave_evolve(u0, u1, variant, lam, att, beta, 1.0)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"max abs diff {np.abs(got - ref).max()}"
    # the previous newest level moved to role 1, untouched
    assert np.array_equal(wave.read_role(1), u0)


@pytest.mark.parametrize("w,h,ch", [(63, 17, 1), (130, 9, 1), (64, 64, 4), (50, 33, 4)])
def test_single_step_bit_exact_generic_path(cwa, ctx, oracle, w, h, ch):
    wave = cwa.StencilImage2DTripleBuffered(ctx, w, h, ch, cwa.WAVE_COUPLED)
    u0, u1 = random_field(h, w, 3, ch), random_field(h, w, 4, ch)
    _load_levels(wave, u0, u1)
    wave.Compute(1)
    ref = oracle.wave_evolve(u0, u1, 0, 0.01, 0.985, 0.001, 1.0)
    assert np.array_equal(wave.read_role(0).view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("wtype", [1.0, 0.0, 0.5])
@pytest.mark.parametrize("ch", [1, 4])
def test_init_modes_match_oracle(cwa, ctx, oracle, wtype, ch):
    ctx.set_wave_uniforms(attributes=(0.01, 0.985, 0.001, wtype))
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, ch, cwa.WAVE_COUPLED)
    ref = oracle.wave_init(64, 64, ch, 0, wtype)
    for role in (0, 1):        # Reinit runs INIT twice: both read levels hold the bump
        assert np.array_equal(wave.read_role(role).view(np.uint32), ref.view(np.uint32))
    st = wave.state()
    assert st["read_index"] == [2, 0] and st["write_index"] == 1 and st["unit"] == [2, 0, 1]


def test_simp_variant_init_and_params(cwa, ctx, oracle):
    wave = cwa.StencilImage2DTripleBuffered(ctx, 96, 64, 1, cwa.WAVE_SIMP)
    ref = oracle.wave_init(96, 64, 1, 1, 1.0)
    assert np.array_equal(wave.read_role(0), ref)
    wave.set_params(0.02, 0.99, 0.002)
    u0, u1 = random_field(64, 96, 5), random_field(64, 96, 6)
    _load_levels(wave, u0, u1)
    wave.Compute(1)
    assert np.array_equal(wave.read_role(0), oracle.wave_evolve(u0, u1, 1, 0.02, 0.99, 0.002, 1.0))


def test_wake_source_column(cwa, ctx, oracle):
    ctx.set_wave_uniforms(attributes=(0.01, 0.985, 0.001, 0.5))       # 0 < type < 1: boat wake
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 1, cwa.WAVE_COUPLED)
    u0 = np.abs(random_field(64, 64, 8)) + 0.01
    u1 = np.abs(random_field(64, 64, 9)) * 0.1
    _load_levels(wave, u0, u1)
    wave.Compute(1)
    ref = oracle.wave_evolve(u0, u1, 0, 0.01, 0.985, 0.001, 0.5)
    got = wave.read_role(0)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    plain = oracle.wave_evolve(u0, u1, 0, 0.01, 0.985, 0.001, 1.0)
    assert (ref[:, 32] != plain[:, 32]).any() and np.array_equal(ref[:, :32], plain[:, :32])


def test_multi_step_rotation_and_long_run_statistics(cwa, ctx, oracle):
    """200 steps from the shipped init: bit-exact against the oracle at every checked step (each
    step is bit-exact, so the sequence is), RMS tracked."""
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 1, cwa.WAVE_COUPLED)
    a = oracle.wave_init(64, 64, 1, 0, 1.0)
    b = a.copy()
    for step in range(1, 201):
        c = oracle.wave_evolve(a, b, 0, 0.01, 0.985, 0.001, 1.0)
        a, b = c, a
        wave.Compute(1)
        if step in (1, 2, 3, 4, 50, 200):
            got = wave.read_role(0)
            assert np.array_equal(got.view(np.uint32), a.view(np.uint32)), step
            assert np.array_equal(wave.read_role(1).view(np.uint32), b.view(np.uint32)), step
    assert abs(np.sqrt(np.mean(wave.read_role(0) ** 2)) - np.sqrt(np.mean(a ** 2))) == 0.0


def test_energy_conserved_without_damping(cwa, ctx, oracle):
    """beta = 0, atten = 1: the leapfrog scheme conserves the discrete energy to round-off."""
    wave = cwa.StencilImage2DTripleBuffered(ctx, 128, 128, 1, cwa.WAVE_SIMP)
    wave.set_params(0.2, 1.0, 0.0)
    u = smooth_field(128, 128)
    _load_levels(wave, u, u)

    def energy(un, uo, lam=0.2):
        def lap_dot(p, q):
            px = np.pad(p, 1, mode="edge")
            return -np.sum((px[1:-1, 2:] - p) * (np.pad(q, 1, mode="edge")[1:-1, 2:] - q)) - np.sum(
                (px[2:, 1:-1] - p) * (np.pad(q, 1, mode="edge")[2:, 1:-1] - q))
        un = un.astype(np.float64); uo = uo.astype(np.float64)
        return np.sum((un - uo) ** 2) - lam * lap_dot(un, uo)
    e0 = None
    for it in range(50):
        wave.Compute(10)
        e = energy(wave.read_role(0), wave.read_role(1))
        e0 = e if e0 is None else e0
        assert abs(e - e0) <= 1e-4 * abs(e0), (it, e, e0)


def test_evolve_checkbox_and_test_mode_are_no_ops(cwa, ctx):
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 1, cwa.WAVE_COUPLED)
    before, st = wave.read_role(0), wave.state()
    wave.set_evolve(False)
    wave.Compute(3)                                   # mEvolve == false -> silent no-op (.cpp:81)
    assert wave.state() == st and np.array_equal(wave.read_role(0), before)
    cs = cwa.ComputeShader(ctx, "wave_comp.glsl")
    cs.bind_object(wave)
    cs.SetMode(cwa.MODE_TEST)
    cs.Dispatch(2, 2, 1)                              # MODE_TEST does nothing (wave_comp.glsl:71-72)
    assert np.array_equal(wave.read_role(0), before)


def test_compute_shader_dispatch_equals_compute_minus_pingpong(cwa, ctx, oracle):
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 1, cwa.WAVE_COUPLED)
    u0, u1 = random_field(64, 64, 12), random_field(64, 64, 13)
    _load_levels(wave, u0, u1)
    cs = cwa.ComputeShader(ctx, "wave_comp.glsl")
    cs.bind_object(wave)
    cs.SetMode(cwa.MODE_EVOLVE)
    cs.Dispatch(2, 2, 1)
    ref = oracle.wave_evolve(u0, u1, 0, 0.01, 0.985, 0.001, 1.0)
    assert np.array_equal(wave.read_role(2), ref)     # written to the unit-2 image, roles not rotated
    wave.PingPong()
    assert np.array_equal(wave.read_role(0), ref)


def test_reinit_from_texture(cwa, ctx):
    wave = cwa.StencilImage2DTripleBuffered(ctx, 32, 16, 4, cwa.WAVE_COUPLED)
    tex = np.random.default_rng(2).random((16, 64, 4)).astype(np.float32)
    wave.ReinitFromTexture(tex)
    got = wave.read_role(0)
    assert np.array_equal(got, tex[:, ::2, :])        # texelFetch(uInitImage, coord*ivec2(2,1)) (wave_comp.glsl:78)


def test_as_shipped_unit_rotation_schedule(cwa, ctx):
    """Appendix B: which physical image display() leaves on texture unit 0."""
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 1, cwa.WAVE_COUPLED)
    seen = []
    for frame in range(1, 8):
        seen.append(wave.state()["tex_unit0"])        # what the SPH passes of this frame sample
        wave.Compute(1)
        wave.bind_texture_unit()
    assert seen == [-1, 0, 0, 0, 0, 0, 0]
    # role-correct rotation: unit 0 always holds the newest level
    assert wave.state()["unit"][wave.role_image(0)] == 0


def test_full_size_c3_linearity_and_symmetry(cwa, ctx):
    """4096^2 (config C3) through size-independent properties: linearity of the update and mirror
    symmetry of the stencil with clamp-to-edge borders."""
    n = 4096
    wave = cwa.StencilImage2DTripleBuffered(ctx, n, n, 1, cwa.WAVE_SIMP)
    rng = np.random.default_rng(0)
    a0 = rng.standard_normal((n, n)).astype(np.float32)
    a1 = rng.standard_normal((n, n)).astype(np.float32)

    def step(u0, u1):
        wave.write_role(0, u0); wave.write_role(1, u1)
        wave.Compute(1)
        return wave.read_role(0)
    ya = step(a0, a1)
    yf = step(a0[::-1, :].copy(), a1[::-1, :].copy())
    assert np.array_equal(yf, ya[::-1, :]), "y mirror symmetry (n+s commutes): exact, incl. top/bottom borders"
    yx = step(a0[:, ::-1].copy(), a1[:, ::-1].copy())
    assert np.allclose(yx, ya[:, ::-1], rtol=0, atol=2e-6), "x mirror symmetry up to the e/w summation order"
    y2 = step(2.0 * a0, 2.0 * a1)
    assert np.array_equal(y2, 2.0 * ya), "scaling by 2 is exact in FP32"
    # interior spot check against the formula
    i, j = 1777, 2999
    lam, att, beta = np.float32(0.01), np.float32(0.9995), np.float32(0.001)
    s = np.float32(np.float32(np.float32(a0[i + 1, j] + a0[i - 1, j]) + a0[i, j + 1]) + a0[i, j - 1])
    kc = np.float32(np.float32(np.float32(2.0) - np.float32(4.0) * lam) - beta)
    v = np.float32(np.float32(np.float32(kc * a0[i, j]) + np.float32(lam * s)) - np.float32(np.float32(1.0 - beta) * a1[i, j]))
    assert ya[i, j] == np.float32(v * att)
