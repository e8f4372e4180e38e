"""GPU parity of the 1-D wave substrates: Shallow1D_cs / Wave1D_cs driven by ImageStencil, against the CPU
restatement.  The iteration is integer-free FP32 in the GLSL's association order -> bit-exact; the initial profiles go through
exp() (implementation-defined precision in GLSL, evaluated in double and rounded once on both sides) -> 1 ulp."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits_equal(a, b):
    """Bit for bit, except that any NaN equals any NaN (x86 produces the negative quiet NaN for 0/0, the GPU the canonical one)."""
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def _ulp_close(a, b, ulps=1):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    same_nan = np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    return same_nan and bool(np.all(np.abs(ia - ib)[ok] <= ulps))


def _sync_images(gpu, orc):
    """Give the oracle object the GPU object's images (so later steps compare bit for bit whatever exp() rounded to)."""
    for i in range(orc.num_images):
        orc.image[i][:] = gpu.read_image(i)


@pytest.mark.parametrize("shader,width", [(0, 128), (0, 512), (0, 3), (0, 2), (1, 1024), (1, 129), (1, 3), (1, 2), (0, 20000), (1, 20000)])
def test_init_state_and_bookkeeping(cwa, ctx, oracle, shader, width):
    g = cwa.ImageStencil(ctx, shader, width)
    o = oracle.ImageStencil(shader, width)
    st = g.state()
    assert st["read_index"] == o.read_index and st["write_index"] == o.write_index and st["unit"] == o.unit
    for i in range(o.num_images):
        assert _ulp_close(g.read_image(i), o.image[i]), f"image {i} after Init()"


@pytest.mark.parametrize("bc", [0, 1, 2])
@pytest.mark.parametrize("shader,width,frames", [(0, 128, 40), (0, 512, 7), (0, 5, 3), (1, 1024, 5), (1, 37, 3), (0, 20000, 2), (1, 20000, 1)])
def test_compute_is_bit_exact(cwa, ctx, oracle, shader, width, frames, bc):
    g = cwa.ImageStencil(ctx, shader, width)
    o = oracle.ImageStencil(shader, width)
    lam, p3 = (0.001, 0.1) if shader == 0 else (0.01, 0.9995)
    g.set_params(lam, p3, 0.001, (0.3, -0.2), bc)
    o.prm.bc = bc; o.prm.boundary[0] = 0.3; o.prm.boundary[1] = -0.2
    _sync_images(g, o)
    g.Compute(frames)
    o.compute(frames)
    st = g.state()
    assert st["read_index"] == o.read_index and st["write_index"] == o.write_index and st["unit"] == o.unit
    for i in range(o.num_images):
        assert _bits_equal(g.read_image(i), o.image[i]), f"image {i} after {frames} frames (bc {bc})"
    # frame by frame (one launch each) equals all frames in one launch
    g2 = cwa.ImageStencil(ctx, shader, width)
    g3 = cwa.ImageStencil(ctx, shader, width)
    for gg in (g2, g3):
        gg.set_params(lam, p3, 0.001, (0.3, -0.2), bc)
    for _ in range(frames):
        g2.Compute(1)
    g3.Compute(frames)
    assert g2.state() == g3.state()
    for i in range(o.num_images):
        assert _bits_equal(g2.read_image(i), g3.read_image(i))


@pytest.mark.parametrize("bc", [0, 1, 2])
@pytest.mark.parametrize("shader,name,width,frames", [(0, "shallow", 128, 40), (1, "wave", 1024, 5)])
def test_committed_golden_vectors(cwa, ctx, shader, name, width, frames, bc):
    """From the committed one-frame state to the committed many-frames state (tests/golden/golden_v2_stencil1d.npz), bit for bit."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2_stencil1d.npz"))
    g = cwa.ImageStencil(ctx, shader, width)
    lam, p3 = (0.001, 0.1) if shader == 0 else (0.01, 0.9995)
    g.set_params(lam, p3, 0.001, (0.3, -0.2), bc)
    start, want = gold[f"{name}_bc{bc}_1frame_images"], gold[f"{name}_bc{bc}_{frames}frames_images"]
    # a fresh object has done Reinit (nread dispatches); the golden state has done one frame more.  Bring the roles in line first.
    g.Compute(1)
    st = g.state()
    assert st["unit"] + st["read_index"] + [st["write_index"]] == gold[f"{name}_bc{bc}_1frame_state"].tolist()
    for i in range(g.num_images):
        g.write_image(i, start[i])
    g.Compute(frames - 1)
    st = g.state()
    assert st["unit"] + st["read_index"] + [st["write_index"]] == gold[f"{name}_bc{bc}_{frames}frames_state"].tolist()
    for i in range(g.num_images):
        assert _bits_equal(g.read_image(i), want[i]), f"image {i}"


def test_splash_compute_func_and_reinit_from_texture(cwa, ctx, oracle):
    g = cwa.ImageStencil(ctx, cwa.STENCIL1D_SHALLOW, 128)
    o = oracle.ImageStencil(oracle.STENCIL1D_SHALLOW, 128)
    _sync_images(g, o)
    g.Compute(10); o.compute(10)
    g.ComputeFunc(1); o.compute_func(1)                      # Splash
    assert g.state()["unit"] == o.unit
    assert _ulp_close(g.GetReadImage(0), o.read_image(0), 2)
    _sync_images(g, o)
    g.Compute(5); o.compute(5)
    assert _bits_equal(g.GetReadImage(0), o.read_image(0))
    tex = np.random.default_rng(3).uniform(4.0, 6.0, (100, 4)).astype(np.float32)    # narrower than the image: the rest reads 0
    g.ReinitFromTexture(tex); o.reinit_from_texture(tex)
    assert g.state()["read_index"] == o.read_index and g.state()["unit"] == o.unit
    assert _bits_equal(g.GetReadImage(0), o.read_image(0))
    g.set_iterate(False)
    before = g.GetReadImage(0)
    g.Compute(3)
    assert _bits_equal(before, g.GetReadImage(0))


@pytest.mark.parametrize("shader", [0, 1])
def test_compute_shader_dispatch_plus_pingpong_equals_compute_func(cwa, ctx, shader):
    """Driving the object the way ImageStencil::ComputeFunc does -- SetMode, Dispatch, PingPong (StencilImage2D.cpp:107-120)."""
    name = "Shallow1D_cs.glsl" if shader == 0 else "Wave1D_cs.glsl"
    a = cwa.ImageStencil(ctx, shader, 256)
    b = cwa.ImageStencil(ctx, shader, 256)
    cs = cwa.ComputeShader(ctx, name)
    cs.bind_object(b)
    modes = [2, 3, 2, 3, 1] if shader == 0 else [2, 2, 2]
    for m in modes:
        a.ComputeFunc(m)
        cs.SetMode(m)
        cs.Dispatch(1, 1, 1)
        b.PingPong()
    assert a.state() == b.state()
    for i in range(a.num_images):
        assert _bits_equal(a.read_image(i), b.read_image(i))
    with pytest.raises(cwa.CwaError):
        other = cwa.ComputeShader(ctx, "Wave1D_cs.glsl" if shader == 0 else "Shallow1D_cs.glsl")
        other.bind_object(b)
        other.Dispatch(1, 1, 1)


def test_shallow_water_stays_finite_and_drains_only_through_the_free_ends(cwa, ctx):
    g = cwa.ImageStencil(ctx, cwa.STENCIL1D_SHALLOW, 128)
    h0 = g.GetReadImage(0)[:, 0].astype(np.float64)
    g.Compute(2000)
    r = g.GetReadImage(0)
    assert np.isfinite(r[:, :2]).all()
    m0, m1 = h0[1:-1].sum(), r[1:-1, 0].astype(np.float64).sum()
    assert m1 <= m0 * (1 + 1e-5) and m1 > 0.95 * m0            # free boundary: the initial bump (2.5 % of the mass) runs out of the domain
    assert 4.0 < r[:, 0].min() and r[:, 0].max() < 5.8


def test_sph2d_samples_the_evolving_shallow_wave(cwa, ctx, oracle):
    """The complete 2-D frame of SphWave2D/Main.cpp:237-240: sph2d.Compute, wave1d.Compute, GetReadImage(0).BindTextureUnit()."""
    from util import assert_close
    n = 4096
    ext = ((0.0, 0.0), (9.6, 9.6), (32, 32))
    prm = oracle.default_params2(oracle.SPH2_WAVE)
    grid = cwa.UniformGrid(ctx, 2, *ext, n)
    sph = cwa.SphUgrid(ctx, n, grid, cwa.SPH2_WAVE, substeps=2)
    wave = cwa.ImageStencil(ctx, cwa.STENCIL1D_SHALLOW, 128)
    o_wave = oracle.ImageStencil(oracle.STENCIL1D_SHALLOW, 128)
    _sync_images(wave, o_wave)
    b0, b1 = sph.download(), np.zeros(n, cwa.PARTICLE2D)
    og = oracle.grid2(*ext)
    r = 0
    sph.bind_wave1d(wave.read_buffer(0), 128)
    tex = o_wave.read_image(0).copy()
    for _ in range(3):
        sph.Compute(1)
        wave.Compute(1)
        sph.bind_wave1d(wave.read_buffer(0), 128)
        r, _g = oracle.sph2_step(b0, b1, r, 2, prm, tex.reshape(1, 128, 4), og)
        o_wave.compute(1)
        tex = o_wave.read_image(0).copy()
    assert _bits_equal(wave.GetReadImage(0), o_wave.read_image(0))
    got, ref = sph.download(), (b0, b1)[r]
    assert_close(got["acc"][:, 3], ref["acc"][:, 3], rtol=1e-3, what="rho after 3 coupled frames")
    assert_close(got["pos"][:, :2], ref["pos"][:, :2], rtol=1e-3, scale=1e-2, what="pos after 3 coupled frames")
    assert np.array_equal(got["pos"][:, 3], ref["pos"][:, 3]), "enable/disable flag follows the wave"
