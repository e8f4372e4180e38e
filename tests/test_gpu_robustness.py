"""Round-2 robustness fixes: per-context tuning state, failure paths that must leave objects intact, contexts on several devices."""
import os

import numpy as np
import pytest

from util import jittered_block

pytestmark = pytest.mark.gpu


def _small_scene(cwa, ctx, oracle):
    prm = oracle.default_params3()
    prm.upper[0] = prm.upper[2] = 0.25
    ctx.set_params_from_oracle(prm)
    p = jittered_block(oracle, 24, 5, 24, prm)
    grid = cwa.UniformGrid(ctx, 3, (0.0, -0.02, 0.0), (0.26, 0.18, 0.26), (25, 19, 25), p.size)
    sph = cwa.Sph(ctx, p.size, grid, particles=p)
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 1, cwa.WAVE_COUPLED)
    return p, grid, sph, wave


def test_two_contexts_hold_different_tuning(cwa, oracle):
    """cwa_set_tuning is per context: one context runs the shared-memory-staged lanes kernels without pipelining, the other the list
    kernels with pipelining, interleaved in one process -- both must produce what they produce alone."""
    a, b = cwa.Context(0), cwa.Context(0)
    a.set_tuning(nb_config=0, pipeline=0, scan_config=0, wave_transpose=0)
    b.set_tuning(nb_config=8, pipeline=3, scan_config=2, wave_transpose=1)
    pa, ga, sa, wa = _small_scene(cwa, a, oracle)
    pb, gb, sb, wb = _small_scene(cwa, b, oracle)
    for _ in range(3):
        sa.coupled_step(wa, 2, cwa.COUPLING_LATEST)
        sb.coupled_step(wb, 2, cwa.COUPLING_LATEST)
    ra, rb = sa.download(), sb.download()
    # reference runs, each alone in a fresh context with the same knobs
    out = []
    for knobs in (dict(nb_config=0, pipeline=0, scan_config=0, wave_transpose=0), dict(nb_config=8, pipeline=3, scan_config=2, wave_transpose=1)):
        c = cwa.Context(0)
        c.set_tuning(**knobs)
        p, g, s, w = _small_scene(cwa, c, oracle)
        for _ in range(3):
            s.coupled_step(w, 2, cwa.COUPLING_LATEST)
        out.append(s.download())
        c.close()
    assert np.array_equal(ra.view(np.uint8), out[0].view(np.uint8)), "context A kept ITS knobs while B ran with others"
    assert np.array_equal(rb.view(np.uint8), out[1].view(np.uint8))
    assert not np.array_equal(ra.view(np.uint8), rb.view(np.uint8)), "the two kernel families differ in the last bits (summation order): the test would be vacuous otherwise"
    a.close(); b.close()


def test_truncated_checkpoint_restores_nothing(cwa, ctx, oracle, tmp_path):
    p, grid, sph, wave = _small_scene(cwa, ctx, oracle)
    sph.coupled_step(wave, 2, cwa.COUPLING_LATEST)
    path = str(tmp_path / "state.ckpt")
    ctx.checkpoint_save(sph, wave, 2, path)
    sph.coupled_step(wave, 2, cwa.COUPLING_LATEST)
    before_p, before_w, before_state = sph.download(), [wave.read_image(i) for i in range(3)], wave.state()
    full = os.path.getsize(path)
    with open(path, "r+b") as f:
        f.truncate(full - 4096)                      # cut inside the last wave image: everything before it is present
    with pytest.raises(cwa.CwaError, match="truncated"):
        ctx.checkpoint_load(sph, wave, path)
    assert np.array_equal(sph.download().view(np.uint8), before_p.view(np.uint8)), "particles untouched"
    for i in range(3):
        assert np.array_equal(wave.read_image(i), before_w[i]), "wave images untouched"
    assert wave.state() == before_state


def test_resize_of_a_row_block_fails_without_freeing_it(cwa, ctx):
    wave = cwa.StencilImage2DTripleBuffered.create_block(ctx, 64, 256, 32, 96, 1, cwa.WAVE_COUPLED)
    before = wave.read_role(0)
    with pytest.raises(cwa.CwaError, match="row-block"):
        wave.resize(128, 128)
    wave.height, wave.w = 96, 64
    wave.Compute(3)                                   # still alive: images, descriptors and buffers were not freed
    assert wave.read_role(0).shape == before.shape and np.isfinite(wave.read_role(0)).all()


def test_default_sim_block_is_48_bytes(cwa, ctx):
    b = ctx.default_ubo(cwa.UBO_SIM)
    v = b.read(np.float32)
    assert v.size == 12 and v[0] == np.float32(0.005) and v[7] == 2.0


def test_contexts_on_two_devices_in_one_thread(cwa, oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    a, b = cwa.Context(0), cwa.Context(1)
    pa, ga, sa, wa = _small_scene(cwa, a, oracle)
    pb, gb, sb, wb = _small_scene(cwa, b, oracle)
    for _ in range(2):
        sa.coupled_step(wa, 2, cwa.COUPLING_LATEST)
        sb.coupled_step(wb, 2, cwa.COUPLING_LATEST)
    assert np.array_equal(sa.download().view(np.uint8), sb.download().view(np.uint8)), "same scene, same bits on either device"
    a.close(); b.close()


def test_gl_interop_entry_points_fail_cleanly_without_a_gl_context(cwa, ctx):
    """the interop layer is compiled in (cuda_gl_interop.h); this box has no GL context, so registering a GL name must
    come back as an error code with the CUDA error text -- never a crash -- and the context must stay usable."""
    import ctypes as C
    lib = ctx.lib
    assert lib.cwa_gl_available() == 1
    r = C.c_int(-7)
    assert lib.cwa_gl_register_buffer(ctx.h, 1, C.byref(r)) != 0 and r.value == -1
    assert b"cudaGraphicsGLRegisterBuffer" in lib.cwa_last_error()
    assert lib.cwa_gl_register_image(ctx.h, 1, 0, C.byref(r)) != 0 and r.value == -1
    b = C.c_int()
    assert lib.cwa_gl_map_buffer(ctx.h, 0, C.byref(b)) != 0 and lib.cwa_gl_unmap(ctx.h, 0) != 0 and lib.cwa_gl_unregister(ctx.h, 0) != 0
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 1, cwa.WAVE_COUPLED)
    wave.Compute(2)
    assert np.isfinite(wave.read_role(0)).all()


def test_reinit_from_decoded_png_bytes(cwa, ctx):
    """ReinitFromTexture on an RGBA8 init texture (StencilImage2DTripleBuffered.cpp:61-77, wave_comp.glsl:76-80):
    texel = texelFetch(uInitImage, coord * ivec2(2, 1)) with normalised 8-bit channels (c / 255)."""
    import ctypes as C
    rng = np.random.default_rng(11)
    tex = rng.integers(0, 256, (64, 128, 4), dtype=np.uint8)
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 4, cwa.WAVE_COUPLED)
    cwa.check(ctx.lib.cwa_wave_reinit_from_rgba8(ctx.h, wave.h, tex.ctypes.data_as(C.c_void_p), 128, 64))
    got = wave.read_role(0)
    want = (tex.astype(np.float32) / np.float32(255.0))[:, ::2, :]
    assert np.array_equal(got, want)
