"""GPU: checkpoint / resume and parameter reflection (8f-4)."""
import os

import numpy as np
import pytest

from util import jittered_block

pytestmark = pytest.mark.gpu
GRID = ((0.0, -0.02, 0.0), (0.26, 0.18, 0.26), (25, 19, 25))


def _scene(cwa, ctx, oracle):
    prm = oracle.default_params3()
    prm.upper[0] = prm.upper[2] = 0.25
    ctx.set_params_from_oracle(prm)
    p = jittered_block(oracle, 24, 5, 24, prm, seed=7, vel=0.5)
    grid = cwa.UniformGrid(ctx, 3, *GRID, p.size)
    sph = cwa.Sph(ctx, p.size, grid, particles=p)
    wave = cwa.StencilImage2DTripleBuffered(ctx, 96, 64, 1, cwa.WAVE_COUPLED)
    return sph, wave


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


@pytest.mark.parametrize("coupling", [0, 1])
def test_checkpoint_resumes_bit_for_bit(cwa, oracle, tmp_path, coupling):
    from coupledwateranimation_b200 import checkpoint
    path = str(tmp_path / "run.ckpt")
    with cwa.Context(0) as a:
        sph, wave = _scene(cwa, a, oracle)
        a.param_set("visc", 2500.0)
        sph.coupled_step(wave, 7, coupling)           # 7 frames: the triple buffer and the stale-texture schedule are mid-cycle
        a.checkpoint_save(sph, wave, 7, path)
        saved_p, saved_state = sph.download(), wave.state()
        saved_imgs = [wave.read_image(i) for i in range(3)]
        sph.coupled_step(wave, 5, coupling)
        ref_p, ref_w, ref_state = sph.download(), [wave.read_image(i) for i in range(3)], wave.state()
    ck = checkpoint.read(path)                        # the pure-numpy reader sees what the device held
    assert ck["frame"] == 7 and np.array_equal(_bits(ck["particles"]), _bits(saved_p))
    assert all(np.array_equal(_bits(x), _bits(y)) for x, y in zip(ck["images"], saved_imgs))
    assert list(ck["header"]["unit"]) == saved_state["unit"] and int(ck["header"]["tex_unit0"]) == saved_state["tex_unit0"]
    assert ck["header"]["constants"][2] == np.float32(2500.0)
    with cwa.Context(0) as b:                         # a fresh context with default parameters and a fresh scene
        sph, wave = _scene(cwa, b, oracle)
        sph.coupled_step(wave, 2, coupling)           # ... in some other state
        assert b.checkpoint_load(sph, wave, path) == 7
        assert b.param_get("visc") == 2500.0
        assert wave.state() == saved_state
        sph.coupled_step(wave, 5, coupling)
        assert wave.state() == ref_state
        assert np.array_equal(_bits(sph.download()), _bits(ref_p))
        for i in range(3):
            assert np.array_equal(_bits(wave.read_image(i)), _bits(ref_w[i]))
        # a checkpoint written by the numpy writer loads the same way
        path2 = str(tmp_path / "made.ckpt")
        checkpoint.write(path2, 3, ck["particles"], ck["images"], read_index=ck["header"]["read_index"], write_index=int(ck["header"]["write_index"]),
                         unit=ck["header"]["unit"], tex_unit0=int(ck["header"]["tex_unit0"]), constants=ck["header"]["constants"],
                         boundary=ck["header"]["boundary"], wave=ck["header"]["wave"], sim=ck["header"]["sim"])
        assert b.checkpoint_load(sph, wave, path2) == 3
        sph.coupled_step(wave, 5, coupling)
        assert np.array_equal(_bits(sph.download()), _bits(ref_p))


def test_checkpoint_rejects_mismatched_or_broken_files(cwa, ctx, oracle, tmp_path):
    sph, wave = _scene(cwa, ctx, oracle)
    path = str(tmp_path / "a.ckpt")
    ctx.checkpoint_save(sph, wave, 0, path)
    other = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 1, cwa.WAVE_COUPLED)
    with pytest.raises(cwa.CwaError, match="checkpoint holds"):
        ctx.checkpoint_load(sph, other, path)
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[: len(raw) // 2])
    with pytest.raises(cwa.CwaError, match="truncated"):
        ctx.checkpoint_load(sph, wave, path)
    open(path, "wb").write(b"not a checkpoint" * 32)
    with pytest.raises(cwa.CwaError, match="not a version-1 checkpoint"):
        ctx.checkpoint_load(sph, wave, path)
    with pytest.raises(cwa.CwaError, match="cannot open"):
        ctx.checkpoint_load(sph, wave, os.path.join(str(tmp_path), "missing.ckpt"))


def test_parameter_table_covers_the_blocks_and_edits_take_effect(cwa, ctx, oracle):
    names = [n for n, *_ in ctx.params()]
    for must in ("mass", "smoothing_coeff", "visc", "resting_rho", "gas_const", "dt", "gravity_y", "upper.x", "lower.y", "wave.atten", "torque_coeff"):
        assert must in names
    assert len(set(names)) == len(names)
    # defaults == the reference's values (Main.cpp:184-204 + shader constants)
    for n, v in (("mass", 0.02), ("smoothing_coeff", 2.0), ("visc", 3000.0), ("resting_rho", 1000.0), ("upper.x", 0.48), ("lower.y", -0.02),
                 ("wave.lambda", 0.01), ("wave.atten", 0.985), ("gas_const", 4000.0), ("dt", 0.00005), ("gravity_y", -9806.65), ("uv_scale", 2.0)):
        assert ctx.param_get(n) == np.float32(v), n
    with pytest.raises(cwa.CwaError, match="unknown parameter"):
        ctx.param_set("no_such_parameter", 1.0)
    # an edit through the table is what the kernels see: same result as writing the block directly
    sph, wave = _scene(cwa, ctx, oracle)
    p0 = sph.download()
    ctx.param_set("dt", 0.00002); ctx.param_set("gravity_y", -5000.0)
    sph.coupled_step(wave, 2, cwa.COUPLING_LATEST)
    a = sph.download()
    prm = oracle.default_params3()
    prm.upper[0] = prm.upper[2] = 0.25
    prm.dt = 0.00002; prm.gravity_y = -5000.0
    ctx.set_params_from_oracle(prm)
    sph.upload(p0)
    wave2 = cwa.StencilImage2DTripleBuffered(ctx, 96, 64, 1, cwa.WAVE_COUPLED)
    sph.coupled_step(wave2, 2, cwa.COUPLING_LATEST)
    assert np.array_equal(_bits(a), _bits(sph.download()))
