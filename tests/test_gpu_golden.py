"""GPU parity against the COMMITTED golden fixtures (tests/golden/golden_v1.npz): the CUDA path is
checked against recorded oracle outputs, so the comparison does not depend on the oracle binary that
happens to be on the GPU box."""
import os

import numpy as np
import pytest

from util import assert_close

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
SMALL_GRID3 = ((0.0, -0.02, 0.0), (0.12, 0.1, 0.12), (6, 6, 6))


def _small_box(ctx):
    ctx.set_boundary(upper=(0.11, 1.0, 0.11, 500.0), lower=(0.0, -0.02, 0.0, 50.0))


def test_scan_goldens(cwa, ctx):
    for k in ("kat1", "kat2", "rand"):
        x = GOLD[f"scan_{k}_in"]
        src = cwa.Buffer(ctx, data=x)
        dst = cwa.Buffer(ctx, nbytes=(x.size + 1) * 4)
        ctx.scan_exclusive(src, dst, x.size)
        assert np.array_equal(dst.read(np.int32, x.size), GOLD[f"scan_{k}_out"]), k


def test_grid_goldens(cwa, ctx):
    rows = np.zeros((512, 12), np.float32); rows[:, :2] = GOLD["grid2_pos"]
    g2 = cwa.UniformGrid(ctx, 2, (0.0, 0.0), (9.6, 9.6), (32, 32), 512)
    g2.build(cwa.Buffer(ctx, data=rows), 48, 512)
    assert np.array_equal(g2.read(cwa.GRID_CELL_OF, 512), GOLD["grid2_cell"])
    assert np.array_equal(g2.read(cwa.GRID_COUNTER, 1024), GOLD["grid2_cnt"])
    assert np.array_equal(g2.read(cwa.GRID_OFFSET, 1024), GOLD["grid2_off"])
    m = int(GOLD["grid2_cnt"].sum())
    assert np.array_equal(g2.read(cwa.GRID_INDEX_LIST, m), GOLD["grid2_idx"][:m])
    pos = GOLD["grid3_pos"]
    g3 = cwa.UniformGrid(ctx, 3, *SMALL_GRID3, pos.shape[0])
    g3.build(cwa.Buffer(ctx, data=np.ascontiguousarray(pos)), 16, pos.shape[0])
    assert np.array_equal(g3.read(cwa.GRID_CELL_OF, pos.shape[0]), GOLD["grid3_cell"])
    assert np.array_equal(g3.read(cwa.GRID_COUNTER, 216), GOLD["grid3_cnt"])
    assert np.array_equal(g3.read(cwa.GRID_OFFSET, 216), GOLD["grid3_off"])
    m = int(GOLD["grid3_cnt"].sum())
    assert np.array_equal(g3.read(cwa.GRID_INDEX_LIST, m), GOLD["grid3_idx"][:m])


def test_wave_goldens_bit_exact(cwa, ctx):
    for wtype in (1.0, 0.0, 0.5):
        ctx.set_wave_uniforms(attributes=(0.01, 0.985, 0.001, wtype))
        w = cwa.StencilImage2DTripleBuffered(ctx, 32, 32, 1, cwa.WAVE_COUPLED)
        assert np.array_equal(w.read_role(0), GOLD[f"wave_init_t{wtype}"]), wtype
    ctx.set_wave_uniforms()
    w = cwa.StencilImage2DTripleBuffered(ctx, 48, 32, 1, cwa.WAVE_SIMP)
    assert np.array_equal(w.read_role(0), GOLD["wave_init_simp"])
    u0, u1 = GOLD["wave_u0"], GOLD["wave_u1"]
    for variant, key, a0, a1 in ((0, "wave_step_coupled", u0, u1), (1, "wave_step_simp", u0, u1)):
        w = cwa.StencilImage2DTripleBuffered(ctx, 52, 40, 1, variant)
        w.write_role(0, a0); w.write_role(1, a1)
        w.Compute(1)
        assert np.array_equal(w.read_role(0).view(np.uint32), GOLD[key].view(np.uint32)), key
    ctx.set_wave_uniforms(attributes=(0.01, 0.985, 0.001, 0.5))
    w = cwa.StencilImage2DTripleBuffered(ctx, 52, 40, 1, 0)
    w.write_role(0, np.abs(u0)); w.write_role(1, 0.1 * np.abs(u1))
    w.Compute(1)
    assert np.array_equal(w.read_role(0).view(np.uint32), GOLD["wave_step_wake"].view(np.uint32))
    ctx.set_wave_uniforms()
    w = cwa.StencilImage2DTripleBuffered(ctx, 32, 32, 1, 0)
    w.Compute(25)
    assert np.array_equal(w.read_role(0).view(np.uint32), GOLD["wave_25_steps"].view(np.uint32))


@pytest.mark.parametrize("use_grid", [False, True])
def test_sph3_pass_goldens(cwa, ctx, use_grid):
    _small_box(ctx)
    p = GOLD["sph3_in"]
    grid = cwa.UniformGrid(ctx, 3, *SMALL_GRID3, p.size) if use_grid else None
    sph = cwa.Sph(ctx, p.size, grid, particles=p.copy())
    wave = cwa.StencilImage2DTripleBuffered(ctx, 32, 32, 1, cwa.WAVE_COUPLED)
    wave.write_image(0, GOLD["sph3_tex"])
    sph.bind_wave(wave, 0)
    assert np.array_equal(sph.neighbour_count(), GOLD["sph3_neighbours"])
    sph.rho_pres()
    got = sph.download()
    assert_close(got["extras"][:, :2], GOLD["sph3_after_rho"]["extras"][:, :2], what="rho/p")
    sph.force()
    got = sph.download()
    assert_close(got["force"][:, :3], GOLD["sph3_after_force"]["force"][:, :3], what="force")
    assert np.array_equal(got["force"][:, 3], GOLD["sph3_after_force"]["force"][:, 3])
    sph.integrate()
    got = sph.download()
    ref = GOLD["sph3_after_integrate"]
    assert_close(got["pos"][:, :3], ref["pos"][:, :3], scale=1e-3, what="pos")
    assert_close(got["vel"][:, :3], ref["vel"][:, :3], what="vel")


@pytest.mark.parametrize("name,mode", [("as_shipped", 0), ("latest", 1)])
@pytest.mark.parametrize("use_grid", [False, True])
def test_coupled_goldens(cwa, ctx, name, mode, use_grid):
    _small_box(ctx)
    start = GOLD["coupled_start"]
    grid = cwa.UniformGrid(ctx, 3, *SMALL_GRID3, start.size) if use_grid else None
    sph = cwa.Sph(ctx, start.size, grid, particles=start.copy())
    wave = cwa.StencilImage2DTripleBuffered(ctx, 32, 32, 1, cwa.WAVE_COUPLED)
    sph.coupled_step(wave, 6, mode)
    got, ref = sph.download(), GOLD[f"coupled_{name}_particles"]
    assert np.array_equal(wave.read_role(0).view(np.uint32), GOLD[f"coupled_{name}_wave"].view(np.uint32))
    assert_close(got["extras"][:, 0], ref["extras"][:, 0], rtol=1e-3, what="rho after 6 frames")
    assert_close(got["pos"][:, :3], ref["pos"][:, :3], rtol=1e-3, scale=1e-3, what="pos after 6 frames")
    assert_close(got["vel"][:, :3], ref["vel"][:, :3], rtol=1e-3, what="vel after 6 frames")


@pytest.mark.parametrize("variant", [0, 1])
def test_sph2_goldens(cwa, ctx, variant):
    p = GOLD[f"sph2_v{variant}_in"]
    grid = cwa.UniformGrid(ctx, 2, (0.0, 0.0), (9.6, 9.6), (32, 32), p.size)
    s = cwa.SphUgrid(ctx, p.size, grid, variant, substeps=2)
    s.upload(p.copy())
    s.Compute(1)
    got, ref = s.download(), GOLD[f"sph2_v{variant}_out"]
    assert_close(got["acc"][:, 3], ref["acc"][:, 3], what="rho")
    assert_close(got["pos"][:, :2], ref["pos"][:, :2], scale=1e-2, what="pos")
    assert_close(got["vel"][:, :2], ref["vel"][:, :2], what="vel")
