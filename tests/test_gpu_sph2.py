"""GPU parity: 2-D Koschier SPH on the uniform grid (SphWave2D, config C2) against the oracle."""
import numpy as np
import pytest

from util import assert_close

pytestmark = pytest.mark.gpu

EXT = ((0.0, 0.0), (9.6, 9.6), (32, 32))          # SphUgrid ctor, StencilBuffer.cpp:138


def _wave1d(width, seed=0):
    x = np.linspace(0.0, 1.0, width, dtype=np.float32)
    w = np.zeros((width, 4), np.float32)
    w[:, 0] = 4.8 + 0.96 * np.exp(-(x - 0.5) ** 2 / 0.005)          # h   (Shallow1D InitWave shape)
    w[:, 1] = 0.5 * 0.96 * np.exp(-(x - 0.5) ** 2 / 0.005) * np.sign(x - 0.5)   # uh
    return w


@pytest.mark.parametrize("variant", [0, 1])
def test_init_lattice(cwa, ctx, oracle, variant):
    n = 4096
    grid = cwa.UniformGrid(ctx, 2, *EXT, n)
    s = cwa.SphUgrid(ctx, n, grid, variant)
    prm = oracle.default_params2(variant)
    ref = oracle.sph2_init(n, prm)
    assert np.array_equal(s.download().view(np.uint8), ref.view(np.uint8))


@pytest.mark.parametrize("variant,bound", [(0, False), (1, False), (1, True)])
def test_one_frame_two_substeps(cwa, ctx, oracle, variant, bound):
    n = 4096
    prm = oracle.default_params2(variant)
    prm.time = 0.25
    grid = cwa.UniformGrid(ctx, 2, *EXT, n)
    s = cwa.SphUgrid(ctx, n, grid, variant, substeps=2)
    s.set_uniforms(time=0.25)
    p0 = oracle.sph2_init(n, prm)
    rng = np.random.default_rng(5)
    p0["pos"][:, :2] += rng.uniform(-0.004, 0.004, (n, 2)).astype(np.float32)
    p0["vel"][:, :2] = rng.uniform(-0.3, 0.3, (n, 2)).astype(np.float32)
    p0["pos"][::50, 0] -= np.float32(0.2)          # a few particles outside the extents / inside the wall
    s.upload(p0)
    w1d = _wave1d(128) if bound else None
    if bound:
        wb = cwa.Buffer(ctx, data=w1d)
        s.bind_wave1d(wb, 128)
    s.Compute(1)
    got = s.download()
    b0, b1 = p0.copy(), np.zeros_like(p0)
    g = oracle.grid2(*EXT)
    r, _ = oracle.sph2_step(b0, b1, 0, 2, prm, w1d.reshape(1, 128, 4) if bound else None, g)
    ref = (b0, b1)[r]
    assert_close(got["acc"][:, 3], ref["acc"][:, 3], what="rho")
    assert_close(got["vel"][:, 3], ref["vel"][:, 3], what="pressure")
    assert_close(got["pos"][:, :2], ref["pos"][:, :2], scale=1e-2, what="pos")
    assert_close(got["vel"][:, :2], ref["vel"][:, :2], what="vel")
    assert_close(got["acc"][:, :2], ref["acc"][:, :2], what="acc")
    assert np.array_equal(got["pos"][:, 3], ref["pos"][:, 3]), "enable/disable flag"


def test_c2_size_64k_particles_runs_and_conserves_count(cwa, ctx, oracle):
    # 512 x 128 lattice at the shipped pitch -> tank width 38.4, 128 x 32 cells of 0.3
    n = 65536
    ext = ((0.0, 0.0), (38.4, 9.6), (128, 32))
    grid = cwa.UniformGrid(ctx, 2, *ext, n)
    s = cwa.SphUgrid(ctx, n, grid, cwa.SPH2_WAVE, substeps=2)
    s.set_uniforms(init_width=512, view_width=38.4)
    s.Reinit()
    prm = oracle.default_params2(1)
    prm.init_width = 512
    prm.view_width = 38.4
    ref0 = oracle.sph2_init(n, prm)
    assert np.array_equal(s.download().view(np.uint8), ref0.view(np.uint8))
    s.Compute(3)
    got = s.download()
    assert got.size == n and np.isfinite(got["pos"][:, :2]).all()
    b0, b1 = ref0.copy(), np.zeros_like(ref0)
    g = oracle.grid2(*ext)
    r = 0
    for _ in range(3):
        r, _g = oracle.sph2_step(b0, b1, r, 2, prm, None, g)
    ref = (b0, b1)[r]
    assert_close(got["acc"][:, 3], ref["acc"][:, 3], rtol=1e-3, what="rho after 3 frames")
    assert_close(got["pos"][:, :2], ref["pos"][:, :2], rtol=1e-3, scale=1e-2, what="pos after 3 frames")


def test_frame_graph_replays_the_same_frames(cwa, ctx, oracle):
    """cwa_sph2_compute captures a CUDA graph of the frame the second time the same frame is asked for and replays it while buffers, uniforms
    and grid stay put; a changed uniform falls back to plain launches (and a new capture).  Bit-identical to the plain launch sequence."""
    n = 4096
    ext = ((0.0, 0.0), (9.6, 9.6), (32, 32))

    def run(graph):
        ctx.set_tuning(graph=graph)
        grid = cwa.UniformGrid(ctx, 2, *ext, n)
        s = cwa.SphUgrid(ctx, n, grid, cwa.SPH2_WAVE, substeps=2)
        s.Reinit()
        s.Compute(5)                       # frames 3..5 are replays when graphs are on
        s.set_uniforms(bottom=0.35)        # a uniform changes: plain launch, then a new capture
        s.Compute(4)
        s.Compute(1)
        out = s.download()
        s.destroy(); grid.destroy()
        return out

    try:
        a, b = run(0), run(1)
    finally:
        ctx.set_tuning(graph=1)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert np.isfinite(a["pos"][:, :2]).all()
