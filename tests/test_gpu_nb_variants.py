"""GPU parity of every neighbour-kernel variant (cwa_set_tuning: "lanes" kernels 0..6, neighbour-list kernels 7)
on several grids -- cells of 2h, cells slightly larger than h (27-cell queries), cells of exactly h and
cells smaller than h (generic wide query) -- with full, partial and no shared-memory staging, against the
all-pairs oracle.  Tolerance 1e-4 relative (summation order), neighbour sets identical."""
import numpy as np
import pytest

from test_gpu_sph3 import NX, NY, NZ, _params
from util import assert_close, jittered_block, smooth_field

pytestmark = pytest.mark.gpu

GRIDS = {
    "2h": ((0.0, -0.02, 0.0), (0.26, 0.18, 0.26), (13, 10, 13)),
    "h+": ((0.0, -0.02, 0.0), (0.26, 0.18, 0.26), (25, 19, 25)),        # cells 0.0104 x 0.0105: 3 x 3 rows of 3 cells
    "h": ((0.0, -0.02, 0.0), (0.26, 0.18, 0.26), (26, 20, 26)),         # cells of exactly h: conservative ranges may reach 4 cells
    "h/1.5": ((0.0, -0.02, 0.0), (0.26, 0.18, 0.26), (39, 30, 39)),     # h > cell: generic (wide) query
    "columns": ((0.0, -0.02, 0.0), (0.26, 0.18, 0.26), (25, 1, 25)),    # ONE layer in y: cells are columns over the sheet, 3 x 3 columns per query
}


@pytest.fixture()
def tuned(ctx):
    yield ctx
    ctx.set_tuning(nb_config=8, nb_cap_d=2048, nb_cap_f=1536)


def _scene(cwa, ctx, oracle, grid_key, cluster=False, seed=1234):
    prm = _params(oracle)
    ctx.set_params_from_oracle(prm)
    p = jittered_block(oracle, NX, NY, NZ, prm, seed=seed, vel=0.5)
    rng = np.random.default_rng(seed + 1)
    p["force"] = rng.uniform(-1e4, 1e4, (p.size, 4)).astype(np.float32)
    p["pos"][::7, 1] += np.float32(0.02)
    if cluster:
        # 300 particles inside a ball of radius 0.004: > 24 neighbours each (list drains), one very long row
        c = np.array([0.101, 0.021, 0.099], np.float32)
        d = rng.normal(size=(300, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        p["pos"][1000:1300, :3] = c + (d * rng.uniform(0.0005, 0.004, (300, 1))).astype(np.float32)
    tex = smooth_field(64, 64, 1, amp=0.02)
    grid = cwa.UniformGrid(ctx, 3, *GRIDS[grid_key], p.size)
    sph = cwa.Sph(ctx, p.size, grid, particles=p)
    wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 1, cwa.WAVE_COUPLED)
    wave.write_image(0, tex)
    sph.bind_wave(wave, 0)
    return prm, p, tex, sph


def _check(oracle, prm, p, tex, sph):
    sph.rho_pres()
    sph.force()
    got = sph.download()
    ref = p.copy()
    oracle.sph3_rho_pres(ref, prm, tex)
    oracle.sph3_force(ref, prm, tex)
    assert_close(got["extras"][:, 0], ref["extras"][:, 0], what="rho")
    assert_close(got["extras"][:, 1], ref["extras"][:, 1], what="pressure")
    assert_close(got["force"][:, :3], ref["force"][:, :3], what="force.xyz")


@pytest.mark.parametrize("grid_key", list(GRIDS))
@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_variant_matches_oracle(cwa, tuned, oracle, cfg, grid_key):
    tuned.set_tuning(nb_config=cfg)
    prm, p, tex, sph = _scene(cwa, tuned, oracle, grid_key)
    got = sph.neighbour_count()
    assert np.array_equal(got, oracle.sph3_neighbour_count(p, 0.01))
    _check(oracle, prm, p, tex, sph)


@pytest.mark.parametrize("cap", [0, 96, 1536])
@pytest.mark.parametrize("grid_key", ["2h", "h+"])
@pytest.mark.parametrize("cfg", [1, 7, 8])
def test_variant_with_partial_or_no_staging(cwa, tuned, oracle, cfg, grid_key, cap):
    tuned.set_tuning(nb_config=cfg, nb_cap_d=cap, nb_cap_f=cap)
    prm, p, tex, sph = _scene(cwa, tuned, oracle, grid_key)
    _check(oracle, prm, p, tex, sph)


@pytest.mark.parametrize("grid_key", ["2h", "h+", "h/1.5"])
@pytest.mark.parametrize("cfg", [1, 7, 8])
def test_variant_dense_cluster(cwa, tuned, oracle, cfg, grid_key):
    """A 300-particle clump: neighbour counts far above the neighbour-list capacity (K = 64 -> the force pass
    re-scans the grid for those targets) and rows that overflow the staging budget of the lanes kernels."""
    tuned.set_tuning(nb_config=cfg)
    prm, p, tex, sph = _scene(cwa, tuned, oracle, grid_key, cluster=True)
    nb = oracle.sph3_neighbour_count(p, 0.01)
    assert nb.max() >= 300
    _check(oracle, prm, p, tex, sph)


@pytest.mark.parametrize("grid_key", ["h+", "columns"])
def test_queued_clump_targets_eight_lanes_each(cwa, tuned, oracle, grid_key):
    """The force pass of queued clump targets has a second shape -- four targets per warp, eight lanes each -- that the library takes when
    the queue is longer than the launch (never in a scene the oracle finishes in seconds); heavy_sub_warp = 2 forces it.  Same clump as
    above against the oracle, with targets beyond `inplace_max` candidates (whole-warp fallback inside that path) and a thin queue."""
    try:
        tuned.set_tuning(nb_config=8, heavy_sub_warp=2, extreme_candidates=40, inplace_max=250)
        prm, p, tex, sph = _scene(cwa, tuned, oracle, grid_key, cluster=True)
        nb = oracle.sph3_neighbour_count(p, 0.01)
        assert nb.max() >= 300
        _check(oracle, prm, p, tex, sph)
        sph.upload(p)
        sph.step(3)                                            # and through full frames: the one-warp-per-target shape is the reference
        a = sph.download()
        tuned.set_tuning(heavy_sub_warp=0)
        sph.upload(p)
        sph.step(3)
        b = sph.download()
        assert_close(a["pos"][:, :3], b["pos"][:, :3], scale=1e-3, rtol=1e-3, what="pos")
        assert_close(a["vel"][:, :3], b["vel"][:, :3], rtol=1e-3, what="vel")
    finally:
        tuned.set_tuning(heavy_sub_warp=1, extreme_candidates=192, inplace_max=640)


@pytest.mark.parametrize("cfg", [7, 8])
def test_rows_variant_full_frames_equal_lanes_variant(cwa, tuned, oracle, cfg):
    """3 fused frames: rows kernels vs the lanes kernels (same physics, summation order differs)."""
    tuned.set_tuning(nb_config=1)
    prm, p, tex, sph = _scene(cwa, tuned, oracle, "h+")
    sph.step(3)
    a = sph.download()
    tuned.set_tuning(nb_config=cfg)
    sph.upload(p)
    sph.step(3)
    b = sph.download()
    assert_close(b["pos"][:, :3], a["pos"][:, :3], scale=1e-3, rtol=1e-3, what="pos")
    assert_close(b["vel"][:, :3], a["vel"][:, :3], rtol=1e-3, what="vel")


@pytest.mark.parametrize("cluster", [False, True])
@pytest.mark.parametrize("fused", [0, 1])
def test_fused_order_reorder_is_canonical(cwa, tuned, oracle, fused, cluster):
    """The SPH snapshot path may fuse the canonical per-cell ordering into the reorder pass; the index list
    it leaves behind must be the same bit-exact list (ascending id inside a cell) as the stand-alone build."""
    tuned.set_tuning(fused_order=fused, nb_config=8)
    try:
        prm, p, tex, sph = _scene(cwa, tuned, oracle, "h+", cluster=cluster)
        p["pos"][5, 0] = np.nan                    # a NaN particle is left out of the grid
        sph.upload(p)
        sph.rho_pres()
        g = oracle.grid3(*GRIDS["h+"])
        cell_of, cnt, off, idx = oracle.grid3_build(g, p["pos"])
        n_ins = int(cnt.sum())
        assert n_ins == p.size - 1
        assert np.array_equal(sph.grid.read(cwa.GRID_COUNTER, cnt.size), cnt)
        assert np.array_equal(sph.grid.read(cwa.GRID_OFFSET, off.size), off)
        assert np.array_equal(sph.grid.read(cwa.GRID_INDEX_LIST, n_ins), idx[:n_ins])
        sph.force()
        got = sph.download()
        ref = p.copy()
        oracle.sph3_rho_pres(ref, prm, tex, grid=(g, cnt, off, idx))
        oracle.sph3_force(ref, prm, tex, grid=(g, cnt, off, idx))
        ok = np.arange(p.size) != 5
        assert_close(got["extras"][ok, 0], ref["extras"][ok, 0], what="rho")
        assert_close(got["force"][ok, :3], ref["force"][ok, :3], what="force.xyz")
    finally:
        tuned.set_tuning(fused_order=1)


@pytest.mark.parametrize("cluster", [False, True])
def test_fused_integrate_tail_equals_separate_kernels(cwa, tuned, oracle, cluster):
    """Full step: the force kernels may finish the particle themselves (force epilogue + integrate + record
    write-back); the result must be the one of the separate integrate kernel (same device functions)."""
    try:
        tuned.set_tuning(nb_config=8, fused_integrate=0)
        prm, p, tex, sph = _scene(cwa, tuned, oracle, "h+", cluster=cluster)
        sph.step(2)
        a = sph.download()
        tuned.set_tuning(fused_integrate=1)
        sph.upload(p)
        sph.step(2)
        b = sph.download()
        # (clump targets: the separate path sums a queued target's pairs over 8 lanes, the fused one over 32 -- another order of the same terms)
        for f in ("pos", "vel", "force", "extras"):
            assert_close(b[f], a[f], rtol=2e-5 if cluster else 5e-6, what=f)
        ref = p.copy()
        for _ in range(2):
            oracle.sph3_rho_pres(ref, prm, tex); oracle.sph3_force(ref, prm, tex); oracle.sph3_integrate(ref, prm, tex)
        if not cluster:
            assert_close(b["pos"][:, :3], ref["pos"][:, :3], scale=1e-3, rtol=1e-3, what="pos vs oracle")
    finally:
        tuned.set_tuning(fused_integrate=0)


@pytest.mark.parametrize("coupling", ["as_shipped", "latest"])
@pytest.mark.parametrize("cluster", [False, True])
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_pipelined_frames_equal_the_plain_sequence(cwa, tuned, oracle, mode, cluster, coupling):
    """cwa_coupled_step(K > 1) pipelines the frames of one call (wave stencil of frame f on a side stream next to
    the grid build of frame f+1; integrate of frame f counts the particles into the cells of frame f+1).  Nothing
    an application can read afterwards may differ from the plain sequence: particle records, all three wave
    levels and the triple-buffer state, and every grid array -- bit for bit."""
    cpl = cwa.COUPLING_AS_SHIPPED if coupling == "as_shipped" else cwa.COUPLING_LATEST
    frames = 7

    def run(pipeline):
        tuned.set_tuning(nb_config=8, pipeline=pipeline)
        prm, p, tex, sph = _scene(cwa, tuned, oracle, "h+", cluster=cluster)
        p["pos"][5, 0] = np.nan                    # never inserted
        p["vel"][::11, :3] *= np.float32(40.0)     # some particles change cells every frame
        sph.upload(p)
        wave = cwa.StencilImage2DTripleBuffered(tuned, 96, 64, 1, cwa.WAVE_COUPLED)
        sph.coupled_step(wave, frames, cpl)
        g = sph.grid
        ncell = g.num_cells_total
        out = dict(p=sph.download(), state=wave.state(), waves=[wave.read_image(i) for i in range(3)],
                   counter=g.read(cwa.GRID_COUNTER, ncell), offset=g.read(cwa.GRID_OFFSET, ncell + 1),
                   cell_of=g.read(cwa.GRID_CELL_OF, p.size))
        n_ins = int(out["offset"][-1])
        out["index"] = g.read(cwa.GRID_INDEX_LIST, n_ins)
        # one more call continues from the pipelined state exactly like from the plain one
        sph.coupled_step(wave, 2, cpl)
        out["p2"] = sph.download()
        return out

    try:
        a = run(0)
        b = run(mode)
    finally:
        tuned.set_tuning(pipeline=3)
    assert a["state"] == b["state"]
    for i in range(3):
        assert np.array_equal(a["waves"][i].view(np.uint32), b["waves"][i].view(np.uint32)), f"wave image {i}"
    for k in ("counter", "offset", "cell_of", "index"):
        assert np.array_equal(a[k], b[k]), k
    assert int(a["counter"].sum()) == int(a["offset"][-1])
    for k in ("p", "p2"):
        for f in ("pos", "vel", "force", "extras"):
            assert np.array_equal(a[k][f].view(np.uint32), b[k][f].view(np.uint32)), f"{k}.{f}"


@pytest.mark.parametrize("coupling", ["as_shipped", "latest"])
def test_transposed_sampling_copy_is_bit_identical(cwa, tuned, oracle, coupling):
    """The SPH passes may sample a transposed copy of the bound wave level (the grid runs fastest along z = the
    texture's t axis).  Same texels, same arithmetic: every particle bit must agree with direct sampling, across
    wave steps, image uploads and buffer-level writes that change the sampled level."""
    cpl = cwa.COUPLING_AS_SHIPPED if coupling == "as_shipped" else cwa.COUPLING_LATEST

    def run(on):
        tuned.set_tuning(nb_config=8, wave_transpose=on, pipeline=3)
        prm, p, tex, sph = _scene(cwa, tuned, oracle, "h+")
        wave = cwa.StencilImage2DTripleBuffered(tuned, 320, 272, 1, cwa.WAVE_COUPLED)   # non-square, above the 256^2 threshold
        for i in range(3):
            wave.write_image(i, smooth_field(272, 320, 1, amp=0.01 * (i + 1)))
        wave.bind_texture_unit()
        outs = []
        sph.coupled_step(wave, 5, cpl)
        outs.append(sph.download())
        st = wave.state()
        bound = st["tex_unit0"] if st["tex_unit0"] >= 0 else wave.role_image(0)
        wave.write_image(bound, smooth_field(272, 320, 1, amp=0.025))      # the sampled level changes behind the copy
        sph.coupled_step(wave, 1, cpl)
        outs.append(sph.download())
        sph.bind_wave(wave, wave.role_image(0))
        sph.rho_pres(); sph.force(); sph.integrate()          # separately dispatched passes sample the same copy
        outs.append(sph.download())
        outs.append(wave.read_role(0))
        return outs

    try:
        a = run(0)
        b = run(1)
    finally:
        tuned.set_tuning(wave_transpose=1)
    for k in range(3):
        for f in ("pos", "vel", "force", "extras"):
            assert np.array_equal(a[k][f].view(np.uint32), b[k][f].view(np.uint32)), f"stage {k} {f}"
    assert np.array_equal(a[3].view(np.uint32), b[3].view(np.uint32))


@pytest.mark.parametrize("coupling", ["as_shipped", "latest"])
def test_pipelined_frames_all_pairs_mode(cwa, tuned, oracle, coupling):
    """The as-shipped configuration (no uniform grid): the all-pairs passes sample the wave field from their first kernel on, so the
    wait for the previous frame's stencil (side stream) sits in front of them.  Pipelined == plain, bit for bit."""
    cpl = cwa.COUPLING_AS_SHIPPED if coupling == "as_shipped" else cwa.COUPLING_LATEST

    def run(pipeline):
        tuned.set_tuning(pipeline=pipeline)
        prm = _params(oracle)
        tuned.set_params_from_oracle(prm)
        p = jittered_block(oracle, NX, NY, NZ, prm, seed=11, vel=0.5)
        sph = cwa.Sph(tuned, p.size, None, particles=p)
        wave = cwa.StencilImage2DTripleBuffered(tuned, 1024, 768, 1, cwa.WAVE_COUPLED)
        for i in range(3):
            wave.write_image(i, smooth_field(768, 1024, 1, amp=0.01 * (i + 1)))
        sph.coupled_step(wave, 6, cpl)
        return sph.download(), [wave.read_image(i) for i in range(3)], wave.state()

    try:
        a = run(0)
        b = run(3)
    finally:
        tuned.set_tuning(pipeline=3)
    assert a[2] == b[2]
    for f in ("pos", "vel", "force", "extras"):
        assert np.array_equal(a[0][f].view(np.uint32), b[0][f].view(np.uint32)), f
    for x, y in zip(a[1], b[1]):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
