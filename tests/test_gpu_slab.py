"""GPU (one device): the primitives of the multi-GPU decomposition (csrc/multi.cu) against numpy -- message packing (separately
and fused into the integrate pass), unpacking, compaction, copy_if.  The two-GPU run itself is tools/dist_check.py."""
import ctypes as C

import numpy as np
import pytest

from util import jittered_block

pytestmark = pytest.mark.gpu
CAP_MIG, CAP_GHOST = 2048, 4096
MSG = (1 + CAP_MIG + CAP_GHOST)
GRID = ((0.0, -0.02, 0.0), (0.26, 0.18, 0.26), (25, 19, 25))


def _particles(cwa, oracle, seed=3):
    prm = oracle.default_params3()
    prm.upper[0] = prm.upper[2] = 0.25
    p = jittered_block(oracle, 24, 5, 24, prm, seed=seed, vel=0.5)
    p["extras"][:, 3] = np.arange(p.size, dtype=np.float32)            # id
    return prm, p


def _dead(q):
    return (q["pos"][:, 3] == np.float32(-1.0)) & np.isnan(q["pos"][:, 0])


def _read_msg(cwa, buf):
    m = buf.read(cwa.PARTICLE, MSG)
    hdr = m[:1].view(np.int32).ravel()
    return int(hdr[0]), int(hdr[1]), int(hdr[2]), m[1:1 + hdr[0]], m[1 + CAP_MIG:1 + CAP_MIG + hdr[1]]


def _ids(q):
    return np.sort(q["extras"][:, 3].astype(np.int64))


def _expect(p, n_owned, z_lo, z_hi, band):
    z = p["pos"][:n_owned, 2]
    live = ~_dead(p[:n_owned])
    to_l = live & (z < np.float32(z_lo + band))
    to_r = live & ~to_l & (z >= np.float32(z_hi - band))
    return (to_l & (z < z_lo), to_l & ~(z < z_lo), to_r & (z >= z_hi), to_r & ~(z >= z_hi))


def test_pack_selects_migrants_and_ghosts_and_marks_migrants_dead(cwa, ctx, oracle):
    prm, p = _particles(cwa, oracle)
    z_lo, z_hi, band = np.float32(0.05), np.float32(0.15), np.float32(0.02)
    n_owned = p.size - 100                                                # the tail is "ghosts of the last frame": never packed
    buf = cwa.Buffer(ctx, data=p)
    ml, mr = cwa.Buffer(ctx, nbytes=MSG * 64), cwa.Buffer(ctx, nbytes=MSG * 64)
    cwa.check(ctx.lib.cwa_slab_pack(ctx.h, buf.h, n_owned, z_lo, z_hi, band, ml.h, mr.h, CAP_MIG, CAP_GHOST))
    mig_l, gh_l, mig_r, gh_r = _expect(p, n_owned, z_lo, z_hi, band)
    for mbuf, mig, gh in ((ml, mig_l, gh_l), (mr, mig_r, gh_r)):
        nm, ng, ovf, M, G = _read_msg(cwa, mbuf)
        assert ovf == 0 and nm == int(mig.sum()) > 0 and ng == int(gh.sum()) > 0
        assert np.array_equal(_ids(M), _ids(p[:n_owned][mig])) and np.array_equal(_ids(G), _ids(p[:n_owned][gh]))
        by_id = {int(r["extras"][3]): r for r in M}
        k = int(p[:n_owned][mig]["extras"][0, 3])
        assert by_id[k].tobytes() == p[k].tobytes(), "records travel whole"
    after = buf.read(cwa.PARTICLE, p.size)
    assert np.array_equal(_dead(after[:n_owned]), mig_l | mig_r), "migrants are marked dead in place, nothing else changes"
    keep = ~(np.concatenate([mig_l | mig_r, np.zeros(100, bool)]))
    assert np.array_equal(after[keep].view(np.uint8), p[keep].view(np.uint8))
    # one-sided rank (no left neighbour): nothing goes left
    buf.sub_data(p)
    cwa.check(ctx.lib.cwa_slab_pack(ctx.h, buf.h, n_owned, -3e38, z_hi, band, -1, mr.h, CAP_MIG, CAP_GHOST))
    nm, ng, ovf, M, G = _read_msg(cwa, mr)
    assert nm == int(mig_r.sum()) and ng == int(gh_r.sum())


def test_pack_overflow_is_reported_and_leaves_the_particle_in_place(cwa, ctx, oracle):
    prm, p = _particles(cwa, oracle)
    buf = cwa.Buffer(ctx, data=p)
    ml = cwa.Buffer(ctx, nbytes=(1 + 4 + 8) * 64)
    cwa.check(ctx.lib.cwa_slab_pack(ctx.h, buf.h, p.size, np.float32(0.1), np.float32(3e38), np.float32(0.02), ml.h, -1, 4, 8))
    hdr = ml.read(np.int32, 4)
    assert hdr[2] == 1 and hdr[0] > 4 and hdr[1] > 8                      # counts keep counting, the flag is raised
    after = buf.read(cwa.PARTICLE, p.size)
    assert int(_dead(after).sum()) == 4                                   # only the migrants that fitted left


def test_unpack_appends_migrants_then_ghosts_and_reports_counts(cwa, ctx, oracle):
    prm, p = _particles(cwa, oracle)
    n_owned = 1000
    buf = cwa.Buffer(ctx, nbytes=4000 * 64)
    buf.sub_data(p[:n_owned])
    rng = np.random.default_rng(1)

    def message(nm, ng, base):
        m = np.zeros(MSG, cwa.PARTICLE)
        m[:1].view(np.int32).ravel()[:2] = (nm, ng)
        m["extras"][1:1 + nm, 3] = base + np.arange(nm); m["pos"][1:1 + nm, :3] = rng.uniform(0, 1, (nm, 3))
        m["extras"][1 + CAP_MIG:1 + CAP_MIG + ng, 3] = base + 500 + np.arange(ng)
        return m
    rl, rr, sl, sr = message(3, 5, 10000), message(2, 7, 20000), message(4, 0, 30000), message(1, 0, 40000)
    bufs = [cwa.Buffer(ctx, data=m) for m in (rl, rr, sl, sr)]
    counts = (C.c_int * 4)()
    cwa.check(ctx.lib.cwa_slab_unpack(ctx.h, buf.h, n_owned, bufs[0].h, bufs[1].h, bufs[2].h, bufs[3].h, CAP_MIG, CAP_GHOST, counts))
    assert list(counts) == [n_owned + 5, n_owned + 5 + 12 + 5, 0, 5]        # owned, owned + ghosts (5+7 received, 4+1 own sent), flags, adopted
    got = buf.read(cwa.PARTICLE, counts[1])
    ids = got["extras"][n_owned:, 3].astype(int).tolist()
    assert ids == [10000, 10001, 10002, 20000, 20001] + [10500 + i for i in range(5)] + [20500 + i for i in range(7)] + [30000 + i for i in range(4)] + [40000]
    assert np.array_equal(got[:n_owned].view(np.uint8), p[:n_owned].view(np.uint8))
    # capacity exceeded -> flag 2, nothing written
    small = cwa.Buffer(ctx, nbytes=(n_owned + 8) * 64)
    small.sub_data(p[:n_owned])
    cwa.check(ctx.lib.cwa_slab_unpack(ctx.h, small.h, n_owned, bufs[0].h, bufs[1].h, bufs[2].h, bufs[3].h, CAP_MIG, CAP_GHOST, counts))
    assert counts[2] & 2


def test_compact_and_copy_if_are_stable(cwa, ctx, oracle):
    prm, p = _particles(cwa, oracle)
    kill = np.zeros(p.size, bool); kill[::7] = True
    p["pos"][kill] = (np.nan, np.nan, np.nan, -1.0)
    buf, scratch = cwa.Buffer(ctx, data=p), cwa.Buffer(ctx, nbytes=p.size * 64)
    n = C.c_int()
    cwa.check(ctx.lib.cwa_slab_compact(ctx.h, buf.h, p.size, scratch.h, C.byref(n)))
    assert n.value == int((~kill).sum())
    assert np.array_equal(buf.read(cwa.PARTICLE, n.value).view(np.uint8), p[~kill].view(np.uint8)), "order kept"
    prm, q = _particles(cwa, oracle, seed=9)
    src, dst = cwa.Buffer(ctx, data=q), cwa.Buffer(ctx, nbytes=(q.size + 10) * 64)
    x = q["pos"][:, 2]
    for kind, a, b, ref in ((0, 0.05, 0.1, (x >= np.float32(0.05)) & (x < np.float32(0.1))), (1, 0.05, 0.0, x < np.float32(0.05)),
                            (2, 0.12, 0.0, x >= np.float32(0.12)), (3, 0.05, 0.12, ~(x < np.float32(0.05)) & ~(x >= np.float32(0.12)))):
        cwa.check(ctx.lib.cwa_particles_copy_if(ctx.h, src.h, q.size, 2, kind, a, b, dst.h, 10, C.byref(n)))
        assert n.value == int(ref.sum())
        assert np.array_equal(dst.read(cwa.PARTICLE, n.value, offset=10 * 64).view(np.uint8), q[ref].view(np.uint8)), kind


def test_step_slab_packs_what_the_separate_pass_would(cwa, ctx, oracle):
    """cwa_sph_step_slab == cwa_sph_step + cwa_slab_pack: same SSBO, same message sets, and the cwa_slab_pack call that follows is a no-op."""
    prm, p = _particles(cwa, oracle)
    ctx.set_params_from_oracle(prm)
    p["vel"][:, 2] = np.random.default_rng(4).uniform(-60.0, 60.0, p.size).astype(np.float32)     # particles cross the faces
    z_lo, z_hi, band = np.float32(0.06), np.float32(0.14), np.float32(0.02)
    n_owned = p.size - 300                                                                      # the last 300 slots play ghosts

    def run(fused):
        grid = cwa.UniformGrid(ctx, 3, *GRID, p.size)
        sph = cwa.Sph(ctx, p.size, grid, particles=p)
        wave = cwa.StencilImage2DTripleBuffered(ctx, 64, 64, 1, cwa.WAVE_COUPLED)
        sph.bind_wave(wave, wave.role_image(0))
        ml, mr = cwa.Buffer(ctx, nbytes=MSG * 64), cwa.Buffer(ctx, nbytes=MSG * 64)
        if fused:
            cwa.check(ctx.lib.cwa_sph_step_slab(ctx.h, sph.h, n_owned, z_lo, z_hi, band, ml.h, mr.h, CAP_MIG, CAP_GHOST))
            launches = ctx.launch_count
            cwa.check(ctx.lib.cwa_slab_pack(ctx.h, sph.buffer.h, n_owned, z_lo, z_hi, band, ml.h, mr.h, CAP_MIG, CAP_GHOST))
            assert ctx.launch_count == launches, "the pack call after the fused step launches nothing"
        else:
            sph.step(1)
            cwa.check(ctx.lib.cwa_slab_pack(ctx.h, sph.buffer.h, n_owned, z_lo, z_hi, band, ml.h, mr.h, CAP_MIG, CAP_GHOST))
        out = [sph.download()]
        for m in (ml, mr):
            nm, ng, ovf, M, G = _read_msg(cwa, m)
            assert ovf == 0
            out.append((M[np.argsort(M["extras"][:, 3])], G[np.argsort(G["extras"][:, 3])]))
        # a pack with OTHER arguments after a fused step is not elided
        if fused:
            cwa.check(ctx.lib.cwa_sph_step_slab(ctx.h, sph.h, n_owned, z_lo, z_hi, band, ml.h, mr.h, CAP_MIG, CAP_GHOST))
            launches = ctx.launch_count
            cwa.check(ctx.lib.cwa_slab_pack(ctx.h, sph.buffer.h, n_owned, z_lo, z_hi, np.float32(0.03), ml.h, mr.h, CAP_MIG, CAP_GHOST))
            assert ctx.launch_count > launches
        return out

    a, b = run(True), run(False)
    assert np.array_equal(a[0].view(np.uint8), b[0].view(np.uint8)), "SSBO after step + pack (dead marks included)"
    for side in (1, 2):
        for part in (0, 1):
            assert a[side][part].size == b[side][part].size and a[side][part].size > 0
            assert np.array_equal(a[side][part].view(np.uint8), b[side][part].view(np.uint8))
