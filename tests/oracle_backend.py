"""CPU stand-in for coupledwateranimation_b200.distributed.CudaBackend, built on the oracle.
TEST INFRASTRUCTURE ONLY: lets the multi-rank protocol (ghost width, migration, wave halos, global
last row, texture schedule) run on CPU with the gloo backend."""
import math

import numpy as np
import torch

from oracle import oracle as O

PB = 64


class OracleBackend:
    device = torch.device("cpu")

    def __init__(self, plan, capacity, prm, grid_def, wtype=1.0):
        self.plan, self.capacity, self.prm, self.grid_def = plan, capacity, prm, grid_def
        self.p = np.zeros(capacity, O.PARTICLE3)
        self.msgs = None
        self.n_owned = self.n_ghost = 0
        W, H = plan.wave_w, plan.wave_h
        full = O.wave_init(W, H, 1, O.WAVE_COUPLED, wtype)
        # StencilImage2DTripleBuffered bookkeeping after Init() (two INIT ping-pongs), SURVEY Appendix B
        self.img = [np.zeros((plan.rows_stored, W), np.float32) for _ in range(3)]
        self.last = [np.zeros(W, np.float32) for _ in range(3)]
        self.read_index, self.write_index, self.unit = [0, 1], 2, [0, 1, 2]
        for _ in range(2):
            out = self.unit.index(2)
            self.img[out][:] = full[plan.store_lo:plan.store_hi]
            self._pingpong()
        self._tex0 = -1

    def _pingpong(self):
        self.write_index, self.read_index[0] = self.read_index[0], self.write_index
        self.read_index[0], self.read_index[1] = self.read_index[1], self.read_index[0]
        u = self.unit
        u[self.write_index], u[self.read_index[0]] = u[self.read_index[0]], u[self.write_index]
        u[self.read_index[0]], u[self.read_index[1]] = u[self.read_index[1]], u[self.read_index[0]]

    # ---- particles (same message layout as cwa_slab_pack / cwa_slab_unpack) --------------------------------
    CAP_MIG, CAP_GHOST = 256, 2048

    def _msg(self):
        return np.zeros(1 + self.CAP_MIG + self.CAP_GHOST, O.PARTICLE3)

    def upload_owned(self, particles):
        self.p[:particles.size] = particles
        self.n_owned, self.n_ghost = particles.size, 0
        self.msgs = {k: self._msg() for k in ("sl", "sr", "rl", "rr")}

    def _dead(self, q):
        return (q["pos"][:, 3] == np.float32(-1.0)) & np.isnan(q["pos"][:, 0])

    def download_owned(self):
        q = self.p[:self.n_owned]
        return q[~self._dead(q)].copy()

    def no_exchange(self):
        self.n_ghost = 0

    def comm_stream(self):
        import contextlib
        return contextlib.nullcontext()

    fused_pack = True          # like CudaBackend: sph_step(pack_next=True) packs the next exchange's messages, pack() then finds its work done
    _packed_for = None

    def pack(self, z_lo, z_hi, band, has_left, has_right):
        if self._packed_for == (self.n_owned, z_lo, z_hi, band, has_left, has_right):      # cwa_slab_pack after cwa_sph_step_slab: no-op
            self._packed_for = None
            f = lambda k: torch.from_numpy(self.msgs[k].view(np.uint8).reshape(-1))
            return (f("sl") if has_left else None, f("sr") if has_right else None)
        self._packed_for = None
        return self._pack_now(z_lo, z_hi, band, has_left, has_right)

    def _pack_now(self, z_lo, z_hi, band, has_left, has_right):
        q = self.p[:self.n_owned]
        live = ~self._dead(q)
        z = q["pos"][:, 2]
        out = []
        with np.errstate(invalid="ignore"):
            zl, zh, bd = np.float32(max(z_lo, -3e38)), np.float32(min(z_hi, 3e38)), np.float32(band)
            to_l = live & (z < zl + bd) if has_left else np.zeros(q.size, bool)
            to_r = live & ~to_l & (z >= zh - bd) if has_right else np.zeros(q.size, bool)
            for key, sel, mig in (("sl", to_l, to_l & (z < zl)), ("sr", to_r, to_r & (z >= zh))):
                m = self.msgs[key]
                m[:] = 0
                migrants, ghosts = q[mig].copy(), q[sel & ~mig].copy()
                assert migrants.size <= self.CAP_MIG and ghosts.size <= self.CAP_GHOST
                hdr = m[:1].view(np.int32)
                hdr[0], hdr[1] = migrants.size, ghosts.size
                m[1:1 + migrants.size] = migrants
                m[1 + self.CAP_MIG:1 + self.CAP_MIG + ghosts.size] = ghosts
                q["pos"][mig] = (np.nan, np.nan, np.nan, -1.0)
                out.append(torch.from_numpy(m.view(np.uint8).reshape(-1)))
        return (out[0] if has_left else None, out[1] if has_right else None)

    def recv_buffers(self, has_left, has_right):
        f = lambda k: torch.from_numpy(self.msgs[k].view(np.uint8).reshape(-1))
        return (f("rl") if has_left else None, f("rr") if has_right else None)

    def unpack(self, has_left, has_right):
        lists = []
        for key, has in (("rl", has_left), ("rr", has_right)):
            if not has:
                lists.append((np.zeros(0, O.PARTICLE3), np.zeros(0, O.PARTICLE3)))
                continue
            m = self.msgs[key]
            hdr = m[:1].view(np.int32)
            lists.append((m[1:1 + hdr[0]].copy(), m[1 + self.CAP_MIG:1 + self.CAP_MIG + hdr[1]].copy()))
        n = self.n_owned
        for part in (lists[0][0], lists[1][0]):
            self.p[n:n + part.size] = part; n += part.size
        self.n_owned = n
        own_sent = []                                # my own outgoing migrants are still neighbours here this frame
        for key, has in (("sl", has_left), ("sr", has_right)):
            m = self.msgs[key]
            own_sent.append(m[1:1 + m[:1].view(np.int32)[0]].copy() if has else np.zeros(0, O.PARTICLE3))
        for part in (lists[0][1], lists[1][1], own_sent[0], own_sent[1]):
            assert n + part.size <= self.capacity
            self.p[n:n + part.size] = part; n += part.size
        self.n_ghost = n - self.n_owned

    # ---- simulation -------------------------------------------------------------------------------
    def _global_texture(self, image):
        if image < 0:
            return None
        pl = self.plan
        tex = np.full((pl.wave_h, pl.wave_w), np.nan, np.float32)     # NaN: any sample outside halos + last row poisons the result
        tex[pl.store_lo:pl.store_hi] = self.img[image]
        if pl.store_hi < pl.wave_h:
            tex[pl.wave_h - 1] = self.last[image]
        return tex

    def sph_step(self, image, pack_next=False):
        n = self.n_owned + self.n_ghost
        tex = self._global_texture(image)
        q = self.p[:n].copy()
        g = O.grid3(*self.grid_def)
        _, cnt, off, idx = O.grid3_build(g, q["pos"])
        grid = (g, cnt, off, idx)
        O.sph3_rho_pres(q, self.prm, tex, grid)
        O.sph3_force(q, self.prm, tex, grid)
        O.sph3_integrate(q, self.prm, tex)
        self.p[:n] = q
        self._packed_for = None
        pl = self.plan
        if pack_next and self.fused_pack and pl.world > 1:        # cwa_sph_step_slab: the integrate pass packs the owned range
            self._pack_now(pl.z_lo, pl.z_hi, pl.ghost_width, pl.has_left, pl.has_right)
            self._packed_for = (self.n_owned, pl.z_lo, pl.z_hi, pl.ghost_width, pl.has_left, pl.has_right)

    def wave_step(self):
        in0, in1, out = self.unit.index(0), self.unit.index(1), self.unit.index(2)
        a = self.prm.attributes
        self.img[out][:] = O.wave_evolve(self.img[in0], self.img[in1], O.WAVE_COUPLED, a[0], a[1], a[2], a[3])
        self._pingpong()

    def bind_texture_unit(self):
        ri0 = self.read_index[0]
        if self.unit[ri0] == 0:
            self._tex0 = ri0

    def newest_image(self):
        return self.unit.index(0)

    def wave_written(self, image):
        pass

    def tex_unit0(self):
        return self._tex0

    def wave_rows(self, image, global_row, nrows):
        lo = global_row - self.plan.store_lo
        return torch.from_numpy(self.img[image][lo:lo + nrows].reshape(-1).view(np.uint8))

    def last_row(self, image):
        return torch.from_numpy(self.last[image].view(np.uint8))

    def copy_own_last_row(self, image):
        self.last[image][:] = self.img[image][self.plan.wave_h - 1 - self.plan.store_lo]

    def full_wave(self):
        """owned rows of the newest level (for gathering in tests)"""
        pl = self.plan
        return self.img[self.newest_image()][pl.row_lo - pl.store_lo:pl.row_hi - pl.store_lo].copy()
