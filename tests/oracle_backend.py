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
        self.scratch = {k: np.zeros(capacity, O.PARTICLE3) for k in ("l", "r", "rl", "rr")}
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

    # ---- particles ------------------------------------------------------------------------------
    def upload_owned(self, particles):
        self.p[:particles.size] = particles
        self.n_owned, self.n_ghost = particles.size, 0

    def download_owned(self):
        return self.p[:self.n_owned].copy()

    def _pred(self, kind, a, b):
        z = self.p["pos"][:self.n_owned, 2]
        with np.errstate(invalid="ignore"):
            if kind == 0:
                return (z >= a) & (z < b)
            if kind == 1:
                return z < a
            if kind == 2:
                return z >= a
            return ~(z < a) & ~(z >= b)

    def select(self, kind, a, b, slot):
        m = self._pred(kind, np.float32(a), np.float32(b))
        sel = self.p[:self.n_owned][m]
        self.scratch[slot][:sel.size] = sel
        return torch.from_numpy(self.scratch[slot][:sel.size].view(np.uint8).reshape(-1))

    def keep(self, z_lo, z_hi):
        m = self._pred(3, np.float32(max(z_lo, -3e38)), np.float32(min(z_hi, 3e38)))
        sel = self.p[:self.n_owned][m].copy()
        self.p[:sel.size] = sel
        self.n_owned = sel.size

    def recv_tensor(self, side, nbytes):
        return torch.from_numpy(self.scratch["r" + side].view(np.uint8).reshape(-1)[:nbytes])

    def _append(self, tensors, base):
        n = base
        for side, t in (("l", tensors[0]), ("r", tensors[1])):
            if t is not None and t.numel():
                m = t.numel() // PB
                assert n + m <= self.capacity
                self.p[n:n + m] = self.scratch["r" + side][:m]
                n += m
        return n

    def set_ghosts(self, recv_l, recv_r):
        self.n_ghost = self._append((recv_l, recv_r), self.n_owned) - self.n_owned

    def append_owned(self, recv_l, recv_r):
        self.n_owned = self._append((recv_l, recv_r), self.n_owned)
        self.n_ghost = 0

    def before_comm(self):
        pass

    def comm_done(self):
        pass

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

    def sph_step(self, image):
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
