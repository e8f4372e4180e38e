"""Multi-GPU coupled step: one process per GPU over torch.distributed.

Decomposition
    particles  -- slabs in z.  Rank r owns z in [z_lo, z_hi); the slab faces are the images of its
                  wave-row block so every particle samples rows held locally.  Each frame:
                  (1) GHOSTS: owned particles within 2h of a face are copied to that neighbour
                      (2h because ghost densities are recomputed locally: a ghost within h of the face
                      needs its own neighbours, which lie within 2h);
                  (2) the three SPH passes run on owned + ghost particles, ghost results are dropped;
                  (3) MIGRATION: owned particles that left [z_lo, z_hi) move to the neighbour.
    wave field -- row blocks.  Rank r stores its owned rows plus SAMPLING halos (ghost reach 2h, the
                  WaveVelocity tap uv + 0.01, one row for the bilinear footprint); after every stencil
                  step the halo rows are overwritten by the neighbours' owned rows (one contiguous
                  send per neighbour) and the global last row -- WaveNormal's uv + (0,1) tap clamps to
                  it from everywhere (force_comp.glsl:136) -- is broadcast by the last rank.
The reference is single-GPU, so the contract is "N ranks reproduce the 1-rank state".

`SlabPlan` is pure host logic; `DistributedCoupled` drives a *backend* (the CUDA library, or -- in
the CPU tests only -- the oracle) and moves bytes with torch.distributed (NCCL on GPUs, gloo on CPU).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

COUPLING_AS_SHIPPED, COUPLING_LATEST = 0, 1
PARTICLE_BYTES = 64


@dataclass
class SlabPlan:
    """Row blocks of the wave field and the matching z slabs of the particles."""
    world: int
    rank: int
    wave_w: int
    wave_h: int
    uv_scale: float
    h: float                      # SPH support radius (smoothing_coeff * particle_radius)
    row_lo: int = 0               # owned rows [row_lo, row_hi)
    row_hi: int = 0
    z_lo: float = -math.inf       # owned particles: z_lo <= z < z_hi
    z_hi: float = math.inf
    halo_lo: int = 0              # sampling halo rows stored below row_lo / above row_hi
    halo_hi: int = 0
    store_lo: int = 0             # stored rows [store_lo, store_hi)
    store_hi: int = 0

    @staticmethod
    def make(world: int, rank: int, wave_w: int, wave_h: int, uv_scale: float, h: float) -> "SlabPlan":
        p = SlabPlan(world, rank, wave_w, wave_h, uv_scale, h)
        base, rem = divmod(wave_h, world)
        p.row_lo = rank * base + min(rank, rem)
        p.row_hi = p.row_lo + base + (1 if rank < rem else 0)
        # texture t = uv_scale * z, texel row = t * H - 0.5: row boundary b  <->  z = b / (H * uv_scale)
        scale = wave_h * uv_scale
        p.z_lo = -math.inf if rank == 0 else p.row_lo / scale
        p.z_hi = math.inf if rank == world - 1 else p.row_hi / scale
        ghost_rows = 2.0 * h * scale                      # ghost layer (2h) in texel rows
        p.halo_lo = 0 if rank == 0 else int(math.ceil(ghost_rows)) + 2
        p.halo_hi = 0 if rank == world - 1 else int(math.ceil(ghost_rows + 0.01 * wave_h)) + 3
        p.store_lo = max(0, p.row_lo - p.halo_lo)
        p.store_hi = min(wave_h, p.row_hi + p.halo_hi)
        return p

    @property
    def ghost_width(self) -> float:
        return 2.0 * self.h

    @property
    def rows_stored(self) -> int:
        return self.store_hi - self.store_lo

    def validate(self):
        assert self.row_hi > self.row_lo, "more ranks than wave rows"
        if self.world > 1:
            width = (self.row_hi - self.row_lo) / (self.wave_h * self.uv_scale)
            assert width > 4.0 * self.h, f"slab of width {width} is thinner than two ghost layers (4h = {4 * self.h})"
            assert self.row_hi - self.row_lo >= max(self.halo_lo, self.halo_hi), "row block smaller than its neighbours' halos"


class DistributedCoupled:
    """nframes x (ghost exchange, rho -> force -> integrate, wave stencil + halo refresh, display bind, migration)."""

    def __init__(self, backend, plan: SlabPlan, dist=None):
        self.b = backend
        self.plan = plan
        self.dist = dist                       # torch.distributed module (None: single rank)
        self.world, self.rank = plan.world, plan.rank
        plan.validate()

    # ---- point-to-point helpers --------------------------------------------------------------
    def _exchange(self, send_left, send_right):
        """Send byte tensors to rank-1 / rank+1, receive theirs.  Returns (from_left, from_right)."""
        import torch
        d = self.dist
        left = self.rank - 1 if self.rank > 0 else None
        right = self.rank + 1 if self.rank < self.world - 1 else None
        dev = self.b.device
        cnt_out = {n: torch.tensor([t.numel() if t is not None else 0], dtype=torch.int64, device=dev) for n, t in (("l", send_left), ("r", send_right))}
        cnt_in = {n: torch.zeros(1, dtype=torch.int64, device=dev) for n in ("l", "r")}
        ops = []
        if left is not None:
            ops += [d.P2POp(d.isend, cnt_out["l"], left), d.P2POp(d.irecv, cnt_in["l"], left)]
        if right is not None:
            ops += [d.P2POp(d.isend, cnt_out["r"], right), d.P2POp(d.irecv, cnt_in["r"], right)]
        for w in d.batch_isend_irecv(ops):
            w.wait()
        n_l, n_r = int(cnt_in["l"].item()), int(cnt_in["r"].item())
        recv_l = self.b.recv_tensor("l", n_l) if left is not None else None
        recv_r = self.b.recv_tensor("r", n_r) if right is not None else None
        ops = []
        if left is not None:
            if send_left is not None and send_left.numel():
                ops.append(d.P2POp(d.isend, send_left, left))
            if n_l:
                ops.append(d.P2POp(d.irecv, recv_l, left))
        if right is not None:
            if send_right is not None and send_right.numel():
                ops.append(d.P2POp(d.isend, send_right, right))
            if n_r:
                ops.append(d.P2POp(d.irecv, recv_r, right))
        if ops:
            for w in d.batch_isend_irecv(ops):
                w.wait()
        self.b.comm_done()
        return recv_l, recv_r

    # ---- one frame -----------------------------------------------------------------------------
    def _ghost_exchange(self):
        p, b = self.plan, self.b
        if self.world == 1:
            b.set_ghosts(None, None)
            return
        send_l = b.select(0, p.z_lo, p.z_lo + p.ghost_width, "l") if self.rank > 0 else None
        send_r = b.select(0, p.z_hi - p.ghost_width, p.z_hi, "r") if self.rank < self.world - 1 else None
        b.before_comm()
        recv_l, recv_r = self._exchange(send_l, send_r)
        b.set_ghosts(recv_l, recv_r)

    def _migrate(self):
        p, b = self.plan, self.b
        if self.world == 1:
            return
        send_l = b.select(1, p.z_lo, 0.0, "l") if self.rank > 0 else None
        send_r = b.select(2, p.z_hi, 0.0, "r") if self.rank < self.world - 1 else None
        b.keep(p.z_lo if self.rank > 0 else -math.inf, p.z_hi if self.rank < self.world - 1 else math.inf)
        b.before_comm()
        recv_l, recv_r = self._exchange(send_l, send_r)
        b.append_owned(recv_l, recv_r)

    def _wave_halo_refresh(self):
        """After a stencil step: overwrite the halo rows of the newest level and refresh the global last row."""
        p, b = self.plan, self.b
        if self.world == 1:
            return
        d = self.dist
        img = b.newest_image()
        b.before_comm()
        ops = []
        left = self.rank - 1 if self.rank > 0 else None
        right = self.rank + 1 if self.rank < self.world - 1 else None
        # my top owned rows fill the left neighbour's upper halo (its halo_hi rows above its row_hi == my row_lo)
        if left is not None:
            left_plan = SlabPlan.make(p.world, left, p.wave_w, p.wave_h, p.uv_scale, p.h)
            n_up = left_plan.store_hi - left_plan.row_hi
            ops.append(d.P2POp(d.isend, b.wave_rows(img, p.row_lo, n_up), left))
            ops.append(d.P2POp(d.irecv, b.wave_rows(img, p.store_lo, p.row_lo - p.store_lo), left))
        if right is not None:
            right_plan = SlabPlan.make(p.world, right, p.wave_w, p.wave_h, p.uv_scale, p.h)
            n_dn = right_plan.row_lo - right_plan.store_lo
            ops.append(d.P2POp(d.isend, b.wave_rows(img, p.row_hi - n_dn, n_dn), right))
            ops.append(d.P2POp(d.irecv, b.wave_rows(img, p.row_hi, p.store_hi - p.row_hi), right))
        for w in d.batch_isend_irecv(ops):
            w.wait()
        last = b.last_row(img)
        if self.rank == self.world - 1:
            b.copy_own_last_row(img)
        d.broadcast(last, src=self.world - 1)
        b.comm_done()

    def step(self, nframes: int = 1, coupling: int = COUPLING_AS_SHIPPED):
        b = self.b
        for _ in range(nframes):
            self._ghost_exchange()
            image = b.newest_image() if coupling == COUPLING_LATEST else b.tex_unit0()
            b.sph_step(image)                      # idle(): rho_pres, force, integrate  (Main.cpp:549-557)
            b.wave_step()                          # Module::sComputeAll               (Main.cpp:560)
            self._wave_halo_refresh()
            b.bind_texture_unit()                  # display(): GetReadImage(0).BindTextureUnit()  (Main.cpp:413)
            self._migrate()

    def init_wave_halos(self):
        """Init() wrote both read levels from global coordinates, so halos and last rows only need the broadcast."""
        if self.world == 1:
            return
        for img in range(3):
            last = self.b.last_row(img)
            if self.rank == self.world - 1:
                self.b.copy_own_last_row(img)
            self.b.before_comm()
            self.dist.broadcast(last, src=self.world - 1)
            self.b.comm_done()


# ====================================================================================================
# backend: the CUDA library
# ====================================================================================================
class _DevPtr:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class CudaBackend:
    """Owns the per-rank library objects: particle SSBO (owned + ghosts), grid, row-block wave object."""

    def __init__(self, cwa, ctx, plan: SlabPlan, capacity: int, grid_min, grid_max, grid_cells, wave_ch=1):
        import torch
        self.torch = torch
        self.cwa, self.ctx, self.plan = cwa, ctx, plan
        self.device = torch.device("cuda", ctx.device)
        self.capacity = capacity
        self.buffer = cwa.Buffer(ctx, nbytes=capacity * PARTICLE_BYTES)
        self.grid = cwa.UniformGrid(ctx, 3, grid_min, grid_max, grid_cells, capacity, compact_index=True)
        self.sph = cwa.Sph(ctx, capacity, self.grid, buffer=self.buffer)
        self.scratch = {k: cwa.Buffer(ctx, nbytes=capacity * PARTICLE_BYTES) for k in ("l", "r", "keep", "rl", "rr")}
        self.wave = cwa.StencilImage2DTripleBuffered.create_block(ctx, plan.wave_w, plan.wave_h, plan.store_lo, plan.rows_stored, wave_ch)
        self.wave_ch = wave_ch
        self.n_owned = 0
        self.n_ghost = 0
        self._ptr = self.buffer.device_ptr()

    def _tensor(self, buf, offset_bytes: int, nbytes: int):
        if nbytes == 0:
            return self.torch.empty(0, dtype=self.torch.uint8, device=self.device)
        return self.torch.as_tensor(_DevPtr(buf.device_ptr() + offset_bytes, nbytes), device=self.device)

    # ---- particles ------------------------------------------------------------------------------
    def upload_owned(self, particles: np.ndarray):
        assert particles.size <= self.capacity
        if particles.size:
            self.buffer.sub_data(particles)
        self.n_owned, self.n_ghost = particles.size, 0

    def download_owned(self) -> np.ndarray:
        return self.buffer.read(self.cwa.PARTICLE, self.n_owned)

    def select(self, kind: int, a: float, b: float, slot: str):
        """copy_if over the OWNED particles into the send scratch `slot`; returns a byte tensor view."""
        import ctypes as C
        cnt = C.c_int()
        a = max(min(a, 3.0e38), -3.0e38); b = max(min(b, 3.0e38), -3.0e38)
        self.cwa.check(self.ctx.lib.cwa_particles_copy_if(self.ctx.h, self.buffer.h, self.n_owned, 2, kind, a, b, self.scratch[slot].h, 0, C.byref(cnt)))
        return self._tensor(self.scratch[slot], 0, cnt.value * PARTICLE_BYTES)

    def keep(self, z_lo: float, z_hi: float):
        import ctypes as C
        cnt = C.c_int()
        a = max(z_lo, -3.0e38); b = min(z_hi, 3.0e38)
        self.cwa.check(self.ctx.lib.cwa_particles_copy_if(self.ctx.h, self.buffer.h, self.n_owned, 2, 3, a, b, self.scratch["keep"].h, 0, C.byref(cnt)))
        if cnt.value:
            self.cwa.check(self.ctx.lib.cwa_buffer_copy(self.ctx.h, self.scratch["keep"].h, self.buffer.h, 0, 0, cnt.value * PARTICLE_BYTES))
        self.n_owned = cnt.value

    def recv_tensor(self, side: str, nbytes: int):
        return self._tensor(self.scratch["r" + side], 0, nbytes)

    def _append(self, tensors, base: int) -> int:
        n = base
        for side, t in (("l", tensors[0]), ("r", tensors[1])):
            if t is not None and t.numel():
                m = t.numel() // PARTICLE_BYTES
                assert n + m <= self.capacity, f"rank {self.plan.rank}: particle capacity {self.capacity} exceeded"
                self.cwa.check(self.ctx.lib.cwa_buffer_copy(self.ctx.h, self.scratch["r" + side].h, self.buffer.h, 0, n * PARTICLE_BYTES, t.numel()))
                n += m
        return n

    def set_ghosts(self, recv_l, recv_r):
        self.n_ghost = self._append((recv_l, recv_r), self.n_owned) - self.n_owned

    def append_owned(self, recv_l, recv_r):
        self.n_owned = self._append((recv_l, recv_r), self.n_owned)
        self.n_ghost = 0

    # ---- stream ordering between the library's stream and torch's communication streams ---------
    def before_comm(self):
        self.ctx.synchronize()

    def comm_done(self):
        self.torch.cuda.synchronize(self.device)

    # ---- simulation -------------------------------------------------------------------------------
    def sph_step(self, image: int):
        n = self.n_owned + self.n_ghost
        self.cwa.check(self.ctx.lib.cwa_sph_set_count(self.ctx.h, self.sph.h, n))
        self.sph.bind_wave(self.wave if image >= 0 else None, image)
        self.sph.step(1)

    def wave_step(self):
        self.wave.Compute(1)

    def bind_texture_unit(self):
        self.wave.bind_texture_unit()

    def newest_image(self) -> int:
        return self.wave.role_image(0)

    def tex_unit0(self) -> int:
        return self.wave.state()["tex_unit0"]

    def wave_rows(self, image: int, global_row: int, nrows: int):
        row_bytes = self.plan.wave_w * self.wave_ch * 4
        off = (global_row - self.plan.store_lo) * row_bytes
        return self._tensor(self.wave.image_buffer(image), off, nrows * row_bytes)

    def last_row(self, image: int):
        return self._tensor(self.wave.last_row_buffer(image), 0, self.plan.wave_w * self.wave_ch * 4)

    def copy_own_last_row(self, image: int):
        row_bytes = self.plan.wave_w * self.wave_ch * 4
        src, dst = self.wave.image_buffer(image), self.wave.last_row_buffer(image)
        off = (self.plan.wave_h - 1 - self.plan.store_lo) * row_bytes
        self.cwa.check(self.ctx.lib.cwa_buffer_copy(self.ctx.h, src.h, dst.h, off, 0, row_bytes))
