"""Multi-GPU coupled step: one process per GPU over torch.distributed.

Decomposition
    particles  -- slabs in z.  Rank r owns z in [z_lo, z_hi); the slab faces are the images of its
                  wave-row block so every particle samples rows held locally.  At the start of every
                  frame each rank sends ONE fixed-size message to each neighbour:
                    * MIGRANTS: owned particles that left the slab through that face during the last
                      integrate (they are marked dead in place -- NaN position, skipped by the grid);
                    * GHOSTS: owned particles within 2h of the face.  2h because ghost densities are
                      recomputed locally: a ghost within h of the face needs its own neighbours,
                      which lie within 2h.
                  The counts travel in the message header, so there is no size negotiation; the
                  receiver appends migrants to its owned range and ghosts behind it, runs the three
                  SPH passes on owned + ghosts and drops the ghost results.
    wave field -- row blocks.  Rank r stores its owned rows plus SAMPLING halos (ghost reach 2h, the
                  WaveVelocity tap uv + 0.01, one row for the bilinear footprint); after every stencil
                  step the halo rows are overwritten by the neighbours' owned rows (one contiguous
                  send per neighbour) and the global last row -- WaveNormal's uv + (0,1) tap clamps to
                  it from everywhere (force_comp.glsl:136) -- is broadcast by the last rank.
All communication is enqueued stream-ordered behind the library's kernels (NCCL on the context's
stream); a frame has one host synchronisation (reading the two particle counts).
The reference is single-GPU, so the contract is "N ranks reproduce the 1-rank state".

`SlabPlan` is pure host logic; `DistributedCoupled` drives a *backend* (the CUDA library, or -- in
the CPU tests only -- the oracle) and moves bytes with torch.distributed (NCCL on GPUs, gloo on CPU).
"""
from __future__ import annotations

import contextlib
import math
from dataclasses import dataclass

import numpy as np

COUPLING_AS_SHIPPED, COUPLING_LATEST = 0, 1
PARTICLE_BYTES = 64
DEAD_W = -1.0


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
    row_bounds: tuple = None      # the row blocks of all ranks when they are not the even split

    @staticmethod
    def make(world: int, rank: int, wave_w: int, wave_h: int, uv_scale: float, h: float, row_bounds=None) -> "SlabPlan":
        """row_bounds: optional world+1 ascending row indices (0 ... wave_h) = the row blocks of the ranks, e.g. chosen so
        that the z slabs hold equal particle counts (balanced_row_bounds); default: equal row counts."""
        p = SlabPlan(world, rank, wave_w, wave_h, uv_scale, h)
        if row_bounds is not None:
            rb = [int(v) for v in row_bounds]
            assert len(rb) == world + 1 and rb[0] == 0 and rb[-1] == wave_h and all(b > a for a, b in zip(rb, rb[1:])), \
                f"row_bounds must be {world + 1} ascending rows from 0 to {wave_h}: {rb}"
            p.row_lo, p.row_hi = rb[rank], rb[rank + 1]
            p.row_bounds = tuple(rb)
        else:
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

    @staticmethod
    def balanced_row_bounds(world: int, wave_h: int, uv_scale: float, z_sorted_sample: np.ndarray):
        """Row blocks whose z slabs hold about equal shares of the particles (z_sorted_sample: ascending z of the particles or
        of a representative sample).  The wave rows follow the particles: a few per cent more stencil rows on some ranks
        cost microseconds, a few per cent more particles cost tens."""
        scale = wave_h * uv_scale
        rb = [0]
        n = len(z_sorted_sample)
        for r in range(1, world):
            zc = float(z_sorted_sample[min(n - 1, (n * r) // world)])
            row = int(round(zc * scale))
            rb.append(min(max(row, rb[-1] + 1), wave_h - (world - r)))
        rb.append(wave_h)
        return rb

    @property
    def ghost_width(self) -> float:
        return 2.0 * self.h

    @property
    def rows_stored(self) -> int:
        return self.store_hi - self.store_lo

    @property
    def has_left(self) -> bool:
        return self.rank > 0

    @property
    def has_right(self) -> bool:
        return self.rank < self.world - 1

    def validate(self):
        assert self.row_hi > self.row_lo, "more ranks than wave rows"
        if self.world > 1:
            width = (self.row_hi - self.row_lo) / (self.wave_h * self.uv_scale)
            assert width > 4.0 * self.h, f"slab of width {width} is thinner than two ghost layers (4h = {4 * self.h})"
            assert self.row_hi - self.row_lo >= max(self.halo_lo, self.halo_hi), "row block smaller than its neighbours' halos"


class DistributedCoupled:
    """nframes x (migrant+ghost exchange, rho -> force -> integrate, wave stencil + halo refresh, display bind)."""

    def __init__(self, backend, plan: SlabPlan, dist=None):
        self.b = backend
        self.plan = plan
        self.dist = dist                       # torch.distributed module (None: single rank)
        self.world, self.rank = plan.world, plan.rank
        plan.validate()
        self._halo_pending = False
        self._nbr_plans = {r: SlabPlan.make(plan.world, r, plan.wave_w, plan.wave_h, plan.uv_scale, plan.h, plan.row_bounds)
                           for r in (plan.rank - 1, plan.rank + 1) if 0 <= r < plan.world}

    def _p2p(self, pairs):
        """pairs: list of (send_tensor | None, recv_tensor | None, peer).  Stream-ordered on CUDA."""
        d = self.dist
        ops = []
        for send, recv, peer in pairs:
            if send is not None and send.numel():
                ops.append(d.P2POp(d.isend, send, peer))
            if recv is not None and recv.numel():
                ops.append(d.P2POp(d.irecv, recv, peer))
        if ops:
            for w in d.batch_isend_irecv(ops):
                w.wait()

    # ---- one frame -----------------------------------------------------------------------------
    def _wave_pairs(self):
        """P2P operations that refresh the halo rows of the newest level and distribute the global last row."""
        p, b = self.plan, self.b
        img = b.newest_image()
        pairs = []
        if p.has_left:
            lp = self._nbr_plans[self.rank - 1]
            n_up = lp.store_hi - lp.row_hi           # my first owned rows are the left neighbour's upper halo
            pairs.append((b.wave_rows(img, p.row_lo, n_up), b.wave_rows(img, p.store_lo, p.row_lo - p.store_lo), self.rank - 1))
        if p.has_right:
            rp = self._nbr_plans[self.rank + 1]
            n_dn = rp.row_lo - rp.store_lo           # my last owned rows are the right neighbour's lower halo
            pairs.append((b.wave_rows(img, p.row_hi - n_dn, n_dn), b.wave_rows(img, p.row_hi, p.store_hi - p.row_hi), self.rank + 1))
        # global last row: WaveNormal's uv + (0,1) tap clamps to it from everywhere (force_comp.glsl:136); the last rank
        # sends it to every other rank inside the same group (a handful of small messages, no separate collective)
        last = self.world - 1
        if self.rank == last:
            b.copy_own_last_row(img)
            pairs += [(b.last_row(img), None, r) for r in range(last)]
        else:
            pairs.append((None, b.last_row(img), last))
        return pairs

    def _exchange(self, particles: bool, wave: bool):
        """ONE grouped NCCL call per frame: migrants + ghosts of this frame and -- deferred from the end of the previous
        frame -- the wave halo rows / last row the SPH passes are about to sample."""
        p, b = self.plan, self.b
        if self.world == 1:
            if particles:
                b.no_exchange()
            return
        with b.comm_stream():
            pairs = []
            if particles:
                send_l, send_r = b.pack(p.z_lo, p.z_hi, p.ghost_width, p.has_left, p.has_right)
                recv_l, recv_r = b.recv_buffers(p.has_left, p.has_right)
                if p.has_left:
                    pairs.append((send_l, recv_l, self.rank - 1))
                if p.has_right:
                    pairs.append((send_r, recv_r, self.rank + 1))
            if wave:
                pairs += self._wave_pairs()
            self._p2p(pairs)
            if wave:
                b.wave_written(b.newest_image())       # halo rows of the newest level were received into the image
            if particles:
                b.unpack(p.has_left, p.has_right)

    def _particle_exchange(self):
        self._exchange(True, False)

    def _wave_halo_refresh(self):
        self._exchange(False, True)

    def step(self, nframes: int = 1, coupling: int = COUPLING_AS_SHIPPED):
        b = self.b
        for f in range(nframes):
            # the halo refresh of the previous frame's stencil step travels with this frame's particles
            self._exchange(True, self._halo_pending)
            self._halo_pending = False
            image = b.newest_image() if coupling == COUPLING_LATEST else b.tex_unit0()
            # idle(): rho_pres, force, integrate (Main.cpp:549-557).  When another frame of this call follows, the integrate
            # pass may already pack that frame's migrant / ghost messages (the last frame leaves every particle in its
            # owner's buffer, so reads and checks between calls see a complete state)
            b.sph_step(image, pack_next=(f + 1 < nframes))
            b.wave_step()                          # Module::sComputeAll               (Main.cpp:560)
            b.bind_texture_unit()                  # display(): GetReadImage(0).BindTextureUnit()  (Main.cpp:413)
            self._halo_pending = True
        if self._halo_pending:                     # leave a consistent field behind (reads, checks, the next call)
            self._exchange(False, True)
            self._halo_pending = False

    def init_wave_halos(self):
        """Init() wrote both read levels from global coordinates, so only the last rows need the broadcast."""
        if self.world == 1:
            return
        with self.b.comm_stream():
            for img in range(3):
                if self.rank == self.world - 1:
                    self.b.copy_own_last_row(img)
                self.dist.broadcast(self.b.last_row(img), src=self.world - 1)


# ====================================================================================================
# backend: the CUDA library
# ====================================================================================================
class _DevPtr:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class CudaBackend:
    """Owns the per-rank library objects: particle SSBO (owned + ghosts), grid, row-block wave object."""

    def __init__(self, cwa, ctx, plan: SlabPlan, capacity: int, grid_min, grid_max, grid_cells, wave_ch=1,
                 cap_mig: int = 16384, cap_ghost: int = 65536):
        import torch
        self.torch = torch
        self.cwa, self.ctx, self.plan = cwa, ctx, plan
        self.device = torch.device("cuda", ctx.device)
        self.capacity = capacity
        self._views = {}
        self.cap_mig, self.cap_ghost = cap_mig, cap_ghost
        self.buffer = cwa.Buffer(ctx, nbytes=capacity * PARTICLE_BYTES)
        self.grid = cwa.UniformGrid(ctx, 3, grid_min, grid_max, grid_cells, capacity, compact_index=True)
        self.sph = cwa.Sph(ctx, capacity, self.grid, buffer=self.buffer)
        msg_bytes = (1 + cap_mig + cap_ghost) * PARTICLE_BYTES
        self.msg = {k: cwa.Buffer(ctx, nbytes=msg_bytes) for k in ("sl", "sr", "rl", "rr")}
        self.msg_t = {k: self._tensor(v, 0, msg_bytes) for k, v in self.msg.items()}
        self.scratch = None                       # allocated on the first compaction
        self.wave = cwa.StencilImage2DTripleBuffered.create_block(ctx, plan.wave_w, plan.wave_h, plan.store_lo, plan.rows_stored, wave_ch)
        self.wave_ch = wave_ch
        self.n_owned = 0                          # owned RANGE (may contain dead slots)
        self.n_ghost = 0
        self.migrated_in = 0
        self._stream = torch.cuda.ExternalStream(ctx.stream, device=self.device)
        import os
        self.fused_pack = os.environ.get("CWA_FUSED_PACK", "1") != "0"   # integrate packs the next exchange's messages (cwa_sph_step_slab)

    def _tensor(self, buf, offset_bytes: int, nbytes: int):
        if nbytes == 0:
            return self.torch.empty(0, dtype=self.torch.uint8, device=self.device)
        key = (buf.h, offset_bytes, nbytes)
        t = self._views.get(key)
        if t is None:                             # views of library-owned memory; built once, reused every frame
            t = self._views[key] = self.torch.as_tensor(_DevPtr(buf.device_ptr() + offset_bytes, nbytes), device=self.device)
        return t

    def comm_stream(self):
        """torch collectives issued inside are ordered behind / ahead of the library's kernels on ITS stream."""
        return self.torch.cuda.stream(self._stream)

    # ---- particles ------------------------------------------------------------------------------
    def upload_owned(self, particles: np.ndarray):
        assert particles.size <= self.capacity
        if particles.size:
            self.buffer.sub_data(particles)
        self.n_owned, self.n_ghost = particles.size, 0

    def download_owned(self) -> np.ndarray:
        p = self.buffer.read(self.cwa.PARTICLE, self.n_owned)
        dead = (p["pos"][:, 3] == np.float32(DEAD_W)) & np.isnan(p["pos"][:, 0])
        return p[~dead]

    def no_exchange(self):
        self.n_ghost = 0

    def _maybe_compact(self):
        reserve = 2 * (2 * self.cap_mig + self.cap_ghost)
        if self.n_owned + reserve <= self.capacity:
            return
        import ctypes as C
        if self.scratch is None:
            self.scratch = self.cwa.Buffer(self.ctx, nbytes=self.capacity * PARTICLE_BYTES)
        n = C.c_int()
        self.cwa.check(self.ctx.lib.cwa_slab_compact(self.ctx.h, self.buffer.h, self.n_owned, self.scratch.h, C.byref(n)))
        self.n_owned = n.value
        assert self.n_owned + reserve <= self.capacity, f"rank {self.plan.rank}: particle capacity {self.capacity} exhausted"

    def pack(self, z_lo, z_hi, band, has_left, has_right):
        self._maybe_compact()
        f = lambda v: max(min(v, 3.0e38), -3.0e38)
        self.cwa.check(self.ctx.lib.cwa_slab_pack(self.ctx.h, self.buffer.h, self.n_owned, f(z_lo), f(z_hi), band,
                                                  self.msg["sl"].h if has_left else -1, self.msg["sr"].h if has_right else -1,
                                                  self.cap_mig, self.cap_ghost))
        return (self.msg_t["sl"] if has_left else None, self.msg_t["sr"] if has_right else None)

    def recv_buffers(self, has_left, has_right):
        return (self.msg_t["rl"] if has_left else None, self.msg_t["rr"] if has_right else None)

    def unpack(self, has_left, has_right):
        import ctypes as C
        counts = (C.c_int * 4)()
        self.cwa.check(self.ctx.lib.cwa_slab_unpack(self.ctx.h, self.buffer.h, self.n_owned, self.msg["rl"].h if has_left else -1,
                                                    self.msg["rr"].h if has_right else -1, self.msg["sl"].h if has_left else -1,
                                                    self.msg["sr"].h if has_right else -1, self.cap_mig, self.cap_ghost, counts))
        if counts[2]:
            hdrs = {k: self.msg[k].read(np.int32, 4).tolist() for k in ("sl", "sr", "rl", "rr")}
            raise RuntimeError(f"rank {self.plan.rank}: exchange overflow (flags {counts[2]}, counts {list(counts)}, n_owned {self.n_owned}, "
                               f"capacity {self.capacity}, caps {self.cap_mig}/{self.cap_ghost}, headers {hdrs}): raise cap_mig/cap_ghost/capacity")
        self.n_ghost = counts[1] - counts[0]
        self.n_owned = counts[0]
        self.migrated_in += counts[3]

    # ---- simulation -------------------------------------------------------------------------------
    def sph_step(self, image: int, pack_next: bool = False):
        n = self.n_owned + self.n_ghost
        self.cwa.check(self.ctx.lib.cwa_sph_set_count(self.ctx.h, self.sph.h, n))
        self.sph.bind_wave(self.wave if image >= 0 else None, image)
        p = self.plan
        if p.world > 1 and self.fused_pack and pack_next:
            # the integrate pass packs the migrant / ghost messages of the NEXT exchange (pack() then finds its work done)
            f = lambda v: max(min(v, 3.0e38), -3.0e38)
            self.cwa.check(self.ctx.lib.cwa_sph_step_slab(self.ctx.h, self.sph.h, self.n_owned, f(p.z_lo), f(p.z_hi), p.ghost_width,
                                                          self.msg["sl"].h if p.has_left else -1, self.msg["sr"].h if p.has_right else -1,
                                                          self.cap_mig, self.cap_ghost))
        else:
            self.sph.step(1)

    def wave_step(self):
        self.wave.Compute(1)

    def bind_texture_unit(self):
        self.wave.bind_texture_unit()

    def newest_image(self) -> int:
        return self.wave.role_image(0)

    def wave_written(self, image: int):
        """Rows of `image` were written through its raw device pointer (NCCL receive): derived copies are stale."""
        self.cwa.check(self.ctx.lib.cwa_wave_mark_written(self.ctx.h, self.wave.h, image))

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
