"""Multi-GPU coupled step.

The product path is `SlabRank`: a thin ctypes front of the C-ABI slab object (csrc/slab.cu) -- peer stores into the neighbours'
mailboxes over NVLink, device-side arrival flags, device-resident particle counts; no host synchronisation and no collective
library on the per-frame path.  torch.distributed is only used ONCE, at set-up, to hand the 64-byte IPC handles of the mailboxes
to the other processes (`connect_processes`).

`SlabPlan` / `DistributedCoupled` below are the host-side statement of the same decomposition, kept as the protocol the CPU tests
(world_size 2 / 3 over gloo, oracle as the compute backend) check: "N ranks reproduce the 1-rank state".

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

    @staticmethod
    def cost_balanced_row_bounds(row_bounds, costs, min_rows: int = 8, damping: float = 1.0):
        """Row blocks re-cut so that every rank carries the same MEASURED cost.  `costs[r]` is the busy time of rank r under `row_bounds`
        (its kernels without the waits on its neighbours); the cost is taken as uniform inside a rank's block, the cumulative cost over
        the rows is inverted at k / world.  A rank that owns a wall -- where the sheet is pressed flat and the clump kernels run -- ends
        up with fewer rows.  damping < 1 moves only part of the way (the cost of a wall region does not shrink with the block)."""
        rb = [int(v) for v in row_bounds]
        world = len(rb) - 1
        assert len(costs) == world and all(c > 0 for c in costs)
        cum = [0.0]
        for c in costs:
            cum.append(cum[-1] + float(c))
        total = cum[-1]
        out = [rb[0]]
        for k in range(1, world):
            target = total * k / world
            r = max(i for i in range(world) if cum[i] <= target)
            frac = (target - cum[r]) / (cum[r + 1] - cum[r])
            row = rb[r] + frac * (rb[r + 1] - rb[r])
            row = rb[k] + damping * (row - rb[k])
            out.append(int(round(row)))
        out.append(rb[-1])
        for k in range(1, world):                       # keep every block at least min_rows tall, bounds ascending
            out[k] = max(out[k], out[k - 1] + min_rows)
        for k in range(world - 1, 0, -1):
            out[k] = min(out[k], out[k + 1] - min_rows)
        return out

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
# the CUDA library: one rank of the slab decomposition through the C ABI (csrc/slab.cu)
# ====================================================================================================
def plan_desc(lib, world: int, rank: int, wave_w: int, wave_h: int, wave_ch: int, uv_scale_z: float, h: float, row_bounds=None,
              cap_mig: int = 16384, cap_ghost: int = 65536, capacity: int = 0, timeout_ms: int = 10000):
    """cwa_slab_plan: the row block / z slab of `rank` as the library plans it (same rule as SlabPlan.make)."""
    import ctypes as C
    from . import _capi
    d = _capi.SlabDesc()
    rb = (C.c_int * (world + 1))(*[int(v) for v in row_bounds]) if row_bounds is not None else None
    _capi.check(lib.cwa_slab_plan(world, rank, wave_w, wave_h, wave_ch, float(uv_scale_z), float(h), rb, C.byref(d)))
    d.cap_mig, d.cap_ghost, d.capacity, d.timeout_ms = int(cap_mig), int(cap_ghost), int(capacity), int(timeout_ms)
    return d


class SlabRank:
    """Per-rank library objects of the slab-decomposed coupled frame: particle SSBO (owned + ghosts), slab-local uniform grid,
    row-block wave object and the cwa_slab object that steps them."""

    def __init__(self, cwa, ctx, desc, grid_min, grid_max, grid_cells):
        import ctypes as C
        self.cwa, self.ctx, self.desc = cwa, ctx, desc
        self.capacity = desc.capacity
        self.buffer = cwa.Buffer(ctx, nbytes=desc.capacity * PARTICLE_BYTES)
        self.grid = cwa.UniformGrid(ctx, 3, grid_min, grid_max, grid_cells, desc.capacity, compact_index=True)
        self.sph = cwa.Sph(ctx, desc.capacity, self.grid, buffer=self.buffer)
        self.wave = cwa.StencilImage2DTripleBuffered.create_block(ctx, desc.wave_w, desc.wave_h, desc.store_lo, desc.store_hi - desc.store_lo, desc.wave_ch)
        h = C.c_int(-1)
        cwa.check(ctx.lib.cwa_slab_create(ctx.h, C.byref(desc), self.sph.h, self.wave.h, C.byref(h)))
        self.h = h.value

    # ---- wiring -----------------------------------------------------------------------------------
    def export(self) -> bytes:
        import ctypes as C
        buf = (C.c_ubyte * 64)()
        self.cwa.check(self.ctx.lib.cwa_slab_export(self.ctx.h, self.h, buf))
        return bytes(buf)

    def mailbox(self):
        import ctypes as C
        p, n = C.c_void_p(), C.c_size_t()
        self.cwa.check(self.ctx.lib.cwa_slab_mailbox(self.ctx.h, self.h, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def connect(self, peer_rank: int, handle: bytes | None = None, ptr: int | None = None):
        import ctypes as C
        hb = (C.c_ubyte * 64)(*handle) if handle is not None else None
        self.cwa.check(self.ctx.lib.cwa_slab_connect(self.ctx.h, self.h, peer_rank, hb, C.c_void_p(ptr) if ptr else None))

    # ---- state ------------------------------------------------------------------------------------
    def upload_owned(self, particles: np.ndarray):
        assert particles.size <= self.capacity
        if particles.size:
            self.buffer.sub_data(particles)
        self.cwa.check(self.ctx.lib.cwa_slab_set_owned(self.ctx.h, self.h, int(particles.size)))

    def counts(self) -> dict:
        import ctypes as C
        c = (C.c_int * 8)()
        self.cwa.check(self.ctx.lib.cwa_slab_counts(self.ctx.h, self.h, c))
        return {"n_owned": c[0], "n_total": c[1], "n_ghost": c[1] - c[0], "free": c[2], "err": c[3],
                "migrated_in": (c[4] & 0xffffffff) | (c[5] << 32), "particle_msgs": c[6], "wave_msgs": c[7]}

    def download_owned(self) -> np.ndarray:
        n = self.counts()["n_owned"]
        p = self.buffer.read(self.cwa.PARTICLE, n)
        dead = (p["pos"][:, 3] == np.float32(DEAD_W)) & np.isnan(p["pos"][:, 0])
        return p[~dead]

    def owned_wave_rows(self, role: int = 0) -> np.ndarray:
        d = self.desc
        return self.wave.read_role(role)[d.row_lo - d.store_lo:d.row_hi - d.store_lo]

    # ---- simulation -------------------------------------------------------------------------------
    def step(self, nframes: int = 1, coupling: int = COUPLING_AS_SHIPPED):
        self.cwa.check(self.ctx.lib.cwa_slab_step(self.ctx.h, self.h, nframes, coupling))

    def check(self):
        """Raise when the device-side protocol reported an error (message overflow, capacity, a peer that never answered)."""
        c = self.counts()
        if c["err"]:
            names = [n for b, n in ((1, "a sender overflowed its message (raise cap_mig / cap_ghost)"), (2, "particle capacity exceeded"),
                                    (4, "timeout waiting for a neighbour's particles"), (8, "timeout waiting for a neighbour's wave rows")) if c["err"] & b]
            raise RuntimeError(f"rank {self.desc.rank}: slab exchange error bits {c['err']}: " + "; ".join(names) + f" ({c})")
        return c


def connect_local(ranks):
    """Ranks that are contexts of THIS process (any mix of devices): hand every rank the others' mailbox pointers."""
    for a in ranks:
        for b in ranks:
            if a is not b:
                a.connect(b.desc.rank, ptr=b.mailbox()[0])


def group_step(ranks, nframes: int = 1, coupling: int = COUPLING_AS_SHIPPED):
    """cwa_slab_group_step: all ranks of one process stepped by one host thread."""
    import ctypes as C
    n = len(ranks)
    ctxs = (C.c_void_p * n)(*[r.ctx.h for r in ranks])
    slabs = (C.c_int * n)(*[r.h for r in ranks])
    ranks[0].cwa.check(ranks[0].ctx.lib.cwa_slab_group_step(ctxs, slabs, n, nframes, coupling))


def connect_processes(rank_obj: SlabRank, dist):
    """One process per GPU: exchange the mailboxes' IPC handles once (torch.distributed is plumbing here, nothing per frame)."""
    world = rank_obj.desc.world
    if world == 1:
        return
    handles = [None] * world
    dist.all_gather_object(handles, rank_obj.export())
    for r, hnd in enumerate(handles):
        if r != rank_obj.desc.rank:
            rank_obj.connect(r, handle=hnd)
    dist.barrier()
