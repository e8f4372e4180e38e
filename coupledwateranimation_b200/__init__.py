"""coupledwateranimation_b200 -- B200-native (sm_100a CUDA) per-frame simulation step of
VarunRamakri7/CoupledWaterAnimation behind the reference's Module / ComputeShader / Buffer /
StencilImage2DTripleBuffered surface.

The product is the C-ABI shared library ``libcwa_b200.so`` (include/cwa_b200.h); the C++ mirrors
of the reference classes live in include/cwa/.  This Python package is the thin ctypes harness
the tests and bench.py drive the library with -- object names follow the reference:

    Context                      the GL-context stand-in (one CUDA stream)
    Buffer                       Buffer                        (SphWave2D/Buffer.h)
    UniformGrid                  UniformGridSph2D / UgridParticles3D
    StencilImage2DTripleBuffered wave height field, triple buffered
    Sph                          the three 3-D SPH compute programs on the particle SSBO
    SphUgrid                     2-D Koschier SPH on the grid (SphWave2D/StencilBuffer.h)
    ComputeShader                ComputeShader::Init/SetMode/Dispatch by GLSL file name

There is no CPU / PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import CwaError, check

PARTICLE = np.dtype([("pos", "<f4", 4), ("vel", "<f4", 4), ("force", "<f4", 4), ("extras", "<f4", 4)])
PARTICLE2D = np.dtype([("pos", "<f4", 4), ("vel", "<f4", 4), ("acc", "<f4", 4)])

TARGET_SSBO, TARGET_UBO = 0, 1
UBO_CONSTANTS, UBO_BOUNDARY, UBO_WAVE, UBO_SIM = 1, 2, 3, 4
MODE_INIT, MODE_INIT_FROM_TEXTURE, MODE_EVOLVE, MODE_TEST = 0, 1, 2, 10
WAVE_COUPLED, WAVE_SIMP = 0, 1
COUPLING_AS_SHIPPED, COUPLING_LATEST = 0, 1
SPH2_KOSCHIER, SPH2_WAVE = 0, 1
GRID_COUNTER, GRID_OFFSET, GRID_INDEX_LIST, GRID_CELL_OF = 0, 1, 2, 3
STENCIL1D_SHALLOW, STENCIL1D_WAVE = 0, 1
BC_REFLECT, BC_FREE, BC_FIXED = 0, 1, 2

__all__ = [
    "Context", "Buffer", "UniformGrid", "StencilImage2DTripleBuffered", "ImageStencil", "Sph", "SphUgrid", "ComputeShader",
    "CwaError", "PARTICLE", "PARTICLE2D",
]


def _vp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """cwa_create / cwa_destroy.  Fails loudly when no sm_100 device is present."""

    def __init__(self, device: int = 0):
        self.lib = _capi.load()
        h = C.c_void_p()
        check(self.lib.cwa_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.cwa_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self.lib.cwa_synchronize(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.cwa_stream(self.h) or 0)

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        check(self.lib.cwa_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}

    def timer_begin(self):
        check(self.lib.cwa_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_float()
        check(self.lib.cwa_timer_end(self.h, C.byref(ms)))
        return ms.value

    @property
    def launch_count(self) -> int:
        return int(self.lib.cwa_launch_count(self.h))

    def set_tuning(self, **kv):
        """cwa_set_tuning: nb_config / nb_cap_d / nb_cap_f / nb_cap_r (process-wide kernel-variant knobs)."""
        for k, v in kv.items():
            check(self.lib.cwa_set_tuning(self.h, k.encode(), int(v)))

    def profile_begin(self):
        check(self.lib.cwa_profile_begin(self.h))

    def profile_end(self) -> dict:
        """{kernel name: (total ms, launches)} for every launch since profile_begin()."""
        k = self.lib.cwa_profile_kernel_count()
        ms = (C.c_float * k)(); cnt = (C.c_int * k)()
        check(self.lib.cwa_profile_end(self.h, ms, cnt, k))
        return {self.lib.cwa_profile_kernel_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(k) if cnt[i] > 0}

    # -- parameter blocks (UBO bindings 1..4) -----------------------------------------------------
    def default_ubo(self, binding: int) -> "Buffer":
        b = C.c_int()
        check(self.lib.cwa_default_ubo(self.h, binding, C.byref(b)))
        sizes = {1: 16, 2: 32, 3: 32, 4: 48}
        return Buffer(self, handle=b.value, nbytes=sizes[binding])

    def set_constants(self, mass=0.02, smoothing_coeff=2.0, visc=3000.0, resting_rho=1000.0):
        a = np.array([mass, smoothing_coeff, visc, resting_rho], np.float32)
        self.default_ubo(UBO_CONSTANTS).sub_data(a)

    def set_boundary(self, upper=(0.48, 1.0, 0.48, 500.0), lower=(0.0, -0.02, 0.0, 50.0)):
        a = np.array(list(upper) + list(lower), np.float32)
        self.default_ubo(UBO_BOUNDARY).sub_data(a)

    def set_wave_uniforms(self, attributes=(0.01, 0.985, 0.001, 1.0), mesh_ws_pos=(2.0, 0.35, -1.0, 0.0)):
        a = np.array(list(attributes) + list(mesh_ws_pos), np.float32)
        self.default_ubo(UBO_WAVE).sub_data(a)

    def set_sim_constants(self, particle_radius=0.005, gas_const=4000.0, dt=0.00005, gravity_y=-9806.65,
                          damping=0.3, crest_threshold=0.01, foam_speed=25.0, uv_scale=2.0, uv_scale_z=0.0, torque_coeff=0.0):
        a = np.array([particle_radius, gas_const, dt, gravity_y, damping, crest_threshold, foam_speed, uv_scale, uv_scale_z, torque_coeff, 0.0, 0.0], np.float32)
        self.default_ubo(UBO_SIM).sub_data(a)

    def set_params_from_oracle(self, prm):
        """Copy an oracle Params3 struct (tests only pass it in; nothing here imports the oracle)."""
        self.set_constants(prm.mass, prm.smoothing_coeff, prm.visc, prm.resting_rho)
        self.set_boundary(tuple(prm.upper), tuple(prm.lower))
        self.set_wave_uniforms(tuple(prm.attributes), tuple(prm.mesh_ws_pos))
        self.set_sim_constants(prm.particle_radius, prm.gas_const, prm.dt, prm.gravity_y, prm.damping,
                               prm.crest_threshold, prm.foam_speed, prm.uv_scale, getattr(prm, "uv_scale_z", 0.0), getattr(prm, "torque_coeff", 0.0))

    # ---- parameter reflection (name -> field of the bound parameter blocks) ---------------------------------------------
    def params(self) -> list:
        """[(name, ubo_binding, byte_offset, origin)] of every run-time parameter."""
        out = []
        for i in range(self.lib.cwa_param_count()):
            name, origin = C.c_char_p(), C.c_char_p()
            ubo, off = C.c_int(), C.c_int()
            check(self.lib.cwa_param_info(i, C.byref(name), C.byref(ubo), C.byref(off), C.byref(origin)))
            out.append((name.value.decode(), ubo.value, off.value, origin.value.decode()))
        return out

    def param_set(self, name: str, value: float):
        check(self.lib.cwa_param_set(self.h, name.encode(), float(value)))

    def param_get(self, name: str) -> float:
        v = C.c_float()
        check(self.lib.cwa_param_get(self.h, name.encode(), C.byref(v)))
        return v.value

    def checkpoint_save(self, sph: "Sph", wave: "StencilImage2DTripleBuffered", frame: int, path: str):
        check(self.lib.cwa_checkpoint_save(self.h, sph.h, wave.h, int(frame), path.encode()))

    def checkpoint_load(self, sph: "Sph", wave: "StencilImage2DTripleBuffered", path: str) -> int:
        fr = C.c_ulonglong()
        check(self.lib.cwa_checkpoint_load(self.h, sph.h, wave.h, path.encode(), C.byref(fr)))
        return fr.value

    def bind_scene(self, sph: "Sph | None", wave: "StencilImage2DTripleBuffered | None"):
        check(self.lib.cwa_bind_scene(self.h, sph.h if sph else -1, wave.h if wave else -1))

    def sph_step(self, nsteps=1):
        check(self.lib.sph_step(self.h, nsteps))

    def wave_step(self, nsteps=1):
        check(self.lib.wave_step(self.h, nsteps))

    def scan_exclusive(self, src: "Buffer", dst: "Buffer", n: int):
        check(self.lib.cwa_scan_exclusive(self.h, src.h, dst.h, n))


class Buffer:
    """Buffer::Init / BufferSubData / BindBufferBase / DebugRead* (SphWave2D/Buffer.cpp:5-83)."""

    def __init__(self, ctx: Context, nbytes: int | None = None, data: np.ndarray | None = None, handle: int | None = None,
                 device_ptr: int | None = None):
        self.ctx = ctx
        self.owned = handle is None
        if handle is not None:
            self.h = handle
            self.nbytes = nbytes
            return
        b = C.c_int(-1)
        if device_ptr is not None:
            check(ctx.lib.cwa_buffer_wrap(ctx.h, C.c_void_p(device_ptr), nbytes, C.byref(b)))
        else:
            if data is not None:
                data = np.ascontiguousarray(data)
                nbytes = data.nbytes
            check(ctx.lib.cwa_buffer_create(ctx.h, nbytes, _vp(data) if data is not None else None, C.byref(b)))
            if data is not None:
                ctx.synchronize()     # the H2D copy reads `data` asynchronously
        self.h = b.value
        self.nbytes = nbytes

    def sub_data(self, data: np.ndarray, offset: int = 0):
        data = np.ascontiguousarray(data)
        check(self.ctx.lib.cwa_buffer_sub_data(self.ctx.h, self.h, offset, data.nbytes, _vp(data)))
        self.ctx.synchronize()

    def read(self, dtype=np.uint8, count: int | None = None, offset: int = 0) -> np.ndarray:
        dtype = np.dtype(dtype)
        if count is None:
            count = (self.nbytes - offset) // dtype.itemsize
        out = np.empty(count, dtype)
        check(self.ctx.lib.cwa_buffer_read(self.ctx.h, self.h, offset, out.nbytes, _vp(out)))
        return out

    def bind_base(self, target: int, binding: int):
        check(self.ctx.lib.cwa_buffer_bind_base(self.ctx.h, target, binding, self.h))

    def device_ptr(self) -> int:
        p, n = C.c_void_p(), C.c_size_t()
        check(self.ctx.lib.cwa_buffer_device_ptr(self.ctx.h, self.h, C.byref(p), C.byref(n)))
        return int(p.value or 0)

    def destroy(self):
        if self.owned and self.h >= 0:
            check(self.ctx.lib.cwa_buffer_destroy(self.ctx.h, self.h))
            self.h = -1


class UniformGrid:
    """UniformGridSph2D (dim 2) / UgridParticles3D (dim 3): ctor + Init + CollisionQuery/BuildGrid."""

    def __init__(self, ctx: Context, dim: int, mn, mx, num_cells, max_particles: int, compact_index: bool = False):
        self.ctx, self.dim, self.max_particles = ctx, dim, max_particles
        g = C.c_int(-1)
        mn_a = (C.c_float * dim)(*mn); mx_a = (C.c_float * dim)(*mx); nc_a = (C.c_int * dim)(*num_cells)
        # compact_index: k-stride Nz instead of the reference's Nx (CWA_GRID_COMPACT_INDEX), for slab-local grids
        check(ctx.lib.cwa_grid_create(ctx.h, dim | (16 if compact_index else 0), mn_a, mx_a, nc_a, max_particles, C.byref(g)))
        self.h = g.value
        info, total = _capi.GridInfo(), C.c_int()
        check(ctx.lib.cwa_grid_get_info(ctx.h, self.h, C.byref(info), C.byref(total)))
        self.info = info
        self.num_cells_total = total.value
        self.num_cells = tuple(info.num_cells[:dim])
        self.cell_size = tuple(info.cell_size[:dim])

    def build(self, particles: Buffer, stride_bytes: int, n: int):
        check(self.ctx.lib.cwa_grid_build(self.ctx.h, self.h, particles.h, stride_bytes, n))

    def read(self, which: int, count: int) -> np.ndarray:
        out = np.empty(count, np.int32)
        check(self.ctx.lib.cwa_grid_read(self.ctx.h, self.h, which, _vp(out), count))
        return out

    def buffer(self, which: int) -> Buffer:
        b = C.c_int()
        check(self.ctx.lib.cwa_grid_buffer(self.ctx.h, self.h, which, C.byref(b)))
        return Buffer(self.ctx, handle=b.value, nbytes=None)

    def destroy(self):
        check(self.ctx.lib.cwa_grid_destroy(self.ctx.h, self.h))


class StencilImage2DTripleBuffered:
    """StencilImage2DTripleBuffered (CoupledWaterAnimation/StencilImage2DTripleBuffered.cpp:4-95)."""

    def __init__(self, ctx: Context, w=64, h=64, channels=1, variant=WAVE_COUPLED):
        self.ctx, self.w, self.height, self.ch, self.variant = ctx, w, h, channels, variant
        o = C.c_int(-1)
        check(ctx.lib.cwa_wave_create(ctx.h, w, h, channels, variant, C.byref(o)))   # Init() incl. Reinit()
        self.h = o.value

    @classmethod
    def create_block(cls, ctx: Context, w: int, h_global: int, row0: int, rows: int, channels=1, variant=WAVE_COUPLED):
        """Row block [row0, row0+rows) of a field h_global rows tall (multi-GPU decomposition)."""
        self = cls.__new__(cls)
        self.ctx, self.w, self.height, self.ch, self.variant = ctx, w, rows, channels, variant
        self.h_global, self.row0 = h_global, row0
        o = C.c_int(-1)
        check(ctx.lib.cwa_wave_create_block(ctx.h, w, h_global, row0, rows, channels, variant, C.byref(o)))
        self.h = o.value
        return self

    def last_row_buffer(self, image: int) -> Buffer:
        b = C.c_int()
        check(self.ctx.lib.cwa_wave_last_row_buffer(self.ctx.h, self.h, image, C.byref(b)))
        return Buffer(self.ctx, handle=b.value, nbytes=self.w * self.ch * 4)

    def _shape(self):
        return (self.height, self.w) if self.ch == 1 else (self.height, self.w, self.ch)

    def Reinit(self):
        check(self.ctx.lib.cwa_wave_reinit(self.ctx.h, self.h))

    def ReinitFromTexture(self, rgba: np.ndarray):
        rgba = np.ascontiguousarray(rgba, np.float32)
        th, tw = rgba.shape[0], rgba.shape[1]
        check(self.ctx.lib.cwa_wave_reinit_from_texture(self.ctx.h, self.h, _vp(rgba), tw, th))

    def Compute(self, nsteps=1):
        check(self.ctx.lib.cwa_wave_compute(self.ctx.h, self.h, nsteps))

    def PingPong(self):
        check(self.ctx.lib.cwa_wave_pingpong(self.ctx.h, self.h))

    def set_evolve(self, on: bool):
        check(self.ctx.lib.cwa_wave_set_evolve(self.ctx.h, self.h, int(on)))

    def set_params(self, lam, atten, beta):
        check(self.ctx.lib.cwa_wave_set_params(self.ctx.h, self.h, lam, atten, beta))

    def resize(self, w, h):
        check(self.ctx.lib.cwa_wave_resize(self.ctx.h, self.h, w, h))
        self.w, self.height = w, h

    def state(self):
        ri = (C.c_int * 2)(); wi = C.c_int(); un = (C.c_int * 3)(); t0 = C.c_int()
        check(self.ctx.lib.cwa_wave_state(self.ctx.h, self.h, ri, C.byref(wi), un, C.byref(t0)))
        return {"read_index": list(ri), "write_index": wi.value, "unit": list(un), "tex_unit0": t0.value}

    def bind_texture_unit(self):
        check(self.ctx.lib.cwa_wave_bind_texture_unit(self.ctx.h, self.h))

    def role_image(self, role: int) -> int:
        i = C.c_int()
        check(self.ctx.lib.cwa_wave_role_image(self.ctx.h, self.h, role, C.byref(i)))
        return i.value

    def read_image(self, image: int) -> np.ndarray:
        out = np.empty(self._shape(), np.float32)
        check(self.ctx.lib.cwa_wave_read_image(self.ctx.h, self.h, image, _vp(out)))
        return out

    def write_image(self, image: int, data: np.ndarray):
        data = np.ascontiguousarray(data, np.float32)
        assert data.shape == self._shape(), (data.shape, self._shape())
        check(self.ctx.lib.cwa_wave_write_image(self.ctx.h, self.h, image, _vp(data)))
        self.ctx.synchronize()

    def read_role(self, role: int) -> np.ndarray:
        return self.read_image(self.role_image(role))

    def write_role(self, role: int, data: np.ndarray):
        self.write_image(self.role_image(role), data)

    def image_buffer(self, image: int) -> Buffer:
        b = C.c_int()
        check(self.ctx.lib.cwa_wave_image_buffer(self.ctx.h, self.h, image, C.byref(b)))
        return Buffer(self.ctx, handle=b.value, nbytes=self.w * self.height * self.ch * 4)

    def destroy(self):
        check(self.ctx.lib.cwa_wave_destroy(self.ctx.h, self.h))


class Sph:
    """The particle SSBO + the three 3-D compute programs of Main.cpp:540-557."""

    def __init__(self, ctx: Context, n: int, grid: UniformGrid | None = None, particles: np.ndarray | None = None,
                 buffer: Buffer | None = None):
        self.ctx, self.n, self.grid = ctx, n, grid
        if buffer is None:
            if particles is not None:
                assert particles.dtype == PARTICLE and particles.size == n
                buffer = Buffer(ctx, data=particles)
            else:
                buffer = Buffer(ctx, nbytes=max(n, 1) * PARTICLE.itemsize)
        self.buffer = buffer
        o = C.c_int(-1)
        check(ctx.lib.cwa_sph_create(ctx.h, buffer.h, n, grid.h if grid is not None else -1, C.byref(o)))
        self.h = o.value

    def upload(self, particles: np.ndarray):
        assert particles.dtype == PARTICLE and particles.size == self.n
        self.buffer.sub_data(particles)

    def download(self) -> np.ndarray:
        return self.buffer.read(PARTICLE, self.n)

    def bind_wave(self, wave: StencilImage2DTripleBuffered | None, image: int = -1):
        check(self.ctx.lib.cwa_sph_bind_wave(self.ctx.h, self.h, wave.h if wave else -1, image))

    def rho_pres(self):
        check(self.ctx.lib.cwa_sph_rho_pres(self.ctx.h, self.h))

    def force(self):
        check(self.ctx.lib.cwa_sph_force(self.ctx.h, self.h))

    def integrate(self):
        check(self.ctx.lib.cwa_sph_integrate(self.ctx.h, self.h))

    def step(self, nsteps=1):
        check(self.ctx.lib.cwa_sph_step(self.ctx.h, self.h, nsteps))

    def neighbour_count(self) -> np.ndarray:
        out = np.empty(self.n, np.int32)
        check(self.ctx.lib.cwa_sph_neighbour_count(self.ctx.h, self.h, _vp(out)))
        return out

    def init_cube(self, nx, ny, nz):
        check(self.ctx.lib.cwa_sph_init_cube(self.ctx.h, self.h, nx, ny, nz))

    def coupled_step(self, wave: StencilImage2DTripleBuffered, nframes=1, coupling=COUPLING_AS_SHIPPED):
        check(self.ctx.lib.cwa_coupled_step(self.ctx.h, self.h, wave.h, nframes, coupling))

    def destroy(self):
        check(self.ctx.lib.cwa_sph_destroy(self.ctx.h, self.h))


class SphUgrid:
    """SphUgrid (SphWave2D/StencilBuffer.cpp:138-179): 2-D Koschier SPH on the uniform grid."""

    def __init__(self, ctx: Context, n: int, grid: UniformGrid, variant=SPH2_WAVE, substeps=2):
        self.ctx, self.n, self.grid, self.variant = ctx, n, grid, variant
        o = C.c_int(-1)
        check(ctx.lib.cwa_sph2_create(ctx.h, n, variant, grid.h, C.byref(o)))
        self.h = o.value
        self.SetSubsteps(substeps)

    def SetSubsteps(self, s: int):
        check(self.ctx.lib.cwa_sph2_set_substeps(self.ctx.h, self.h, s))

    def set_uniforms(self, time=0.0, bottom=0.3, psi=-1.0, init_width=0, view_width=None):
        check(self.ctx.lib.cwa_sph2_set_uniforms(self.ctx.h, self.h, time, bottom, psi, init_width))
        if view_width is not None:
            check(self.ctx.lib.cwa_sph2_set_view_width(self.ctx.h, self.h, view_width))

    def Reinit(self):
        check(self.ctx.lib.cwa_sph2_reinit(self.ctx.h, self.h))

    def bind_wave1d(self, buf: Buffer | None, width: int = 0):
        check(self.ctx.lib.cwa_sph2_bind_wave1d(self.ctx.h, self.h, buf.h if buf else -1, width))

    def Compute(self, nframes=1):
        check(self.ctx.lib.cwa_sph2_compute(self.ctx.h, self.h, nframes))

    def download(self) -> np.ndarray:
        out = np.empty(self.n, PARTICLE2D)
        check(self.ctx.lib.cwa_sph2_read(self.ctx.h, self.h, _vp(out)))
        return out

    def upload(self, p: np.ndarray):
        assert p.dtype == PARTICLE2D and p.size == self.n
        p = np.ascontiguousarray(p)
        check(self.ctx.lib.cwa_sph2_write(self.ctx.h, self.h, _vp(p)))
        self.ctx.synchronize()

    def destroy(self):
        check(self.ctx.lib.cwa_sph2_destroy(self.ctx.h, self.h))


class ImageStencil:
    """ImageStencil (SphWave2D/StencilImage2D.h:10-66) on a 1-D RGBA32F image, driving Shallow1D_cs.glsl (double buffered,
    modes 2,3 per frame) or Wave1D_cs.glsl (triple buffered, 10 substeps of mode 2) -- the wave the 2-D SPH couples to."""

    def __init__(self, ctx: Context, shader: int, width: int):
        self.ctx, self.shader, self.w = ctx, shader, width
        o = C.c_int(-1)
        check(ctx.lib.cwa_stencil1d_create(ctx.h, shader, width, C.byref(o)))
        self.h = o.value
        self.num_images = self.state()["num_images"]

    def Reinit(self):
        check(self.ctx.lib.cwa_stencil1d_reinit(self.ctx.h, self.h))

    def ReinitFromTexture(self, rgba: np.ndarray):
        rgba = np.ascontiguousarray(rgba, np.float32).reshape(-1, 4)
        check(self.ctx.lib.cwa_stencil1d_reinit_from_texture(self.ctx.h, self.h, _vp(rgba), rgba.shape[0]))

    def Compute(self, nframes=1):
        check(self.ctx.lib.cwa_stencil1d_compute(self.ctx.h, self.h, nframes))

    def ComputeFunc(self, mode: int):
        check(self.ctx.lib.cwa_stencil1d_compute_func(self.ctx.h, self.h, mode))

    def PingPong(self):
        check(self.ctx.lib.cwa_stencil1d_pingpong(self.ctx.h, self.h))

    def SetSubsteps(self, s: int):
        check(self.ctx.lib.cwa_stencil1d_set_substeps(self.ctx.h, self.h, s))

    def set_iterate(self, on: bool):
        check(self.ctx.lib.cwa_stencil1d_set_iterate(self.ctx.h, self.h, int(on)))

    def set_params(self, lam, dx_or_atten, beta, boundary=(0.0, 0.0), bc=BC_FREE):
        check(self.ctx.lib.cwa_stencil1d_set_params(self.ctx.h, self.h, lam, dx_or_atten, beta, boundary[0], boundary[1], bc))

    def state(self):
        n = C.c_int(); ri = (C.c_int * 2)(); wi = C.c_int(); un = (C.c_int * 3)()
        check(self.ctx.lib.cwa_stencil1d_state(self.ctx.h, self.h, C.byref(n), ri, C.byref(wi), un))
        return {"num_images": n.value, "read_index": list(ri)[:n.value - 1], "write_index": wi.value, "unit": list(un)[:n.value]}

    def read_image(self, image: int) -> np.ndarray:
        out = np.empty((self.w, 4), np.float32)
        check(self.ctx.lib.cwa_stencil1d_read_image(self.ctx.h, self.h, image, _vp(out)))
        return out

    def write_image(self, image: int, rgba: np.ndarray):
        rgba = np.ascontiguousarray(rgba, np.float32)
        assert rgba.shape == (self.w, 4)
        check(self.ctx.lib.cwa_stencil1d_write_image(self.ctx.h, self.h, image, _vp(rgba)))

    def GetReadImage(self, i: int = 0) -> np.ndarray:
        return self.read_image(self.state()["read_index"][i])

    def read_buffer(self, i: int = 0) -> Buffer:
        """GetReadImage(i) as a Buffer (what SphUgrid.bind_wave1d takes: wave1d.GetReadImage(0).BindTextureUnit(), Main.cpp:240)."""
        b = C.c_int()
        check(self.ctx.lib.cwa_stencil1d_image_buffer(self.ctx.h, self.h, self.state()["read_index"][i], C.byref(b)))
        return Buffer(self.ctx, handle=b.value, nbytes=self.w * 16)

    def destroy(self):
        check(self.ctx.lib.cwa_stencil1d_destroy(self.ctx.h, self.h))


class ComputeShader:
    """ComputeShader::Init / SetMode / Dispatch by GLSL file name (ComputeShader.cpp:9-56)."""

    def __init__(self, ctx: Context, filename: str):
        self.ctx, self.filename = ctx, filename
        o = C.c_int(-1)
        check(ctx.lib.cwa_shader_create(ctx.h, filename.encode(), C.byref(o)))
        self.h = o.value

    def SetMode(self, mode: int):
        check(self.ctx.lib.cwa_shader_set_mode(self.ctx.h, self.h, mode))

    def set_uniform_i(self, loc: int, v: int):
        check(self.ctx.lib.cwa_shader_set_uniform_i(self.ctx.h, self.h, loc, v))

    def set_uniform_f(self, loc: int, v: float):
        check(self.ctx.lib.cwa_shader_set_uniform_f(self.ctx.h, self.h, loc, v))

    def bind_object(self, obj):
        check(self.ctx.lib.cwa_shader_bind_object(self.ctx.h, self.h, obj.h))

    def Dispatch(self, gx=1, gy=1, gz=1):
        check(self.ctx.lib.cwa_shader_dispatch(self.ctx.h, self.h, gx, gy, gz))
