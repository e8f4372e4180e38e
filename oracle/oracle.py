"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of the reference's own CPU grid
(oracle/_ref/libref_grid.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

PARTICLE3 = np.dtype([("pos", "<f4", 4), ("vel", "<f4", 4), ("force", "<f4", 4), ("extras", "<f4", 4)])
PARTICLE2 = np.dtype([("pos", "<f4", 4), ("vel", "<f4", 4), ("acc", "<f4", 4)])

COUPLING_AS_SHIPPED, COUPLING_LATEST = 0, 1
WAVE_COUPLED, WAVE_SIMP = 0, 1
SPH2_KOSCHIER, SPH2_WAVE = 0, 1


class Params3(C.Structure):
    _fields_ = [
        ("mass", C.c_float), ("smoothing_coeff", C.c_float), ("visc", C.c_float), ("resting_rho", C.c_float),
        ("upper", C.c_float * 4), ("lower", C.c_float * 4),
        ("attributes", C.c_float * 4), ("mesh_ws_pos", C.c_float * 4),
        ("particle_radius", C.c_float), ("gas_const", C.c_float), ("dt", C.c_float), ("gravity_y", C.c_float),
        ("damping", C.c_float), ("crest_threshold", C.c_float), ("foam_speed", C.c_float), ("uv_scale", C.c_float),
        ("uv_scale_z", C.c_float), ("torque_coeff", C.c_float),
    ]


class Stencil1DParams(C.Structure):
    _fields_ = [("lam", C.c_float), ("dx_or_atten", C.c_float), ("beta", C.c_float), ("boundary", C.c_float * 2), ("bc", C.c_int)]


class Params2(C.Structure):
    _fields_ = [("variant", C.c_int), ("time", C.c_float), ("bottom", C.c_float), ("psi", C.c_float),
                ("init_width", C.c_int), ("view_width", C.c_float)]


class Tex(C.Structure):
    _fields_ = [("data", C.c_void_p), ("w", C.c_int), ("h", C.c_int), ("ch", C.c_int)]


class Grid2(C.Structure):
    _fields_ = [("min", C.c_float * 2), ("max", C.c_float * 2), ("ncells", C.c_int * 2), ("cell", C.c_float * 2)]


class Grid3(C.Structure):
    _fields_ = [("min", C.c_float * 4), ("max", C.c_float * 4), ("ncells", C.c_int * 4), ("cell", C.c_float * 4)]


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/ when the reference tree is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("cwa_oracle.c", "cwa_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale or (os.path.isdir("/root/reference") and not os.path.exists(os.path.join(_HERE, "_ref", "libref_grid.so"))):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


_lib = None
_ref = None


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.orc_tex_bilinear.restype = C.c_float
        L.orc_tex_bilinear.argtypes = [C.POINTER(Tex), C.c_float, C.c_float]
        L.orc_scan_blelloch.restype = C.c_int
        L.orc_grid2_cell_index.restype = C.c_int
        L.orc_grid2_cell_index.argtypes = [C.POINTER(Grid2), C.c_float, C.c_float]
        L.orc_grid3_cell_index.restype = C.c_int
        L.orc_grid3_cell_index.argtypes = [C.POINTER(Grid3), C.c_float, C.c_float, C.c_float]
        L.orc_wave_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        L.orc_wave_evolve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_float, C.c_float, C.c_float, C.c_float]
        L.orc_sph3_neighbour_count.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
        L.orc_coupled_create.restype = C.c_void_p
        L.orc_coupled_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Params3), C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_coupled_destroy.argtypes = [C.c_void_p]
        L.orc_stencil1d_params_default.argtypes = [C.POINTER(Stencil1DParams), C.c_int]
        L.orc_shallow1d_dispatch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Stencil1DParams)]
        L.orc_wave1d_dispatch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Stencil1DParams)]
        L.orc_coupled_particles.restype = C.c_void_p
        L.orc_coupled_particles.argtypes = [C.c_void_p]
        L.orc_coupled_wave.restype = C.c_void_p
        L.orc_coupled_wave.argtypes = [C.c_void_p, C.c_int]
        L.orc_coupled_wave_reinit.argtypes = [C.c_void_p]
        L.orc_coupled_step.argtypes = [C.c_void_p, C.c_int]
        L.orc_coupled_sampled_image.restype = C.c_int
        L.orc_coupled_sampled_image.argtypes = [C.c_void_p]
        L.orc_coupled_set_params.argtypes = [C.c_void_p, C.POINTER(Params3)]
        L.orc_sph2_step.restype = C.c_int
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def ref_lib():
    """The reference's own CPU grid (None when oracle/_ref was never built)."""
    global _ref
    if _ref is None:
        p = os.path.join(_HERE, "_ref", "libref_grid.so")
        if not os.path.exists(p):
            try:
                build()
            except Exception:
                pass
        if not os.path.exists(p):
            return None
        _ref = C.CDLL(p)
        _ref.ref_grid2d_build.restype = C.c_int
        _ref.ref_grid2d_query.restype = C.c_int
    return _ref


# ------------------------------------------------------------------------------------------------
# thin numpy-level helpers
# ------------------------------------------------------------------------------------------------
def default_params3() -> Params3:
    p = Params3()
    lib().orc_params3_default(C.byref(p))
    return p


def default_params2(variant: int) -> Params2:
    p = Params2()
    lib().orc_params2_default(C.byref(p), C.c_int(variant))
    return p


def make_tex(arr) -> Tex:
    """arr: float32 [H,W] or [H,W,C] (or [W,C] for 1-D) or None (unbound)."""
    t = Tex()
    if arr is None:
        t.data, t.w, t.h, t.ch = None, 1, 1, 1
        return t
    assert arr.dtype == np.float32 and arr.flags.c_contiguous
    if arr.ndim == 2:
        h, w, ch = arr.shape[0], arr.shape[1], 1
    else:
        h, w, ch = arr.shape
    t.data, t.w, t.h, t.ch = arr.ctypes.data, w, h, ch
    t._keep = arr
    return t


def tex_bilinear(arr, s: float, t: float) -> float:
    tx = make_tex(arr)
    return float(lib().orc_tex_bilinear(C.byref(tx), C.c_float(s), C.c_float(t)))


def scan_blelloch(x: np.ndarray) -> np.ndarray:
    y = np.ascontiguousarray(x, dtype=np.int32).copy()
    rc = lib().orc_scan_blelloch(_ptr(y), C.c_int(y.size))
    if rc != 0:
        raise ValueError("n must be a power of two >= 2 (ParallelScan.cpp:15-16)")
    return y


def scan_exclusive(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.int32)
    y = np.empty_like(x)
    lib().orc_scan_exclusive(_ptr(x), _ptr(y), C.c_int(x.size))
    return y


def grid2(mn, mx, n) -> Grid2:
    g = Grid2()
    lib().orc_grid2_init(C.byref(g), (C.c_float * 2)(*mn), (C.c_float * 2)(*mx), (C.c_int * 2)(*n))
    return g


def grid3(mn, mx, n) -> Grid3:
    g = Grid3()
    lib().orc_grid3_init(C.byref(g), (C.c_float * 3)(*mn), (C.c_float * 3)(*mx), (C.c_int * 3)(*n))
    return g


def grid2_build(g: Grid2, pos: np.ndarray):
    """pos: float32 [n, stride>=2] rows; returns cell_of, counter, offset, index_list (-1 padded)."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    n, stride = pos.shape
    ncell = g.ncells[0] * g.ncells[1]
    cell_of = np.empty(n, np.int32); counter = np.empty(ncell, np.int32); offset = np.empty(ncell, np.int32)
    index_list = np.full(n, -1, np.int32)
    lib().orc_grid2_build(C.byref(g), _ptr(pos), C.c_int(stride), C.c_int(n), _ptr(cell_of), _ptr(counter),
                          _ptr(offset), _ptr(index_list))
    return cell_of, counter, offset, index_list


def grid3_build(g: Grid3, pos: np.ndarray):
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    n, stride = pos.shape
    ncell = g.ncells[0] * g.ncells[1] * g.ncells[2]
    cell_of = np.empty(n, np.int32); counter = np.empty(ncell, np.int32); offset = np.empty(ncell, np.int32)
    index_list = np.full(n, -1, np.int32)
    lib().orc_grid3_build(C.byref(g), _ptr(pos), C.c_int(stride), C.c_int(n), _ptr(cell_of), _ptr(counter),
                          _ptr(offset), _ptr(index_list))
    return cell_of, counter, offset, index_list


def wave_init(w, h, ch=1, variant=WAVE_COUPLED, wtype=1.0) -> np.ndarray:
    out = np.empty((h, w) if ch == 1 else (h, w, ch), np.float32)
    lib().orc_wave_init(_ptr(out), w, h, ch, variant, C.c_float(wtype))
    return out


def wave_evolve(u0, u1, variant=WAVE_COUPLED, lam=0.01, atten=0.985, beta=0.001, wtype=1.0) -> np.ndarray:
    u0 = np.ascontiguousarray(u0, np.float32); u1 = np.ascontiguousarray(u1, np.float32)
    h, w = u0.shape[0], u0.shape[1]
    ch = 1 if u0.ndim == 2 else u0.shape[2]
    out = np.empty_like(u0)
    lib().orc_wave_evolve(_ptr(u0), _ptr(u1), _ptr(out), w, h, ch, variant, C.c_float(lam), C.c_float(atten),
                          C.c_float(beta), C.c_float(wtype))
    return out


def make_cube(nx, ny, nz, prm: Params3 | None = None) -> np.ndarray:
    prm = prm or default_params3()
    p = np.zeros(nx * ny * nz, PARTICLE3)
    lib().orc_make_cube(_ptr(p), nx, ny, nz, C.byref(prm))
    return p


def _grid_args(grid):
    if grid is None:
        return None, None, None, None
    g, counter, offset, index_list = grid
    return C.byref(g), _ptr(counter), _ptr(offset), _ptr(index_list)


def sph3_rho_pres(p, prm, tex_arr, grid=None):
    tx = make_tex(tex_arr); g, c, o, l = _grid_args(grid)
    lib().orc_sph3_rho_pres(_ptr(p), C.c_int(p.size), C.byref(prm), C.byref(tx), g, c, o, l)


def sph3_force(p, prm, tex_arr, grid=None):
    tx = make_tex(tex_arr); g, c, o, l = _grid_args(grid)
    lib().orc_sph3_force(_ptr(p), C.c_int(p.size), C.byref(prm), C.byref(tx), g, c, o, l)


def sph3_integrate(p, prm, tex_arr):
    tx = make_tex(tex_arr)
    lib().orc_sph3_integrate(_ptr(p), C.c_int(p.size), C.byref(prm), C.byref(tx))


def sph3_neighbour_count(p, h, grid=None) -> np.ndarray:
    out = np.empty(p.size, np.int32)
    g, c, o, l = _grid_args(grid)
    lib().orc_sph3_neighbour_count(_ptr(p), C.c_int(p.size), C.c_float(h), g, c, o, l, _ptr(out))
    return out


class Coupled:
    """orc_coupled driver: one frame = rho -> force -> integrate -> wave evolve -> display bind."""

    def __init__(self, n, wave_w, wave_h, wave_ch=1, prm=None, coupling=COUPLING_AS_SHIPPED, grid=None):
        self.prm = prm or default_params3()
        self.n, self.w, self.h, self.ch = n, wave_w, wave_h, wave_ch
        if grid is None:
            self._h = lib().orc_coupled_create(n, wave_w, wave_h, wave_ch, C.byref(self.prm), coupling, 0, None, None, None)
        else:
            mn, mx, nc = grid
            self._h = lib().orc_coupled_create(n, wave_w, wave_h, wave_ch, C.byref(self.prm), coupling, 1,
                                               (C.c_float * 3)(*mn), (C.c_float * 3)(*mx), (C.c_int * 3)(*nc))

    @property
    def particles(self) -> np.ndarray:
        ptr = lib().orc_coupled_particles(self._h)
        buf = (C.c_char * (self.n * PARTICLE3.itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=PARTICLE3)

    def wave(self, role=0) -> np.ndarray:
        ptr = lib().orc_coupled_wave(self._h, role)
        cnt = self.w * self.h * self.ch
        buf = (C.c_float * cnt).from_address(ptr)
        a = np.frombuffer(buf, dtype=np.float32)
        return a.reshape((self.h, self.w) if self.ch == 1 else (self.h, self.w, self.ch))

    def step(self, nframes=1):
        lib().orc_coupled_step(self._h, nframes)

    def reinit_wave(self):
        lib().orc_coupled_wave_reinit(self._h)

    def sampled_image(self) -> int:
        return lib().orc_coupled_sampled_image(self._h)

    def set_params(self, prm):
        self.prm = prm
        lib().orc_coupled_set_params(self._h, C.byref(prm))

    def close(self):
        if self._h:
            lib().orc_coupled_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sph2_init(n, prm: Params2) -> np.ndarray:
    p = np.zeros(n, PARTICLE2)
    lib().orc_sph2_init(_ptr(p), C.c_int(n), C.byref(prm))
    return p


def sph2_density(pin, prm, wave1d, g, counter, offset, index_list):
    out = np.zeros_like(pin); tx = make_tex(wave1d)
    lib().orc_sph2_density(_ptr(pin), _ptr(out), C.c_int(pin.size), C.byref(prm), C.byref(tx), C.byref(g),
                           _ptr(counter), _ptr(offset), _ptr(index_list))
    return out


def sph2_forces(pin, prm, wave1d, g, counter, offset, index_list):
    out = np.zeros_like(pin); tx = make_tex(wave1d)
    lib().orc_sph2_forces(_ptr(pin), _ptr(out), C.c_int(pin.size), C.byref(prm), C.byref(tx), C.byref(g),
                          _ptr(counter), _ptr(offset), _ptr(index_list))
    return out


def sph2_step(buf0, buf1, read_index, substeps, prm, wave1d, g):
    n = buf0.size
    ncell = g.ncells[0] * g.ncells[1]
    counter = np.zeros(ncell, np.int32); offset = np.zeros(ncell, np.int32)
    index_list = np.full(n, -1, np.int32); cell_of = np.zeros(n, np.int32)
    tx = make_tex(wave1d)
    r = lib().orc_sph2_step(_ptr(buf0), _ptr(buf1), C.c_int(read_index), C.c_int(n), C.c_int(substeps),
                            C.byref(prm), C.byref(tx), C.byref(g), _ptr(counter), _ptr(offset), _ptr(index_list),
                            _ptr(cell_of))
    return r, (cell_of, counter, offset, index_list)


# ------------------------------------------------------------------------------------------------
# reference CPU grid (oracle/_ref)
# ------------------------------------------------------------------------------------------------
def ref_grid2d_build(xy: np.ndarray, mn, mx, ncells):
    R = ref_lib()
    if R is None:
        return None
    xy = np.ascontiguousarray(xy, np.float32)
    n = xy.shape[0]
    ncell = ncells[0] * ncells[1]
    counter = np.empty(ncell, np.int32); offset = np.empty(ncell, np.int32); index_list = np.empty(n, np.int32)
    cs = np.empty(2, np.float32)
    rc = R.ref_grid2d_build(_ptr(xy), C.c_int(n), (C.c_float * 2)(*mn), (C.c_float * 2)(*mx), (C.c_int * 2)(*ncells),
                            _ptr(counter), _ptr(offset), _ptr(index_list), _ptr(cs))
    if rc != 0:
        raise RuntimeError("reference grid build failed")
    return counter, offset, index_list, cs


# ------------------------------------------------------------------------------------------------
# 1-D wave substrates + ImageStencil (SphWave2D/StencilImage2D.cpp:67-164)
# ------------------------------------------------------------------------------------------------
STENCIL1D_SHALLOW, STENCIL1D_WAVE = 0, 1
BC_REFLECT, BC_FREE, BC_FIXED = 0, 1, 2


def default_stencil1d_params(shader: int) -> Stencil1DParams:
    p = Stencil1DParams()
    lib().orc_stencil1d_params_default(C.byref(p), C.c_int(shader))
    return p


def shallow1d_dispatch(inp: np.ndarray, out: np.ndarray, mode: int, prm: Stencil1DParams) -> None:
    """One dispatch of Shallow1D_cs.glsl; `out` ([w,4] float32) is updated in place (unwritten texels keep their value)."""
    assert inp.dtype == np.float32 and out.dtype == np.float32 and inp.shape == out.shape and inp.shape[1] == 4
    lib().orc_shallow1d_dispatch(_ptr(np.ascontiguousarray(inp)), _ptr(out), inp.shape[0], mode, C.byref(prm))


def wave1d_dispatch(in0: np.ndarray, in1: np.ndarray, out: np.ndarray, mode: int, prm: Stencil1DParams) -> None:
    assert in0.shape == in1.shape == out.shape and out.shape[1] == 4
    lib().orc_wave1d_dispatch(_ptr(np.ascontiguousarray(in0)), _ptr(np.ascontiguousarray(in1)), _ptr(out), out.shape[0], mode, C.byref(prm))


class ImageStencil:
    """Host bookkeeping of ImageStencil restated line by line: N RGBA32F 1-D images, mReadIndex / mWriteIndex, per-image
    unit (SwapUnits), Reinit / Compute / ComputeFunc / ReinitFromTexture.  The shader addresses images by UNIT
    (Shallow1D: input 0, output 1; Wave1D: input0 0, input1 1, output 2), which is what the dispatch below does."""

    def __init__(self, shader: int, width: int, prm: Stencil1DParams | None = None):
        self.shader, self.w = shader, width
        self.prm = prm if prm is not None else default_stencil1d_params(shader)
        self.substeps = 1
        self.iterate = True
        if shader == STENCIL1D_SHALLOW:                       # InitShallowWaterEquation, SphWave2D/Main.cpp:77-97
            n, self.mode_iter_first, self.mode_iter_last = 2, 2, 3
        else:                                                 # InitWaveEquation, SphWave2D/Main.cpp:63-75
            n, self.mode_iter_first, self.mode_iter_last = 3, 2, 2
            self.substeps = 10
        self.mode_init_first = 0
        self.set_num_buffers(n)
        self.image = [np.zeros((width, 4), np.float32) for _ in range(self.num_images)]     # fresh texture storage: zeros
        self.unit = list(range(self.num_images))              # Init(): mImage[i].SetUnit(i)
        self.reinit()

    def set_num_buffers(self, n: int):                        # StencilImage2D.cpp:38-65
        self.num_images = max(1, n)
        if self.num_images == 1:
            self.read_index, self.write_index = [0], 0
        else:
            self.read_index, self.write_index = list(range(self.num_images - 1)), self.num_images - 1

    def pingpong(self):                                       # :67-83
        if self.num_images == 1:
            return
        r = self.read_index
        self.write_index, r[0] = r[0], self.write_index
        for i in range(len(r) - 1):
            r[i], r[i + 1] = r[i + 1], r[i]
        u = self.unit
        u[self.write_index], u[r[0]] = u[r[0]], u[self.write_index]
        for i in range(len(r) - 1):
            u[r[i]], u[r[i + 1]] = u[r[i + 1]], u[r[i]]

    def _with_unit(self, k: int) -> int:
        return self.unit.index(k)

    def dispatch(self, mode: int):
        if self.shader == STENCIL1D_SHALLOW:
            shallow1d_dispatch(self.image[self._with_unit(0)], self.image[self._with_unit(1)], mode, self.prm)
        else:
            wave1d_dispatch(self.image[self._with_unit(0)], self.image[self._with_unit(1)], self.image[self._with_unit(2)], mode, self.prm)

    def compute_func(self, mode: int):                        # :107-120
        self.dispatch(mode)
        self.pingpong()

    def reinit(self):                                         # :85-105
        for i in range(len(self.read_index)):
            self.compute_func(self.mode_init_first + i)

    def reinit_from_texture(self, rgba: np.ndarray):          # :122-140, mode -1: texelFetch(uInitImage, coord)
        out = self.image[self._with_unit(len(self.unit) - 1 if self.shader == STENCIL1D_WAVE else 1)]
        n = min(self.w, rgba.shape[0])
        out[:] = 0.0
        out[:n] = rgba[:n]
        self.pingpong()

    def compute(self, nframes: int = 1):                      # :142-164
        if not self.iterate:
            return
        for _ in range(nframes):
            for _ in range(self.substeps):
                for m in range(self.mode_iter_first, self.mode_iter_last + 1):
                    self.compute_func(m)

    def read_image(self, i: int) -> np.ndarray:               # GetReadImage(i)
        return self.image[self.read_index[i]]

    def write_image(self) -> np.ndarray:
        return self.image[self.write_index]
