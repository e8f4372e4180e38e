"""ctypes declarations of the C ABI in include/cwa_b200.h.

The library is the product; this module only loads it.  There is no fallback: if
libcwa_b200.so is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CWA_LIB_PATH") or os.path.join(HERE, "libcwa_b200.so")   # CWA_LIB_PATH: another build of the same library (tuning runs)


class CwaError(RuntimeError):
    pass


class GridInfo(C.Structure):
    _fields_ = [("min", C.c_float * 4), ("max", C.c_float * 4), ("num_cells", C.c_int * 4), ("cell_size", C.c_float * 4)]


class ConstantsUniform(C.Structure):
    _fields_ = [("mass", C.c_float), ("smoothing_coeff", C.c_float), ("visc", C.c_float), ("resting_rho", C.c_float)]


class BoundaryUniform(C.Structure):
    _fields_ = [("upper", C.c_float * 4), ("lower", C.c_float * 4)]


class WaveUniforms(C.Structure):
    _fields_ = [("attributes", C.c_float * 4), ("mesh_ws_pos", C.c_float * 4)]


class SimConstants(C.Structure):
    _fields_ = [("particle_radius", C.c_float), ("gas_const", C.c_float), ("dt", C.c_float), ("gravity_y", C.c_float),
                ("damping", C.c_float), ("crest_threshold", C.c_float), ("foam_speed", C.c_float), ("uv_scale", C.c_float),
                ("uv_scale_z", C.c_float), ("torque_coeff", C.c_float), ("pad1", C.c_float), ("pad2", C.c_float)]


class SlabDesc(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("wave_w", C.c_int), ("wave_h", C.c_int), ("wave_ch", C.c_int),
                ("row_lo", C.c_int), ("row_hi", C.c_int), ("store_lo", C.c_int), ("store_hi", C.c_int),
                ("left_store_hi", C.c_int), ("right_store_lo", C.c_int), ("halo_rows_max", C.c_int),
                ("z_lo", C.c_float), ("z_hi", C.c_float), ("band", C.c_float),
                ("cap_mig", C.c_int), ("cap_ghost", C.c_int), ("capacity", C.c_int), ("timeout_ms", C.c_int)]


# name -> (restype, argtypes); every symbol include/cwa_b200.h declares
_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_Z = C.c_size_t
_IP = C.POINTER(C.c_int)
SIGNATURES = {
    "cwa_last_error": (C.c_char_p, []),
    "cwa_version": (_I, []),
    "cwa_create": (_I, [_I, C.POINTER(_P)]),
    "cwa_destroy": (None, [_P]),
    "cwa_synchronize": (_I, [_P]),
    "cwa_stream": (_P, [_P]),
    "cwa_device_info": (_I, [_P, _IP, _IP, _IP, C.POINTER(_Z)]),
    "cwa_timer_begin": (_I, [_P]),
    "cwa_timer_end": (_I, [_P, C.POINTER(_F)]),
    "cwa_launch_count": (C.c_ulonglong, [_P]),
    "cwa_set_tuning": (_I, [_P, C.c_char_p, _I]),
    "cwa_profile_kernel_count": (_I, []),
    "cwa_profile_kernel_name": (C.c_char_p, [_I]),
    "cwa_profile_begin": (_I, [_P]),
    "cwa_profile_end": (_I, [_P, C.POINTER(_F), _IP, _I]),
    "cwa_buffer_create": (_I, [_P, _Z, _P, _IP]),
    "cwa_buffer_wrap": (_I, [_P, _P, _Z, _IP]),
    "cwa_buffer_destroy": (_I, [_P, _I]),
    "cwa_buffer_sub_data": (_I, [_P, _I, _Z, _Z, _P]),
    "cwa_buffer_read": (_I, [_P, _I, _Z, _Z, _P]),
    "cwa_buffer_read_async": (_I, [_P, _I, _Z, _Z, _P]),
    "cwa_buffer_copy": (_I, [_P, _I, _I, _Z, _Z, _Z]),
    "cwa_buffer_bind_base": (_I, [_P, _I, _I, _I]),
    "cwa_buffer_device_ptr": (_I, [_P, _I, C.POINTER(_P), C.POINTER(_Z)]),
    "cwa_default_ubo": (_I, [_P, _I, _IP]),
    "cwa_scan_exclusive": (_I, [_P, _I, _I, _I]),
    "cwa_grid_create": (_I, [_P, _I, _P, _P, _P, _I, _IP]),
    "cwa_grid_destroy": (_I, [_P, _I]),
    "cwa_grid_get_info": (_I, [_P, _I, C.POINTER(GridInfo), _IP]),
    "cwa_grid_build": (_I, [_P, _I, _I, _I, _I]),
    "cwa_grid_read": (_I, [_P, _I, _I, _P, _I]),
    "cwa_grid_buffer": (_I, [_P, _I, _I, _IP]),
    "cwa_wave_create": (_I, [_P, _I, _I, _I, _I, _IP]),
    "cwa_wave_destroy": (_I, [_P, _I]),
    "cwa_wave_reinit": (_I, [_P, _I]),
    "cwa_wave_reinit_from_texture": (_I, [_P, _I, _P, _I, _I]),
    "cwa_wave_compute": (_I, [_P, _I, _I]),
    "cwa_wave_pingpong": (_I, [_P, _I]),
    "cwa_wave_set_evolve": (_I, [_P, _I, _I]),
    "cwa_wave_set_params": (_I, [_P, _I, _F, _F, _F]),
    "cwa_wave_resize": (_I, [_P, _I, _I, _I]),
    "cwa_wave_state": (_I, [_P, _I, _IP, _IP, _IP, _IP]),
    "cwa_wave_bind_texture_unit": (_I, [_P, _I]),
    "cwa_wave_read_image": (_I, [_P, _I, _I, _P]),
    "cwa_wave_write_image": (_I, [_P, _I, _I, _P]),
    "cwa_wave_mark_written": (_I, [_P, _I, _I]),
    "cwa_wave_read_image_async": (_I, [_P, _I, _I, _P]),
    "cwa_wave_role_image": (_I, [_P, _I, _I, _IP]),
    "cwa_wave_image_buffer": (_I, [_P, _I, _I, _IP]),
    "cwa_wave_size": (_I, [_P, _I, _IP, _IP, _IP]),
    "cwa_sph_create": (_I, [_P, _I, _I, _I, _IP]),
    "cwa_sph_destroy": (_I, [_P, _I]),
    "cwa_sph_bind_wave": (_I, [_P, _I, _I, _I]),
    "cwa_sph_rho_pres": (_I, [_P, _I]),
    "cwa_sph_force": (_I, [_P, _I]),
    "cwa_sph_integrate": (_I, [_P, _I]),
    "cwa_sph_step": (_I, [_P, _I, _I]),
    "cwa_sph_neighbour_count": (_I, [_P, _I, _P]),
    "cwa_sph_init_cube": (_I, [_P, _I, _I, _I, _I]),
    "cwa_coupled_step": (_I, [_P, _I, _I, _I, _I]),
    "cwa_bind_scene": (_I, [_P, _I, _I]),
    "sph_step": (_I, [_P, _I]),
    "wave_step": (_I, [_P, _I]),
    "cwa_sph_set_count": (_I, [_P, _I, _I]),
    "cwa_particles_copy_if": (_I, [_P, _I, _I, _I, _I, _F, _F, _I, _I, _IP]),
    "cwa_stencil1d_create": (_I, [_P, _I, _I, _P]),
    "cwa_stencil1d_destroy": (_I, [_P, _I]),
    "cwa_stencil1d_reinit": (_I, [_P, _I]),
    "cwa_stencil1d_reinit_from_texture": (_I, [_P, _I, _P, _I]),
    "cwa_stencil1d_compute": (_I, [_P, _I, _I]),
    "cwa_stencil1d_compute_func": (_I, [_P, _I, _I]),
    "cwa_stencil1d_set_params": (_I, [_P, _I, _F, _F, _F, _F, _F, _I]),
    "cwa_stencil1d_pingpong": (_I, [_P, _I]),
    "cwa_stencil1d_set_substeps": (_I, [_P, _I, _I]),
    "cwa_stencil1d_set_iterate": (_I, [_P, _I, _I]),
    "cwa_stencil1d_state": (_I, [_P, _I, _P, _P, _P, _P]),
    "cwa_stencil1d_image_buffer": (_I, [_P, _I, _I, _P]),
    "cwa_stencil1d_read_image": (_I, [_P, _I, _I, _P]),
    "cwa_stencil1d_write_image": (_I, [_P, _I, _I, _P]),
    "cwa_param_count": (_I, []),
    "cwa_param_info": (_I, [_I, _P, _P, _P, _P]),
    "cwa_param_set": (_I, [_P, C.c_char_p, _F]),
    "cwa_param_get": (_I, [_P, C.c_char_p, _P]),
    "cwa_checkpoint_save": (_I, [_P, _I, _I, C.c_ulonglong, C.c_char_p]),
    "cwa_checkpoint_load": (_I, [_P, _I, _I, C.c_char_p, _P]),
    "cwa_slab_pack": (_I, [_P, _I, _I, _F, _F, _F, _I, _I, _I, _I]),
    "cwa_sph_step_slab": (_I, [_P, _I, _I, _F, _F, _F, _I, _I, _I, _I]),
    "cwa_slab_unpack": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _IP]),
    "cwa_slab_compact": (_I, [_P, _I, _I, _I, _IP]),
    "cwa_slab_plan": (_I, [_I, _I, _I, _I, _I, C.c_double, C.c_double, _IP, C.POINTER(SlabDesc)]),
    "cwa_slab_create": (_I, [_P, C.POINTER(SlabDesc), _I, _I, _IP]),
    "cwa_slab_destroy": (_I, [_P, _I]),
    "cwa_slab_export": (_I, [_P, _I, _P]),
    "cwa_slab_mailbox": (_I, [_P, _I, C.POINTER(_P), C.POINTER(_Z)]),
    "cwa_slab_connect": (_I, [_P, _I, _I, _P, _P]),
    "cwa_slab_set_owned": (_I, [_P, _I, _I]),
    "cwa_slab_step": (_I, [_P, _I, _I, _I]),
    "cwa_slab_group_step": (_I, [C.POINTER(_P), _IP, _I, _I, _I]),
    "cwa_slab_counts": (_I, [_P, _I, _IP]),
    "cwa_slab_counts_async": (_I, [_P, _I, _P]),
    "cwa_gl_available": (_I, []),
    "cwa_gl_register_buffer": (_I, [_P, C.c_uint, _IP]),
    "cwa_gl_map_buffer": (_I, [_P, _I, _IP]),
    "cwa_gl_unmap": (_I, [_P, _I]),
    "cwa_gl_register_image": (_I, [_P, C.c_uint, C.c_uint, _IP]),
    "cwa_gl_copy_wave_to_image": (_I, [_P, _I, _I, _I]),
    "cwa_gl_unregister": (_I, [_P, _I]),
    "cwa_wave_reinit_from_rgba8": (_I, [_P, _I, _P, _I, _I]),
    "cwa_wave_create_block": (_I, [_P, _I, _I, _I, _I, _I, _I, _IP]),
    "cwa_wave_last_row_buffer": (_I, [_P, _I, _I, _IP]),
    "cwa_sph2_create": (_I, [_P, _I, _I, _I, _IP]),
    "cwa_sph2_destroy": (_I, [_P, _I]),
    "cwa_sph2_reinit": (_I, [_P, _I]),
    "cwa_sph2_set_substeps": (_I, [_P, _I, _I]),
    "cwa_sph2_set_uniforms": (_I, [_P, _I, _F, _F, _F, _I]),
    "cwa_sph2_set_view_width": (_I, [_P, _I, _F]),
    "cwa_sph2_bind_wave1d": (_I, [_P, _I, _I, _I]),
    "cwa_sph2_compute": (_I, [_P, _I, _I]),
    "cwa_sph2_read": (_I, [_P, _I, _P]),
    "cwa_sph2_write": (_I, [_P, _I, _P]),
    "cwa_sph2_read_buffer": (_I, [_P, _I, _IP]),
    "cwa_shader_create": (_I, [_P, C.c_char_p, _IP]),
    "cwa_shader_set_mode": (_I, [_P, _I, _I]),
    "cwa_shader_set_uniform_i": (_I, [_P, _I, _I, _I]),
    "cwa_shader_set_uniform_f": (_I, [_P, _I, _I, _F]),
    "cwa_shader_bind_object": (_I, [_P, _I, _I]),
    "cwa_shader_dispatch": (_I, [_P, _I, _I, _I, _I]),
}

_lib = None


def load():
    """dlopen the in-tree library; no fallback of any kind."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CwaError(
            f"{LIB_PATH} is missing: build it with `python -m coupledwateranimation_b200.build` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the simulation step.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().cwa_last_error()
        raise CwaError(f"libcwa_b200 error {rc}: {msg.decode() if msg else '?'}")
