"""The C++ host side: the mirrored reference classes (include/cwa/*.h) compile, link against the
C-ABI library and -- on the GPU box -- drive the as-shipped scene exactly like Main.cpp's idle()."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "coupledwateranimation_b200")
EXE = os.path.join(ROOT, "examples", "coupled_main")
EXE2D = os.path.join(ROOT, "examples", "sphwave2d_main")


def _build(name="coupled_main"):
    from coupledwateranimation_b200 import build as B
    B.build()
    exe = os.path.join(ROOT, "examples", name)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", name + ".cpp"),
           "-L" + PKG, "-lcwa_b200", "-Wl,-rpath," + PKG, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_compiles_and_links_against_the_c_abi():
    exe = _build()
    assert os.path.exists(exe)
    assert os.path.exists(_build("sphwave2d_main"))       # SphUgrid + ImageStencil: the 2-D app's simulation half
    assert os.path.exists(_build("slab_main"))            # N ranks of the slab-decomposed frame through cwa_slab_* (plain C ABI)
    # C (not C++) consumers must be able to include the ABI header too
    r = subprocess.run(["gcc", "-std=c11", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "cwa_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_cpp_host_runs_the_shipped_scene_like_main_cpp(oracle):
    exe = _build()
    r = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"frames=2 particles=20480 nan=(\d+) mean_rho=([\d.]+) mean_y=([\d.eE+-]+)", r.stdout)
    assert m, r.stdout
    prm = oracle.default_params3()
    oc = oracle.Coupled(20480, 64, 64, 4, prm, oracle.COUPLING_AS_SHIPPED)
    oc.particles[:] = oracle.make_cube(64, 5, 64, prm)
    oc.step(2)
    P = oc.particles
    ok = ~np.isnan(P["pos"][:, 1])
    assert int(m.group(1)) == int((~ok).sum())
    assert abs(float(m.group(2)) - float(P["extras"][ok, 0].astype(np.float64).mean())) <= 1e-3 * float(P["extras"][ok, 0].mean())
    assert abs(float(m.group(3)) - float(P["pos"][ok, 1].astype(np.float64).mean())) <= 1e-6
    oc.close()


@pytest.mark.gpu
def test_cpp_host_runs_the_2d_app_like_sphwave2d_main_cpp(oracle):
    """examples/sphwave2d_main.cpp: SphUgrid (2 substeps) + ImageStencil (Shallow1D, modes 2,3) + the sampler re-bind, per frame."""
    exe = _build("sphwave2d_main")
    frames = 3
    r = subprocess.run([exe, str(frames)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"frames=3 particles=4096 enabled=(\d+) mean_x=([\d.eE+-]+) mean_y=([\d.eE+-]+) mean_rho=([\d.eE+-]+) wave_h_sum=([\d.eE+-]+)", r.stdout)
    assert m, r.stdout
    n = 4096
    prm = oracle.default_params2(oracle.SPH2_WAVE)
    b0, b1 = oracle.sph2_init(n, prm), np.zeros(n, oracle.PARTICLE2)
    g = oracle.grid2((0.0, 0.0), (9.6, 9.6), (32, 32))
    wave = oracle.ImageStencil(oracle.STENCIL1D_SHALLOW, 128)
    ri, tex = 0, None                                    # frame 1 samples the unbound texture: the bind follows the first compute
    for _ in range(frames):
        ri, _g = oracle.sph2_step(b0, b1, ri, 2, prm, tex, g)
        wave.compute(1)
        tex = wave.read_image(0).copy().reshape(1, 128, 4)
    P = (b0, b1)[ri]
    assert int(m.group(1)) == int((P["pos"][:, 3] > 0.5).sum())
    for got, ref in ((m.group(2), P["pos"][:, 0]), (m.group(3), P["pos"][:, 1]), (m.group(4), P["acc"][:, 3])):
        ref = float(ref.astype(np.float64).mean())
        assert abs(float(got) - ref) <= 1e-5 * abs(ref) + 1e-6, (got, ref)
    assert abs(float(m.group(5)) - float(wave.read_image(0)[:, 0].astype(np.float64).sum())) <= 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [2, 3])
def test_cpp_host_drives_several_ranks_through_the_slab_abi(ranks):
    """examples/slab_main.cpp: a C++ host (no Python, no torch) plans the slabs, creates one context per rank, wires the mailboxes and steps
    the ranks with cwa_slab_group_step; then runs the same scene with cwa_coupled_step and compares.  On one GPU the ranks share it."""
    exe = _build("slab_main")
    r = subprocess.run([exe, str(ranks), "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"ranks=(\d+) frames=8 particles=(\d+) ids_ok=1 nan_diff=0 migrated=(\d+) wave_bit_exact=1 max_pos_diff=([\d.eE+-]+) max_vel_rel=([\d.eE+-]+)", r.stdout)
    assert m, r.stdout
    assert int(m.group(1)) == ranks and int(m.group(2)) == 64 * 5 * 160
    assert int(m.group(3)) > 0, "the fixture must move particles across slab faces"
    assert float(m.group(4)) <= 1e-5 and float(m.group(5)) <= 2e-3, r.stdout
