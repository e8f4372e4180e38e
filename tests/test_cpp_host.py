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


def _build():
    from coupledwateranimation_b200 import build as B
    B.build()
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "coupled_main.cpp"),
           "-L" + PKG, "-lcwa_b200", "-Wl,-rpath," + PKG, "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_cpp_host_compiles_and_links_against_the_c_abi():
    exe = _build()
    assert os.path.exists(exe)
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
