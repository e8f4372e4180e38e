"""CPU: the workload definitions of bench.py (no GPU, no library): C4, the weak-scaling scenes and the slab plans built on them."""
import json
import math
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_c4_is_the_survey_config():
    sc = bench.scaled_scene(1)
    assert sc["nx"] * sc["ny"] * sc["nz"] == 1003520 == bench.N_PARTICLES
    assert (sc["wave_w"], sc["wave_h"]) == (2048, 2048) and sc["torque"] == 0.0 and sc["uv_z"] == 0.0
    assert abs(sc["uv"] - 2.0 / 7.0) < 1e-12
    cell = [(sc["gmax"][a] - sc["gmin"][a]) / sc["gn"][a] for a in range(3)]
    assert all(0.01 * 1.0025 <= c <= 0.0125 for c in cell), "cells of ~h: the 27-cell query is exact (DESIGN 5, list_query)"
    assert "1003520 particles" in bench.WORKLOAD and "2048^2" in bench.WORKLOAD


def test_weak_scaling_scene_keeps_the_per_gpu_share_and_the_torque_gain():
    base = bench.scaled_scene(1)
    per_gpu = base["nx"] * base["ny"] * base["nz"]
    for world in (2, 4, 8):
        sc = bench.scaled_scene(world)
        n = sc["nx"] * sc["ny"] * sc["nz"]
        assert abs(n / world - per_gpu) <= 0.005 * per_gpu
        assert abs(sc["wave_w"] * sc["wave_h"] / world - 2048 * 2048) <= 0.005 * 2048 * 2048
        cells = sc["gn"][0] * sc["gn"][1] * sc["gn"][2]
        assert abs(cells / world - 384 * 31 * 384) <= 0.01 * 384 * 31 * 384
        # lattice pitch, texel size and cell size are C4's
        assert abs(sc["box_x"] / sc["gn"][0] - base["box_x"] / base["gn"][0]) < 5e-6          # 384 sqrt(N) cells, rounded to a whole number
        assert abs(1.0 / (sc["uv"] * sc["wave_w"]) - 1.0 / (base["uv"] * base["wave_w"])) < 2e-6
        # far-corner feedback gain of the torque term, 0.25 |pos| at C4: unchanged
        gain = sc["torque"] * math.hypot(sc["box_x"], sc["box_z"])
        assert abs(gain - 0.25 * math.hypot(base["box_x"], base["box_z"])) < 1e-9
        assert f"sqrt({world})" in bench.scaled_workload(world)


def test_weak_scaling_slabs_are_balanced_and_valid():
    from coupledwateranimation_b200.distributed import SlabPlan
    for world in (2, 4, 8):
        sc = bench.scaled_scene(world)
        sp = np.float32(np.float32(2.0 * 0.85) * np.float32(0.005))
        ks = np.arange(sc["nz"], dtype=np.float32) * sp
        rb = SlabPlan.balanced_row_bounds(world, sc["wave_h"], sc["uv"], ks - np.float32(0.5) * sp)
        plans = [SlabPlan.make(world, r, sc["wave_w"], sc["wave_h"], sc["uv"], 0.01, rb) for r in range(world)]
        layers = [int(((ks >= p.z_lo) & (ks < p.z_hi)).sum()) for p in plans]
        assert sum(layers) == sc["nz"] and max(layers) - min(layers) <= 1
        for p in plans:
            p.validate()


def test_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` runs the CPU restatement here too (2 frames of C4 take a few seconds on 8 threads)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "particle_updates_per_sec" and line["unit"] == "particle-updates/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["config"]["workload"] == bench.WORKLOAD
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_c5_is_baseline_config_5():
    for world in (2, 4, 8):
        sc = bench.c5_scene(world, "torque_scaled")
        assert sc["nx"] * sc["ny"] * sc["nz"] == 16056320 and (sc["wave_w"], sc["wave_h"]) == (8192, 8192)
        assert sc["scaling"] == "strong" and "16056320 particles" in sc["name"]
        base = bench.scaled_scene(1)
        assert abs(sc["box_x"] / sc["gn"][0] - base["box_x"] / base["gn"][0]) < 1e-9       # C4's cell size
        assert abs(1.0 / (sc["uv"] * sc["wave_w"]) - 1.0 / (base["uv"] * base["wave_w"])) < 1e-9   # C4's texel size
        gain = sc["torque"] * math.hypot(sc["box_x"], sc["box_z"])
        assert abs(gain - 0.25 * math.hypot(base["box_x"], base["box_z"])) < 1e-9
        assert bench.c5_scene(world, "literal")["torque"] == 0.0                            # 0 = the shader's literal 0.25


def test_reference_arm_of_the_other_configs():
    """`bench.py --impl reference --config D|C2|C3`: the CPU restatement of configs 1-3 with the contract keys and the workload text of the
    native arm (bench_configs.py); bounded to a few steps here."""
    import bench_configs
    from oracle import oracle as O
    O.build()
    for cfg, unit in (("D", "particle-updates/s"), ("C2", "particle-updates/s"), ("C3", "cell-updates/s")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", cfg, "--steps", "2", "--warmup", "0"],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        line = json.loads(r.stdout.strip().splitlines()[-1])
        assert line["impl"] == "reference" and line["unit"].startswith(unit) and line["value"] > 0
        assert line["config"]["workload"].startswith(cfg + ":") and line["cpu_baseline"]["kind"] == "port"
        assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert set(bench_configs.RUNNERS) == {"D", "C2", "C3"}


def test_bench_cli_names_every_baseline_config():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    for word in ("--config", "C4", "C2", "C3", "--state-frame", "--scene", "--impl"):
        assert word in r.stdout
