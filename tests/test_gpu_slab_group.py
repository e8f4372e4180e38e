"""The slab-decomposed coupled frame through the C ABI (csrc/slab.cu) with all ranks inside ONE process: the whole device-side
protocol -- pack fused into integrate, peer stores into the neighbour's mailbox, flag waits, free-list adoption of migrants,
device-resident counts, wave halo / last-row exchange on the side stream -- runs on a single GPU (ranks = contexts sharing the
device) and is compared with the single-GPU path.  With more devices the ranks are spread over them (NVLink peer stores)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _check(res, world):
    assert res["count_conserved"], res
    assert res["wave_bit_exact"], res
    assert res["nan_sets_equal"], res
    if world > 1:
        assert res["migrated"] > 0, "the fixture must move particles across slab faces"
    assert res["ok"], res


def test_one_rank_slab_equals_plain_path_bit_for_bit():
    import dist_check
    res = dist_check.run_group(1, frames=8, coupling=0, calls=2)
    _check(res, 1)
    assert res["pos_max_rel"] == 0.0 and res["vel_max_rel"] == 0.0, res


@pytest.mark.parametrize("world,coupling", [(2, 0), (2, 1), (3, 0), (4, 1)])
def test_ranks_in_one_process_reproduce_one_gpu(world, coupling):
    import dist_check
    res = dist_check.run_group(world, frames=12, coupling=coupling, calls=3)
    _check(res, world)


def test_free_slots_are_reused():
    """Migrants are adopted into the slots earlier emigrants left: the owned range must not grow by the traffic."""
    import dist_check
    res = dist_check.run_group(2, frames=24, coupling=0, calls=1)
    _check(res, 2)
    sc, p = dist_check.scene("small")
    assert res["migrated"] > 300
    # without reuse every adopted migrant would be appended: owned ranges would add up to particles + migrants
    assert sum(res["owned_range"]) < p.size + res["migrated"] // 2, res


def test_count_ahead_across_the_exchange_changes_nothing(monkeypatch):
    """Frames in the middle of a call skip the cell hash: the integrate pass counts the particles that stay, the next frame's unpack
    kernel the arrivals (csrc/slab.cu).  Same cells, same canonical order inside a cell -> the same bits as the plain grid build."""
    import dist_check
    out = {}
    for ahead in ("0", "1"):
        monkeypatch.setenv("CWA_SLAB_AHEAD", ahead)
        out[ahead] = dist_check.run_group(3, frames=12, coupling=0, calls=2)
        _check(out[ahead], 3)
    for key in ("pos_max_rel", "vel_max_rel", "migrated", "owned_range", "free_slots"):
        assert out["0"][key] == out["1"][key], (key, out["0"][key], out["1"][key])
