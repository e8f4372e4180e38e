"""Multi-GPU parity on real devices: runs tools/dist_check.py under torchrun when the box has >= 2 GPUs
(the protocol itself is covered on CPU by tests/test_distributed_cpu.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_single_rank_driver_matches_plain_path():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dist_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"] and res["pos_max_rel"] == 0.0 and res["wave_bit_exact"]


@pytest.mark.parametrize("coupling", [0, 1])
def test_two_gpus_reproduce_one_gpu(coupling):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (validated with `gpurun --gpus 2`, see profiles/r1_multigpu_parity.md)")
    env = dict(os.environ, CWA_DIST_COUPLING=str(coupling))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"] and res["migrated"] > 0 and res["count_conserved"] and res["wave_bit_exact"]
