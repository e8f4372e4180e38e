"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank simulates its z slab of a small scene through the CUDA library
(coupledwateranimation_b200.distributed); rank 0 additionally runs the whole scene on its own GPU
through the single-GPU path and compares the gathered states particle by particle (ids ride in the
unused extras.w).  The wave field must be bit-identical."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupledwateranimation_b200 as cwa  # noqa: E402
from coupledwateranimation_b200.distributed import CudaBackend, DistributedCoupled, SlabPlan  # noqa: E402

NX, NY, NZ = 96, 5, 192
WAVE_W, WAVE_H = 256, 512
UV = 0.6
BOX = (0.9, 1.0, 1.7)
FRAMES = int(os.environ.get("CWA_DIST_FRAMES", "12"))


def scene():
    sp = np.float32(0.0085)
    i, j, k = np.meshgrid(np.arange(NX), np.arange(NY), np.arange(NZ), indexing="ij")
    p = np.zeros(NX * NY * NZ, cwa.PARTICLE)
    p["pos"][:, 0] = (i.ravel().astype(np.float32) * sp)
    p["pos"][:, 1] = (j.ravel().astype(np.float32) * sp)
    p["pos"][:, 2] = (k.ravel().astype(np.float32) * sp)
    p["pos"][:, 3] = 1.0
    rng = np.random.default_rng(5)
    p["pos"][:, :3] += rng.uniform(-0.1, 0.1, (p.size, 3)).astype(np.float32) * sp
    p["vel"][:, 2] = rng.uniform(-30.0, 30.0, p.size).astype(np.float32)      # 12 frames * 5e-5 * 30 = 0.018: crosses the faces
    p["extras"][:, 0] = 1000.0
    p["extras"][:, 2] = 500.0
    p["extras"][:, 3] = np.arange(p.size, dtype=np.float32)                   # id
    return p


def set_params(ctx):
    ctx.set_boundary(upper=(BOX[0], BOX[1], BOX[2], 500.0), lower=(0.0, -0.02, 0.0, 50.0))
    ctx.set_sim_constants(uv_scale=UV)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    coupling = int(os.environ.get("CWA_DIST_COUPLING", "0"))
    p = scene()
    ctx = cwa.Context(local_rank)
    set_params(ctx)
    plan = SlabPlan.make(world, rank, WAVE_W, WAVE_H, UV, 0.01)
    z = p["pos"][:, 2]
    mine = p[(z >= plan.z_lo) & (z < plan.z_hi)]
    zl = max(0.0, plan.z_lo - 0.06) if rank > 0 else 0.0
    zh = min(BOX[2], plan.z_hi + 0.06) if rank < world - 1 else BOX[2]
    ncz = max(4, int(np.ceil((zh - zl) / 0.02)))
    be = CudaBackend(cwa, ctx, plan, int(mine.size * 1.5) + 40000, (0.0, -0.02, zl), (BOX[0], 1.0, zh), (45, 51, ncz),
                     cap_mig=2048, cap_ghost=8192)
    be.upload_owned(mine)
    drv = DistributedCoupled(be, plan, dist if world > 1 else None)
    drv.init_wave_halos()
    drv.step(FRAMES, coupling)
    ctx.synchronize()
    owned = be.download_owned()
    img = be.newest_image()
    wave_rows = be.wave.read_image(img)[plan.row_lo - plan.store_lo:plan.row_hi - plan.store_lo]
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, (owned, wave_rows, be.migrated_in))
    else:
        gathered = [(owned, wave_rows, be.migrated_in)]
    ok = True
    if rank == 0:
        P = np.concatenate([g[0] for g in gathered])
        P = P[np.argsort(P["extras"][:, 3])]
        W = np.concatenate([g[1] for g in gathered])
        moved = sum(g[2] for g in gathered)
        # single-GPU reference on this rank's device
        grid = cwa.UniformGrid(ctx, 3, (0.0, -0.02, 0.0), (BOX[0], 1.0, BOX[2]), (45, 51, 85), p.size, compact_index=True)
        sph = cwa.Sph(ctx, p.size, grid, particles=p)
        wave = cwa.StencilImage2DTripleBuffered(ctx, WAVE_W, WAVE_H, 1, cwa.WAVE_COUPLED)
        sph.coupled_step(wave, FRAMES, coupling)
        R = sph.download()
        RW = wave.read_role(0)
        res = {"world": world, "frames": FRAMES, "coupling": coupling, "particles": int(P.size), "migrated": int(moved)}
        res["count_conserved"] = bool(P.size == R.size and np.array_equal(P["extras"][:, 3], R["extras"][:, 3]))
        res["wave_bit_exact"] = bool(np.array_equal(W.view(np.uint32), RW.view(np.uint32)))
        nan_r, nan_g = np.isnan(R["pos"]).any(1), np.isnan(P["pos"]).any(1)
        res["nan_sets_equal"] = bool(np.array_equal(nan_r, nan_g))
        good = ~nan_r & ~nan_g
        for f, tol in (("pos", 1e-4), ("vel", 1e-3)):
            a, b = P[f][good, :3].astype(np.float64), R[f][good, :3].astype(np.float64)
            scale = float(np.sqrt(np.mean(b ** 2)))
            err = np.abs(a - b).max(1) / scale
            res[f + "_max_rel"] = float(err.max())
            res[f + "_outliers"] = float((err > tol).mean())
        ok = res["count_conserved"] and res["wave_bit_exact"] and res["nan_sets_equal"] and res["pos_outliers"] <= 0.005 and res["vel_outliers"] <= 0.005
        res["ok"] = bool(ok)
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
