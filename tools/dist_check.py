"""Multi-GPU parity check of the slab-decomposed coupled frame (csrc/slab.cu).

  one process per GPU (CUDA IPC mailboxes):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
  all ranks in ONE process (direct peer pointers; ranks share devices round-robin when there are fewer devices than ranks):
    python tools/dist_check.py --group 3

Every rank simulates its z slab of a scene through the C-ABI slab object; the gathered state is compared particle by particle
(ids ride in the unused extras.w) with the whole scene run through the single-GPU path (cwa_coupled_step).  The wave field must be
bit-identical, every particle must exist exactly once, the NaN sets must agree.

Environment: CWA_DIST_FRAMES (12), CWA_DIST_COUPLING (0 as shipped | 1 latest), CWA_DIST_SCENE (small | big: 1.0 M particles),
CWA_DIST_CALLS (frames are spread over this many step calls, default 2: covers the explicit pack at a call's start)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupledwateranimation_b200 as cwa  # noqa: E402
from coupledwateranimation_b200.distributed import SlabRank, connect_local, connect_processes, group_step, plan_desc  # noqa: E402

FRAMES = int(os.environ.get("CWA_DIST_FRAMES", "12"))
CALLS = int(os.environ.get("CWA_DIST_CALLS", "2"))
SCENES = {
    # name: lattice, wave, uv, box, grid cells of the whole-scene reference (x, y, z), velocity kick
    "small": dict(n=(96, 5, 192), wave=(256, 512), uv=0.6, box=(0.9, 1.0, 1.7), cells=(45, 51, 85), cell=0.02, kick=30.0, cap=(4096, 16384)),
    "big": dict(n=(448, 5, 448), wave=(2048, 2048), uv=2.0 / 7.0, box=(3.85, 1.0, 3.85), cells=(384, 31, 384), cell=3.85 / 384, kick=30.0,
                cap=(32768, 65536), gy=0.30),
}


def scene(name="small"):
    sc = SCENES[name]
    nx, ny, nz = sc["n"]
    sp = np.float32(0.0085)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    p = np.zeros(nx * ny * nz, cwa.PARTICLE)
    p["pos"][:, 0] = (i.ravel().astype(np.float32) * sp)
    p["pos"][:, 1] = (j.ravel().astype(np.float32) * sp)
    p["pos"][:, 2] = (k.ravel().astype(np.float32) * sp)
    p["pos"][:, 3] = 1.0
    rng = np.random.default_rng(5)
    p["pos"][:, :3] += rng.uniform(-0.1, 0.1, (p.size, 3)).astype(np.float32) * sp
    p["vel"][:, 2] = rng.uniform(-sc["kick"], sc["kick"], p.size).astype(np.float32)      # 12 frames * 5e-5 * 30 = 0.018: crosses the faces
    p["extras"][:, 0] = 1000.0
    p["extras"][:, 2] = 500.0
    p["extras"][:, 3] = np.arange(p.size, dtype=np.float32)                   # id
    return sc, p


def set_params(ctx, sc):
    ctx.set_boundary(upper=(sc["box"][0], sc["box"][1], sc["box"][2], 500.0), lower=(0.0, -0.02, 0.0, 50.0))
    ctx.set_sim_constants(uv_scale=sc["uv"])


def make_rank(ctx, sc, p, world, rank, timeout_ms=10000):
    """This rank's share of the scene: SlabRank with its owned particles uploaded."""
    set_params(ctx, sc)
    ww, wh = sc["wave"]
    d = plan_desc(ctx.lib, world, rank, ww, wh, 1, sc["uv"], 0.01, None, cap_mig=sc["cap"][0], cap_ghost=sc["cap"][1], timeout_ms=timeout_ms)
    z = p["pos"][:, 2]
    mine = p[(z >= d.z_lo) & (z < d.z_hi)]
    d.capacity = int(mine.size * 1.5) + 2 * (2 * sc["cap"][0] + sc["cap"][1]) + 1024
    box = sc["box"]
    zl = max(0.0, d.z_lo - 0.06) if rank > 0 else 0.0
    zh = min(box[2], d.z_hi + 0.06) if rank < world - 1 else box[2]
    ncz = max(4, int(np.floor((zh - zl) / sc["cell"] + 1e-6)))   # cells never narrower than the scene's cell (the 3x3x3 query needs cell >= 1.0025 h)
    gy = sc.get("gy", 1.0)
    rk = SlabRank(cwa, ctx, d, (0.0, -0.02, zl), (box[0], gy, zh), (sc["cells"][0], sc["cells"][1], ncz))
    rk.upload_owned(mine)
    return rk


def reference(ctx, sc, p, frames, coupling):
    """The whole scene through the single-GPU path on ctx's device."""
    set_params(ctx, sc)
    box = sc["box"]
    grid = cwa.UniformGrid(ctx, 3, (0.0, -0.02, 0.0), (box[0], sc.get("gy", 1.0), box[2]), sc["cells"], p.size, compact_index=True)
    sph = cwa.Sph(ctx, p.size, grid, particles=p)
    wave = cwa.StencilImage2DTripleBuffered(ctx, sc["wave"][0], sc["wave"][1], 1, cwa.WAVE_COUPLED)
    sph.coupled_step(wave, frames, coupling)
    return sph.download(), wave.read_role(0)


def compare(P, W, R, RW, moved, extra):
    P = P[np.argsort(P["extras"][:, 3])]
    res = dict(extra)
    res.update({"particles": int(P.size), "migrated": int(moved)})
    res["count_conserved"] = bool(P.size == R.size and np.array_equal(P["extras"][:, 3], R["extras"][:, 3]))
    res["wave_bit_exact"] = bool(W.shape == RW.shape and np.array_equal(W.view(np.uint32), RW.view(np.uint32)))
    ok = res["count_conserved"] and res["wave_bit_exact"]
    if res["count_conserved"]:
        nan_r, nan_g = np.isnan(R["pos"]).any(1), np.isnan(P["pos"]).any(1)
        res["nan_sets_equal"] = bool(np.array_equal(nan_r, nan_g))
        good = ~nan_r & ~nan_g
        for f, tol in (("pos", 1e-4), ("vel", 1e-3)):
            a, b = P[f][good, :3].astype(np.float64), R[f][good, :3].astype(np.float64)
            scale = float(np.sqrt(np.mean(b ** 2)))
            err = np.abs(a - b).max(1) / scale
            res[f + "_max_rel"] = float(err.max())
            res[f + "_outliers"] = float((err > tol).mean())
        ok = ok and res["nan_sets_equal"] and res["pos_outliers"] <= 0.005 and res["vel_outliers"] <= 0.005
    res["ok"] = bool(ok)
    return res


def split_frames(frames, calls):
    calls = max(1, min(calls, frames))
    base, rem = divmod(frames, calls)
    return [base + (1 if c < rem else 0) for c in range(calls)]


def run_group(world, frames=FRAMES, coupling=0, scene_name="small", devices=None, calls=CALLS):
    """All ranks in this process (one host thread): rank r on device devices[r % len(devices)]."""
    import torch
    ndev = torch.cuda.device_count()
    devices = devices if devices is not None else list(range(min(world, ndev)))
    sc, p = scene(scene_name)
    ctxs = [cwa.Context(devices[r % len(devices)]) for r in range(world)]
    ranks = [make_rank(ctxs[r], sc, p, world, r) for r in range(world)]
    connect_local(ranks)
    for n in split_frames(frames, calls):
        group_step(ranks, n, coupling)
    for c in ctxs:
        c.synchronize()
    cnt = [rk.check() for rk in ranks]
    P = np.concatenate([rk.download_owned() for rk in ranks])
    W = np.concatenate([rk.owned_wave_rows(0) for rk in ranks])
    moved = sum(c["migrated_in"] for c in cnt)
    ref_ctx = cwa.Context(devices[0])
    R, RW = reference(ref_ctx, sc, p, frames, coupling)
    res = compare(P, W, R, RW, moved, {"mode": "group", "world": world, "devices": devices, "frames": frames, "coupling": coupling, "scene": scene_name,
                                       "free_slots": [c["free"] for c in cnt], "owned_range": [c["n_owned"] for c in cnt]})
    ref_ctx.close()
    for c in ctxs:
        c.close()
    return res


def main_processes():
    import torch
    import torch.distributed as dist
    from datetime import timedelta
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=timedelta(seconds=120))
    coupling = int(os.environ.get("CWA_DIST_COUPLING", "0"))
    scene_name = os.environ.get("CWA_DIST_SCENE", "small")
    sc, p = scene(scene_name)
    ctx = cwa.Context(local_rank)
    rk = make_rank(ctx, sc, p, world, rank)
    if world > 1:
        connect_processes(rk, dist)
    for n in split_frames(FRAMES, CALLS):
        rk.step(n, coupling)
    ctx.synchronize()
    cnt = rk.counts()
    owned = rk.download_owned()
    rows = rk.owned_wave_rows(0)
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, (owned, rows, cnt))
    else:
        gathered = [(owned, rows, cnt)]
    ok = True
    if rank == 0:
        errs = [g[2]["err"] for g in gathered]
        P = np.concatenate([g[0] for g in gathered])
        W = np.concatenate([g[1] for g in gathered])
        moved = sum(g[2]["migrated_in"] for g in gathered)
        R, RW = reference(ctx, sc, p, FRAMES, coupling)
        res = compare(P, W, R, RW, moved, {"mode": "processes", "world": world, "frames": FRAMES, "coupling": coupling, "scene": scene_name, "err_bits": errs})
        res["ok"] = bool(res["ok"] and not any(errs))
        ok = res["ok"]
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", type=int, default=0, help="run N ranks inside this one process instead of one process per GPU")
    args = ap.parse_args()
    if args.group > 0:
        r = run_group(args.group, FRAMES, int(os.environ.get("CWA_DIST_COUPLING", "0")), os.environ.get("CWA_DIST_SCENE", "small"))
        print(json.dumps(r), flush=True)
        sys.exit(0 if r["ok"] else 1)
    main_processes()
