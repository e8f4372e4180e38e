"""Diagnostic: error growth of the slab decomposition against the single-GPU path, frame by frame (group mode, one process)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import coupledwateranimation_b200 as cwa
import dist_check as D
from coupledwateranimation_b200.distributed import connect_local, group_step

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scene_name = sys.argv[2] if len(sys.argv) > 2 else "big"
frames_list = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,2,4,8,12,20").split(",")]
coupling = int(os.environ.get("CWA_DIST_COUPLING", "0"))
sc, p = D.scene(scene_name)
import torch
ndev = torch.cuda.device_count()
devs = [r % ndev for r in range(world)]
ctxs = [cwa.Context(devs[r]) for r in range(world)]
ranks = [D.make_rank(ctxs[r], sc, p, world, r) for r in range(world)]
connect_local(ranks)
rc = cwa.Context(devs[0])
D.set_params(rc, sc)
box = sc["box"]
grid = cwa.UniformGrid(rc, 3, (0.0, -0.02, 0.0), (box[0], sc.get("gy", 1.0), box[2]), sc["cells"], p.size, compact_index=True)
sph = cwa.Sph(rc, p.size, grid, particles=p)
wave = cwa.StencilImage2DTripleBuffered(rc, sc["wave"][0], sc["wave"][1], 1, cwa.WAVE_COUPLED)
done = 0
faces = [r.desc.z_lo for r in ranks[1:]]
for fr in frames_list:
    n = fr - done
    group_step(ranks, n, coupling)
    sph.coupled_step(wave, n, coupling)
    done = fr
    for c in ctxs: c.synchronize()
    cnt = [rk.check() for rk in ranks]
    P = np.concatenate([rk.download_owned() for rk in ranks]); P = P[np.argsort(P["extras"][:, 3])]
    R = sph.download()
    W = np.concatenate([rk.owned_wave_rows(0) for rk in ranks]); RW = wave.read_role(0)
    res = {"frames": fr, "count_ok": bool(P.size == R.size and np.array_equal(P["extras"][:, 3], R["extras"][:, 3])),
           "wave_bit_exact": bool(np.array_equal(W.view(np.uint32), RW.view(np.uint32))), "migrated": sum(c["migrated_in"] for c in cnt)}
    if res["count_ok"]:
        nr, ng = np.isnan(R["pos"]).any(1), np.isnan(P["pos"]).any(1)
        res["nan_ref"], res["nan_got"], res["nan_diff"] = int(nr.sum()), int(ng.sum()), int((nr != ng).sum())
        good = ~nr & ~ng
        for f in ("pos", "vel"):
            a, b = P[f][good, :3].astype(np.float64), R[f][good, :3].astype(np.float64)
            scale = float(np.sqrt(np.mean(b ** 2)))
            err = np.abs(a - b).max(1) / scale
            res[f + "_max"] = float(err.max()); res[f + "_out1e-3"] = float((err > 1e-3).mean())
            if f == "vel":
                bad = err > 1e-3
                zb = R["pos"][good, 2][bad]
                if zb.size and faces:
                    dist_face = np.min(np.abs(zb[:, None] - np.array(faces)[None, :]), axis=1)
                    res["bad_near_face_0.05"] = float((dist_face < 0.05).mean()); res["bad_n"] = int(bad.sum())
                res["vel_rms"] = scale
        rho_a, rho_b = P["extras"][good, 0].astype(np.float64), R["extras"][good, 0].astype(np.float64)
        res["rho_max_rel"] = float(np.max(np.abs(rho_a - rho_b) / np.abs(rho_b)))
    print(json.dumps(res), flush=True)
