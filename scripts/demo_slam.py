"""A miniature of the reference's SLAM loop (src/tracking.py:110-148, src/mapping.py:96-170) on synthetic scans, run entirely through the
drop-in API: per scan  track_frame (captured graph) -> bundle_adjust_frames over the keyframe window (captured graph) -> incremental
map update (MapUpdater) -> publication to the tracker (share.SharedMap); at the end GPU marching cubes.
BASELINE.json config 2 ("KITTI seq 00 tracking+mapping incremental loop") / config 3 ("... + marching cubes") in spirit; prints a
JSON summary (per-stage ms per scan, trajectory error against the synthetic ground truth).

    python scripts/demo_slam.py [--scans 12] [--spacing 0.5] [--json out.json]
"""
import argparse, json, os, sys, time
from types import SimpleNamespace
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfloam_b200 as nl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=12); ap.add_argument("--spacing", type=float, default=0.5)
    ap.add_argument("--init-calls", type=int, default=12); ap.add_argument("--json", default=None)
    ap.add_argument("--ray-selection", default="device", choices=["device", "host"])
    a = ap.parse_args()
    dev = torch.device("cuda"); syn, rh = nl.synthetic, nl.render_helpers
    vs, md, trunc = 0.3, 40.0, 0.3
    crit = nl.criterion.Criterion(SimpleNamespace(criteria={"eiko_weight": 0.1, "sdf_weight": 10000.0, "fs_weight": 1.0, "sdf_truncation": trunc},
                                                  data_specs={"max_depth": md}))
    torch.manual_seed(777)
    dec = nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).to(dev)
    mu = nl.mapping.MapUpdater(vs, device=dev, reserve_nodes=400_000, reserve_rows=400_000)       # embeddings start at zero like the reference
    shared = nl.share.SharedMap(lambda: nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).to(dev))
    mu.readers = shared
    scans = [syn.make_scan(seed=1000 + i, sensor_xyz=(a.spacing * i, 0.05 * np.sin(0.7 * i), 0.0), yaw=0.01 * np.sin(0.5 * i)) for i in range(a.scans)]
    gt = [torch.from_numpy(s[2]) for s in scans]
    T = {k: [] for k in ("track", "map", "update", "publish")}
    def tick():
        torch.cuda.synchronize(); return time.perf_counter()
    mk = lambda i, pose: nl.frame.LidarFrame(i, torch.from_numpy(scans[i][0]), torch.from_numpy(scans[i][1]), pose, new_keyframe=True)
    lr = [0.01, 0.005, 0.001]
    ba = lambda fr, upd_dec, it=25: rh.bundle_adjust_frames(fr, mu.embeddings, ms, dec, crit, vs, 0.5 * vs, N_rays=2048, num_iterations=it, truncation=trunc,
                                                            max_voxel_hit=20, max_distance=md, learning_rate=lr, update_pose=True, update_decoder=upd_dec,
                                                            ray_selection=a.ray_selection)
    # ---- first scan: map it, then train on it while "waiting for the tracker" (mapping.py:100-108) ----
    f0 = mk(0, nl.se3pose.OptimizablePose.from_matrix(gt[0].clone()))
    ms = mu.create_voxels(scans[0][0], gt[0])
    keyframes, est = [f0], [gt[0].clone()]
    t0 = tick()
    for _ in range(a.init_calls):
        ba([f0], True)
    t_init = (tick() - t0) * 1e3
    shared.publish(ms, dec)
    rel = None
    for i in range(1, a.scans):
        last = est[-1]
        guess = last @ rel if rel is not None else last.clone()                            # constant-velocity model (tracking.py:112-117)
        fr = mk(i, nl.se3pose.OptimizablePose.from_matrix(guess.clone()))
        # KITTI's tracking rate (0.06, /3 from the third scan on).  The first tracked scan has no motion prior: 125 iterations, and a rate of
        # 0.01 because frame index < 2 doubles it -- 0.12 m Adam steps on 0.3 m voxels are bistable (scripts/sweep_tracking_lr.py)
        t0 = tick()
        m_t, dec_t, slot = shared.acquire()
        pose, hit = rh.track_frame(fr.pose, fr, m_t, dec_t, crit, vs, N_rays=2048, step_size=0.2 * vs, num_iterations=25 if rel is not None else 125,
                                   truncation=trunc, learning_rate=0.06 if rel is not None else 0.01, max_voxel_hit=20, max_distance=md, ray_selection=a.ray_selection)
        shared.release(slot)
        T["track"].append((tick() - t0) * 1e3)
        if hit is None:
            pose = nl.se3pose.OptimizablePose.from_matrix(guess.clone())
        fr.pose = nl.se3pose.OptimizablePose(pose.data.detach().cpu().clone())
        t0 = tick()
        ba(keyframes[-4:] + [fr], upd_dec=(i < 5))                                          # window_size 4 + the tracked frame, freeze_frame 5
        T["map"].append((tick() - t0) * 1e3)
        t0 = tick()
        ms = mu.create_voxels(scans[i][0], fr.get_pose().detach())
        T["update"].append((tick() - t0) * 1e3)
        t0 = tick()
        shared.publish(ms, dec)
        T["publish"].append((tick() - t0) * 1e3)
        cur = fr.get_pose().detach().cpu()
        rel = torch.linalg.inv(last) @ cur
        est.append(cur)
        if i % 2 == 0:
            keyframes.append(fr)
    t0 = tick()
    verts, faces = nl.mesh.extract_mesh(dec, ms, vs, res=8)
    t_mesh = (tick() - t0) * 1e3
    err = [float((e[:3, 3] - g[:3, 3]).norm()) for e, g in zip(est, gt)]
    rot = [float(torch.linalg.matrix_norm(e[:3, :3] - g[:3, :3])) for e, g in zip(est, gt)]
    out = {"scans": a.scans, "spacing_m": a.spacing, "translation_error_m": {"mean": float(np.mean(err[1:])), "max": float(np.max(err[1:])), "last": err[-1]},
           "rotation_error_fro": {"mean": float(np.mean(rot[1:])), "max": float(np.max(rot[1:]))},
           "ms_per_scan": {k: float(np.median(v)) for k, v in T.items()}, "ms_first_tracked_scan": T["track"][0], "init_mapping_ms": t_init,
           "mesh": {"ms": t_mesh, "vertices": int(verts.shape[0]), "triangles": int(faces.shape[0])},
           "map": {"nodes": ms.n_nodes, "embedding_rows": int(mu.n_rows), "last_update": mu.last_update},
           "per_scan_translation_error_m": [round(e, 4) for e in err]}
    print(json.dumps(out))
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
