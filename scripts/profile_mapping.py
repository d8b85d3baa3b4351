"""ncu target: a few eager mapping iterations (5 keyframes x 2048 rays, device-side ray selection, decoder trained unless FREEZE=1),
to see which kernels make up a bundle_adjust_frames iteration at the reference's real size.
Usage: ncu --metrics gpu__time_duration.sum --csv --log-file out.csv python scripts/profile_mapping.py"""
import os, sys
from types import SimpleNamespace
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfloam_b200 as nl
dev = torch.device("cuda"); syn = nl.synthetic
mu = nl.mapping.MapUpdater(0.3, init_std=0.01, seed=777, device=dev)
frames = []
for i in range(5):
    pts, cos, pose = syn.make_scan(seed=777 + i, sensor_xyz=(1.0 * i, 0.0, 0.0))
    ms = mu.create_voxels(pts, torch.from_numpy(pose))
    frames.append(nl.frame.LidarFrame(i, torch.from_numpy(pts), torch.from_numpy(cos), nl.se3pose.OptimizablePose.from_matrix(torch.from_numpy(pose.copy())), new_keyframe=True))
torch.manual_seed(777)
dec = nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).to(dev)
crit = nl.criterion.Criterion(SimpleNamespace(criteria={"eiko_weight": 0.1, "sdf_weight": 10000.0, "fs_weight": 1.0, "sdf_truncation": 0.3}, data_specs={"max_depth": 40.0}))
for _ in range(2):
    nl.render_helpers.bundle_adjust_frames(frames, mu.embeddings, ms, dec, crit, 0.3, 0.15, N_rays=2048, num_iterations=int(os.environ.get("IT", 3)), truncation=0.3,
                                           max_voxel_hit=20, max_distance=40.0, learning_rate=[0.01, 0.005, 0.001], update_pose=True,
                                           update_decoder=os.environ.get("FREEZE", "0") != "1", ray_selection="device", cuda_graph=False)
torch.cuda.synchronize()
