"""ncu target: a few eager tracking iterations (device-side ray selection) on the headline scan, to see which kernels make up a
tracking iteration.  Usage: ncu --metrics gpu__time_duration.sum --csv --log-file out.csv python scripts/profile_tracking.py"""
import os, sys
from types import SimpleNamespace
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfloam_b200 as nl
dev = torch.device("cuda"); syn = nl.synthetic
pts, cos, pose = syn.make_scan(seed=777)
mu = nl.mapping.MapUpdater(0.3, init_std=0.01, seed=777, device=dev)
ms = mu.insert_voxels(torch.from_numpy(syn.voxelize(pts, pose, 0.3)))
torch.manual_seed(777)
dec = nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).to(dev)
crit = nl.criterion.Criterion(SimpleNamespace(criteria={"eiko_weight": 0.1, "sdf_weight": 10000.0, "fs_weight": 1.0, "sdf_truncation": 0.3}, data_specs={"max_depth": 40.0}))
fr = nl.frame.LidarFrame(5, torch.from_numpy(pts), torch.from_numpy(cos), nl.se3pose.OptimizablePose.from_matrix(torch.from_numpy(pose.copy())), new_keyframe=True)
for _ in range(2):
    nl.render_helpers.track_frame(fr.pose, fr, ms, dec, crit, 0.3, N_rays=2048, step_size=0.06, num_iterations=int(os.environ.get("IT", 5)), truncation=0.3,
                                  learning_rate=0.06, max_voxel_hit=20, max_distance=40.0, ray_selection="device", cuda_graph=False)
torch.cuda.synchronize()
