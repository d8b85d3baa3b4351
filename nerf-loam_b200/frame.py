"""LidarFrame host container with the interface of src/lidarFrame.py:10-57 and the Gumbel top-k ray
selection of src/utils/sample_util.py:4-19 (CPU torch RNG, so seeds select the same rays as the reference)."""
import torch
import torch.nn as nn

from .se3pose import OptimizablePose

POSE_OFFSET = 2000  # lidarFrame.py:18


def sample_rays(mask, num_samples):
    """sample_util.py:12-19: boolean mask [B,H,W] with exactly num_samples True, uniform without replacement."""
    B, H, W = mask.shape
    probs = (mask / (mask.sum() + 1e-9)).reshape(B, -1)
    logp = torch.log(probs + 1e-9)
    gumbel = -torch.log(-torch.log(torch.rand_like(logp) + 1e-7) + 1e-7)
    idx = (logp + gumbel).topk(num_samples, dim=-1)[1]
    return torch.zeros_like(probs).scatter_(-1, idx, 1).reshape(B, H, W) > 0


class LidarFrame(nn.Module):
    def __init__(self, index, points, pointsCos, pose=None, new_keyframe=False):
        super().__init__()
        self.index = index
        self.num_point = len(points)
        self.points = points
        self.pointsCos = pointsCos
        if (not new_keyframe) and (pose is not None):
            pose[:3, 3] += POSE_OFFSET
            self.pose = OptimizablePose.from_matrix(torch.tensor(pose, requires_grad=True, dtype=torch.float32))
        elif new_keyframe:
            self.pose = pose
        self.rays_d = self.get_rays()
        self.rel_pose = None

    def get_pose(self):
        return self.pose.matrix()

    def get_translation(self):
        return self.pose.translation()

    def get_rotation(self):
        return self.pose.rotation()

    def get_points(self):
        return self.points

    def get_pointsCos(self):
        return self.pointsCos

    def set_rel_pose(self, rel_pose):
        self.rel_pose = rel_pose

    def get_rel_pose(self):
        return self.rel_pose

    @torch.no_grad()
    def get_rays(self):
        self.rays_norm = torch.norm(self.points, 2, -1, keepdim=True) + 1e-8
        return (self.points / self.rays_norm).unsqueeze(1).float()

    @torch.no_grad()
    def sample_rays(self, N_rays, track=False):
        self.sample_mask = sample_rays(torch.ones((self.num_point, 1))[None, ...], N_rays)[0, ...]
