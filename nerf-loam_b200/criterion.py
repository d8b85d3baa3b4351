"""Criterion with the interface of src/criterion.py:6-115 (free-space + truncated-SDF L2 loss with
class-balance weights).  `forward` works on a render_rays output dict like the reference; inside
bundle_adjust_frames / track_frame the same loss is evaluated by the fused CUDA kernels (csrc/render.cu
for the masks and counts, csrc/mlp.cu for the squared errors and d loss / d sdf) and only the
configuration stored here is read."""
import torch
import torch.nn as nn


class Criterion(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.eiko_weight = args.criteria["eiko_weight"]
        self.sdf_weight = args.criteria["sdf_weight"]
        self.fs_weight = args.criteria["fs_weight"]
        self.truncation = args.criteria["sdf_truncation"]
        self.max_dpeth = args.data_specs["max_depth"]  # (sic) attribute name kept for drop-in compatibility

    def kernel_config(self):
        return dict(truncation=float(self.truncation), max_depth=float(self.max_dpeth), fs_weight=float(self.fs_weight),
                    sdf_weight=float(self.sdf_weight))

    def get_masks(self, z_vals, depth, epsilon):
        one, zero = torch.ones_like(z_vals), torch.zeros_like(z_vals)
        front = torch.where(z_vals < (depth - epsilon), one, zero)
        back = torch.where(z_vals > (depth + epsilon), one, zero)
        in_range = torch.where((depth > 0.0) & (depth < self.max_dpeth), one, zero)
        sdf_mask = (1.0 - front) * (1.0 - back) * in_range
        n_fs = torch.count_nonzero(front).float()
        n_sdf = torch.count_nonzero(sdf_mask).float()
        n = n_sdf + n_fs
        return front, sdf_mask, 1.0 - n_fs / n, 1.0 - n_sdf / n

    def forward(self, outputs, obs, pointsCos, use_color_loss=True, use_depth_loss=True, compute_sdf_loss=True,
                weight_depth_loss=False, compute_eikonal_loss=False):
        if compute_eikonal_loss:
            raise NotImplementedError("the eikonal branch is never enabled by the reference (criterion.py:18)")
        ray_mask = outputs["ray_mask"]
        cos = pointsCos[ray_mask].view(-1)
        depth = torch.norm(obs[ray_mask], 2, -1) * cos
        z = outputs["z_vals"] * cos.view(-1, 1)
        sdf, valid = outputs["sdf"], outputs["valid_mask"]
        d = depth.unsqueeze(-1).expand(*z.shape)
        front, sdf_mask, w_fs, w_sdf = self.get_masks(z, d, self.truncation)
        fs_loss = torch.mean(torch.square(sdf * front * valid - front)) * w_fs
        sdf_loss = torch.mean(torch.square((z + sdf * self.truncation) * sdf_mask * valid - d * sdf_mask)) * w_sdf
        loss = self.fs_weight * fs_loss + self.sdf_weight * sdf_loss
        return loss, {"fs_loss": fs_loss.item(), "sdf_loss": sdf_loss.item(), "loss": loss.item()}
