"""Floating-point half of the oracle: trilinear embedding lookup, SDF decoder, loss, SE(3) pose,
the render_rays composition and one optimiser step -- restated in plain fp32 PyTorch on the CPU,
gradients by autograd.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restated reference code (paths relative to /root/reference):
  src/variations/render_helpers.py:9-10     ray
  src/variations/render_helpers.py:40-70    trilinear_interp / offset_points / get_embeddings
  src/variations/render_helpers.py:74-93    get_features
  src/variations/render_helpers.py:190-318  render_rays
  src/variations/render_helpers.py:321-425  bundle_adjust_frames (one iteration)
  src/variations/render_helpers.py:428-514  track_frame (one iteration)
  src/variations/lidar.py:80-131            Decoder
  src/criterion.py:16-115                   Criterion
  src/se3pose.py:8-92                       OptimizablePose
Pinned against the reference's own Python (imported from /root/reference) by
tests/golden/make_golden.py -> tests/golden/*.npz and tests/test_oracle_vs_reference.py.
"""
import math

import numpy as np
import torch

from . import kernels as K

MAX_DEPTH = K.MAX_DEPTH


# ------------------------------------------------------------------------------------------------
# a-7  trilinear interpolation (render_helpers.py:40-70)
# ------------------------------------------------------------------------------------------------
def corner_table(dtype=torch.float32):
    """q in {0,1}^3, corner k = 4*kx + 2*ky + kz (offset_points with bits=2, x slowest)."""
    q = [[(k >> 2) & 1, (k >> 1) & 1, k & 1] for k in range(8)]
    return torch.tensor(q, dtype=dtype)


def trilinear_weights(sampled_xyz, centre_xyz, voxel_size):
    p = ((sampled_xyz - centre_xyz) / voxel_size + 0.5).unsqueeze(1)          # [M,1,3]
    q = corner_table(p.dtype).unsqueeze(0)                                     # [1,8,3]
    return (p * q + (1 - p) * (1 - q)).prod(dim=-1, keepdim=True)              # [M,8,1]


def get_embeddings(sampled_xyz, centre_xyz, point_feats, voxel_size):
    """point_feats: [M, 8*E] or [M,8,E] (bf16 or fp32).  Returns fp32 [M,E] (fp64 when the sample positions are fp64: the
    precision-apportioning runs of the parity tests)."""
    w = trilinear_weights(sampled_xyz, centre_xyz, voxel_size)
    if point_feats.dim() == 2:
        point_feats = point_feats.view(point_feats.size(0), 8, -1)
    out = (w * point_feats).sum(1)
    return out if out.dtype == torch.float64 else out.float()


def get_features(sampled_idx, sampled_xyz, centres, vertex_rows, emb, voxel_size, contrib=None):
    """get_features (render_helpers.py:74-93) with the two-level vertex->row indirection already
    composed into `vertex_rows` i64[n,8] (= voxel_id2embedding_id[voxel_vertex_idx]).
    contrib: optional dict; the gathered rows then become a leaf of their own (contrib["point_feats"], [M*8,E] in the working
    dtype, with contrib["rows"]), so that a test can read the per-contribution gradients -- the quantity the reference's bf16
    table rounds to bf16 one by one before accumulating (autograd through `.float()` of a bf16 gather)."""
    c = centres[sampled_idx].to(sampled_xyz.dtype)
    rows = vertex_rows[sampled_idx].reshape(-1)
    if contrib is not None:
        pf = emb.detach()[rows].to(sampled_xyz.dtype).requires_grad_()
        contrib["point_feats"], contrib["rows"] = pf, rows
        feats = pf.view(c.size(0), -1)
    else:
        feats = emb[rows].view(c.size(0), -1)
    return get_embeddings(sampled_xyz, c, feats, voxel_size)


# ------------------------------------------------------------------------------------------------
# a-8  decoder (lidar.py:80-131), 'none' embedder, no skips
# ------------------------------------------------------------------------------------------------
class Decoder(torch.nn.Module):
    def __init__(self, depth=2, width=256, in_dim=16, **kw):
        super().__init__()
        self.pts_linears = torch.nn.ModuleList(
            [torch.nn.Linear(in_dim, width)] + [torch.nn.Linear(width, width) for _ in range(depth - 1)])
        self.sdf_out = torch.nn.Linear(width, 1)

    def get_values(self, x):
        h = x
        for l in self.pts_linears:
            h = torch.relu(l(h))
        return self.sdf_out(h)

    def forward(self, x):
        return {"sdf": self.get_values(x)}


# ------------------------------------------------------------------------------------------------
# a-9  loss (criterion.py)
# ------------------------------------------------------------------------------------------------
def sdf_loss(z_vals, sdf, valid_mask, gt_points, points_cos, truncation, max_depth, fs_weight, sdf_weight):
    """z_vals/sdf/valid_mask: [R_hit,S]; gt_points [R_hit,3]; points_cos [R_hit].
    Returns (loss, dict(fs_loss, sdf_loss, n_fs, n_sdf))."""
    depth = torch.norm(gt_points, 2, -1) * points_cos.view(-1)
    z = z_vals * points_cos.view(-1, 1)
    d = depth.unsqueeze(-1).expand(*z.shape)
    if sdf.dtype == torch.float64:
        # precision-apportioning run: the masks are DISCRETE decisions of the fp32 program (criterion.py:68-82 on fp32 tensors);
        # they are taken in fp32 here too, only the differentiable arithmetic below runs in fp64
        d32 = (torch.norm(gt_points.float(), 2, -1) * points_cos.float().view(-1)).unsqueeze(-1).expand(*z.shape)
        z32 = z_vals.float() * points_cos.float().view(-1, 1)
        one, zero = torch.ones_like(z), torch.zeros_like(z)
        front = torch.where(z32 < (d32 - truncation), one, zero)
        back = torch.where(z32 > (d32 + truncation), one, zero)
        dmask = torch.where((d32 > 0.0) & (d32 < max_depth), one, zero)
        z, d = z32.double(), d32.double()
    else:
        front = torch.where(z < (d - truncation), torch.ones_like(z), torch.zeros_like(z))
        back = torch.where(z > (d + truncation), torch.ones_like(z), torch.zeros_like(z))
        dmask = torch.where((d > 0.0) & (d < max_depth), torch.ones_like(d), torch.zeros_like(d))
    smask = (1.0 - front) * (1.0 - back) * dmask
    n_fs = torch.count_nonzero(front).to(z.dtype)
    n_sdf = torch.count_nonzero(smask).to(z.dtype)
    n = n_sdf + n_fs
    w_fs = 1.0 - n_fs / n
    w_sdf = 1.0 - n_sdf / n
    fs = torch.mean(torch.square(sdf * front * valid_mask - front)) * w_fs
    sl = torch.mean(torch.square((z + sdf * truncation) * smask * valid_mask - d * smask)) * w_sdf
    loss = fs_weight * fs + sdf_weight * sl
    return loss, {"fs_loss": fs, "sdf_loss": sl, "n_fs": n_fs, "n_sdf": n_sdf}


# ------------------------------------------------------------------------------------------------
# a-10  SE(3) pose (se3pose.py)
# ------------------------------------------------------------------------------------------------
def taylor_A(x, nth=10):
    ans = torch.zeros_like(x)
    denom = 1.0
    for i in range(nth + 1):
        if i > 0:
            denom *= (2 * i) * (2 * i + 1)
        ans = ans + (-1) ** i * x ** (2 * i) / denom
    return ans


def taylor_B(x, nth=10):
    ans = torch.zeros_like(x)
    denom = 1.0
    for i in range(nth + 1):
        denom *= (2 * i + 1) * (2 * i + 2)
        ans = ans + (-1) ** i * x ** (2 * i) / denom
    return ans


def skew(w):
    w0, w1, w2 = w.unbind(-1)
    O = torch.zeros_like(w0)
    return torch.stack([torch.stack([O, -w2, w1], -1), torch.stack([w2, O, -w0], -1), torch.stack([-w1, w0, O], -1)], -2)


def pose_rotation(data):
    w = data[3:]
    wx = skew(w)
    theta = w.norm(dim=-1)[..., None, None]
    I = torch.eye(3, dtype=data.dtype)
    return I + taylor_A(theta) * wx + taylor_B(theta) * wx @ wx


def pose_translation(data):
    return data[:3]


def pose_log(R, eps=1e-7):
    trace = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    theta = ((trace - 1) / 2).clamp(-1 + eps, 1 - eps).acos_()[..., None, None] % np.pi
    lnR = 1 / (2 * taylor_A(theta) + 1e-8) * (R - R.transpose(-2, -1))
    return torch.stack([lnR[..., 2, 1], lnR[..., 0, 2], lnR[..., 1, 0]], -1)


def pose_from_matrix(Rt):
    return torch.cat([Rt[:3, 3], pose_log(Rt[:3, :3])], -1)


# ------------------------------------------------------------------------------------------------
# a-5  render_rays (render_helpers.py:190-318)
# ------------------------------------------------------------------------------------------------
def render_rays(rays_o, rays_d, map_np, emb, decoder, step_size, voxel_size, max_distance, deterministic=True,
                noise=None, raw_hits=None, rays_np=None, contrib=None):
    """rays_o/rays_d: torch fp32 [R,3] (may require grad).  map_np: dict of numpy arrays
    centres f32[n,3], structure i32[n,9], vertex_rows i64[n,8].  emb: torch [V,E] (bf16 or fp32).
    Returns dict like the reference (z_vals, sdf, ray_mask, valid_mask, sampled_xyz) plus the raw
    sample tensors, or None."""
    # rays_np: the fp32 rays of the fp32 program, for the discrete part (traversal, sampling) of an fp64 apportioning run
    ro = rays_np[0] if rays_np is not None else rays_o.detach().numpy().astype(np.float32)
    rd = rays_np[1] if rays_np is not None else rays_d.detach().numpy().astype(np.float32)
    inter, hits = K.ray_intersect(ro, rd, map_np["centres"], map_np["structure"], voxel_size, 20, max_distance, raw=raw_hits)
    if hits.sum() <= 0:
        return None
    inter_h = {k: v[hits] for k, v in inter.items()}
    samples = K.ray_sample(inter_h, step_size=step_size, fixed=deterministic, noise=noise)
    if samples is None:
        return None
    hmask = torch.from_numpy(hits)
    if rays_np is not None:
        # straight-through: values of the fp32 program's rays, derivatives of the working-precision expression -- an fp64 run then
        # differentiates the SAME function at the SAME (fp32-rounded) points, and differs from an fp32 run by arithmetic rounding only
        rays_o = rays_o + (torch.from_numpy(ro).to(rays_o.dtype) - rays_o).detach()
        rays_d = rays_d + (torch.from_numpy(rd).to(rays_d.dtype) - rays_d).detach()
    ro_h, rd_h = rays_o[hmask], rays_d[hmask]
    depth = torch.from_numpy(samples["sampled_point_depth"]).to(rays_o.dtype)
    sidx = torch.from_numpy(samples["sampled_point_voxel_idx"]).long()
    smask = sidx.ne(-1)
    if smask.sum() == 0:
        return None
    xyz = ro_h.unsqueeze(1) + rd_h.unsqueeze(1) * depth.unsqueeze(2)          # ray(): mul then add
    if rays_np is not None:      # sample positions as the fp32 program rounds them (|xyz| ~ 2000 m: 1.2e-4 m per ulp), straight-through
        ro32, rd32 = torch.from_numpy(ro)[hmask], torch.from_numpy(rd)[hmask]
        xyz32 = ro32.unsqueeze(1) + rd32.unsqueeze(1) * torch.from_numpy(samples["sampled_point_depth"]).unsqueeze(2)
        xyz = xyz + (xyz32.to(xyz.dtype) - xyz).detach()
    xyz_v = xyz[smask]
    idx_v = sidx[smask]
    centres = torch.from_numpy(map_np["centres"])
    vrows = torch.from_numpy(map_np["vertex_rows"]).long()
    feats = get_features(idx_v, xyz_v, centres, vrows, emb, voxel_size, contrib)
    sdf_v = decoder(feats)["sdf"]
    sdf = torch.ones(smask.shape, dtype=sdf_v.dtype).masked_scatter(smask, sdf_v.squeeze(-1))
    return {"z_vals": depth, "sdf": sdf, "ray_mask": hmask, "valid_mask": smask, "sampled_xyz": xyz_v,
            "sampled_idx": sidx, "feats": feats, "sdf_valid": sdf_v.squeeze(-1), "intersections": inter,
            "samples": samples}


def mapping_iteration(frames, map_np, emb, decoder, cfg, deterministic=True, noise=None, rays_np=None, contrib=None):
    """One iteration of bundle_adjust_frames' loop body (render_helpers.py:356-423) WITHOUT the
    optimiser step: returns (loss, outputs).  frames: list of dict(pose=torch[6] param,
    dirs=torch[N,3] unit ray dirs of the selected rays, points=[N,3], cos=[N])."""
    ro, rd, pts, cos = [], [], [], []
    for f in frames:
        R = pose_rotation(f["pose"])
        t = pose_translation(f["pose"])
        d = f["dirs"] @ R.transpose(-1, -2)
        ro.append(t.reshape(1, -1).expand_as(d))
        rd.append(d)
        pts.append(f["points"])
        cos.append(f["cos"])
    ro, rd, pts, cos = torch.cat(ro), torch.cat(rd), torch.cat(pts), torch.cat(cos)
    out = render_rays(ro, rd, map_np, emb, decoder, cfg["step_size"], cfg["voxel_size"], cfg["max_distance"],
                      deterministic=deterministic, noise=noise, rays_np=rays_np, contrib=contrib)
    if out is None:
        return None, None
    out["rays_np"] = (ro.detach().numpy().astype(np.float32), rd.detach().numpy().astype(np.float32))
    m = out["ray_mask"]
    loss, parts = sdf_loss(out["z_vals"], out["sdf"], out["valid_mask"], pts[m], cos[m], cfg["truncation"],
                           cfg["max_depth"], cfg["fs_weight"], cfg["sdf_weight"])
    out.update(parts)
    return loss, out
