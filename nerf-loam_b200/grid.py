"""Drop-in for the reference's `grid` extension module (third_party/sparse_voxels/src/binding.cpp:12-20).

Live functions (same tensor contracts, same RuntimeError on bad inputs as the CHECK_* macros of
third_party/sparse_voxels/include/utils.h:10-34): svo_intersect, inverse_cdf_sampling.
The other five exports are dead code in NeRF-LOAM (never called) and raise
NotImplementedError here.

    sys.modules["grid"] = nerfloam_b200.grid        # before importing variations.voxel_helpers
"""
import torch

from . import _capi


def _chk(t, name, dtype):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be a {'float' if dtype == torch.float32 else 'int'} tensor")


def svo_intersect(ray_start, ray_dir, points, children, voxelsize, n_max):
    """intersect.cpp:83-112.  ray_start/ray_dir f32[B,K,3], points f32[B,n,3], children i32[B,n,9]
    -> (idx i32[B,K,n_max], min_depth f32, max_depth f32)."""
    _chk(ray_start, "ray_start", torch.float32)
    _chk(ray_dir, "ray_dir", torch.float32)
    _chk(points, "points", torch.float32)
    _chk(children, "children", torch.int32)
    B, K = ray_start.shape[0], ray_start.shape[1]
    n = points.shape[1]
    idx = torch.empty((B, K, n_max), dtype=torch.int32, device=ray_start.device)
    mn = torch.empty((B, K, n_max), dtype=torch.float32, device=ray_start.device)
    mx = torch.empty((B, K, n_max), dtype=torch.float32, device=ray_start.device)
    with torch.cuda.device(ray_start.device):
        _capi.check(_capi.lib().nl_svo_intersect(B, n, K, float(voxelsize), int(n_max), _capi.ptr(ray_start), _capi.ptr(ray_dir),
                                                 _capi.ptr(points), _capi.ptr(children), _capi.ptr(idx), _capi.ptr(mn),
                                                 _capi.ptr(mx), _capi.stream_ptr()), "nl_svo_intersect")
    _capi.LAUNCHES += 1
    return idx, mn, mx


def inverse_cdf_sampling(pts_idx, min_depth, max_depth, uniform_noise, probs, steps, fixed_step_size):
    """sample.cpp:56-95.  [B,K,P] hits + noise f32[B,K,S] -> (idx i32[B,K,S], depth f32, dists f32)."""
    _chk(pts_idx, "pts_idx", torch.int32)
    for t, nm in ((min_depth, "min_depth"), (max_depth, "max_depth"), (uniform_noise, "uniform_noise"), (probs, "probs"),
                  (steps, "steps")):
        _chk(t, nm, torch.float32)
    B, K, P = pts_idx.shape
    S = uniform_noise.shape[-1]
    si = torch.empty((B, K, S), dtype=torch.int32, device=pts_idx.device)
    sd = torch.empty((B, K, S), dtype=torch.float32, device=pts_idx.device)
    sl = torch.empty((B, K, S), dtype=torch.float32, device=pts_idx.device)
    with torch.cuda.device(pts_idx.device):
        _capi.check(_capi.lib().nl_inverse_cdf_sampling(B, K, P, S, float(fixed_step_size), _capi.ptr(pts_idx), _capi.ptr(min_depth),
                                                        _capi.ptr(max_depth), _capi.ptr(uniform_noise), _capi.ptr(probs),
                                                        _capi.ptr(steps), _capi.ptr(si), _capi.ptr(sd), _capi.ptr(sl),
                                                        _capi.stream_ptr()), "nl_inverse_cdf_sampling")
    _capi.LAUNCHES += 1
    return si, sd, sl


def _dead(name):
    def f(*a, **k):
        raise NotImplementedError(f"grid.{name} is never called by NeRF-LOAM (dead NSVF heritage, SURVEY.md 2.3); "
                                  "it is not part of the hot path and is not implemented")
    f.__name__ = name
    return f


ball_intersect = _dead("ball_intersect")
aabb_intersect = _dead("aabb_intersect")
triangle_intersect = _dead("triangle_intersect")
uniform_ray_sampling = _dead("uniform_ray_sampling")
build_octree = _dead("build_octree")
