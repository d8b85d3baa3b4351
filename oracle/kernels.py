"""ctypes bindings to oracle/nl_oracle.c + numpy restatement of the Python wrappers around the two
live reference CUDA kernels.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restated reference code (paths relative to /root/reference):
  src/variations/voxel_helpers.py:92-137   SparseVoxelOctreeRayIntersect.forward
  src/variations/voxel_helpers.py:262-344  InverseCDFRaySampling.forward
  src/variations/voxel_helpers.py:530-567  ray_intersect
  src/variations/voxel_helpers.py:570-598  ray_sample
  src/mapping.py:320-326                   update_grid_features (centres / children glue)
"""
import ctypes
import math

import numpy as np

from . import build as _build

MAX_DEPTH = 80  # voxel_helpers.py:24

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        L = _lib
        L.nlo_octree_create.restype = ctypes.c_void_p
        L.nlo_octree_create.argtypes = [ctypes.c_int]
        L.nlo_octree_destroy.argtypes = [ctypes.c_void_p]
        L.nlo_octree_insert.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.nlo_octree_count.restype = ctypes.c_int64
        L.nlo_octree_count.argtypes = [ctypes.c_void_p]
        L.nlo_octree_count_leaf.restype = ctypes.c_int64
        L.nlo_octree_count_leaf.argtypes = [ctypes.c_void_p]
        L.nlo_octree_export.argtypes = [ctypes.c_void_p] * 4
        L.nlo_encode.restype = ctypes.c_uint64
        L.nlo_encode.argtypes = [ctypes.c_int] * 3
        L.nlo_svo_intersect.restype = ctypes.c_int
        L.nlo_svo_intersect.argtypes = [ctypes.c_int64, ctypes.c_float, ctypes.c_int] + [ctypes.c_void_p] * 7
        L.nlo_inverse_cdf_sampling.argtypes = ([ctypes.c_int] * 4 + [ctypes.c_float] + [ctypes.c_void_p] * 9)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Octree:
    """Oracle octree with the reference method names (bindings.cpp:11-31)."""

    def __init__(self):
        self._h = None

    def init(self, grid_dim, feat_dim, voxel_size):
        self._h = lib().nlo_octree_create(int(grid_dim))
        self.voxel_size = voxel_size

    def insert(self, vox):
        vox = np.ascontiguousarray(np.asarray(vox), dtype=np.int32)
        assert vox.ndim == 2 and vox.shape[1] == 3
        lib().nlo_octree_insert(self._h, _p(vox), vox.shape[0])

    def count_nodes(self):
        return int(lib().nlo_octree_count(self._h))

    def count_leaf_nodes(self):
        return int(lib().nlo_octree_count_leaf(self._h))

    def get_centres_and_children(self):
        n = self.count_nodes()
        voxels = np.zeros((n, 4), np.float32)
        children = np.zeros((n, 8), np.float32)
        features = np.zeros((n, 8), np.int32)
        lib().nlo_octree_export(self._h, _p(voxels), _p(children), _p(features))
        return voxels, children, features

    def __del__(self):
        if self._h is not None and _lib is not None:
            _lib.nlo_octree_destroy(self._h)
            self._h = None


def map_arrays(voxels, children, features, voxel_size):
    """mapping.py:320-326: centres f32[n,3], children i32[n,9] (8 ids + side), vertex idx i32[n,8]."""
    vs = np.float32(voxel_size)
    centres = ((voxels[:, :3] + voxels[:, 3:] / np.float32(2)) * vs).astype(np.float32)
    structure = np.concatenate([children, voxels[:, 3:]], -1).astype(np.int32)
    return centres, structure, features.astype(np.int32)


def svo_intersect(ray_start, ray_dir, centres, structure, voxel_size, n_max=20):
    """intersect.cpp:83-112 + intersect_gpu.cu:193-272 for rays [R,3] (per-ray results do not depend on
    the G-way batching of voxel_helpers.py:97-108)."""
    rs = np.ascontiguousarray(ray_start, np.float32).reshape(-1, 3)
    rd = np.ascontiguousarray(ray_dir, np.float32).reshape(-1, 3)
    R = rs.shape[0]
    idx = np.empty((R, n_max), np.int32)
    mn = np.empty((R, n_max), np.float32)
    mx = np.empty((R, n_max), np.float32)
    c = np.ascontiguousarray(centres, np.float32)
    s = np.ascontiguousarray(structure, np.int32)
    depth = lib().nlo_svo_intersect(R, float(voxel_size), int(n_max), _p(rs), _p(rd), _p(c), _p(s), _p(idx), _p(mn), _p(mx))
    assert depth >= 0, "DFS stack overflow (reference asserts ptr < 256)"
    return idx, mn, mx


def ray_intersect(ray_start, ray_dir, centres, structure, voxel_size, max_hits=None, max_distance=MAX_DEPTH, raw=None):
    """voxel_helpers.py:530-567 (max_hits argument is ignored there too: hard-coded 20).
    raw: optional precomputed (idx, min_depth, max_depth) [R,20] of svo_intersect (e.g. from the GPU kernel that is
    itself checked bit-exact against the compiled reference), so that everything downstream sees identical inputs."""
    if raw is not None:
        pts_idx, min_depth, max_depth = [np.array(a, copy=True) for a in raw]
    else:
        pts_idx, min_depth, max_depth = svo_intersect(ray_start, ray_dir, centres, structure, voxel_size, 20)
    md = np.float32(max_distance)
    min_depth[pts_idx == -1] = md
    max_depth[pts_idx == -1] = md
    order = np.argsort(min_depth, axis=-1, kind="stable")  # torch.sort is not guaranteed stable: ties are
    min_depth = np.take_along_axis(min_depth, order, -1)   # excluded from exact comparisons in the tests
    max_depth = np.take_along_axis(max_depth, order, -1)
    pts_idx = np.take_along_axis(pts_idx, order, -1)
    pts_idx[max_depth > np.float32(2 * max_distance)] = -1
    pts_idx[min_depth > md] = -1
    min_depth[pts_idx == -1] = md
    max_depth[pts_idx == -1] = md
    mh = int((pts_idx != -1).sum(-1).max()) if pts_idx.shape[0] else 0
    min_depth, max_depth, pts_idx = min_depth[:, :mh], max_depth[:, :mh], pts_idx[:, :mh]
    hits = (pts_idx != -1).any(-1)
    return {"min_depth": min_depth, "max_depth": max_depth, "intersected_voxel_idx": pts_idx}, hits


def seq_sum(x):
    """Left-to-right fp32 sum over the last axis (the order the fused CUDA path uses; torch's own
    reduction order for `dists.sum(-1)` is not sequential, so probs/steps of the reference can differ
    from these in the last bit -- the op-level test feeds identical probs/steps instead)."""
    s = np.zeros(x.shape[:-1], np.float32)
    for c in range(x.shape[-1]):
        s = (s + x[..., c]).astype(np.float32)
    return s


def inverse_cdf_sampling(pts_idx, min_depth, max_depth, probs, steps, fixed_step_size=-1, deterministic=False,
                         noise=None):
    """voxel_helpers.py:262-344.  Inputs [N,P] / [N].  `noise` (optional) must have the padded layout
    (200, ceil(N/200), max_steps) the reference generates at :297-301."""
    G, N, P = 200, pts_idx.shape[0], pts_idx.shape[1]
    H = int(np.ceil(N / G)) * G
    if H > N:
        def pad(a):
            return np.concatenate([a, np.broadcast_to(a[:1], (H - N,) + a.shape[1:])], 0)
        pts_idx, min_depth, max_depth, probs, steps = map(pad, (pts_idx, min_depth, max_depth, probs, steps))
    pts_idx = pts_idx.reshape(G, -1, P)
    min_depth = min_depth.reshape(G, -1, P)
    max_depth = max_depth.reshape(G, -1, P)
    probs = probs.reshape(G, -1, P)
    steps = steps.reshape(G, -1)
    max_steps = int(np.ceil(steps).astype(np.int64).max()) + P
    if noise is None:
        assert deterministic, "pass noise explicitly for the stochastic case"
        noise = np.full(min_depth.shape[:-1] + (max_steps,), 0.5, np.float32)
    assert noise.shape == min_depth.shape[:-1] + (max_steps,)
    chunk = 4 * G
    outs = []
    for i in range(0, min_depth.shape[1], chunk):
        a = [np.ascontiguousarray(x[:, i:i + chunk]) for x in (pts_idx, min_depth, max_depth, noise, probs, steps)]
        b, nr = a[0].shape[0], a[0].shape[1]
        si = np.empty((b, nr, max_steps), np.int32)
        sd = np.empty((b, nr, max_steps), np.float32)
        sl = np.empty((b, nr, max_steps), np.float32)
        lib().nlo_inverse_cdf_sampling(b, nr, P, max_steps, float(fixed_step_size),
                                       _p(a[0].astype(np.int32)), _p(a[1].astype(np.float32)), _p(a[2].astype(np.float32)),
                                       _p(a[3].astype(np.float32)), _p(a[4].astype(np.float32)), _p(a[5].astype(np.float32)),
                                       _p(si), _p(sd), _p(sl))
        outs.append((si, sd, sl))
    sampled_idx, sampled_depth, sampled_dists = [np.concatenate([o[i] for o in outs], 1) for i in range(3)]
    sampled_idx = sampled_idx.reshape(H, -1)[:N]
    sampled_depth = sampled_depth.reshape(H, -1)[:N]
    sampled_dists = sampled_dists.reshape(H, -1)[:N]
    max_len = int((sampled_idx != -1).sum(-1).max())
    return sampled_idx[:, :max_len], sampled_depth[:, :max_len], sampled_dists[:, :max_len]


def ray_sample(intersection_outputs, step_size=0.01, fixed=False, noise=None, sequential_sum=True):
    """voxel_helpers.py:570-598."""
    idx = intersection_outputs["intersected_voxel_idx"]
    dists = (intersection_outputs["max_depth"] - intersection_outputs["min_depth"]).astype(np.float32)
    dists[idx == -1] = 0
    tot = seq_sum(dists) if sequential_sum else dists.sum(-1, dtype=np.float32)
    probs = (dists / tot[:, None]).astype(np.float32)
    steps = (tot / np.float32(step_size)).astype(np.float32)
    if tot.max() > 10 * MAX_DEPTH:
        return None
    si, sd, sl = inverse_cdf_sampling(idx, intersection_outputs["min_depth"], intersection_outputs["max_depth"],
                                      probs, steps, -1, fixed, noise=noise)
    sl = np.maximum(sl, np.float32(0))
    sd = sd.copy()
    sd[si == -1] = MAX_DEPTH
    sl[si == -1] = 0.0
    return {"sampled_point_depth": sd, "sampled_point_distance": sl, "sampled_point_voxel_idx": si,
            "probs": probs, "steps": steps}
