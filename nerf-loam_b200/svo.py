"""`Octree` with the method set of the reference's torch.classes.svo.Octree
(third_party/sparse_octree/src/bindings.cpp:11-31), backed by the flat-array host octree of
libnerfloam_b200.so (csrc/octree_host.cpp).

Differences from the reference class, all deliberate:
  * node ids are per-tree (the reference's counter is a process global, octree.h:62, so a second
    instance corrupts its exports);
  * `get_centres_and_children()` is a direct array pass (no per-node torch dispatch);
  * `export_map()` returns the arrays already in the layout the hot path consumes (mapping.py:320-326);
  * pickling replays the inserted point tensors like the reference (bindings.cpp:23-31) and is wired.
"""
import ctypes as C

import numpy as np
import torch

from . import _capi


def encode(x, y, z):
    """svo.encode (bindings.cpp:6): 63-bit Morton key of integer voxel coordinates (utils.h:106-109)."""
    return int(_capi.lib().nl_morton_encode(int(x), int(y), int(z)))


class Octree:
    def __init__(self):
        self._h = None
        self._all_pts = []
        self._init_args = None

    # -- bindings.cpp:14 init(grid_dim, feat_dim, voxel_size)
    def init(self, grid_dim, feat_dim, voxel_size):
        if self._h is not None:
            _capi.lib().nl_octree_destroy(self._h)
        self._h = _capi.lib().nl_octree_create(int(grid_dim), int(feat_dim), float(voxel_size))
        if not self._h:
            raise _capi.NerfLoamError(_capi.lib().nl_last_error().decode())
        self._init_args = (int(grid_dim), int(feat_dim), float(voxel_size))
        self._all_pts = []

    def _check(self):
        if self._h is None:
            raise RuntimeError("Octree not initialized!")  # octree.cpp:56-59 prints this and carries on

    @staticmethod
    def _as_i32(vox):
        if isinstance(vox, torch.Tensor):
            if vox.dtype != torch.int32:
                raise RuntimeError("expected an int32 tensor (the reference accessor<int,2> throws here)")
            vox = vox.detach().cpu().contiguous().numpy()
        vox = np.ascontiguousarray(vox, dtype=np.int32)
        return vox

    # -- bindings.cpp:15 insert(Tensor int32[N,3])
    def insert(self, vox):
        self._check()
        v = self._as_i32(vox)
        if v.ndim != 2 or v.shape[1] != 3:
            print(f"Point dimensions mismatch: inputs are {v.shape[-1]} expect 3")  # octree.cpp:62-66
            return
        _capi.check(_capi.lib().nl_octree_insert(self._h, v.ctypes.data_as(C.c_void_p), v.shape[0]), "nl_octree_insert")
        self._all_pts.append(torch.from_numpy(v.copy()))

    def try_insert(self, vox):
        self._check()
        v = self._as_i32(vox)
        if v.ndim != 2 or v.shape[1] != 3:
            return -1.0
        return float(_capi.lib().nl_octree_try_insert(self._h, v.ctypes.data_as(C.c_void_p), v.shape[0]))

    def count_nodes(self):
        self._check()
        return int(_capi.lib().nl_octree_count_nodes(self._h))

    def count_leaf_nodes(self):
        self._check()
        return int(_capi.lib().nl_octree_count_leaf_nodes(self._h))

    def has_voxel(self, xyz):
        self._check()
        v = self._as_i32(xyz).reshape(-1)
        if v.shape[0] != 3:
            return False
        return bool(_capi.lib().nl_octree_has_voxel(self._h, v.ctypes.data_as(C.c_void_p)))

    def get_features(self, pts):
        """Empty body in the reference (octree.cpp:208-210)."""
        return None

    def get_voxels(self):
        self._check()
        n = self.count_nodes()
        out = np.empty((n, 4), np.float32)
        rows = _capi.lib().nl_octree_get_voxels(self._h, out.ctypes.data_as(C.c_void_p), n)
        return torch.from_numpy(out[:rows])

    def get_leaf_voxels(self):
        self._check()
        n = self.count_leaf_nodes()
        out = np.empty((max(n, 1), 3), np.float32)
        rows = _capi.lib().nl_octree_get_leaf_voxels(self._h, out.ctypes.data_as(C.c_void_p), n)
        return torch.from_numpy(out[:rows])

    # -- bindings.cpp:22 get_centres_and_children() -> (voxels f32[n,4], children f32[n,8], features i32[n,8])
    def get_centres_and_children(self):
        self._check()
        n = int(_capi.lib().nl_octree_count_export_nodes(self._h))
        voxels = np.empty((n, 4), np.float32)
        children = np.empty((n, 8), np.float32)
        features = np.empty((n, 8), np.int32)
        _capi.check(_capi.lib().nl_octree_export(self._h, voxels.ctypes.data_as(C.c_void_p), children.ctypes.data_as(C.c_void_p),
                                                 features.ctypes.data_as(C.c_void_p)), "nl_octree_export")
        return torch.from_numpy(voxels), torch.from_numpy(children), torch.from_numpy(features)

    def export_map(self):
        """(centres f32[n,3], structure i32[n,9], vertex i32[n,8]) == what mapping.py:320-326 derives."""
        self._check()
        n = int(_capi.lib().nl_octree_count_export_nodes(self._h))
        centres = np.empty((n, 3), np.float32)
        structure = np.empty((n, 9), np.int32)
        vertex = np.empty((n, 8), np.int32)
        _capi.check(_capi.lib().nl_octree_export_map(self._h, centres.ctypes.data_as(C.c_void_p),
                                                     structure.ctypes.data_as(C.c_void_p), vertex.ctypes.data_as(C.c_void_p)),
                    "nl_octree_export_map")
        return torch.from_numpy(centres), torch.from_numpy(structure), torch.from_numpy(vertex)

    def export_dirty(self, clear=True):
        """Incremental export: (ids i32[m] ascending, centres f32[m,3], structure i32[m,9], vertex i32[m,8]) of the rows that
        changed since the last clearing call -- scattering them into the previous export gives the full export bit for bit."""
        self._check()
        m = int(_capi.lib().nl_octree_dirty_count(self._h))
        ids = np.empty(m, np.int32)
        centres = np.empty((m, 3), np.float32)
        structure = np.empty((m, 9), np.int32)
        vertex = np.empty((m, 8), np.int32)
        _capi.check(_capi.lib().nl_octree_export_dirty(self._h, ids.ctypes.data_as(C.c_void_p), centres.ctypes.data_as(C.c_void_p),
                                                       structure.ctypes.data_as(C.c_void_p), vertex.ctypes.data_as(C.c_void_p), int(bool(clear))),
                    "nl_octree_export_dirty")
        return ids, centres, structure, vertex

    def count_export_nodes(self):
        self._check()
        return int(_capi.lib().nl_octree_count_export_nodes(self._h))

    # -- pickle (bindings.cpp:23-31): state = (size, feat_dim, voxel_size, all inserted tensors), replayed on load
    def __getstate__(self):
        return {"init": self._init_args, "all_pts": self._all_pts}

    def __setstate__(self, state):
        self._h = None
        self._all_pts = []
        self._init_args = None
        if state["init"] is not None:
            self.init(*state["init"])
            for p in state["all_pts"]:
                self.insert(p)

    def __del__(self):
        try:
            if self._h is not None:
                _capi.lib().nl_octree_destroy(self._h)
                self._h = None
        except Exception:
            pass
