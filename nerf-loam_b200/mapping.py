"""Map update glue between the octree and the optimisation loop:
Mapping.create_voxels / update_grid_features / get_embeddings of src/mapping.py:283-339.

What changes relative to the reference:
  * no 8 GB `voxel_id2embedding_id` CPU lookup (mapping.py:76): a compact i32[n_nodes] vertex->row table;
  * one embedding row per distinct vertex (the reference allocates a row per *reference* to a new vertex,
    ~70 % orphan rows); values are identical because new rows are zero-initialised;
  * the bf16 table grows in place on the device (capacity doubling) instead of CPU cat + full re-upload;
  * the octree export is already in hot-path layout (centres / structure / vertex) -- svo.Octree.export_map.
`map_states` keeps the reference's dict keys so the reference's tracker/mesher can read it, plus
"_mapstate" (engine.MapState) which the fused renderer uses directly.
"""
import ctypes as C

import numpy as np
import torch

from . import _capi
from .engine import MapState
from .svo import Octree


class MapUpdater:
    def __init__(self, voxel_size, embed_dim=16, grid_dim=256 * 256 * 4, device="cuda", init_std=0.0, seed=0):
        assert embed_dim == 16
        self.voxel_size = float(voxel_size)
        self.device = torch.device(device)
        self.svo = Octree()
        self.svo.init(grid_dim, embed_dim, voxel_size)     # mapping.py:81-82
        self.vertex2row = np.full(0, -1, np.int32)
        self.n_rows = 0
        self._emb_buf = torch.zeros((1024, embed_dim), dtype=torch.bfloat16, device=self.device)
        self.init_std = init_std
        self._gen = torch.Generator(device="cpu").manual_seed(seed)
        self.map_states = None

    @property
    def embeddings(self):
        return self._emb_buf[:self.n_rows]

    def create_voxels(self, points, pose):
        """mapping.py:283-291: sensor-frame points [N,3] + 4x4 pose -> voxel coordinates -> octree -> map_states."""
        pts = torch.as_tensor(points, dtype=torch.float32, device=self.device)
        T = torch.as_tensor(pose, dtype=torch.float32, device=self.device)
        world = pts @ T[:3, :3].transpose(-1, -2) + T[:3, 3]
        voxels = torch.div(world, self.voxel_size, rounding_mode="floor")
        self.svo.insert(voxels.cpu().int())
        return self.update_grid_features()

    def insert_voxels(self, voxels_i32):
        self.svo.insert(voxels_i32)
        return self.update_grid_features()

    def update_grid_features(self):
        centres, structure, vertex = self.svo.export_map()                     # mapping.py:321-326
        n = centres.shape[0]
        if self.vertex2row.shape[0] < n:
            self.vertex2row = np.concatenate([self.vertex2row, np.full(n - self.vertex2row.shape[0], -1, np.int32)])
        v = vertex.numpy()
        new_rows = int(_capi.lib().nl_assign_embedding_rows(v.ctypes.data_as(C.c_void_p), n,
                                                            self.vertex2row.ctypes.data_as(C.c_void_p), self.n_rows))
        if new_rows < 0:
            raise _capi.NerfLoamError(_capi.lib().nl_last_error().decode())
        if new_rows > self._emb_buf.shape[0]:                                  # grow in place (mapping.py:309-314 re-uploads everything)
            cap = max(new_rows, 2 * self._emb_buf.shape[0])
            buf = torch.zeros((cap, 16), dtype=torch.bfloat16, device=self.device)
            buf[:self.n_rows] = self._emb_buf[:self.n_rows]
            self._emb_buf = buf
        if self.init_std > 0 and new_rows > self.n_rows:                       # reference: zeros (mapping.py:305-307)
            add = torch.randn((new_rows - self.n_rows, 16), generator=self._gen) * self.init_std
            self._emb_buf[self.n_rows:new_rows] = add.to(torch.bfloat16).to(self.device)
        self.n_rows = new_rows
        vox2row = np.where(v >= 0, self.vertex2row[np.clip(v, 0, None)], -1).astype(np.int32)
        ms = MapState(centres, structure, torch.from_numpy(vox2row), self.embeddings, self.device)
        id2 = torch.from_numpy(self.vertex2row[:n].copy()).view(-1, 1)
        self.map_states = {"voxel_vertex_idx": vertex, "voxel_center_xyz": centres, "voxel_structure": structure,
                           "voxel_vertex_emb": self.embeddings, "voxel_id2embedding_id": id2, "_mapstate": ms}
        return ms
