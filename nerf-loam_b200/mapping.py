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
    """Device-resident map, updated incrementally.

    All hot-path arrays -- centres, structure, vox2row, the packed traversal image, the bf16 table -- live in capacity-doubling
    DEVICE buffers whose base addresses stay put between doublings.  A map update uploads only the rows the insertion touched
    (svo.Octree.export_dirty: new nodes, nodes that gained a child, leaves that turned SURFACE and their parents), scatters them
    in place, numbers the new vertices and re-packs the touched nodes; the reference re-exports the whole octree with a torch
    dispatcher call per node and re-uploads everything every frame (octree.cpp:293-342, mapping.py:309-337).  Stable addresses are
    also what lets the captured CUDA graphs of the optimisation loops (render_helpers._MapGraph / _TrackGraph) survive map updates.
    `incremental=False` keeps the full-export path (used by the tests as the cross-check)."""

    def __init__(self, voxel_size, embed_dim=16, grid_dim=256 * 256 * 4, device="cuda", init_std=0.0, seed=0, incremental=True,
                 reserve_nodes=0, reserve_rows=0):
        assert embed_dim == 16
        self.voxel_size = float(voxel_size)
        self.device = torch.device(device)
        self.svo = Octree()
        self.svo.init(grid_dim, embed_dim, voxel_size)     # mapping.py:81-82
        self.vertex2row = np.full(0, -1, np.int32)
        self.n_rows = 0
        self.n_nodes = 0
        self._emb_buf = torch.zeros((1024, embed_dim), dtype=torch.bfloat16, device=self.device)
        self.init_std = init_std
        self._gen = torch.Generator(device="cpu").manual_seed(seed)
        self.map_states = None
        self.incremental = bool(incremental) and self.device.type == "cuda"
        self._cap = 0
        self._centres = self._structure = self._vox2row = self._packed = None
        self._host = None            # host mirrors (centres, structure, vertex) for the reference-format dict
        self.generation = 0
        self.last_update = {}
        self.readers = None          # optional share.SharedMap: in-place updates wait (on the stream) for its readers
        # reserve_*: initial capacities.  Every capacity doubling moves the buffers (one device copy) and invalidates captured graphs;
        # a run that knows its map size (a KITTI sequence at 0.3 m: ~10^6 nodes = 0.3 GB of the 180 GB) reserves it once
        if self.incremental and reserve_nodes:
            self._grow_nodes(int(reserve_nodes))
        if reserve_rows and reserve_rows > self._emb_buf.shape[0]:
            self._emb_buf = torch.zeros((int(reserve_rows), embed_dim), dtype=torch.bfloat16, device=self.device)

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

    # ---------------------------------------------------------------------------------------------------------------
    def _grow_rows(self, new_rows):
        if new_rows > self._emb_buf.shape[0]:                                  # grow in place (mapping.py:309-314 re-uploads everything)
            cap = max(new_rows, 2 * self._emb_buf.shape[0])
            buf = torch.zeros((cap, 16), dtype=torch.bfloat16, device=self.device)
            buf[:self.n_rows] = self._emb_buf[:self.n_rows]
            self._emb_buf = buf
        if self.init_std > 0 and new_rows > self.n_rows:                       # reference: zeros (mapping.py:305-307)
            add = torch.randn((new_rows - self.n_rows, 16), generator=self._gen) * self.init_std
            self._emb_buf[self.n_rows:new_rows] = add.to(torch.bfloat16).to(self.device)
        self.n_rows = new_rows

    def _grow_nodes(self, n):
        if n <= self._cap:
            return
        cap = max(n, 2 * self._cap, 4096)
        d = self.device
        new = (torch.zeros((cap, 3), dtype=torch.float32, device=d), torch.full((cap, 9), -1, dtype=torch.int32, device=d),
               torch.full((cap, 8), -1, dtype=torch.int32, device=d), torch.zeros(cap * 128, dtype=torch.uint8, device=d))
        if self._cap:
            k = self.n_nodes
            new[0][:k] = self._centres[:k]; new[1][:k] = self._structure[:k]; new[2][:k] = self._vox2row[:k]
            new[3][:k * 128] = self._packed[:k * 128]
        self._centres, self._structure, self._vox2row, self._packed = new
        self._cap = cap

    def update_grid_features(self):
        """mapping.py:320-339.  Returns the engine.MapState (also stored, with the reference's dict keys, in self.map_states)."""
        return self._update_incremental() if self.incremental else self._update_full()

    def _update_incremental(self):
        ids, c, s, v = self.svo.export_dirty()
        n = self.svo.count_export_nodes()
        m = ids.shape[0]
        if self.vertex2row.shape[0] < n:
            self.vertex2row = np.concatenate([self.vertex2row, np.full(n - self.vertex2row.shape[0], -1, np.int32)])
        rows = np.empty((m, 8), np.int32)
        new_rows = int(_capi.lib().nl_assign_embedding_rows_subset(v.ctypes.data_as(C.c_void_p), m, n, self.vertex2row.ctypes.data_as(C.c_void_p),
                                                                   self.n_rows, rows.ctypes.data_as(C.c_void_p)))
        if new_rows < 0:
            raise _capi.NerfLoamError(_capi.lib().nl_last_error().decode())
        self._grow_rows(new_rows)
        self._grow_nodes(n)
        if self.readers is not None:      # a tracker on another stream may still traverse the arrays that are patched below
            for ev in self.readers.reader_events():
                torch.cuda.current_stream(self.device).wait_event(ev)
        if m:
            d = self.device
            idx = torch.from_numpy(ids).to(d, non_blocking=True)
            idl = idx.long()
            self._centres.index_copy_(0, idl, torch.from_numpy(c).to(d, non_blocking=True))
            self._structure.index_copy_(0, idl, torch.from_numpy(s).to(d, non_blocking=True))
            self._vox2row.index_copy_(0, idl, torch.from_numpy(rows).to(d, non_blocking=True))
            _capi.check(_capi.lib().nl_octree_pack_children_rows(m, _capi.ptr(idx), _capi.ptr(self._centres), _capi.ptr(self._structure),
                                                                 _capi.ptr(self._packed), _capi.stream_ptr()), "nl_octree_pack_children_rows")
            _capi.LAUNCHES += 1
        # host mirrors for the reference-format dict (CPU tensors like mapping.py:321-337), patched in place as well
        if self._host is None or self._host[0].shape[0] < n:
            cap = max(n, 2 * (self._host[0].shape[0] if self._host is not None else 0))
            hc, hs, hv = np.zeros((cap, 3), np.float32), np.full((cap, 9), -1, np.int32), np.full((cap, 8), -1, np.int32)
            if self._host is not None:
                k = self.n_nodes
                hc[:k], hs[:k], hv[:k] = self._host[0][:k], self._host[1][:k], self._host[2][:k]
            self._host = (hc, hs, hv)
        self._host[0][ids], self._host[1][ids], self._host[2][ids] = c, s, v
        self.n_nodes = n
        self.generation += 1
        self.last_update = {"dirty_rows": int(m), "nodes": int(n), "embedding_rows": int(new_rows)}
        ms = MapState.__new__(MapState)
        ms.centres, ms.structure, ms.vox2row = self._centres[:n], self._structure[:n], self._vox2row[:n]
        ms.emb = self.embeddings
        ms.emb_full = self._emb_buf              # capacity-sized table at a stable address (rows >= n_rows: zero, never referenced)
        ms.n_nodes = n
        ms._packed = self._packed[:n * 128]
        ms.stable = True                          # base addresses survive map updates until a capacity doubling
        centres_h, structure_h, vertex_h = (torch.from_numpy(a[:n]) for a in self._host)
        id2 = torch.from_numpy(self.vertex2row[:n]).view(-1, 1)
        self.map_states = {"voxel_vertex_idx": vertex_h, "voxel_center_xyz": centres_h, "voxel_structure": structure_h,
                           "voxel_vertex_emb": self.embeddings, "voxel_id2embedding_id": id2, "_mapstate": ms}
        return ms

    def _update_full(self):
        self.svo.export_dirty()                                                # keep the dirty list from growing without bound
        centres, structure, vertex = self.svo.export_map()                     # mapping.py:321-326
        n = centres.shape[0]
        if self.vertex2row.shape[0] < n:
            self.vertex2row = np.concatenate([self.vertex2row, np.full(n - self.vertex2row.shape[0], -1, np.int32)])
        v = vertex.numpy()
        new_rows = int(_capi.lib().nl_assign_embedding_rows(v.ctypes.data_as(C.c_void_p), n,
                                                            self.vertex2row.ctypes.data_as(C.c_void_p), self.n_rows))
        if new_rows < 0:
            raise _capi.NerfLoamError(_capi.lib().nl_last_error().decode())
        self._grow_rows(new_rows)
        self.n_nodes = n
        self.generation += 1
        vox2row = np.where(v >= 0, self.vertex2row[np.clip(v, 0, None)], -1).astype(np.int32)
        ms = MapState(centres, structure, torch.from_numpy(vox2row), self.embeddings, self.device)
        id2 = torch.from_numpy(self.vertex2row[:n].copy()).view(-1, 1)
        self.map_states = {"voxel_vertex_idx": vertex, "voxel_center_xyz": centres, "voxel_structure": structure,
                           "voxel_vertex_emb": self.embeddings, "voxel_id2embedding_id": id2, "_mapstate": ms}
        return ms
