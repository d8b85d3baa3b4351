"""CPU-only tests of the product's host side: the C-ABI library loads and exports every declared symbol,
the host octree is bit-exact against the reference goldens, map-update glue, API surface, and the rule that
the product never touches oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from util import ROOT, golden


@pytest.fixture(scope="module")
def nl():
    import nerfloam_b200 as nl
    return nl


def test_library_exports_every_declared_symbol(nl):
    hdr = open(os.path.join(ROOT, "include", "nerfloam_b200.h")).read()
    declared = set(re.findall(r"NL_API [^;(]*?\b(nl_[a-z0-9_]+)\(", hdr))
    assert len(declared) >= 30
    L = ctypes.CDLL(nl._capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/nerfloam_b200.h but not exported"
    assert set(nl._capi.EXPORTED_SYMBOLS) == declared
    assert nl._capi.lib().nl_version() == 100
    assert ctypes.sizeof(nl._capi.RenderStats) == 160 and nl._capi.RenderStats.n_samples.offset == 12


def test_ctypes_signatures_match_the_header(nl):
    """Every prototype in the header has the same number of parameters as the ctypes binding (catches ABI drift between
    include/nerfloam_b200.h, the .cu/.cpp definitions behind it and _capi.py)."""
    hdr = open(os.path.join(ROOT, "include", "nerfloam_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = dict(re.findall(r"NL_API [^;(]*?\b(nl_[a-z0-9_]+)\(([^;]*?)\);", hdr, flags=re.S))
    assert set(protos) == set(nl._capi._SIGNATURES)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(nl._capi._SIGNATURES[name][1]), (name, n, len(nl._capi._SIGNATURES[name][1]))
    sizes = (ctypes.c_int32 * 4)()
    nl._capi.lib().nl_abi_sizes(ctypes.byref(sizes))
    assert list(sizes)[:3] == [ctypes.sizeof(nl._capi.RenderStats), nl._capi.RenderStats.n_samples.offset, ctypes.sizeof(nl._capi.RenderArgs)]


def test_octree_bit_exact_vs_reference_golden(nl):
    z = golden("octree.npz")
    o = nl.svo.Octree()
    o.init(256 * 256 * 4, 16, 0.3)
    o.insert(torch.from_numpy(z["v1"]))
    v, c, f = o.get_centres_and_children()
    assert v.dtype == torch.float32 and c.dtype == torch.float32 and f.dtype == torch.int32
    assert np.array_equal(v.numpy(), z["voxels1"]) and np.array_equal(c.numpy(), z["children1"]) and np.array_equal(f.numpy(), z["features1"])
    assert [o.count_nodes(), o.count_leaf_nodes()] == list(z["count1"])
    o.insert(torch.from_numpy(z["v2"]))                                  # incremental insert keeps creation-order ids
    v, c, f = o.get_centres_and_children()
    assert np.array_equal(v.numpy(), z["voxels2"]) and np.array_equal(c.numpy(), z["children2"]) and np.array_equal(f.numpy(), z["features2"])
    assert [o.count_nodes(), o.count_leaf_nodes()] == list(z["count2"])
    # hot-path layout == the reference's Python glue (mapping.py:322-326) applied to the reference export
    ce, st, ve = o.export_map()
    assert torch.equal(ce, ((v[:, :3] + v[:, -1:] / 2) * 0.3).float())
    assert torch.equal(st, torch.cat([c, v[:, -1:]], -1).int()) and torch.equal(ve, f)


def test_octree_8cubed_config0_and_queries(nl):
    z = golden("octree_8.npz")
    o = nl.svo.Octree()
    o.init(8, 16, 1.0)
    o.insert(torch.from_numpy(z["v0"]))
    v, c, f = o.get_centres_and_children()
    assert np.array_equal(v.numpy(), z["voxels"]) and np.array_equal(c.numpy(), z["children"]) and np.array_equal(f.numpy(), z["features"])
    assert [o.count_nodes(), o.count_leaf_nodes()] == list(z["count"])
    p = z["v0"][0]
    assert o.has_voxel(torch.tensor(p, dtype=torch.int32)) and o.has_voxel(torch.tensor(p + 1, dtype=torch.int32))
    lv = o.get_leaf_voxels().numpy()
    assert lv.shape == (o.count_leaf_nodes(), 3)
    assert set(map(tuple, lv.astype(int))) == set(map(tuple, np.unique(z["v0"], axis=0)))
    assert o.get_voxels().shape == (o.count_nodes(), 4)
    assert o.try_insert(torch.from_numpy(z["v0"])) == 1.0
    far = torch.tensor([[5, 5, 5]], dtype=torch.int32)
    assert 0.0 <= o.try_insert(far) <= 1.0
    with pytest.raises(nl._capi.NerfLoamError):
        o.insert(torch.tensor([[7, 0, 0]], dtype=torch.int32))           # corner x+1 = 8 leaves the grid
    with pytest.raises(RuntimeError):
        o.insert(torch.zeros((2, 3), dtype=torch.int64))                 # wrong dtype, like accessor<int,2>


def test_octree_edge_cases_and_pickle(nl):
    import pickle
    o = nl.svo.Octree()
    with pytest.raises(RuntimeError):
        o.count_nodes()                                                  # not initialised
    o.init(64, 16, 0.5)
    assert o.count_nodes() == 1 and o.count_leaf_nodes() == 0
    v, c, f = o.get_centres_and_children()                               # empty tree: just the root
    assert v.tolist() == [[0.0, 0.0, 0.0, 64.0]] and (c == -1).all() and (f == -1).all()
    o.insert(torch.zeros((0, 3), dtype=torch.int32))                     # empty insert is a no-op
    pts = torch.tensor([[3, 4, 5], [3, 4, 5], [3, 4, 6]], dtype=torch.int32)   # duplicates + shared corners
    o.insert(pts)
    n1 = o.count_nodes()
    o.insert(pts)                                                        # re-inserting changes nothing
    assert o.count_nodes() == n1 and o.count_leaf_nodes() == 2
    o2 = pickle.loads(pickle.dumps(o))
    for a, b in zip(o.get_centres_and_children(), o2.get_centres_and_children()):
        assert torch.equal(a, b)
    assert nl.svo.encode(1, 2, 3) == nl.svo.encode(1, 2, 3) and nl.svo.encode(1, 0, 0) == 1 and nl.svo.encode(0, 1, 0) == 2


def test_octree_matches_oracle_on_random_trees(nl):
    from oracle import kernels as OK
    rng = np.random.default_rng(3)
    for trial in range(5):
        size = [16, 64, 1024][trial % 3]
        pts = rng.integers(0, size - 1, size=(int(rng.integers(1, 400)), 3)).astype(np.int32)
        a = nl.svo.Octree(); a.init(size, 16, 0.25)
        b = OK.Octree(); b.init(size, 16, 0.25)
        for part in np.array_split(pts, 3):
            a.insert(torch.from_numpy(np.ascontiguousarray(part))); b.insert(part)
        va, ca, fa = a.get_centres_and_children()
        vb, cb, fb = b.get_centres_and_children()
        assert np.array_equal(va.numpy(), vb) and np.array_equal(ca.numpy(), cb) and np.array_equal(fa.numpy(), fb)


def test_map_updater_rows_and_layout(nl):
    z = golden("octree.npz")
    mu = nl.mapping.MapUpdater(0.3, device="cpu")
    ms = mu.insert_voxels(torch.from_numpy(z["v1"]))
    vertex = mu.map_states["voxel_vertex_idx"]
    flat = vertex.reshape(-1)
    used = flat[flat >= 0]
    assert mu.n_rows == len(torch.unique(used))                          # one row per distinct vertex
    first = {}
    for v in used.tolist():
        first.setdefault(v, len(first))
    assert all(mu.vertex2row[v] == r for v, r in first.items())          # numbered by first appearance (mapping.py:296-317)
    rows_before = mu.vertex2row.copy()
    n_before = mu.n_rows
    mu.embeddings[:] = 0.5
    mu.insert_voxels(torch.from_numpy(z["v2"]))                          # growth keeps existing rows and values
    assert np.array_equal(mu.vertex2row[:len(rows_before)][rows_before >= 0], rows_before[rows_before >= 0])
    assert mu.n_rows > n_before and float(mu.embeddings[:n_before].float().min()) == 0.5
    assert float(mu.embeddings[n_before:].float().abs().max()) == 0.0    # new rows are zero like mapping.py:305
    ms = mu.map_states["_mapstate"]
    surf = (mu.map_states["voxel_structure"][:, 8] == 1) & (mu.map_states["voxel_vertex_idx"][:, 0] >= 0)
    assert bool((ms.vox2row[surf] >= 0).all()) and int(ms.vox2row.max()) == mu.n_rows - 1
    assert set(mu.map_states) >= {"voxel_vertex_idx", "voxel_center_xyz", "voxel_structure", "voxel_vertex_emb", "voxel_id2embedding_id"}


def test_api_surface_matches_reference(nl):
    import inspect
    rh = nl.render_helpers
    ref_sigs = {
        "render_rays": ["rays_o", "rays_d", "map_states", "sdf_network", "step_size", "voxel_size", "truncation", "max_voxel_hit",
                        "max_distance", "chunk_size", "profiler", "return_raw"],
        "bundle_adjust_frames": ["keyframe_graph", "embeddings", "map_states", "sdf_network", "loss_criteria", "voxel_size", "step_size",
                                 "N_rays", "num_iterations", "truncation", "max_voxel_hit", "max_distance", "learning_rate",
                                 "update_pose", "update_decoder", "profiler"],
        "track_frame": ["frame_pose", "curr_frame", "map_states", "sdf_network", "loss_criteria", "voxel_size", "N_rays", "step_size",
                        "num_iterations", "truncation", "learning_rate", "max_voxel_hit", "max_distance", "profiler", "depth_variance"],
        "get_scores": ["sdf_network", "map_states", "voxel_size", "bits"],
    }
    for name, params in ref_sigs.items():
        got = list(inspect.signature(getattr(rh, name)).parameters)
        assert got[:len(params)] == params, name
    dec = nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0)
    assert list(dec.state_dict()) == ["pts_linears.0.weight", "pts_linears.0.bias", "pts_linears.1.weight", "pts_linears.1.bias",
                                      "sdf_out.weight", "sdf_out.bias"]
    assert sum(p.numel() for p in dec.parameters()) == 70401
    with pytest.raises(RuntimeError):
        dec(torch.zeros(4, 16))                                          # no CPU fallback
    with pytest.raises(NotImplementedError):
        nl.lidar.Decoder(depth=8, width=256, in_dim=16, skips=[4], embedder="none")
    for fn in ("svo_intersect", "inverse_cdf_sampling", "ball_intersect", "aabb_intersect", "triangle_intersect",
               "uniform_ray_sampling", "build_octree"):
        assert callable(getattr(nl.grid, fn))
    with pytest.raises(RuntimeError):
        nl.grid.svo_intersect(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3), torch.zeros(1, 1, 3), torch.zeros(1, 1, 9, dtype=torch.int32), 0.3, 20)


def test_pose_and_frame_host_api(nl):
    z = golden("pose.npz")
    P = nl.se3pose.OptimizablePose.from_matrix(torch.from_numpy(z["before"]))
    np.testing.assert_allclose(P.data.detach().numpy(), z["data"], atol=1e-6)
    np.testing.assert_allclose(P.rotation().detach().numpy(), z["R"], atol=1e-6)
    np.testing.assert_allclose(P.matrix().detach().numpy(), z["after"], atol=1e-6)
    torch.manual_seed(5)
    pts = torch.randn(500, 3) * 10
    f = nl.frame.LidarFrame(3, pts, torch.ones(500), np.eye(4))
    assert float(f.pose.data[0]) == 2000.0 and f.rays_d.shape == (500, 1, 3)
    f.sample_rays(64)
    assert f.sample_mask.shape == (500, 1) and int(f.sample_mask.sum()) == 64
    # same seed -> same rays as the reference's selection (sample_util.py:4-19 restated literally: Gumbel top-k on log-probabilities)
    def reference_mask(n, k):
        mask = torch.ones((n, 1))[None, ...]
        probs = (mask / (mask.sum() + 1e-9)).reshape(1, -1)
        logp = torch.log(probs + 1e-9)
        scores = logp + (-torch.log(-torch.log(torch.rand_like(logp) + 1e-7) + 1e-7))
        idx = scores.topk(k, dim=-1)[1]
        return (torch.zeros_like(probs).scatter_(-1, idx, 1).reshape(1, n, 1) > 0)[0, ...]
    for seed in (0, 7):
        torch.manual_seed(seed)
        ref = reference_mask(500, 64)
        torch.manual_seed(seed)
        f.sample_rays(64)
        assert torch.equal(f.sample_mask, ref)
    np.testing.assert_array_equal(f.rays_norm.numpy(), (torch.norm(pts, 2, -1, keepdim=True) + 1e-8).numpy())


def test_criterion_host_forward_vs_reference_golden(nl):
    from util import Args
    z = golden("criterion.npz")
    valid = torch.from_numpy(z["valid"])
    sv = torch.from_numpy(z["sdf_valid"]).requires_grad_()
    sdf = torch.ones(valid.shape).masked_scatter(valid, sv)
    crit = nl.criterion.Criterion(Args())
    loss, d = crit({"sdf": sdf, "z_vals": torch.from_numpy(z["z"]), "ray_mask": torch.from_numpy(z["ray_mask"]), "valid_mask": valid,
                    "sampled_xyz": None}, torch.from_numpy(z["points"]), torch.from_numpy(z["cos"]))
    np.testing.assert_allclose(float(loss), float(z["loss"]), rtol=1e-6)
    np.testing.assert_allclose(torch.autograd.grad(loss, sv)[0].numpy(), z["grad_sdf"], rtol=1e-5, atol=1e-7)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under nerf-loam_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "nerf-loam_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, fn), errors="ignore").read()
                assert "oracle" not in src.replace("oracle/ is", "").lower() or fn == "__init__.py" and False, f"{fn} mentions oracle"


def test_torchscript_svo_shim_is_a_drop_in_for_the_reference_extension(nl):
    """`torch.classes.load_library(<svo_b200.so>)` + `torch.classes.svo.Octree()` -- the two statements of src/mapping.py:19-20,81 --
    give an object with the reference's methods whose exports are bit-identical to the executed reference (goldens)."""
    import io
    import sys
    sys.path.insert(0, os.path.join(ROOT, "nerf-loam_b200"))
    import build as nlbuild
    torch.classes.load_library(nlbuild.build_svo_shim())
    z = golden("octree.npz")
    o = torch.classes.svo.Octree()
    o.init(256 * 256 * 4, 16, 0.3)
    o.insert(torch.from_numpy(z["v1"]))
    v, c, f = o.get_centres_and_children()
    assert np.array_equal(v.numpy(), z["voxels1"]) and np.array_equal(c.numpy(), z["children1"]) and np.array_equal(f.numpy(), z["features1"])
    assert [o.count_nodes(), o.count_leaf_nodes()] == list(z["count1"])
    o.insert(torch.from_numpy(z["v2"]))
    v, c, f = o.get_centres_and_children()
    assert np.array_equal(v.numpy(), z["voxels2"]) and np.array_equal(c.numpy(), z["children2"]) and np.array_equal(f.numpy(), z["features2"])
    assert [o.count_nodes(), o.count_leaf_nodes()] == list(z["count2"])
    p = nl.svo.Octree()
    p.init(256 * 256 * 4, 16, 0.3)
    p.insert(torch.from_numpy(z["v1"])); p.insert(torch.from_numpy(z["v2"]))
    assert torch.equal(o.get_voxels(), p.get_voxels()) and torch.equal(o.get_leaf_voxels(), p.get_leaf_voxels())
    pt = torch.from_numpy(z["v1"][0])
    assert o.has_voxel(pt) and not o.has_voxel(torch.tensor([1, 1, 1], dtype=torch.int32))
    assert o.try_insert(torch.from_numpy(z["v1"][:10])) == p.try_insert(torch.from_numpy(z["v1"][:10]))
    assert torch.ops.svo.encode(torch.tensor([3, 5, 7], dtype=torch.int32)) == nl.svo.encode(3, 5, 7)
    buf = io.BytesIO()                                                   # def_pickle: state = (size, feat_dim, voxel_size, inserted tensors)
    torch.save(o, buf)
    buf.seek(0)
    o2 = torch.load(buf, weights_only=False)
    v2, c2, f2 = o2.get_centres_and_children()
    assert torch.equal(v2, v) and torch.equal(c2, c) and torch.equal(f2, f)


def test_c_abi_rejects_bad_arguments_without_touching_the_gpu(nl):
    """Every entry point validates its arguments before the first CUDA call and reports through a status + nl_last_error()
    (the reference's kernels exit(-1) on errors, cuda_utils.h:37-48; its CHECK_* macros raise on bad tensors)."""
    L = nl._capi.lib()
    err = lambda: L.nl_last_error().decode()
    assert L.nl_gather_trilinear_fwd(-1, None, None, None, None, None, None, 0.3, None, None) != 0 and "negative" in err()
    assert L.nl_gather_trilinear_fwd(5, None, None, None, None, None, None, 0.3, None, None) != 0 and "null" in err()
    assert L.nl_gather_trilinear_fwd(0, None, None, None, None, None, None, 0.3, None, None) == 0          # empty input is fine
    assert L.nl_mlp_tc_forward(7, None, None, None, None, None, None, None, None, None) != 0 and "null" in err()
    assert L.nl_mlp_tc_forward(0, None, None, None, None, None, None, None, None, None) == 0
    assert L.nl_adam_f32(3, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 1, None) != 0 and "null" in err()
    assert L.nl_adam_f32(3, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 0, None) != 0                 # Adam steps start at 1
    assert L.nl_octree_pack_children(0, None, None, None, None) != 0 and "positive" in err()
    assert L.nl_octree_pack_children(1 << 26, None, None, None, None) != 0 and "2^26" in err()
    assert L.nl_render_samples(None, None) != 0 and "null" in err()
    a = nl._capi.RenderArgs()
    a.n_rays, a.n_nodes = 0, 5
    assert L.nl_render_samples(ctypes.byref(a), None) != 0 and "positive" in err()
    assert L.nl_render_workspace_bytes(-1) == -1 and L.nl_render_workspace_bytes(1000) > 1000 * 240
    assert L.nl_octree_create(0, 16, 0.3) is None and err() != ""                                        # grid_dim must be positive
    with pytest.raises(nl._capi.NerfLoamError):
        nl._capi.check(L.nl_svo_intersect(1, 1, 1, 0.3, 20, None, None, None, None, None, None, None, None), "nl_svo_intersect")
    # round-2 entry points
    assert L.nl_select_rays(0, 100, 10, None, None, 1, None, None, None, None, None, None, None, None) != 0 and "positive" in err()
    assert L.nl_select_rays(1, 100, 10, None, None, 1, None, None, None, None, None, None, None, None) != 0 and "null" in err()
    assert L.nl_peer_reduce_adam_bf16(32, 2, 2, None, None, None, None, None, None, 1e-2, 0.9, 0.999, 1e-8, None, 0, None, None, None, None) != 0 and "sizes" in err()
    assert L.nl_peer_reduce_adam_bf16(24, 0, 2, None, None, None, None, None, None, 1e-2, 0.9, 0.999, 1e-8, None, 0, None, None, None, None) != 0   # whole rows of 16
    assert L.nl_peer_reduce_adam_bf16(32, 0, 2, None, None, None, None, None, None, 1e-2, 0.9, 0.999, 1e-8, None, 6, None, None, None, None) != 0 and "header" in err()
    assert L.nl_peer_reduce_adam_bf16(32, 0, 2, None, None, None, None, None, None, 1e-2, 0.9, 0.999, 1e-8, None, 0, None, None, None, None) != 0 and "null" in err()


def test_incremental_octree_export_equals_full_export():
    """nl_octree_export_dirty: scattering only the rows an insertion touched into the previous export reproduces the
    full export bit for bit, and the incremental vertex numbering equals the full-pass numbering."""
    import ctypes as C
    import nerfloam_b200 as nl
    syn = nl.synthetic
    o = nl.svo.Octree(); o.init(256 * 256 * 4, 16, 0.3)
    acc = [np.zeros((0, 3), np.float32), np.zeros((0, 9), np.int32), np.zeros((0, 8), np.int32)]
    v2r_inc, v2r_full, rows_inc, rows_full = np.full(0, -1, np.int32), np.full(0, -1, np.int32), 0, 0
    touched = []
    for i in range(4):
        pts, cos, pose = syn.make_scan(n_beams=16, n_az=150, seed=50 + i, sensor_xyz=(1.0 * i, 0.2 * i, 0.0))
        o.insert(torch.from_numpy(syn.voxelize(pts, pose, 0.3)))
        ids, c, s, v = o.export_dirty()
        n = o.count_export_nodes()
        assert np.all(np.diff(ids) > 0) and o.export_dirty()[0].shape[0] == 0            # ascending, and cleared by the call
        k = acc[0].shape[0]
        acc = [np.concatenate([acc[0], np.zeros((n - k, 3), np.float32)]), np.concatenate([acc[1], np.full((n - k, 9), -1, np.int32)]),
               np.concatenate([acc[2], np.full((n - k, 8), -1, np.int32)])]
        acc[0][ids], acc[1][ids], acc[2][ids] = c, s, v
        fc, fs, fv = [t.numpy() for t in o.export_map()]
        assert np.array_equal(acc[0], fc) and np.array_equal(acc[1], fs) and np.array_equal(acc[2], fv)
        touched.append(ids.shape[0])
        v2r_inc = np.concatenate([v2r_inc, np.full(n - v2r_inc.shape[0], -1, np.int32)])
        v2r_full = np.concatenate([v2r_full, np.full(n - v2r_full.shape[0], -1, np.int32)])
        sub = np.empty((ids.shape[0], 8), np.int32)
        rows_inc = nl._capi.lib().nl_assign_embedding_rows_subset(v.ctypes.data_as(C.c_void_p), ids.shape[0], n, v2r_inc.ctypes.data_as(C.c_void_p),
                                                                  rows_inc, sub.ctypes.data_as(C.c_void_p))
        rows_full = nl._capi.lib().nl_assign_embedding_rows(fv.ctypes.data_as(C.c_void_p), n, v2r_full.ctypes.data_as(C.c_void_p), rows_full)
        assert rows_inc == rows_full and np.array_equal(v2r_inc, v2r_full)
        assert np.array_equal(sub, np.where(v >= 0, v2r_full[np.clip(v, 0, None)], -1))
    assert touched[0] == acc[0].shape[0] - sum(0 for _ in ()) or touched[0] > 0
    assert all(t < 0.6 * acc[0].shape[0] for t in touched[1:])                            # later scans touch a fraction of the tree


def test_frame_device_arrays_are_cached_per_frame_object_and_fill_the_graph_buffers(nl):
    """render_helpers._frame_arrays: one upload per LidarFrame object (weak side table, nothing attached to the frame), re-done when a
    tensor of the frame is replaced; the upload helpers of the captured tracking / mapping iterations copy those arrays into their
    static buffers (exercised here on CPU tensors: the same code runs on the device)."""
    import copy, gc, pickle
    from types import SimpleNamespace
    rh, syn = nl.render_helpers, nl.synthetic
    frames = []
    for i, (nb, na) in enumerate(((8, 50), (8, 40))):
        pts, cos, pose = syn.make_scan(n_beams=nb, n_az=na, seed=5 + i)
        frames.append(nl.frame.LidarFrame(i, torch.from_numpy(pts), torch.from_numpy(cos), nl.se3pose.OptimizablePose.from_matrix(torch.from_numpy(pose)),
                                          new_keyframe=True))
    f = frames[0]
    d, c, g = rh._frame_arrays(f, "cpu")
    n = f.points.shape[0]
    assert d.shape == (n, 3) and c.shape == (n,) and g.shape == (n,)
    torch.testing.assert_close(d, f.rays_d.reshape(-1, 3).float())
    torch.testing.assert_close(g, torch.norm(f.points.float(), 2, -1) * f.pointsCos.float().view(-1))       # criterion.py:30-32
    d2, c2, g2 = rh._frame_arrays(f, torch.device("cpu"))
    assert d2 is d and c2 is c and g2 is g                                   # cached
    assert not any(isinstance(v, tuple) and any(x is d for x in v) for v in vars(f).values())     # nothing attached to the frame
    assert len(pickle.dumps(f)) < 40 * n + 20000                             # ... so a pickled frame stays its own size
    f.points = f.points.clone()                                              # a replaced tensor invalidates the entry
    d3, _, g3 = rh._frame_arrays(f, "cpu")
    assert d3 is not d and torch.equal(g3, g)
    fb = rh._FrameBatch(frames, torch.device("cpu"))
    assert fb.dirs[0] is d3 and fb.gt[1].shape[0] == frames[1].points.shape[0]
    # the graphs' upload helpers (static buffers of capacity `cap`, ragged scans)
    cap = 512
    mg = SimpleNamespace(dirs=torch.zeros(2, cap, 3), cos=torch.zeros(2, cap), gt=torch.zeros(2, cap), n_dev=torch.zeros(2, 1, dtype=torch.int64))
    rh._MapGraph._upload(mg, frames)
    for i, fr in enumerate(frames):
        k = fr.points.shape[0]
        assert int(mg.n_dev[i]) == k and torch.equal(mg.dirs[i, :k], fb.dirs[i]) and torch.equal(mg.gt[i, :k], fb.gt[i]) and torch.equal(mg.cos[i, :k], fb.cos[i])
        assert float(mg.dirs[i, k:].abs().sum()) == 0.0
    tg = SimpleNamespace(dirs=torch.zeros(cap, 3), cos=torch.zeros(cap), gt=torch.zeros(cap), n_dev=torch.zeros(1, dtype=torch.int64))
    rh._TrackGraph._upload(tg, frames[1])
    k = frames[1].points.shape[0]
    assert int(tg.n_dev) == k and torch.equal(tg.dirs[:k], fb.dirs[1]) and torch.equal(tg.gt[:k], fb.gt[1])
    n_before = len(rh._FRAME_ARRAYS)
    del fb, frames, f, fr
    gc.collect()
    assert len(rh._FRAME_ARRAYS) < n_before                                  # entries die with their frames


def test_octree_membership_sets_under_growth_and_duplicates(nl):
    """The host octree answers "seen before?" through flat open-addressing key sets (csrc/octree_host.cpp, KeySet): a few hundred
    thousand points with heavy repetition, inserted in several calls, must give exactly one SURFACE leaf per distinct voxel, and
    try_insert must report the exact overlap of the corner sets."""
    rng = np.random.default_rng(7)
    o = nl.svo.Octree()
    o.init(1024, 16, 0.2)
    pts = rng.integers(0, 1022, size=(60_000, 3), dtype=np.int32)
    pts = pts[rng.integers(0, pts.shape[0], size=400_000)]                 # every voxel ~7 times, shuffled
    for part in np.array_split(pts, 5):
        o.insert(torch.from_numpy(np.ascontiguousarray(part)))
    uniq = np.unique(pts, axis=0)
    assert o.count_leaf_nodes() == uniq.shape[0]
    assert all(o.has_voxel(torch.from_numpy(v)) for v in uniq[:50])
    inside = torch.from_numpy(np.ascontiguousarray(uniq[:1000]))
    assert o.try_insert(inside) == 1.0                                      # every corner of known voxels is known
    far = torch.from_numpy(np.ascontiguousarray(uniq[:1000] % 7 + np.array([[0, 0, 0]], np.int32)))   # a 8^3 corner of the grid
    r = o.try_insert(far)
    assert 0.0 <= r <= 1.0

