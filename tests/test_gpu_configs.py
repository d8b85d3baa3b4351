"""GPU parity of the fused path against the oracle on the other BASELINE.json configurations (shapes of MaiCity /
KITTI-incremental / Newer College, and the 8^3 / 2x32 plumbing case), plus edge cases of the reference's behaviour."""
import numpy as np
import pytest
import torch

from util import Args

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nl():
    import nerfloam_b200 as nl
    assert torch.cuda.is_available()
    return nl


def _oracle_map(ms):
    return {"centres": ms.centres.cpu().numpy(), "structure": ms.structure.cpu().numpy(),
            "vertex_rows": ms.vox2row.cpu().numpy().astype(np.int64)}


def _oracle_decoder(dec, width):
    from oracle import chain as OC
    d = OC.Decoder(depth=2, width=width, in_dim=16)
    d.load_state_dict({k: v.cpu() for k, v in dec.state_dict().items()})
    return d


def _run_case(nl, scans, voxel_size, step_frac, max_depth, width=256, n_rays=1500, grid_dim=256 * 256 * 4, seed=3):
    """Map from `scans`, one mapping iteration over rays of the last scan: fused kernels vs oracle autograd."""
    from oracle import chain as OC
    syn = nl.synthetic
    dev = torch.device("cuda")
    mu = nl.mapping.MapUpdater(voxel_size, init_std=0.02, seed=seed, grid_dim=grid_dim)
    for pts, cos, pose in scans:
        ms = mu.insert_voxels(torch.from_numpy(syn.voxelize(pts, pose, voxel_size)))
    torch.manual_seed(seed)
    dec = nl.lidar.Decoder(depth=2, width=width, in_dim=16, skips=[], embedder="none", multires=0).to(dev)
    pts, cos, pose = scans[-1]
    sel = np.sort(np.random.default_rng(seed).choice(pts.shape[0], min(n_rays, pts.shape[0]), replace=False))
    P, Cn = torch.from_numpy(pts[sel]), torch.from_numpy(cos[sel])
    dirs = P / (P.norm(dim=-1, keepdim=True) + 1e-8)
    pose_o = torch.nn.Parameter(nl.se3pose.OptimizablePose.from_matrix(torch.from_numpy(pose)).data.detach().clone())
    cfg = dict(step_size=step_frac * voxel_size, voxel_size=voxel_size, max_distance=max_depth, truncation=0.3, max_depth=max_depth,
               fs_weight=1.0, sdf_weight=10000.0)
    R = dirs.shape[0]
    eng = nl.engine.SDFEngine(R, R * 64)
    bufs = nl.engine.DecoderBuffers(dec, dev)
    pose6 = pose_o.detach().reshape(1, 6).to(dev).contiguous()
    d_dirs, d_cos = dirs.to(dev).contiguous(), Cn.to(dev)
    gt = torch.norm(P.to(dev), 2, -1) * d_cos
    eng.rays_from_poses(pose6, d_dirs, None)
    ro_g, rd_g = eng.ray_o[:R].clone(), eng.ray_d[:R].clone()
    eng.forward_backward(ms, bufs, R, cfg, gt, d_cos, dir_local=d_dirs, ray_frame=None, n_frames=1, update_decoder=True, update_emb=True,
                         update_pose=True, pose6=pose6)
    st = eng.read_stats()
    assert st.error == 0
    # oracle: rays from the pose through the differentiable SE(3) restatement; traversal hits from the drop-in kernel
    # (bit-exact against the compiled reference, test_gpu_ops.py) so that everything downstream sees identical hits
    ri, rmn, rmx = nl.grid.svo_intersect(ro_g[None].contiguous(), rd_g[None].contiguous(), ms.centres[None].contiguous(),
                                         ms.structure[None].contiguous(), voxel_size, 20)
    raw = (ri[0].cpu().numpy(), rmn[0].cpu().numpy(), rmx[0].cpu().numpy())
    Rm, t = OC.pose_rotation(pose_o), OC.pose_translation(pose_o)
    rd_o = dirs @ Rm.transpose(-1, -2)
    ro_o = t.reshape(1, -1).expand_as(rd_o)
    np.testing.assert_allclose(rd_g.cpu().numpy(), rd_o.detach().numpy(), atol=2e-7)
    emb_o = ms.emb.cpu().float().requires_grad_()
    dec_o = _oracle_decoder(dec, width)
    out = OC.render_rays(ro_o, rd_o, _oracle_map(ms), emb_o, dec_o, cfg["step_size"], voxel_size, max_depth, deterministic=True, raw_hits=raw)
    assert out is not None
    M = st.n_samples
    assert M == int(out["valid_mask"].sum()) and st.n_hit_rays == int(out["ray_mask"].sum()) and st.max_samples == out["z_vals"].shape[1]
    assert np.array_equal(eng.s_vox[:M].cpu().numpy(), out["sampled_idx"][out["valid_mask"]].numpy())      # bit-exact ids
    assert np.array_equal(eng.s_depth[:M].cpu().numpy(), out["z_vals"][out["valid_mask"]].numpy())          # bit-exact depths
    np.testing.assert_allclose(eng.sdf[:M].cpu().numpy(), out["sdf_valid"].detach().numpy(), atol=1e-5)     # north star: 1e-5
    hm = out["ray_mask"]
    loss, parts = OC.sdf_loss(out["z_vals"], out["sdf"], out["valid_mask"], P[hm], Cn[hm], cfg["truncation"], cfg["max_depth"],
                              cfg["fs_weight"], cfg["sdf_weight"])
    np.testing.assert_allclose(st.loss, float(loss), rtol=1e-4)
    loss.backward()
    for g, p in zip(bufs.grads, dec_o.parameters()):
        ref = p.grad.numpy()
        np.testing.assert_allclose(g.cpu().numpy(), ref, rtol=5e-3, atol=1e-4 * np.abs(ref).max())
    ref_pose = pose_o.grad.numpy()
    np.testing.assert_allclose(eng.pose_grad.cpu().numpy()[0], ref_pose, rtol=1e-2, atol=1e-3 * np.abs(ref_pose).max())
    ge, re_ = eng.grad_emb.cpu().numpy(), emb_o.grad.numpy()
    assert np.mean(np.abs(ge - re_) > 1e-2 * np.abs(re_) + 1e-3 * np.abs(re_).max()) < 1e-3
    return st


def test_config1_maicity_shape(nl):
    """voxel 0.2, mapper step 0.5*0.2, max_depth 50, min 1.5 (configs/maicity/maicity.yaml, maicity_01.yaml)."""
    scan = nl.synthetic.make_scan(n_beams=32, n_az=300, seed=11, min_depth=1.5, max_depth=50.0)
    _run_case(nl, [scan], 0.2, 0.5, 50.0)


def test_config2_kitti_incremental_map(nl):
    """KITTI shape, map grown incrementally over 6 scans along +x (create_voxels per frame): node ids / rows of earlier
    scans must survive every update, and the last scan renders against the grown map like the oracle."""
    syn = nl.synthetic
    scans = [syn.make_scan(n_beams=32, n_az=240, seed=100 + i, sensor_xyz=(1.0 * i, 0.0, 0.0)) for i in range(6)]
    from oracle import kernels as OK
    mu = nl.mapping.MapUpdater(0.3)
    orc = OK.Octree(); orc.init(256 * 256 * 4, 16, 0.3)
    prev_rows = None
    for pts, cos, pose in scans:
        vox = syn.voxelize(pts, pose, 0.3)
        ms = mu.insert_voxels(torch.from_numpy(vox))
        orc.insert(vox)
        ov, oc, of = orc.get_centres_and_children()
        c, s, f = OK.map_arrays(ov, oc, of, 0.3)
        assert np.array_equal(ms.centres.cpu().numpy(), c) and np.array_equal(ms.structure.cpu().numpy(), s)   # bit-exact ids
        assert np.array_equal(mu.map_states["voxel_vertex_idx"].numpy(), f)
        if prev_rows is not None:
            n0 = prev_rows.shape[0]
            keep = prev_rows >= 0
            assert np.array_equal(ms.vox2row.cpu().numpy()[:n0][keep], prev_rows[keep])      # rows never renumbered
        prev_rows = ms.vox2row.cpu().numpy().copy()
    _run_case(nl, scans, 0.3, 0.5, 40.0)


def test_config3_newer_college_shape(nl):
    """voxel 0.2, mapper step 0.2*0.2 (ncd.yaml): ~2.5x more samples per ray."""
    scan = nl.synthetic.make_scan(n_beams=32, n_az=200, seed=21, min_depth=1.0, max_depth=40.0)
    st = _run_case(nl, [scan], 0.2, 0.2, 40.0, n_rays=800)
    assert st.max_samples > 20


def test_config0_plumbing_8cubed_width32(nl):
    """BASELINE config 0: tiny scan, 8^3 octree, 2x32 decoder (fp32 CUDA-core decoder path, width 32)."""
    rng = np.random.default_rng(7)
    # points on a plane patch in front of the sensor, all voxel coordinates inside [0, 6]
    n = 1000
    world = np.stack([rng.uniform(1.0, 6.0, n), rng.uniform(1.0, 6.0, n), np.full(n, 2.3) + rng.normal(0, 0.02, n)], -1)
    sensor = np.array([3.5, 3.5, 5.8])
    pts = (world - sensor).astype(np.float32)
    pose = np.eye(4, dtype=np.float32); pose[:3, 3] = sensor
    cos = np.ones(n, np.float32)
    _run_case(nl, [(pts, cos, pose)], 1.0, 0.5, 40.0, width=32, n_rays=1000, grid_dim=8)


def test_edge_cases(nl):
    dev = torch.device("cuda")
    syn = nl.synthetic
    pts, cos, pose = syn.make_scan(n_beams=16, n_az=100, seed=5)
    mu = nl.mapping.MapUpdater(0.3, init_std=0.01)
    ms = mu.insert_voxels(torch.from_numpy(syn.voxelize(pts, pose, 0.3)))
    torch.manual_seed(0)
    dec = nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).to(dev)
    crit = nl.criterion.Criterion(Args())
    o = torch.tensor(pose[:3, 3]).to(dev)
    # 1. rays that miss everything -> None, like render_helpers.py:216-217
    up = torch.tensor([[0.0, 0.0, 1.0]] * 7, device=dev)
    assert nl.render_helpers.render_rays(o.expand(7, 3)[None].contiguous(), up[None].contiguous(), ms, dec, 0.15, 0.3, 0.3, 20, 40.0) is None
    # 2. rays starting inside a surface voxel: min_depth = 0 (f_low starts at 0, intersect_gpu.cu:84)
    c = ms.centres[ms.structure[:, 8] == 1][:5]
    d = torch.nn.functional.normalize(torch.tensor([[1.0, 0.2, -0.1]], device=dev)).expand(5, 3)
    out = nl.render_helpers.render_rays(c[None].contiguous(), d[None].contiguous(), ms, dec, 0.15, 0.3, 0.3, 20, 40.0, deterministic=True)
    assert out is not None and bool(out["ray_mask"].all())
    assert float(out["z_vals"][:, 0].min()) >= 0.0 and float(out["z_vals"][:, 0].max()) < 0.3
    # 3. a single ray, and a ragged batch mixing hits and misses; invalid cells read sdf = 1, z = MAX_DEPTH
    P = torch.from_numpy(pts[:50]).to(dev)
    dirs = P / P.norm(dim=-1, keepdim=True)
    rd = torch.cat([dirs, up], 0)
    out = nl.render_helpers.render_rays(o.expand(rd.shape[0], 3)[None].contiguous(), rd[None].contiguous(), ms, dec, 0.15, 0.3, 0.3, 20, 40.0,
                                        deterministic=True)
    assert out["ray_mask"].shape == (1, 57) and int(out["ray_mask"].sum()) == out["z_vals"].shape[0] <= 50
    inv = ~out["valid_mask"]
    assert bool((out["sdf"][inv] == 1.0).all()) and bool((out["z_vals"][inv] == 80.0).all())
    one = nl.render_helpers.render_rays(o[None, None].contiguous(), dirs[:1][None].contiguous(), ms, dec, 0.15, 0.3, 0.3, 20, 40.0, deterministic=True)
    assert one is not None and one["z_vals"].shape[0] == 1
    # 4. stochastic sampling is reproducible under torch.manual_seed and differs between seeds
    torch.manual_seed(5); a = nl.render_helpers.render_rays(o.expand(50, 3)[None].contiguous(), dirs[None].contiguous(), ms, dec, 0.15, 0.3, 0.3, 20, 40.0)
    torch.manual_seed(5); b = nl.render_helpers.render_rays(o.expand(50, 3)[None].contiguous(), dirs[None].contiguous(), ms, dec, 0.15, 0.3, 0.3, 20, 40.0)
    torch.manual_seed(6); c2 = nl.render_helpers.render_rays(o.expand(50, 3)[None].contiguous(), dirs[None].contiguous(), ms, dec, 0.15, 0.3, 0.3, 20, 40.0)
    assert torch.equal(a["z_vals"], b["z_vals"]) and not torch.equal(a["z_vals"], c2["z_vals"])
    # 5. reference_compat=0 (tail segment for every ray) never yields fewer samples than the quirk-compatible sampler
    q1 = nl.render_helpers.render_rays(o.expand(50, 3)[None].contiguous(), dirs[None].contiguous(), ms, dec, 0.15, 0.3, 0.3, 20, 40.0, deterministic=True)
    q0 = nl.render_helpers.render_rays(o.expand(50, 3)[None].contiguous(), dirs[None].contiguous(), ms, dec, 0.15, 0.3, 0.3, 20, 40.0, deterministic=True,
                                       reference_compat=False)
    assert int(q0["valid_mask"].sum()) >= int(q1["valid_mask"].sum())
    # 6. tracking against a map that the rays never see: (pose, None) like render_helpers.py:488-491 / tracking.py:136
    far = nl.frame.LidarFrame(3, torch.from_numpy(pts), torch.from_numpy(cos),
                              nl.se3pose.OptimizablePose(torch.tensor([5000.0, 5000.0, 5000.0, 0.0, 0.0, 0.0])), new_keyframe=True)
    pose_out, hit = nl.render_helpers.track_frame(far.pose, far, ms, dec, crit, 0.3, N_rays=64, step_size=0.06, num_iterations=2, truncation=0.3,
                                                  learning_rate=0.06, max_voxel_hit=20, max_distance=40.0)
    assert hit is None and torch.equal(pose_out.data.detach().cpu(), far.pose.data.detach())
    # 7. get_scores: res^3 lattice per voxel, values equal to a direct decoder evaluation at the voxel centre lattice point
    sc = nl.render_helpers.get_scores(dec, ms, 0.3, bits=3)
    assert sc.shape == (ms.n_nodes, 3, 3, 3, 1) and bool(torch.isfinite(sc).all())
    # 8. the Decoder module survives deepcopy / pickle / state_dict round trips and still evaluates through the kernels
    import copy
    import pickle
    x = torch.randn(100, 16, device=dev) * 0.05
    y0 = dec(x)["sdf"]
    for clone in (copy.deepcopy(dec), pickle.loads(pickle.dumps(dec)).to(dev)):
        assert torch.equal(clone(x)["sdf"], y0)


def test_cooperative_traversal_equals_thread_per_ray(nl):
    """The warp-cooperative traversal over the packed octree image and the thread-per-ray kernel over the reference's arrays
    produce the same sample list bit for bit (full 82 k-ray scan, plus rays from inside the map and rays that miss)."""
    syn = nl.synthetic
    pts, cos, pose = syn.make_scan()
    mu = nl.mapping.MapUpdater(0.3)
    ms = mu.insert_voxels(torch.from_numpy(syn.voxelize(pts, pose, 0.3)))
    dev = torch.device("cuda")
    P = torch.from_numpy(pts).to(dev)
    d = P / P.norm(dim=-1, keepdim=True)
    Rm = torch.from_numpy(pose[:3, :3]).to(dev)
    rd = torch.cat([d @ Rm.T, torch.tensor([[0.0, 0.0, 1.0], [0.6, 0.8, 0.0]], device=dev)]).contiguous()
    ro = torch.from_numpy(pose[:3, 3]).to(dev).expand(rd.shape[0], 3).contiguous().clone()
    ro[1000:1200] = ms.centres[ms.structure[:, 8] == 1][:200]          # rays starting inside surface voxels
    R = rd.shape[0]
    cfg = dict(step_size=0.15, voxel_size=0.3, max_distance=40.0)
    outs = []
    for packed in (True, False):
        eng = nl.engine.SDFEngine(R, R * 24)
        eng.use_packed_octree = packed
        eng.render_samples(ms, R, cfg, ro, rd)
        st = eng.read_stats()
        assert st.error == 0
        M = st.n_samples
        outs.append((st.n_hit_rays, st.max_hits, st.max_samples, M, eng.hit_rank[:R].clone(), eng.ray_nsamp[:R].clone(),
                     eng.s_vox[:M].clone(), eng.s_depth[:M].clone(), eng.s_ray[:M].clone()))
    a, b = outs
    assert a[:4] == b[:4] and a[3] > 700000
    for x, y in zip(a[4:], b[4:]):
        assert torch.equal(x, y)


def test_track_frame_device_selection_and_cuda_graph(nl):
    """Tracking with rays drawn on the GPU, eagerly and as a captured + replayed CUDA graph.  With N_rays == number of points
    (every ray selected) and deterministic sampling the two run the same arithmetic, so the optimised poses must agree up to
    float atomics / the device-side evaluation of Adam's bias correction.  Three iterations: Adam's first steps are
    lr * g / |g| per component, so over many iterations a component whose gradient passes through zero amplifies atomics noise
    into +-lr differences -- that is a property of Adam, not of the graph."""
    syn = nl.synthetic
    pts, cos, pose = syn.make_scan(n_beams=16, n_az=120, seed=9)
    N = pts.shape[0]            # every ray selected in both modes
    mu = nl.mapping.MapUpdater(0.3, init_std=0.02, seed=4)
    ms = mu.insert_voxels(torch.from_numpy(syn.voxelize(pts, pose, 0.3)))
    dev = torch.device("cuda")
    torch.manual_seed(3)
    dec = nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).to(dev)
    crit = nl.criterion.Criterion(Args())
    start6 = nl.se3pose.OptimizablePose.from_matrix(torch.from_numpy(pose)).data.detach().clone()
    start6[:3] += torch.tensor([0.05, -0.03, 0.02])
    outs = []
    for graph in (False, True):
        fr = nl.frame.LidarFrame(5, torch.from_numpy(pts), torch.from_numpy(cos), nl.se3pose.OptimizablePose(start6.clone()), new_keyframe=True)
        out, hit = nl.render_helpers.track_frame(fr.pose, fr, ms, dec, crit, 0.3, N_rays=N, step_size=0.06, num_iterations=3, truncation=0.3,
                                                 learning_rate=0.03, max_voxel_hit=20, max_distance=40.0, ray_selection="device", cuda_graph=graph,
                                                 deterministic=True)
        assert hit is not None and hit.shape == (N,) and int(hit.sum()) > 0.5 * N
        assert torch.equal(fr.pose.data.detach(), start6)          # the input pose object is not modified
        outs.append(out.data.detach().cpu().clone())
    assert float((outs[0] - start6).abs().max()) > 1e-3            # the Adam steps moved the pose
    torch.testing.assert_close(outs[1], outs[0], rtol=0, atol=2e-4)
    # a second scan through the cached graph (same map and decoder): identical result
    fr = nl.frame.LidarFrame(6, torch.from_numpy(pts), torch.from_numpy(cos), nl.se3pose.OptimizablePose(start6.clone()), new_keyframe=True)
    out2, _ = nl.render_helpers.track_frame(fr.pose, fr, ms, dec, crit, 0.3, N_rays=N, step_size=0.06, num_iterations=3, truncation=0.3,
                                            learning_rate=0.03, max_voxel_hit=20, max_distance=40.0, ray_selection="device", cuda_graph=True,
                                            deterministic=True)
    torch.testing.assert_close(out2.data.detach().cpu(), outs[1], rtol=0, atol=2e-5)


def _ba_case(nl, far=False):
    syn = nl.synthetic
    scans = [syn.make_scan(n_beams=32, n_az=300, seed=200 + i, sensor_xyz=(1.0 * i, 0.0, 0.0)) for i in range(2)]
    mu = nl.mapping.MapUpdater(0.3, init_std=0.01, seed=2)
    for pts, cos, pose in scans:
        ms = mu.insert_voxels(torch.from_numpy(syn.voxelize(pts, pose, 0.3)))
    torch.manual_seed(5)
    dec = nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).cuda()
    frames = []
    for i, (pts, cos, pose) in enumerate(scans):
        T = torch.from_numpy(pose.copy())
        if far:
            T[:3, 3] += 500.0                     # the sensor is nowhere near the map: no ray hits anything
        frames.append(nl.frame.LidarFrame(i, torch.from_numpy(pts), torch.from_numpy(cos), nl.se3pose.OptimizablePose.from_matrix(T), new_keyframe=True))
    return ms, dec, frames


def _run_ba(nl, ms, dec, frames, n_it=3, **kw):
    emb = ms.emb.clone()
    torch.manual_seed(9)
    nl.render_helpers.bundle_adjust_frames(frames, emb, ms, dec, nl.criterion.Criterion(Args()), 0.3, 0.15, N_rays=512, num_iterations=n_it,
                                           truncation=0.3, max_voxel_hit=20, max_distance=40.0, learning_rate=[0.01, 0.005, 0.001],
                                           deterministic=True, **{"ray_selection": "host", **kw})
    torch.cuda.synchronize()
    return emb, {k: v.detach().clone() for k, v in dec.state_dict().items()}, torch.stack([f.pose.data.detach().cpu() for f in frames])


def test_sample_capacity_overflow_is_undone_and_redone(nl, monkeypatch):
    """A call whose sample capacity is exceeded must not train on a truncated sample set: the parameters and the RNG state are
    restored and the call is redone with a larger engine -- same result as with enough capacity from the start."""
    rh = nl.render_helpers
    ms, dec, frames = _ba_case(nl)
    e_ok, d_ok, p_ok = _run_ba(nl, ms, dec, frames)
    rh._ENGINES.clear()
    monkeypatch.setattr(rh, "SAMPLES_PER_RAY_MAP", 1)       # ~10 samples per ray are needed: every iteration overflows
    small = {"n": 0}
    real_engine = rh._engine

    def tiny_engine(n_rays, min_samples, device):             # bypass the 2^16-sample floor of the engine cache for this test
        if min_samples <= n_rays:
            small["n"] += 1
            return nl.engine.SDFEngine(n_rays, min_samples, device)
        return real_engine(n_rays, min_samples, device)
    monkeypatch.setattr(rh, "_engine", tiny_engine)
    ms2, dec2, frames2 = _ba_case(nl)
    e2, d2, p2 = _run_ba(nl, ms2, dec2, frames2)
    assert small["n"] == 1                                     # the first attempt really ran on the too-small engine
    torch.testing.assert_close(p2, p_ok, rtol=0, atol=2e-5)
    for k in d_ok:
        torch.testing.assert_close(d2[k], d_ok[k], rtol=0, atol=2e-5)
    assert float((e2.float() - e_ok.float()).abs().gt(2e-3).float().mean()) < 1e-3


def test_iterations_without_hits_skip_the_optimiser_step(nl, capsys):
    """render_helpers.py:405-409: when nothing is hit the reference `continue`s -- no optimiser step, no change of Adam's state.
    Here the Adam kernels read a device-side skip flag; nothing may move and the reference's message is printed per skipped iteration."""
    ms, dec, frames = _ba_case(nl, far=True)
    emb0 = ms.emb.clone()
    dec0 = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    p0 = torch.stack([f.pose.data.detach().cpu() for f in frames])
    e, d, p = _run_ba(nl, ms, dec, frames, n_it=4)
    assert torch.equal(e, emb0) and torch.equal(p, p0)
    assert all(torch.equal(d[k], dec0[k]) for k in d)
    assert capsys.readouterr().out.count("Encouter a bug while Mapping") == 4


def test_mapstate_cache_is_keyed_on_content(nl):
    """a reference-format map_states dict whose tensors changed in place (or were re-created at the same address with the
    same shapes) must not hit the cache."""
    syn = nl.synthetic
    pts, cos, pose = syn.make_scan(n_beams=16, n_az=200, seed=3)
    mu = nl.mapping.MapUpdater(0.3)
    mu.insert_voxels(torch.from_numpy(syn.voxelize(pts, pose, 0.3)))
    ms = {k: v for k, v in mu.map_states.items() if k != "_mapstate"}
    a = nl.engine.MapState.from_map_states(ms, "cuda")
    b = nl.engine.MapState.from_map_states(ms, "cuda")
    assert a.vox2row.data_ptr() == b.vox2row.data_ptr()                      # unchanged content: cache hit
    ms["voxel_structure"][0, 0] = -1                                         # in-place edit, same address, same shape
    c = nl.engine.MapState.from_map_states(ms, "cuda")
    assert int(c.structure[0, 0]) == -1 and c.structure.data_ptr() != a.structure.data_ptr()


def test_bundle_adjust_frames_cuda_graph_equals_eager(nl):
    """The captured + replayed mapping iteration (_MapGraph, the default path) against the eager loop.  With N_rays == number of
    points of every scan (all rays selected in both modes) and deterministic sampling the two run the same arithmetic."""
    syn = nl.synthetic
    pts, cos, pose = syn.make_scan(n_beams=16, n_az=160, seed=12)
    N = pts.shape[0]
    res = []
    for graph in (False, True):
        mu = nl.mapping.MapUpdater(0.3, init_std=0.01, seed=2)
        ms = mu.insert_voxels(torch.from_numpy(syn.voxelize(pts, pose, 0.3)))
        torch.manual_seed(5)
        dec = nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).cuda()
        frames = []
        for i in range(2):
            T = torch.from_numpy(pose.copy())
            T[:3, 3] += torch.tensor([0.03 * i, -0.02 * i, 0.01 * i])
            frames.append(nl.frame.LidarFrame(i, torch.from_numpy(pts), torch.from_numpy(cos), nl.se3pose.OptimizablePose.from_matrix(T), new_keyframe=True))
        emb = ms.emb.clone()
        e_init = emb.float().cpu()
        for call in range(2):             # the second call goes through the cached graph
            nl.render_helpers.bundle_adjust_frames(frames, emb, ms, dec, nl.criterion.Criterion(Args()), 0.3, 0.15, N_rays=N, num_iterations=3,
                                                   truncation=0.3, max_voxel_hit=20, max_distance=40.0, learning_rate=[0.01, 0.005, 0.001],
                                                   deterministic=True, ray_selection="device", cuda_graph=graph)
        torch.cuda.synchronize()
        res.append((emb.float().cpu(), {k: v.detach().cpu().clone() for k, v in dec.state_dict().items()},
                    torch.stack([f.pose.data.detach().cpu() for f in frames])))
    (e0, d0, p0), (e1, d1, p1) = res
    assert torch.equal(p0[0], p1[0])                                       # frame index 0 stays frozen in both
    torch.testing.assert_close(p1, p0, rtol=0, atol=1e-4)
    for k in d0:
        torch.testing.assert_close(d1[k], d0[k], rtol=0, atol=1e-4)
    assert float((e1 - e0).abs().gt(2e-3).float().mean()) < 2e-3
    assert float((e0 - e_init).abs().max()) > 1e-3                         # and it did train


def test_eager_mapping_with_device_selection_never_synchronises_and_stays_finite(nl):
    """The default drop-in path for reference-format maps: eager bundle_adjust_frames, rays selected on the device, the decoder's
    weight gradients and Adam pipelined on the side stream -- a loop without a single host synchronisation.  Repeated calls (fresh
    optimiser state from recycled device blocks each time) must train like the captured path does."""
    syn = nl.synthetic
    pts, cos, pose = syn.make_scan(n_beams=32, n_az=400, seed=12)
    res = {}
    for graph in (False, True):
        mu = nl.mapping.MapUpdater(0.3, init_std=0.01, seed=2)
        ms = mu.insert_voxels(torch.from_numpy(syn.voxelize(pts, pose, 0.3)))
        torch.manual_seed(5)
        dec = nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).cuda()
        fr = [nl.frame.LidarFrame(0, torch.from_numpy(pts), torch.from_numpy(cos), nl.se3pose.OptimizablePose.from_matrix(torch.from_numpy(pose.copy())),
                                  new_keyframe=True)]
        for call in range(8):
            junk = [torch.full((n,), float("nan"), device="cuda") for n in (256, 256, 4096, 4096, 65536, 65536)]   # what recycled blocks may hold
            del junk
            nl.render_helpers.bundle_adjust_frames(fr, mu.embeddings, ms, dec, nl.criterion.Criterion(Args()), 0.3, 0.15, N_rays=1024, num_iterations=25,
                                                   truncation=0.3, max_voxel_hit=20, max_distance=40.0, learning_rate=[0.01, 0.005, 0.001],
                                                   ray_selection="device", cuda_graph=graph)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(mu.embeddings.float()).all()) and all(bool(torch.isfinite(p).all()) for p in dec.parameters())
        res[graph] = float(mu.embeddings.float().abs().mean())
    assert abs(res[False] - res[True]) < 0.1 * res[True], res


def test_incremental_map_update_equals_full_and_graphs_survive_it(nl):
    """mapping.MapUpdater: the device-resident, incrementally patched map equals the fully re-exported one after every
    scan, its arrays keep their addresses while the map grows, and the captured mapping graph is therefore reused across map updates."""
    syn = nl.synthetic
    rh = nl.render_helpers
    scans = [syn.make_scan(n_beams=16, n_az=160, seed=300 + i, sensor_xyz=(0.8 * i, 0.1 * i, 0.0)) for i in range(4)]
    inc = nl.mapping.MapUpdater(0.3, init_std=0.01, seed=3, reserve_nodes=100_000, reserve_rows=100_000)
    ful = nl.mapping.MapUpdater(0.3, init_std=0.01, seed=3, incremental=False)
    torch.manual_seed(5)
    dec = nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).cuda()
    crit = nl.criterion.Criterion(Args())
    ptrs, graphs, frames = set(), [], []
    rows_before = 0
    N = min(s[0].shape[0] for s in scans)
    for i, (pts, cos, pose) in enumerate(scans):
        vox = torch.from_numpy(syn.voxelize(pts, pose, 0.3))
        a, b = inc.insert_voxels(vox), ful.insert_voxels(vox)
        assert a.n_nodes == b.n_nodes and torch.equal(a.centres, b.centres) and torch.equal(a.structure, b.structure)
        assert torch.equal(a.vox2row, b.vox2row) and a.emb.shape == b.emb.shape
        assert torch.equal(a.emb[rows_before:], b.emb[rows_before:])              # rows handed out by this update (older ones are trained below)
        rows_before = a.emb.shape[0]
        assert torch.equal(a.packed_children(), b.packed_children())
        for k in ("voxel_vertex_idx", "voxel_center_xyz", "voxel_structure", "voxel_id2embedding_id"):
            assert torch.equal(inc.map_states[k], ful.map_states[k]), k
        assert inc.last_update["dirty_rows"] <= a.n_nodes and (i == 0 or inc.last_update["dirty_rows"] < 0.7 * a.n_nodes)
        ptrs.add((a.centres.data_ptr(), a.structure.data_ptr(), a.vox2row.data_ptr(), a.packed_children().data_ptr(), a.emb.data_ptr()))
        frames.append(nl.frame.LidarFrame(i + 1, torch.from_numpy(pts[:N]), torch.from_numpy(cos[:N]), nl.se3pose.OptimizablePose.from_matrix(torch.from_numpy(pose.copy())),
                                          new_keyframe=True))
        if i >= 1:      # a mapping call on the growing map after every update, through the default (captured) path
            rh.bundle_adjust_frames(frames[-2:], inc.embeddings, a, dec, crit, 0.3, 0.15, N_rays=512, num_iterations=2, truncation=0.3,
                                    max_voxel_hit=20, max_distance=40.0, learning_rate=[0.01, 0.005, 0.001])
            graphs.append(rh._MapGraph._cache["g"])
    assert len(ptrs) == 1                                        # reserved capacity: no buffer ever moved while the map grew
    assert all(g is graphs[0] for g in graphs)                   # ONE capture served every map version
    torch.cuda.synchronize()
    assert bool(torch.isfinite(inc.embeddings.float()).all())


def test_tracker_mapper_handoff_on_device(nl):
    """share.SharedMap: the tracker reads a published device snapshot (embedding rows + decoder parameters) on its own
    stream while the mapper keeps optimising the live table; the result equals tracking against a frozen copy of the state at
    publication time."""
    syn = nl.synthetic
    rh = nl.render_helpers
    pts, cos, pose = syn.make_scan(n_beams=16, n_az=200, seed=31)
    mu = nl.mapping.MapUpdater(0.3, init_std=0.02, seed=4)
    ms = mu.insert_voxels(torch.from_numpy(syn.voxelize(pts, pose, 0.3)))
    dev = torch.device("cuda")
    torch.manual_seed(3)
    mk = lambda: nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).to(dev)
    dec = mk()
    crit = nl.criterion.Criterion(Args())
    shared = nl.share.SharedMap(mk)
    mu.readers = shared
    start6 = nl.se3pose.OptimizablePose.from_matrix(torch.from_numpy(pose)).data.detach().clone()
    start6[:3] += torch.tensor([0.04, -0.03, 0.02])
    N = pts.shape[0]
    frame = lambda i: nl.frame.LidarFrame(i, torch.from_numpy(pts), torch.from_numpy(cos), nl.se3pose.OptimizablePose(start6.clone()), new_keyframe=True)
    # reference result: tracking against a frozen deep copy of the published state
    import copy
    emb0, dec0 = ms.emb.clone(), copy.deepcopy(dec)
    frozen = nl.engine.MapState(ms.centres, ms.structure, ms.vox2row, emb0, dev)
    kw = dict(N_rays=N, step_size=0.06, num_iterations=3, truncation=0.3, learning_rate=0.03, max_voxel_hit=20, max_distance=40.0,
              ray_selection="device", cuda_graph=False, deterministic=True)
    want, _ = rh.track_frame(frame(5).pose, frame(5), frozen, dec0, crit, 0.3, **kw)
    # publish, then let the mapper train the live table + decoder on the main stream while the tracker runs on another one
    shared.publish(ms, dec)
    t_stream = torch.cuda.Stream()
    t_stream.wait_stream(torch.cuda.current_stream())
    fr_map = [frame(0), frame(1)]
    rh.bundle_adjust_frames(fr_map, mu.embeddings, ms, dec, crit, 0.3, 0.15, N_rays=512, num_iterations=3, truncation=0.3, max_voxel_hit=20,
                            max_distance=40.0, learning_rate=[0.01, 0.005, 0.001], cuda_graph=False)
    with torch.cuda.stream(t_stream):
        m_t, dec_t, slot = shared.acquire()
        got, hit = rh.track_frame(frame(5).pose, frame(5), m_t, dec_t, crit, 0.3, **kw)
        shared.release(slot)
    torch.cuda.synchronize()
    assert hit is not None
    assert float((ms.emb.float() - emb0.float()).abs().max()) > 1e-3                    # the live table moved meanwhile
    torch.testing.assert_close(got.data.detach().cpu(), want.data.detach().cpu(), rtol=0, atol=2e-5)
    # a later map update waits for the reader event and still matches a fresh full export
    pts2, cos2, pose2 = syn.make_scan(n_beams=16, n_az=200, seed=32, sensor_xyz=(1.0, 0.0, 0.0))
    ms2 = mu.insert_voxels(torch.from_numpy(syn.voxelize(pts2, pose2, 0.3)))
    c, s, v = mu.svo.export_map()
    assert torch.equal(ms2.centres.cpu(), c) and torch.equal(ms2.structure.cpu(), s)
    # ShareData drop-in: device tensors in, device tensors out, decoder rebuilt from the flat snapshot
    sd = nl.share.ShareData()
    sd.decoder = dec
    sd.states = mu.map_states
    d2 = sd.decoder
    assert all(torch.equal(a, b) for a, b in zip(d2.state_dict().values(), dec.state_dict().values()))
    m2 = sd.map_state()
    assert m2.emb.is_cuda and torch.equal(m2.vox2row, ms2.vox2row)


def _ipc_child(q_in, q_out):
    """Runs in a spawned process: receives share.ShareData's device tensors (CUDA IPC handles under the hood), reads them and
    writes into one of them in place."""
    import torch
    st, dec_flat = q_in.get(timeout=120)
    emb = st["emb"]
    assert emb.is_cuda and st["vox2row"].is_cuda
    q_out.put((float(emb.float().sum()), int(st["vox2row"].long().clamp(min=0).sum()), float(dec_flat.sum())))
    emb[0, 0] = 123.0                                    # visible to the parent only if the memory is really shared
    torch.cuda.synchronize()
    q_out.put("done")
    q_in.get(timeout=120)                                # keep the mapping alive until the parent has looked


def test_share_data_crosses_processes_as_cuda_ipc(nl):
    """share.ShareData's payload goes to another process as CUDA IPC handles: no host staging, and the receiver sees the very
    same device memory (the reference pickles CPU copies of every tensor through a Manager, mapping.py:227-232)."""
    import torch.multiprocessing as mp
    syn = nl.synthetic
    pts, cos, pose = syn.make_scan(n_beams=16, n_az=100, seed=41)
    mu = nl.mapping.MapUpdater(0.3, init_std=0.02, seed=4)
    mu.insert_voxels(torch.from_numpy(syn.voxelize(pts, pose, 0.3)))
    torch.manual_seed(3)
    dec = nl.lidar.Decoder(depth=2, width=256, in_dim=16, skips=[], embedder="none", multires=0).cuda()
    sd = nl.share.ShareData()
    sd.decoder = dec
    sd.states = mu.map_states
    st = sd.states
    flat = sd._decoder_flat
    want = (float(st["emb"].float().sum()), int(st["vox2row"].long().clamp(min=0).sum()), float(flat.sum()))
    ctx = mp.get_context("spawn")
    q_in, q_out = ctx.Queue(), ctx.Queue()
    p = ctx.Process(target=_ipc_child, args=(q_in, q_out))
    p.start()
    try:
        q_in.put(({"emb": st["emb"], "vox2row": st["vox2row"]}, flat))
        got = q_out.get(timeout=180)
        assert got[1] == want[1] and abs(got[0] - want[0]) < 1e-3 and abs(got[2] - want[2]) < 1e-3
        assert q_out.get(timeout=60) == "done"
        torch.cuda.synchronize()
        assert float(st["emb"][0, 0]) == 123.0             # the child's in-place write landed in the parent's tensor
    finally:
        q_in.put("bye")
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    assert p.exitcode == 0


def test_slam_loop_end_to_end_tracks_the_synthetic_trajectory():
    """The reference's loop in miniature through the drop-in API only (scripts/demo_slam.py: track_frame -> bundle_adjust_frames over
    the keyframe window -> incremental map update -> device-side publication, then GPU marching cubes) on 6 synthetic scans 0.5 m
    apart: the estimated trajectory stays within a fraction of a voxel (0.3 m) of the ground truth although every scan is tracked from
    the constant-velocity guess against a map that was itself built from the estimated poses."""
    import json, os, subprocess, sys
    from util import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "demo_slam.py"), "--scans", "6", "--init-calls", "8"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["translation_error_m"]["max"] < 0.10 and d["rotation_error_fro"]["max"] < 0.01, d
    assert d["mesh"]["triangles"] > 10000
    assert d["map"]["last_update"]["dirty_rows"] < 0.5 * d["map"]["nodes"]
