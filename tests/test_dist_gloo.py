"""World-size-2 gloo test (CPU) of the ray-sharded path's host logic: the statistics exchange of
nerf-loam_b200/dist.py reproduces the unsharded loss normalisation and loss (criterion.py:84-100)."""
import ctypes
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_case(seed=5, R=300, S=14):
    """Synthetic padded render output with ragged rows (like tests/golden/make_golden.py::golden_criterion)."""
    g = torch.Generator().manual_seed(seed)
    pts = torch.randn(R, 3, generator=g) * torch.tensor([10.0, 6.0, 1.0])
    cos = torch.rand(R, generator=g).clamp(min=0.05)
    cos[::3] = 1.0
    depth = pts.norm(dim=-1)
    nvalid = torch.randint(0, S + 1, (R,), generator=g)           # some rays have no samples at all (missed rays)
    z = depth[:, None] + (torch.rand(R, S, generator=g) - 0.6) * 2.0
    valid = torch.arange(S)[None, :] < nvalid[:, None]
    sdf_v = torch.randn(int(valid.sum()), generator=g) * 0.3
    return pts, cos, z, valid, sdf_v


def _raw_counters(pts, cos, z, valid, trunc=0.3, max_depth=40.0):
    """Per-shard raw statistics exactly as csrc/render.cu accumulates them (rows = hit rays of the shard)."""
    hit = valid.any(-1)
    d = (pts.norm(dim=-1) * cos)[hit]
    c = cos[hit]
    zz = (z * cos[:, None])[hit]
    v = valid[hit]
    ns = v.sum(-1)

    def flags(zc, dd):
        front = zc < (dd - trunc)
        back = zc > (dd + trunc)
        dm = (dd > 0) & (dd < max_depth)
        return front, (~front) & (~back) & dm
    f, s = flags(zz, d[:, None].expand_as(zz))
    fp, sp = flags(80.0 * c, d)
    return dict(n_hit=int(hit.sum()), max_samples=int(ns.max()) if ns.numel() else 0, cnt_fs=int((f & v).sum()), cnt_sdf=int((s & v).sum()),
                pad_fs_rays=int(fp.sum()), pad_fs_nsamp=int(ns[fp].sum()), pad_sdf_rays=int(sp.sum()), pad_sdf_nsamp=int(ns[sp].sum()),
                pad_sdf_d2=float((d[sp].double() ** 2).sum()), pad_sdf_d2_nsamp=float((d[sp].double() ** 2 * ns[sp]).sum()))


def _prepare(st):
    """k_loss_prepare (csrc/render.cu) in Python."""
    S = st.max_samples
    n_fs = st.cnt_fs_valid + st.pad_fs_rays * S - st.pad_fs_nsamp
    n_sdf = st.cnt_sdf_valid + st.pad_sdf_rays * S - st.pad_sdf_nsamp
    return n_fs, n_sdf, st.n_hit_rays * S


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nerfloam_b200 as nl
    from nerfloam_b200 import dist as nldist
    pts, cos, z, valid, sdf_v = _make_case()
    R = pts.shape[0]
    lo, hi = nldist.shard_bounds(R, rank, world)
    rc = _raw_counters(pts[lo:hi], cos[lo:hi], z[lo:hi], valid[lo:hi])
    st = nl._capi.RenderStats()
    st.n_hit_rays, st.max_samples = rc["n_hit"], rc["max_samples"]
    st.cnt_fs_valid, st.cnt_sdf_valid = rc["cnt_fs"], rc["cnt_sdf"]
    st.pad_fs_rays, st.pad_fs_nsamp, st.pad_sdf_rays, st.pad_sdf_nsamp = rc["pad_fs_rays"], rc["pad_fs_nsamp"], rc["pad_sdf_rays"], rc["pad_sdf_nsamp"]
    st.pad_sdf_d2, st.pad_sdf_d2_nsamp = rc["pad_sdf_d2"], rc["pad_sdf_d2_nsamp"]
    st.fs_sum, st.sdf_sum = 1.5 + rank + 1e-9 * (rank + 1), 2.25 * (rank + 1)
    st.error = 2 if rank == 1 else 0                     # one rank ran out of sample capacity: every rank must learn it (OR)
    buf = torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8).clone()
    # exchange 1: ONE collective for every statistic (counters SUM, R_hit SUM, S_max MAX, error bits OR)
    xbuf = nldist.allreduce_sample_stats(buf)
    assert xbuf.numel() == nldist.stats_words(world)
    # exchange 2: ONE collective for the flat fp32 gradient buffer, the loss sums riding in its 4-float header as (hi, lo) pairs
    gradflat = torch.full((16 + 16 + 7 * 16,), float(rank + 1))
    gradflat[:16] = 0
    nldist.allreduce_grads_with_loss(buf, gradflat)
    ok_flat = bool((gradflat[16:] == 3.0).all())
    dec_flat = torch.full((70464,), float(rank + 1) * 0.5)
    nldist.allreduce_flat(dec_flat)
    ok_flat = ok_flat and bool((dec_flat == 1.5).all())
    out = nl._capi.RenderStats.from_buffer_copy(buf.numpy().tobytes())
    q.put((rank, _prepare(out), out.fs_sum, out.sdf_sum, 3.0 if ok_flat else -1.0, (lo, hi), out.error))
    dist.destroy_process_group()


def test_sharded_statistics_equal_unsharded():
    import nerfloam_b200 as nl
    from oracle import chain as OC
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # unsharded truth from the oracle's Criterion restatement
    pts, cos, z, valid, sdf_v = _make_case()
    hit = valid.any(-1)
    zh = torch.where(valid, z, torch.full_like(z, 80.0))[hit]
    S = int(valid.sum(-1).max())
    sdf = torch.ones(valid.shape).masked_scatter(valid, sdf_v)[hit][:, :S]
    loss, parts = OC.sdf_loss(zh[:, :S], sdf, valid[hit][:, :S], pts[hit], cos[hit], 0.3, 40.0, 1.0, 10000.0)
    for rank, (n_fs, n_sdf, N), fs_sum, sdf_sum, g00, (lo, hi), err in res:
        assert n_fs == int(parts["n_fs"]) and n_sdf == int(parts["n_sdf"])       # global mask counts
        assert N == int(hit.sum()) * S                                            # global mean denominator
        assert abs(fs_sum - (1.5 + 2.5 + 3e-9)) < 1e-12 and sdf_sum == 2.25 + 4.5   # (hi, lo) float pairs keep ~48 bits of the f64 sums
        assert g00 == 3.0
        assert err == 2
    assert res[0][5][1] == res[1][5][0] and res[0][5][0] == 0 and res[1][5][1] == pts.shape[0]   # contiguous cover
