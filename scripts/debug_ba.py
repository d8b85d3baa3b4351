"""GPU debug: reference vs drop-in bundle_adjust_frames with the decoder frozen -- where do the embedding updates diverge?"""
import copy, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nerfloam_b200 as nl
from oracle import ref_harness as H
import test_gpu_dropin_reference as T
ref = H.load()
crit0 = ref.Criterion(H.args(T.MD, T.TR))
class RecCrit(ref.Criterion):
    def forward(self, *a, **k):
        loss, d = super().forward(*a, **k)
        self.rec.append(float(loss.detach()))
        return loss, d
for upd in (False, True):
  for impl in ("tc", "simt"):
    os.environ["NL_MLP_IMPL"] = impl
    for n_it in (1, 2, 3):
        scans, ms0, dec0 = T._state(ref, nl, table_rows=4_000_000)
        kw = dict(voxel_size=T.VS, step_size=0.5 * T.VS, N_rays=1024, num_iterations=n_it, truncation=T.TR, max_voxel_hit=20, max_distance=T.MD,
                  learning_rate=T.LR, update_pose=True, update_decoder=upd)
        rc = RecCrit(H.args(T.MD, T.TR)); rc.rec = []
        ms_r, dec_r, fr_r = T._clone_ms(ms0), copy.deepcopy(dec0), T._ref_frames(ref, scans)
        torch.manual_seed(11)
        with H.pinned(ref):
            ref.orig["bundle_adjust_frames"](fr_r, ms_r["voxel_vertex_emb"], ms_r, dec_r, rc, **kw)
        ms_p, dec_p, fr_p = T._clone_ms(ms0), copy.deepcopy(dec0), T._ref_frames(ref, scans)
        ll = []
        torch.manual_seed(11)
        nl.render_helpers.bundle_adjust_frames(fr_p, ms_p["voxel_vertex_emb"], ms_p, dec_p, crit0, deterministic=True, ray_selection="host", loss_log=ll, **kw)
        torch.cuda.synchronize()
        e_r, e_p, e_0 = (t["voxel_vertex_emb"].detach().float().cpu() for t in (ms_r, ms_p, ms0))
        u_r, u_p = e_r - e_0, e_p - e_0
        print(f"upd_dec={upd} impl={impl} it={n_it} ref loss {rc.rec} ours {ll} update_rel {float((u_p-u_r).norm()/u_r.norm()):.3e} "
              f"frac>2e-3 {float(((e_p-e_r).abs()>2e-3).float().mean()):.3e} max {float((e_p-e_r).abs().max()):.3e} |u_r| {float(u_r.norm()):.3f} "
              f"pose diff {float(torch.stack([a.pose.data.detach().cpu()-b.pose.data.detach().cpu() for a,b in zip(fr_r,fr_p)]).abs().max()):.2e}")
