"""GPU debug: where do the reference's and the drop-in's tracking runs diverge?"""
import copy, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nerfloam_b200 as nl
from oracle import ref_harness as H
import test_gpu_dropin_reference as T
ref = H.load()
scans, ms0, dec0 = T._state(ref, nl)
crit = ref.Criterion(H.args(T.MD, T.TR))
VS = T.VS
kw = dict(voxel_size=VS, N_rays=1024, step_size=0.2 * VS, truncation=T.TR, learning_rate=0.06, max_voxel_hit=20, max_distance=T.MD, depth_variance=True)

class RecCrit(ref.Criterion):
    def forward(self, *a, **k):
        loss, d = super().forward(*a, **k)
        self.rec.append(float(loss))
        return loss, d
rc = RecCrit(H.args(T.MD, T.TR)); rc.rec = []

def frame():
    f = T._ref_frames(ref, scans)[1]; f.index = 5
    return f
for n_it in (1, 2, 3):
    f_r = frame(); rc.rec = []
    torch.manual_seed(21)
    with H.pinned(ref):
        pose_r, hit_r = ref.orig["track_frame"](copy.deepcopy(f_r.pose), f_r, T._clone_ms(ms0), copy.deepcopy(dec0), rc, num_iterations=n_it, **kw)
    f_p = frame(); ll = []
    torch.manual_seed(21)
    pose_p, hit_p = nl.render_helpers.track_frame(copy.deepcopy(f_p.pose), f_p, T._clone_ms(ms0), copy.deepcopy(dec0), crit, num_iterations=n_it, deterministic=True, loss_log=ll, **kw)
    print(f"it={n_it} ref losses {rc.rec}\n       our losses {ll}")
    print("   ref pose", pose_r.data.detach().cpu().numpy(), "\n   our pose", pose_p.data.detach().cpu().numpy())
    print("   diff", (pose_r.data.detach().cpu() - pose_p.data.detach().cpu()).numpy())

# single-iteration gradient: reference autograd vs kernels vs fp64 oracle
f = frame()
torch.manual_seed(21)
f.sample_rays(1024, track=True)
mask = f.sample_mask
pose = copy.deepcopy(f.pose).cuda(); pose.requires_grad_(True)
ray_dirs = f.rays_d[mask].unsqueeze(0).cuda()
pts = f.points.unsqueeze(1).cuda()[mask]; pc = f.pointsCos.unsqueeze(1).cuda()[mask]
rd = (ray_dirs.squeeze(0) @ pose.rotation().transpose(-1, -2)).unsqueeze(0)
ro = pose.translation().reshape(1, 1, -1).expand_as(rd).cuda().contiguous()
with H.pinned(ref):
    out = ref.orig["render_rays"](ro, rd, T._clone_ms(ms0), copy.deepcopy(dec0), 0.2 * VS, VS, T.TR, 20, T.MD, chunk_size=-2)
    hm = out["ray_mask"].view(1024); out["ray_mask"] = hm
    loss, _ = crit(out, pts, pc, weight_depth_loss=True)
loss.backward()
g_ref = pose.data.grad.detach().cpu().double() if pose.data.grad is not None else [p.grad for p in pose.parameters()][0].detach().cpu().double()
print("ref loss", float(loss), "grad", g_ref.numpy())
dev = torch.device("cuda")
m = nl.engine.MapState.from_map_states(T._clone_ms(ms0), dev)
cfg = nl.render_helpers._cfg(0.2 * VS, VS, T.MD, crit)
dirs = ray_dirs[0].contiguous(); cosv = pc.view(-1).contiguous(); gt = (torch.norm(pts, 2, -1) * cosv).contiguous()
pose6 = f.pose.data.detach().reshape(1, 6).cuda().contiguous()
for impl in ("tc", "simt"):
    os.environ["NL_MLP_IMPL"] = impl
    bufs = nl.engine.DecoderBuffers(copy.deepcopy(dec0), dev); bufs.refresh_transposes()
    eng = nl.engine.SDFEngine(1024, 1024 * 64, dev)
    eng.rays_from_poses(pose6, dirs, None)
    eng.forward_backward(m, bufs, 1024, cfg, gt, cosv, dir_local=dirs, ray_frame=None, n_frames=1, update_decoder=False, update_emb=False, update_pose=True, pose6=pose6, refresh_weights=False)
    st = eng.read_stats()
    g_our = eng.pose_grad[0].cpu().double()
    print(impl, "our loss", st.loss, "grad", g_our.numpy(), "samples", st.n_samples, "ref samples", int(out["valid_mask"].sum()))
os.environ.pop("NL_MLP_IMPL")
# fp64 oracle on the kernel's rays
from oracle import chain as OC
map_np = {"centres": m.centres.cpu().numpy(), "structure": m.structure.cpu().numpy(), "vertex_rows": m.vox2row.cpu().numpy().astype(np.int64)}
rays = (eng.ray_o[:1024].cpu().numpy().copy(), eng.ray_d[:1024].cpu().numpy().copy())
for dt in (torch.float32, torch.float64):
    dec_o = OC.Decoder(); dec_o.load_state_dict({k: v.cpu() for k, v in dec0.state_dict().items()}); dec_o = dec_o.to(dt)
    fr = [dict(pose=f.pose.data.detach().clone().to(dt).requires_grad_(), dirs=dirs.cpu().to(dt), points=pts.cpu().to(dt), cos=cosv.cpu().to(dt))]
    cfg_o = dict(step_size=0.2 * VS, voxel_size=VS, max_distance=T.MD, truncation=T.TR, max_depth=T.MD, fs_weight=1, sdf_weight=10000.0)
    l, o = OC.mapping_iteration(fr, map_np, m.emb.cpu().float().to(dt), dec_o, cfg_o, deterministic=True, rays_np=rays)
    l.backward()
    print(dt, "oracle loss", float(l), "grad", fr[0]["pose"].grad.double().numpy(), "samples", int(o["valid_mask"].sum()))
