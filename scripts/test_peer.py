"""torchrun --nproc-per-node N scripts/test_peer.py : the fused peer-memory reduce+Adam path against the NCCL all-reduce + Adam path."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerfloam_b200 as nl
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cap = nl._capi
V, F = 24099, 1
torch.manual_seed(0)
p0 = (torch.randn(V, 16, device=dev) * 0.01).to(torch.bfloat16)          # identical on all ranks
ctl = torch.zeros(cap.CTL_WORDS, dtype=torch.int32, device=dev)
peer = nl.dist.PeerReduceAdam(dist.group.WORLD, dev, V, F, lr=0.01)
peer.param.copy_(p0)
print(f"rank {rank}: multicast={peer.multicast}", flush=True)
# reference path: NCCL all-reduce + fused Adam kernel
p_ref = p0.clone()
g_ref = torch.zeros(V * 16 + 32, device=dev)
opt = nl.engine.FusedAdam([dict(param=p_ref, grad=g_ref[32:].view(V, 16), lr=0.01)], ctl=ctl)
stats = torch.zeros(nl.engine.STATS_BYTES, dtype=torch.uint8, device=dev)
for step in range(1, 4):
    torch.manual_seed(100 * step + rank)
    g = torch.randn(V, 16, device=dev) * (0.1 if step < 3 else 1e-3)
    g[::5] = 0
    ctl[cap.CTL_ADAM_STEP] = step
    # fused path
    peer.grad.zero_(); peer.grad[peer.n_hdr:].view(V, 16).copy_(g); peer.grad[16:28] = float(rank + 1)
    stats.view(torch.float64)[16:18] = torch.tensor([1.5 + rank, 2.25 * (rank + 1)], dtype=torch.float64, device=dev)
    st2 = stats.clone()
    pose = torch.zeros(1, 12, device=dev)
    peer.step(st2, ctl, pose)
    # NCCL path
    g_ref.zero_(); g_ref[32:].view(V, 16).copy_(g)
    dist.all_reduce(g_ref)
    opt.step()
    torch.cuda.synchronize()
    d = (peer.param.float() - p_ref.float()).abs()
    sums = st2.view(torch.float64)[16:18].tolist()
    print(f"rank {rank} step {step}: max|param diff| {float(d.max()):.3e} frac differing {float((d > 0).float().mean()):.2e} pose {pose[0,0].item()} loss sums {sums}", flush=True)
    assert float((d > 0).float().mean()) < 2e-3 and float(d.max()) < 2e-3
    assert pose[0, 0].item() == world * (world + 1) / 2
    # replicas identical
    chk = peer.param.float().sum().double().reshape(1).clone(); lo = chk.clone(); hi = chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert float(lo) == float(hi)
# timing
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
t_peer = timed(lambda: peer.step(stats, ctl, None))
def nccl():
    dist.all_reduce(g_ref); opt.step()
t_nccl = timed(nccl)
if rank == 0:
    print(f"RESULT world={world} multicast={peer.multicast} fused peer reduce+adam {t_peer:.1f} us/step vs nccl allreduce + adam {t_nccl:.1f} us/step ({V} rows)", flush=True)
dist.barrier(); dist.destroy_process_group()
