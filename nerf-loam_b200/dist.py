"""Collectives of the ray-sharded data-parallel hot path (SURVEY.md section 8 e).

Rays are independent up to the loss, whose normalisation is global (criterion.py:84-100): the mean runs over R_hit x S_max
cells and the class-balance weights use global mask counts.  So per iteration there are exactly

  1. ONE tiny all-reduce before the backward pass: every statistic that has to become global is packed into one f64 vector
     reduced with SUM -- counters and sums as they are; S_max (needs MAX) and the kernel error bits (need OR) in one slot per
     rank, so that a SUM transports them (nl_stats_pack / nl_stats_unpack, include/nerfloam_b200.h section 9);
  2. ONE all-reduce after it over a flat fp32 buffer [loss sums (hi, lo) | pose accumulators | embedding-gradient table]
     (engine.SDFEngine.gradflat);
  3. while the decoder is trained, ONE more over its flat gradient buffer (engine.DecoderBuffers.gradflat), issued on the
     side stream where the weight-gradient kernels run.

The octree, embedding table and decoder are replicated and every rank applies the identical Adam step.
On CUDA tensors packing runs as the C-ABI kernels; the same layout in torch ops serves CPU tensors (the gloo tests).
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _capi

# slots inside nl_render_stats (checked against the C struct at import of _capi)
_I32_N_HIT, _I32_MAX_SAMPLES, _I32_ERROR = 0, 2, 4   # int32 slots
_I64_COUNTERS = slice(4, 10)                         # cnt_fs_valid .. pad_sdf_nsamp
_F64_PAD = slice(10, 12)                             # pad_sdf_d2, pad_sdf_d2_nsamp
_F64_LOSS_SUMS = slice(16, 18)                       # fs_sum, sdf_sum
FIXED = _capi.STATS_PACK_FIXED


def stats_words(world):
    return FIXED + 2 * world


def pack_stats(stats_u8, buf, rank, world, phase=0):
    """phase 0: buf f64[stats_words(world)] <- sample statistics; phase 1: buf f32[4] <- squared-error sums as (hi, lo) pairs."""
    if stats_u8.is_cuda:
        _capi.check(_capi.lib().nl_stats_pack(_capi.ptr(stats_u8), _capi.ptr(buf), int(rank), int(world), int(phase), _capi.stream_ptr()),
                    "nl_stats_pack")
        _capi.LAUNCHES += 1
        return
    i32, i64, f64 = stats_u8.view(torch.int32), stats_u8.view(torch.int64), stats_u8.view(torch.float64)
    if phase == 0:
        buf.zero_()
        buf[0:6] = i64[_I64_COUNTERS].double()
        buf[6:8] = f64[_F64_PAD]
        buf[8] = float(i32[_I32_N_HIT])
        buf[FIXED + rank] = float(i32[_I32_MAX_SAMPLES])
        buf[FIXED + world + rank] = float(i32[_I32_ERROR])
    else:
        s = f64[_F64_LOSS_SUMS]
        hi = s.float()
        buf[0], buf[2] = hi[0], hi[1]
        lo = (s - hi.double()).float()
        buf[1], buf[3] = lo[0], lo[1]


def unpack_stats(stats_u8, buf, world, phase=0, fs_weight=0.0, sdf_weight=0.0):
    """Inverse of pack_stats after the SUM all-reduce; phase 0 on CUDA also re-derives the loss constants (nl_loss_prepare)."""
    if stats_u8.is_cuda:
        _capi.check(_capi.lib().nl_stats_unpack(_capi.ptr(stats_u8), _capi.ptr(buf), int(world), int(phase), float(fs_weight), float(sdf_weight),
                                                _capi.stream_ptr()), "nl_stats_unpack")
        _capi.LAUNCHES += 2 if phase == 0 else 1
        return
    i32, i64, f64 = stats_u8.view(torch.int32), stats_u8.view(torch.int64), stats_u8.view(torch.float64)
    if phase == 0:
        i64[_I64_COUNTERS] = buf[0:6].long()
        f64[_F64_PAD] = buf[6:8]
        i32[_I32_N_HIT] = int(buf[8])
        i32[_I32_MAX_SAMPLES] = int(buf[FIXED:FIXED + world].max())
        err = 0
        for e in buf[FIXED + world:FIXED + 2 * world].tolist():
            err |= int(e)
        i32[_I32_ERROR] = err
    else:
        f64[_F64_LOSS_SUMS] = torch.stack([buf[0].double() + buf[1].double(), buf[2].double() + buf[3].double()])


def allreduce_sample_stats(stats_u8, group=None, buf=None, fs_weight=0.0, sdf_weight=0.0):
    """Make the loss-mask statistics global with one collective (counters SUM, R_hit SUM, S_max MAX, error bits OR)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if buf is None:
        buf = torch.empty(stats_words(world), dtype=torch.float64, device=stats_u8.device)
    pack_stats(stats_u8, buf, rank, world, 0)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    unpack_stats(stats_u8, buf, world, 0, fs_weight, sdf_weight)
    return buf


def allreduce_flat(flat, group=None):
    """One SUM all-reduce over a flat gradient buffer."""
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)


def allreduce_grads_with_loss(stats_u8, gradflat, group=None):
    """Post-backward exchange: the squared-error sums ride in the 4-float header of the flat fp32 gradient buffer
    [header(16) | pose accumulators | embedding-gradient table], so loss and gradients are ONE collective."""
    world = dist.get_world_size(group)
    pack_stats(stats_u8, gradflat[:4], 0, world, 1)
    dist.all_reduce(gradflat, op=dist.ReduceOp.SUM, group=group)
    unpack_stats(stats_u8, gradflat[:4], world, 1)


def shard_bounds(n, rank, world):
    """Contiguous ray shard [lo, hi) of rank `rank` (SURVEY.md 8 e: rank r <- rays[r*R/W : (r+1)*R/W])."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


class PeerReduceAdam:
    """Embedding-gradient reduction fused with the table's Adam step over NVLink peer memory (csrc/peer.cu): per iteration ONE
    kernel does reduce-scatter -> Adam -> all-gather -- through the NVLS multicast addresses of the two symmetric buffers when the
    fabric offers them (multimem.ld_reduce / multimem.st, the sum happens inside the NVSwitch), through plain P2P loads and stores
    otherwise -- instead of an NCCL all-reduce followed by a separate Adam launch.  Optimiser state is sharded over the ranks.

    The flat fp32 gradient buffer has the engine's layout [header(16) | pose accumulators | table V*16] (SDFEngine.adopt_gradflat
    makes the scatter kernels write straight into it) and the bf16 table is the map's embedding table (MapState.emb must be
    `self.param`).  step() = barrier, small reduce (loss sums + pose accumulators), fused kernel, barrier."""

    def __init__(self, group, device, n_rows, n_frames, lr, betas=(0.9, 0.999), eps=1e-8):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n_rows, self.n_frames = int(n_rows), int(n_frames)
        self.npose = (n_frames * 12 + 15) // 16 * 16
        self.n_hdr = 16 + self.npose
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.grad = symm.empty(self.n_hdr + self.n_rows * 16, dtype=torch.float32, device=device)
        self.param = symm.empty((self.n_rows, 16), dtype=torch.bfloat16, device=device)
        self.grad.zero_(); self.param.zero_()
        self.h_grad = symm.rendezvous(self.grad, self.group)
        self.h_param = symm.rendezvous(self.param, self.group)
        off = self.n_hdr * 4
        i64 = lambda xs: torch.tensor(list(xs), dtype=torch.int64, device=device)
        self._hdr_peers = i64(self.h_grad.buffer_ptrs)
        self._grad_peers = i64(p + off for p in self.h_grad.buffer_ptrs)
        self._param_peers = i64(self.h_param.buffer_ptrs)
        mc_g, mc_p = int(self.h_grad.multicast_ptr or 0), int(self.h_param.multicast_ptr or 0)
        self.multicast = bool(mc_g and mc_p)
        self._grad_mc = C.c_void_p(mc_g + off) if self.multicast else None
        self._hdr_mc = C.c_void_p(mc_g) if self.multicast else None
        self._param_mc = C.c_void_p(mc_p) if self.multicast else None
        self.m = torch.zeros((self.n_rows, 16), dtype=torch.bfloat16, device=device)
        self.v = torch.zeros((self.n_rows, 16), dtype=torch.bfloat16, device=device)
        self.hdr_out = torch.zeros(self.n_hdr, dtype=torch.float32, device=device)

    def step(self, stats_u8, ctl, pose_acc=None, defer_barrier=False):
        """After the local scatter: make loss sums / pose accumulators / embedding gradients global and apply the table's Adam step.
        defer_barrier: leave out the closing barrier (every rank's slice has landed in every table); the caller runs wait_params()
        before the table is read again -- after the next iteration's traversal and sampling, which do not touch it."""
        lib, st = _capi.lib(), _capi.stream_ptr()
        pack_stats(stats_u8, self.grad[:4], 0, self.world, 1)                       # loss sums ride in the header
        self.h_grad.barrier(channel=0)                                              # every rank's scatter + header are complete
        _capi.check(lib.nl_peer_reduce_adam_bf16(self.n_rows * 16, self.rank, self.world, _capi.ptr(self._grad_peers), self._grad_mc,
                                                 _capi.ptr(self._param_peers), self._param_mc, _capi.ptr(self.m), _capi.ptr(self.v), self.lr,
                                                 self.betas[0], self.betas[1], self.eps, _capi.ptr(ctl), self.n_hdr, _capi.ptr(self._hdr_peers),
                                                 self._hdr_mc, _capi.ptr(self.hdr_out), st), "nl_peer_reduce_adam_bf16")
        _capi.LAUNCHES += 1
        unpack_stats(stats_u8, self.hdr_out[:4], self.world, 1)                     # local results: no need to wait for the peers
        if pose_acc is not None:
            pose_acc.copy_(self.hdr_out[16:16 + pose_acc.numel()].view(pose_acc.shape))
        self._pending_barrier = True
        if not defer_barrier:
            self.wait_params()

    def wait_params(self):
        if getattr(self, "_pending_barrier", False):
            self.h_param.barrier(channel=1)
            self._pending_barrier = False


class PeerStats:
    """The pre-backward statistics exchange through symmetric memory: pack into this rank's buffer, cross-rank barrier, every rank
    sums all the vectors itself (nl_stats_unpack_peers) -- three launches on the caller's stream, no collective-library call."""

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.buf = symm.empty(stats_words(self.world), dtype=torch.float64, device=device)
        self.buf.zero_()
        self.h = symm.rendezvous(self.buf, self.group)
        self._peers = torch.tensor(list(self.h.buffer_ptrs), dtype=torch.int64, device=device)

    def exchange(self, stats_u8, fs_weight, sdf_weight):
        pack_stats(stats_u8, self.buf, self.rank, self.world, 0)
        self.h.barrier(channel=0)
        _capi.check(_capi.lib().nl_stats_unpack_peers(_capi.ptr(stats_u8), _capi.ptr(self._peers), self.world, float(fs_weight), float(sdf_weight),
                                                      _capi.stream_ptr()), "nl_stats_unpack_peers")
        _capi.LAUNCHES += 2
