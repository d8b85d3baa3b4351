"""Collectives of the ray-sharded data-parallel hot path (SURVEY.md section 8 e).

Rays are independent up to the loss, whose normalisation is global (criterion.py:84-100): the mean runs
over R_hit x S_max cells and the class-balance weights use global mask counts.  So one tiny exchange
before the backward pass (raw counters SUM, S_max MAX) and one gradient reduction after it; the octree,
embedding table and decoder are replicated and every rank applies the identical Adam step.
The functions work on the raw nl_render_stats byte block (any device: NCCL on GPU, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

# byte offsets inside nl_render_stats (checked against the C struct at import of _capi)
_I32_N_HIT, _I32_MAX_SAMPLES = 0, 2          # int32 slots
_I64_COUNTERS = slice(4, 10)                 # cnt_fs_valid .. pad_sdf_nsamp
_F64_PAD = slice(10, 12)                     # pad_sdf_d2, pad_sdf_d2_nsamp
_F64_LOSS_SUMS = slice(16, 18)               # fs_sum, sdf_sum


def allreduce_sample_stats(stats_u8, group=None):
    """Make the loss-mask statistics global: counters SUM, R_hit SUM, S_max MAX.  nl_loss_prepare must be
    re-run afterwards."""
    i32, i64, f64 = stats_u8.view(torch.int32), stats_u8.view(torch.int64), stats_u8.view(torch.float64)
    dist.all_reduce(i64[_I64_COUNTERS], op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(f64[_F64_PAD], op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(i32[_I32_N_HIT:_I32_N_HIT + 1], op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(i32[_I32_MAX_SAMPLES:_I32_MAX_SAMPLES + 1], op=dist.ReduceOp.MAX, group=group)


def allreduce_loss_sums(stats_u8, group=None):
    dist.all_reduce(stats_u8.view(torch.float64)[_F64_LOSS_SUMS], op=dist.ReduceOp.SUM, group=group)


def allreduce_grads(tensors, group=None):
    """Gradient reduction: embedding-gradient table [V,16] fp32, decoder grads, pose accumulators [F,12] -- all fp32, one
    coalesced all-reduce (a single grouped NCCL launch instead of eight latency-bound ones)."""
    ts = [t for t in tensors if t is not None]
    if not ts:
        return
    if len(ts) > 1 and all(t.dtype == ts[0].dtype for t in ts):
        try:
            from torch.distributed.distributed_c10d import _coalescing_manager
            with _coalescing_manager(group=group, async_ops=False):
                for t in ts:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return
        except (ImportError, RuntimeError, ValueError, NotImplementedError):
            from torch.distributed import distributed_c10d as _c10d
            _c10d._world.pg_coalesce_state.pop(group or _c10d._get_default_group(), None)   # leave no half-open coalescing state
    for t in ts:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def shard_bounds(n, rank, world):
    """Contiguous ray shard [lo, hi) of rank `rank` (SURVEY.md 8 e: rank r <- rays[r*R/W : (r+1)*R/W])."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi
