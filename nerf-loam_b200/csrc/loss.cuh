// The loss read-out of one iteration (criterion.py:84-100 after the masked sums): shared by nl_loss_finalize (render.cu) and the fused
// pose step (pose.cu).
#pragma once
#include "nl_cuda.cuh"

__device__ __forceinline__ void nl_loss_finalize_dev(nl_render_stats *s, float fs_weight, float sdf_weight) {
    const double N = (double)((long long)s->n_hit_rays * (long long)s->max_samples);
    s->fs_loss = (float)((s->fs_sum + (double)s->pad_fs_sum) / N) * s->w_fs;
    s->sdf_loss = (float)((s->sdf_sum + (double)s->pad_sdf_sum) / N) * s->w_sdf;
    s->loss = fs_weight * s->fs_loss + sdf_weight * s->sdf_loss;
}
