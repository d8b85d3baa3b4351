// Drop-in replacements for the two live kernels of the reference's `grid` CUDA extension
// (third_party/sparse_voxels/src/binding.cpp:12-20): svo_intersect and inverse_cdf_sampling.
// Same tensor layouts in and out, results bit-identical to the reference kernels on the same GPU.
#include "traverse.cuh"

namespace {

// grid = (ceil(m/128), b); one thread per ray (intersect_gpu.cu:193-272 uses one block per batch).
__global__ void __launch_bounds__(128) svo_intersect_kernel(int n, int m, float voxelsize, int n_max,
                                                             const float *__restrict__ ray_start,
                                                             const float *__restrict__ ray_dir,
                                                             const float *__restrict__ points,
                                                             const int32_t *__restrict__ children, int32_t *__restrict__ idx,
                                                             float *__restrict__ min_depth, float *__restrict__ max_depth) {
    const int batch = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    points += (size_t)batch * n * 3;
    children += (size_t)batch * n * 9;
    const size_t r = (size_t)batch * m + j;
    const float *o = ray_start + r * 3, *d = ray_dir + r * 3;
    int32_t *oi = idx + r * n_max;
    float *omin = min_depth + r * n_max, *omax = max_depth + r * n_max;
    for (int l = 0; l < n_max; ++l) { oi[l] = -1; omin[l] = 0.f; omax[l] = 0.f; }  // intersect.cpp:98-106 + :227-230
    const NlRay ray = nl_make_ray(o[0], o[1], o[2], d[0], d[1], d[2]);
    const float half_voxel = voxelsize * 0.5f;
    int cnt = 0;
    nl_traverse(ray, points, children, half_voxel, n_max, [&](int k, float lo, float hi) {
        oi[cnt] = k; omin[cnt] = lo; omax[cnt] = hi; ++cnt;
    });
}

// sample_gpu.cu:133-239, one thread per ray, grid = (ceil(num_rays/128), b).
// `min + u*(max-min)` is one FMA in the reference binary (nvcc -fmad default), written explicitly here.
__global__ void __launch_bounds__(128) inverse_cdf_sampling_kernel(
    int num_rays, int max_hits, int max_steps, float fixed_step_size, const int32_t *__restrict__ pts_idx,
    const float *__restrict__ min_depth, const float *__restrict__ max_depth, const float *__restrict__ uniform_noise,
    const float *__restrict__ probs, const float *__restrict__ steps, int32_t *__restrict__ sampled_idx,
    float *__restrict__ sampled_depth, float *__restrict__ sampled_dists) {
    const int batch = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= num_rays) return;
    pts_idx += (size_t)batch * num_rays * max_hits;
    min_depth += (size_t)batch * num_rays * max_hits;
    max_depth += (size_t)batch * num_rays * max_hits;
    probs += (size_t)batch * num_rays * max_hits;
    steps += (size_t)batch * num_rays;
    uniform_noise += (size_t)batch * num_rays * max_steps;
    sampled_idx += (size_t)batch * num_rays * max_steps;
    sampled_depth += (size_t)batch * num_rays * max_steps;
    sampled_dists += (size_t)batch * num_rays * max_steps;

    const int H = j * max_hits;
    const size_t K = (size_t)j * max_steps;
    for (int s = 0; s < max_steps; ++s) { sampled_idx[K + s] = -1; sampled_depth[K + s] = 0.f; sampled_dists[K + s] = 0.f; }  // sample.cpp:82-90
    int curr_bin = 0, s = 0;
    float curr_min_depth = min_depth[H];
    float curr_max_depth = max_depth[H];
    float curr_min_cdf = 0.f;
    float curr_max_cdf = probs[H];
    float step_size = __frcp_rn(steps[j]);  // (float)(1.0 / (double)x) == correctly rounded 1/x
    float z_low = curr_min_depth;
    const int total_steps = (int)ceilf(steps[j]);
    bool done = false;
    if (fixed_step_size > 0.0f) step_size = fixed_step_size;
    for (int curr_step = 0; curr_step < total_steps; ++curr_step) {
        const float curr_cdf = __fmul_rn(__fadd_rn((float)curr_step, uniform_noise[K + curr_step]), step_size);
        while (curr_cdf > curr_max_cdf) {
            sampled_idx[K + s] = pts_idx[H + curr_bin];
            sampled_dists[K + s] = __fsub_rn(curr_max_depth, z_low);
            sampled_depth[K + s] = __fmul_rn(__fadd_rn(curr_max_depth, z_low), 0.5f);
            ++curr_bin;
            ++s;
            if ((curr_bin >= max_hits) || (pts_idx[H + curr_bin] == -1)) { done = true; break; }
            curr_min_depth = min_depth[H + curr_bin];
            curr_max_depth = max_depth[H + curr_bin];
            curr_min_cdf = curr_max_cdf;
            curr_max_cdf = __fadd_rn(curr_max_cdf, probs[H + curr_bin]);
            z_low = curr_min_depth;
        }
        if (done) break;
        const float u = __fdiv_rn(__fsub_rn(curr_cdf, curr_min_cdf), __fsub_rn(curr_max_cdf, curr_min_cdf));
        const float z = __fmaf_rn(u, __fsub_rn(curr_max_depth, curr_min_depth), curr_min_depth);
        sampled_idx[K + s] = pts_idx[H + curr_bin];
        sampled_dists[K + s] = __fsub_rn(z, z_low);
        sampled_depth[K + s] = __fmul_rn(__fadd_rn(z, z_low), 0.5f);
        z_low = z;
        ++s;
    }
    // sample_gpu.cu:224-238, quirks kept: `num_rays > H + curr_bin` and `pts_idx[curr_bin]` (no +H)
    while ((z_low < curr_max_depth) && (!done) && (num_rays > (H + curr_bin))) {
        sampled_idx[K + s] = pts_idx[H + curr_bin];
        sampled_dists[K + s] = __fsub_rn(curr_max_depth, z_low);
        sampled_depth[K + s] = __fmul_rn(__fadd_rn(curr_max_depth, z_low), 0.5f);
        ++curr_bin;
        ++s;
        if ((curr_bin >= max_hits) || (pts_idx[curr_bin] == -1)) break;
        curr_min_depth = min_depth[H + curr_bin];
        curr_max_depth = max_depth[H + curr_bin];
        z_low = curr_min_depth;
    }
}

}  // namespace

extern "C" int nl_svo_intersect(int b, int n, int m, float voxelsize, int n_max, const float *ray_start, const float *ray_dir,
                                const float *points, const int32_t *children, int32_t *idx, float *min_depth,
                                float *max_depth, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || n_max <= 0) return nl_set_error("nl_svo_intersect: bad sizes");
    if (b == 0 || m == 0) return NL_OK;
    if (!ray_start || !ray_dir || !points || !children || !idx || !min_depth || !max_depth)
        return nl_set_error("nl_svo_intersect: null pointer");
    if (b > 65535) return nl_set_error("nl_svo_intersect: more than 65535 batches");
    dim3 grid(nl_div_up(m, 128), b);
    svo_intersect_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(n, m, voxelsize, n_max, ray_start, ray_dir, points, children,
                                                                 idx, min_depth, max_depth);
    NL_CHECK_LAUNCH("nl_svo_intersect");
    return NL_OK;
}

extern "C" int nl_inverse_cdf_sampling(int b, int num_rays, int max_hits, int max_steps, float fixed_step_size,
                                       const int32_t *pts_idx, const float *min_depth, const float *max_depth,
                                       const float *noise, const float *probs, const float *steps, int32_t *sampled_idx,
                                       float *sampled_depth, float *sampled_dists, void *stream) {
    if (b < 0 || num_rays < 0 || max_hits <= 0 || max_steps <= 0) return nl_set_error("nl_inverse_cdf_sampling: bad sizes");
    if (b == 0 || num_rays == 0) return NL_OK;
    if (!pts_idx || !min_depth || !max_depth || !noise || !probs || !steps || !sampled_idx || !sampled_depth || !sampled_dists)
        return nl_set_error("nl_inverse_cdf_sampling: null pointer");
    if (b > 65535) return nl_set_error("nl_inverse_cdf_sampling: more than 65535 batches");
    dim3 grid(nl_div_up(num_rays, 128), b);
    inverse_cdf_sampling_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(num_rays, max_hits, max_steps, fixed_step_size, pts_idx,
                                                                        min_depth, max_depth, noise, probs, steps, sampled_idx,
                                                                        sampled_depth, sampled_dists);
    NL_CHECK_LAUNCH("nl_inverse_cdf_sampling");
    return NL_OK;
}
