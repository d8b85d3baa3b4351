// Drop-in replacements for the two live kernels of the reference's `grid` CUDA extension
// (third_party/sparse_voxels/src/binding.cpp:12-20): svo_intersect and inverse_cdf_sampling.
// Same tensor layouts in and out, results bit-identical to the reference kernels on the same GPU.
#include "sampler.cuh"
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

// grid.inverse_cdf_sampling: one thread per ray, grid = (ceil(num_rays/128), b); outputs padded with -1 / 0 like sample.cpp:82-90.
// The walk itself is nl_inverse_cdf_walk (sampler.cuh), the same routine the fused sampler runs; this wrapper only adapts the
// reference's [b, num_rays, max_hits] / [b, num_rays, max_steps] tensors and states its two tail quirks in their original form:
// the tail is flushed only while  num_rays > ray * max_hits + bin,  and its continuation test reads pts_idx[bin] of the
// batch's FIRST ray (sample_gpu.cu:224-231).
struct RayBinsGlobal {
    const int32_t *i; const float *a, *b, *p;
    __device__ __forceinline__ int idx(int c) const { return i[c]; }
    __device__ __forceinline__ float lo(int c) const { return a[c]; }
    __device__ __forceinline__ float hi(int c) const { return b[c]; }
    __device__ __forceinline__ float prob(int c) const { return p[c]; }
};

__global__ void __launch_bounds__(128) k_inverse_cdf_dropin(int num_rays, int max_hits, int max_steps, float fixed_step_size,
                                                             const int32_t *__restrict__ pts_idx, const float *__restrict__ min_depth,
                                                             const float *__restrict__ max_depth, const float *__restrict__ uniform_noise,
                                                             const float *__restrict__ probs, const float *__restrict__ steps,
                                                             int32_t *__restrict__ sampled_idx, float *__restrict__ sampled_depth,
                                                             float *__restrict__ sampled_dists) {
    const int ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= num_rays) return;
    const size_t batch_hits = (size_t)blockIdx.y * num_rays * max_hits, batch_steps = (size_t)blockIdx.y * num_rays * max_steps;
    const int32_t *first_ray_idx = pts_idx + batch_hits;
    const size_t h0 = batch_hits + (size_t)ray * max_hits, s0 = batch_steps + (size_t)ray * max_steps;
    const RayBinsGlobal bins = {pts_idx + h0, min_depth + h0, max_depth + h0, probs + h0};
    int32_t *o_idx = sampled_idx + s0;
    float *o_depth = sampled_depth + s0, *o_len = sampled_dists + s0;
    for (int s = 0; s < max_steps; ++s) { o_idx[s] = -1; o_depth[s] = 0.f; o_len[s] = 0.f; }
    const float *nz = uniform_noise + s0;
    const int flat = ray * max_hits;
    int n = 0;
    nl_inverse_cdf_walk(
        max_hits, bins, steps[(size_t)blockIdx.y * num_rays + ray], fixed_step_size, [&](int i) { return nz[i]; },
        [&](int vox, float z0, float z1) {
            o_idx[n] = vox;
            o_len[n] = __fsub_rn(z1, z0);
            o_depth[n] = __fmul_rn(__fadd_rn(z1, z0), 0.5f);
            ++n;
        },
        [&](int bin) { return num_rays > flat + bin; }, [&](int bin) { return first_ray_idx[bin]; });
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
    k_inverse_cdf_dropin<<<grid, 128, 0, (cudaStream_t)stream>>>(num_rays, max_hits, max_steps, fixed_step_size, pts_idx, min_depth, max_depth,
                                                                 noise, probs, steps, sampled_idx, sampled_depth, sampled_dists);
    NL_CHECK_LAUNCH("nl_inverse_cdf_sampling");
    return NL_OK;
}
