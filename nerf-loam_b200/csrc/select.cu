// Per-iteration ray selection on the device: N distinct points of a scan, uniformly at random, in the scan's point order --
// the distribution of LidarFrame.sample_rays / sampling_without_replacement (src/lidarFrame.py:55-57, src/utils/sample_util.py:4-19:
// Gumbel top-k over constant weights = a uniform N-subset, returned as a mask, i.e. in point order) -- plus the gather of the
// selected rays' direction / range / cosine, as ONE launch.
//
// The torch formulation of the same thing (rand -> top-k of 131 072 keys -> sort -> three gathers) is ~10 kernels and 140-160 us,
// a third of a 2048-ray tracking iteration whose other 20 kernels take 270 us.  Here one block per frame keeps a bitmap of the
// scan's points in shared memory: every thread draws indices from a counter-based generator and claims them with atomicOr,
// redrawing when the bit was already taken (rejection of duplicates = sampling without replacement; whatever order the threads
// interleave in corresponds to some sequential order, so the subset is uniform); a block-wide scan of the words' popcounts then
// yields the chosen indices in ascending order together with their output slots.
#include "nl_cuda.cuh"

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

constexpr int SEL_THREADS = 1024;

__global__ void __launch_bounds__(SEL_THREADS) k_select_rays(int cap, int n_select, const int64_t *__restrict__ n_points, const uint32_t *__restrict__ seed_dev,
                                                              uint32_t seed_host, const float *__restrict__ dirs_all, const float *__restrict__ gt_all,
                                                              const float *__restrict__ cos_all, float *__restrict__ dirs, float *__restrict__ gt,
                                                              float *__restrict__ cosv, int32_t *__restrict__ idx_out) {
    extern __shared__ uint32_t bits[];                 // ceil(cap / 32) words
    __shared__ int s_warp[SEL_THREADS / 32];
    const int f = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int n = (int)min((long long)n_points[f], (long long)cap);
    const int N = min(n_select, n);
    const int words = (cap + 31) >> 5;
    for (int i = t; i < words; i += SEL_THREADS) bits[i] = 0u;
    __syncthreads();
    uint32_t seed = seed_dev ? *seed_dev : seed_host;
    seed = mix32(seed ^ 0x9E3779B9u * (uint32_t)(f + 1));
    // more than half of the points wanted: draw the ones to LEAVE OUT instead (the complement of a uniform subset is uniform),
    // so the rejection loop never fishes for the last few free bits
    const bool invert = N > n / 2;
    const int draws = invert ? n - N : N;
    for (int j = t; j < draws; j += SEL_THREADS) {
        uint32_t ctr = 0;
        while (true) {
            const uint32_t r = mix32(seed ^ mix32((uint32_t)j * 0x85EBCA77u + ctr * 0xC2B2AE3Du + 0x27D4EB2Fu));
            ++ctr;
            const uint32_t i = (uint32_t)(((unsigned long long)r * (unsigned long long)n) >> 32);      // uniform in [0, n)
            const uint32_t bit = 1u << (i & 31);
            if (!(atomicOr(&bits[i >> 5], bit) & bit)) break;                                          // claimed a free point
        }
    }
    __syncthreads();
    if (invert) {
        for (int i = t; i < words; i += SEL_THREADS) {
            const int base = i * 32;
            const uint32_t valid = base + 32 <= n ? 0xffffffffu : (base < n ? (1u << (n - base)) - 1u : 0u);
            bits[i] = ~bits[i] & valid;
        }
        __syncthreads();
    }
    // ascending order: exclusive scan of the words' popcounts (each thread owns a contiguous run of words)
    const int per = (words + SEL_THREADS - 1) / SEL_THREADS;
    const int w0 = t * per, w1 = min(w0 + per, words);
    int mine = 0;
    for (int i = w0; i < w1; ++i) mine += __popc(bits[i]);
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += y;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    int pre = 0;
    for (int q = 0; q < w; ++q) pre += s_warp[q];
    int slot = pre + incl - mine;
    const size_t in0 = (size_t)f * cap, out0 = (size_t)f * n_select;
    for (int i = w0; i < w1; ++i) {
        uint32_t b = bits[i];
        while (b) {
            const int k = __ffs(b) - 1;
            b &= b - 1;
            const int p = i * 32 + k;
            const size_t o = out0 + slot;
            dirs[o * 3] = dirs_all[(in0 + p) * 3]; dirs[o * 3 + 1] = dirs_all[(in0 + p) * 3 + 1]; dirs[o * 3 + 2] = dirs_all[(in0 + p) * 3 + 2];
            gt[o] = gt_all[in0 + p];
            cosv[o] = cos_all[in0 + p];
            if (idx_out) idx_out[o] = p;
            ++slot;
        }
    }
}

NlPerDevice g_sel_attr;

}  // namespace

extern "C" int nl_select_rays(int n_frames, int cap, int n_select, const int64_t *d_n_points, const uint32_t *d_seed, uint32_t seed,
                              const float *d_dirs_all, const float *d_gt_all, const float *d_cos_all, float *d_dirs, float *d_gt, float *d_cos,
                              int32_t *d_idx, void *stream) {
    if (n_frames <= 0 || cap <= 0 || n_select <= 0) return nl_set_error("nl_select_rays: sizes must be positive");
    if (cap > (1 << 20)) return nl_set_error("nl_select_rays: at most 2^20 points per scan (bitmap in shared memory)");
    if (!d_n_points || !d_dirs_all || !d_gt_all || !d_cos_all || !d_dirs || !d_gt || !d_cos) return nl_set_error("nl_select_rays: null pointer");
    const int smem = ((cap + 31) >> 5) * 4;
    if (smem > 48 * 1024) {
        const cudaError_t e = g_sel_attr.once([] { return cudaFuncSetAttribute(k_select_rays, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); });
        if (e != cudaSuccess) return nl_set_error_code(NL_ERR_CUDA, cudaGetErrorString(e));
    }
    k_select_rays<<<n_frames, SEL_THREADS, smem, (cudaStream_t)stream>>>(cap, n_select, d_n_points, d_seed, seed, d_dirs_all, d_gt_all, d_cos_all, d_dirs,
                                                                         d_gt, d_cos, d_idx);
    NL_CHECK_LAUNCH("nl_select_rays");
    return NL_OK;
}
