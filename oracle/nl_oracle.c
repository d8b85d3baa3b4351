/*
 * nl_oracle.c -- CPU restatement of the NeRF-LOAM hot-path integer/index kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (nerf-loam_b200/, bench.py's GPU arm)
 * may link, load or call this file.  It exists so that tests/ can check the CUDA kernels
 * against an independent, scalar, easy-to-read statement of what the reference computes.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 * Pinning status:
 *   - octree (nlo_octree_*):  pinned bit-exact against the compiled, unmodified reference
 *     `svo` (oracle/_ref/svo, tests/test_oracle_vs_reference.py) and frozen in tests/golden/.
 *   - svo_intersect / inverse_cdf_sampling: the reference implementation is CUDA-only, so
 *     here (no GPU) it is pinned only through hand-checked small cases; on the GPU box
 *     tests/test_gpu_reference_grid.py pins it against the compiled, unmodified reference
 *     `grid` (oracle/_ref/grid).
 *
 * Build:  gcc -O2 -ffp-contract=off -shared -fPIC -o _build/libnl_oracle.so nl_oracle.c -lm
 * (-ffp-contract=off: every fused multiply-add below is written explicitly with fmaf()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================================
 * 1. Sparse voxel octree  (third_party/sparse_octree/include/octree.h:12-129,
 *    src/octree.cpp:34-111 insert, :151-171 find_octant, :293-342 export,
 *    include/utils.h:64-109 Morton encode/decode)
 * ====================================================================================== */

enum { NLO_NONLEAF = -1, NLO_SURFACE = 0, NLO_FEATURE = 1 }; /* octree.h:5-10 */

typedef struct nlo_octant {
    uint64_t code;              /* octree.h:51 */
    unsigned side;              /* octree.h:53 */
    int index;                  /* octree.h:57: creation counter */
    int is_leaf;
    int type;
    struct nlo_octant *child[8]; /* octree.h:61, addressed as x + 2y + 4z (octree.h:42-45) */
} nlo_octant;

typedef struct {
    int size, max_level;
    int next_index;             /* octree.h:62 / octree.cpp:9 (process-global there) */
    nlo_octant *root;
} nlo_octree;

/* utils.h:64-74 */
static uint64_t nlo_expand(uint64_t value) {
    uint64_t x = value & 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}
/* utils.h:76-86 */
static uint64_t nlo_compact(uint64_t value) {
    uint64_t x = value & 0x1249249249249249ULL;
    x = (x | x >> 2) & 0x10c30c30c30c30c3ULL;
    x = (x | x >> 4) & 0x100f00f00f00f00fULL;
    x = (x | x >> 8) & 0x1f0000ff0000ffULL;
    x = (x | x >> 16) & 0x1f00000000ffffULL;
    x = (x | x >> 32) & 0x1fffff;
    return x;
}
/* utils.h:41-62: MASK[i] keeps the top 3*(i+1) Morton bits below bit 63 */
static uint64_t nlo_mask(int i) {
    uint64_t m = 0x7000000000000000ULL, acc = m;
    for (int k = 1; k <= i; ++k) acc |= (m >> (3 * k));
    return acc;
}
/* utils.h:88-109 */
uint64_t nlo_encode(int x, int y, int z) {
    uint64_t code = nlo_expand((uint64_t)(long long)x) | (nlo_expand((uint64_t)(long long)y) << 1) |
                    (nlo_expand((uint64_t)(long long)z) << 2);
    return code & nlo_mask(20);
}
void nlo_decode(uint64_t code, int out[3]) {
    out[0] = (int)nlo_compact(code >> 0);
    out[1] = (int)nlo_compact(code >> 1);
    out[2] = (int)nlo_compact(code >> 2);
}

static nlo_octant *nlo_new_octant(nlo_octree *t) { /* octree.h:15-31 */
    nlo_octant *n = (nlo_octant *)calloc(1, sizeof(nlo_octant));
    n->index = t->next_index++;
    n->type = NLO_NONLEAF;
    return n;
}

nlo_octree *nlo_octree_create(int grid_dim) { /* octree.cpp:34-49 */
    nlo_octree *t = (nlo_octree *)calloc(1, sizeof(nlo_octree));
    t->size = grid_dim;
    t->max_level = (int)log2((double)grid_dim);
    t->root = nlo_new_octant(t);
    t->root->side = (unsigned)grid_dim;
    return t;
}

static void nlo_free_rec(nlo_octant *n) {
    if (!n) return;
    for (int i = 0; i < 8; ++i) nlo_free_rec(n->child[i]);
    free(n);
}
void nlo_octree_destroy(nlo_octree *t) {
    if (!t) return;
    nlo_free_rec(t->root);
    free(t);
}

static const int NLO_INCR_X[8] = {0, 0, 0, 0, 1, 1, 1, 1}; /* octree.cpp:12 */
static const int NLO_INCR_Y[8] = {0, 0, 1, 1, 0, 0, 1, 1}; /* octree.cpp:13 */
static const int NLO_INCR_Z[8] = {0, 1, 0, 1, 0, 1, 0, 1}; /* octree.cpp:14 */

void nlo_octree_insert(nlo_octree *t, const int32_t *pts, int64_t npts) { /* octree.cpp:51-111 */
    const int MAX_BITS = 21;
    for (int64_t i = 0; i < npts; ++i) {
        for (int j = 0; j < 8; ++j) {
            int x = pts[i * 3 + 0] + NLO_INCR_X[j];
            int y = pts[i * 3 + 1] + NLO_INCR_Y[j];
            int z = pts[i * 3 + 2] + NLO_INCR_Z[j];
            uint64_t key = nlo_encode(x, y, z);
            const unsigned shift = (unsigned)(MAX_BITS - t->max_level - 1);
            nlo_octant *n = t->root;
            unsigned edge = (unsigned)t->size / 2;
            for (int d = 1; d <= t->max_level; edge /= 2, ++d) {
                const int childid = ((x & edge) > 0) + 2 * ((y & edge) > 0) + 4 * ((z & edge) > 0);
                nlo_octant *tmp = n->child[childid];
                if (!tmp) {
                    tmp = nlo_new_octant(t);
                    tmp->code = key & nlo_mask(d + (int)shift);
                    tmp->side = edge;
                    tmp->is_leaf = (d == t->max_level);
                    tmp->type = tmp->is_leaf ? (j == 0 ? NLO_SURFACE : NLO_FEATURE) : NLO_NONLEAF;
                    n->child[childid] = tmp;
                } else if (tmp->type == NLO_FEATURE && j == 0) {
                    tmp->type = NLO_SURFACE;
                }
                n = tmp;
            }
        }
    }
}

static nlo_octant *nlo_find(nlo_octree *t, int x, int y, int z) { /* octree.cpp:151-171 */
    nlo_octant *n = t->root;
    unsigned edge = (unsigned)t->size / 2;
    for (int d = 1; d <= t->max_level; edge /= 2, ++d) {
        const int childid = ((x & edge) > 0) + 2 * ((y & edge) > 0) + 4 * ((z & edge) > 0);
        nlo_octant *tmp = n->child[childid];
        if (!tmp) return NULL;
        n = tmp;
    }
    return n;
}

static int64_t nlo_count_rec(nlo_octant *n) { /* octree.cpp:275-291: leaf counts 1 and stops */
    if (!n) return 0;
    if (n->is_leaf) return 1;
    int64_t s = 1;
    for (int i = 0; i < 8; ++i) s += nlo_count_rec(n->child[i]);
    return s;
}
int64_t nlo_octree_count(nlo_octree *t) { return nlo_count_rec(t->root); }

static int64_t nlo_count_surface_rec(nlo_octant *n) { /* octree.cpp:372-389 */
    if (!n) return 0;
    if (n->type == NLO_SURFACE) return 1;
    int64_t s = 0;
    for (int i = 0; i < 8; ++i) s += nlo_count_surface_rec(n->child[i]);
    return s;
}
int64_t nlo_octree_count_leaf(nlo_octree *t) { return nlo_count_surface_rec(t->root); }

/* octree.cpp:293-342.  voxels f32[n,4] (zero-init), children f32[n,8] (-1 init), features i32[n,8] (-1 init);
 * n = nlo_octree_count().  BFS over children whose type != FEATURE. */
void nlo_octree_export(nlo_octree *t, float *voxels, float *children, int32_t *features) {
    int64_t n = nlo_octree_count(t);
    for (int64_t i = 0; i < n * 4; ++i) voxels[i] = 0.f;
    for (int64_t i = 0; i < n * 8; ++i) { children[i] = -1.f; features[i] = -1; }
    nlo_octant **queue = (nlo_octant **)malloc(sizeof(nlo_octant *) * (size_t)(n + 1));
    int64_t head = 0, tail = 0;
    queue[tail++] = t->root;
    while (head < tail) {
        nlo_octant *node = queue[head++];
        int xyz[3];
        nlo_decode(node->code, xyz);
        float coords[4] = {(float)xyz[0], (float)xyz[1], (float)xyz[2], (float)node->side};
        memcpy(voxels + (int64_t)node->index * 4, coords, sizeof(coords));
        if (node->type == NLO_SURFACE) {
            for (int i = 0; i < 8; ++i) {
                float vx = coords[0] + NLO_INCR_X[i], vy = coords[1] + NLO_INCR_Y[i], vz = coords[2] + NLO_INCR_Z[i];
                nlo_octant *v = nlo_find(t, (int)vx, (int)vy, (int)vz);
                if (v) features[(int64_t)node->index * 8 + i] = v->index;
            }
        }
        for (int i = 0; i < 8; ++i) {
            nlo_octant *c = node->child[i];
            if (c && c->type != NLO_FEATURE) {
                queue[tail++] = c;
                children[(int64_t)node->index * 8 + i] = (float)c->index;
            }
        }
    }
    free(queue);
}

/* ======================================================================================
 * 2. Ray / octree intersection
 *    third_party/sparse_voxels/src/intersect_gpu.cu:77-142 (RayAABBIntersection),
 *    :193-272 (svo_intersect_point_kernel)
 * ====================================================================================== */

/* intersect_gpu.cu:77-142.  __fdividef(1, x) is restated as 1.0f/x (<= 2 ulp apart; the GPU
 * tests compare indices exactly and depths to a relative tolerance). */
static int nlo_ray_aabb(const float ori[3], const float dir[3], const float center[3], float half_voxel,
                        float *out_low, float *out_high) {
    float f_low = 0.f, f_high = 100000.f;
    for (int d = 0; d < 3; ++d) {
        float inv_ray_dir = 1.0f / dir[d];
        float start = ori[d], aabb = center[d];
        float f_dim_low = (aabb - half_voxel - start) * inv_ray_dir;
        float f_dim_high = (aabb + half_voxel - start) * inv_ray_dir;
        if (f_dim_high < f_dim_low) { float tmp = f_dim_low; f_dim_low = f_dim_high; f_dim_high = tmp; }
        if (f_dim_high < f_low) return 0;
        if (f_dim_low > f_high) return 0;
        f_low = (f_dim_low > f_low) ? f_dim_low : f_low;
        f_high = (f_dim_high < f_high) ? f_dim_high : f_high;
        if (f_low > f_high) return 0;
    }
    *out_low = f_low;
    *out_high = f_high;
    return 1; /* the kernel tests depths.x > -1.0f; a hit always has f_low >= 0 */
}

/* intersect_gpu.cu:193-272 for one batch (the G-way replication of the octree in
 * voxel_helpers.py:97-108 does not change per-ray results).  points f32[n,3], children i32[n,9],
 * outputs [m,n_max]; idx pre-set to -1, depths zero (intersect.cpp:98-106).
 * Returns the maximum DFS stack occupancy seen (the reference asserts < 256). */
int nlo_svo_intersect(int64_t m, float voxelsize, int n_max, const float *ray_start, const float *ray_dir,
                      const float *points, const int32_t *children, int32_t *idx, float *min_depth,
                      float *max_depth) {
    float half_voxel = (float)(voxelsize * 0.5); /* intersect_gpu.cu:222 (double mul, exact) */
    int max_ptr = 0;
    for (int64_t j = 0; j < m; ++j) {
        for (int l = 0; l < n_max; ++l) { idx[j * n_max + l] = -1; min_depth[j * n_max + l] = 0.f; max_depth[j * n_max + l] = 0.f; }
        int stack[256];
        int ptr = 0, cnt = 0, k = -1;
        stack[ptr] = 0;
        while (ptr > -1 && cnt < n_max) {
            k = stack[ptr];
            float lo = -1.f, hi = -1.f;
            int hit = nlo_ray_aabb(ray_start + j * 3, ray_dir + j * 3, points + (int64_t)k * 3,
                                   half_voxel * (float)children[(int64_t)k * 9 + 8], &lo, &hi);
            ptr--;
            if (hit && lo > -1.0f) {
                if (children[(int64_t)k * 9 + 8] == 1) {
                    idx[j * n_max + cnt] = k;
                    min_depth[j * n_max + cnt] = lo;
                    max_depth[j * n_max + cnt] = hi;
                    ++cnt;
                    continue;
                }
                for (int u = 0; u < 8; ++u) {
                    if (children[(int64_t)k * 9 + u] > -1) {
                        ptr++;
                        if (ptr > max_ptr) max_ptr = ptr;
                        if (ptr >= 256) return -1;
                        stack[ptr] = children[(int64_t)k * 9 + u];
                    }
                }
            }
        }
    }
    return max_ptr + 1;
}

/* ======================================================================================
 * 3. Inverse-CDF ray sampling
 *    third_party/sparse_voxels/src/sample_gpu.cu:133-239 (inverse_cdf_sampling_kernel), one call
 *    = one launch over [b, num_rays, ...]; outputs pre-initialised as in sample.cpp:82-90
 *    (idx -1, depth 0, dists 0).  Both index quirks are kept verbatim.
 *    nvcc's default -fmad=true contracts `min + u*(max-min)` into one FMA on the device; that
 *    contraction is written explicitly here (fmaf) and nowhere else.
 * ====================================================================================== */
void nlo_inverse_cdf_sampling(int b, int num_rays, int max_hits, int max_steps, float fixed_step_size,
                              const int32_t *pts_idx_, const float *min_depth_, const float *max_depth_,
                              const float *uniform_noise_, const float *probs_, const float *steps_,
                              int32_t *sampled_idx_, float *sampled_depth_, float *sampled_dists_) {
    for (int64_t i = 0; i < (int64_t)b * num_rays * max_steps; ++i) { sampled_idx_[i] = -1; sampled_depth_[i] = 0.f; sampled_dists_[i] = 0.f; }
    for (int batch_index = 0; batch_index < b; ++batch_index) {
        const int32_t *pts_idx = pts_idx_ + (int64_t)batch_index * num_rays * max_hits;
        const float *min_depth = min_depth_ + (int64_t)batch_index * num_rays * max_hits;
        const float *max_depth = max_depth_ + (int64_t)batch_index * num_rays * max_hits;
        const float *probs = probs_ + (int64_t)batch_index * num_rays * max_hits;
        const float *steps = steps_ + (int64_t)batch_index * num_rays;
        const float *uniform_noise = uniform_noise_ + (int64_t)batch_index * num_rays * max_steps;
        int32_t *sampled_idx = sampled_idx_ + (int64_t)batch_index * num_rays * max_steps;
        float *sampled_depth = sampled_depth_ + (int64_t)batch_index * num_rays * max_steps;
        float *sampled_dists = sampled_dists_ + (int64_t)batch_index * num_rays * max_steps;
        for (int j = 0; j < num_rays; ++j) {
            int H = j * max_hits, K = j * max_steps;
            int curr_bin = 0, s = 0;
            float curr_min_depth = min_depth[H];
            float curr_max_depth = max_depth[H];
            float curr_min_cdf = 0;
            float curr_max_cdf = probs[H];
            float step_size = (float)(1.0 / (double)steps[j]);
            float z_low = curr_min_depth;
            int total_steps = (int)ceil((double)steps[j]);
            int done = 0;
            if (fixed_step_size > 0.0) step_size = fixed_step_size;
            for (int curr_step = 0; curr_step < total_steps; curr_step++) {
                float curr_cdf = ((float)curr_step + uniform_noise[K + curr_step]) * step_size;
                while (curr_cdf > curr_max_cdf) {
                    sampled_idx[K + s] = pts_idx[H + curr_bin];
                    sampled_dists[K + s] = (curr_max_depth - z_low);
                    sampled_depth[K + s] = (float)((double)(curr_max_depth + z_low) * .5);
                    curr_bin++;
                    s++;
                    if ((curr_bin >= max_hits) || (pts_idx[H + curr_bin] == -1)) { done = 1; break; }
                    curr_min_depth = min_depth[H + curr_bin];
                    curr_max_depth = max_depth[H + curr_bin];
                    curr_min_cdf = curr_max_cdf;
                    curr_max_cdf = curr_max_cdf + probs[H + curr_bin];
                    z_low = curr_min_depth;
                }
                if (done) break;
                float u = (curr_cdf - curr_min_cdf) / (curr_max_cdf - curr_min_cdf);
                float z = fmaf(u, (curr_max_depth - curr_min_depth), curr_min_depth);
                sampled_idx[K + s] = pts_idx[H + curr_bin];
                sampled_dists[K + s] = (z - z_low);
                sampled_depth[K + s] = (float)((double)(z + z_low) * .5);
                z_low = z;
                s++;
            }
            /* sample_gpu.cu:224-238: quirk 1 (num_rays > H + curr_bin) and quirk 2 (pts_idx[curr_bin]) */
            while ((z_low < curr_max_depth) && (!done) && (num_rays > (H + curr_bin))) {
                sampled_idx[K + s] = pts_idx[H + curr_bin];
                sampled_dists[K + s] = (curr_max_depth - z_low);
                sampled_depth[K + s] = (float)((double)(curr_max_depth + z_low) * .5);
                curr_bin++;
                s++;
                if ((curr_bin >= max_hits) || (pts_idx[curr_bin] == -1)) break;
                curr_min_depth = min_depth[H + curr_bin];
                curr_max_depth = max_depth[H + curr_bin];
                z_low = curr_min_depth;
            }
        }
    }
}
