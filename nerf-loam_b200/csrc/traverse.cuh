// Ray / sparse-voxel-octree traversal shared by the `grid.svo_intersect` drop-in (grid_dropin.cu)
// and the fused renderer (render.cu).
//
// Semantics follow third_party/sparse_voxels/src/intersect_gpu.cu:77-142 (slab test) and :225-271
// (DFS): start from node 0, pop a node, slab-test it with half extent 0.5*voxelsize*side, record it
// if it is a leaf (side == 1), otherwise push its existing children in slot order 0..7 (so they pop
// 7..0); stop after n_max leaves.  The reciprocal is __fdividef(1, d) like the reference, the slab
// arithmetic is sub, sub, mul (nothing for the compiler to contract), so depths are bit-identical to
// the reference kernel compiled for the same GPU.
//
// Differences that do not change results: one copy of the octree (the reference's wrapper physically
// replicates it up to 256x, voxel_helpers.py:97-108); the reciprocals are computed once per ray; the
// children of a node are slab-tested when the node is expanded and only the hit ones are pushed, so
// misses never reach the stack -- the emission order is unchanged because a missed child emits nothing
// and pushes nothing; a child's half extent is 0.5 * its parent's (sides are powers of two, so
// half_voxel*side is reproduced exactly).
#pragma once
#include "nl_cuda.cuh"

#define NL_STACK_CAP 128  // worst case 1 + 7 * 18 levels (every child of every expanded node hit); typical <= 3 per level

struct NlRay {
    float ox, oy, oz;
    float ix, iy, iz;  // __fdividef(1, d)
};

__device__ __forceinline__ NlRay nl_make_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
    NlRay r;
    r.ox = ox; r.oy = oy; r.oz = oz;
    r.ix = __fdividef(1.0f, dx);
    r.iy = __fdividef(1.0f, dy);
    r.iz = __fdividef(1.0f, dz);
    return r;
}

// intersect_gpu.cu:77-142.  Returns true on hit and the clipped interval [lo, hi], lo >= 0.
__device__ __forceinline__ bool nl_slab(const NlRay &r, float cx, float cy, float cz, float half, float &lo, float &hi) {
    float f_low = 0.f, f_high = 100000.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float inv = d == 0 ? r.ix : (d == 1 ? r.iy : r.iz);
        const float start = d == 0 ? r.ox : (d == 1 ? r.oy : r.oz);
        const float aabb = d == 0 ? cx : (d == 1 ? cy : cz);
        float a = __fmul_rn(__fsub_rn(__fsub_rn(aabb, half), start), inv);
        float b = __fmul_rn(__fsub_rn(__fadd_rn(aabb, half), start), inv);
        if (b < a) { float t = a; a = b; b = t; }
        if (b < f_low) return false;
        if (a > f_high) return false;
        f_low = (a > f_low) ? a : f_low;
        f_high = (b < f_high) ? b : f_high;
        if (f_low > f_high) return false;
    }
    lo = f_low;
    hi = f_high;
    return true;
}

// Traverse; calls emit(node, lo, hi) for each leaf hit in reference order; returns the number of
// leaves emitted (<= n_max) or -1 on stack overflow.
template <typename Emit>
__device__ __forceinline__ int nl_traverse(const NlRay &ray, const float *__restrict__ centres,
                                           const int32_t *__restrict__ structure, float half_voxel, int n_max, Emit emit) {
    int stack[NL_STACK_CAP];
    float s_lo[NL_STACK_CAP], s_hi[NL_STACK_CAP];
    int ptr = -1, cnt = 0;
    {   // root (node 0) is tested like any popped node
        float lo, hi;
        const float half = __fmul_rn(half_voxel, (float)structure[8]);
        if (nl_slab(ray, centres[0], centres[1], centres[2], half, lo, hi)) { ptr = 0; stack[0] = 0; s_lo[0] = lo; s_hi[0] = hi; }
    }
    while (ptr > -1 && cnt < n_max) {
        const int k = stack[ptr];
        const float lo = s_lo[ptr], hi = s_hi[ptr];
        --ptr;
        const int32_t *st = structure + (size_t)k * 9;
        const int side = st[8];
        if (side == 1) {  // terminal node
            emit(k, lo, hi);
            ++cnt;
            continue;
        }
        const float half = __fmul_rn(__fmul_rn(half_voxel, (float)side), 0.5f);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = st[u];
            if (c > -1) {
                float clo, chi;
                if (nl_slab(ray, centres[(size_t)c * 3], centres[(size_t)c * 3 + 1], centres[(size_t)c * 3 + 2], half, clo, chi)) {
                    if (ptr + 1 >= NL_STACK_CAP) return -1;
                    ++ptr;
                    stack[ptr] = c; s_lo[ptr] = clo; s_hi[ptr] = chi;
                }
            }
        }
    }
    return cnt;
}
