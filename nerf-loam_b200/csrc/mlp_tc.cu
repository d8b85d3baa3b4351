// Tensor-core version of the SDF decoder MLP (16 -> 256 -> 256 -> 1, ReLU) for sm_100a:
// tcgen05.mma kind::tf32 with a 3-term hi/lo split (3xTF32: a*b ~= ah*bh + ah*bl + al*bh, fp32
// accumulation in TMEM), which keeps the result within ~1e-6 of an fp32 evaluation -- single-pass
// TF32/BF16 (1e-3) would break the 1e-5 SDF parity bar.
//
// One persistent CTA per SM, 6 warps:
//   warp 0   producer : streams pre-swizzled weight panels (hi+lo, 64 KB per K-block of 32) from L2 into a
//                       2-stage shared-memory ring with cp.async.bulk (1-D TMA) + mbarrier complete_tx
//   warp 1   MMA      : one elected thread issues tcgen05.mma (M=128 samples, N=256, K=8 per instruction),
//                       releases ring stages with tcgen05.commit; owns the TMEM allocation (512 columns:
//                       D1 = layer-1 accumulator, D2 = layer-2 accumulator)
//   warps 2-5 epilogue: thread = sample row (TMEM lane).  Loads x, and for every K-block of layer 2 pulls 32
//                       columns of D1 out of TMEM (tcgen05.ld), applies bias+ReLU, splits into tf32 hi/lo and
//                       writes the A operand tile straight into the 128B-swizzled K-major layout the tensor
//                       core reads; finally reduces relu(D2 + b1) . w2 + b2 per row.
// Layer 1 (K = 16) runs through the same ring as one extra K-block with 2 k-steps.
// Shared memory: 2 x 64 KB weight stages + 2 x 32 KB activation stages = 192 KB.
#include "nl_cuda.cuh"

namespace tc {

constexpr int TM = 128;               // samples per tile (UMMA M)
constexpr int WN = 256;               // hidden width (UMMA N)
constexpr int PANEL_A = TM * 128;     // 16 KB : 128 rows x 128 B (32 tf32 of K)
constexpr int PANEL_B = WN * 128;     // 32 KB : 256 rows x 128 B
constexpr int STAGE_A = 2 * PANEL_A;  // hi + lo
constexpr int STAGE_B = 2 * PANEL_B;
constexpr int NSTAGE = 2;
constexpr int STEPS_FWD = 9;          // layer 1 + 8 K-blocks of layer 2
constexpr int SMEM_DATA = NSTAGE * (STAGE_A + STAGE_B);
constexpr int SMEM_TOTAL = SMEM_DATA + 4096 + 1024;  // + biases/barriers + alignment slack
constexpr int NTHREADS = 192;

// byte offset of 16-byte chunk c (0..7) of row r inside a K-major SWIZZLE_128B panel
__host__ __device__ __forceinline__ int panel_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// ---- mbarrier helpers -------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30)
// (ignored for swizzled K-major, set to 1), SBO>>4 = 1024>>4 [32,46), version 1 [46,48), layout SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 [4,6), a/b_format TF32 = 2 [7,10)/[10,13),
// a/b K-major = 0, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// Weight panels: the exact shared-memory image (K-major, 128B swizzle, tf32 hi then lo) of every weight
// K-block, laid out contiguously in global memory so one bulk copy fills a ring stage.
//   stage 0      : W0  [256 x 16]  (row j, k = e; only the first 64 B of each 128 B row are meaningful)
//   stage 1 + kb : W1[:, 32kb : 32kb+32]
// ------------------------------------------------------------------------------------------------
__global__ void k_tc_prepare(const float *__restrict__ W0, const float *__restrict__ W1, uint8_t *__restrict__ panels) {
    const int stage = blockIdx.y;                         // 0..8
    const int t = blockIdx.x * blockDim.x + threadIdx.x;  // (row, chunk)
    if (t >= WN * 8) return;
    const int r = t >> 3, c = t & 7;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (stage == 0) {
        if (c < 4) for (int i = 0; i < 4; ++i) v[i] = W0[r * 16 + c * 4 + i];
    } else {
        const int k0 = (stage - 1) * 32 + c * 4;
        for (int i = 0; i < 4; ++i) v[i] = W1[(size_t)r * WN + k0 + i];
    }
    float hi[4], lo[4];
    for (int i = 0; i < 4; ++i) { hi[i] = tf32_rna(v[i]); lo[i] = tf32_rna(v[i] - hi[i]); }
    uint8_t *base = panels + (size_t)stage * STAGE_B;
    *reinterpret_cast<float4 *>(base + panel_off(r, c)) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4 *>(base + PANEL_B + panel_off(r, c)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

struct FwdParams {
    long long M_host;
    const int32_t *M_dev;
    const float *feats;       // [M,16]
    const uint8_t *panels;    // 9 stages x 64 KB
    const float *b0, *b1, *w2, *b2;
    float *sdf;
};

__global__ void __launch_bounds__(NTHREADS, 1) k_mlp_tc_fwd(FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B panels need 1024-byte alignment
    uint8_t *sm = smem_raw + (base - raw);
    const uint32_t sB = base, sA = base + NSTAGE * STAGE_B;
    uint8_t *A_gen = sm + NSTAGE * STAGE_B;        // generic pointer to the activation stages
    float *b0s = reinterpret_cast<float *>(sm + SMEM_DATA);
    float *b1s = b0s + WN;
    float *w2s = b1s + WN;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + SMEM_DATA + 3 * WN * 4);
    // barriers: 0,1 b_full  2,3 b_empty  4,5 a_full  6,7 a_empty  8 d1_full  9 d2_full  10 d2_empty
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long M = p.M_dev ? min((long long)*p.M_dev, p.M_host) : p.M_host;
    const long long ntiles = (M + TM - 1) / TM;

    for (int i = tid; i < WN; i += NTHREADS) { b0s[i] = p.b0[i]; b1s[i] = p.b1[i]; w2s[i] = p.w2[i]; }
    if (tid == 0) {
        mbar_init(BAR(0), 1); mbar_init(BAR(1), 1);      // b_full: producer's expect_tx arrive
        mbar_init(BAR(2), 1); mbar_init(BAR(3), 1);      // b_empty: tcgen05.commit
        mbar_init(BAR(4), 4); mbar_init(BAR(5), 4);      // a_full: one arrive per epilogue warp
        mbar_init(BAR(6), 1); mbar_init(BAR(7), 1);      // a_empty: tcgen05.commit
        mbar_init(BAR(8), 1); mbar_init(BAR(9), 1);      // d1_full, d2_full: tcgen05.commit
        mbar_init(BAR(10), 4);                           // d2_empty: one arrive per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (D1 at column 0, D2 at column 256)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t D1 = tmem, D2 = tmem + 256;

    if (warp == 0) {
        // ===================== producer: weight panels -> ring =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int step = 0; step < STEPS_FWD; ++step, ++it) {
                    const uint32_t s = it & 1, ph = (it >> 1) & 1;
                    mbar_wait(BAR(2 + s), ph ^ 1);
                    mbar_expect_tx(BAR(0 + s), STAGE_B);
                    const uint8_t *src = p.panels + (size_t)step * STAGE_B;
                    bulk_g2s(sB + s * STAGE_B, src, PANEL_B, BAR(0 + s));
                    bulk_g2s(sB + s * STAGE_B + PANEL_B, src + PANEL_B, PANEL_B, BAR(0 + s));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TM, WN);
            uint32_t it = 0, tl = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
                for (int step = 0; step < STEPS_FWD; ++step, ++it) {
                    const uint32_t s = it & 1, ph = (it >> 1) & 1;
                    if (step == 1) mbar_wait(BAR(10), (tl & 1) ^ 1);   // D2 of the previous tile fully read
                    mbar_wait(BAR(0 + s), ph);                          // weights landed
                    mbar_wait(BAR(4 + s), ph);                          // activations written
                    tc_fence_after();
                    const uint32_t a_hi = sA + s * STAGE_A, a_lo = a_hi + PANEL_A;
                    const uint32_t b_hi = sB + s * STAGE_B, b_lo = b_hi + PANEL_B;
                    const uint32_t d = (step == 0) ? D1 : D2;
                    const int nk = (step == 0) ? 2 : 4;                 // k-steps of 8 (32 B) in this K-block
                    uint32_t acc = (step <= 1) ? 0u : 1u;               // first MMA of a layer overwrites
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const uint64_t ad = make_desc(term == 2 ? a_lo : a_hi);
                        const uint64_t bd = make_desc(term == 1 ? b_lo : b_hi);
                        for (int ks = 0; ks < nk; ++ks) {
                            mma_tf32(d, ad + 2 * ks, bd + 2 * ks, idesc, acc);
                            acc = 1u;
                        }
                    }
                    tc_commit(BAR(2 + s));                              // ring stages free once these MMAs retire
                    tc_commit(BAR(6 + s));
                    if (step == 0) tc_commit(BAR(8));                   // D1 complete
                    if (step == STEPS_FWD - 1) tc_commit(BAR(9));       // D2 complete
                }
            }
        }
    } else {
        // ===================== epilogue warps: thread = sample row =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const float b2 = p.b2[0];
        uint32_t it = 0, tl = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
            const long long m = tile * TM + row;
            // ---- step 0: x (16 features = 4 chunks) -> A stage ----
            {
                const uint32_t s = it & 1, ph = (it >> 1) & 1;
                mbar_wait(BAR(6 + s), ph ^ 1);
                uint8_t *dst = A_gen + s * STAGE_A;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (m < M) v = *reinterpret_cast<const float4 *>(p.feats + (size_t)m * 16 + c * 4);
                    const float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
                    const float4 lo = make_float4(tf32_rna(v.x - hi.x), tf32_rna(v.y - hi.y), tf32_rna(v.z - hi.z), tf32_rna(v.w - hi.w));
                    *reinterpret_cast<float4 *>(dst + panel_off(row, c)) = hi;
                    *reinterpret_cast<float4 *>(dst + PANEL_A + panel_off(row, c)) = lo;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(4 + s));
                ++it;
            }
            // ---- steps 1..8: h1 K-block kb = relu(D1[:, 32kb:32kb+32] + b0) -> A stage ----
            mbar_wait(BAR(8), tl & 1);
            tc_fence_after();
            for (int kb = 0; kb < 8; ++kb, ++it) {
                const uint32_t s = it & 1, ph = (it >> 1) & 1;
                uint32_t v[32];
                tmem_ld32(D1 + lane_addr + kb * 32, v);
                mbar_wait(BAR(6 + s), ph ^ 1);
                uint8_t *dst = A_gen + s * STAGE_A;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float h[4], hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        h[i] = fmaxf(__uint_as_float(v[c * 4 + i]) + b0s[kb * 32 + c * 4 + i], 0.f);
                        hi[i] = tf32_rna(h[i]);
                        lo[i] = tf32_rna(h[i] - hi[i]);
                    }
                    *reinterpret_cast<float4 *>(dst + panel_off(row, c)) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4 *>(dst + PANEL_A + panel_off(row, c)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(4 + s));
            }
            // ---- output layer: sdf = relu(D2 + b1) . w2 + b2 ----
            mbar_wait(BAR(9), tl & 1);
            tc_fence_after();
            float acc4[4] = {0.f, 0.f, 0.f, 0.f};
            for (int cb = 0; cb < 8; ++cb) {
                uint32_t v[32];
                tmem_ld32(D2 + lane_addr + cb * 32, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float h = fmaxf(__uint_as_float(v[i]) + b1s[cb * 32 + i], 0.f);
                    acc4[i & 3] = fmaf(h, w2s[cb * 32 + i], acc4[i & 3]);
                }
            }
            const float acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(10));
            if (m < M) p.sdf[m] = acc + b2;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace tc

extern "C" int64_t nl_mlp_tc_panel_bytes(void) { return (int64_t)tc::STEPS_FWD * tc::STAGE_B; }

extern "C" int nl_mlp_tc_prepare(const float *W0, const float *W1, void *panels, void *stream) {
    if (!W0 || !W1 || !panels) return nl_set_error("nl_mlp_tc_prepare: null pointer");
    dim3 grid(nl_div_up(tc::WN * 8, 256), tc::STEPS_FWD);
    tc::k_tc_prepare<<<grid, 256, 0, (cudaStream_t)stream>>>(W0, W1, (uint8_t *)panels);
    NL_CHECK_LAUNCH("nl_mlp_tc_prepare");
    return NL_OK;
}

extern "C" int nl_mlp_tc_forward(int64_t M, const int32_t *d_M_dev, const float *feats, const void *panels, const float *b0,
                                 const float *b1, const float *w2, const float *b2, float *sdf, void *stream) {
    if (M < 0) return nl_set_error("nl_mlp_tc_forward: negative M");
    if (M == 0) return NL_OK;
    if (!feats || !panels || !b0 || !b1 || !w2 || !b2 || !sdf) return nl_set_error("nl_mlp_tc_forward: null pointer");
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc::k_mlp_tc_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_TOTAL);
        if (e != cudaSuccess) return nl_set_error_code(NL_ERR_CUDA, cudaGetErrorString(e));
        configured = true;
    }
    tc::FwdParams p;
    p.M_host = M; p.M_dev = d_M_dev; p.feats = feats; p.panels = (const uint8_t *)panels;
    p.b0 = b0; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.sdf = sdf;
    const long long ntiles = (M + tc::TM - 1) / tc::TM;
    const int grid = (int)(ntiles < (long long)nl_num_sms() ? ntiles : (long long)nl_num_sms());
    tc::k_mlp_tc_fwd<<<grid, tc::NTHREADS, tc::SMEM_TOTAL, (cudaStream_t)stream>>>(p);
    NL_CHECK_LAUNCH("nl_mlp_tc_forward");
    return NL_OK;
}
