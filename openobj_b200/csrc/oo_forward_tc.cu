// Blackwell-native forward of one hidden-32 OccupancyMap + UniDirsEmbed at free query points -- the compute of
// Trainer.eval_points / the meshing grid (objnerf/trainer.py:46-69,104-128; embedding.py:46-55; model.py:61-103;
// render_rays.py:6-14) -- on the 5th-generation tensor cores: tcgen05.mma kind::tf32, accumulators AND the A operands in
// tensor memory, the object's weights pre-split and resident in shared memory.
//
//   * one CTA per SM, 128 threads, a tile = 128 query points = the 128 lanes of tensor memory: thread p owns point p;
//   * the layers are chained through TMEM: an MMA leaves D[point][unit] (fp32) in TMEM, the owning thread reads its row with
//     tcgen05.ld, applies bias + ReLU, splits the result into (hi, lo) TF32 halves and writes them back with tcgen05.st as
//     the A operand [point][k] of the next layer -- activations never touch shared memory;
//   * fp32-level accuracy with TF32 inputs by three-term error compensation, exactly like the mma.sync path of the training
//     tile (oo_tile.h): x = hi + lo, a.b ~ lo_a hi_b + hi_a lo_b + hi_a hi_b accumulated in fp32 -> 3 tcgen05.mma per
//     8-wide k-step.  The weights are constant per object, so their (hi, lo) copies are formed ONCE when the CTA starts and
//     stay in shared memory in the canonical K-major core-matrix layout the MMA's shared-memory descriptor addresses
//     ([k / 4][32 rows][4]: 8 rows x 16 bytes per core matrix, SBO = 128 B between 8-row groups, LBO = 512 B between 16-byte
//     k-chunks); nothing is re-split per tile;
//   * the encoder and the two tiny output layers (out_alpha 32 -> 1, out_color 32 -> 3) run in registers of the owning
//     thread in the operation order of the training tile (phase 1 / phase 7 of oo_tile.h).
//
// Per 128 points: 132 tcgen05.mma (M128 N32 K8) = 4.3 M tensor MACs for 1.43 M algorithmic MACs.
#include "../../include/openobj_b200.h"
#include "oo_common.cuh"
#include "oo_tile.h"

using namespace oo;

namespace {

constexpr int TC_M = 128, TC_THREADS = 128;
// tensor-memory columns (32-bit each, 128 lanes)
constexpr int C_E1H = 0, C_E1L = 88, C_E2H = 176, C_E2L = 224, C_HAH = 272, C_HAL = 304, C_HBH = 336, C_HBL = 368, C_D = 400;
constexpr int TC_COLS = 512;
// shared memory (floats): weight blocks in canonical layout (hi copies, then lo copies), small vectors, the mbarrier
constexpr int K_IN = 88, K_H = 32, K_CATB = 88, K_CLB = 48;
constexpr int WB_IN = 0, WB_M1 = WB_IN + H * K_IN, WB_CATA = WB_M1 + H * K_H, WB_CATB = WB_CATA + H * K_H,
              WB_M2 = WB_CATB + H * K_CATB, WB_CLA = WB_M2 + H * K_H, WB_CLB = WB_CLA + H * K_H, WB_TOTAL = WB_CLB + H * K_CLB;
constexpr int SB_BIAS = 2 * WB_TOTAL;                 // in, m1, cat, m2, cl: 5 x 32
constexpr int SB_WA = SB_BIAS + 5 * H;                // out_alpha.weight [32], bias at +32
constexpr int SB_WOC = SB_WA + 36;                    // out_color.weight [3][32], bias at +96
constexpr int SB_PE = SB_WOC + 100;                   // B_layer.weight [21][3]
constexpr int SB_BAR = SB_PE + 64;                    // mbarrier (8 bytes, 8-byte aligned: SB_BAR is even)
constexpr int SB_FLOATS = SB_BAR + 4;
// at least half of the SM's shared memory is requested so that two CTAs can never share an SM: each allocates all 512
// tensor-memory columns, and a second CTA would wait for them forever
constexpr size_t TC_SMEM = 120 * 1024;
static_assert(SB_FLOATS * 4 <= TC_SMEM && SB_BAR % 2 == 0, "shared-memory map");

// instruction descriptor of tcgen05.mma kind::tf32: D fp32 (bits 4-5 = 1), A and B TF32 (bits 7-9 / 10-12 = 2), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ uint64_t smem_desc(uint32_t byte_addr) {
    // K-major, no swizzle: start >> 4 | LBO (512 B, next 16-byte k-chunk) >> 4 at bit 16 | SBO (128 B, next 8 rows) >> 4 at
    // bit 32 | descriptor version 1 (sm_100) at bit 46
    return (uint64_t)((byte_addr & 0x3FFFFu) >> 4) | ((uint64_t)(512u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(IDESC), "r"(accumulate) : "memory");
}

// D (+)= A [128 x K] * W^T with A's (hi, lo) halves in TMEM columns a_hi / a_lo and W's in the shared blocks w_hi / w_lo
__device__ __forceinline__ void layer_mma(uint32_t tm, int a_hi, int a_lo, uint32_t w_hi, uint32_t w_lo, int K, bool first) {
    for (int s = 0; s < K / 8; ++s) {
        const uint64_t bh = smem_desc(w_hi + (uint32_t)(2 * s) * 512u), bl = smem_desc(w_lo + (uint32_t)(2 * s) * 512u);
        umma_ts(tm + C_D, tm + a_lo + 8 * s, bh, (first && s == 0) ? 0u : 1u);      // small terms first
        umma_ts(tm + C_D, tm + a_hi + 8 * s, bl, 1u);
        umma_ts(tm + C_D, tm + a_hi + 8 * s, bh, 1u);
    }
}

// tcgen05.st / tcgen05.ld of 8 / 32 consecutive columns of this thread's lane WITHOUT the completion wait: a layer's stores
// are awaited once (publish_tmem), its 32 accumulator columns arrive with one load and one wait
__device__ __forceinline__ void tm_st8_nw(uint32_t addr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tm_ld32(uint32_t addr, float* v) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
                 "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void split(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// one weight matrix [32][K_real] (row-major in theta, `ld` floats per row, columns col0 ..) -> canonical (hi, lo) blocks
__device__ void stage_block(int tid, float* sm, int blk, const float* __restrict__ W, int ld, int col0, int k_real, int K) {
    for (int i = tid; i < H * K; i += TC_THREADS) {
        const int j = i / K, k = i - j * K;
        const float v = k < k_real ? W[j * ld + col0 + k] : 0.f;
        float hi, lo;
        split(v, hi, lo);
        const int e = (k >> 2) * (H * 4) + j * 4 + (k & 3);
        sm[blk + e] = hi;
        sm[WB_TOTAL + blk + e] = lo;
    }
}

// bias + ReLU on the 32 accumulator columns of this thread's point, result into v; then (hi, lo) -> TMEM columns
__device__ __forceinline__ void epilogue(uint32_t tm_lane, const float* __restrict__ bias, float* v) {
    tm_ld32(tm_lane + C_D, v);
#pragma unroll
    for (int j = 0; j < H; ++j) v[j] = fmaxf(v[j] + bias[j], 0.f);
}
__device__ __forceinline__ void put_split(uint32_t tm_lane, int col_hi, int col_lo, const float* v) {
#pragma unroll
    for (int q = 0; q < H; q += 8) {
        float hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split(v[q + i], hi[i], lo[i]);
        tm_st8_nw(tm_lane + col_hi + q, hi);
        tm_st8_nw(tm_lane + col_lo + q, lo);
    }
}

// wait for phase `parity` of the mbarrier; bounded (a descriptor mistake must end as an error code, not as a hung GPU)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
        if (clock64() - t0 > 400000000LL) return false;       // ~0.2 s
    }
}

// all threads: everything written to TMEM so far is visible to the MMA the elected thread issues after the barrier
__device__ __forceinline__ void publish_tmem() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_forward_tc(const float* __restrict__ theta, const float* __restrict__ pts,
                                                              long long n_pts, float scale, float* __restrict__ occ,
                                                              float* __restrict__ alpha_out, float* __restrict__ color,
                                                              int* __restrict__ err) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint32_t tm_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    // ---- once per CTA: the object's weights as canonical (hi, lo) blocks, small vectors, TMEM, the mbarrier
    stage_block(tid, sm, WB_IN, theta + OFF_IN_W, E1, 0, E1, K_IN);
    stage_block(tid, sm, WB_M1, theta + OFF_M1_W, H, 0, H, K_H);
    stage_block(tid, sm, WB_CATA, theta + OFF_CAT_W, H + E1, 0, H, K_H);
    stage_block(tid, sm, WB_CATB, theta + OFF_CAT_W, H + E1, H, E1, K_CATB);
    stage_block(tid, sm, WB_M2, theta + OFF_M2_W, H, 0, H, K_H);
    stage_block(tid, sm, WB_CLA, theta + OFF_CL_W, H + E2, 0, H, K_H);
    stage_block(tid, sm, WB_CLB, theta + OFF_CL_W, H + E2, H, E2, K_CLB);
    if (tid < H) {
        sm[SB_BIAS + tid] = theta[OFF_IN_B + tid];
        sm[SB_BIAS + H + tid] = theta[OFF_M1_B + tid];
        sm[SB_BIAS + 2 * H + tid] = theta[OFF_CAT_B + tid];
        sm[SB_BIAS + 3 * H + tid] = theta[OFF_M2_B + tid];
        sm[SB_BIAS + 4 * H + tid] = theta[OFF_CL_B + tid];
        sm[SB_WA + tid] = theta[OFF_A_W + tid];
    }
    if (tid < 3 * H) sm[SB_WOC + tid] = theta[OFF_OC_W + tid];
    if (tid < 3) sm[SB_WOC + 3 * H + tid] = theta[OFF_OC_B + tid];
    if (tid == 0) sm[SB_WA + H] = theta[OFF_A_B];
    if (tid < NDIR * 3) sm[SB_PE + tid] = theta[OFF_PE_B + tid];
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(sm + SB_BAR);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tm_base_s)), "r"((uint32_t)TC_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the weight blocks are read by the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tm_base_s;
    const uint32_t tm_lane = tm + ((uint32_t)(32 * (warp & 3)) << 16);
    const uint32_t w_hi = (uint32_t)__cvta_generic_to_shared(sm), w_lo = w_hi + WB_TOTAL * 4u;
    uint32_t phase = 0;
    bool ok = true;
    const float* bias = sm + SB_BIAS;

    const long long n_tiles = (n_pts + TC_M - 1) / TC_M;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p = tile * TC_M + tid;
        const bool live = p < n_pts;
        // ---- encoder (embedding.py:46-55; arithmetic of oo_tile.h phase 0 / 1): e = [t, sin(pi 2^k B t)] -> (hi, lo) A operands
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        if (live) {
            t0 = pts[3 * p] / scale; t1 = pts[3 * p + 1] / scale; t2 = pts[3 * p + 2] / scale;
        }
        {
            float sn[NDIR], cs[NDIR];
#pragma unroll
            for (int d = 0; d < NDIR; ++d) {
                const float proj = sm[SB_PE + 3 * d] * t0 + sm[SB_PE + 3 * d + 1] * t1 + sm[SB_PE + 3 * d + 2] * t2;
                sincosf(proj * PI_F, &sn[d], &cs[d]);
            }
            float hi[8], lo[8];
            // columns 0 .. 135: the 88 columns of e1 (87 values + one zero), then the 48 columns of e2 (42 values + six zeros);
            // value q of e1 is t (q < 3) or band (q - 3) / 21 of direction (q - 3) % 21; e2 continues with bands 4 and 5
#pragma unroll
            for (int q = 0; q < 136; ++q) {
                const int row = q < 88 ? q : q - 1;            // embedding row of this column (column 87 is e1's zero pad)
                float v = 0.f;
                if (q < 3) v = q == 0 ? t0 : q == 1 ? t1 : t2;
                else if (q != 87 && row < EMB) {
                    const int d = (row - 3) % NDIR;
                    v = sn[d];
                    const float s2 = 2.f * sn[d] * cs[d], c2 = (cs[d] - sn[d]) * (cs[d] + sn[d]);   // next band of this direction
                    sn[d] = s2; cs[d] = c2;
                }
                split(v, hi[q & 7], lo[q & 7]);
                if ((q & 7) == 7) {
                    const int c0 = q - 7;
                    if (c0 < 88) {
                        tm_st8_nw(tm_lane + C_E1H + c0, hi);
                        tm_st8_nw(tm_lane + C_E1L + c0, lo);
                    } else {
                        tm_st8_nw(tm_lane + C_E2H + c0 - 88, hi);
                        tm_st8_nw(tm_lane + C_E2L + c0 - 88, lo);
                    }
                }
            }
        }
        float v[H];
        // ---- in_layer: relu(W_in e1 + b) -> HA
        publish_tmem();
        if (tid == 0) {
            layer_mma(tm, C_E1H, C_E1L, w_hi + WB_IN * 4u, w_lo + WB_IN * 4u, K_IN, true);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        ok &= mbar_wait(bar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        epilogue(tm_lane, bias, v);
        put_split(tm_lane, C_HAH, C_HAL, v);
        // ---- mid1 -> HB (fc2)
        publish_tmem();
        if (tid == 0) {
            layer_mma(tm, C_HAH, C_HAL, w_hi + WB_M1 * 4u, w_lo + WB_M1 * 4u, K_H, true);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        ok &= mbar_wait(bar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        epilogue(tm_lane, bias + H, v);
        put_split(tm_lane, C_HBH, C_HBL, v);
        // ---- cat_layer on [fc2, e1] -> HA (fc3)
        publish_tmem();
        if (tid == 0) {
            layer_mma(tm, C_HBH, C_HBL, w_hi + WB_CATA * 4u, w_lo + WB_CATA * 4u, K_H, true);
            layer_mma(tm, C_E1H, C_E1L, w_hi + WB_CATB * 4u, w_lo + WB_CATB * 4u, K_CATB, false);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        ok &= mbar_wait(bar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        epilogue(tm_lane, bias + 2 * H, v);
        put_split(tm_lane, C_HAH, C_HAL, v);
        // ---- mid2 -> HB (fc4); out_alpha (x10, model.py:88) in registers
        publish_tmem();
        if (tid == 0) {
            layer_mma(tm, C_HAH, C_HAL, w_hi + WB_M2 * 4u, w_lo + WB_M2 * 4u, K_H, true);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        ok &= mbar_wait(bar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        epilogue(tm_lane, bias + 3 * H, v);
        put_split(tm_lane, C_HBH, C_HBL, v);
        {
            float r = sm[SB_WA + H];
#pragma unroll
            for (int j = 0; j < H; ++j) r += sm[SB_WA + j] * v[j];
            const float a = r * 10.f;
            if (live) {
                if (alpha_out != nullptr) alpha_out[p] = a;
                if (occ != nullptr) occ[p] = 1.f / (1.f + expf(-a));            // render_rays.py:13 without distances
            }
        }
        // ---- color_linear on [fc4, e2], out_color + sigmoid (model.py:94-96) in registers
        publish_tmem();
        if (tid == 0) {
            layer_mma(tm, C_HBH, C_HBL, w_hi + WB_CLA * 4u, w_lo + WB_CLA * 4u, K_H, true);
            layer_mma(tm, C_E2H, C_E2L, w_hi + WB_CLB * 4u, w_lo + WB_CLB * 4u, K_CLB, false);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        ok &= mbar_wait(bar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        epilogue(tm_lane, bias + 4 * H, v);
        if (live) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float r = sm[SB_WOC + 3 * H + ch];
#pragma unroll
                for (int j = 0; j < H; ++j) r += sm[SB_WOC + ch * H + j] * v[j];
                color[3 * p + ch] = sigmoidf_(r);
            }
        }
        // the accumulator columns are re-used by the next tile's first MMA: every thread is past its loads here.  A missed
        // completion (never observed) ends the CTA's work for all of its threads together.
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        ok = __syncthreads_and(ok ? 1 : 0) != 0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!ok) break;
    }
    if (!ok && tid == 0 && err != nullptr) atomicExch(err, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"((uint32_t)TC_COLS) : "memory");
}

}  // namespace

// Blackwell-native (tcgen05 / TMEM) variant of oo_eval_points without the part-feature output.  err_flag: device int[1],
// set to 1 if a tensor-core completion never arrived (never observed; the wait is bounded so that it cannot hang the GPU).
extern "C" int oo_eval_points_tc(const float* theta, const float* pts, long long n_pts, float pe_scale, float* occ, float* alpha,
                                 float* color, int* err_flag, void* stream) {
    OO_REQUIRE(theta && pts && color && (occ || alpha), "oo_eval_points_tc: null argument");
    OO_REQUIRE(n_pts > 0, "oo_eval_points_tc: empty query");
    static PerDevice attr_set;
    if (!attr_set.cur()) {
        OO_CUDA(cudaFuncSetAttribute(k_forward_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        attr_set.cur() = 1;
    }
    int dev = 0, n_sm = 148;
    OO_CUDA(cudaGetDevice(&dev));
    OO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const long long n_tiles = (n_pts + TC_M - 1) / TC_M;
    const int grid = (int)(n_tiles < n_sm ? n_tiles : n_sm);
    k_forward_tc<<<grid, TC_THREADS, TC_SMEM, (cudaStream_t)stream>>>(theta, pts, n_pts, pe_scale, occ, alpha, color, err_flag);
    OO_LAUNCH_CHECK();
    return 0;
}
