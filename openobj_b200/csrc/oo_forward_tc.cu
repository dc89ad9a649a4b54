// Blackwell-native forward of one hidden-32 OccupancyMap + UniDirsEmbed on the 5th-generation tensor cores: tcgen05.mma
// kind::tf32, accumulators AND the A operands in tensor memory, the object's weights pre-split and resident in shared memory.
// Two entry points share the kernel:
//   * oo_eval_points_tc : Trainer.eval_points / the meshing grid (objnerf/trainer.py:46-69,104-128) at free query points;
//   * oo_render_object (feat == NULL) : render_2D_syn (vmap.py:604-685, trainer.py:130-198) -- the 149 midpoints of every hit
//     ray through the network, occupancy -> termination compositing along the ray (render_rays.py:32-63), masked depth / rgb
//     maps and the compact per-hit record {S = sum_i T_i hp_i, opacity} the winner-only feature path consumes.
//
//   - one CTA per SM, 512 threads, a tile = 128 points = the 128 lanes of tensor memory; FOUR threads share a point (warp w:
//     lane quarter w % 4, sub-thread w / 4): they split the encoder by projection direction and every epilogue by column;
//   - the layers are chained through TMEM: an MMA leaves D[point][unit] (fp32) in TMEM, a thread reads its columns of its
//     point's row with tcgen05.ld, applies bias + ReLU, splits the result into (hi, lo) TF32 halves and writes them back with
//     tcgen05.st as the A operand [point][k] of the next layer -- hidden activations never touch shared memory;
//   - fp32-level accuracy with TF32 inputs by three-term error compensation, exactly like the mma.sync path of the training
//     tile (oo_tile.h): x = hi + lo, a.b ~ lo_a hi_b + hi_a lo_b + hi_a hi_b accumulated in fp32 -> 3 tcgen05.mma per
//     8-wide k-step.  The weights are constant per object, so their (hi, lo) copies are formed ONCE when the CTA starts and
//     stay in shared memory in the canonical K-major core-matrix layout the MMA's shared-memory descriptor addresses
//     ([k / 4][rows][4]: 8 rows x 16 bytes per core matrix, SBO = 128 B between 8-row groups, LBO = rows x 16 B between 16-byte
//     k-chunks); nothing is re-split per tile;
//   - the two tiny output layers (out_alpha 32 -> 1, out_color 32 -> 3) are dot products in registers, partial sums of the
//     four sub-threads combined through shared memory in a fixed order.
// Per 128 points: 99 tcgen05.mma (33 x N64 for [in_layer ; cat_layer's e1 part], 36 x N32, 30 x N32 / N64 for the heads).
// Measured (B200, gpurun_out/tc_probe.json): a M128 N32 K8 kind::tf32 MMA costs ~54 cycles when issued back to back like
// this (skipping the MMAs takes the 256^3 grid from 7.1 to 3.8 ms), so merging the two e1 consumers into N = 64 MMAs pays.
#include "../../include/openobj_b200.h"
#include "oo_common.cuh"
#include "oo_tile.h"

#include <stdlib.h>

using namespace oo;

namespace {

constexpr int TC_M = 128, TC_THREADS = 512;
// tensor-memory columns (32-bit each, 128 lanes)
constexpr int C_E1H = 0, C_E1L = 88, C_E2H = 176, C_E2L = 224, C_HAH = 272, C_HAL = 304, C_HBH = 336, C_HBL = 368, C_D = 400,
              C_D2 = 464;                              // C_D: 64 columns; C_D2: 32 columns for mid1 / mid2
constexpr int TC_COLS = 512;
constexpr int E_COLS = 136;                          // 88 columns of e1 (87 values + a zero) + 48 of e2 (42 + six zeros)
// shared memory (floats): weight blocks in canonical layout (hi copies, then lo copies)
constexpr int K_IN = 88, K_H = 32, K_CATB = 88, K_HDB = 48, N_HD = 64;
// W_E1 = [in_layer ; cat_layer's e1 columns] (64 x 88): both consume e1, so ONE group of N = 64 MMAs forms in_layer's
// pre-activation (accumulator columns 0..31) and cat_layer's e1 contribution (columns 32..63, completed later by the fc2 part)
constexpr int WB_E1 = 0, WB_M1 = WB_E1 + N_HD * K_IN, WB_CATA = WB_M1 + H * K_H,
              WB_M2 = WB_CATA + H * K_H, WB_HDA = WB_M2 + H * K_H, WB_HDB = WB_HDA + N_HD * K_H,
              WB_TOTAL = WB_HDB + N_HD * K_HDB;
constexpr int SB_BIAS = 2 * WB_TOTAL;                 // in, m1, cat, m2 (4 x 32), heads (64: color_linear, clip_linear)
constexpr int SB_WA = SB_BIAS + 4 * H + N_HD;         // out_alpha.weight [32], bias at +32
constexpr int SB_WOC = SB_WA + 36;                    // out_color.weight [3][32], bias at +96
constexpr int SB_PE = SB_WOC + 100;                   // B_layer.weight [21][3]
constexpr int SB_BAR = SB_PE + 64;                    // mbarrier (8 bytes; SB_BAR is even)
constexpr int SB_PT = SB_BAR + 4;                     // per point: t0, t1, t2, z  [4][128]
constexpr int SB_STG = SB_PT + 4 * TC_M;              // encoder staging [136][128]; later the per-point results (see R_*)
// per-point results, aliasing the staging area once the embedding is in tensor memory
constexpr int R_APART = 0;                            // [4][128]    partial out_alpha dot products of the four sub-threads
constexpr int R_CPART = R_APART + 4 * TC_M;           // [4][3][128] partial out_color dot products
constexpr int R_OCC = R_CPART + 12 * TC_M;            // [128] occupancy
constexpr int R_COL = R_OCC + TC_M;                   // [3][128] colour
constexpr int R_T = R_COL + 3 * TC_M;                 // [128] termination weights
constexpr int R_HP = R_T + TC_M;                      // [32][128] clip_linear activations
constexpr int R_SPART = R_HP + H * TC_M;              // [16 warps][2 segments][32] partial S sums
constexpr int R_END = R_SPART + 16 * 2 * H;
static_assert(R_END <= E_COLS * TC_M, "per-point results must fit in the staging area");
constexpr int SB_OPEN = SB_STG + E_COLS * TC_M;       // [2][40] open-ray accumulators {depth, opac, c0, c1, c2, carry, -, -, S[32]}
constexpr int SB_FLOATS = SB_OPEN + 80;
constexpr size_t TC_SMEM = (size_t)SB_FLOATS * 4;
// more than half of the SM's shared memory: two CTAs can never share an SM (each allocates all 512 tensor-memory columns; a
// second CTA would wait for them forever)
static_assert(TC_SMEM > 116 * 1024 && TC_SMEM <= 227 * 1024 && SB_BAR % 2 == 0, "shared-memory map");

// instruction descriptor of tcgen05.mma kind::tf32: D fp32 (bits 4-5 = 1), A and B TF32 (bits 7-9 / 10-12 = 2), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t byte_addr, uint32_t lbo_bytes) {
    // K-major, no swizzle: start >> 4 | LBO (next 16-byte k-chunk) >> 4 at bit 16 | SBO (128 B, next 8 rows) >> 4 at bit 32 |
    // descriptor version 1 (sm_100) at bit 46
    return (uint64_t)((byte_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t id, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(id), "r"(accumulate) : "memory");
}

// D (+)= A [128 x K] * W^T with A's (hi, lo) halves in TMEM columns a_hi / a_lo and W's [n rows] x K in the shared blocks
// w_hi / w_lo (rows_blk = rows the block was staged with: its k-chunk stride is rows_blk x 16 bytes)
__device__ __forceinline__ void layer_mma(uint32_t tm_d, uint32_t tm, int a_hi, int a_lo, uint32_t w_hi, uint32_t w_lo, int K, int n,
                                          int rows_blk, bool first, int dbg = 0) {
    if (dbg == 1) return;
    const uint32_t lbo = (uint32_t)rows_blk * 16u, id = idesc(n);
    for (int s = 0; s < K / 8; ++s) {
        const uint64_t bh = smem_desc(w_hi + (uint32_t)(2 * s) * lbo, lbo), bl = smem_desc(w_lo + (uint32_t)(2 * s) * lbo, lbo);
        umma_ts(tm_d, tm + a_lo + 8 * s, bh, id, (first && s == 0) ? 0u : 1u);      // small terms first
        umma_ts(tm_d, tm + a_hi + 8 * s, bl, id, 1u);
        umma_ts(tm_d, tm + a_hi + 8 * s, bh, id, 1u);
    }
}

// tcgen05.st of 8 consecutive columns of this thread's lane; stores are awaited once per layer (publish_tmem)
__device__ __forceinline__ void tm_st8_nw(uint32_t addr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}

__device__ __forceinline__ void split(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// one weight matrix [rows][k_real] (row-major in theta, `ld` floats per row, columns col0 ..) -> rows row0 .. of a canonical
// (hi, lo) block staged for rows_blk rows
__device__ void stage_block(int tid, float* sm, int blk, const float* __restrict__ W, int ld, int col0, int k_real, int K, int rows,
                            int row0, int rows_blk) {
    for (int i = tid; i < rows * K; i += TC_THREADS) {
        const int j = i / K, k = i - j * K;
        const float v = k < k_real ? W[j * ld + col0 + k] : 0.f;
        float hi, lo;
        split(v, hi, lo);
        const int e = (k >> 2) * (rows_blk * 4) + (row0 + j) * 4 + (k & 3);
        sm[blk + e] = hi;
        sm[WB_TOTAL + blk + e] = lo;
    }
}

// 8 accumulator columns c0 .. c0+7 of this thread's point: bias + ReLU into v
__device__ __forceinline__ void epilogue8(uint32_t tm_lane, int c0, const float* __restrict__ bias, float* v, int d_col = C_D) {
    tm_ld<8>(tm_lane + d_col + c0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j] + bias[c0 + j], 0.f);
}
__device__ __forceinline__ void put_split8(uint32_t tm_lane, int col_hi, int col_lo, const float* v) {
    float hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split(v[i], hi[i], lo[i]);
    tm_st8_nw(tm_lane + col_hi, hi);
    tm_st8_nw(tm_lane + col_lo, lo);
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
// Only warp 0 watches the mbarrier of the MMA group (a spin loop costs issue slots: 15 more warps polling would slow down the
// very thread that is still issuing the MMAs); everybody else sleeps in the block barrier until warp 0 arrives.
__device__ __forceinline__ void layer_done() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float wsum32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct TcArgs {
    const float* theta;
    float scale;
    int* err;
    // eval-points mode
    const float* pts; long long n_pts; float* occ; float* alpha; float* color;
    // render mode: n_obj objects; object o owns the pooled hits [obj_start[o], obj_start[o+1]) (pixels in hit_pix, by pixel
    // order); a CTA takes an equal share of the pooled hits x (n_bins - 1) samples and walks the objects its share touches
    int n_bins, n_obj, W, H;
    const float* const* theta_tab;       // [n_obj] parameter blocks, or NULL: one object with `theta`
    const int* obj_start;                // [n_obj + 1] (device)
    const int* hit_pix;                  // pooled
    const float* T_oc;                   // [n_obj][16] = inv(T_WO) @ T_WC
    const float* half_extent;            // [n_obj][3]
    const float* lin; const float* jitter;
    int jitter_by_rank;
    const float* rays_dir; const float* T_wc;
    uint8_t* mask; float* depth_out; uint8_t* rgb; float* opacity;      // [n_obj][W*H](x3)
    float* ray_rec;                      // pooled [..][OO_RENDER_REC] or NULL
    long long pool_rows;
    int dbg;                             // profiling experiments (OO_TC_DEBUG): 1 = skip the MMAs, 2 = skip the encoder math
};

// ray / oriented-box slab test in the box frame (trainer.py:151-169, utils.py:309-319); near clipped at 0, far + 0.2
__device__ __forceinline__ bool slab(const float* __restrict__ T, const float* __restrict__ he, const float* __restrict__ dc,
                                     float& near, float& far) {
    const float dx = dc[0], dy = dc[1], dz = dc[2];
    const float d[3] = {T[0] * dx + T[1] * dy + T[2] * dz, T[4] * dx + T[5] * dy + T[6] * dz, T[8] * dx + T[9] * dy + T[10] * dz};
    const float o[3] = {T[3], T[7], T[11]};
    near = -INFINITY; far = INFINITY;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const float tmin = __fdiv_rn(-he[ax] - o[ax], d[ax]), tmax = __fdiv_rn(he[ax] - o[ax], d[ax]);   // utils.py:310-311
        near = fmaxf(near, fminf(tmin, tmax));
        far = fminf(far, fmaxf(tmin, tmax));
    }
    const bool hit = (near <= far) && (far > 0.f);                                                      // utils.py:316-318
    near = fmaxf(near, 0.f);                                                                             // trainer.py:168
    far = far + 0.2f;                                                                                    // trainer.py:169
    return hit;
}

template <bool RENDER>
__global__ void __launch_bounds__(TC_THREADS, 1) k_forward_tc(const TcArgs a) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint32_t tm_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = warp >> 2, p = 32 * (warp & 3) + lane;             // four threads (sub 0..3) share point p
    const float* theta = a.theta;
    long long q_begin = 0, q_end = 0;                                   // the current range of points
    long long g0 = 0, g1 = 0;                                           // render: this CTA's share of the pooled hits
    int nmid = 1, obj = 0;
    if (RENDER) {
        nmid = a.n_bins - 1;
        long long total = a.obj_start[a.n_obj];
        if (total > a.pool_rows) total = a.pool_rows;                   // (the host reports the overflow)
        g0 = total * blockIdx.x / gridDim.x; g1 = total * (blockIdx.x + 1) / gridDim.x;
        if (g0 >= g1) return;
        while (obj + 1 < a.n_obj && a.obj_start[obj + 1] <= g0) ++obj;
    } else {
        const long long n_tiles = (a.n_pts + TC_M - 1) / TC_M;
        const long long t0 = n_tiles * blockIdx.x / gridDim.x, t1 = n_tiles * (blockIdx.x + 1) / gridDim.x;
        if (t0 >= t1) return;
        q_begin = t0 * TC_M; q_end = t1 * TC_M < a.n_pts ? t1 * TC_M : a.n_pts;
    }
    // ---- once per CTA: TMEM, the mbarrier
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tm_base_s;
    const uint32_t tm_lane = tm + ((uint32_t)(32 * (warp & 3)) << 16);
    const uint32_t w_hi = (uint32_t)__cvta_generic_to_shared(sm), w_lo = w_hi + WB_TOTAL * 4u;
    uint32_t phase = 0;
    bool ok = true;
    const float* bias = sm + SB_BIAS;
    float* pt = sm + SB_PT;
    float* stg = sm + SB_STG;
    float* open = sm + SB_OPEN;
    int cur = 0;                                                        // which open-ray slot continues from the previous tile
    const float* Tw = a.T_wc;
    const size_t npix = (size_t)a.W * a.H;

  // ---- one pass per object this CTA's share touches (eval-points mode: exactly one pass)
  for (;;) {
    long long seg_g0 = 0;                                               // pool index of local hit 0 of this object
    if (RENDER) {
        if (obj >= a.n_obj || (long long)a.obj_start[obj] >= g1) break;
        seg_g0 = a.obj_start[obj];
        const long long s0 = g0 > seg_g0 ? g0 : seg_g0, s1 = g1 < (long long)a.obj_start[obj + 1] ? g1 : (long long)a.obj_start[obj + 1];
        const int n_hit_o = a.obj_start[obj + 1] - a.obj_start[obj];
        if (s0 >= s1 || n_hit_o <= 1) { ++obj; continue; }              // trainer.py:167: "<= 1 hit -> the object misses"
        q_begin = (s0 - seg_g0) * nmid; q_end = (s1 - seg_g0) * nmid;
        theta = a.theta_tab != nullptr ? a.theta_tab[obj] : a.theta;
    }
    // ---- the object's weights as canonical (hi, lo) blocks, small vectors
    stage_block(tid, sm, WB_E1, theta + OFF_IN_W, E1, 0, E1, K_IN, H, 0, N_HD);              // in_layer            rows 0..31
    stage_block(tid, sm, WB_E1, theta + OFF_CAT_W, H + E1, H, E1, K_IN, H, H, N_HD);         // cat_layer, e1 part  rows 32..63
    stage_block(tid, sm, WB_M1, theta + OFF_M1_W, H, 0, H, K_H, H, 0, H);
    stage_block(tid, sm, WB_CATA, theta + OFF_CAT_W, H + E1, 0, H, K_H, H, 0, H);
    stage_block(tid, sm, WB_M2, theta + OFF_M2_W, H, 0, H, K_H, H, 0, H);
    stage_block(tid, sm, WB_HDA, theta + OFF_CL_W, H + E2, 0, H, K_H, H, 0, N_HD);          // color_linear rows 0..31
    stage_block(tid, sm, WB_HDA, theta + OFF_CP_W, H + E2, 0, H, K_H, H, H, N_HD);          // clip_linear  rows 32..63
    stage_block(tid, sm, WB_HDB, theta + OFF_CL_W, H + E2, H, E2, K_HDB, H, 0, N_HD);
    stage_block(tid, sm, WB_HDB, theta + OFF_CP_W, H + E2, H, E2, K_HDB, H, H, N_HD);
    if (tid < H) {
        sm[SB_BIAS + tid] = theta[OFF_IN_B + tid];
        sm[SB_BIAS + H + tid] = theta[OFF_M1_B + tid];
        sm[SB_BIAS + 2 * H + tid] = theta[OFF_CAT_B + tid];
        sm[SB_BIAS + 3 * H + tid] = theta[OFF_M2_B + tid];
        sm[SB_BIAS + 4 * H + tid] = theta[OFF_CL_B + tid];
        sm[SB_BIAS + 5 * H + tid] = theta[OFF_CP_B + tid];
        sm[SB_WA + tid] = theta[OFF_A_W + tid];
    }
    if (tid < 3 * H) sm[SB_WOC + tid] = theta[OFF_OC_W + tid];
    if (tid < 3) sm[SB_WOC + 3 * H + tid] = theta[OFF_OC_B + tid];
    if (tid == 0) sm[SB_WA + H] = theta[OFF_A_B];
    if (tid < NDIR * 3) sm[SB_PE + tid] = theta[OFF_PE_B + tid];
    if (tid < 80) sm[SB_OPEN + tid] = (tid == 5 || tid == 45) ? 1.f : 0.f;      // carry (free-probability product) starts at 1
    cur = 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the weight blocks are read by the tensor core (async proxy)
    __syncthreads();
    const float* Toc = RENDER ? a.T_oc + 16 * (size_t)obj : nullptr;
    const float* he = RENDER ? a.half_extent + 3 * (size_t)obj : nullptr;

    for (long long q0 = q_begin; q0 < q_end; q0 += TC_M) {
        const int npts = (int)((q_end - q0) < TC_M ? (q_end - q0) : TC_M);
        // ---- (0) the tile's points: scaled coordinates t = x / scale (embedding.py:47) and, when rendering, the sample depth
        if (tid < TC_M) {
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, zm = 0.f;
            if (tid < npts) {
                if (RENDER) {
                    // midpoints of the jittered stratified bins (trainer.py:174-178, utils.py:342-379), points = o + d z
                    const long long q = q0 + tid;
                    const int j = (int)(q / nmid), kk = (int)(q - (long long)j * nmid);
                    const int pix = a.hit_pix[seg_g0 + j];
                    float near, far;
                    slab(Toc, he, a.rays_dir + (size_t)pix * 3, near, far);
                    const float range = __fsub_rn(far, near), blen = __fdiv_rn(range, (float)a.n_bins);
                    const float* u = a.jitter + (size_t)(a.jitter_by_rank ? j : pix) * a.n_bins;
                    const float za = __fadd_rn(__fadd_rn(__fmul_rn(range, a.lin[kk]), near), __fmul_rn(u[kk], blen));
                    const float zb = __fadd_rn(__fadd_rn(__fmul_rn(range, a.lin[kk + 1]), near), __fmul_rn(u[kk + 1], blen));
                    zm = __fmul_rn(0.5f, __fadd_rn(zb, za));
                    const float* dc = a.rays_dir + (size_t)pix * 3;
                    const float dx = dc[0], dy = dc[1], dz = dc[2];
                    const float wx = Tw[0] * dx + Tw[1] * dy + Tw[2] * dz, wy = Tw[4] * dx + Tw[5] * dy + Tw[6] * dz,
                                wz = Tw[8] * dx + Tw[9] * dy + Tw[10] * dz;
                    t0 = (Tw[3] + wx * zm) / a.scale; t1 = (Tw[7] + wy * zm) / a.scale; t2 = (Tw[11] + wz * zm) / a.scale;
                } else {
                    const float* x = a.pts + 3 * (q0 + tid);
                    t0 = x[0] / a.scale; t1 = x[1] / a.scale; t2 = x[2] / a.scale;
                }
            }
            pt[tid] = t0; pt[TC_M + tid] = t1; pt[2 * TC_M + tid] = t2; pt[3 * TC_M + tid] = zm;
        }
        __syncthreads();
        // ---- (1) encoder (embedding.py:46-55; arithmetic of oo_tile.h phase 1): sub-thread s takes directions s, s + 4, ...;
        // value (band k, direction d) -> staging column 3 + 21 k + d (+ 1 past e1's zero pad)
        {
            const float t0 = pt[p], t1 = pt[TC_M + p], t2 = pt[2 * TC_M + p];
            if (sub == 0) {
                stg[0 * TC_M + p] = t0; stg[1 * TC_M + p] = t1; stg[2 * TC_M + p] = t2;
            } else if (sub == 1) {
                stg[87 * TC_M + p] = 0.f;
#pragma unroll
                for (int c = 130; c < E_COLS; ++c) stg[c * TC_M + p] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int d = sub + 4 * u;
                if (d < NDIR) {
                    const float proj = sm[SB_PE + 3 * d] * t0 + sm[SB_PE + 3 * d + 1] * t1 + sm[SB_PE + 3 * d + 2] * t2;
                    float sn = proj, cs = proj;
                    if (a.dbg != 2) sincosf(proj * PI_F, &sn, &cs);
#pragma unroll
                    for (int k = 0; k < NBAND; ++k) {
                        const int row = 3 + NDIR * k + d;
                        stg[(row < E1 ? row : row + 1) * TC_M + p] = sn;
                        const float s2 = 2.f * sn * cs, c2 = (cs - sn) * (cs + sn);
                        sn = s2; cs = c2;
                    }
                }
            }
        }
        __syncthreads();
        // ---- (2) staging -> (hi, lo) A operands in tensor memory, 8 columns at a time (sub-thread s: groups s, s + 4, ...)
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const int g = sub + 4 * u;
            if (g < E_COLS / 8) {
                float e8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) e8[i] = stg[(8 * g + i) * TC_M + p];
                if (8 * g < 88) put_split8(tm_lane, C_E1H + 8 * g, C_E1L + 8 * g, e8);
                else put_split8(tm_lane, C_E2H + 8 * g - 88, C_E2L + 8 * g - 88, e8);
            }
        }
        float v[16];
        // ---- in_layer: relu(W_in e1 + b) -> HA
        publish_tmem();
        if (warp == 0) {
            if (lane == 0) {
                layer_mma(tm + C_D, tm, C_E1H, C_E1L, w_hi + WB_E1 * 4u, w_lo + WB_E1 * 4u, K_IN, N_HD, N_HD, true, a.dbg);
                commit(bar);
            }
            __syncwarp();
            ok &= mbar_wait(bar, phase);
        }
        phase ^= 1;
        layer_done();
        epilogue8(tm_lane, 8 * sub, bias, v);
        put_split8(tm_lane, C_HAH + 8 * sub, C_HAL + 8 * sub, v);
        // ---- mid1 -> HB (fc2)
        publish_tmem();
        if (warp == 0) {
            if (lane == 0) {
                layer_mma(tm + C_D2, tm, C_HAH, C_HAL, w_hi + WB_M1 * 4u, w_lo + WB_M1 * 4u, K_H, H, H, true, a.dbg);
                commit(bar);
            }
            __syncwarp();
            ok &= mbar_wait(bar, phase);
        }
        phase ^= 1;
        layer_done();
        epilogue8(tm_lane, 8 * sub, bias + H, v, C_D2);
        put_split8(tm_lane, C_HBH + 8 * sub, C_HBL + 8 * sub, v);
        // ---- cat_layer on [fc2, e1] -> HA (fc3)
        publish_tmem();
        if (warp == 0) {
            if (lane == 0) {
                layer_mma(tm + C_D + H, tm, C_HBH, C_HBL, w_hi + WB_CATA * 4u, w_lo + WB_CATA * 4u, K_H, H, H, false, a.dbg);
                commit(bar);
            }
            __syncwarp();
            ok &= mbar_wait(bar, phase);
        }
        phase ^= 1;
        layer_done();
        epilogue8(tm_lane, 8 * sub, bias + 2 * H, v, C_D + H);
        put_split8(tm_lane, C_HAH + 8 * sub, C_HAL + 8 * sub, v);
        // ---- mid2 -> HB (fc4); this sub-thread's share of out_alpha (model.py:86-88)
        publish_tmem();
        if (warp == 0) {
            if (lane == 0) {
                layer_mma(tm + C_D2, tm, C_HAH, C_HAL, w_hi + WB_M2 * 4u, w_lo + WB_M2 * 4u, K_H, H, H, true, a.dbg);
                commit(bar);
            }
            __syncwarp();
            ok &= mbar_wait(bar, phase);
        }
        phase ^= 1;
        layer_done();
        epilogue8(tm_lane, 8 * sub, bias + 3 * H, v, C_D2);
        put_split8(tm_lane, C_HBH + 8 * sub, C_HBL + 8 * sub, v);
        {
            float r = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) r += sm[SB_WA + 8 * sub + j] * v[j];
            stg[R_APART + sub * TC_M + p] = r;
        }
        // ---- heads on [fc4, e2]: color_linear (and clip_linear when rendering) in one MMA group
        publish_tmem();
        if (warp == 0) {
            if (lane == 0) {
                layer_mma(tm + C_D, tm, C_HBH, C_HBL, w_hi + WB_HDA * 4u, w_lo + WB_HDA * 4u, K_H, RENDER ? N_HD : H, N_HD, true, a.dbg);
                layer_mma(tm + C_D, tm, C_E2H, C_E2L, w_hi + WB_HDB * 4u, w_lo + WB_HDB * 4u, K_HDB, RENDER ? N_HD : H, N_HD, false, a.dbg);
                commit(bar);
            }
            __syncwarp();
            ok &= mbar_wait(bar, phase);
        }
        phase ^= 1;
        layer_done();
        if (RENDER) {
            // sub-threads 0, 1: columns 0..31 = color_linear; 2, 3: columns 32..63 = clip_linear (kept per point for S)
            epilogue8(tm_lane, 16 * sub, bias + 4 * H, v);
            epilogue8(tm_lane, 16 * sub + 8, bias + 4 * H, v + 8);
            if (sub < 2) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float r = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) r += sm[SB_WOC + ch * H + 16 * sub + j] * v[j];
                    stg[R_CPART + (sub * 3 + ch) * TC_M + p] = r;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) stg[R_HP + (16 * (sub - 2) + j) * TC_M + p] = v[j];
            }
        } else {
            epilogue8(tm_lane, 8 * sub, bias + 4 * H, v);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float r = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) r += sm[SB_WOC + ch * H + 8 * sub + j] * v[j];
                stg[R_CPART + (sub * 3 + ch) * TC_M + p] = r;
            }
        }
        // the accumulator columns are re-used by the next tile's first MMA: every thread is past its loads here.  A missed
        // completion (never observed) ends the CTA's work for all of its threads together.
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        ok = __syncthreads_and(ok ? 1 : 0) != 0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!ok) break;
        // ---- out_alpha (x10) -> occupancy (render_rays.py:13), out_color -> sigmoid (model.py:96): partial sums in a fixed order
        if (tid < TC_M) {
            const float ra = sm[SB_WA + H] + ((stg[R_APART + tid] + stg[R_APART + TC_M + tid]) +
                                              (stg[R_APART + 2 * TC_M + tid] + stg[R_APART + 3 * TC_M + tid]));
            const float al = ra * 10.f, oc_ = 1.f / (1.f + expf(-al));
            float col[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float r = sm[SB_WOC + 3 * H + ch];
                if (RENDER) r += stg[R_CPART + ch * TC_M + tid] + stg[R_CPART + (3 + ch) * TC_M + tid];
                else r += (stg[R_CPART + ch * TC_M + tid] + stg[R_CPART + (3 + ch) * TC_M + tid]) +
                          (stg[R_CPART + (6 + ch) * TC_M + tid] + stg[R_CPART + (9 + ch) * TC_M + tid]);
                col[ch] = sigmoidf_(r);
            }
            if (RENDER) {
                stg[R_OCC + tid] = oc_;
                stg[R_COL + tid] = col[0]; stg[R_COL + TC_M + tid] = col[1]; stg[R_COL + 2 * TC_M + tid] = col[2];
            } else if (tid < npts) {
                const long long gq = q0 + tid;
                if (a.alpha != nullptr) a.alpha[gq] = al;
                if (a.occ != nullptr) a.occ[gq] = oc_;
                a.color[3 * gq] = col[0]; a.color[3 * gq + 1] = col[1]; a.color[3 * gq + 2] = col[2];
            }
        }
        __syncthreads();
        if (RENDER) {
            // ---- compositing: the tile holds the tail of the open ray (segment A) and maybe the head of the next (B)
            const int jA = (int)(q0 / nmid);
            const int kA = (int)(q0 - (long long)jA * nmid);
            const int lenA = npts < nmid - kA ? npts : nmid - kA;
            const int lenB = npts - lenA;                                // < nmid because nmid > 128
            if (warp < 2) {
                const int pa = warp == 0 ? 0 : lenA, len = warp == 0 ? lenA : lenB;
                float* acc_o = open + (warp == 0 ? cur : cur ^ 1) * 40;
                if (len > 0) {
                    float carry = acc_o[5];
                    float sd = 0.f, so = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
                    for (int b0 = 0; b0 < len; b0 += 32) {
                        const int pp = pa + b0 + lane;
                        const bool in = b0 + lane < len;
                        const float o = in ? stg[R_OCC + pp] : 0.f;
                        float inc = in ? (1.f - o + 1e-10f) : 1.f;                      // render_rays.py:41
#pragma unroll
                        for (int s = 1; s < 32; s <<= 1) {
                            const float t = __shfl_up_sync(0xffffffffu, inc, s);
                            if (lane >= s) inc *= t;
                        }
                        float ex = __shfl_up_sync(0xffffffffu, inc, 1);
                        if (lane == 0) ex = 1.f;
                        const float T = o * carry * ex;                                  // render_rays.py:43
                        carry *= __shfl_sync(0xffffffffu, inc, 31);
                        if (in) {
                            stg[R_T + pp] = T;
                            sd += T * pt[3 * TC_M + pp];
                            so += T;
                            s0 += T * stg[R_COL + pp];
                            s1 += T * stg[R_COL + TC_M + pp];
                            s2 += T * stg[R_COL + 2 * TC_M + pp];
                        }
                    }
                    sd = wsum32(sd); so = wsum32(so); s0 = wsum32(s0); s1 = wsum32(s1); s2 = wsum32(s2);
                    if (lane == 0) {
                        acc_o[0] += sd; acc_o[1] += so; acc_o[2] += s0; acc_o[3] += s1; acc_o[4] += s2; acc_o[5] = carry;
                    }
                }
            }
            __syncthreads();
            // S[j] += sum_p T_p hp[j][p] of both segments with all 16 warps: warp w sums points 8 w .. 8 w + 7 (lane = hidden unit),
            // then 64 threads add the 16 partials in a fixed order
            {
                float sa = 0.f, sb = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int pp = 8 * warp + i;
                    const float tv = pp < npts ? stg[R_T + pp] * stg[R_HP + lane * TC_M + pp] : 0.f;
                    if (pp < lenA) sa += tv; else sb += tv;
                }
                stg[R_SPART + (warp * 2 + 0) * H + lane] = sa;
                stg[R_SPART + (warp * 2 + 1) * H + lane] = sb;
            }
            __syncthreads();
            if (tid < 2 * H) {
                const int seg = tid >> 5;
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < 16; ++w) t += stg[R_SPART + (w * 2 + seg) * H + lane];
                open[(seg == 0 ? cur : cur ^ 1) * 40 + 8 + lane] += t;
            }
            __syncthreads();
            // ---- a ray that ended in this tile: mask test (vmap.py:665,672), outputs, compact record; its slot is reset
            if (lenA > 0 && kA + lenA == nmid) {
                const float* acc_o = open + cur * 40;
                const int pix = a.hit_pix[seg_g0 + jA];
                const float d = acc_o[0], op = acc_o[1];
                if (tid < OO_RENDER_REC && a.ray_rec != nullptr)
                    a.ray_rec[(size_t)(seg_g0 + jA) * OO_RENDER_REC + tid] = tid < H ? acc_o[8 + tid] : (tid == H ? op : 0.f);
                if (tid == 0) {
                    float near, far;
                    slab(Toc, he, a.rays_dir + (size_t)pix * 3, near, far);
                    const bool bad = d < near || d > far || op < 0.9f;
                    const size_t o_pix = (size_t)obj * npix + pix;
                    a.mask[o_pix] = bad ? 0 : 1;
                    a.depth_out[o_pix] = bad ? 0.f : d;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const float vv = __fmul_rn(acc_o[2 + ch], 255.f);                            // vmap.py:671
                        a.rgb[o_pix * 3 + ch] = bad ? 0 : (uint8_t)(int)vv;
                    }
                    if (a.opacity) a.opacity[o_pix] = op;
                }
                __syncthreads();
                if (tid < 40) open[cur * 40 + tid] = tid == 5 ? 1.f : 0.f;
                cur ^= 1;                                               // segment B (if any) is now the open ray
                __syncthreads();
            }
        }
    }
    if (!RENDER || !ok) break;
    ++obj;
    __syncthreads();                                                    // the next object's weights replace these
  }
    if (!ok && tid == 0 && a.err != nullptr) atomicExch(a.err, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"((uint32_t)TC_COLS) : "memory");
}

int n_sm_of_current_device(int* out) {
    int dev = 0, n = 148;
    OO_CUDA(cudaGetDevice(&dev));
    OO_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    *out = n;
    return 0;
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
        OO_CUDA(cudaFuncSetAttribute(k_forward_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        attr_set.cur() = 1;
    }
    int n_sm = 148;
    if (int rc = n_sm_of_current_device(&n_sm)) return rc;
    const long long n_tiles = (n_pts + TC_M - 1) / TC_M;
    const int grid = (int)(n_tiles < n_sm ? n_tiles : n_sm);
    TcArgs a = {};
    a.theta = theta; a.scale = pe_scale; a.err = err_flag;
    a.pts = pts; a.n_pts = n_pts; a.occ = occ; a.alpha = alpha; a.color = color;
    if (const char* e = getenv("OO_TC_DEBUG")) a.dbg = atoi(e);
    k_forward_tc<false><<<grid, TC_THREADS, TC_SMEM, (cudaStream_t)stream>>>(a);
    OO_LAUNCH_CHECK();
    return 0;
}

// ---- hit lists of MANY objects in three parallel passes (the single-object path of oo_render.cu scans with one CTA):
//   k_hit_count : slab test of every (object, pixel), hits per 256-pixel block
//   k_hit_scan  : per object, exclusive prefix over its blocks + the object's total; then (last CTA) the prefix over objects
//   k_hit_fill  : the slab test again, rank inside the block by ballot -> pixel ids at obj_start[o] + block prefix + rank
// Pixels appear in increasing order inside an object's list (the reference's rank order, trainer.py:166-176).
namespace {
constexpr int HB = 256;
__global__ void __launch_bounds__(HB) k_hit_count(int n_pix, const float* __restrict__ T_oc, const float* __restrict__ half_extent,
                                                  const float* __restrict__ rays_dir, int* __restrict__ blk_cnt) {
    const int o = blockIdx.y, i = blockIdx.x * HB + threadIdx.x;
    float near, far;
    const bool hit = i < n_pix && slab(T_oc + 16 * (size_t)o, half_extent + 3 * (size_t)o, rays_dir + (size_t)i * 3, near, far);
    const int n = __syncthreads_count(hit ? 1 : 0);
    if (threadIdx.x == 0) blk_cnt[(size_t)o * gridDim.x + blockIdx.x] = n;
}
__global__ void __launch_bounds__(1024) k_hit_scan(int n_blk, int n_obj, int* __restrict__ blk_cnt, int* __restrict__ obj_start,
                                                   int* __restrict__ done) {
    __shared__ int sh[32];
    __shared__ int s_last;
    const int o = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    int* c = blk_cnt + (size_t)o * n_blk;
    const int per = (n_blk + 1023) / 1024, b = min(tid * per, n_blk), e = min(b + per, n_blk);
    int n = 0;
    for (int i = b; i < e; ++i) n += c[i];
    int sc = n;
#pragma unroll
    for (int q = 1; q < 32; q <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, sc, q);
        if (lane >= q) sc += t;
    }
    if (lane == 31) sh[wv] = sc;
    __syncthreads();
    if (wv == 0) {
        int v = sh[lane];
        const int own = v;
#pragma unroll
        for (int q = 1; q < 32; q <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, q);
            if (lane >= q) v += t;
        }
        sh[lane] = v - own;
        if (lane == 31) obj_start[o + 1] = v;                 // the object's total for now; turned into a prefix below
    }
    __syncthreads();
    int run = sh[wv] + sc - n;
    for (int i = b; i < e; ++i) {
        const int v = c[i];
        c[i] = run;
        run += v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(done, 1) == n_obj - 1 ? 1 : 0;
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        int acc = 0;
        obj_start[0] = 0;
        for (int q = 0; q < n_obj; ++q) {
            acc += reinterpret_cast<volatile int*>(obj_start)[q + 1];
            obj_start[q + 1] = acc;
        }
        *done = 0;
    }
}
__global__ void __launch_bounds__(HB) k_hit_fill(int n_pix, const float* __restrict__ T_oc, const float* __restrict__ half_extent,
                                                 const float* __restrict__ rays_dir, const int* __restrict__ blk_cnt,
                                                 const int* __restrict__ obj_start, long long pool_rows, int* __restrict__ hit_pix) {
    __shared__ int s_w[HB / 32];
    const int o = blockIdx.y, i = blockIdx.x * HB + threadIdx.x, lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    float near, far;
    const bool hit = i < n_pix && slab(T_oc + 16 * (size_t)o, half_extent + 3 * (size_t)o, rays_dir + (size_t)i * 3, near, far);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_w[wv] = __popc(m);
    __syncthreads();
    int off = 0;
    for (int w = 0; w < wv; ++w) off += s_w[w];
    if (hit) {
        const long long g = (long long)obj_start[o] + blk_cnt[(size_t)o * gridDim.x + blockIdx.x] + off + __popc(m & ((1u << lane) - 1u));
        if (g < pool_rows) hit_pix[g] = i;
    }
}
__global__ void k_set_start(const int* __restrict__ n_hit, int* __restrict__ obj_start) {
    obj_start[0] = 0;
    obj_start[1] = n_hit[0];
}
}  // namespace

// K5 on the tensor cores for ONE object: called by oo_render_object (oo_render.cu) when no dense feature map is requested;
// `list` / n_hit come from its own compaction, obj_start2 = 2 ints of device scratch.
namespace oo {
int render_tc_launch(const oo_render_args* ra, const int* list, int* obj_start2, const float* lin, cudaStream_t st) {
    OO_REQUIRE(ra->n_bins - 1 > TC_M, "oo_render_object: the tensor-core path needs more than %d samples per ray", TC_M);
    static PerDevice attr_set;
    if (!attr_set.cur()) {
        OO_CUDA(cudaFuncSetAttribute(k_forward_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        attr_set.cur() = 1;
    }
    int n_sm = 148;
    if (int rc = n_sm_of_current_device(&n_sm)) return rc;
    k_set_start<<<1, 1, 0, st>>>(ra->n_hit, obj_start2);
    OO_LAUNCH_CHECK();
    TcArgs a = {};
    a.theta = ra->theta1; a.scale = ra->scale; a.err = ra->tc_err;
    a.n_bins = ra->n_bins; a.n_obj = 1; a.W = ra->W; a.H = ra->H;
    a.obj_start = obj_start2; a.hit_pix = list; a.T_oc = ra->T_oc; a.half_extent = ra->half_extent;
    a.lin = lin; a.jitter = ra->jitter; a.jitter_by_rank = ra->jitter_by_rank; a.rays_dir = ra->rays_dir; a.T_wc = ra->T_wc;
    a.mask = ra->mask; a.depth_out = ra->depth; a.rgb = ra->rgb; a.opacity = ra->opacity; a.ray_rec = ra->ray_rec;
    a.pool_rows = (long long)ra->W * ra->H;
    k_forward_tc<true><<<n_sm, TC_THREADS, TC_SMEM, st>>>(a);
    OO_LAUNCH_CHECK();
    return 0;
}
}  // namespace oo

// ---- a19 for ALL objects of a rank in one call (BASELINE config 5): hit lists of every object (three parallel passes), then
// ONE launch of the tcgen05 kernel over the pooled hits.  See include/openobj_b200.h.
extern "C" int oo_render_frame(const oo_render_frame_args* f, void* stream) {
    OO_REQUIRE(f && f->theta && f->T_wc && f->T_oc && f->half_extent && f->rays_dir && f->jitter && f->lin,
               "oo_render_frame: null input");
    OO_REQUIRE(f->mask && f->depth && f->rgb && f->hit_pix && f->obj_start && f->scratch, "oo_render_frame: null output / scratch");
    OO_REQUIRE(f->n_obj >= 1 && f->W > 0 && f->H > 0 && f->pool_rows > 0, "oo_render_frame: bad size");
    OO_REQUIRE(f->n_bins - 1 > TC_M && f->n_bins <= 160, "oo_render_frame: need %d < n_bins - 1, n_bins <= 160", TC_M);
    cudaStream_t st = (cudaStream_t)stream;
    const int n_pix = f->W * f->H, n_blk = (n_pix + HB - 1) / HB;
    OO_REQUIRE((long long)f->scratch_ints >= (long long)f->n_obj * n_blk + 1, "oo_render_frame: scratch too small");
    int* blk_cnt = f->scratch;
    int* done = f->scratch + (size_t)f->n_obj * n_blk;
    static PerDevice attr_set;
    if (!attr_set.cur()) {
        OO_CUDA(cudaFuncSetAttribute(k_forward_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        attr_set.cur() = 1;
    }
    int n_sm = 148;
    if (int rc = n_sm_of_current_device(&n_sm)) return rc;
    OO_CUDA(cudaMemsetAsync(done, 0, sizeof(int), st));
    k_hit_count<<<dim3(n_blk, f->n_obj), HB, 0, st>>>(n_pix, f->T_oc, f->half_extent, f->rays_dir, blk_cnt);
    OO_LAUNCH_CHECK();
    k_hit_scan<<<f->n_obj, 1024, 0, st>>>(n_blk, f->n_obj, blk_cnt, f->obj_start, done);
    OO_LAUNCH_CHECK();
    k_hit_fill<<<dim3(n_blk, f->n_obj), HB, 0, st>>>(n_pix, f->T_oc, f->half_extent, f->rays_dir, blk_cnt, f->obj_start, f->pool_rows,
                                                     f->hit_pix);
    OO_LAUNCH_CHECK();
    TcArgs a = {};
    a.scale = f->scale; a.err = f->tc_err;
    a.n_bins = f->n_bins; a.n_obj = f->n_obj; a.W = f->W; a.H = f->H;
    a.theta_tab = f->theta; a.obj_start = f->obj_start; a.hit_pix = f->hit_pix; a.T_oc = f->T_oc; a.half_extent = f->half_extent;
    a.lin = f->lin; a.jitter = f->jitter; a.jitter_by_rank = 0; a.rays_dir = f->rays_dir; a.T_wc = f->T_wc;
    a.mask = f->mask; a.depth_out = f->depth; a.rgb = f->rgb; a.opacity = nullptr; a.ray_rec = f->ray_rec;
    a.pool_rows = f->pool_rows;
    k_forward_tc<true><<<n_sm, TC_THREADS, TC_SMEM, st>>>(a);
    OO_LAUNCH_CHECK();
    return 0;
}
