// One fused tile of the per-object NeRF training step: 10 rays x 10 samples of ONE object.
//
//   encode (UniDirsEmbed, embedding.py:46-55) -> OccupancyMap MLP (model.py:61-103)
//   -> occupancy/termination compositing (render_rays.py:6-63) -> losses (loss.py:5-103)
//   -> hand-derived backward of all of it, weight gradients accumulated in registers.
//
// The tile is written as a list of PHASES; inside a phase every thread works on private outputs and the
// only communication is through shared memory, with a block barrier between phases.  That makes the
// same source compilable for the device (oo_train.cu: phase<k>(threadIdx.x); __syncthreads();) and for
// the host-side tile emulator the CPU tests use to check the index algebra (tests/emu: for every tid:
// phase<k>(tid)).  The emulator is test infrastructure; the product path is the CUDA kernel only.
//
// Data layout (see oo_layout.h): activations feature-major  A[row][p]  (p = 10*ray + sample),
// weights row-major [out][in] with padded row strides so that 8 consecutive rows hit 8 distinct
// 16-byte bank groups; all inner loops are 4x4 register tiles fed by 128-bit shared loads
// (8 FFMA per LDS.128).  The 512-wide out_clip layer is linear and only ever consumed through the
// compositing sum, so it is applied once per RAY to S = sum_i T_i hp_i instead of once per sample
// (SURVEY 8d "legal algebraic restructure"): feat = W S + b*opacity.
#pragma once
#include <math.h>
#include <stdint.h>

#include "oo_layout.h"

#ifdef __CUDACC__
#define OO_DEV __device__ __forceinline__
#define OO_LDG(p) __ldg(p)
// 16-byte asynchronous global -> shared copy (LDGSTS); completion awaited with OO_CP_ASYNC_WAIT before a barrier
#define OO_CP_ASYNC16(dst, src)                                                                          \
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src))
#define OO_CP_ASYNC4(dst, src)                                                                           \
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src))
#define OO_CP_ASYNC_WAIT() asm volatile("cp.async.wait_all;" ::: "memory")
#else
#define OO_DEV inline
#define OO_LDG(p) (*(p))
#define OO_CP_ASYNC16(dst, src) memcpy((dst), (src), 16)
#define OO_CP_ASYNC4(dst, src) memcpy((dst), (src), 4)
#define OO_CP_ASYNC_WAIT() ((void)0)
#include <string.h>
struct alignas(16) float4 {
    float x, y, z, w;
};
#endif

namespace oo {

OO_DEV float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
OO_DEV void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
OO_DEV float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
OO_DEV float sgnf_(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
#ifdef __CUDACC__
OO_DEV float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

// Weight-gradient accumulators of one thread, live across all tiles of an object.  The five weight matrices are
// accumulated as m16n8k8 C fragments: output tile u = warp + NWARPS i of a GEMM with MT row tiles (16 output rows j each)
// is (m = u % MT, n = u / MT); register r of fragment i holds row j = 16 m + lane/4 + 8 (r >> 1) and input column
// k = 8 n + 2 (lane % 4) + (r & 1)   (see wfrag_row / wfrag_col).
//
// On the device they do NOT live in registers between phases: each warp owns 64 columns x 32 lanes of TENSOR MEMORY
// (tcgen05.alloc, 256 columns per CTA: lane quarter = warp % 4, column block = warp / 4) and a phase that updates an
// accumulator loads it with tcgen05.ld at its start and stores it back with tcgen05.st at its end (OO_ACC / OO_ACC_PUT).
// That frees 52 registers per thread for the GEMM loops (128 registers x 512 threads is the whole register file).
// The host build (CPU tile emulator, tests only) keeps plain arrays.
constexpr int AC_IN = 0, AC_CAT = 8, AC_HD = 16, AC_M1 = 32, AC_M2 = 36, AC_GM = 40, AC_OC = 44, AC_B = 45, AC_PE = 46,
              AC_MV = 47, AC_LOSS = 48, AC_TRIG = 52, AC_COLS = 64;
struct TileAcc {
#ifdef __CUDACC__
    uint32_t tm;       // tensor-memory address of this warp's column block (lane field = 32 * (warp % 4))
#else
    float w_in[8];     // in_layer   : 2 x 11 tiles -> 2 fragments
    float w_cat[8];    // cat_layer  : 2 x 15 tiles -> 2 fragments
    float w_m1[4];     // mid1       : 2 x 4 tiles, two warps per tile (k-split, see gemm_bwd_w32)
    float w_m2[4];     // mid2
    float w_hd[12];    // [color_linear ; clip_linear] : 4 x 10 tiles -> 3 fragments (part features off: 2 x 10 -> 2)
    float g_oc[1];     // out_color.weight (tid<96), out_alpha.weight (96<=tid<128)
    float g_b[1];      // biases of the six hidden layers (tid<192), out_color.bias (192..194), out_alpha.bias (195)
    float g_pe[1];     // B_layer.weight (tid<63)
    float g_gm[4];     // M = sum_r B_r S_r S_r^T, row tid/8, cols 4(tid%8)..+3   (out_clip gradient, see phase 13)
    float g_mv[1];     // m = sum_r B_r opac_r S_r (tid<32), beta = sum_r B_r opac_r^2 (tid 32)
    float loss[4];     // per-ray loss partials (thread 32 r, r < 10): depth, colour, opacity, feature
#endif
};

#ifdef __CUDACC__
template <int N>
OO_DEV void tm_ld(uint32_t addr, float* v) {
    static_assert(N == 1 || N == 2 || N == 4 || N == 8 || N == 12, "tm_ld sizes");
    uint32_t r[12];
    if constexpr (N == 1) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(addr));
    } else if constexpr (N == 2) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
    } else if constexpr (N == 4) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
        if constexpr (N == 12)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]) : "r"(addr + 8));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __uint_as_float(r[i]);
}
template <int N>
OO_DEV void tm_st(uint32_t addr, const float* v) {
    static_assert(N == 1 || N == 2 || N == 4 || N == 8 || N == 12, "tm_st sizes");
    uint32_t r[12];
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = __float_as_uint(v[i]);
    if constexpr (N == 1) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr), "r"(r[0]) : "memory");
    } else if constexpr (N == 2) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(r[0]), "r"(r[1]) : "memory");
    } else if constexpr (N == 4) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                     "r"(r[3]) : "memory");
    } else {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]), "r"(r[1]),
                     "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        if constexpr (N == 12)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr + 8), "r"(r[8]), "r"(r[9]),
                         "r"(r[10]), "r"(r[11]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// a phase's view of one accumulator: `name` is a local array on the device (loaded from tensor memory), the member itself on the host
#define OO_ACC(name, N, COL) float name[N]; tm_ld<N>(a.tm + (COL), name)
#define OO_ACC_PUT(name, N, COL) tm_st<N>(a.tm + (COL), name)
#else
#define OO_ACC(name, N, COL) float* name = a.name
#define OO_ACC_PUT(name, N, COL) ((void)0)
#endif

OO_DEV void acc_zero(TileAcc& a) {
#ifdef __CUDACC__
    float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < AC_COLS; c += 8) tm_st<8>(a.tm + c, z);
#else
    a = TileAcc{};
#endif
}

// everything a tile needs that is uniform across the block
struct TileCtx {
    const float* pcs;          // first ray of the tile: [nrays][S][3]
    const float* z;            // [nrays][S]
    const float* gt_depth;     // [nrays]
    const uint8_t* gt_rgb;     // [nrays][3]
    const uint8_t* labels;     // [nrays]
    const int32_t* feat_row;   // [nrays] or nullptr
    const float* feat_table;   // [rows][512]
    const float* theta;        // this object's parameter block
    const float* derived;      // this CTA's scratch holding the current object's G = W^T W, wb = W^T b, bb (gram_stage)
    float* rayrec;             // [nrays][RAYREC] per-ray record for K4, first ray of the tile
    float* slab;               // gradient slab of the current (CTA, object) slot
    int nrays;                 // <= RT
    int npts;                  // valid points in the tile (nrays * S when training)
    int flags;                 // OO_FLAG_NO_OBJ (2) / OO_FLAG_NO_SEM (4) for this step
    float scale;               // UniDirsEmbed scale
    float inv1, invs;          // 1/(n(label==1)+1e-10), 1/(n(label!=2)+1e-10) of this object in this step
    float cs, os, fs;          // colour / opacity / feature scaling (loss.py:6)
    // general-upstream backward (oo_forward_bwd: the autograd surface of vmap(fc_model)), first point of the tile:
    const float* up_alpha;     // [npts]      dL/d alpha (alpha = 10 x out_alpha's output, model.py:88)
    const float* up_color;     // [npts][3]   dL/d color (after the sigmoid)
    const float* up_hp;        // [npts][32]  dL/d hp = d_clip W_ocl (k_clip_dhp), or nullptr
    float* hp_out;             // [npts][32]  clip_linear's output (post-ReLU) for the out_clip weight gradient, or nullptr
};

// ------------------------------------------------------------------------------------------------
// tensor-core building blocks
//
// Every contraction of the tile runs on the tensor pipe as mma.sync m16n8k8 TF32 with three-term error compensation
// (x = hi + lo, hi = x with the low 13 mantissa bits cleared, lo = x - hi exactly; a.b ~ lo_a hi_b + hi_a lo_b +
// hi_a hi_b, fp32 accumulate): the result agrees with an fp32 FMA chain to ~1e-6, which the parity tolerance
// (rel 1e-4 on losses) needs and plain TF32 does not give.  Fragment element maps (g = lane / 4, t = lane % 4):
//   A 16x8: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B 8x8: b0 (t, g) b1 (t+4, g);
//   C 16x8: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
// The k index inside a k-step is a dummy, so a fragment "slot" may hold any k as long as A and B agree: the forward
// GEMM puts k0+2t in slot t and k0+2t+1 in slot t+4, which makes the weight fragment one 64-bit load and the activation
// fragment bank-conflict free with the row stride PS = 100.
// Work split: a GEMM with MT row tiles and NT column tiles has MT*NT output tiles; warp w owns tiles u = w + NWARPS i
// (m = u % MT, n = u / MT).  The host build (CPU tile emulator, tests only) replaces the fragment code by plain loops
// with the same thread -> element ownership.
// ------------------------------------------------------------------------------------------------
constexpr int NT_P = (P + 7) / 8;      // 13 point tiles of 8 (the last one is half empty)

OO_HOSTDEV inline constexpr int wfrag_units(int MT, int NT) { return (MT * NT + NWARPS - 1) / NWARPS; }
// which weight-gradient element register r of fragment i of thread tid holds (MT row tiles, NT column tiles)
OO_HOSTDEV inline int wfrag_row(int tid, int i, int r, int MT) { return 16 * (((tid >> 5) + NWARPS * i) % MT) + ((tid & 31) >> 2) + 8 * (r >> 1); }
OO_HOSTDEV inline int wfrag_col(int tid, int i, int r, int MT) { return 8 * (((tid >> 5) + NWARPS * i) / MT) + 2 * (tid & 3) + (r & 1); }

#ifdef __CUDACC__
struct FragA { uint32_t hi[4], lo[4]; };
struct FragB { uint32_t hi[2], lo[2]; };

OO_DEV void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
OO_DEV void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
OO_DEV void mma3(float (&c)[4], const FragA& a, const FragB& b) {
    mma_tf32(c, a.lo, b.hi);      // small terms first
    mma_tf32(c, a.hi, b.lo);
    mma_tf32(c, a.hi, b.hi);
}
OO_DEV void frag_a(FragA& f, float a0, float a1, float a2, float a3) {
    tf32_split(a0, f.hi[0], f.lo[0]); tf32_split(a1, f.hi[1], f.lo[1]);
    tf32_split(a2, f.hi[2], f.lo[2]); tf32_split(a3, f.hi[3], f.lo[3]);
}
OO_DEV void frag_b(FragB& f, float b0, float b1) {
    tf32_split(b0, f.hi[0], f.lo[0]); tf32_split(b1, f.hi[1], f.lo[1]);
}
#endif

// Y[j][p] = act(b[j] + sum_k W[j][k] X[k][p]), 16*MT output rows, P points, K % 8 == 0.  M = j, N = p.
// A warp owns NU or NU-1 output tiles; the body is instantiated for both counts so that it has no branches and the
// independent accumulator chains interleave.
#ifdef __CUDACC__
template <int K, int WS, int MT, bool RELU, int NU>
OO_DEV void gemm_fwd_body(int warp, int lane, const float* __restrict__ W, const float* __restrict__ bias,
                          const float* __restrict__ X, float* __restrict__ Y) {
    constexpr int NSTEP = NWARPS / MT;
    const int g = lane >> 2, t = lane & 3;
    const int m = warp % MT, n0 = warp / MT;
    // three accumulator chains per output tile (lo*hi, hi*lo, hi*hi): the MMAs of one k-step are independent
    float acc[NU][4], sm1[NU][4], sm2[NU][4];
    {
        const float b_lo = bias[16 * m + g], b_hi = bias[16 * m + g + 8];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            acc[i][0] = acc[i][1] = b_lo; acc[i][2] = acc[i][3] = b_hi;
            sm1[i][0] = sm1[i][1] = sm1[i][2] = sm1[i][3] = 0.f;
            sm2[i][0] = sm2[i][1] = sm2[i][2] = sm2[i][3] = 0.f;
        }
    }
    const float* wp = W + (16 * m + g) * WS + 2 * t;
    const float* xp = X + 2 * t * PS + 8 * n0 + g;
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 8) {
        const float2 w_lo = *reinterpret_cast<const float2*>(wp + k0);
        const float2 w_hi = *reinterpret_cast<const float2*>(wp + 8 * WS + k0);
        FragA a;
        frag_a(a, w_lo.x, w_hi.x, w_lo.y, w_hi.y);
        FragB b[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) frag_b(b[i], xp[k0 * PS + 8 * NSTEP * i], xp[(k0 + 1) * PS + 8 * NSTEP * i]);
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            mma_tf32(sm1[i], a.lo, b[i].hi);
            mma_tf32(sm2[i], a.hi, b[i].lo);
            mma_tf32(acc[i], a.hi, b[i].hi);
        }
    }
#pragma unroll
    for (int i = 0; i < NU; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[i][r] += sm1[i][r] + sm2[i][r];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const int p = 8 * (n0 + NSTEP * i) + 2 * t;
        if (p < P) {
            float2 y0 = {acc[i][0], acc[i][1]}, y1 = {acc[i][2], acc[i][3]};
            if (RELU) {
                y0.x = fmaxf(y0.x, 0.f); y0.y = fmaxf(y0.y, 0.f); y1.x = fmaxf(y1.x, 0.f); y1.y = fmaxf(y1.y, 0.f);
            }
            *reinterpret_cast<float2*>(Y + (16 * m + g) * PS + p) = y0;
            *reinterpret_cast<float2*>(Y + (16 * m + g + 8) * PS + p) = y1;
        }
    }
}
#endif

template <int K, int WS, int MT, bool RELU>
OO_DEV void gemm_fwd(int tid, const float* __restrict__ W, const float* __restrict__ bias,
                     const float* __restrict__ X, float* __restrict__ Y) {
#ifdef __CUDACC__
    static_assert(NWARPS % MT == 0 && K % 8 == 0 && WS % 2 == 0, "gemm_fwd tiling");
    constexpr int NU = wfrag_units(MT, NT_P), NSTEP = NWARPS / MT;
    const int warp = tid >> 5, lane = tid & 31;
    if (warp / MT + NSTEP * (NU - 1) < NT_P) gemm_fwd_body<K, WS, MT, RELU, NU>(warp, lane, W, bias, X, Y);
    else if (NU > 1) gemm_fwd_body<K, WS, MT, RELU, (NU > 1 ? NU - 1 : 1)>(warp, lane, W, bias, X, Y);
#else
    if (tid != 0) return;
    for (int j = 0; j < 16 * MT; ++j)
        for (int p = 0; p < P; ++p) {
            float acc = bias[j];
            for (int k = 0; k < K; ++k) acc += W[j * WS + k] * X[k * PS + p];
            Y[j * PS + p] = RELU ? (acc > 0.f ? acc : 0.f) : acc;
        }
#endif
}

// DX[k][p] = [k < relu_rows ? (X[k][p] > 0) : 1] * ( sum_{j<J0} W0[j][k] DY0[j][p] + sum_{j<J1} W1[j][k] DY1[j][p]
//            + [k < 32 && wa] wa[k] * draw[p] ),  written IN PLACE over X (rows 0..K-1).  M = k (MT = ceil(K/16) row
// tiles; rows >= K are computed from whatever follows the weights and never stored), N = p, contraction over j.
// An output tile reads and writes only its own 16 x 8 block of X, so in-place is safe across warps.
// Work split: warps 0..11 own one point tile each with all row tiles (each DY fragment is split once per k-step and
// reused for every row tile); the row tiles of the 13th point tile are shared out between warps 12..15, so every
// scheduler carries 3.25 point tiles.
#ifdef __CUDACC__
struct BwdDataArgs {
    const float *W0, *DY0, *W1, *DY1;
    float* X;
    int relu_rows;
    const float *wa, *draw;
};

// row tiles [mb, mb + MTN) of point tile n0
template <int K, int WS0, int J0, int WS1, int J1, int MTN>
OO_DEV void gemm_bwd_data_body(int lane, int mb, int n0, const BwdDataArgs& q) {
    constexpr int MG = MTN < 3 ? MTN : 3;            // row tiles whose weight fragments are live together
    const int g = lane >> 2, t = lane & 3;
    float acc[MTN][4], sml[MTN][4];          // hi*hi chain and the chain of the two small products
#pragma unroll
    for (int m = 0; m < MTN; ++m) {
        acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
        sml[m][0] = sml[m][1] = sml[m][2] = sml[m][3] = 0.f;
    }
#pragma unroll
    for (int term = 0; term < (J1 > 0 ? 2 : 1); ++term) {
        const int WS = term == 0 ? WS0 : WS1, J = term == 0 ? J0 : J1;
        const float* wp = (term == 0 ? q.W0 : q.W1) + t * WS + 16 * mb + g;
        const float* dp = (term == 0 ? q.DY0 : q.DY1) + t * PS + 8 * n0 + g;
#pragma unroll 2
        for (int j0 = 0; j0 < J; j0 += 8) {
            FragB b;
            frag_b(b, dp[j0 * PS], dp[(j0 + 4) * PS]);
#pragma unroll
            for (int m0 = 0; m0 < MTN; m0 += MG) {
                FragA a[MG];
#pragma unroll
                for (int m = 0; m < MG; ++m)
                    if (m0 + m < MTN)
                        frag_a(a[m], wp[j0 * WS + 16 * (m0 + m)], wp[j0 * WS + 16 * (m0 + m) + 8], wp[(j0 + 4) * WS + 16 * (m0 + m)],
                               wp[(j0 + 4) * WS + 16 * (m0 + m) + 8]);
#pragma unroll
                for (int m = 0; m < MG; ++m)
                    if (m0 + m < MTN) { mma_tf32(sml[m0 + m], a[m].lo, b.hi); mma_tf32(acc[m0 + m], a[m].hi, b.hi); }
#pragma unroll
                for (int m = 0; m < MG; ++m)
                    if (m0 + m < MTN) mma_tf32(sml[m0 + m], a[m].hi, b.lo);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MTN; ++m)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[m][r] += sml[m][r];
    const int p = 8 * n0 + 2 * t;
    if (p >= P) return;
    float2 dr = {0.f, 0.f};
    if (q.wa != nullptr) dr = *reinterpret_cast<const float2*>(q.draw + p);
#pragma unroll
    for (int m = 0; m < MTN; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 16 * (mb + m) + g + 8 * h;
            if (k < K) {
                float2 o = {acc[m][2 * h], acc[m][2 * h + 1]};
                if (q.wa != nullptr && k < H) {
                    o.x += q.wa[k] * dr.x; o.y += q.wa[k] * dr.y;
                }
                float2* xp = reinterpret_cast<float2*>(q.X + k * PS + p);
                if (k < q.relu_rows) {
                    const float2 hv = *xp;
                    o.x = hv.x > 0.f ? o.x : 0.f; o.y = hv.y > 0.f ? o.y : 0.f;
                }
                *xp = o;
            }
        }
}
#endif

template <int K, int WS0, int J0, int WS1, int J1>
OO_DEV void gemm_bwd_data(int tid, const float* __restrict__ W0, const float* __restrict__ DY0,
                          const float* __restrict__ W1, const float* __restrict__ DY1,
                          float* X, int relu_rows, const float* wa, const float* draw) {
#ifdef __CUDACC__
    static_assert(J0 % 8 == 0 && J1 % 8 == 0 && NWARPS == 16 && NT_P == 13, "gemm_bwd_data tiling");
    constexpr int MT = (K + 15) / 16, MSPLIT = (MT + 3) / 4, MREM = MT % MSPLIT ? MT % MSPLIT : MSPLIT;
    const int warp = tid >> 5, lane = tid & 31;
    const BwdDataArgs q = {W0, DY0, W1, DY1, X, relu_rows, wa, draw};
    if (warp < NT_P - 1) {
        gemm_bwd_data_body<K, WS0, J0, WS1, J1, MT>(lane, 0, warp, q);
    } else {                                   // a quarter of the row tiles of the last point tile
        const int mb = (warp - (NT_P - 1)) * MSPLIT;
        if (mb + MSPLIT <= MT) gemm_bwd_data_body<K, WS0, J0, WS1, J1, MSPLIT>(lane, mb, NT_P - 1, q);
        else if (mb < MT) gemm_bwd_data_body<K, WS0, J0, WS1, J1, MREM>(lane, mb, NT_P - 1, q);
    }
#else
    if (tid != 0) return;
    for (int k = 0; k < K; ++k)
        for (int p = 0; p < P; ++p) {
            float acc = 0.f;
            for (int j = 0; j < J0; ++j) acc += W0[j * WS0 + k] * DY0[j * PS + p];
            for (int j = 0; j < J1; ++j) acc += W1[j * WS1 + k] * DY1[j * PS + p];
            if (wa != nullptr && k < H) acc += wa[k] * draw[p];
            if (k < relu_rows && !(X[k * PS + p] > 0.f)) acc = 0.f;
            X[k * PS + p] = acc;
        }
#endif
}

// acc fragment i (+)= sum_p DY[j][p] X[k][p] over the tile's points, for this thread's elements of output tiles
// u = warp + NWARPS i of the [16 MT] x [8 NT] weight gradient.  M = j, N = k, contraction over points (the four points
// of the last k-step that lie beyond P contribute zero).
#ifdef __CUDACC__
template <int MT, int NT, int NU>
OO_DEV void gemm_bwd_w_body(float* acc, int warp, int lane, const float* __restrict__ DY, const float* __restrict__ X,
                            int ks0 = 0, int ks1 = NT_P) {
    constexpr int NSTEP = NWARPS / MT;
    const int g = lane >> 2, t = lane & 3;
    const int m = warp % MT, n0 = warp / MT;
    // three accumulator chains per output tile (lo*hi, hi*lo, hi*hi): the MMAs of one k-step are independent
    float c[3][NU][4];
#pragma unroll
    for (int i = 0; i < NU; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) { c[0][i][r] = acc[4 * i + r]; c[1][i][r] = 0.f; c[2][i][r] = 0.f; }
    const float* dp = DY + (16 * m + g) * PS + t;
    const float* xp = X + (8 * n0 + g) * PS + t;
#pragma unroll 2
    for (int ks = ks0; ks < ks1; ++ks) {
        const int p0 = 8 * ks;
        const bool tail = p0 + t + 4 >= P;               // only in the last k-step
        FragA a;
        frag_a(a, dp[p0], dp[8 * PS + p0], tail ? 0.f : dp[p0 + 4], tail ? 0.f : dp[8 * PS + p0 + 4]);
        FragB b[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) frag_b(b[i], xp[8 * NSTEP * i * PS + p0], tail ? 0.f : xp[8 * NSTEP * i * PS + p0 + 4]);
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            mma_tf32(c[1][i], a.lo, b[i].hi);
            mma_tf32(c[2][i], a.hi, b[i].lo);
            mma_tf32(c[0][i], a.hi, b[i].hi);
        }
    }
#pragma unroll
    for (int i = 0; i < NU; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[4 * i + r] = c[0][i][r] + (c[1][i][r] + c[2][i][r]);
}
#endif

template <int MT, int NT>
OO_DEV void gemm_bwd_w(float* acc, int tid, const float* __restrict__ DY, const float* __restrict__ X) {
#ifdef __CUDACC__
    static_assert(NWARPS % MT == 0, "gemm_bwd_w tiling");
    constexpr int NU = wfrag_units(MT, NT), NSTEP = NWARPS / MT;
    const int warp = tid >> 5, lane = tid & 31;
    if (warp / MT + NSTEP * (NU - 1) < NT) gemm_bwd_w_body<MT, NT, NU>(acc, warp, lane, DY, X);
    else if (NU > 1 && warp / MT < NT) gemm_bwd_w_body<MT, NT, (NU > 1 ? NU - 1 : 1)>(acc, warp, lane, DY, X);
#else
    for (int i = 0; i < wfrag_units(MT, NT); ++i)
        for (int r = 0; r < 4; ++r) {
            const int j = wfrag_row(tid, i, r, MT), k = wfrag_col(tid, i, r, MT);
            if (k >= 8 * NT) continue;
            float s = 0.f;
            for (int p = 0; p < P; ++p) s += DY[j * PS + p] * X[k * PS + p];
            acc[4 * i + r] += s;
        }
#endif
}

// The two 32 x 32 weight gradients have only 8 output tiles: two warps share a tile, each contracting over half of the
// point k-steps (warp w and w + 8 own tile w; the halves are added when the slot is flushed).
constexpr int KS_SPLIT = 7;            // k-steps [0, 7) and [7, 13)
OO_HOSTDEV inline int wfrag32_row(int tid, int r) { return 16 * (((tid >> 5) & 7) % 2) + ((tid & 31) >> 2) + 8 * (r >> 1); }
OO_HOSTDEV inline int wfrag32_col(int tid, int r) { return 8 * (((tid >> 5) & 7) / 2) + 2 * (tid & 3) + (r & 1); }

OO_DEV void gemm_bwd_w32(float* acc, int tid, const float* __restrict__ DY, const float* __restrict__ X) {
    static_assert(NWARPS == 16 && H == 32, "gemm_bwd_w32 pairs warps w and w + 8");
    const int half = tid >> 8;
#ifdef __CUDACC__
    gemm_bwd_w_body<2, H / 8, 1>(acc, (tid >> 5) & 7, tid & 31, DY, X, half ? KS_SPLIT : 0, half ? NT_P : KS_SPLIT);
#else
    for (int r = 0; r < 4; ++r) {
        const int j = wfrag32_row(tid, r), k = wfrag32_col(tid, r);
        float s = 0.f;
        for (int p = half ? 8 * KS_SPLIT : 0; p < (half ? P : 8 * KS_SPLIT); ++p) s += DY[j * PS + p] * X[k * PS + p];
        acc[r] += s;
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// staging an object's weights into shared memory (padded rows, zero pad columns)
// ------------------------------------------------------------------------------------------------
// every element travels as a 4-byte asynchronous copy (the rows of the reference's tensors are not 16-byte aligned): all
// of a thread's copies are in flight together instead of one L2 round trip per loop iteration; awaited in stage_weights
OO_DEV void copy_rows(int tid, float* dst, int ws, const float* __restrict__ src, int rows, int cols, int colsp) {
    for (int i = tid; i < rows * colsp; i += NTHREADS) {
        const int r = i / colsp, c = i - r * colsp;
        if (c < cols) OO_CP_ASYNC4(dst + r * ws + c, src + r * cols + c);
        else dst[r * ws + c] = 0.f;
    }
}

OO_DEV void stage_weights(int tid, float* sm, const float* __restrict__ th) {
    float* w = sm + SM_W;
    copy_rows(tid, w + W_IN, WS_IN, th + OFF_IN_W, H, E1, KP_IN);
    copy_rows(tid, w + W_M1, WS_H, th + OFF_M1_W, H, H, H);
    copy_rows(tid, w + W_CAT, WS_CAT, th + OFF_CAT_W, H, H + E1, KP_CAT);
    copy_rows(tid, w + W_M2, WS_H, th + OFF_M2_W, H, H, H);
    copy_rows(tid, w + W_CL, WS_HD, th + OFF_CL_W, H, H + E2, KP_HD);
    copy_rows(tid, w + W_CP, WS_HD, th + OFF_CP_W, H, H + E2, KP_HD);
    for (int i = tid; i < H; i += NTHREADS) {
        OO_CP_ASYNC4(w + W_A + i, th + OFF_A_W + i);
        OO_CP_ASYNC4(w + B_IN + i, th + OFF_IN_B + i);
        OO_CP_ASYNC4(w + B_M1 + i, th + OFF_M1_B + i);
        OO_CP_ASYNC4(w + B_CAT + i, th + OFF_CAT_B + i);
        OO_CP_ASYNC4(w + B_M2 + i, th + OFF_M2_B + i);
        OO_CP_ASYNC4(w + B_CL + i, th + OFF_CL_B + i);
        OO_CP_ASYNC4(w + B_CP + i, th + OFF_CP_B + i);
    }
    for (int i = tid; i < 3 * H; i += NTHREADS) OO_CP_ASYNC4(w + W_OC + i, th + OFF_OC_W + i);
    for (int i = tid; i < 64; i += NTHREADS) {
        if (i < NDIR * 3) OO_CP_ASYNC4(w + W_PE + i, th + OFF_PE_B + i);
        else w[W_PE + i] = 0.f;
    }
    if (tid < 4) {
        w[B_A + tid] = tid < 1 ? OO_LDG(th + OFF_A_B) : 0.f;
        w[B_OC + tid] = tid < 3 ? OO_LDG(th + OFF_OC_B + tid) : 0.f;
    }
    OO_CP_ASYNC_WAIT();        // the caller's block barrier publishes the copies
}

// Per-object constants of the out_clip layer, computed once per object and step by k_gram (one CTA per object, before
// k_train): [W | b] (512 x 33, rows padded to 36 floats) is staged in shared memory with 16-byte async copies, then
// G' = [W | b]^T [W | b] (33 x 33; G = W^T W, wb = W^T b, bb = b.b) is formed with 4x4 register tiles, the 512 rows split
// over NGG thread groups whose partials are summed through shared memory.  Result -> the object's row of `derived` (global).
constexpr int GS = 36;                       // row stride of the staged [W | b]
constexpr int SM_GPART = 512 * GS;           // [NGG groups][36 x 36] partial products (floats, inside the activation area)
constexpr int NGG = NTHREADS / 64;           // thread groups of 64, each reduces C / NGG rows of [W | b]
static_assert(SM_GPART + NGG * 36 * 36 <= SM_W, "gram staging must fit in the activation area");

template <int STEP>
OO_DEV void gram_stage(int tid, float* __restrict__ sm, const float* __restrict__ th, float* __restrict__ der) {
    float* wst = sm + SM_ACT;
    float* part = sm + SM_ACT + SM_GPART;
    if constexpr (STEP == -1) {
        // issue only: the 64 KB fetch overlaps the (synchronous) staging of the other weights
        for (int q = tid; q < C * (H / 4); q += NTHREADS) {                       // 4096 x 16 B
            const int row = q >> 3, v4 = q & 7;
            OO_CP_ASYNC16(wst + row * GS + 4 * v4, th + OFF_OCL_W + row * H + 4 * v4);
        }
    } else if constexpr (STEP == 0) {
        for (int row = tid; row < C; row += NTHREADS) {
            wst[row * GS + H] = OO_LDG(th + OFF_OCL_B + row);
            wst[row * GS + H + 1] = wst[row * GS + H + 2] = wst[row * GS + H + 3] = 0.f;
        }
        OO_CP_ASYNC_WAIT();
    } else if constexpr (STEP == 1) {
        // thread (group g = tid>>6, tile t = tid&63): 9 x 9 tiles of 4x4 cover 36 x 36; tiles 64..80 are taken by the
        // first 17 threads of each group in a second pass
        const int g = tid >> 6;
        for (int t = tid & 63; t < 81; t += 64) {
            const int k4 = 4 * (t / 9), j4 = 4 * (t % 9);
            float acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x) acc[x][0] = acc[x][1] = acc[x][2] = acc[x][3] = 0.f;
            const float* base = wst + (size_t)((C / NGG) * g) * GS;
#pragma unroll 4
            for (int cc = 0; cc < C / NGG; ++cc) {
                const float4 a4 = ld4(base + cc * GS + k4), b4 = ld4(base + cc * GS + j4);
                acc[0][0] += a4.x * b4.x; acc[0][1] += a4.x * b4.y; acc[0][2] += a4.x * b4.z; acc[0][3] += a4.x * b4.w;
                acc[1][0] += a4.y * b4.x; acc[1][1] += a4.y * b4.y; acc[1][2] += a4.y * b4.z; acc[1][3] += a4.y * b4.w;
                acc[2][0] += a4.z * b4.x; acc[2][1] += a4.z * b4.y; acc[2][2] += a4.z * b4.z; acc[2][3] += a4.z * b4.w;
                acc[3][0] += a4.w * b4.x; acc[3][1] += a4.w * b4.y; acc[3][2] += a4.w * b4.z; acc[3][3] += a4.w * b4.w;
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
                st4(part + g * 1296 + (k4 + x) * 36 + j4, float4{acc[x][0], acc[x][1], acc[x][2], acc[x][3]});
        }
    } else {
        for (int q = tid; q < 33 * 33; q += NTHREADS) {
            const int k = q / 33, j = q - 33 * k;
            float v = 0.f;
#pragma unroll
            for (int gg = 0; gg < NGG; gg += 2) v += part[gg * 1296 + k * 36 + j] + part[(gg + 1) * 1296 + k * 36 + j];
            if (k < H && j < H) der[DER_G + k * H + j] = v;
            else if (k < H) der[DER_WB + k] = v;            // column 32: W^T b
            else if (j == H) der[DER_BB] = v;               // b . b
        }
    }
}

// W_ocl^T b_ocl and b.b of the object a CTA starts working on: global (k_gram) -> the per-object rows of the ray-value area
OO_DEV void stage_derived(int tid, float* __restrict__ sm, const float* __restrict__ der) {
    if (tid <= H) sm[SM_RV + V_WB * RP + tid] = OO_LDG(der + DER_WB + tid);      // DER_BB == DER_WB + H
    if (tid < H * H / 4) OO_CP_ASYNC16(sm + SM_G + 4 * tid, der + DER_G + 4 * tid);
    OO_CP_ASYNC_WAIT();        // the caller's block barrier publishes the copies
}

// zero the pad rows of e1 / e2 and the spare rows once per kernel
OO_DEV void zero_pad_rows(int tid, float* sm) {
    for (int i = tid; i < PS; i += NTHREADS) {
        sm[(R_E1 + E1) * PS + i] = 0.f;
#pragma unroll
        for (int q = E2; q < E2P; ++q) sm[(R_E2 + q) * PS + i] = 0.f;
        sm[(R_T + 3) * PS + i] = 0.f;
        sm[(R_MISC + M_HU) * PS + i] = 0.f;
        sm[(R_MISC + M_HU + 1) * PS + i] = 0.f;
    }
    for (int i = tid; i < 2 * H * RP + NRV_TILE * RP; i += NTHREADS) sm[SM_ST + i] = 0.f;
}

#ifdef __CUDACC__
// Phase 0 of a training tile with its global loads taken out: while a tile is being processed every thread holds ONE value
// of the next tile in a register (issued before the tile starts, so the HBM round trip hides behind ~100k cycles of work):
//   tid < 300: pcs, 300..399: z, 400..449: gt depth / label / r / g / b of the rays, 450..459: feature-table rows.
static_assert(4 * P + 6 * RT <= NTHREADS, "one prefetched value per thread");
template <bool PART>
OO_DEV float tile_prefetch(int tid, const TileCtx& c) {
    float v = 0.f;
    if (tid < 3 * P) {
        if (tid < c.npts * 3) v = OO_LDG(c.pcs + tid);
    } else if (tid < 4 * P) {
        if (tid - 3 * P < c.npts) v = OO_LDG(c.z + tid - 3 * P);
    } else if (tid < 4 * P + 5 * RT) {
        const int q = (tid - 4 * P) / RT, r = (tid - 4 * P) - q * RT;
        if (r < c.nrays) v = q == 0 ? OO_LDG(c.gt_depth + r) : q == 1 ? (float)c.labels[r] : (float)c.gt_rgb[3 * r + q - 2];
    } else if (PART && tid < 4 * P + 6 * RT) {
        const int r = tid - (4 * P + 5 * RT);
        if (r < c.nrays) v = __int_as_float(OO_LDG(c.feat_row + r));
    }
    return v;
}

template <bool PART>
OO_DEV void tile_phase0_pre(int tid, float* __restrict__ sm, const TileCtx& c, float pre) {
    float* act = sm + SM_ACT;
    if (tid < 3 * P) {
        const int p = tid / 3, ch = tid - 3 * p;
        const float t = pre / c.scale;
        act[(R_T + ch) * PS + p] = t;
        act[(R_E1 + ch) * PS + p] = t;
    } else if (tid < 4 * P) {
        act[(R_MISC + M_Z) * PS + tid - 3 * P] = pre;
    } else if (tid < 4 * P + 5 * RT) {
        const int q = (tid - 4 * P) / RT, r = (tid - 4 * P) - q * RT;
        sm[SM_RV + (V_GTD + q) * RP + r] = pre;
    } else if (PART && tid < 4 * P + 6 * RT) {
        reinterpret_cast<int*>(sm + SM_FROW)[tid - (4 * P + 5 * RT)] = __float_as_int(pre);
    }
}
#endif

// ------------------------------------------------------------------------------------------------
// phases
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// device: the eleven small phases between the forward heads and the backward GEMMs are fused into three (40, 41, 42)
constexpr int N_TRAIN_PHASES = 22;
#else
constexpr int N_TRAIN_PHASES = 30;
#endif
constexpr int N_FWD_PHASES = 8;    // phases 0..7 are shared with the standalone forward kernel
// execution order of the training tile (phases 32..35 were split out of their neighbours later)
#ifdef __CUDACC__
constexpr int kTrainOrder[N_TRAIN_PHASES] = {0, 1, 2, 3, 4, 5, 6, 40, 41, 42, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31};
#else
// host (CPU tile emulator): the unfused formulation, one small phase per step, every thread private within a phase
constexpr int kTrainOrder[N_TRAIN_PHASES] = {0, 1, 2, 3, 4, 5, 6, 7, 33, 8, 10, 11, 32, 13, 16, 17, 18, 19,
                                             20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31};
#endif

template <int PH, bool PART>
OO_DEV void tile_phase(int tid, float* __restrict__ sm, const TileCtx& c, TileAcc& a) {
    float* act = sm + SM_ACT;
    float* w = sm + SM_W;
    float* misc = act + R_MISC * PS;
    float* rv = sm + SM_RV;

    if constexpr (PH == 0) {
        // scaled coordinates t = x / scale (embedding.py:47) -> T rows and e1[0:3]
        const int n = c.npts * 3;
        for (int i = tid; i < P * 3; i += NTHREADS) {
            const float x = i < n ? OO_LDG(c.pcs + i) : 0.f;
            const int p = i / 3, ch = i - 3 * p;
            const float t = x / c.scale;
            act[(R_T + ch) * PS + p] = t;
            act[(R_E1 + ch) * PS + p] = t;
        }
        if (c.nrays > 0) {
            // training tile: z of the points and the gt values of the rays, so that no later phase waits on HBM
            for (int p = tid; p < P; p += NTHREADS) misc[M_Z * PS + p] = p < c.npts ? OO_LDG(c.z + p) : 0.f;
            if (tid >= 128 && tid < 128 + 5 * RT) {
                const int q = (tid - 128) / RT, r = (tid - 128) - q * RT;
                float v = 0.f;
                if (r < c.nrays) v = q == 0 ? OO_LDG(c.gt_depth + r) : q == 1 ? (float)c.labels[r] : (float)c.gt_rgb[3 * r + q - 2];
                rv[(V_GTD + q) * RP + r] = v;
            }
        }
        if (PART && tid >= 256 && tid < 256 + RT)         // feature-table rows of the tile's rays (consumed by phase 1)
            reinterpret_cast<int*>(sm + SM_FROW)[tid - 256] = tid - 256 < c.nrays ? OO_LDG(c.feat_row + tid - 256) : 0;
    } else if constexpr (PH == 1) {
        if (PART) {
            // gt part features of the tile's rays -> Y [10][512]; asynchronous, awaited at the end of phase 8
            const int* frow = reinterpret_cast<const int*>(sm + SM_FROW);
            for (int i = tid; i < RT * (C / 4); i += NTHREADS) {
                const int r = i / (C / 4), q = i - r * (C / 4);
                float* dst = sm + SM_FEAT + r * YSTR + 4 * q;
                if (r < c.nrays) {
                    OO_CP_ASYNC16(dst, c.feat_table + (size_t)frow[r] * C + 4 * q);
                } else {
                    dst[0] = dst[1] = dst[2] = dst[3] = 0.f;
                }
            }
        }
        // proj = B t ; e[3 + 21 k + d] = sin(pi * 2^k * proj) (embedding.py:48-53).  The reference's argument for band k
        // is fl(2^k proj * pi_f) = 2^k * fl(proj * pi_f) exactly (power-of-two scaling commutes with rounding), so all
        // six bands follow from one sincosf by angle doubling: s' = 2 s c, c' = (c - s)(c + s)  (norm error only doubles).
        // the band-0 (sin, cos) of this thread's items stay in its tensor-memory lane for the backward encoder (phase 30 walks
        // the same items with the same thread), which then needs neither the projection nor a second sincosf
        constexpr int NTRIG = (NDIR * P + NTHREADS - 1) / NTHREADS;
        static_assert(NTRIG == 5, "the sin/cos stash holds five items per thread");
        float trig[2 * NTRIG];
#pragma unroll
        for (int u = 0; u < NTRIG; ++u) {
            const int i = tid + NTHREADS * u;
            trig[u] = trig[NTRIG + u] = 0.f;
            if (i < NDIR * P) {
                const int d = i / P, p = i - d * P;
                const float t0 = act[(R_T + 0) * PS + p], t1 = act[(R_T + 1) * PS + p], t2 = act[(R_T + 2) * PS + p];
                const float proj = w[W_PE + 3 * d] * t0 + w[W_PE + 3 * d + 1] * t1 + w[W_PE + 3 * d + 2] * t2;
                float sn, cs;
                sincosf(proj * PI_F, &sn, &cs);
                trig[u] = sn;
                trig[NTRIG + u] = cs;
#pragma unroll
                for (int k = 0; k < NBAND; ++k) {
                    const int row = 3 + NDIR * k + d;
                    if (row < E1) act[(R_E1 + row) * PS + p] = sn;
                    else act[(R_E2 + row - E1) * PS + p] = sn;
                    const float s2 = 2.f * sn * cs, c2 = (cs - sn) * (cs + sn);
                    sn = s2;
                    cs = c2;
                }
            }
        }
#ifdef __CUDACC__
        if (c.nrays > 0) {                       // training tiles only (the forward kernels allocate no tensor memory)
            tm_st<8>(a.tm + AC_TRIG, trig);
            tm_st<2>(a.tm + AC_TRIG + 8, trig + 8);
        }
#endif
    } else if constexpr (PH == 2) {
        gemm_fwd<KP_IN, WS_IN, 2, true>(tid, w + W_IN, w + B_IN, act + R_E1 * PS, act + R_H1 * PS);
    } else if constexpr (PH == 3) {
        gemm_fwd<H, WS_H, 2, true>(tid, w + W_M1, w + B_M1, act + R_H1 * PS, act + R_H2 * PS);
    } else if constexpr (PH == 4) {
        gemm_fwd<KP_CAT, WS_CAT, 2, true>(tid, w + W_CAT, w + B_CAT, act + R_H2 * PS, act + R_H3 * PS);
    } else if constexpr (PH == 5) {
        gemm_fwd<H, WS_H, 2, true>(tid, w + W_M2, w + B_M2, act + R_H3 * PS, act + R_H4 * PS);
    } else if constexpr (PH == 6) {
        // [color_linear ; clip_linear] share their input: one GEMM with 64 output rows when the clip head is live
        if (PART) gemm_fwd<KP_HD, WS_HD, 4, true>(tid, w + W_CL, w + B_CL, act + R_H4 * PS, act + R_HC * PS);
        else gemm_fwd<KP_HD, WS_HD, 2, true>(tid, w + W_CL, w + B_CL, act + R_H4 * PS, act + R_HC * PS);
#ifdef __CUDACC__
        if (PART) OO_CP_ASYNC_WAIT();      // Y rows issued in phase 1 are complete for this thread; the barrier publishes them
#endif
    } else if constexpr (PH == 7) {
        // out_alpha (x10, model.py:88) -> occupancy = sigmoid (render_rays.py:13); out_color -> sigmoid (model.py:96)
        for (int i = tid; i < 4 * P; i += NTHREADS) {
            const int o = i / P, p = i - o * P;
            if (o == 0) {
                float r = w[B_A];
#pragma unroll 8
                for (int j = 0; j < H; ++j) r += w[W_A + j] * act[(R_H4 + j) * PS + p];
                misc[M_DRAW * PS + p] = r * 10.f;                  // alpha (kept for the forward-only kernel)
                misc[M_OCC * PS + p] = sigmoidf_(r * 10.f);
            } else {
                const int ch = o - 1;
                float r = w[B_OC + ch];
#pragma unroll 8
                for (int j = 0; j < H; ++j) r += w[W_OC + ch * H + j] * act[(R_HC + j) * PS + p];
                misc[(M_COL + ch) * PS + p] = sigmoidf_(r);
            }
        }
#ifdef __CUDACC__
    } else if constexpr (PH == 40) {
        // ---- fused: out_alpha / out_color heads (phase 7) + v = W^T y partials (tensor part of phase 10).  The out_clip
        // fragments come from L2: they are requested first and the head outputs are computed while they fly.
        constexpr int CW = C / NWARPS;
        static_assert(CW == 32, "each warp takes four 8-wide k-steps of gt-feature columns");
        const int wv = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
        float bw[4][4][2], bq[4][2];
        if (PART) {
            const float* wg = c.theta + OFF_OCL_W + (size_t)(wv * CW) * H;
            const float* bg = c.theta + OFF_OCL_B + wv * CW;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    bw[ks][nt][0] = OO_LDG(wg + (8 * ks + t) * H + 8 * nt + g);
                    bw[ks][nt][1] = OO_LDG(wg + (8 * ks + t + 4) * H + 8 * nt + g);
                }
                bq[ks][0] = g == 0 ? OO_LDG(bg + 8 * ks + t) : 0.f;
                bq[ks][1] = g == 0 ? OO_LDG(bg + 8 * ks + t + 4) : 0.f;
            }
        }
        if (tid < 4 * P) {
            const int o = tid / P, p = tid - o * P;
            if (o == 0) {
                float r = w[B_A];
#pragma unroll 8
                for (int j = 0; j < H; ++j) r += w[W_A + j] * act[(R_H4 + j) * PS + p];
                misc[M_OCC * PS + p] = sigmoidf_(r * 10.f);
            } else {
                const int ch = o - 1;
                float r = w[B_OC + ch];
#pragma unroll 8
                for (int j = 0; j < H; ++j) r += w[W_OC + ch * H + j] * act[(R_HC + j) * PS + p];
                misc[(M_COL + ch) * PS + p] = sigmoidf_(r);
            }
        }
        if (PART) {
            float* y = sm + SM_FEAT + wv * CW;
            const bool hi_row = g + 8 < RT;
            float vacc[5][4];
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) vacc[nt][0] = vacc[nt][1] = vacc[nt][2] = vacc[nt][3] = 0.f;
            float yy0 = 0.f, yy1 = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const float a0 = y[g * YSTR + 8 * ks + t], a2 = y[g * YSTR + 8 * ks + t + 4];
                const float a1 = hi_row ? y[(g + 8) * YSTR + 8 * ks + t] : 0.f, a3 = hi_row ? y[(g + 8) * YSTR + 8 * ks + t + 4] : 0.f;
                yy0 += a0 * a0 + a2 * a2;
                yy1 += a1 * a1 + a3 * a3;
                FragA fa;
                frag_a(fa, a0, a1, a2, a3);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    FragB fb;
                    frag_b(fb, bw[ks][nt][0], bw[ks][nt][1]);
                    mma3(vacc[nt], fa, fb);
                }
                FragB fq;
                frag_b(fq, bq[ks][0], bq[ks][1]);
                mma3(vacc[4], fa, fq);
            }
            yy0 += __shfl_xor_sync(0xffffffffu, yy0, 1); yy0 += __shfl_xor_sync(0xffffffffu, yy0, 2);
            yy1 += __shfl_xor_sync(0xffffffffu, yy1, 1); yy1 += __shfl_xor_sync(0xffffffffu, yy1, 2);
            if (t == 0) {
                sm[SM_YS + (g * 16 + wv) * 2] = vacc[4][0];
                sm[SM_YS + (g * 16 + wv) * 2 + 1] = yy0;
                if (hi_row) {
                    sm[SM_YS + ((g + 8) * 16 + wv) * 2] = vacc[4][2];
                    sm[SM_YS + ((g + 8) * 16 + wv) * 2 + 1] = yy1;
                }
            }
            __syncwarp();
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                *reinterpret_cast<float2*>(y + g * YSTR + 8 * nt + 2 * t) = float2{vacc[nt][0], vacc[nt][1]};
                if (hi_row) *reinterpret_cast<float2*>(y + (g + 8) * YSTR + 8 * nt + 2 * t) = float2{vacc[nt][2], vacc[nt][3]};
            }
        }
    } else if constexpr (PH == 41) {
        // ---- fused per-ray chain: warp r owns ray r from the termination weights to dL/d(alpha, colour) of its samples
        // (phases 33, 8, 10 (S), 11, 32, 13 (U, record), 16, 17 of the unfused formulation; same operation order).  Lanes are
        // samples for the compositing parts and hidden units for the feature parts; everything stays in registers / shuffles,
        // shared memory only receives what later phases read.
        const int r = tid >> 5, lane = tid & 31;
        if (r < RT) {
            constexpr unsigned FULL = 0xffffffffu;
            const bool live_ray = r < c.nrays, smp = lane < S, live = live_ray && smp;
            const int p = r * S + (smp ? lane : 0);
            // termination: exclusive product of the free probabilities in torch.cumprod's order (render_rays.py:36-43)
            const float occ = smp ? misc[M_OCC * PS + p] : 0.f;
            const float fq = 1.f - occ + 1e-10f;
            float freep = 1.f;
#pragma unroll
            for (int q = 0; q < S - 1; ++q) {
                const float f = __shfl_sync(FULL, fq, q);
                if (q < lane) freep *= f;
            }
            const float fp = live ? freep : 0.f;
            const float t = live ? occ * freep : 0.f;
            if (smp) misc[M_TERM * PS + p] = t;
            // rendered depth / variance / colour / opacity (render_rays.py:56-63)
            const float zv = live ? misc[M_Z * PS + p] : 0.f;
            const float k0 = misc[(M_COL + 0) * PS + p], k1 = misc[(M_COL + 1) * PS + p], k2 = misc[(M_COL + 2) * PS + p];
            const float depth = warp_sum(t * zv), opac = warp_sum(t);
            const float c0 = warp_sum(t * k0), c1 = warp_sum(t * k1), c2 = warp_sum(t * k2);
            const float dz = zv - depth;
            const float var = warp_sum(t * (dz * dz));
            float gd = 0.f, go = 0.f, gc0 = 0.f, gc1 = 0.f, gc2 = 0.f, cf = 0.f;
            float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
            if (lane == 0 && live_ray) {                  // loss terms and their derivatives (loss.py:27-75)
                const int lab = (int)rv[V_LAB * RP + r];
                const bool is1 = lab == 1, sem = lab != 2;
                const float tgt = lab != 0 ? 1.f : 0.f;
                if (is1 && !(c.flags & 2)) {
                    const float wgt = 1.f / (sqrtf(var) + 1e-4f);           // render_rays.py:95-100, var detached
                    const float dd = depth - rv[V_GTD * RP + r];
                    l0 = fabsf(dd) * wgt;
                    gd = sgnf_(dd) * wgt * c.inv1;
                    const float e0 = c0 - rv[(V_RGB + 0) * RP + r] / 255.f;   // train.py:373 `/ 255.`
                    const float e1 = c1 - rv[(V_RGB + 1) * RP + r] / 255.f;
                    const float e2 = c2 - rv[(V_RGB + 2) * RP + r] / 255.f;
                    l1 = fabsf(e0) + fabsf(e1) + fabsf(e2);                      // loss.py:61 sum over channels
                    gc0 = sgnf_(e0) * c.cs * c.inv1;
                    gc1 = sgnf_(e1) * c.cs * c.inv1;
                    gc2 = sgnf_(e2) * c.cs * c.inv1;
                    cf = PART ? c.fs * c.inv1 : 0.f;
                }
                if (sem && !(c.flags & 4)) {
                    const float eo = opac - tgt;                                 // loss.py:71
                    l2 = fabsf(eo);
                    go = sgnf_(eo) * c.os * c.invs;
                }
            }
            gd = __shfl_sync(FULL, gd, 0); go = __shfl_sync(FULL, go, 0);
            gc0 = __shfl_sync(FULL, gc0, 0); gc1 = __shfl_sync(FULL, gc1, 0); gc2 = __shfl_sync(FULL, gc2, 0);
            float bgv = 0.f;
            if (PART) {
                // S_j = sum_i T_i hp_i[j]; v_j = sum of the 16 partial W^T y; gs_j = (G S)_j + opac wb_j
                float sj = 0.f;
#pragma unroll
                for (int q = 0; q < S; ++q) sj += __shfl_sync(FULL, t, q) * act[(R_HP + lane) * PS + r * S + q];
                sm[SM_ST + lane * RP + r] = sj;
                float vj = 0.f;
#pragma unroll
                for (int wv = 0; wv < NWARPS; ++wv) vj += sm[SM_FEAT + r * YSTR + wv * (C / NWARPS) + lane];
                float tot = 0.f;
                if (lane < 2) {                               // totals of y.b (lane 0) and y.y (lane 1)
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int ch = 0; ch < 16; ch += 2) {
                        s0 += sm[SM_YS + (r * 16 + ch) * 2 + lane];
                        s1 += sm[SM_YS + (r * 16 + ch + 1) * 2 + lane];
                    }
                    tot = s0 + s1;
                }
                const float yb = __shfl_sync(FULL, tot, 0), yy = __shfl_sync(FULL, tot, 1);
                const float wbj = rv[V_WB * RP + lane];
                float gk[H];
#pragma unroll
                for (int k = 0; k < H; ++k) gk[k] = sm[SM_G + k * H + lane];             // G is symmetric: conflict-free over j
                float gs = opac * wbj;
#pragma unroll
                for (int k = 0; k < H; ++k) gs += gk[k] * __shfl_sync(FULL, sj, k);
                const float sv = warp_sum(sj * vj), swb = warp_sum(sj * wbj), sgs = warp_sum(sj * gs);
                float A = 0.f, B = 0.f;
                if (lane == 0) {
                    const float bb = rv[V_WB * RP + H];
                    const float xb = swb + opac * bb;                     // b . x
                    const float xy = sv + opac * yb, xx = sgs + opac * xb;
                    const float nxr = sqrtf(fmaxf(xx, 0.f)), nyr = sqrtf(yy);
                    const float nx = fmaxf(nxr, 1e-8f), ny = fmaxf(nyr, 1e-8f);   // F.cosine_similarity eps clamp
                    const float cosv = xy / (nx * ny);
                    if (cf != 0.f) {
                        l3 = 1.f - cosv;
                        A = -cf / (nx * ny);                              // d(1-cos)/dx = -y/(nx ny) + [|x|>eps] cos x / nx^2
                        B = nxr > 1e-8f ? cf * cosv / (nx * nx) : 0.f;
                    }
                    bgv = A * yb + B * xb;                                // b_ocl . dL/dx
                    rv[V_B * RP + r] = B;
                }
                A = __shfl_sync(FULL, A, 0); B = __shfl_sync(FULL, B, 0); bgv = __shfl_sync(FULL, bgv, 0);
                sm[SM_UT + lane * RP + r] = A * vj + B * gs;             // dL/dS
                if (live_ray) {                                           // per-ray record for K4 (out_clip gradient)
                    c.rayrec[r * RAYREC + REC_S + lane] = sj;
                    if (lane < 2) c.rayrec[r * RAYREC + lane] = lane == REC_A ? A : opac;
                }
                __syncwarp();
            }
            if (lane == 0) rv[V_OPAC * RP + r] = opac;
            // dL/dT of each sample, then back through the compositing sums and the exclusive product
            float g = 0.f;
            if (live) {
                g = zv * gd + go + k0 * gc0 + k1 * gc1 + k2 * gc2;
                if (PART) {
                    float hu = bgv;
#pragma unroll 8
                    for (int j = 0; j < H; ++j) hu += act[(R_HP + j) * PS + p] * sm[SM_UT + j * RP + r];
                    g += hu;
                }
            }
            const float gT = g * t;
            float suffix = 0.f;
#pragma unroll
            for (int q = S - 1; q >= 1; --q) {
                const float v = __shfl_sync(FULL, gT, q);
                if (q > lane) suffix += v;
            }
            if (smp) {
                float draw = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
                if (live_ray) {
                    const float docc = g * fp - suffix / (1.f - occ + 1e-10f);
                    draw = docc * occ * (1.f - occ) * 10.f;
                    d0 = t * gc0 * k0 * (1.f - k0);
                    d1 = t * gc1 * k1 * (1.f - k1);
                    d2 = t * gc2 * k2 * (1.f - k2);
                }
                misc[(M_DCOL + 0) * PS + p] = d0;
                misc[(M_DCOL + 1) * PS + p] = d1;
                misc[(M_DCOL + 2) * PS + p] = d2;
                misc[M_DRAW * PS + p] = draw;
            }
            OO_ACC(loss, 4, AC_LOSS);
            loss[0] += l0; loss[1] += l1; loss[2] += l2; loss[3] += l3;      // non-zero in lane 0 only
            OO_ACC_PUT(loss, 4, AC_LOSS);
        }
    } else if constexpr (PH == 42) {
        // ---- fused: M / m / beta accumulators (phase 13), out_color / out_alpha weight gradients and d(hp_pre) (phase 18),
        // then (after a block barrier: it overwrites the hc rows the out_color weight gradient reads) d(hc_pre) (phase 19)
        if (PART) {
            if (tid >= M_T0 && tid < M_T0 + 8 * H) {
                OO_ACC(g_gm, 4, AC_GM);
                const int k = (tid - M_T0) >> 3, j4 = 4 * ((tid - M_T0) & 7);
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const float bs = rv[V_B * RP + r] * sm[SM_ST + k * RP + r];
                    g_gm[0] += bs * sm[SM_ST + (j4 + 0) * RP + r];
                    g_gm[1] += bs * sm[SM_ST + (j4 + 1) * RP + r];
                    g_gm[2] += bs * sm[SM_ST + (j4 + 2) * RP + r];
                    g_gm[3] += bs * sm[SM_ST + (j4 + 3) * RP + r];
                }
                OO_ACC_PUT(g_gm, 4, AC_GM);
            }
            if (tid < 2 * 32) {
                OO_ACC(g_mv, 1, AC_MV);
                if (tid <= H) {
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        const float bo = rv[V_B * RP + r] * rv[V_OPAC * RP + r];
                        g_mv[0] += bo * (tid < H ? sm[SM_ST + tid * RP + r] : rv[V_OPAC * RP + r]);
                    }
                }
                OO_ACC_PUT(g_mv, 1, AC_MV);
            }
        }
        if (tid >= OC_T0 && tid < OC_T0 + 4 * H) {
            OO_ACC(g_oc, 1, AC_OC);
            const int o = (tid - OC_T0) >> 5, j = tid & 31;
            const float* dy = misc + (o < 3 ? (M_DCOL + o) : M_DRAW) * PS;
            const float* x = act + ((o < 3 ? R_HC : R_H4) + j) * PS;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;      // four independent chains
#pragma unroll
            for (int p0 = 0; p0 < P; p0 += 4) {
                const float4 d = ld4(dy + p0), h = ld4(x + p0);
                s0 += d.x * h.x; s1 += d.y * h.y; s2 += d.z * h.z; s3 += d.w * h.w;
            }
            g_oc[0] += (s0 + s1) + (s2 + s3);
            OO_ACC_PUT(g_oc, 1, AC_OC);
        }
        if (PART) {
            for (int i = tid; i < H * (P / 4); i += NTHREADS) {      // d(hp_pre) = T_p * U_r * [hp > 0] in place
                const int j = i / (P / 4), p0 = 4 * (i - j * (P / 4));
                float* hp = act + (R_HP + j) * PS + p0;
                const float4 h = ld4(hp), t = ld4(misc + M_TERM * PS + p0);
                const float* ut = sm + SM_UT + j * RP;
                float4 o;
                o.x = h.x > 0.f ? t.x * ut[(p0 + 0) / S] : 0.f;
                o.y = h.y > 0.f ? t.y * ut[(p0 + 1) / S] : 0.f;
                o.z = h.z > 0.f ? t.z * ut[(p0 + 2) / S] : 0.f;
                o.w = h.w > 0.f ? t.w * ut[(p0 + 3) / S] : 0.f;
                st4(hp, o);
            }
        }
        __syncthreads();       // the out_color weight gradient above read hc; d(hc_pre) now replaces it in place
        for (int i = tid; i < H * (P / 4); i += NTHREADS) {          // d(hc_pre) = (W_oc^T dcol_pre) * [hc > 0]
            const int j = i / (P / 4), p0 = 4 * (i - j * (P / 4));
            float* hc = act + (R_HC + j) * PS + p0;
            const float4 h = ld4(hc);
            const float4 d0 = ld4(misc + (M_DCOL + 0) * PS + p0), d1 = ld4(misc + (M_DCOL + 1) * PS + p0),
                         d2 = ld4(misc + (M_DCOL + 2) * PS + p0);
            const float w0 = w[W_OC + j], w1 = w[W_OC + H + j], w2 = w[W_OC + 2 * H + j];
            float4 o;
            o.x = h.x > 0.f ? w0 * d0.x + w1 * d1.x + w2 * d2.x : 0.f;
            o.y = h.y > 0.f ? w0 * d0.y + w1 * d1.y + w2 * d2.y : 0.f;
            o.z = h.z > 0.f ? w0 * d0.z + w1 * d1.z + w2 * d2.z : 0.f;
            o.w = h.w > 0.f ? w0 * d0.w + w1 * d1.w + w2 * d2.w : 0.f;
            st4(hc, o);
        }
    } else if constexpr (PH == 43) {
        // ---- general upstream (oo_forward_bwd), after forward phases 2..7: gradients w.r.t. the raw head outputs.
        //   d raw_alpha = 10 d_alpha (model.py:88);  d col_pre = d_color * col (1 - col) (model.py:96);
        //   d hp_pre = (d_clip W_ocl) * [hp > 0], the product having been formed by k_clip_dhp; hp itself leaves for the
        //   out_clip weight gradient (k_clip_dw) before it is overwritten.  Points beyond npts contribute nothing.
        for (int p = tid; p < P; p += NTHREADS) {
            const bool in = p < c.npts;
            misc[M_DRAW * PS + p] = in ? 10.f * OO_LDG(c.up_alpha + p) : 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float k = misc[(M_COL + ch) * PS + p];
                misc[(M_DCOL + ch) * PS + p] = in ? OO_LDG(c.up_color + 3 * p + ch) * k * (1.f - k) : 0.f;
            }
        }
        for (int i = tid; i < P * H; i += NTHREADS) {
            const int p = i / H, j = i - p * H;
            float* hp = act + (R_HP + j) * PS + p;
            const float h = *hp;
            const bool in = p < c.npts;
            if (in && c.hp_out != nullptr) c.hp_out[i] = h;
            *hp = (in && c.up_hp != nullptr && h > 0.f) ? OO_LDG(c.up_hp + i) : 0.f;
        }
    } else if constexpr (PH == 44) {
        // ---- out_color / out_alpha weight gradients, then (after a block barrier) d(hc_pre) in place: phase 42 without the
        // compositing-specific parts
        if (tid >= OC_T0 && tid < OC_T0 + 4 * H) {
            OO_ACC(g_oc, 1, AC_OC);
            const int o = (tid - OC_T0) >> 5, j = tid & 31;
            const float* dy = misc + (o < 3 ? (M_DCOL + o) : M_DRAW) * PS;
            const float* x = act + ((o < 3 ? R_HC : R_H4) + j) * PS;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int p0 = 0; p0 < P; p0 += 4) {
                const float4 d = ld4(dy + p0), h = ld4(x + p0);
                s0 += d.x * h.x; s1 += d.y * h.y; s2 += d.z * h.z; s3 += d.w * h.w;
            }
            g_oc[0] += (s0 + s1) + (s2 + s3);
            OO_ACC_PUT(g_oc, 1, AC_OC);
        }
        __syncthreads();
        for (int i = tid; i < H * (P / 4); i += NTHREADS) {          // d(hc_pre) = (W_oc^T dcol_pre) * [hc > 0]
            const int j = i / (P / 4), p0 = 4 * (i - j * (P / 4));
            float* hc = act + (R_HC + j) * PS + p0;
            const float4 h = ld4(hc);
            const float4 d0 = ld4(misc + (M_DCOL + 0) * PS + p0), d1 = ld4(misc + (M_DCOL + 1) * PS + p0),
                         d2 = ld4(misc + (M_DCOL + 2) * PS + p0);
            const float w0 = w[W_OC + j], w1 = w[W_OC + H + j], w2 = w[W_OC + 2 * H + j];
            float4 o;
            o.x = h.x > 0.f ? w0 * d0.x + w1 * d1.x + w2 * d2.x : 0.f;
            o.y = h.y > 0.f ? w0 * d0.y + w1 * d1.y + w2 * d2.y : 0.f;
            o.z = h.z > 0.f ? w0 * d0.z + w1 * d1.z + w2 * d2.z : 0.f;
            o.w = h.w > 0.f ? w0 * d0.w + w1 * d1.w + w2 * d2.w : 0.f;
            st4(hc, o);
        }
#endif
    } else if constexpr (PH == 33) {
        // per point: exclusive product of the free probabilities in torch.cumprod's order, termination weight
        // (render_rays.py:36-43)
        for (int p = tid; p < P; p += NTHREADS) {
            const int r = p / S, i = p - r * S;
            float freep = 1.f;
            for (int q = 0; q < i; ++q) freep *= (1.f - misc[M_OCC * PS + r * S + q] + 1e-10f);
            const bool live = r < c.nrays;
            misc[M_FREE * PS + p] = live ? freep : 0.f;
            misc[M_TERM * PS + p] = live ? misc[M_OCC * PS + p] * freep : 0.f;
        }
    } else if constexpr (PH == 8) {
        // per ray: rendered depth / variance / colour / opacity, loss terms and their derivatives w.r.t. the rendered
        // quantities (render_rays.py:56-63, loss.py:27-75)
        // warp r renders ray r: lanes = samples, sums by shuffle; lane 0 (thread 32 r) owns the ray's loss partials
        const int r = tid >> 5, li = tid & 31;
        if (r < RT) {
            OO_ACC(loss, 4, AC_LOSS);
            float gd = 0.f, go = 0.f, gc0 = 0.f, gc1 = 0.f, gc2 = 0.f, cf = 0.f;
            float depth = 0.f, opac = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, var = 0.f;
#ifdef __CUDACC__
            {
                const bool live = r < c.nrays && li < S;
                const int p = r * S + (li < S ? li : 0);
                const float t = live ? misc[M_TERM * PS + p] : 0.f;
                const float zv = live ? misc[M_Z * PS + p] : 0.f;
                depth = warp_sum(t * zv);
                opac = warp_sum(t);
                c0 = warp_sum(t * misc[(M_COL + 0) * PS + p]);
                c1 = warp_sum(t * misc[(M_COL + 1) * PS + p]);
                c2 = warp_sum(t * misc[(M_COL + 2) * PS + p]);
                const float dz = zv - depth;
                var = warp_sum(t * (dz * dz));
            }
#else
            if (li == 0 && r < c.nrays) {
                for (int i = 0; i < S; ++i) {
                    const int p = r * S + i;
                    const float t = misc[M_TERM * PS + p];
                    depth += t * misc[M_Z * PS + p];
                    opac += t;
                    c0 += t * misc[(M_COL + 0) * PS + p];
                    c1 += t * misc[(M_COL + 1) * PS + p];
                    c2 += t * misc[(M_COL + 2) * PS + p];
                }
                for (int i = 0; i < S; ++i) {
                    const float dz = misc[M_Z * PS + r * S + i] - depth;
                    var += misc[M_TERM * PS + r * S + i] * (dz * dz);
                }
            }
#endif
            if (li == 0) {
                if (r < c.nrays) {
                    const int lab = (int)rv[V_LAB * RP + r];
                    const bool is1 = lab == 1, sem = lab != 2;
                    const float tgt = lab != 0 ? 1.f : 0.f;
                    if (is1 && !(c.flags & 2)) {
                        const float wgt = 1.f / (sqrtf(var) + 1e-4f);           // render_rays.py:95-100, var detached
                        const float dd = depth - rv[V_GTD * RP + r];
                        loss[0] += fabsf(dd) * wgt;
                        gd = sgnf_(dd) * wgt * c.inv1;
                        const float e0 = c0 - rv[(V_RGB + 0) * RP + r] / 255.f;   // train.py:373 `/ 255.`
                        const float e1 = c1 - rv[(V_RGB + 1) * RP + r] / 255.f;
                        const float e2 = c2 - rv[(V_RGB + 2) * RP + r] / 255.f;
                        loss[1] += fabsf(e0) + fabsf(e1) + fabsf(e2);                // loss.py:61 sum over channels
                        gc0 = sgnf_(e0) * c.cs * c.inv1;
                        gc1 = sgnf_(e1) * c.cs * c.inv1;
                        gc2 = sgnf_(e2) * c.cs * c.inv1;
                        cf = PART ? c.fs * c.inv1 : 0.f;
                    }
                    if (sem && !(c.flags & 4)) {
                        const float eo = opac - tgt;                                 // loss.py:71
                        loss[2] += fabsf(eo);
                        go = sgnf_(eo) * c.os * c.invs;
                    }
                }
                rv[V_DEPTH * RP + r] = depth;
                rv[V_OPAC * RP + r] = opac;
                rv[(V_COL + 0) * RP + r] = c0; rv[(V_COL + 1) * RP + r] = c1; rv[(V_COL + 2) * RP + r] = c2;
                rv[V_GD * RP + r] = gd; rv[V_GO * RP + r] = go;
                rv[(V_GC + 0) * RP + r] = gc0; rv[(V_GC + 1) * RP + r] = gc1; rv[(V_GC + 2) * RP + r] = gc2;
                rv[V_CF * RP + r] = cf;
                rv[V_BG * RP + r] = 0.f;
            }
            OO_ACC_PUT(loss, 4, AC_LOSS);
        }
        if (PART) OO_CP_ASYNC_WAIT();      // Y rows issued in phase 0 are complete for this thread; the barrier publishes them
    } else if constexpr (PH == 10) {
        // The rendered feature x_r = W S_r + b opac_r is never formed.  With per-object constants G = W^T W,
        // wb = W^T b, bb = b.b (k_gram) the cosine loss and its gradient need only
        //     v_r = W^T y_r,  yb_r = b.y_r,  yy_r = y_r.y_r        (y_r = gt feature row, staged in Y)
        //   x.y = S.v + opac yb ;  x.x = S.(G S + opac wb) + opac (S.wb + opac bb) ;  dL/dS = A v + B (G S + opac wb).
        if (PART) {
            // S[j][r] = sum_i T_i hp_i[j]   (render of the clip-head hidden activations); independent of the rest of the phase
            for (int i = tid; i < H * RT; i += NTHREADS) {
                const int j = i / RT, r = i - j * RT;
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < S; ++q) s += misc[M_TERM * PS + r * S + q] * act[(R_HP + j) * PS + r * S + q];
                sm[SM_ST + j * RP + r] = s;
            }
            // warp wv owns gt-feature columns [32 wv, 32 wv + 32): partial v_r[j] = sum_c y_r[c] W[c][j] as a 16 x 32 x 32
            // tensor-core product (M = ray, N = hidden unit, K = column; W fragments straight from L2), y.b as a fifth
            // column tile whose column 0 is b, y.y on the side.  The partial v replaces the warp's own columns of Y
            // (nobody else reads them); phase 11 sums the 16 partials.
            constexpr int CW = C / NWARPS;
            static_assert(CW == 32, "phase 10 gives each warp four 8-wide k-steps of gt-feature columns");
#ifdef __CUDACC__
            const int wv = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
            const float* wg = c.theta + OFF_OCL_W + (size_t)(wv * CW) * H;
            const float* bg = c.theta + OFF_OCL_B + wv * CW;
            float bw[4][4][2], bq[4][2];            // global loads in flight before the first use
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    bw[ks][nt][0] = OO_LDG(wg + (8 * ks + t) * H + 8 * nt + g);
                    bw[ks][nt][1] = OO_LDG(wg + (8 * ks + t + 4) * H + 8 * nt + g);
                }
                bq[ks][0] = g == 0 ? OO_LDG(bg + 8 * ks + t) : 0.f;
                bq[ks][1] = g == 0 ? OO_LDG(bg + 8 * ks + t + 4) : 0.f;
            }
            float* y = sm + SM_FEAT + wv * CW;
            const bool hi_row = g + 8 < RT;
            float vacc[5][4];
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) vacc[nt][0] = vacc[nt][1] = vacc[nt][2] = vacc[nt][3] = 0.f;
            float yy0 = 0.f, yy1 = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const float a0 = y[g * YSTR + 8 * ks + t], a2 = y[g * YSTR + 8 * ks + t + 4];
                const float a1 = hi_row ? y[(g + 8) * YSTR + 8 * ks + t] : 0.f, a3 = hi_row ? y[(g + 8) * YSTR + 8 * ks + t + 4] : 0.f;
                yy0 += a0 * a0 + a2 * a2;
                yy1 += a1 * a1 + a3 * a3;
                FragA fa;
                frag_a(fa, a0, a1, a2, a3);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    FragB fb;
                    frag_b(fb, bw[ks][nt][0], bw[ks][nt][1]);
                    mma3(vacc[nt], fa, fb);
                }
                FragB fq;
                frag_b(fq, bq[ks][0], bq[ks][1]);
                mma3(vacc[4], fa, fq);
            }
            yy0 += __shfl_xor_sync(0xffffffffu, yy0, 1); yy0 += __shfl_xor_sync(0xffffffffu, yy0, 2);
            yy1 += __shfl_xor_sync(0xffffffffu, yy1, 1); yy1 += __shfl_xor_sync(0xffffffffu, yy1, 2);
            if (t == 0) {
                sm[SM_YS + (g * 16 + wv) * 2] = vacc[4][0];
                sm[SM_YS + (g * 16 + wv) * 2 + 1] = yy0;
                if (hi_row) {
                    sm[SM_YS + ((g + 8) * 16 + wv) * 2] = vacc[4][2];
                    sm[SM_YS + ((g + 8) * 16 + wv) * 2 + 1] = yy1;
                }
            }
            __syncwarp();
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                *reinterpret_cast<float2*>(y + g * YSTR + 8 * nt + 2 * t) = float2{vacc[nt][0], vacc[nt][1]};
                if (hi_row) *reinterpret_cast<float2*>(y + (g + 8) * YSTR + 8 * nt + 2 * t) = float2{vacc[nt][2], vacc[nt][3]};
            }
#else
            if (tid == 0) {      // host emulation: same partial sums, sequential order inside a warp's column block
                for (int wv = 0; wv < NWARPS; ++wv) {
                    float* y = sm + SM_FEAT + wv * CW;
                    float vv[RT][CW];
                    for (int r = 0; r < RT; ++r) {
                        float yb = 0.f, yy = 0.f;
                        for (int q = 0; q < CW; ++q) {
                            yb += y[r * YSTR + q] * c.theta[OFF_OCL_B + wv * CW + q];
                            yy += y[r * YSTR + q] * y[r * YSTR + q];
                        }
                        sm[SM_YS + (r * 16 + wv) * 2] = yb;
                        sm[SM_YS + (r * 16 + wv) * 2 + 1] = yy;
                        for (int j = 0; j < H; ++j) {
                            float v = 0.f;
                            for (int q = 0; q < CW; ++q) v += y[r * YSTR + q] * c.theta[OFF_OCL_W + (size_t)(wv * CW + q) * H + j];
                            vv[r][j] = v;
                        }
                    }
                    for (int r = 0; r < RT; ++r)
                        for (int j = 0; j < H; ++j) y[r * YSTR + j] = vv[r][j];
                }
            }
#endif
        }
    } else if constexpr (PH == 11) {
        if (PART) {
            for (int i = tid; i < H * RT; i += NTHREADS) {      // v[j][r] -> SM_UT
                const int r = i / H, j = i - r * H;
                float v = 0.f;
#pragma unroll
                for (int wv = 0; wv < NWARPS; ++wv) v += sm[SM_FEAT + r * YSTR + wv * (C / NWARPS) + j];
                sm[SM_UT + j * RP + r] = v;
            }
            if (tid >= 128 && tid < 128 + 2 * RT) {             // totals of y.b and y.y
                const int q = (tid - 128) / RT, r = (tid - 128) - q * RT;
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int ch = 0; ch < 16; ch += 2) {
                    s0 += sm[SM_YS + (r * 16 + ch) * 2 + q];
                    s1 += sm[SM_YS + (r * 16 + ch + 1) * 2 + q];
                }
                sm[SM_YS + 320 + q * RP + r] = s0 + s1;
            }
            // gs[j][r] = sum_k G[j][k] S[k][r] + opac_r wb[j]   (S from phase 10)
            for (int i = tid; i < H * RT; i += NTHREADS) {
                const int r = i / H, j = i - r * H;
                float g[H];
#pragma unroll
                for (int k = 0; k < H; ++k) g[k] = c.derived[DER_G + k * H + j];      // G is symmetric: coalesced over j
                float acc = rv[V_OPAC * RP + r] * rv[V_WB * RP + j];
#pragma unroll
                for (int k = 0; k < H; ++k) acc += g[k] * sm[SM_ST + k * RP + r];
                sm[SM_UPART + j * RP + r] = acc;
            }
        }
    } else if constexpr (PH == 32) {
        // per ray: cosine, loss partial, and the two coefficients of dL/dx = A y + B x
        if (PART) {
            // warp r: lanes = hidden units, three dot products by shuffle; lane 0 (thread 32 r) owns the ray's loss partial
            const int r = tid >> 5, lj = tid & 31;
            float sv = 0.f, swb = 0.f, sgs = 0.f;
            float lossf = 0.f;
#ifdef __CUDACC__
            if (r < RT) {
                const float sj = sm[SM_ST + lj * RP + r];
                sv = warp_sum(sj * sm[SM_UT + lj * RP + r]);
                swb = warp_sum(sj * rv[V_WB * RP + lj]);
                sgs = warp_sum(sj * sm[SM_UPART + lj * RP + r]);
            }
#else
            if (r < RT && lj == 0)
                for (int j = 0; j < H; ++j) {
                    const float sj = sm[SM_ST + j * RP + r];
                    sv += sj * sm[SM_UT + j * RP + r];
                    swb += sj * rv[V_WB * RP + j];
                    sgs += sj * sm[SM_UPART + j * RP + r];
                }
#endif
            if (r < RT && lj == 0) {
                const float opac = rv[V_OPAC * RP + r];
                const float bb = rv[V_WB * RP + H];
                const float yb = sm[SM_YS + 320 + r], yy = sm[SM_YS + 320 + RP + r];
                const float xb = swb + opac * bb;                     // b . x
                const float xy = sv + opac * yb, xx = sgs + opac * xb;
                const float cf = rv[V_CF * RP + r];
                const float nxr = sqrtf(fmaxf(xx, 0.f)), nyr = sqrtf(yy);
                const float nx = fmaxf(nxr, 1e-8f), ny = fmaxf(nyr, 1e-8f);   // F.cosine_similarity eps clamp
                const float cosv = xy / (nx * ny);
                float A = 0.f, B = 0.f;
                if (cf != 0.f) {
                    lossf = 1.f - cosv;
                    // d(1-cos)/dx = -y/(nx ny) + [|x|>eps] cos * x / nx^2
                    A = -cf / (nx * ny);
                    B = nxr > 1e-8f ? cf * cosv / (nx * nx) : 0.f;
                }
                rv[V_A * RP + r] = A;
                rv[V_B * RP + r] = B;
                rv[V_BG * RP + r] = A * yb + B * xb;       // b_ocl . dL/dx
            }
            if (r < RT) {                              // warp-uniform: the accumulator access is a warp-wide operation
                OO_ACC(loss, 4, AC_LOSS);
                loss[3] += lossf;
                OO_ACC_PUT(loss, 4, AC_LOSS);
            }
        }
    } else if constexpr (PH == 13) {
        if (PART) {
            // dL/dS -> SM_UT (in place over v); per-ray record for K4; M, m, beta accumulators of the out_clip gradient:
            //   dW = sum_r (A_r y_r + B_r x_r) S_r^T = sum_r A_r y_r S_r^T + W M + b m^T,  M = sum_r B_r S_r S_r^T, m = sum_r B_r opac_r S_r
            for (int i = tid; i < H * RT; i += NTHREADS) {
                const int r = i / H, j = i - r * H;
                sm[SM_UT + j * RP + r] = rv[V_A * RP + r] * sm[SM_UT + j * RP + r] + rv[V_B * RP + r] * sm[SM_UPART + j * RP + r];
            }
            for (int i = tid; i < c.nrays * (H + 2); i += NTHREADS) {
                const int r = i / (H + 2), q = i - r * (H + 2);
                if (q < 2) c.rayrec[r * RAYREC + q] = q == REC_A ? rv[V_A * RP + r] : rv[V_OPAC * RP + r];
                else c.rayrec[r * RAYREC + REC_S + q - 2] = sm[SM_ST + (q - 2) * RP + r];
            }
            if (tid >= M_T0 && tid < M_T0 + 8 * H) {
                OO_ACC(g_gm, 4, AC_GM);
                const int k = (tid - M_T0) >> 3, j4 = 4 * ((tid - M_T0) & 7);
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const float bs = rv[V_B * RP + r] * sm[SM_ST + k * RP + r];
                    g_gm[0] += bs * sm[SM_ST + (j4 + 0) * RP + r];
                    g_gm[1] += bs * sm[SM_ST + (j4 + 1) * RP + r];
                    g_gm[2] += bs * sm[SM_ST + (j4 + 2) * RP + r];
                    g_gm[3] += bs * sm[SM_ST + (j4 + 3) * RP + r];
                }
                OO_ACC_PUT(g_gm, 4, AC_GM);
            }
            if (tid < 2 * 32) {                        // warps 0 and 1 (thread 32 holds beta): warp-uniform accumulator access
                OO_ACC(g_mv, 1, AC_MV);
                if (tid <= H) {
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        const float bo = rv[V_B * RP + r] * rv[V_OPAC * RP + r];
                        g_mv[0] += bo * (tid < H ? sm[SM_ST + tid * RP + r] : rv[V_OPAC * RP + r]);
                    }
                }
                OO_ACC_PUT(g_mv, 1, AC_MV);
            }
        }
    } else if constexpr (PH == 16) {
        // per point: g = dL/dT = z gd + go + col . gc (+ hp . U_r + b_ocl . g_r : what one unit of termination weight
        // at this point adds to the feature loss).  g -> M_DRAW row, g*T -> M_HU row.
        for (int p = tid; p < P; p += NTHREADS) {
            const int r = p / S;
            float g = 0.f;
            if (r < c.nrays) {
                g = misc[M_Z * PS + p] * rv[V_GD * RP + r] + rv[V_GO * RP + r] +
                    misc[(M_COL + 0) * PS + p] * rv[(V_GC + 0) * RP + r] + misc[(M_COL + 1) * PS + p] * rv[(V_GC + 1) * RP + r] +
                    misc[(M_COL + 2) * PS + p] * rv[(V_GC + 2) * RP + r];
                if (PART) {
                    float hu = rv[V_BG * RP + r];
#pragma unroll 8
                    for (int j = 0; j < H; ++j) hu += act[(R_HP + j) * PS + p] * sm[SM_UT + j * RP + r];
                    g += hu;
                }
            }
            misc[M_DRAW * PS + p] = g;
            misc[M_HU * PS + p] = g * misc[M_TERM * PS + p];
        }
    } else if constexpr (PH == 17) {
        // back through the compositing sums and the exclusive product, one thread per point:
        //   g_i = dL/dT_i ; dL/do_i = g_i P_i - (sum_{k>i} g_k T_k) / (1 - o_i + 1e-10) ; alpha = 10 raw
        // phase 34 stored g_i T_i in the M_HU row (after adding the feature term), g_i in M_DRAW.
        for (int p = tid; p < P; p += NTHREADS) {
            const int r = p / S, i = p - r * S;
            float draw = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
            if (r < c.nrays) {
                float suffix = 0.f;
                for (int q = S - 1; q > i; --q) suffix += misc[M_HU * PS + r * S + q];
                const float t = misc[M_TERM * PS + p], o = misc[M_OCC * PS + p], fp = misc[M_FREE * PS + p];
                const float k0 = misc[(M_COL + 0) * PS + p], k1 = misc[(M_COL + 1) * PS + p], k2 = misc[(M_COL + 2) * PS + p];
                const float g = misc[M_DRAW * PS + p];
                const float docc = g * fp - suffix / (1.f - o + 1e-10f);
                draw = docc * o * (1.f - o) * 10.f;
                const float gc0 = rv[(V_GC + 0) * RP + r], gc1 = rv[(V_GC + 1) * RP + r], gc2 = rv[(V_GC + 2) * RP + r];
                d0 = t * gc0 * k0 * (1.f - k0);
                d1 = t * gc1 * k1 * (1.f - k1);
                d2 = t * gc2 * k2 * (1.f - k2);
            }
            misc[(M_DCOL + 0) * PS + p] = d0;
            misc[(M_DCOL + 1) * PS + p] = d1;
            misc[(M_DCOL + 2) * PS + p] = d2;
            misc[M_DRAW * PS + p] = draw;      // only this thread read g from this element
        }
    } else if constexpr (PH == 18) {
        // out_color / out_alpha weight gradients (reduce over points) ...
        if (tid >= OC_T0 && tid < OC_T0 + 4 * H) {
            OO_ACC(g_oc, 1, AC_OC);
            const int o = (tid - OC_T0) >> 5, j = tid & 31;
            const float* dy = misc + (o < 3 ? (M_DCOL + o) : M_DRAW) * PS;
            const float* x = act + ((o < 3 ? R_HC : R_H4) + j) * PS;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;      // four independent chains
#pragma unroll
            for (int p0 = 0; p0 < P; p0 += 4) {
                const float4 d = ld4(dy + p0), h = ld4(x + p0);
                s0 += d.x * h.x; s1 += d.y * h.y; s2 += d.z * h.z; s3 += d.w * h.w;
            }
            g_oc[0] += (s0 + s1) + (s2 + s3);
            OO_ACC_PUT(g_oc, 1, AC_OC);
        }
        // ... and d(hp_pre) = T_p * U_r * [hp > 0] in place
        if (PART) {
            for (int i = tid; i < H * (P / 4); i += NTHREADS) {
                const int j = i / (P / 4), p0 = 4 * (i - j * (P / 4));
                float* hp = act + (R_HP + j) * PS + p0;
                const float4 h = ld4(hp), t = ld4(misc + M_TERM * PS + p0);
                const float* ut = sm + SM_UT + j * RP;
                float4 o;
                o.x = h.x > 0.f ? t.x * ut[(p0 + 0) / S] : 0.f;
                o.y = h.y > 0.f ? t.y * ut[(p0 + 1) / S] : 0.f;
                o.z = h.z > 0.f ? t.z * ut[(p0 + 2) / S] : 0.f;
                o.w = h.w > 0.f ? t.w * ut[(p0 + 3) / S] : 0.f;
                st4(hp, o);
            }
        }
    } else if constexpr (PH == 19) {
        // d(hc_pre) = (W_oc^T dcol_pre) * [hc > 0] in place
        for (int i = tid; i < H * (P / 4); i += NTHREADS) {
            const int j = i / (P / 4), p0 = 4 * (i - j * (P / 4));
            float* hc = act + (R_HC + j) * PS + p0;
            const float4 h = ld4(hc);
            const float4 d0 = ld4(misc + (M_DCOL + 0) * PS + p0), d1 = ld4(misc + (M_DCOL + 1) * PS + p0),
                         d2 = ld4(misc + (M_DCOL + 2) * PS + p0);
            const float w0 = w[W_OC + j], w1 = w[W_OC + H + j], w2 = w[W_OC + 2 * H + j];
            float4 o;
            o.x = h.x > 0.f ? w0 * d0.x + w1 * d1.x + w2 * d2.x : 0.f;
            o.y = h.y > 0.f ? w0 * d0.y + w1 * d1.y + w2 * d2.y : 0.f;
            o.z = h.z > 0.f ? w0 * d0.z + w1 * d1.z + w2 * d2.z : 0.f;
            o.w = h.w > 0.f ? w0 * d0.w + w1 * d1.w + w2 * d2.w : 0.f;
            st4(hc, o);
        }
    } else if constexpr (PH == 20) {
        // [color_linear ; clip_linear] weight gradient: rows = [d_hc ; d_hp], cols = [h4 ; e2]
        OO_ACC(w_hd, 12, AC_HD);
        if (PART) gemm_bwd_w<4, KP_HD / 8>(w_hd, tid, act + R_HC * PS, act + R_H4 * PS);
        else gemm_bwd_w<2, KP_HD / 8>(w_hd, tid, act + R_HC * PS, act + R_H4 * PS);
        OO_ACC_PUT(w_hd, 12, AC_HD);
    } else if constexpr (PH == 21) {
        // d[h4 ; e2] = W_cl^T d_hc + W_cp^T d_hp (+ W_a draw on the h4 rows), ReLU mask on the h4 rows; in place
        if (PART)
            gemm_bwd_data<KP_HD, WS_HD, 2 * H, 8, 0>(tid, w + W_CL, act + R_HC * PS, nullptr, nullptr,
                                                     act + R_H4 * PS, H, w + W_A, misc + M_DRAW * PS);
        else
            gemm_bwd_data<KP_HD, WS_HD, H, 8, 0>(tid, w + W_CL, act + R_HC * PS, nullptr, nullptr,
                                                 act + R_H4 * PS, H, w + W_A, misc + M_DRAW * PS);
    } else if constexpr (PH == 22) {
        OO_ACC(w_m2, 4, AC_M2);
        gemm_bwd_w32(w_m2, tid, act + R_H4 * PS, act + R_H3 * PS);
        OO_ACC_PUT(w_m2, 4, AC_M2);
    } else if constexpr (PH == 23) {
        gemm_bwd_data<H, WS_H, H, 8, 0>(tid, w + W_M2, act + R_H4 * PS, nullptr, nullptr, act + R_H3 * PS, H,
                                        nullptr, nullptr);
    } else if constexpr (PH == 24) {
        OO_ACC(w_cat, 8, AC_CAT);
        gemm_bwd_w<2, KP_CAT / 8>(w_cat, tid, act + R_H3 * PS, act + R_H2 * PS);
        OO_ACC_PUT(w_cat, 8, AC_CAT);
    } else if constexpr (PH == 25) {
        gemm_bwd_data<H, WS_CAT, H, 8, 0>(tid, w + W_CAT, act + R_H3 * PS, nullptr, nullptr, act + R_H2 * PS, H,
                                          nullptr, nullptr);
    } else if constexpr (PH == 26) {
        OO_ACC(w_m1, 4, AC_M1);
        gemm_bwd_w32(w_m1, tid, act + R_H2 * PS, act + R_H1 * PS);
        OO_ACC_PUT(w_m1, 4, AC_M1);
    } else if constexpr (PH == 27) {
        gemm_bwd_data<H, WS_H, H, 8, 0>(tid, w + W_M1, act + R_H2 * PS, nullptr, nullptr, act + R_H1 * PS, H,
                                        nullptr, nullptr);
    } else if constexpr (PH == 28) {
        OO_ACC(w_in, 8, AC_IN);
        gemm_bwd_w<2, KP_IN / 8>(w_in, tid, act + R_H1 * PS, act + R_E1 * PS);
        OO_ACC_PUT(w_in, 8, AC_IN);
    } else if constexpr (PH == 29) {
        // d e1 = W_cat[:, 32:]^T d_h3 + W_in^T d_h1, in place over e1 (no mask)
        gemm_bwd_data<KP_IN, WS_CAT, H, WS_IN, H>(tid, w + W_CAT + H, act + R_H3 * PS, w + W_IN, act + R_H1 * PS,
                                                  act + R_E1 * PS, 0, nullptr, nullptr);
    } else if constexpr (PH == 30) {
        // d proj[d][p] = sum_k d e[3+21k+d][p] * pi 2^k cos(pi 2^k proj); stored over e1 row 3+d
#ifdef __CUDACC__
        float trig[10];
        tm_ld<8>(a.tm + AC_TRIG, trig);
        tm_ld<2>(a.tm + AC_TRIG + 8, trig + 8);
#endif
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const int i = tid + NTHREADS * u;
            if (i >= NDIR * P) continue;
            const int d = i / P, p = i - d * P;
            float sn, cs;
#ifdef __CUDACC__
            sn = trig[u];
            cs = trig[5 + u];
#else
            const float t0 = act[(R_T + 0) * PS + p], t1 = act[(R_T + 1) * PS + p], t2 = act[(R_T + 2) * PS + p];
            const float proj = w[W_PE + 3 * d] * t0 + w[W_PE + 3 * d + 1] * t1 + w[W_PE + 3 * d + 2] * t2;
            sincosf(proj * PI_F, &sn, &cs);
#endif
            float band = PI_F, dp = 0.f;
#pragma unroll
            for (int k = 0; k < NBAND; ++k) {
                const int row = 3 + NDIR * k + d;
                const float de = row < E1 ? act[(R_E1 + row) * PS + p] : act[(R_E2 + row - E1) * PS + p];
                dp += de * (cs * band);
                const float s2 = 2.f * sn * cs, c2 = (cs - sn) * (cs + sn);
                sn = s2;
                cs = c2;
                band *= 2.f;
            }
            act[(R_E1 + 3 + d) * PS + p] = dp;
        }
    } else if constexpr (PH == 31) {
        if (tid < 64) {                                // warps 0, 1: warp-uniform accumulator access
          OO_ACC(g_pe, 1, AC_PE);
          if (tid < NDIR * 3) {
            const int d = tid / 3, ch = tid - 3 * d;
            const float* dp = act + (R_E1 + 3 + d) * PS;
            const float* tt = act + (R_T + ch) * PS;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int p0 = 0; p0 < P; p0 += 4) {
                const float4 x = ld4(dp + p0), y = ld4(tt + p0);
                s0 += x.x * y.x; s1 += x.y * y.y; s2 += x.z * y.z; s3 += x.w * y.w;
            }
            g_pe[0] += (s0 + s1) + (s2 + s3);
          }
          OO_ACC_PUT(g_pe, 1, AC_PE);
        }
        if (tid < 7 * H) {                             // warps 0..6: warp-uniform accumulator access
          OO_ACC(g_b, 1, AC_B);
          if (tid < 6 * H + 4) {
            const float* row;
            if (tid < 6 * H) {
                const int l = tid >> 5, j = tid & 31;
                const int base = l == 0 ? R_H1 : l == 1 ? R_H2 : l == 2 ? R_H3 : l == 3 ? R_H4 : l == 4 ? R_HC : R_HP;
                row = act + (base + j) * PS;
            } else {
                const int o = tid - 6 * H;
                row = misc + (o < 3 ? (M_DCOL + o) : M_DRAW) * PS;
            }
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int p0 = 0; p0 < P; p0 += 4) {
                const float4 x = ld4(row + p0);
                s0 += x.x; s1 += x.y; s2 += x.z; s3 += x.w;
            }
            if (PART || tid < 5 * H || tid >= 6 * H) g_b[0] += (s0 + s1) + (s2 + s3);
          }
          OO_ACC_PUT(g_b, 1, AC_B);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// flush one (CTA, object) slot: registers -> slab in the reference's tensor layouts
// ------------------------------------------------------------------------------------------------
// step 0: every register accumulator -> slab, ray loss partials -> smem; step 1: loss sum -> slot_loss.
// A block barrier follows every step.
template <int MT, int NT>
OO_DEV void flush_wfrags(int tid, float* __restrict__ slab, const float* acc, int off_lo, int off_hi, int cols) {
    // rows j < 32 go to the tensor at off_lo, rows 32..63 (clip_linear) to off_hi; both [32][cols] row-major
#pragma unroll
    for (int i = 0; i < wfrag_units(MT, NT); ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = wfrag_row(tid, i, r, MT), k = wfrag_col(tid, i, r, MT);
            if (k < cols && k < 8 * NT) slab[(j < H ? off_lo + j * cols : off_hi + (j - H) * cols) + k] = acc[4 * i + r];
        }
}

template <int STEP, bool PART>
OO_DEV void tile_flush(int tid, float* __restrict__ sm, float* __restrict__ slab, float* __restrict__ slot_loss,
                       TileAcc& a) {
    if constexpr (STEP == 0) {
        {
            OO_ACC(w_in, 8, AC_IN);
            flush_wfrags<2, KP_IN / 8>(tid, slab, w_in, OFF_IN_W, 0, E1);
        }
        {
            OO_ACC(w_cat, 8, AC_CAT);
            flush_wfrags<2, KP_CAT / 8>(tid, slab, w_cat, OFF_CAT_W, 0, H + E1);
        }
        {
            OO_ACC(w_m1, 4, AC_M1);
            OO_ACC(w_m2, 4, AC_M2);
            if (tid >= NTHREADS / 2) {                   // second k-half of the 32 x 32 gradients -> scratch (Y area, free here)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int e = wfrag32_row(tid, r) * H + wfrag32_col(tid, r);
                    sm[SM_FEAT + e] = w_m1[r];
                    sm[SM_FEAT + H * H + e] = w_m2[r];
                }
            }
        }
        {
            OO_ACC(w_hd, 12, AC_HD);
            if (PART) flush_wfrags<4, KP_HD / 8>(tid, slab, w_hd, OFF_CL_W, OFF_CP_W, H + E2);
            else flush_wfrags<2, KP_HD / 8>(tid, slab, w_hd, OFF_CL_W, 0, H + E2);
        }
        OO_ACC(g_oc, 1, AC_OC);
        OO_ACC(g_b, 1, AC_B);
        OO_ACC(g_pe, 1, AC_PE);
        if (tid >= OC_T0 && tid < OC_T0 + 3 * H) slab[OFF_OC_W + tid - OC_T0] = g_oc[0];
        else if (tid >= OC_T0 + 3 * H && tid < OC_T0 + 4 * H) slab[OFF_A_W + tid - OC_T0 - 3 * H] = g_oc[0];
        if (tid < 6 * H) {
            const int l = tid >> 5, j = tid & 31;
            const int off = l == 0 ? OFF_IN_B : l == 1 ? OFF_M1_B : l == 2 ? OFF_CAT_B : l == 3 ? OFF_M2_B
                                                                      : l == 4 ? OFF_CL_B : OFF_CP_B;
            if (PART || l < 5) slab[off + j] = g_b[0];
        } else if (tid < 6 * H + 3) {
            slab[OFF_OC_B + tid - 6 * H] = g_b[0];
        } else if (tid == 6 * H + 3) {
            slab[OFF_A_B] = g_b[0];
        }
        if (tid < NDIR * 3) slab[OFF_PE_B + tid] = g_pe[0];
        if (PART) {
            OO_ACC(g_gm, 4, AC_GM);
            OO_ACC(g_mv, 1, AC_MV);
            if (tid >= M_T0 && tid < M_T0 + 8 * H) st4(slab + SLAB_M + 4 * (tid - M_T0), float4{g_gm[0], g_gm[1], g_gm[2], g_gm[3]});
            if (tid <= H) slab[SLAB_MV + tid] = g_mv[0];     // m[0..31], beta at SLAB_MV + 32 == SLAB_BETA
        }
        OO_ACC(loss, 4, AC_LOSS);
        if ((tid & 31) == 0 && (tid >> 5) < RT) {       // thread 32 r owns ray slot r (phases 8 and 32)
            float* rvl = sm + SM_RV;
            const int r = tid >> 5;
            rvl[V_LD * RP + r] = loss[0];
            rvl[V_LC * RP + r] = loss[1];
            rvl[V_LO * RP + r] = loss[2];
            rvl[V_LF * RP + r] = loss[3];
        }
    } else {
        OO_ACC(w_m1, 4, AC_M1);
        OO_ACC(w_m2, 4, AC_M2);
        if (tid < NTHREADS / 2) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int e = wfrag32_row(tid, r) * H + wfrag32_col(tid, r);
                slab[OFF_M1_W + e] = w_m1[r] + sm[SM_FEAT + e];
                slab[OFF_M2_W + e] = w_m2[r] + sm[SM_FEAT + H * H + e];
            }
        }
        if (tid < 4) {
            const float* rvl = sm + SM_RV + (V_LD + tid) * RP;
            float sum = 0.f;
            for (int r = 0; r < RT; ++r) sum += rvl[r];
            slot_loss[tid] = sum;
        }
    }
}

constexpr int N_FLUSH_STEPS = 2;   // barrier after each

}  // namespace oo
