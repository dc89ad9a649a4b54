// a17: the background model (objnerf/train.py:300-315,379-388,447-463; vmap.py:43-47): ONE OccupancyMap of hidden
// width 128 + UniDirsEmbed(scale 5) trained on 1200 rays x 14 samples per step beside the object ensemble.
// M = 16 800 points with K <= 215 are ordinary GEMMs, so this path is layer by layer: an FP32 shared-memory-tiled
// GEMM with fused bias / activation / ReLU-mask epilogues (forward, backward-data, split-M backward-weight with a
// fixed-order reduction), the encoder forward / backward, and the standalone compositing + loss kernels (K3,
// oo_composite.cu) in between.  Hidden width is a run-time argument (any multiple of 4).
#include "../../include/openobj_b200.h"
#include "oo_common.cuh"
#include "oo_layout.h"

using namespace oo;

namespace {

// ------------------------------------------------------------------------------------------------
// parameter layout of a model of hidden width h: the reference's named_parameters() order, offsets rounded to 4 floats
// ------------------------------------------------------------------------------------------------
struct BgLayout {
    int off[NT], size[NT], total;
};

BgLayout bg_layout(int h) {
    const int sz[NT] = {h * E1, h, h * h, h, h * (h + E1), h, h * h, h, h, 1, h * (h + E2), h, 3 * h, 3,
                        h * (h + E2), h, C * h, C, NDIR * 3};
    BgLayout L;
    int o = 0;
    for (int i = 0; i < NT; ++i) {
        L.off[i] = o;
        L.size[i] = sz[i];
        o += (sz[i] + 3) & ~3;
    }
    L.total = o;
    return L;
}

enum { T_IN_W, T_IN_B, T_M1_W, T_M1_B, T_CAT_W, T_CAT_B, T_M2_W, T_M2_B, T_A_W, T_A_B, T_CL_W, T_CL_B, T_OC_W, T_OC_B,
       T_CP_W, T_CP_B, T_OCL_W, T_OCL_B, T_PE };

// ------------------------------------------------------------------------------------------------
// generic FP32 GEMM:  C(i,j) (+)= epilogue( mult * sum_c A(i,c) B(j,c) + bias[j] ),  A(i,c) = A[i*sai + c*sac] etc.
// CTA tile 128 x 64 x 16, 256 threads, 8 x 4 outputs per thread (32 FFMA per three 128-bit shared loads).
// split > 1: grid.z chunks of the contraction write raw partial sums to part[z][I][J]; k_gemm_reduce finishes.
// ------------------------------------------------------------------------------------------------
struct GemmOp {
    const float* A; long long sai, sac;
    const float* B; long long sbj, sbc;
    float* C; long long sci, scj;
    int I, J, K;
    const float* bias;          // [J] or null
    float mult, post;           // v = (mult * acc + bias) * post
    int act;                    // 0 none, 1 relu, 2 sigmoid
    const float* mask; long long smi, smj; int mask_cols;   // C(i,j) = 0 where j < mask_cols and mask(i,j) <= 0
    int accumulate;             // C += result
    int split, chunk;           // contraction chunks (chunk is a multiple of BK)
    float* part;
    float* ones_out;            // non-null: one more virtual column j == J with B(J, c) = 1, written to ones_out[i]
                                // (the bias gradient = column sums rides along with the weight gradient)
};

__device__ __forceinline__ int gemm_je(const GemmOp& g) { return g.J + (g.ones_out != nullptr ? 1 : 0); }
__device__ __forceinline__ float* gemm_dst(const GemmOp& g, int i, int j) {
    return j == g.J ? g.ones_out + i : g.C + i * g.sci + j * g.scj;
}

constexpr int BI = 128, BJ = 64, BK = 16, LDA_S = BI + 4, LDB_S = BJ + 4;

__device__ __forceinline__ float gemm_epilogue(const GemmOp& g, int i, int j, float v) {
    v *= g.mult;
    if (g.bias && j < g.J) v += g.bias[j];
    v *= g.post;
    if (g.act == 1) v = fmaxf(v, 0.f);
    else if (g.act == 2) v = 1.f / (1.f + expf(-v));
    if (g.mask && j < g.mask_cols && !(g.mask[i * g.smi + j * g.smj] > 0.f)) v = 0.f;
    return v;
}

// AC / BC: the contraction index is the contiguous one of A / B (compile-time so that the tile loaders have no
// run-time index arithmetic).  The next k-tile is fetched into registers while the current one is multiplied.
template <bool AC, bool BC>
__global__ void __launch_bounds__(256, 2) k_gemm(const GemmOp g) {
    __shared__ __align__(16) float As[BK][LDA_S];
    __shared__ __align__(16) float Bs[BK][LDB_S];
    const int tid = threadIdx.x, i0 = blockIdx.x * BI, j0 = blockIdx.y * BJ;
    const int c_begin = blockIdx.z * g.chunk, c_end = min(g.K, c_begin + g.chunk);
    const int ty = tid >> 4, tx = tid & 15;
    constexpr int NA = BI * BK / 256, NB_ = BJ * BK / 256;
    // this thread's elements of the A / B tiles: (ii, cc) pairs and their global offsets (without the k-tile offset)
    int a_ii[NA], a_cc[NA], b_jj[NB_], b_cc[NB_];
    int a_off[NA], b_off[NB_];                    // element offsets fit 32 bits (checked in run_gemm)
    bool a_ok[NA], b_ok[NB_], b_one[NB_];
#pragma unroll
    for (int u = 0; u < NA; ++u) {
        const int e = tid + 256 * u;
        a_ii[u] = AC ? e / BK : e % BI;
        a_cc[u] = AC ? e % BK : e / BI;
        a_ok[u] = i0 + a_ii[u] < g.I;
        a_off[u] = (int)((i0 + a_ii[u]) * g.sai + a_cc[u] * g.sac);
    }
#pragma unroll
    for (int u = 0; u < NB_; ++u) {
        const int e = tid + 256 * u;
        b_jj[u] = BC ? e / BK : e % BJ;
        b_cc[u] = BC ? e % BK : e / BJ;
        const int j = j0 + b_jj[u];
        b_ok[u] = j < g.J;
        b_one[u] = j == g.J && g.ones_out != nullptr;
        b_off[u] = (int)(j * g.sbj + b_cc[u] * g.sbc);
    }
    float acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f;
    float ra[NA], rb[NB_];
    const int sac = (int)g.sac, sbc = (int)g.sbc;
    auto fetch = [&](int c0) {
#pragma unroll
        for (int u = 0; u < NA; ++u)
            ra[u] = (a_ok[u] && c0 + a_cc[u] < c_end) ? g.A[a_off[u] + c0 * sac] : 0.f;
#pragma unroll
        for (int u = 0; u < NB_; ++u) {
            const bool in = c0 + b_cc[u] < c_end;
            rb[u] = in ? (b_ok[u] ? g.B[b_off[u] + c0 * sbc] : (b_one[u] ? 1.f : 0.f)) : 0.f;
        }
    };
    fetch(c_begin);
    for (int c0 = c_begin; c0 < c_end; c0 += BK) {
#pragma unroll
        for (int u = 0; u < NA; ++u) As[a_cc[u]][a_ii[u]] = ra[u];
#pragma unroll
        for (int u = 0; u < NB_; ++u) Bs[b_cc[u]][b_jj[u]] = rb[u];
        __syncthreads();
        if (c0 + BK < c_end) fetch(c0 + BK);          // in flight during the multiply
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][8 * ty]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][8 * ty + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][4 * tx]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                acc[a][0] = fmaf(av[a], b.x, acc[a][0]);
                acc[a][1] = fmaf(av[a], b.y, acc[a][1]);
                acc[a][2] = fmaf(av[a], b.z, acc[a][2]);
                acc[a][3] = fmaf(av[a], b.w, acc[a][3]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int i = i0 + 8 * ty + a;
        if (i >= g.I) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + 4 * tx + q;
            if (j >= gemm_je(g)) continue;
            if (g.split > 1) {
                g.part[((size_t)blockIdx.z * g.I + i) * gemm_je(g) + j] = acc[a][q];
            } else {
                float* dst = gemm_dst(g, i, j);
                const float v = gemm_epilogue(g, i, j, acc[a][q]);
                *dst = g.accumulate ? *dst + v : v;
            }
        }
    }
}

// fixed-order reduction of the split partials: eight lanes per output element, lane q sums z = q, q + 8, ... in order and
// the eight sums are combined by a shuffle tree (deterministic; eight times shorter load chains than one thread per element)
__global__ void __launch_bounds__(256) k_gemm_reduce(const GemmOp g) {
    const int je = gemm_je(g), n = g.I * je;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, q = threadIdx.x & 7;
    float s = 0.f;
    if (e < n)
        for (int z = q; z < g.split; z += 8) s += g.part[(size_t)z * n + e];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (e >= n || q != 0) return;
    const int i = e / je, j = e - i * je;
    float* dst = gemm_dst(g, i, j);
    const float v = gemm_epilogue(g, i, j, s);
    *dst = g.accumulate ? *dst + v : v;
}

int run_gemm(GemmOp g, int split, float* part, cudaStream_t st) {
    g.split = split < 1 ? 1 : split;
    g.chunk = g.K;
    g.part = part;
    if (g.split > 1) {
        g.chunk = ((g.K + g.split - 1) / g.split + BK - 1) / BK * BK;
        g.split = (g.K + g.chunk - 1) / g.chunk;
    }
    const int je = g.J + (g.ones_out != nullptr ? 1 : 0);
    OO_REQUIRE((long long)g.I * (g.sai > 0 ? g.sai : 1) + (long long)g.K * (g.sac > 0 ? g.sac : 1) < (1LL << 31) &&
                   (long long)je * (g.sbj > 0 ? g.sbj : 1) + (long long)g.K * (g.sbc > 0 ? g.sbc : 1) < (1LL << 31),
               "oo_bg gemm: operand larger than 2^31 elements");
    const dim3 grid((g.I + BI - 1) / BI, (je + BJ - 1) / BJ, g.split);
    const bool ac = g.sac == 1, bc = g.sbc == 1;
    if (ac && bc) k_gemm<true, true><<<grid, 256, 0, st>>>(g);
    else if (ac) k_gemm<true, false><<<grid, 256, 0, st>>>(g);
    else if (bc) k_gemm<false, true><<<grid, 256, 0, st>>>(g);
    else k_gemm<false, false><<<grid, 256, 0, st>>>(g);
    OO_LAUNCH_CHECK();
    if (g.split > 1) {
        k_gemm_reduce<<<(g.I * je * 8 + 255) / 256, 256, 0, st>>>(g);
        OO_LAUNCH_CHECK();
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// encoder (embedding.py:46-55): e = [t, sin(pi 2^k B t)], written straight into the three places that consume it:
// in_layer input X1 [M][ld1] (cols 0..86), cat_layer input XC [M][ldc] (cols h..h+86), head input XH [M][ldh]
// (cols h..h+41)
// ------------------------------------------------------------------------------------------------
struct EmbedBufs {
    float *x1, *xc, *xh;     // forward destinations, or the gradient sources in the backward kernel
    int ld1, ldc, ldh, h;
};

__device__ __forceinline__ void emb_store(const EmbedBufs& b, size_t p, int idx, float v) {
    if (idx < E1) {
        b.x1[p * b.ld1 + idx] = v;
        b.xc[p * b.ldc + b.h + idx] = v;
    } else {
        b.xh[p * b.ldh + b.h + idx - E1] = v;
    }
}

__global__ void k_embed_fwd(const float* __restrict__ pcs, const float* __restrict__ Bm, float scale, int n_pts,
                            const EmbedBufs b, float* __restrict__ emb_out) {
    // thread = (point, slot): slots 0..20 = directions, 21..23 = the three scaled coordinates
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t p = e / 24;
    const int d = (int)(e - p * 24);
    if (p >= (size_t)n_pts) return;
    const float t0 = pcs[3 * p] / scale, t1 = pcs[3 * p + 1] / scale, t2 = pcs[3 * p + 2] / scale;
    if (d >= NDIR) {
        const int ch = d - NDIR;
        const float t = ch == 0 ? t0 : ch == 1 ? t1 : t2;
        if (b.x1) emb_store(b, p, ch, t);
        if (emb_out) emb_out[p * EMB + ch] = t;
        return;
    }
    const float proj = Bm[3 * d] * t0 + Bm[3 * d + 1] * t1 + Bm[3 * d + 2] * t2;
    const float arg = proj * PI_F;                  // band k: fl(2^k proj pi_f) == 2^k fl(proj pi_f) exactly
    float band = 1.f;
#pragma unroll
    for (int k = 0; k < NBAND; ++k) {
        const float s = sinf(arg * band);
        if (b.x1) emb_store(b, p, 3 + NDIR * k + d, s);
        if (emb_out) emb_out[p * EMB + 3 + NDIR * k + d] = s;
        band *= 2.f;
    }
}

// d B[d][ch] = sum_p t[p][ch] * sum_k de[3+21k+d][p] * pi 2^k cos(pi 2^k proj): per-block partials [nblk][63], then a
// fixed-order reduction
constexpr int EB_PTS = 32;        // points per block
__global__ void __launch_bounds__(NDIR * 4) k_embed_bwd(const float* __restrict__ pcs, const float* __restrict__ Bm, float scale,
                                                         int n_pts, const EmbedBufs g, float* __restrict__ partial) {
    // thread = (direction d, point lane q of 4); each walks EB_PTS / 4 points
    const int d = threadIdx.x >> 2, q = threadIdx.x & 3;
    const float b0 = Bm[3 * d], b1 = Bm[3 * d + 1], b2 = Bm[3 * d + 2];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    const int p_end = min(n_pts, (int)(blockIdx.x + 1) * EB_PTS);
    for (int p = blockIdx.x * EB_PTS + q; p < p_end; p += 4) {
        const float t0 = pcs[3 * (size_t)p] / scale, t1 = pcs[3 * (size_t)p + 1] / scale, t2 = pcs[3 * (size_t)p + 2] / scale;
        const float arg = (b0 * t0 + b1 * t1 + b2 * t2) * PI_F;
        float band = 1.f, dp = 0.f;
#pragma unroll
        for (int k = 0; k < NBAND; ++k) {
            const int idx = 3 + NDIR * k + d;
            const float de = idx < E1 ? g.x1[(size_t)p * g.ld1 + idx] + g.xc[(size_t)p * g.ldc + g.h + idx]
                                      : g.xh[(size_t)p * g.ldh + g.h + idx - E1];
            dp += de * (cosf(arg * band) * (PI_F * band));
            band *= 2.f;
        }
        s0 += dp * t0; s1 += dp * t1; s2 += dp * t2;
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
    if (q == 0) {
        float* o = partial + (size_t)blockIdx.x * (NDIR * 3) + 3 * d;
        o[0] = s0; o[1] = s1; o[2] = s2;
    }
}

__global__ void __launch_bounds__(256) k_embed_bwd_reduce(const float* __restrict__ partial, int nblk, float* __restrict__ dB) {
    // one warp per output element: lane l sums blocks l, l + 32, ... in order, then a shuffle tree (fixed order)
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (e >= NDIR * 3) return;
    float s = 0.f;
    for (int b = lane; b < nblk; b += 32) s += partial[(size_t)b * (NDIR * 3) + e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) dB[e] = s;
}

// ------------------------------------------------------------------------------------------------
// small elementwise kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_sigmoid_bwd(const float* __restrict__ d, const float* __restrict__ y, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = d[i] * y[i] * (1.f - y[i]);
}

__global__ void k_gt_prepare(const uint8_t* __restrict__ rgb8, const int32_t* __restrict__ feat_row,
                             const float* __restrict__ feat_table, int n_rays, float* __restrict__ rgb,
                             float* __restrict__ feat) {
    // gt colour u8 -> float / 255 (train.py:373); gt part-feature rows gathered from the resident table (train.py:378)
    const int r = blockIdx.x;
    if (threadIdx.x < 3) rgb[3 * r + threadIdx.x] = (float)rgb8[3 * r + threadIdx.x] / 255.f;
    if (feat_row) {
        const float4* src = reinterpret_cast<const float4*>(feat_table + (size_t)feat_row[r] * C);
        float4* dst = reinterpret_cast<float4*>(feat + (size_t)r * C);
        for (int q = threadIdx.x; q < C / 4; q += blockDim.x) dst[q] = src[q];
    }
}

__global__ void k_fill(float* p, float v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// workspace map (floats)
// ------------------------------------------------------------------------------------------------
struct BgWs {
    // forward activations
    float *x1, *xc, *h1, *h3, *xh, *hc, *hp, *alpha, *color, *clip;
    // gradients
    float *d_alpha, *d_color, *d_colpre, *d_clip, *d_hc, *d_hp, *d_xh, *d_h3, *d_xc, *d_h1, *d_x1;
    float *gt_rgb, *gt_feat, *loss_ws, *ones, *part, *emb_part, *grads;
    int ld1, ldc, ldh;
    long long total;
};

constexpr int BG_SPLIT = 64;

BgWs bg_ws_map(float* base, int h, int n_pts, int n_rays) {
    BgWs w;
    w.ld1 = 88; w.ldc = (h + E1 + 3) & ~3; w.ldh = (h + E2 + 3) & ~3;
    long long o = 0;
    auto take = [&](long long n) { float* p = base ? base + o : nullptr; o += (n + 3) & ~3LL; return p; };
    const long long M = n_pts;
    w.x1 = take(M * w.ld1); w.xc = take(M * w.ldc); w.h1 = take(M * h); w.h3 = take(M * h); w.xh = take(M * w.ldh);
    w.hc = take(M * h); w.hp = take(M * h); w.alpha = take(M); w.color = take(M * 3); w.clip = take(M * C);
    w.d_alpha = take(M); w.d_color = take(M * 3); w.d_colpre = take(M * 3); w.d_clip = take(M * C);
    w.d_hc = take(M * h); w.d_hp = take(M * h); w.d_xh = take(M * w.ldh); w.d_h3 = take(M * h); w.d_xc = take(M * w.ldc);
    w.d_h1 = take(M * h); w.d_x1 = take(M * w.ld1);
    w.gt_rgb = take(3LL * n_rays); w.gt_feat = take((long long)C * n_rays);
    w.loss_ws = take((long long)n_rays * oo_loss_ws_per_ray() + 8);
    w.ones = take(4);
    const long long maxij = (long long)C * (h + 1) > (long long)h * (h + E1 + 1) ? (long long)C * (h + 1) : (long long)h * (h + E1 + 1);
    w.part = take((BG_SPLIT + 1) * maxij);
    w.emb_part = take(((M + EB_PTS - 1) / EB_PTS) * (NDIR * 3));
    w.grads = take(bg_layout(h).total);
    w.total = o;
    return w;
}

GemmOp op(const float* A, long long sai, long long sac, const float* B, long long sbj, long long sbc, float* Cp, long long sci,
          long long scj, int I, int J, int K) {
    GemmOp g = {};
    g.A = A; g.sai = sai; g.sac = sac; g.B = B; g.sbj = sbj; g.sbc = sbc; g.C = Cp; g.sci = sci; g.scj = scj;
    g.I = I; g.J = J; g.K = K; g.mult = 1.f; g.post = 1.f;
    return g;
}

#define OO_TRY(expr)              \
    do {                          \
        if (int rc__ = (expr)) return rc__; \
    } while (0)

// forward of the whole model into the workspace (model.py:61-103); returns with ws.alpha (x10 applied), ws.color
// (after sigmoid) and, if want_clip, ws.clip filled
int bg_forward(const float* th, const BgLayout& L, int h, const float* pcs, int M, float scale, const BgWs& w, bool want_clip,
               float* emb_out, cudaStream_t st) {
    const EmbedBufs eb = {w.x1, w.xc, w.xh, w.ld1, w.ldc, w.ldh, h};
    k_embed_fwd<<<(unsigned)(((size_t)M * 24 + 255) / 256), 256, 0, st>>>(pcs, th + L.off[T_PE], scale, M, eb, emb_out);
    OO_LAUNCH_CHECK();
    GemmOp g;
    // fc1 = relu(in_layer(e1))
    g = op(w.x1, w.ld1, 1, th + L.off[T_IN_W], E1, 1, w.h1, h, 1, M, h, E1); g.bias = th + L.off[T_IN_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // fc2 = relu(mid1(fc1)) -> cols 0..h-1 of the cat input
    g = op(w.h1, h, 1, th + L.off[T_M1_W], h, 1, w.xc, w.ldc, 1, M, h, h); g.bias = th + L.off[T_M1_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // fc3 = relu(cat_layer([fc2, e1]))
    g = op(w.xc, w.ldc, 1, th + L.off[T_CAT_W], h + E1, 1, w.h3, h, 1, M, h, h + E1); g.bias = th + L.off[T_CAT_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // fc4 = relu(mid2(fc3)) -> cols 0..h-1 of the head input
    g = op(w.h3, h, 1, th + L.off[T_M2_W], h, 1, w.xh, w.ldh, 1, M, h, h); g.bias = th + L.off[T_M2_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // alpha = 10 * (out_alpha(fc4))   (model.py:86-88)
    g = op(w.xh, w.ldh, 1, th + L.off[T_A_W], h, 1, w.alpha, 1, 1, M, 1, h); g.bias = th + L.off[T_A_B]; g.post = 10.f;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // color = sigmoid(out_color(relu(color_linear([fc4, e2]))))
    g = op(w.xh, w.ldh, 1, th + L.off[T_CL_W], h + E2, 1, w.hc, h, 1, M, h, h + E2); g.bias = th + L.off[T_CL_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    g = op(w.hc, h, 1, th + L.off[T_OC_W], h, 1, w.color, 3, 1, M, 3, h); g.bias = th + L.off[T_OC_B]; g.act = 2;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // clip = out_clip(relu(clip_linear([fc4, e2])))
    g = op(w.xh, w.ldh, 1, th + L.off[T_CP_W], h + E2, 1, w.hp, h, 1, M, h, h + E2); g.bias = th + L.off[T_CP_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    if (want_clip) {
        g = op(w.hp, h, 1, th + L.off[T_OCL_W], h, 1, w.clip, C, 1, M, C, h); g.bias = th + L.off[T_OCL_B];
        OO_TRY(run_gemm(g, 1, nullptr, st));
    }
    return 0;
}

}  // namespace

extern "C" int oo_bg_param_count(int hidden) { return hidden > 0 ? bg_layout(hidden).total : -1; }
extern "C" int oo_bg_param_offset(int hidden, int i) { return (hidden > 0 && i >= 0 && i < NT) ? bg_layout(hidden).off[i] : -1; }
extern "C" int oo_bg_param_size(int hidden, int i) { return (hidden > 0 && i >= 0 && i < NT) ? bg_layout(hidden).size[i] : -1; }
extern "C" int64_t oo_bg_ws_floats(int hidden, int n_pts, int n_rays) {
    if (hidden <= 0 || n_pts <= 0 || n_rays <= 0) return -1;
    return bg_ws_map(nullptr, hidden, n_pts, n_rays).total;
}

extern "C" int oo_bg_forward(const float* theta, int hidden, const float* pcs, int n_pts, float scale, float* alpha,
                             float* color, float* clip, float* emb_out, float* ws, void* stream) {
    OO_REQUIRE(theta && pcs && ws && alpha && color, "oo_bg_forward: null argument");
    OO_REQUIRE(hidden > 0 && hidden % 4 == 0 && n_pts > 0, "oo_bg_forward: hidden must be a positive multiple of 4");
    cudaStream_t st = (cudaStream_t)stream;
    const BgLayout L = bg_layout(hidden);
    const BgWs w = bg_ws_map(ws, hidden, n_pts, 1);
    OO_TRY(bg_forward(theta, L, hidden, pcs, n_pts, scale, w, clip != nullptr, emb_out, st));
    OO_CUDA(cudaMemcpyAsync(alpha, w.alpha, sizeof(float) * n_pts, cudaMemcpyDeviceToDevice, st));
    OO_CUDA(cudaMemcpyAsync(color, w.color, sizeof(float) * 3 * n_pts, cudaMemcpyDeviceToDevice, st));
    if (clip) OO_CUDA(cudaMemcpyAsync(clip, w.clip, sizeof(float) * C * (size_t)n_pts, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int oo_bg_train_step(float* theta, float* adam_m, float* adam_v, int hidden, const float* pcs, const float* z,
                                const float* gt_depth, const uint8_t* gt_rgb, const uint8_t* labels, const int32_t* feat_row,
                                const float* feat_table, int n_rays, int n_samp, float scale, int adam_step, float lr,
                                float weight_decay, float beta1, float beta2, float eps, float color_scaling,
                                float opacity_scaling, float feat_scaling, float* ws, float* terms_out, float* loss_out,
                                int* flags_out, float* grads_out, void* stream) {
    OO_REQUIRE(theta && pcs && z && gt_depth && gt_rgb && labels && ws && terms_out && loss_out && flags_out,
               "oo_bg_train_step: null argument");
    OO_REQUIRE(hidden > 0 && hidden % 4 == 0 && n_rays > 0 && n_samp > 0 && n_samp <= 32, "oo_bg_train_step: bad shape");
    OO_REQUIRE(grads_out || (adam_m && adam_v && adam_step >= 1), "oo_bg_train_step: optimiser state missing");
    OO_REQUIRE((feat_row == nullptr) == (feat_table == nullptr), "oo_bg_train_step: feat_row and feat_table go together");
    cudaStream_t st = (cudaStream_t)stream;
    const int h = hidden, M = n_rays * n_samp;
    const bool part = feat_row != nullptr;
    const BgLayout L = bg_layout(h);
    const BgWs w = bg_ws_map(ws, h, M, n_rays);
    float* G = grads_out ? grads_out : w.grads;
    k_fill<<<1, 32, 0, st>>>(w.ones, 1.f, 4);
    OO_LAUNCH_CHECK();
    k_gt_prepare<<<n_rays, 128, 0, st>>>(gt_rgb, feat_row, feat_table, n_rays, w.gt_rgb, w.gt_feat);
    OO_LAUNCH_CHECK();
    OO_TRY(bg_forward(theta, L, h, pcs, M, scale, w, part, nullptr, st));
    // ---- loss.step_batch_loss on [1, R, S] (train.py:452-462) and its gradient w.r.t. alpha / colour / clip (K3)
    OO_TRY(oo_loss_fwd(w.alpha, w.color, z, gt_depth, w.gt_rgb, labels, part ? w.clip : nullptr, part ? w.gt_feat : nullptr, 1,
                       n_rays, n_samp, part ? C : 0, color_scaling, opacity_scaling, feat_scaling, terms_out, loss_out,
                       flags_out, w.loss_ws, stream));
    OO_TRY(oo_loss_bwd(w.alpha, w.color, z, gt_depth, w.gt_rgb, labels, part ? w.clip : nullptr, part ? w.gt_feat : nullptr, 1,
                       n_rays, n_samp, part ? C : 0, color_scaling, opacity_scaling, feat_scaling, 1.f, flags_out, w.loss_ws,
                       w.d_alpha, w.d_color, part ? w.d_clip : nullptr, stream));
    const float* th = theta;
    GemmOp g;
    OO_CUDA(cudaMemsetAsync(G, 0, sizeof(float) * L.total, st));
    // ---- clip head
    if (part) {
        // d out_clip.weight [C][h] = d_clip^T hp ; bias = column sums ; d hp = (d_clip W_ocl) * [hp > 0]
        g = op(w.d_clip, 1, C, w.hp, 1, h, G + L.off[T_OCL_W], h, 1, C, h, M);
        g.ones_out = G + L.off[T_OCL_B];
        OO_TRY(run_gemm(g, BG_SPLIT, w.part, st));
        g = op(w.d_clip, C, 1, th + L.off[T_OCL_W], 1, h, w.d_hp, h, 1, M, h, C);
        g.mask = w.hp; g.smi = h; g.smj = 1; g.mask_cols = h;
        OO_TRY(run_gemm(g, 1, nullptr, st));
        g = op(w.d_hp, 1, h, w.xh, 1, w.ldh, G + L.off[T_CP_W], h + E2, 1, h, h + E2, M);
        g.ones_out = G + L.off[T_CP_B];
        OO_TRY(run_gemm(g, BG_SPLIT, w.part, st));
    }
    // ---- colour head: sigmoid backward, out_color, color_linear
    k_sigmoid_bwd<<<(3 * M + 255) / 256, 256, 0, st>>>(w.d_color, w.color, w.d_colpre, 3 * M);
    OO_LAUNCH_CHECK();
    g = op(w.d_colpre, 1, 3, w.hc, 1, h, G + L.off[T_OC_W], h, 1, 3, h, M);
    g.ones_out = G + L.off[T_OC_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part, st));
    g = op(w.d_colpre, 3, 1, th + L.off[T_OC_W], 1, h, w.d_hc, h, 1, M, h, 3);
    g.mask = w.hc; g.smi = h; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    g = op(w.d_hc, 1, h, w.xh, 1, w.ldh, G + L.off[T_CL_W], h + E2, 1, h, h + E2, M);
    g.ones_out = G + L.off[T_CL_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part, st));
    // ---- alpha head: alpha = 10 * (W_a fc4 + b_a)
    g = op(w.d_alpha, 1, 1, w.xh, 1, w.ldh, G + L.off[T_A_W], h, 1, 1, h, M); g.mult = 10.f;
    g.ones_out = G + L.off[T_A_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part, st));
    // ---- d [fc4, e2] = d_hc W_cl + d_hp W_cp + 10 d_alpha W_a (cols < h).  The ReLU mask of fc4 is a 0/1 factor, so it is
    // applied to every term as it is added: (a + b + c) m == a m + b m + c m exactly, in the same order
    g = op(w.d_hc, h, 1, th + L.off[T_CL_W], 1, h + E2, w.d_xh, w.ldh, 1, M, h + E2, h);
    g.mask = w.xh; g.smi = w.ldh; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    if (part) {
        g = op(w.d_hp, h, 1, th + L.off[T_CP_W], 1, h + E2, w.d_xh, w.ldh, 1, M, h + E2, h); g.accumulate = 1;
        g.mask = w.xh; g.smi = w.ldh; g.smj = 1; g.mask_cols = h;
        OO_TRY(run_gemm(g, 1, nullptr, st));
    }
    g = op(w.d_alpha, 1, 1, th + L.off[T_A_W], 1, h, w.d_xh, w.ldh, 1, M, h, 1); g.mult = 10.f; g.accumulate = 1;
    g.mask = w.xh; g.smi = w.ldh; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // ---- mid2: d W = d_fc4^T fc3, d fc3 = (d_fc4 W_m2) * [fc3 > 0]
    g = op(w.d_xh, 1, w.ldh, w.h3, 1, h, G + L.off[T_M2_W], h, 1, h, h, M);
    g.ones_out = G + L.off[T_M2_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part, st));
    g = op(w.d_xh, w.ldh, 1, th + L.off[T_M2_W], 1, h, w.d_h3, h, 1, M, h, h);
    g.mask = w.h3; g.smi = h; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // ---- cat_layer: d W = d_fc3^T [fc2, e1], d [fc2, e1] = d_fc3 W_cat with the ReLU mask on the fc2 columns
    g = op(w.d_h3, 1, h, w.xc, 1, w.ldc, G + L.off[T_CAT_W], h + E1, 1, h, h + E1, M);
    g.ones_out = G + L.off[T_CAT_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part, st));
    g = op(w.d_h3, h, 1, th + L.off[T_CAT_W], 1, h + E1, w.d_xc, w.ldc, 1, M, h + E1, h);
    g.mask = w.xc; g.smi = w.ldc; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // ---- mid1
    g = op(w.d_xc, 1, w.ldc, w.h1, 1, h, G + L.off[T_M1_W], h, 1, h, h, M);
    g.ones_out = G + L.off[T_M1_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part, st));
    g = op(w.d_xc, w.ldc, 1, th + L.off[T_M1_W], 1, h, w.d_h1, h, 1, M, h, h);
    g.mask = w.h1; g.smi = h; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // ---- in_layer
    g = op(w.d_h1, 1, h, w.x1, 1, w.ld1, G + L.off[T_IN_W], E1, 1, h, E1, M);
    g.ones_out = G + L.off[T_IN_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part, st));
    g = op(w.d_h1, h, 1, th + L.off[T_IN_W], 1, E1, w.d_x1, w.ld1, 1, M, E1, h);
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // ---- encoder: B_layer.weight is trainable (embedding.py:43)
    {
        const EmbedBufs gb = {w.d_x1, w.d_xc, w.d_xh, w.ld1, w.ldc, w.ldh, h};
        const int nblk = (M + EB_PTS - 1) / EB_PTS;
        k_embed_bwd<<<nblk, NDIR * 4, 0, st>>>(pcs, th + L.off[T_PE], scale, M, gb, w.emb_part);
        OO_LAUNCH_CHECK();
        k_embed_bwd_reduce<<<(NDIR * 3 + 7) / 8, 256, 0, st>>>(w.emb_part, nblk, G + L.off[T_PE]);
        OO_LAUNCH_CHECK();
    }
    if (grads_out) return 0;
    // ---- torch.optim.AdamW over the flat block; with part features off the clip head has grad None and is skipped
    // entirely (no decay either; quirk 8) -- its four tensors are contiguous in the layout
    if (part) {
        OO_TRY(oo_adamw_flat(theta, G, adam_m, adam_v, L.total, adam_step, lr, weight_decay, beta1, beta2, eps, stream));
    } else {
        const int a = L.off[T_CP_W], b = L.off[T_PE];
        OO_TRY(oo_adamw_flat(theta, G, adam_m, adam_v, a, adam_step, lr, weight_decay, beta1, beta2, eps, stream));
        OO_TRY(oo_adamw_flat(theta + b, G + b, adam_m + b, adam_v + b, L.total - b, adam_step, lr, weight_decay, beta1, beta2,
                             eps, stream));
    }
    return 0;
}
