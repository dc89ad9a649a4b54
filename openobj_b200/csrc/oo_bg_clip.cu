// Background model, part-feature term without [points x 512] tensors (a17; loss.py:82-87, render_rays.py:56-63,75-76).
//
// out_clip (hidden -> 512) is linear and its output is only consumed through the compositing sum, so -- the same legal
// restructure the fused object tile uses (oo_tile.h) -- the rendered feature of a ray is
//       x_r = sum_i T_i (W hp_i + b) = W S_r + b opac_r,      S_r = sum_i T_i hp_i,  opac_r = sum_i T_i,
// one [rays x (h + 1)] x [(h + 1) x 512] product per step instead of a [points x h] x [h x 512] one (14 samples per ray: 14 x
// fewer flops for the layer), and the backward needs only  d x_r = A_r y_r + B_r x_r  per RAY:
//       dW = sum_r d x_r S_r^T,   db = sum_r opac_r d x_r,   d S_r = W^T d x_r,   d opac_r = b . d x_r,
//       d hp_i = T_i d S_r [hp_i > 0],      dL/dT_i += hp_i . d S_r + d opac_r   (-> d alpha through the compositing backward, K3).
// The [16 800 x 512] tensors clip, d_clip and the three GEMMs on them (0.15 ms of the 0.64 ms step) disappear; K3 runs without
// features.  The kernels here are the per-ray pieces; the four small GEMMs go through run_gemm (oo_bg.cu).
#include <stdint.h>

#include "../../include/openobj_b200.h"
#include "oo_common.cuh"

namespace oo {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + expf(-x)); }
// exclusive prefix product across lanes, as K3 (oo_composite.cu): the termination weights must be the same numbers
__device__ __forceinline__ float warp_excl_prod(float f, int lane) {
    float inc = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc *= t;
    }
    const float ex = __shfl_up_sync(0xffffffffu, inc, 1);
    return lane == 0 ? 1.f : ex;
}
__device__ __forceinline__ float termination(const float* __restrict__ alpha, size_t ray, int S, int lane) {
    const bool act = lane < S;
    const float a = act ? alpha[ray * S + lane] : 0.f;
    const float o = act ? sigmoid_(a) : 0.f;                       // render_rays.py:13
    const float f = act ? (1.f - o + 1e-10f) : 1.f;                // render_rays.py:38
    return o * warp_excl_prod(f, lane);                            // render_rays.py:43
}

constexpr int HMAX = 8;      // hidden <= 256: up to 8 features per lane

// Sx[r] = [ S_r (h) | opac_r | 0 0 0 ]
__global__ void __launch_bounds__(256) k_bg_render_hp(const float* __restrict__ alpha, const float* __restrict__ hp, int n_rays, int S,
                                                      int h, int hs, float* __restrict__ Sx) {
    const int lane = threadIdx.x & 31;
    const size_t ray = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= (size_t)n_rays) return;
    const float T = termination(alpha, ray, S, lane);
    const float opac = warp_sum(T);
    float acc[HMAX];
#pragma unroll
    for (int k = 0; k < HMAX; ++k) acc[k] = 0.f;
    const float* row = hp + ray * S * h;
    for (int i = 0; i < S; ++i) {                                   // accumulation in sample order
        const float Ti = __shfl_sync(0xffffffffu, T, i);
#pragma unroll
        for (int k = 0; k < HMAX; ++k) {
            const int f = lane + 32 * k;
            if (f < h) acc[k] += Ti * row[(size_t)i * h + f];
        }
    }
    float* out = Sx + ray * hs;
#pragma unroll
    for (int k = 0; k < HMAX; ++k) {
        const int f = lane + 32 * k;
        if (f < h) out[f] = acc[k];
    }
    if (lane < hs - h) out[h + lane] = lane == 0 ? opac : 0.f;
}

// per ray: cosine loss partial and d x_r = A y + B x (zero unless the ray has label 1 and the zero-mask rule lets the term live)
// tail = K3's per-object sums (oo_composite.cu): tail[4] = number of label-1 rays
__global__ void __launch_bounds__(256) k_bg_feat_loss(const float* __restrict__ X, const float* __restrict__ gt_feat,
                                                      const uint8_t* __restrict__ labels, const int* __restrict__ flags,
                                                      const float* __restrict__ tail, int n_rays, int C, float fs,
                                                      float* __restrict__ d_x, float* __restrict__ lf) {
    const int lane = threadIdx.x & 31;
    const size_t ray = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= (size_t)n_rays) return;
    const float* x = X + ray * C;
    const float* y = gt_feat + ray * C;
    float xy = 0.f, xx = 0.f, yy = 0.f;
    for (int c4 = 4 * lane; c4 < C; c4 += 128) {
        const float4 a = *reinterpret_cast<const float4*>(x + c4), b = *reinterpret_cast<const float4*>(y + c4);
        xy += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        xx += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        yy += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    }
    xy = warp_sum(xy); xx = warp_sum(xx); yy = warp_sum(yy);
    const int lab = labels[ray], fl = flags[0];
    const float nxr = sqrtf(xx), nx = fmaxf(nxr, 1e-8f), ny = fmaxf(sqrtf(yy), 1e-8f);   // F.cosine_similarity eps clamp
    const float cosv = xy / (nx * ny);
    if (lane == 0) lf[ray] = lab == 1 ? 1.f - cosv : 0.f;                                  // render_rays.py:75-76
    float A = 0.f, B = 0.f;
    if (lab == 1 && !(fl & OO_FLAG_NO_OBJ)) {
        const float cf = fs / (tail[4] + 1e-10f);                                          // render_rays.py:108
        A = -cf / (nx * ny);
        B = nxr > 1e-8f ? cf * cosv / (nx * nx) : 0.f;
    }
    float* d = d_x + ray * C;
    for (int c4 = 4 * lane; c4 < C; c4 += 128) {
        const float4 a = *reinterpret_cast<const float4*>(x + c4), b = *reinterpret_cast<const float4*>(y + c4);
        *reinterpret_cast<float4*>(d + c4) = make_float4(A * b.x + B * a.x, A * b.y + B * a.y, A * b.z + B * a.z, A * b.w + B * a.w);
    }
}

// fixed-order sum of the per-ray partials (the order of K3's k_loss_obj_reduce) -> the feature term joins terms / loss
__global__ void __launch_bounds__(256) k_bg_feat_term(const float* __restrict__ lf, const int* __restrict__ flags,
                                                      const float* __restrict__ tail, int n_rays, float fs, float* __restrict__ terms,
                                                      float* __restrict__ loss) {
    __shared__ float sh[8];
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    float s = 0.f;
    for (int r = threadIdx.x; r < n_rays; r += blockDim.x) s += lf[r];
    s = warp_sum(s);
    if (lane == 0) sh[wv] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sh[w];
        const float q = (flags[0] & OO_FLAG_NO_OBJ) ? 0.f : t / (tail[4] + 1e-10f);
        terms[3] = q;
        loss[0] += q * fs;                                                                 // loss.py:99,101
    }
}

// hu[r][i] = hp_i . dS_r + d opac_r (what one unit of termination weight at sample i adds to the feature loss);
// d hp_i = T_i dS_r [hp_i > 0]
__global__ void __launch_bounds__(256) k_bg_feat_bwd(const float* __restrict__ alpha, const float* __restrict__ hp,
                                                     const float* __restrict__ dSx, int n_rays, int S, int h, int hs,
                                                     float* __restrict__ hu, float* __restrict__ d_hp) {
    const int lane = threadIdx.x & 31;
    const size_t ray = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= (size_t)n_rays) return;
    const float T = termination(alpha, ray, S, lane);
    const float* ds = dSx + ray * hs;
    const float dop = ds[h];
    float g[HMAX];
#pragma unroll
    for (int k = 0; k < HMAX; ++k) {
        const int f = lane + 32 * k;
        g[k] = f < h ? ds[f] : 0.f;
    }
    const float* row = hp + ray * S * h;
    float* drow = d_hp + ray * S * h;
    float mine = 0.f;
    for (int i = 0; i < S; ++i) {
        const float Ti = __shfl_sync(0xffffffffu, T, i);
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < HMAX; ++k) {
            const int f = lane + 32 * k;
            if (f < h) {
                const float v = row[(size_t)i * h + f];
                dot += v * g[k];
                drow[(size_t)i * h + f] = v > 0.f ? Ti * g[k] : 0.f;
            }
        }
        dot = warp_sum(dot);
        if (lane == i) mine = dot + dop;
    }
    if (lane < S) hu[ray * S + lane] = mine;
}

// dWb [C][hs] -> out_clip.weight gradient [C][h] and out_clip.bias gradient [C]
__global__ void k_bg_clip_scatter(const float* __restrict__ dWb, int C, int h, int hs, float* __restrict__ gW, float* __restrict__ gb) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= C * (h + 1)) return;
    const int c = e / (h + 1), j = e - c * (h + 1);
    const float v = dWb[(size_t)c * hs + j];
    if (j < h) gW[(size_t)c * h + j] = v;
    else gb[c] = v;
}

}  // namespace

int bg_clip_render(const float* alpha, const float* hp, int n_rays, int S, int h, int hs, float* Sx, cudaStream_t st) {
    OO_REQUIRE(h <= 32 * HMAX && S <= 32 && hs >= h + 1 && hs <= h + 32, "background clip head: hidden width / samples out of range");
    k_bg_render_hp<<<(n_rays + 7) / 8, 256, 0, st>>>(alpha, hp, n_rays, S, h, hs, Sx);
    OO_LAUNCH_CHECK();
    return 0;
}

int bg_clip_loss(const float* X, const float* gt_feat, const uint8_t* labels, const int* flags, const float* tail, int n_rays, int C,
                 float fs, float* d_x, float* lf, float* terms, float* loss, cudaStream_t st) {
    OO_REQUIRE((C & 3) == 0, "background clip head: feature width must be a multiple of 4");
    k_bg_feat_loss<<<(n_rays + 7) / 8, 256, 0, st>>>(X, gt_feat, labels, flags, tail, n_rays, C, fs, d_x, lf);
    OO_LAUNCH_CHECK();
    k_bg_feat_term<<<1, 256, 0, st>>>(lf, flags, tail, n_rays, fs, terms, loss);
    OO_LAUNCH_CHECK();
    return 0;
}

int bg_clip_bwd(const float* alpha, const float* hp, const float* dSx, int n_rays, int S, int h, int hs, float* hu, float* d_hp,
                cudaStream_t st) {
    k_bg_feat_bwd<<<(n_rays + 7) / 8, 256, 0, st>>>(alpha, hp, dSx, n_rays, S, h, hs, hu, d_hp);
    OO_LAUNCH_CHECK();
    return 0;
}

int bg_clip_scatter(const float* dWb, int C, int h, int hs, float* gW, float* gb, cudaStream_t st) {
    k_bg_clip_scatter<<<(C * (h + 1) + 255) / 256, 256, 0, st>>>(dWb, C, h, hs, gW, gb);
    OO_LAUNCH_CHECK();
    return 0;
}

}  // namespace oo
