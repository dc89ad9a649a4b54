// K3: occupancy -> termination compositing and the loss reductions as standalone HBM-bound kernels
// (the Python surface render_rays.* / loss.step_batch_loss operates on tensors the caller already holds).
// One warp per ray: lanes = samples, exclusive product by warp-shuffle scan, feature rendering streamed with
// 128-bit loads.  Reference: objnerf/render_rays.py:6-117, objnerf/loss.py:5-103.
#include "../../include/openobj_b200.h"
#include "oo_common.cuh"

namespace {

constexpr int CMAX = 512;                    // widest feature the workspace row holds (clip_point_feature_size)
constexpr int WS_PX = 16, WS_PY = 48, WS_X = 80;
constexpr int WSR = WS_X + CMAX;             // floats of workspace per ray
// ray_ws[ray] = {depth, var, opac, col0, col1, col2, xy, xx, yy, l_depth, l_col, l_opac, l_feat, -, -, -,
//                px[32] = pred_feat[i] . x, py[32] = pred_feat[i] . y, x[512] = rendered feature}
// The forward leaves px, py and x so that the backward never re-reads pred_feat (10 x the size of everything else):
//   dL/dpred[i][c] = T_i (A y[c] + B x[c]),   dL/dT_i (feature part) = pred[i] . (A y + B x) = A py[i] + B px[i]
// tail (after n_obj*n_rays*WSR): per object {sum_d, sum_c, sum_o, sum_f, n1, nsem, -, -}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float sgn_(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

// exclusive prefix product across lanes (lane i gets prod_{j<i} f_j); inclusive total in *total
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
// suffix sum excluding self: lane i gets sum_{k>i} v_k
__device__ __forceinline__ float warp_excl_suffix_sum(float v, int lane) {
    float inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc += t;
    }
    return inc - v;
}

struct RayIn {
    const float *alpha, *color, *z, *gt_depth, *gt_color, *pred_feat, *gt_feat;
    const uint8_t* labels;
    int n_rays_total, S, C;
};

// Feature rows reach the warp through a per-lane staging ring in shared memory (cp.async, 16 B per lane per row): the
// S x 512 B of the NEXT 128-column chunk -- of this ray or of the warp's next ray -- are in flight while the current chunk is
// reduced, so a warp keeps 2 x S x 512 B outstanding without holding a second register copy (a register double buffer
// spilled at 128 registers).  Every lane reads back only the 16 bytes it copied itself.
template <int SMAX>
constexpr int loss_fwd_stage_bytes() { return SMAX <= 16 ? 8 * 2 * (SMAX + 1) * 32 * 16 : 0; }

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int SMAX>
__global__ void __launch_bounds__(256, SMAX <= 10 ? 2 : 1) k_loss_ray_fwd(RayIn in, float* __restrict__ ws) {
    extern __shared__ float4 stage4[];                             // [8 warps][2 buffers][SMAX + 1 rows][32 lanes]
    constexpr bool STAGE = SMAX <= 16;
    const int lane = threadIdx.x & 31;
    const int S = in.S, C = in.C, nch = (C + 127) / 128;
    const int wpg = gridDim.x * (blockDim.x >> 5);
    const bool feat = in.pred_feat != nullptr;
    float4* sb = stage4 + (size_t)(threadIdx.x >> 5) * 2 * (SMAX + 1) * 32 + lane;
    // one commit group per (ray, chunk) job, empty when there is nothing to copy, so that wait_group 1 == "the current one landed"
    auto issue = [&](int ray, int ch, int b) {
        const int c4 = ch * 128 + lane * 4;
        if (ray < in.n_rays_total && c4 < C) {
            const float* pf = in.pred_feat + (size_t)ray * S * C + c4;
#pragma unroll
            for (int i = 0; i < SMAX; ++i)
                if (i < S) cp_async16(sb + (b * (SMAX + 1) + i) * 32, pf + (size_t)i * C);
            cp_async16(sb + (b * (SMAX + 1) + SMAX) * 32, in.gt_feat + (size_t)ray * C + c4);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int buf = 0;
    // persistent warps: the grid is sized to one resident wave and every warp strides over the rays (no partial last wave)
    int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (STAGE && feat) issue(ray, 0, 0);
    for (; ray < in.n_rays_total; ray += wpg) {
    const bool act = lane < S;
    const size_t pi = (size_t)ray * S + lane;
    const float a = act ? in.alpha[pi] : 0.f;
    const float zv = act ? in.z[pi] : 0.f;
    const float o = act ? sigmoid_(a) : 0.f;                       // render_rays.py:13
    const float f = act ? (1.f - o + 1e-10f) : 1.f;                // render_rays.py:38
    const float T = o * warp_excl_prod(f, lane);                   // render_rays.py:43
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (act) {
        c0 = in.color[pi * 3 + 0]; c1 = in.color[pi * 3 + 1]; c2 = in.color[pi * 3 + 2];
    }
    const float depth = warp_sum(T * zv);                          // loss.py:31
    const float dz = zv - depth;
    const float var = warp_sum(T * dz * dz);                       // loss.py:32-33
    const float opac = warp_sum(T);                                // loss.py:35
    const float r0 = warp_sum(T * c0), r1 = warp_sum(T * c1), r2 = warp_sum(T * c2);   // loss.py:34
    float xy = 0.f, xx = 0.f, yy = 0.f;
    if (feat) {                                                    // loss.py:82-87
        float px[SMAX], py[SMAX];
#pragma unroll
        for (int i = 0; i < SMAX; ++i) px[i] = py[i] = 0.f;
        for (int ch = 0; ch < nch; ++ch) {
            const int c4 = ch * 128 + lane * 4;
            float4 p[SMAX], y = {0.f, 0.f, 0.f, 0.f};
            if (STAGE) {
                if (ch + 1 < nch) issue(ray, ch + 1, buf ^ 1);
                else issue(ray + wpg, 0, buf ^ 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                if (c4 < C) {
#pragma unroll
                    for (int i = 0; i < SMAX; ++i)
                        if (i < S) p[i] = sb[(buf * (SMAX + 1) + i) * 32];
                    y = sb[(buf * (SMAX + 1) + SMAX) * 32];
                }
                buf ^= 1;
            } else if (c4 < C) {
                // all S rows of this column chunk are requested before the first is used
                const float* pf = in.pred_feat + (size_t)ray * S * C + c4;
#pragma unroll
                for (int i = 0; i < SMAX; ++i)
                    if (i < S) p[i] = *reinterpret_cast<const float4*>(pf + (size_t)i * C);
                y = *reinterpret_cast<const float4*>(in.gt_feat + (size_t)ray * C + c4);
            }
            if (c4 < C) {
                // accumulation stays in sample order
                float4 x = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < SMAX; ++i) {
                    if (i < S) {
                        const float Ti = __shfl_sync(0xffffffffu, T, i);
                        x.x += Ti * p[i].x; x.y += Ti * p[i].y; x.z += Ti * p[i].z; x.w += Ti * p[i].w;
                    }
                }
                xy += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
                xx += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
                yy += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
#pragma unroll
                for (int i = 0; i < SMAX; ++i) {
                    if (i < S) {
                        px[i] += p[i].x * x.x + p[i].y * x.y + p[i].z * x.z + p[i].w * x.w;
                        py[i] += p[i].x * y.x + p[i].y * y.y + p[i].z * y.z + p[i].w * y.w;
                    }
                }
                *reinterpret_cast<float4*>(ws + (size_t)ray * WSR + WS_X + c4) = x;
            }
        }
#pragma unroll
        for (int i = 0; i < SMAX; ++i) {
            if (i < S) {
                const float a_ = warp_sum(px[i]), b_ = warp_sum(py[i]);
                if (lane == i) { ws[(size_t)ray * WSR + WS_PX + i] = a_; ws[(size_t)ray * WSR + WS_PY + i] = b_; }
            }
        }
        xy = warp_sum(xy); xx = warp_sum(xx); yy = warp_sum(yy);
    }
    if (lane == 0) {
        const int lab = in.labels[ray];
        const bool is1 = lab == 1, sem = lab != 2;
        float ld = 0.f, lc = 0.f, lo = 0.f, lf = 0.f;
        if (is1) {
            const float w = 1.f / (sqrtf(var) + 1e-4f);            // render_rays.py:95-100
            ld = fabsf(depth - in.gt_depth[ray]) * w;
            lc = fabsf(r0 - in.gt_color[ray * 3 + 0]) + fabsf(r1 - in.gt_color[ray * 3 + 1]) +
                 fabsf(r2 - in.gt_color[ray * 3 + 2]);             // loss.py:61
            if (in.pred_feat != nullptr) {
                const float nx = fmaxf(sqrtf(xx), 1e-8f), ny = fmaxf(sqrtf(yy), 1e-8f);
                lf = 1.f - xy / (nx * ny);                         // render_rays.py:75-76
            }
        }
        if (sem) lo = fabsf(opac - (lab != 0 ? 1.f : 0.f));       // loss.py:71-73
        float* w = ws + (size_t)ray * WSR;
        w[0] = depth; w[1] = var; w[2] = opac; w[3] = r0; w[4] = r1; w[5] = r2; w[6] = xy; w[7] = xx; w[8] = yy;
        w[9] = ld; w[10] = lc; w[11] = lo; w[12] = lf;
    }
    }
}

// one block per object: fixed-order reduction of the per-ray losses and the mask counts
__global__ void __launch_bounds__(256) k_loss_obj_reduce(const float* __restrict__ ws, const uint8_t* __restrict__ labels,
                                                         int n_rays, float* __restrict__ tail) {
    __shared__ float sh[8][6];
    const int o = blockIdx.x, lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = threadIdx.x; r < n_rays; r += blockDim.x) {
        const size_t ray = (size_t)o * n_rays + r;
        const float* w = ws + ray * WSR;
        s[0] += w[9]; s[1] += w[10]; s[2] += w[11]; s[3] += w[12];
        const int lab = labels[ray];
        s[4] += lab == 1 ? 1.f : 0.f;
        s[5] += lab != 2 ? 1.f : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) s[k] = warp_sum(s[k]);
    if (lane == 0)
        for (int k = 0; k < 6; ++k) sh[wv][k] = s[k];
    __syncthreads();
    if (threadIdx.x < 6) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w][threadIdx.x];
        tail[o * 8 + threadIdx.x] = t;
    }
}

__global__ void k_loss_finalize(const float* __restrict__ tail, int n_obj, int has_feat, float cs, float os, float fs,
                                float* __restrict__ terms, float* __restrict__ loss, int* __restrict__ flags) {
    __shared__ int sflags;
    if (threadIdx.x == 0) sflags = 0;
    __syncthreads();
    int f = 0;
    for (int o = threadIdx.x; o < n_obj; o += blockDim.x) {
        if (tail[o * 8 + 4] == 0.f) f |= OO_FLAG_NO_OBJ;           // render_rays.py:88-89
        if (tail[o * 8 + 5] == 0.f) f |= OO_FLAG_NO_SEM;
    }
    if (f) atomicOr(&sflags, f);
    __syncthreads();
    f = sflags;
    __syncthreads();
    int ex = 0;
    for (int o = threadIdx.x; o < n_obj; o += blockDim.x) {
        const float n1 = tail[o * 8 + 4] + 1e-10f, ns = tail[o * 8 + 5] + 1e-10f;   // render_rays.py:108
        const float d = (f & OO_FLAG_NO_OBJ) ? 0.f : tail[o * 8 + 0] / n1;
        const float c = (f & OO_FLAG_NO_OBJ) ? 0.f : tail[o * 8 + 1] / n1;
        const float p = (f & OO_FLAG_NO_SEM) ? 0.f : tail[o * 8 + 2] / ns;
        const float q = ((f & OO_FLAG_NO_OBJ) || !has_feat) ? 0.f : tail[o * 8 + 3] / n1;
        terms[o * 4 + 0] = d; terms[o * 4 + 1] = c; terms[o * 4 + 2] = p; terms[o * 4 + 3] = q;
        if (d > 100000.f || c > 100000.f || p > 100000.f || q > 100000.f) ex = OO_FLAG_EXPLODE;   // render_rays.py:109
    }
    if (ex) atomicOr(&sflags, ex);
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int o = 0; o < n_obj; ++o)                            // loss.py:79,99,101
            tot += terms[o * 4 + 0] + terms[o * 4 + 1] * cs + terms[o * 4 + 2] * os + terms[o * 4 + 3] * fs;
        loss[0] = tot;
        flags[0] = sflags;
    }
}

template <int SMAX>
__global__ void __launch_bounds__(256, 4) k_loss_ray_bwd(RayIn in, const float* __restrict__ ws, const float* __restrict__ tail,
                                                      const int* __restrict__ flags, int n_rays, float cs, float os,
                                                      float fs, float gl, float* __restrict__ d_alpha,
                                                      float* __restrict__ d_color, float* __restrict__ d_pred,
                                                      const float* __restrict__ hu_extra) {
    const int lane = threadIdx.x & 31;
    for (int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ray < in.n_rays_total; ray += gridDim.x * (blockDim.x >> 5)) {
    const int S = in.S, obj = ray / n_rays;
    const int fl = flags[0];
    const bool act = lane < S;
    const size_t pi = (size_t)ray * S + lane;
    const float a = act ? in.alpha[pi] : 0.f;
    const float zv = act ? in.z[pi] : 0.f;
    const float o = act ? sigmoid_(a) : 0.f;
    const float f = act ? (1.f - o + 1e-10f) : 1.f;
    const float P = warp_excl_prod(f, lane);
    const float T = o * P;
    const float* w = ws + (size_t)ray * WSR;
    const float depth = w[0], var = w[1], opac = w[2];
    const int lab = in.labels[ray];
    const bool is1 = lab == 1 && !(fl & OO_FLAG_NO_OBJ), sem = lab != 2 && !(fl & OO_FLAG_NO_SEM);
    const float inv1 = gl / (tail[obj * 8 + 4] + 1e-10f), invs = gl / (tail[obj * 8 + 5] + 1e-10f);
    float gd = 0.f, go = 0.f, gc0 = 0.f, gc1 = 0.f, gc2 = 0.f, A = 0.f, B = 0.f;
    if (is1) {
        gd = sgn_(depth - in.gt_depth[ray]) / (sqrtf(var) + 1e-4f) * inv1;
        gc0 = sgn_(w[3] - in.gt_color[ray * 3 + 0]) * cs * inv1;
        gc1 = sgn_(w[4] - in.gt_color[ray * 3 + 1]) * cs * inv1;
        gc2 = sgn_(w[5] - in.gt_color[ray * 3 + 2]) * cs * inv1;
        if (in.pred_feat != nullptr) {
            const float nxr = sqrtf(w[7]), nx = fmaxf(nxr, 1e-8f), ny = fmaxf(sqrtf(w[8]), 1e-8f);
            const float cosv = w[6] / (nx * ny), cf = fs * inv1;
            A = -cf / (nx * ny);
            B = nxr > 1e-8f ? cf * cosv / (nx * nx) : 0.f;
        }
    }
    if (sem) go = sgn_(opac - (lab != 0 ? 1.f : 0.f)) * os * invs;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (act) {
        c0 = in.color[pi * 3 + 0]; c1 = in.color[pi * 3 + 1]; c2 = in.color[pi * 3 + 2];
    }
    float hu = 0.f;                                                // lane i: pred_feat[i] . dL/dx = A py[i] + B px[i]
    if (in.pred_feat != nullptr) {
        const float* gy = in.gt_feat + (size_t)ray * in.C;
        const float* xr = w + WS_X;
        float* dp = d_pred + (size_t)ray * S * in.C;
        if (act) hu = A * w[WS_PY + lane] + B * w[WS_PX + lane];
        for (int c4 = lane * 4; c4 < in.C; c4 += 128) {
            const float4 y = *reinterpret_cast<const float4*>(gy + c4);
            const float4 x = *reinterpret_cast<const float4*>(xr + c4);
            float4 g;
            g.x = A * y.x + B * x.x; g.y = A * y.y + B * x.y; g.z = A * y.z + B * x.z; g.w = A * y.w + B * x.w;
#pragma unroll
            for (int i = 0; i < SMAX; ++i) {
                if (i < S) {
                    const float Ti = __shfl_sync(0xffffffffu, T, i);
                    float4 d;
                    d.x = Ti * g.x; d.y = Ti * g.y; d.z = Ti * g.z; d.w = Ti * g.w;
                    *reinterpret_cast<float4*>(dp + (size_t)i * in.C + c4) = d;
                }
            }
        }
    }
    // hu_extra: dL/dT contributions formed outside (the background's factored clip head, oo_bg_clip.cu)
    if (hu_extra != nullptr && act) hu += hu_extra[pi];
    const float g = zv * gd + go + c0 * gc0 + c1 * gc1 + c2 * gc2 + hu;    // dL/dT_lane
    const float suffix = warp_excl_suffix_sum(act ? g * T : 0.f, lane);
    if (act) {
        const float docc = g * P - suffix / f;
        d_alpha[pi] = docc * o * (1.f - o);
        d_color[pi * 3 + 0] = T * gc0;
        d_color[pi * 3 + 1] = T * gc1;
        d_color[pi * 3 + 2] = T * gc2;
    }
    }
}

// grid of one resident wave (occupancy x SM count), capped by the work
template <typename K>
int wave_blocks(K kernel, int needed, size_t smem = 0) {
    int dev = 0, n_sm = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, smem);
    const int wave = n_sm * (occ < 1 ? 1 : occ);
    return needed < wave ? needed : wave;
}

int check_loss_args(const void* a, const void* c, const void* z, const void* gd, const void* gc, const void* lab,
                    const void* pf, const void* gf, int n_obj, int n_rays, int S, int C) {
    OO_REQUIRE(a && c && z && gd && gc && lab, "oo_loss: null input");
    OO_REQUIRE((pf == nullptr) == (gf == nullptr), "oo_loss: pred_feat and gt_feat must both be given or both be NULL");
    OO_REQUIRE(n_obj > 0 && n_rays > 0 && S > 0 && S <= 32, "oo_loss: need 0 < n_samp <= 32");
    OO_REQUIRE(pf == nullptr || (C > 0 && C % 4 == 0 && C <= CMAX), "oo_loss: feature width must be a multiple of 4, at most 512");
    return 0;
}

}  // namespace

extern "C" int oo_loss_ws_per_ray(void) { return WSR; }

extern "C" int oo_loss_fwd(const float* alpha, const float* color, const float* z, const float* gt_depth,
                           const float* gt_color, const uint8_t* labels, const float* pred_feat, const float* gt_feat,
                           int n_obj, int n_rays, int n_samp, int n_feat, float cs, float os, float fs, float* terms_out,
                           float* loss_out, int* flags_out, float* ray_ws, void* stream) {
    if (int rc = check_loss_args(alpha, color, z, gt_depth, gt_color, labels, pred_feat, gt_feat, n_obj, n_rays, n_samp,
                                 n_feat))
        return rc;
    OO_REQUIRE(terms_out && loss_out && flags_out && ray_ws, "oo_loss_fwd: null output");
    cudaStream_t st = (cudaStream_t)stream;
    RayIn in{alpha, color, z, gt_depth, gt_color, pred_feat, gt_feat, labels, n_obj * n_rays, n_samp, n_feat};
    float* tail = ray_ws + (size_t)n_obj * n_rays * WSR;
    const int fblocks = (in.n_rays_total + 7) / 8;
    // the staging ring is only needed (and only paid for in occupancy) when features are rendered
#define OO_LAUNCH_FWD(SM_)                                                                                                   \
    do {                                                                                                                     \
        const size_t sm = pred_feat != nullptr ? (size_t)loss_fwd_stage_bytes<SM_>() : 0;                                    \
        OO_CUDA(cudaFuncSetAttribute(k_loss_ray_fwd<SM_>, cudaFuncAttributeMaxDynamicSharedMemorySize,                       \
                                     loss_fwd_stage_bytes<SM_>()));                                                          \
        k_loss_ray_fwd<SM_><<<wave_blocks(k_loss_ray_fwd<SM_>, fblocks, sm), 256, sm, st>>>(in, ray_ws);                     \
    } while (0)
    if (n_samp <= 10) OO_LAUNCH_FWD(10);
    else if (n_samp <= 16) OO_LAUNCH_FWD(16);
    else OO_LAUNCH_FWD(32);
#undef OO_LAUNCH_FWD
    OO_LAUNCH_CHECK();
    k_loss_obj_reduce<<<n_obj, 256, 0, st>>>(ray_ws, labels, n_rays, tail);
    OO_LAUNCH_CHECK();
    k_loss_finalize<<<1, 256, 0, st>>>(tail, n_obj, pred_feat != nullptr, cs, os, fs, terms_out, loss_out, flags_out);
    OO_LAUNCH_CHECK();
    return 0;
}

// oo_loss_bwd with an extra per-sample dL/dT input (nullptr = none); internal: the C ABI entry below passes nullptr
namespace oo {
int loss_bwd_hu(const float* alpha, const float* color, const float* z, const float* gt_depth,
                           const float* gt_color, const uint8_t* labels, const float* pred_feat, const float* gt_feat,
                           int n_obj, int n_rays, int n_samp, int n_feat, float cs, float os, float fs, float grad_loss,
                           const int* flags, const float* ray_ws, float* d_alpha, float* d_color, float* d_pred_feat,
                           const float* hu_extra, void* stream) {
    if (int rc = check_loss_args(alpha, color, z, gt_depth, gt_color, labels, pred_feat, gt_feat, n_obj, n_rays, n_samp,
                                 n_feat))
        return rc;
    OO_REQUIRE(flags && ray_ws && d_alpha && d_color, "oo_loss_bwd: null argument");
    OO_REQUIRE((pred_feat == nullptr) == (d_pred_feat == nullptr), "oo_loss_bwd: d_pred_feat must match pred_feat");
    cudaStream_t st = (cudaStream_t)stream;
    RayIn in{alpha, color, z, gt_depth, gt_color, pred_feat, gt_feat, labels, n_obj * n_rays, n_samp, n_feat};
    const float* tail = ray_ws + (size_t)n_obj * n_rays * WSR;
    const int blocks = (in.n_rays_total + 7) / 8;
    if (n_samp <= 10)
        k_loss_ray_bwd<10><<<wave_blocks(k_loss_ray_bwd<10>, blocks), 256, 0, st>>>(in, ray_ws, tail, flags, n_rays, cs, os, fs, grad_loss, d_alpha, d_color,
                                                   d_pred_feat, hu_extra);
    else if (n_samp <= 16)
        k_loss_ray_bwd<16><<<wave_blocks(k_loss_ray_bwd<16>, blocks), 256, 0, st>>>(in, ray_ws, tail, flags, n_rays, cs, os, fs, grad_loss, d_alpha, d_color,
                                                   d_pred_feat, hu_extra);
    else
        k_loss_ray_bwd<32><<<wave_blocks(k_loss_ray_bwd<32>, blocks), 256, 0, st>>>(in, ray_ws, tail, flags, n_rays, cs, os, fs, grad_loss, d_alpha, d_color,
                                                   d_pred_feat, hu_extra);
    OO_LAUNCH_CHECK();
    return 0;
}
}  // namespace oo

extern "C" int oo_loss_bwd(const float* alpha, const float* color, const float* z, const float* gt_depth,
                           const float* gt_color, const uint8_t* labels, const float* pred_feat, const float* gt_feat,
                           int n_obj, int n_rays, int n_samp, int n_feat, float cs, float os, float fs, float grad_loss,
                           const int* flags, const float* ray_ws, float* d_alpha, float* d_color, float* d_pred_feat,
                           void* stream) {
    return oo::loss_bwd_hu(alpha, color, z, gt_depth, gt_color, labels, pred_feat, gt_feat, n_obj, n_rays, n_samp, n_feat, cs, os, fs,
                           grad_loss, flags, ray_ws, d_alpha, d_color, d_pred_feat, nullptr, stream);
}
