// General-upstream backward of the ensemble forward: what autograd does for the reference's
//     emb = vmap(pe_model)(pe_param, pe_buffer, pcs);  alpha, color, clip = vmap(fc_model)(fc_param, fc_buffer, emb)
// (objnerf/train.py:424-425, backward at :472) when the caller composes its own loss on (alpha, color, clip) instead of going
// through the fused training step.  oo_forward_bwd recomputes the forward of a 100-point tile (phases 2..7 of the fused tile),
// takes dL/d(alpha, color, clip) from the caller, and runs the tile's own backward phases 20..31; the 512-wide out_clip
// layer is handled by two plain contraction kernels around it (d hp = d_clip W, d W = d_clip^T hp).  oo_embed_bwd is the
// backward of UniDirsEmbed.forward (embedding.py:46-55) for the trainable direction matrix.
#include "../../include/openobj_b200.h"
#include "oo_common.cuh"
#include "oo_tile.h"

using namespace oo;

namespace {

template <int PH, int END>
struct RunPhases {
    static __device__ __forceinline__ void run(int tid, float* sm, const TileCtx& c, TileAcc& a) {
        tile_phase<PH, true>(tid, sm, c, a);
        __syncthreads();
        RunPhases<PH + 1, END>::run(tid, sm, c, a);
    }
};
template <int END>
struct RunPhases<END, END> {
    static __device__ __forceinline__ void run(int, float*, const TileCtx&, TileAcc&) {}
};

// ---- d hp[p][j] = sum_c d_clip[p][c] W_ocl[c][j]  (M x 512 by 512 x 32 per object).
// grid (ceil(n_pts / 64), n_obj), 256 threads: thread (point = tid / 4, jq = tid % 4) owns 8 outputs; d_clip and W travel
// through shared memory in 64-column chunks.
constexpr int CB_PTS = 64, CB_CH = 64, CB_STR = CB_CH + 1;
__global__ void __launch_bounds__(256) k_clip_dhp(const float* __restrict__ theta, const float* __restrict__ d_clip, int n_pts,
                                                  float* __restrict__ d_hp) {
    __shared__ float ds[CB_PTS * CB_STR];
    __shared__ __align__(16) float ws[CB_CH * H];
    const int obj = blockIdx.y, p0 = blockIdx.x * CB_PTS, tid = threadIdx.x;
    const int pl = tid >> 2, jq = tid & 3;
    const float* W = theta + (size_t)obj * PSTRIDE + OFF_OCL_W;
    const float* D = d_clip + ((size_t)obj * n_pts + p0) * C;
    const int np = min(CB_PTS, n_pts - p0);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < C; c0 += CB_CH) {
        __syncthreads();
        for (int i = tid; i < CB_PTS * CB_CH; i += 256) {
            const int p = i / CB_CH, cc = i - p * CB_CH;
            ds[p * CB_STR + cc] = p < np ? D[(size_t)p * C + c0 + cc] : 0.f;
        }
        for (int i = tid; i < CB_CH * H / 4; i += 256)
            reinterpret_cast<float4*>(ws)[i] = reinterpret_cast<const float4*>(W + (size_t)c0 * H)[i];
        __syncthreads();
#pragma unroll 8
        for (int cc = 0; cc < CB_CH; ++cc) {
            const float a = ds[pl * CB_STR + cc];
            const float4 w0 = *reinterpret_cast<const float4*>(ws + cc * H + 8 * jq);
            const float4 w1 = *reinterpret_cast<const float4*>(ws + cc * H + 8 * jq + 4);
            acc[0] += a * w0.x; acc[1] += a * w0.y; acc[2] += a * w0.z; acc[3] += a * w0.w;
            acc[4] += a * w1.x; acc[5] += a * w1.y; acc[6] += a * w1.z; acc[7] += a * w1.w;
        }
    }
    if (pl < np) {
        float* o = d_hp + ((size_t)obj * n_pts + p0 + pl) * H + 8 * jq;
        *reinterpret_cast<float4*>(o) = float4{acc[0], acc[1], acc[2], acc[3]};
        *reinterpret_cast<float4*>(o + 4) = float4{acc[4], acc[5], acc[6], acc[7]};
    }
}

// ---- d W_ocl[c][j] = sum_p d_clip[p][c] hp[p][j],  d b_ocl[c] = sum_p d_clip[p][c]; written straight into the gradient
// block.  grid (512 / 64, n_obj), 256 threads: thread (column = tid / 4, jq = tid % 4) owns 8 outputs and walks ALL points in a
// fixed order (deterministic).
__global__ void __launch_bounds__(256) k_clip_dw(const float* __restrict__ d_clip, const float* __restrict__ hp, int n_pts,
                                                 float* __restrict__ grads) {
    __shared__ float ds[CB_PTS * CB_STR];
    __shared__ __align__(16) float hs[CB_PTS * H];
    const int obj = blockIdx.y, c0 = blockIdx.x * CB_CH, tid = threadIdx.x;
    const int cl = tid >> 2, jq = tid & 3;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, accb = 0.f;
    for (int p0 = 0; p0 < n_pts; p0 += CB_PTS) {
        const int np = min(CB_PTS, n_pts - p0);
        const float* D = d_clip + ((size_t)obj * n_pts + p0) * C + c0;
        const float* Hh = hp + ((size_t)obj * n_pts + p0) * H;
        __syncthreads();
        for (int i = tid; i < CB_PTS * CB_CH; i += 256) {
            const int p = i / CB_CH, cc = i - p * CB_CH;
            ds[p * CB_STR + cc] = p < np ? D[(size_t)p * C + cc] : 0.f;
        }
        for (int i = tid; i < CB_PTS * H; i += 256) hs[i] = i < np * H ? Hh[i] : 0.f;
        __syncthreads();
#pragma unroll 8
        for (int p = 0; p < CB_PTS; ++p) {
            const float a = ds[p * CB_STR + cl];
            const float4 h0 = *reinterpret_cast<const float4*>(hs + p * H + 8 * jq);
            const float4 h1 = *reinterpret_cast<const float4*>(hs + p * H + 8 * jq + 4);
            acc[0] += a * h0.x; acc[1] += a * h0.y; acc[2] += a * h0.z; acc[3] += a * h0.w;
            acc[4] += a * h1.x; acc[5] += a * h1.y; acc[6] += a * h1.z; acc[7] += a * h1.w;
            accb += a;
        }
    }
    float* g = grads + (size_t)obj * PSTRIDE;
    float* o = g + OFF_OCL_W + (size_t)(c0 + cl) * H + 8 * jq;
    *reinterpret_cast<float4*>(o) = float4{acc[0], acc[1], acc[2], acc[3]};
    *reinterpret_cast<float4*>(o + 4) = float4{acc[4], acc[5], acc[6], acc[7]};
    if (jq == 0) g[OFF_OCL_B + c0 + cl] = accb;
}

// ---- the tile kernel: forward recomputation + backward from the caller's upstream gradients.
// grid (gx, n_obj): CTA (x, obj) walks tiles x, x + gx, ... of its object and leaves its weight-gradient partial in slab slot
// obj * gx + x (summed in slot order by k_reduce_slots: deterministic).
struct BwdParams {
    const float* theta;
    const float* emb_in;       // [n_obj][n_pts][129]
    const float* d_alpha;      // [n_obj][n_pts]
    const float* d_color;      // [n_obj][n_pts][3]
    const float* d_hp;         // [n_obj][n_pts][32] or nullptr (no clip gradient)
    float* hp_out;             // [n_obj][n_pts][32] or nullptr
    float* d_emb;              // [n_obj][n_pts][129] or nullptr
    float* slab;               // [n_obj * gx][PSTRIDE]
    float* slot_loss;          // [n_obj * gx][4] (unused partial sums of the shared flush code)
    int n_pts;
};

__global__ void __launch_bounds__(NTHREADS, 1) k_fwdbwd(const BwdParams prm) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, obj = blockIdx.y;
    const float* th = prm.theta + (size_t)obj * PSTRIDE;
    zero_pad_rows(tid, sm);
    for (int i = tid; i < 3 * PS; i += NTHREADS) sm[SM_ACT + R_T * PS + i] = 0.f;      // no encoder here: the PE rows stay zero
    __shared__ uint32_t tm_base_s;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tm_base_s)), "r"((uint32_t)(AC_COLS * NWARPS / 4)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_base = tm_base_s;
    TileAcc acc;
    acc.tm = tm_base + ((uint32_t)(32 * ((tid >> 5) & 3)) << 16) + (uint32_t)(AC_COLS * (tid >> 7));
    acc_zero(acc);
    stage_weights(tid, sm, th);
    __syncthreads();
    TileCtx c = {};
    c.theta = th;
    c.scale = 1.f;
    c.nrays = 0;
    float* act = sm + SM_ACT;
    const int n_tiles = (prm.n_pts + P - 1) / P;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int p0 = t * P;
        c.npts = min(P, prm.n_pts - p0);
        const size_t base = (size_t)obj * prm.n_pts + p0;
        for (int i = tid; i < P * EMB; i += NTHREADS) {
            const int p = i / EMB, e = i - p * EMB;
            const float v = p < c.npts ? prm.emb_in[base * EMB + i] : 0.f;
            if (e < E1) act[(R_E1 + e) * PS + p] = v;
            else act[(R_E2 + e - E1) * PS + p] = v;
        }
        __syncthreads();
        RunPhases<2, N_FWD_PHASES>::run(tid, sm, c, acc);                 // forward: in, mid1, cat, mid2, heads, alpha / colour
        c.up_alpha = prm.d_alpha + base;
        c.up_color = prm.d_color + base * 3;
        c.up_hp = prm.d_hp ? prm.d_hp + base * H : nullptr;
        c.hp_out = prm.hp_out ? prm.hp_out + base * H : nullptr;
        RunPhases<43, 45>::run(tid, sm, c, acc);                          // upstream -> raw head gradients; out_color / out_alpha
        RunPhases<20, 30>::run(tid, sm, c, acc);                          // weight / data gradients down to d e1
        RunPhases<31, 32>::run(tid, sm, c, acc);                          // bias gradients (the PE part sums zero rows)
        if (prm.d_emb != nullptr) {
            for (int i = tid; i < c.npts * EMB; i += NTHREADS) {
                const int p = i / EMB, e = i - p * EMB;
                prm.d_emb[base * EMB + i] = e < E1 ? act[(R_E1 + e) * PS + p] : act[(R_E2 + e - E1) * PS + p];
            }
        }
        __syncthreads();
    }
    const int slot = obj * gridDim.x + blockIdx.x;
    tile_flush<0, true>(tid, sm, prm.slab + (size_t)slot * PSTRIDE, prm.slot_loss + 4 * slot, acc); __syncthreads();
    tile_flush<1, true>(tid, sm, prm.slab + (size_t)slot * PSTRIDE, prm.slot_loss + 4 * slot, acc); __syncthreads();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"((uint32_t)(AC_COLS * NWARPS / 4))
                     : "memory");
}

// grads[o][i] = sum over the object's slots (fixed order); the out_clip region belongs to k_clip_dw (or is zero without a clip
// gradient), the PE directions get no gradient here (oo_embed_bwd), the pad is zero.
__global__ void __launch_bounds__(256) k_reduce_slots(const float* __restrict__ slab, int gx, int have_clip, float* __restrict__ grads) {
    const int o = blockIdx.y;
    const int i = 4 * (blockIdx.x * 256 + threadIdx.x);
    if (i >= PSTRIDE) return;
    float4 g = {0.f, 0.f, 0.f, 0.f};
    const bool clip_region = i >= OFF_OCL_W && i < OFF_PE_B;
    if (clip_region && have_clip) return;
    if (!clip_region && i < OFF_PE_B) {
        for (int s = 0; s < gx; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(slab + ((size_t)o * gx + s) * PSTRIDE + i);
            g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
        }
    }
    *reinterpret_cast<float4*>(grads + (size_t)o * PSTRIDE + i) = g;
}

// ---- UniDirsEmbed backward: d B[d][ch] = sum_p t[p][ch] * sum_k d_emb[p][3 + 21 k + d] * pi 2^k cos(pi 2^k proj[d][p]).
// grid (n_chunks, n_obj), 21 warps: warp d, lanes stride over the chunk's points; same angle-doubling recurrence as the
// forward encoder (oo_tile.h phase 1 / 30).  Partials [n_obj][n_chunks][63] -> k_embed_reduce (fixed order).
constexpr int EB_CHUNK = 256;
__global__ void __launch_bounds__(NDIR * 32) k_embed_bwd_ens(const float* __restrict__ theta, const float* __restrict__ pcs,
                                                              const float* __restrict__ d_emb, int n_pts, float scale,
                                                              float* __restrict__ partial) {
    const int obj = blockIdx.y, d = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* Bm = theta + (size_t)obj * PSTRIDE + OFF_PE_B + 3 * d;
    const float b0 = Bm[0], b1 = Bm[1], b2 = Bm[2];
    const int p_end = min(n_pts, (int)(blockIdx.x + 1) * EB_CHUNK);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int p = blockIdx.x * EB_CHUNK + lane; p < p_end; p += 32) {
        const size_t q = (size_t)obj * n_pts + p;
        const float t0 = pcs[3 * q] / scale, t1 = pcs[3 * q + 1] / scale, t2 = pcs[3 * q + 2] / scale;
        const float proj = b0 * t0 + b1 * t1 + b2 * t2;
        float sn, cs;
        sincosf(proj * PI_F, &sn, &cs);
        float band = PI_F, dp = 0.f;
#pragma unroll
        for (int k = 0; k < NBAND; ++k) {
            dp += d_emb[q * EMB + 3 + NDIR * k + d] * (cs * band);
            const float s2x = 2.f * sn * cs, c2x = (cs - sn) * (cs + sn);
            sn = s2x; cs = c2x;
            band *= 2.f;
        }
        s0 += dp * t0; s1 += dp * t1; s2 += dp * t2;
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) {
        float* o = partial + ((size_t)obj * gridDim.x + blockIdx.x) * (NDIR * 3) + 3 * d;
        o[0] = s0; o[1] = s1; o[2] = s2;
    }
}
__global__ void k_embed_reduce(const float* __restrict__ partial, int n_chunks, float* __restrict__ d_B) {
    const int obj = blockIdx.x, e = threadIdx.x;
    if (e >= NDIR * 3) return;
    float s = 0.f;
    for (int ch = 0; ch < n_chunks; ++ch) s += partial[((size_t)obj * n_chunks + ch) * (NDIR * 3) + e];
    d_B[(size_t)obj * (NDIR * 3) + e] = s;
}

int bwd_gx(int n_obj, int n_pts) {
    const int n_tiles = (n_pts + P - 1) / P;
    int gx = (2 * 148 + n_obj - 1) / n_obj;
    if (gx > n_tiles) gx = n_tiles;
    if (gx > 32) gx = 32;
    if (gx < 1) gx = 1;
    return gx;
}

}  // namespace

extern "C" int64_t oo_forward_bwd_ws_floats(int n_obj, int n_pts) {
    if (n_obj <= 0 || n_pts <= 0) return -1;
    const int gx = bwd_gx(n_obj, n_pts);
    return (int64_t)n_obj * gx * (PSTRIDE + 4) + 2 * (int64_t)n_obj * n_pts * H;
}

extern "C" int oo_forward_bwd(const float* theta, int n_obj, const float* emb_in, int n_pts, const float* d_alpha,
                              const float* d_color, const float* d_clip, float* grads_out, float* d_emb_out, float* ws,
                              void* stream) {
    OO_REQUIRE(theta && emb_in && d_alpha && d_color && grads_out && ws, "oo_forward_bwd: null argument");
    OO_REQUIRE(n_obj > 0 && n_pts > 0, "oo_forward_bwd: empty input");
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = bwd_gx(n_obj, n_pts);
    float* slab = ws;
    float* slot_loss = slab + (size_t)n_obj * gx * PSTRIDE;
    float* d_hp = slot_loss + (size_t)n_obj * gx * 4;
    float* hp = d_hp + (size_t)n_obj * n_pts * H;
    if (d_clip != nullptr) {
        k_clip_dhp<<<dim3((n_pts + CB_PTS - 1) / CB_PTS, n_obj), 256, 0, st>>>(theta, d_clip, n_pts, d_hp);
        OO_LAUNCH_CHECK();
    }
    const size_t smem = (size_t)SM_TOTAL * sizeof(float);
    static PerDevice attr_set;
    if (!attr_set.cur()) {
        OO_CUDA(cudaFuncSetAttribute(k_fwdbwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set.cur() = 1;
    }
    BwdParams prm;
    prm.theta = theta; prm.emb_in = emb_in; prm.d_alpha = d_alpha; prm.d_color = d_color;
    prm.d_hp = d_clip ? d_hp : nullptr;
    prm.hp_out = d_clip ? hp : nullptr;
    prm.d_emb = d_emb_out; prm.slab = slab; prm.slot_loss = slot_loss; prm.n_pts = n_pts;
    k_fwdbwd<<<dim3(gx, n_obj), NTHREADS, smem, st>>>(prm);
    OO_LAUNCH_CHECK();
    if (d_clip != nullptr) {
        k_clip_dw<<<dim3(C / CB_CH, n_obj), 256, 0, st>>>(d_clip, hp, n_pts, grads_out);
        OO_LAUNCH_CHECK();
    }
    k_reduce_slots<<<dim3(PSTRIDE / 4 / 256, n_obj), 256, 0, st>>>(slab, gx, d_clip != nullptr ? 1 : 0, grads_out);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t oo_embed_bwd_ws_floats(int n_obj, int n_pts) {
    if (n_obj <= 0 || n_pts <= 0) return -1;
    return (int64_t)n_obj * ((n_pts + EB_CHUNK - 1) / EB_CHUNK) * (NDIR * 3);
}

extern "C" int oo_embed_bwd(const float* theta, int n_obj, const float* pcs, int n_pts, float scale, const float* d_emb,
                            float* d_B, float* ws, void* stream) {
    OO_REQUIRE(theta && pcs && d_emb && d_B && ws, "oo_embed_bwd: null argument");
    OO_REQUIRE(n_obj > 0 && n_pts > 0, "oo_embed_bwd: empty input");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_chunks = (n_pts + EB_CHUNK - 1) / EB_CHUNK;
    k_embed_bwd_ens<<<dim3(n_chunks, n_obj), NDIR * 32, 0, st>>>(theta, pcs, d_emb, n_pts, scale, ws);
    OO_LAUNCH_CHECK();
    k_embed_reduce<<<n_obj, 64, 0, st>>>(ws, n_chunks, d_B);
    OO_LAUNCH_CHECK();
    return 0;
}
