// Standalone ensemble forward: vmap(pe_model) -> vmap(fc_model) (objnerf/train.py:424-425, embedding.py:46-55,
// model.py:61-103).  Same tile phases 0..7 as the fused training step; the 512-wide out_clip layer is applied
// per point here because this entry point returns the reference's per-point `clip` tensor.
#include "../../include/openobj_b200.h"
#include "oo_common.cuh"
#include "oo_tile.h"

using namespace oo;

namespace {

template <int PH, int END>
struct FwdPhases {
    static __device__ __forceinline__ void run(int tid, float* sm, const TileCtx& c, TileAcc& a) {
        tile_phase<PH, true>(tid, sm, c, a);
        __syncthreads();
        FwdPhases<PH + 1, END>::run(tid, sm, c, a);
    }
};
template <int END>
struct FwdPhases<END, END> {
    static __device__ __forceinline__ void run(int, float*, const TileCtx&, TileAcc&) {}
};

// grid = (tile groups, n_obj); each CTA stages its object's weights once and loops over 100-point tiles
__global__ void __launch_bounds__(NTHREADS, 1) k_forward(const float* __restrict__ theta, const float* __restrict__ pcs,
                                                         const float* __restrict__ emb_in, int n_pts, float scale,
                                                         float* __restrict__ alpha,
                                                         float* __restrict__ color, float* __restrict__ clip,
                                                         float* __restrict__ emb, float* __restrict__ occ) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, obj = blockIdx.y;
    const float* th = theta + (size_t)obj * PSTRIDE;
    zero_pad_rows(tid, sm);
    stage_weights(tid, sm, th);
    __syncthreads();
    TileAcc acc;        // unused by phases 0..7
    TileCtx c;
    c.scale = scale;
    c.theta = th;
    float* act = sm + SM_ACT;
    float* misc = act + R_MISC * PS;
    // this thread's out_clip row (C == NTHREADS) stays in registers for the whole CTA
    static_assert(C == NTHREADS, "k_forward maps one out_clip row to each thread");
    float w0[H], b0 = 0.f;
    if (clip != nullptr) {
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            const float4 a = *reinterpret_cast<const float4*>(th + OFF_OCL_W + tid * H + j);
            w0[j] = a.x; w0[j + 1] = a.y; w0[j + 2] = a.z; w0[j + 3] = a.w;
        }
        b0 = th[OFF_OCL_B + tid];
    }
    const int n_tiles = (n_pts + P - 1) / P;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int p0 = t * P;
        c.npts = min(P, n_pts - p0);
        c.nrays = 0;
        const size_t base = (size_t)obj * n_pts + p0;
        if (emb_in != nullptr) {
            // OccupancyMap.forward on a caller-supplied embedding: fill e1 / e2 rows directly
            for (int i = tid; i < P * EMB; i += NTHREADS) {
                const int p = i / EMB, e = i - p * EMB;
                const float v = p < c.npts ? emb_in[base * EMB + i] : 0.f;
                if (e < E1) act[(R_E1 + e) * PS + p] = v;
                else act[(R_E2 + e - E1) * PS + p] = v;
            }
            __syncthreads();
            FwdPhases<2, N_FWD_PHASES>::run(tid, sm, c, acc);
        } else {
            c.pcs = pcs + base * 3;
            if (color == nullptr) FwdPhases<0, 2>::run(tid, sm, c, acc);      // encoder only (vmap(pe_model) on its own)
            else FwdPhases<0, N_FWD_PHASES>::run(tid, sm, c, acc);
        }
        if (alpha != nullptr)
            for (int i = tid; i < c.npts; i += NTHREADS) alpha[base + i] = misc[M_DRAW * PS + i];
        if (occ != nullptr)      // render_rays.occupancy_activation without distances: sigmoid(alpha) (render_rays.py:6-14)
            for (int i = tid; i < c.npts; i += NTHREADS) occ[base + i] = 1.f / (1.f + expf(-misc[M_DRAW * PS + i]));
        if (color != nullptr)
            for (int i = tid; i < c.npts * 3; i += NTHREADS) {
                const int p = i / 3, ch = i - 3 * p;
                color[base * 3 + i] = misc[(M_COL + ch) * PS + p];
            }
        if (emb != nullptr) {
            for (int i = tid; i < c.npts * EMB; i += NTHREADS) {
                const int p = i / EMB, e = i - p * EMB;
                emb[base * EMB + i] = e < E1 ? act[(R_E1 + e) * PS + p] : act[(R_E2 + e - E1) * PS + p];
            }
        }
        if (clip != nullptr) {
            for (int p = 0; p < c.npts; ++p) {
                float f0 = b0;
#pragma unroll
                for (int j = 0; j < H; ++j) f0 += w0[j] * act[(R_HP + j) * PS + p];
                clip[(base + p) * C + tid] = f0;
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int oo_forward(const float* theta, int n_obj, const float* pcs, const float* emb_in, int n_pts, float scale,
                          float* alpha, float* color, float* clip, float* emb_out, void* stream) {
    OO_REQUIRE(theta, "oo_forward: null argument");
    OO_REQUIRE((alpha && color) || (!alpha && !color && !clip && emb_out && pcs),
               "oo_forward: alpha and color go together; without them only the encoder runs (pcs -> emb_out)");
    OO_REQUIRE((pcs != nullptr) != (emb_in != nullptr), "oo_forward: give exactly one of pcs / emb_in");
    OO_REQUIRE(n_obj > 0 && n_pts > 0, "oo_forward: empty input");
    const size_t smem = (size_t)SM_TOTAL * sizeof(float);
    static PerDevice attr_set;
    if (!attr_set.cur()) {
        OO_CUDA(cudaFuncSetAttribute(k_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set.cur() = 1;
    }
    const int n_tiles = (n_pts + P - 1) / P;
    int gx = (4 * 148 + n_obj - 1) / n_obj;
    if (gx > n_tiles) gx = n_tiles;
    if (gx < 1) gx = 1;
    k_forward<<<dim3(gx, n_obj), NTHREADS, smem, (cudaStream_t)stream>>>(theta, pcs, emb_in, n_pts, scale, alpha, color, clip, emb_out,
                                                                         nullptr);
    OO_LAUNCH_CHECK();
    return 0;
}

// ---- f3: Trainer.eval_points (objnerf/trainer.py:104-128) and the query grid of Trainer.meshing (trainer.py:46-66) ----
namespace {
// render_rays.make_3D_grid (render_rays.py:119-146): p = R (t_ijk * scale) + trans, then meshing's `grid_pc -= obj_center`
// (trainer.py:64); separate roundings in the reference's order: products, ((a + b) + c), + translation, - centre.
__global__ void k_make_grid(const oo_grid g, float* __restrict__ pts) {
    const long long n = (long long)g.dim * g.dim * g.dim;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(e % g.dim), j = (int)((e / g.dim) % g.dim), i = (int)(e / ((long long)g.dim * g.dim));
        const float x = __fmul_rn(g.t[i], g.scale[0]), y = __fmul_rn(g.t[j], g.scale[1]), z = __fmul_rn(g.t[k], g.scale[2]);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float* T = g.transform + 4 * r;
            float v = __fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), __fmul_rn(T[2], z));
            v = __fsub_rn(__fadd_rn(v, T[3]), g.center[r]);
            pts[3 * e + r] = v;
        }
    }
}
}  // namespace

namespace {
__global__ void k_occupancy(const float* __restrict__ alpha, const float* __restrict__ dist, long long n, float* __restrict__ occ) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        occ[e] = dist != nullptr ? 1.f - expf(-alpha[e] * dist[e]) : 1.f / (1.f + expf(-alpha[e]));
}
}  // namespace

extern "C" int oo_occupancy_activation(const float* alpha, const float* distances, long long n, float* occ, void* stream) {
    OO_REQUIRE(alpha && occ && n > 0, "oo_occupancy_activation: bad argument");
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    k_occupancy<<<blocks, 256, 0, (cudaStream_t)stream>>>(alpha, distances, n, occ);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_make_grid(const oo_grid* g, float* pts_out, void* stream) {
    OO_REQUIRE(g && g->t && pts_out, "oo_make_grid: null argument");
    OO_REQUIRE(g->dim > 0 && g->dim <= 1024, "oo_make_grid: grid_dim out of range");
    const long long n = (long long)g->dim * g->dim * g->dim;
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    k_make_grid<<<blocks, 256, 0, (cudaStream_t)stream>>>(*g, pts_out);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_eval_points(const float* theta, const float* pts, long long n_pts, float pe_scale, float* occ, float* color,
                              float* clip, void* stream) {
    OO_REQUIRE(theta && pts && occ && color, "oo_eval_points: null argument");
    OO_REQUIRE(n_pts > 0 && n_pts < (1LL << 31) - P, "oo_eval_points: point count out of range");
    const size_t smem = (size_t)SM_TOTAL * sizeof(float);
    static PerDevice attr_set;
    if (!attr_set.cur()) {
        OO_CUDA(cudaFuncSetAttribute(k_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set.cur() = 1;
    }
    const long long n_tiles = (n_pts + P - 1) / P;
    const int gx = (int)(n_tiles < 4 * 148 ? n_tiles : 4 * 148);
    k_forward<<<dim3(gx, 1), NTHREADS, smem, (cudaStream_t)stream>>>(theta, pts, nullptr, (int)n_pts, pe_scale, nullptr, color, clip, nullptr,
                                                                    occ);
    OO_LAUNCH_CHECK();
    return 0;
}
