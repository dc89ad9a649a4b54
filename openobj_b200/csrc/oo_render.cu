// K5/K6: full-frame evaluation render of one object (objnerf/vmap.py:604-685, trainer.py:130-198,
// utils.py:309-319) and the sequential depth-test merge across objects (train.py:577-594).
//   k_hit     : ray / OBB slab test per pixel            (trainer.py:151-169, utils.py:309-319)
//   k_compact : ranks of the hit rays (the reference draws its jitter rows only for hit rays, by rank)
//   k_render  : persistent CTAs stream the 149 midpoints of their rays through the fused forward tile
//               (phases 1..7 of oo_tile.h), composite with a warp-shuffle product scan carried across tiles,
//               apply out_clip once per ray, and write masked depth / rgb / feature maps.
#include "../../include/openobj_b200.h"
#include "oo_common.cuh"
#include "oo_tile.h"

using namespace oo;

namespace oo {
// K5 on the tensor cores (oo_forward_tc.cu): tcgen05.mma / TMEM forward + compositing, used when no dense feature map is asked for
int render_tc_launch(const oo_render_args* ra, const int* list, int* obj_start2, const float* lin, cudaStream_t st);
}  // namespace oo

namespace {

constexpr int MAXBINS = 160;


struct RenderK {
    oo_render_args a;
    int n_pix;
    int n_cta;
    // scratch (device): hit flags, near, far, rank->pixel list
    uint8_t* hit;
    float* near_;
    float* far_;
    int* list;
    const float* lin;      // device copy of torch.linspace(0,1,n_bins+1)
};

__global__ void k_hit(RenderK k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k.n_pix) return;
    const float* dc = k.a.rays_dir + (size_t)i * 3;
    const float* T = k.a.T_oc;
    const float dx = dc[0], dy = dc[1], dz = dc[2];
    const float d[3] = {T[0] * dx + T[1] * dy + T[2] * dz, T[4] * dx + T[5] * dy + T[6] * dz, T[8] * dx + T[9] * dy + T[10] * dz};
    const float o[3] = {T[3], T[7], T[11]};
    float near = -INFINITY, far = INFINITY;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const float he = k.a.half_extent[ax];
        const float tmin = __fdiv_rn(-he - o[ax], d[ax]), tmax = __fdiv_rn(he - o[ax], d[ax]);   // utils.py:310-311
        near = fmaxf(near, fminf(tmin, tmax));
        far = fminf(far, fmaxf(tmin, tmax));
    }
    const bool hit = (near <= far) && (far > 0.f);                                              // utils.py:316-318
    k.hit[i] = hit ? 1 : 0;
    k.near_[i] = fmaxf(near, 0.f);                                                               // trainer.py:168
    k.far_[i] = far + 0.2f;                                                                      // trainer.py:169
}

__global__ void __launch_bounds__(1024, 1) k_compact(RenderK k) {
    __shared__ int sh[32];
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int cpt = (k.n_pix + 1023) / 1024;
    const int b = min(tid * cpt, k.n_pix), e = min(b + cpt, k.n_pix);
    int n = 0;
    for (int i = b; i < e; ++i) n += k.hit[i];
    int s = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) sh[wv] = s;
    __syncthreads();
    if (wv == 0) {
        int c = sh[lane];
        const int own = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, c, o);
            if (lane >= o) c += t;
        }
        sh[lane] = c - own;
        if (lane == 31) k.a.n_hit[0] = c;
    }
    __syncthreads();
    int rank = sh[wv] + s - n;
    for (int i = b; i < e; ++i)
        if (k.hit[i]) k.list[rank++] = i;
}

// shared-memory extras of the render kernel live in the per-ray region of the training tile (unused here)
constexpr int SM_OPEN = SM_UT;            // [2][40]: open-ray accumulators {depth, opac, c0, c1, c2, carry, -, -, S[32]}
constexpr int SM_FIN = SM_RV;             // finished batch: rows V_DEPTH.. (see below) [NRV][12]
constexpr int F_DEPTH = 0, F_OPAC = 1, F_C0 = 2, F_NEAR = 5, F_FAR = 6, F_PIX = 7, F_J = 8;

template <int PH, int END>
struct RPhases {
    static __device__ __forceinline__ void run(int tid, float* sm, const TileCtx& c, TileAcc& a) {
        tile_phase<PH, true>(tid, sm, c, a);
        __syncthreads();
        RPhases<PH + 1, END>::run(tid, sm, c, a);
    }
};
template <int END>
struct RPhases<END, END> {
    static __device__ __forceinline__ void run(int, float*, const TileCtx&, TileAcc&) {}
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// write the finished batch of up to RT rays: feature = W_ocl S + b*opacity, mask test, outputs
__device__ void emit_batch(int tid, float* sm, const RenderK& k, const float* th, int n_fin) {
    float* fin = sm + SM_FIN;
    const bool want_feat = k.a.feat != nullptr;
    if (want_feat) {
        static_assert(C == NTHREADS, "emit_batch maps one out_clip row to each thread");
        float f[RT];
        const int c0 = tid;
        const float b0 = th[OFF_OCL_B + c0];
#pragma unroll
        for (int r = 0; r < RT; ++r) f[r] = b0 * fin[F_OPAC * RP + r];
        const float4* w0 = reinterpret_cast<const float4*>(th + OFF_OCL_W + c0 * H);
#pragma unroll
        for (int j4 = 0; j4 < H / 4; ++j4) {
            const float4 a = w0[j4];
            const float wa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* srow = sm + SM_ST + (4 * j4 + q) * RP;
#pragma unroll
                for (int r = 0; r < RT; ++r) f[r] += wa[q] * srow[r];
            }
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            if (r < n_fin) {
                const float d = fin[F_DEPTH * RP + r], op = fin[F_OPAC * RP + r];
                const bool bad = d < fin[F_NEAR * RP + r] || d > fin[F_FAR * RP + r] || op < 0.9f;   // vmap.py:665,672
                const size_t pix = (size_t)__float_as_int(fin[F_PIX * RP + r]);
                k.a.feat[pix * C + c0] = bad ? 0.f : f[r];
            }
        }
    }
    if (k.a.ray_rec != nullptr) {
        // compact per-hit record {S[32] = sum_i T_i hp_i, opacity, -, -, -}: the 512-wide out_clip layer is applied later, and
        // only for the pixels this object wins in the depth-test merge (oo_winner_features)
        for (int i = tid; i < n_fin * OO_RENDER_REC; i += NTHREADS) {
            const int r = i / OO_RENDER_REC, e = i - r * OO_RENDER_REC;
            const size_t j = (size_t)__float_as_int(fin[F_J * RP + r]);
            k.a.ray_rec[j * OO_RENDER_REC + e] = e < H ? sm[SM_ST + e * RP + r] : (e == H ? fin[F_OPAC * RP + r] : 0.f);
        }
    }
    if (tid < n_fin) {
        const int r = tid;
        const float d = fin[F_DEPTH * RP + r], op = fin[F_OPAC * RP + r];
        const bool bad = d < fin[F_NEAR * RP + r] || d > fin[F_FAR * RP + r] || op < 0.9f;
        const size_t pix = (size_t)__float_as_int(fin[F_PIX * RP + r]);
        k.a.mask[pix] = bad ? 0 : 1;
        k.a.depth[pix] = bad ? 0.f : d;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float v = __fmul_rn(fin[(F_C0 + ch) * RP + r], 255.f);                             // vmap.py:671
            k.a.rgb[pix * 3 + ch] = bad ? 0 : (uint8_t)(int)v;
        }
        if (k.a.opacity) k.a.opacity[pix] = op;
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) k_render(const RenderK k) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int n_hit = k.a.n_hit[0];
    if (n_hit <= 1) return;                                             // trainer.py:167 "<= 1 -> miss"
    const int nmid = k.a.n_bins - 1;
    const int j0 = (int)(((long long)n_hit * blockIdx.x) / gridDim.x), j1 = (int)(((long long)n_hit * (blockIdx.x + 1)) / gridDim.x);
    if (j0 >= j1) return;
    const float* th = k.a.theta1;
    zero_pad_rows(tid, sm);
    stage_weights(tid, sm, th);
    float* act = sm + SM_ACT;
    float* misc = act + R_MISC * PS;
    float* open = sm + SM_OPEN;
    float* fin = sm + SM_FIN;
    if (tid < 80) open[tid] = (tid == 5 || tid == 45) ? 1.f : 0.f;      // carry (free-probability product) starts at 1
    __syncthreads();
    TileAcc acc;
    TileCtx c;
    c.scale = k.a.scale;
    c.theta = th;
    const float* Tw = k.a.T_wc;
    const float ox = Tw[3], oy = Tw[7], oz = Tw[11];
    int cur = 0;          // which open-ray slot continues from the previous tile
    int n_fin = 0;
    const long long q_begin = (long long)j0 * nmid, q_end = (long long)j1 * nmid;
    for (long long q0 = q_begin; q0 < q_end; q0 += P) {
        const int npts = (int)min((long long)P, q_end - q0);
        // ---- points of this tile: midpoints of the jittered bins (trainer.py:174-178) ------------------
        for (int p = tid; p < P; p += NTHREADS) {
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, zm = 0.f;
            if (p < npts) {
                const long long q = q0 + p;
                const int j = (int)(q / nmid), kk = (int)(q - (long long)j * nmid);
                const int pix = k.list[j];
                const float near = k.near_[pix], far = k.far_[pix];
                const float range = __fsub_rn(far, near), blen = __fdiv_rn(range, (float)k.a.n_bins);
                const float* u = k.a.jitter + (size_t)(k.a.jitter_by_rank ? j : pix) * k.a.n_bins;
                const float za = __fadd_rn(__fadd_rn(__fmul_rn(range, k.lin[kk]), near), __fmul_rn(u[kk], blen));
                const float zb = __fadd_rn(__fadd_rn(__fmul_rn(range, k.lin[kk + 1]), near), __fmul_rn(u[kk + 1], blen));
                zm = __fmul_rn(0.5f, __fadd_rn(zb, za));
                const float* dc = k.a.rays_dir + (size_t)pix * 3;
                const float dx = dc[0], dy = dc[1], dz = dc[2];
                const float wx = Tw[0] * dx + Tw[1] * dy + Tw[2] * dz, wy = Tw[4] * dx + Tw[5] * dy + Tw[6] * dz,
                            wz = Tw[8] * dx + Tw[9] * dy + Tw[10] * dz;
                t0 = (ox + wx * zm) / c.scale; t1 = (oy + wy * zm) / c.scale; t2 = (oz + wz * zm) / c.scale;
            }
            act[(R_T + 0) * PS + p] = t0; act[(R_T + 1) * PS + p] = t1; act[(R_T + 2) * PS + p] = t2;
            act[(R_E1 + 0) * PS + p] = t0; act[(R_E1 + 1) * PS + p] = t1; act[(R_E1 + 2) * PS + p] = t2;
            misc[M_HU * PS + p] = zm;
        }
        __syncthreads();
        RPhases<1, N_FWD_PHASES>::run(tid, sm, c, acc);
        // ---- compositing: the tile holds the tail of the open ray (segment A) and maybe the head of the next (B)
        const int jA = (int)(q0 / nmid);
        const int kA = (int)(q0 - (long long)jA * nmid);
        const int lenA = min(npts, nmid - kA);
        const int lenB = npts - lenA;                                    // < nmid because nmid > P
        if (wv < 2) {
            const int pa = wv == 0 ? 0 : lenA, len = wv == 0 ? lenA : lenB;
            float* acc_o = open + (wv == 0 ? cur : cur ^ 1) * 40;
            if (len > 0) {
                float carry = acc_o[5];
                float sd = 0.f, so = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
                for (int b0 = 0; b0 < len; b0 += 32) {
                    const int p = pa + b0 + lane;
                    const bool in = b0 + lane < len;
                    const float o = in ? misc[M_OCC * PS + p] : 0.f;
                    float inc = in ? (1.f - o + 1e-10f) : 1.f;                          // render_rays.py:41
#pragma unroll
                    for (int s = 1; s < 32; s <<= 1) {
                        const float t = __shfl_up_sync(0xffffffffu, inc, s);
                        if (lane >= s) inc *= t;
                    }
                    float ex = __shfl_up_sync(0xffffffffu, inc, 1);
                    if (lane == 0) ex = 1.f;
                    const float T = o * carry * ex;                                      // render_rays.py:43
                    carry *= __shfl_sync(0xffffffffu, inc, 31);
                    if (in) {
                        misc[M_TERM * PS + p] = T;
                        sd += T * misc[M_HU * PS + p];
                        so += T;
                        s0 += T * misc[(M_COL + 0) * PS + p];
                        s1 += T * misc[(M_COL + 1) * PS + p];
                        s2 += T * misc[(M_COL + 2) * PS + p];
                    }
                }
                sd = wsum(sd); so = wsum(so); s0 = wsum(s0); s1 = wsum(s1); s2 = wsum(s2);
                __syncwarp();
                // S[j] += sum_p T_p hp[j][p]  (lane = hidden unit)
                float sj = 0.f;
                for (int p = pa; p < pa + len; ++p) sj += misc[M_TERM * PS + p] * act[(R_HP + lane) * PS + p];
                acc_o[8 + lane] += sj;
                if (lane == 0) {
                    acc_o[0] += sd; acc_o[1] += so; acc_o[2] += s0; acc_o[3] += s1; acc_o[4] += s2; acc_o[5] = carry;
                }
            }
        }
        __syncthreads();
        // ---- rays that ended in this tile move to the finished batch ---------------------------------------
        const bool endA = lenA > 0 && kA + lenA == nmid;
        const bool endB = false;   // B can only end if nmid <= P, excluded
        (void)endB;
        if (endA) {
            const float* acc_o = open + cur * 40;
            if (tid < H) sm[SM_ST + tid * RP + n_fin] = acc_o[8 + tid];
            if (tid == 0) {
                const int pix = k.list[jA];
                fin[F_DEPTH * RP + n_fin] = acc_o[0];
                fin[F_OPAC * RP + n_fin] = acc_o[1];
                fin[(F_C0 + 0) * RP + n_fin] = acc_o[2];
                fin[(F_C0 + 1) * RP + n_fin] = acc_o[3];
                fin[(F_C0 + 2) * RP + n_fin] = acc_o[4];
                fin[F_NEAR * RP + n_fin] = k.near_[pix];
                fin[F_FAR * RP + n_fin] = k.far_[pix];
                fin[F_PIX * RP + n_fin] = __int_as_float(pix);
                fin[F_J * RP + n_fin] = __int_as_float(jA);
            }
            __syncthreads();
            if (tid < 40) open[cur * 40 + tid] = tid == 5 ? 1.f : 0.f;   // reset the slot for a later ray
            cur ^= 1;                                                   // segment B (if any) is now the open ray
            ++n_fin;
            __syncthreads();
            if (n_fin == RT) {
                emit_batch(tid, sm, k, th, n_fin);
                n_fin = 0;
                __syncthreads();
            }
        }
    }
    if (n_fin > 0) emit_batch(tid, sm, k, th, n_fin);
}

__global__ void k_zmerge(const uint8_t* __restrict__ masks, const float* __restrict__ depths, const uint8_t* __restrict__ rgbs,
                         const uint8_t* __restrict__ is_bg, int n_obj, int64_t n_pix, float* __restrict__ depth_out,
                         uint8_t* __restrict__ rgb_out, int32_t* __restrict__ winner) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    float dbuf = 100.f;                                                 // train.py:562
    int win = -1;
    uint8_t r = 0, g = 0, b = 0;
    for (int o = 0; o < n_obj; ++o) {
        const size_t q = (size_t)o * n_pix + i;
        if (!masks[q]) continue;
        const float d = depths[q];
        if (dbuf > d) {                                                 // train.py:582 strict test
            r = rgbs[q * 3]; g = rgbs[q * 3 + 1]; b = rgbs[q * 3 + 2];
            win = o;
            if (!is_bg[o]) dbuf = d;                                    // train.py:593-594
        }
    }
    depth_out[i] = dbuf;
    rgb_out[i * 3] = r; rgb_out[i * 3 + 1] = g; rgb_out[i * 3 + 2] = b;
    winner[i] = win;
}

// the same merge over per-object POINTERS (objects of different ranks sit in different places of an all-gathered buffer):
// object q of the global insertion order has mask / depth / rgb maps at masks[q] / depths[q] / rgbs[q]
__global__ void k_zmerge_ptr(const uint8_t* const* __restrict__ masks, const float* const* __restrict__ depths,
                             const uint8_t* const* __restrict__ rgbs, const uint8_t* __restrict__ is_bg, int n_obj, int64_t n_pix,
                             float* __restrict__ depth_out, uint8_t* __restrict__ rgb_out, int32_t* __restrict__ winner) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    float dbuf = 100.f;                                                 // train.py:562
    int win = -1;
    uint8_t r = 0, g = 0, b = 0;
    for (int o = 0; o < n_obj; ++o) {
        if (!masks[o][i]) continue;
        const float d = depths[o][i];
        if (dbuf > d) {                                                 // train.py:582 strict test
            const uint8_t* c = rgbs[o] + i * 3;
            r = c[0]; g = c[1]; b = c[2];
            win = o;
            if (!is_bg[o]) dbuf = d;                                    // train.py:593-594
        }
    }
    depth_out[i] = dbuf;
    rgb_out[i * 3] = r; rgb_out[i * 3 + 1] = g; rgb_out[i * 3 + 2] = b;
    winner[i] = win;
}

// Part features of the pixels ONE object won in the merge: feat = W_ocl S + b_ocl * opacity (the legal restructure of
// out_clip is linear and only consumed through the compositing sum) for the hits whose pixel has winner == k.
// One warp per hit: lanes = 16 output channels each; rows are appended to a compact list (the order of the list is not
// deterministic, its content is: every row carries its pixel).
__global__ void __launch_bounds__(256) k_winner_features(const float* __restrict__ theta1, const float* __restrict__ ray_rec,
                                                         const int32_t* __restrict__ hit_pix, const int* __restrict__ n_hit,
                                                         const int32_t* __restrict__ winner, int k_global, int64_t cap,
                                                         float* __restrict__ rows, int32_t* __restrict__ row_pix,
                                                         int* __restrict__ n_rows) {
    __shared__ float S[8][OO_RENDER_REC];
    const int wv = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = n_hit[0];
    for (int j = blockIdx.x * 8 + wv; j < n; j += gridDim.x * 8) {
        const int pix = hit_pix[j];
        if (winner[pix] != k_global) continue;                          // warp-uniform
        int slot = 0;
        if (lane == 0) slot = atomicAdd(n_rows, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= cap) continue;
        for (int e = lane; e < OO_RENDER_REC; e += 32) S[wv][e] = ray_rec[(size_t)j * OO_RENDER_REC + e];
        __syncwarp();
        const float op = S[wv][H];
        for (int c = lane; c < C; c += 32) {
            float f = theta1[OFF_OCL_B + c] * op;
            const float4* w = reinterpret_cast<const float4*>(theta1 + OFF_OCL_W + c * H);
#pragma unroll
            for (int j4 = 0; j4 < H / 4; ++j4) {
                const float4 a = w[j4];
                f += a.x * S[wv][4 * j4] + a.y * S[wv][4 * j4 + 1] + a.z * S[wv][4 * j4 + 2] + a.w * S[wv][4 * j4 + 3];
            }
            rows[(size_t)slot * C + c] = f;
        }
        if (lane == 0) row_pix[slot] = pix;
        __syncwarp();
    }
}

}  // namespace

namespace {
// the same for ALL local objects in one launch over the pooled hits of oo_render_frame: hit g belongs to the object o with
// obj_start[o] <= g < obj_start[o + 1] (binary search), whose global ensemble index is k_of[o]
__global__ void __launch_bounds__(256) k_winner_features_pool(const float* const* __restrict__ theta, const float* __restrict__ ray_rec,
                                                              const int32_t* __restrict__ hit_pix, const int* __restrict__ obj_start,
                                                              int n_obj, const int32_t* __restrict__ k_of,
                                                              const int32_t* __restrict__ winner, int64_t pool_rows, int64_t cap,
                                                              float* __restrict__ rows, int32_t* __restrict__ row_pix,
                                                              int* __restrict__ n_rows) {
    __shared__ float S[8][OO_RENDER_REC];
    const int wv = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long n = obj_start[n_obj];
    if (n > pool_rows) n = pool_rows;
    for (long long g = (long long)blockIdx.x * 8 + wv; g < n; g += (long long)gridDim.x * 8) {
        int lo = 0, hi = n_obj - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((long long)obj_start[mid] <= g) lo = mid; else hi = mid - 1;
        }
        const int pix = hit_pix[g];
        if (winner[pix] != k_of[lo]) continue;                          // warp-uniform
        int slot = 0;
        if (lane == 0) slot = atomicAdd(n_rows, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= cap) continue;
        for (int e = lane; e < OO_RENDER_REC; e += 32) S[wv][e] = ray_rec[(size_t)g * OO_RENDER_REC + e];
        __syncwarp();
        const float* th = theta[lo];
        const float op = S[wv][H];
        for (int c = lane; c < C; c += 32) {
            float f = th[OFF_OCL_B + c] * op;
            const float4* w = reinterpret_cast<const float4*>(th + OFF_OCL_W + c * H);
#pragma unroll
            for (int j4 = 0; j4 < H / 4; ++j4) {
                const float4 a = w[j4];
                f += a.x * S[wv][4 * j4] + a.y * S[wv][4 * j4 + 1] + a.z * S[wv][4 * j4 + 2] + a.w * S[wv][4 * j4 + 3];
            }
            rows[(size_t)slot * C + c] = f;
        }
        if (lane == 0) row_pix[slot] = pix;
        __syncwarp();
    }
}

}  // namespace

extern "C" int oo_winner_features_frame(const float* const* theta, const float* ray_rec, const int32_t* hit_pix, const int* obj_start,
                                        int n_obj, const int32_t* k_of, const int32_t* winner, int64_t pool_rows, int64_t cap_rows,
                                        float* rows, int32_t* row_pix, int* n_rows, void* stream) {
    OO_REQUIRE(theta && ray_rec && hit_pix && obj_start && k_of && winner && rows && row_pix && n_rows && n_obj >= 1,
               "oo_winner_features_frame: null argument");
    k_winner_features_pool<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(theta, ray_rec, hit_pix, obj_start, n_obj, k_of, winner, pool_rows,
                                                                      cap_rows, rows, row_pix, n_rows);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_zmerge_ptr(const uint8_t* const* masks, const float* const* depths, const uint8_t* const* rgbs,
                             const uint8_t* is_bg, int n_obj, int64_t n_pix, float* depth_out, uint8_t* rgb_out,
                             int32_t* winner_out, void* stream) {
    OO_REQUIRE(masks && depths && rgbs && is_bg && depth_out && rgb_out && winner_out, "oo_zmerge_ptr: null argument");
    OO_REQUIRE(n_obj >= 0 && n_pix > 0, "oo_zmerge_ptr: bad size");
    k_zmerge_ptr<<<(unsigned)((n_pix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(masks, depths, rgbs, is_bg, n_obj, n_pix,
                                                                                    depth_out, rgb_out, winner_out);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_winner_features(const float* theta1, const float* ray_rec, const int32_t* hit_pix, const int* n_hit,
                                  const int32_t* winner, int k_global, int64_t cap_rows, float* rows, int32_t* row_pix,
                                  int* n_rows, void* stream) {
    OO_REQUIRE(theta1 && ray_rec && hit_pix && n_hit && winner && rows && row_pix && n_rows, "oo_winner_features: null argument");
    k_winner_features<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(theta1, ray_rec, hit_pix, n_hit, winner, k_global, cap_rows, rows,
                                                                 row_pix, n_rows);
    OO_LAUNCH_CHECK();
    return 0;
}

// scratch for the render call is carved from one cudaMallocAsync allocation on `stream`
extern "C" int oo_render_object(const oo_render_args* a, void* stream) {
    OO_REQUIRE(a && a->theta1 && a->T_wc && a->T_oc && a->half_extent && a->rays_dir && a->jitter,
               "oo_render_object: null input");
    OO_REQUIRE(a->mask && a->depth && a->rgb && a->n_hit, "oo_render_object: null output");
    OO_REQUIRE(a->n_bins > P + 1 && a->n_bins <= MAXBINS, "oo_render_object: need %d < n_bins <= %d", P + 1, MAXBINS);
    OO_REQUIRE(a->lin_host != nullptr, "oo_render_object: null linspace table");
    OO_REQUIRE((a->ray_rec == nullptr) == (a->hit_pix == nullptr), "oo_render_object: ray_rec and hit_pix go together");
    cudaStream_t st = (cudaStream_t)stream;
    RenderK k;
    k.a = *a;
    k.n_pix = a->W * a->H;
    const size_t n = (size_t)k.n_pix;
    const size_t bytes = n * (1 + 4 + 4 + 4) + MAXBINS * 4 + 256 + 64;
    char* scratch = nullptr;
    OO_CUDA(cudaMallocAsync((void**)&scratch, bytes, st));
    k.near_ = (float*)scratch;
    k.far_ = k.near_ + n;
    k.list = (int*)(k.far_ + n);
    float* lin = (float*)(k.list + n);
    k.lin = lin;
    k.hit = (uint8_t*)(lin + MAXBINS);
    OO_CUDA(cudaMemcpyAsync(lin, a->lin_host, (a->n_bins + 1) * sizeof(float), cudaMemcpyHostToDevice, st));
    OO_CUDA(cudaMemsetAsync(a->mask, 0, n, st));
    OO_CUDA(cudaMemsetAsync(a->depth, 0, n * 4, st));
    OO_CUDA(cudaMemsetAsync(a->rgb, 0, n * 3, st));
    if (a->feat) OO_CUDA(cudaMemsetAsync(a->feat, 0, n * C * 4, st));
    if (a->opacity) OO_CUDA(cudaMemsetAsync(a->opacity, 0, n * 4, st));
    k_hit<<<(k.n_pix + 255) / 256, 256, 0, st>>>(k);
    OO_LAUNCH_CHECK();
    k_compact<<<1, 1024, 0, st>>>(k);
    OO_LAUNCH_CHECK();
    if (a->hit_pix != nullptr) OO_CUDA(cudaMemcpyAsync(a->hit_pix, k.list, n * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (a->feat == nullptr && !a->force_mma_sync) {
        // no dense [W][H][512] map wanted (the winner-only feature path, or no features at all): the tcgen05 / TMEM kernel
        if (int rc = render_tc_launch(a, k.list, reinterpret_cast<int*>(k.hit + ((n + 15) & ~(size_t)15)), k.lin, st)) return rc;
        OO_CUDA(cudaFreeAsync(scratch, st));
        return 0;
    }
    const size_t smem = (size_t)SM_TOTAL * sizeof(float);
    static PerDevice attr_set;
    if (!attr_set.cur()) {
        OO_CUDA(cudaFuncSetAttribute(k_render, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set.cur() = 1;
    }
    int dev = 0, n_sm = 148;
    OO_CUDA(cudaGetDevice(&dev));
    OO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    k.n_cta = n_sm;
    k_render<<<n_sm, NTHREADS, smem, st>>>(k);
    OO_LAUNCH_CHECK();
    OO_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

extern "C" int oo_zmerge(const uint8_t* masks, const float* depths, const uint8_t* rgbs, const uint8_t* is_bg, int n_obj,
                         int64_t n_pix, float* depth_out, uint8_t* rgb_out, int32_t* winner_out, void* stream) {
    OO_REQUIRE(masks && depths && rgbs && is_bg && depth_out && rgb_out && winner_out, "oo_zmerge: null argument");
    OO_REQUIRE(n_obj >= 0 && n_pix > 0, "oo_zmerge: bad size");
    k_zmerge<<<(unsigned)((n_pix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(masks, depths, rgbs, is_bg, n_obj, n_pix,
                                                                                depth_out, rgb_out, winner_out);
    OO_LAUNCH_CHECK();
    return 0;
}
