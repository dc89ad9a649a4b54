// K2: keyframe / pixel draw, gathers, ray transform and depth-guided sample placement for ALL objects in one
// launch (objnerf/vmap.py:386-554, utils.py:324-397), plus the counter-based RNG used for throughput runs.
// One 1024-thread CTA per object: pass A gathers and classifies the object's rays, a block scan gives every ray
// its rank inside its class (the reference consumes its random rows by rank) and the batch-max depth
// (quirk 6), pass B places the samples.  HBM-bound: ~180 B per ray.
#include "../../include/openobj_b200.h"
#include "oo_common.cuh"

namespace {

constexpr int NTH = 1024;
constexpr int MAXB = 33;

struct SampleK {
    oo_sample_args a;
    float lin_s[MAXB], lin_c[MAXB], lin_b[MAXB];
};

// ---- Philox4x32-10 counter RNG: value i of object id `oid` in frame `frame` depends only on (seed, frame, oid, i)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}


constexpr float TWO_M24 = 5.9604644775390625e-8f;
struct Rng {
    uint2 key;
    uint32_t oid, frame;
    __device__ __forceinline__ uint32_t word(uint32_t stream, uint64_t e) const {
        const uint64_t q = e >> 2;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), oid, 8u * frame + stream), key);
        const uint32_t i = (uint32_t)(e & 3);
        return i == 0 ? r.x : i == 1 ? r.y : i == 2 ? r.z : r.w;
    }
    __device__ __forceinline__ float uniform(uint32_t stream, uint64_t e) const { return (float)(word(stream, e) >> 8) * TWO_M24; }
};

struct RayPix {
    int kf, iw, ih;
    float iwf, ihf;
    bool oob;
};

// pixel of a ray from its two uniforms: separate fp32 mul and add, then truncation -- vmap.py:418-422
__device__ __forceinline__ RayPix pixel_from_uniforms(const oo_sample_args& a, int obj, int kf, float uw, float uh) {
    RayPix r;
    r.kf = kf;
    const float* bb = a.store_rgbi != nullptr ? a.slot_bbox + ((size_t)obj * a.kf_cap + kf) * 4 : a.bbox[obj] + 4 * kf;
    r.iwf = __fadd_rn(__fmul_rn(uw, __fsub_rn(bb[1], bb[0])), bb[0]);
    r.ihf = __fadd_rn(__fmul_rn(uh, __fsub_rn(bb[3], bb[2])), bb[2]);
    int iw = (int)r.iwf, ih = (int)r.ihf;
    r.oob = iw < 0 || iw >= a.W || ih < 0 || ih >= a.H;   // quirk 11: the reference would raise an index error
    r.iw = min(max(iw, 0), a.W - 1);
    r.ih = min(max(ih, 0), a.H - 1);
    return r;
}

// ---- pass A for one ray: gathers, per-ray outputs, class (0 invalid depth / 1 this object / 2 other)
__device__ __forceinline__ int gather_ray(const oo_sample_args& a, int obj, int ray, int n_rays, const RayPix& p, float& d_out,
                                          int& oob) {
    oob += p.oob;
    uchar4 c;
    float d;
    if (a.store_rgbi != nullptr) {
        // shared keyframe store: one copy of the frame for all objects; the per-object pixel state of train.py:203-205
        // (1 where inst == id, 2 where inst == -1, else 0) is derived here from the instance map
        const size_t pix = ((size_t)a.slot_frame[(size_t)obj * a.kf_cap + p.kf] * a.W + p.iw) * a.H + p.ih;
        const int2 v = reinterpret_cast<const int2*>(a.store_rgbi)[pix];
        c = make_uchar4((unsigned char)(v.x & 255), (unsigned char)((v.x >> 8) & 255), (unsigned char)((v.x >> 16) & 255),
                        (unsigned char)(v.y == a.obj_ids[obj] ? 1 : (v.y == -1 ? 2 : 0)));
        d = a.store_depth[pix];
    } else {
        const size_t pix = ((size_t)p.kf * a.W + p.iw) * a.H + p.ih;
        c = *reinterpret_cast<const uchar4*>(a.rgbs[obj] + pix * 4);              // vmap.py:424
        d = a.depth[obj][pix];                                                    // vmap.py:425
    }
    const size_t o = (size_t)obj * n_rays + ray;
    a.gt_rgb[o * 3 + 0] = c.x; a.gt_rgb[o * 3 + 1] = c.y; a.gt_rgb[o * 3 + 2] = c.z;
    a.gt_depth[o] = d;
    a.labels[o] = c.w;
    const bool invalid = d <= a.min_bound;                                      // vmap.py:485
    a.valid[o] = invalid ? 0 : 1;
    if (a.pix) {
        a.pix[o * 3 + 0] = p.kf; a.pix[o * 3 + 1] = p.iw; a.pix[o * 3 + 2] = p.ih;
    }
    if (a.feat_row) {                                                           // vmap.py:437-452
        const int pw = min(max((int)floorf(__fdiv_rn(p.iwf, (float)a.part_down)), 0), a.pw - 1);
        const int ph = min(max((int)floorf(__fdiv_rn(p.ihf, (float)a.part_down)), 0), a.ph - 1);
        a.feat_row[o] = (a.part_frame[(size_t)obj * a.kf_cap + p.kf] * a.pw + pw) * a.ph + ph;
    }
    d_out = d;
    return invalid ? 0 : (c.w == 1 ? 1 : 2);
}

// the 25-exchange network for 9 inputs (Floyd; checked exhaustively with the 0-1 principle): columns, merge of the three
// sorted rows of a 3 x 3 arrangement, clean-up
__device__ __forceinline__ void cmpex(float& a, float& b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo; b = hi;
}
__device__ __forceinline__ void sort9(float* v) {
    cmpex(v[0], v[1]); cmpex(v[3], v[4]); cmpex(v[6], v[7]);
    cmpex(v[1], v[2]); cmpex(v[4], v[5]); cmpex(v[7], v[8]);
    cmpex(v[0], v[1]); cmpex(v[3], v[4]); cmpex(v[6], v[7]);
    cmpex(v[0], v[3]); cmpex(v[3], v[6]); cmpex(v[0], v[3]);
    cmpex(v[1], v[4]); cmpex(v[4], v[7]); cmpex(v[1], v[4]);
    cmpex(v[2], v[5]); cmpex(v[5], v[8]); cmpex(v[2], v[5]);
    cmpex(v[1], v[3]); cmpex(v[5], v[7]); cmpex(v[2], v[6]); cmpex(v[4], v[6]);
    cmpex(v[2], v[4]); cmpex(v[2], v[3]); cmpex(v[5], v[6]);
}

// ascending sort of n <= N values held in registers: odd-even transposition network (N rounds of compare-exchange)
template <int N>
__device__ __forceinline__ void sort_regs(float* v, int n) {
    if (N == 9 && n == 9) {
        sort9(v);
        return;
    }
#pragma unroll
    for (int round = 0; round < N; ++round)
#pragma unroll
        for (int i = round & 1; i + 1 < N; i += 2)
            if (i + 1 < n) {
                const float lo = fminf(v[i], v[i + 1]), hi = fmaxf(v[i], v[i + 1]);
                v[i] = lo; v[i + 1] = hi;
            }
}

// ---- pass B for one ray: depth placement along the ray and the sample points.  max_bound = max sampled depth of the
// object's batch (vmap.py:489, quirk 6).  NC / NB = compile-time upper bounds of n_c2s / n_bins (the shipped configurations
// 1 + 9 and 5 + 9 get exact instantiations, so every per-ray array lives in registers); zo [S] / po [S][3]: where the ray's
// depths and points go (global rows, or the CTA's staging tile in the parallel path).  `dr` supplies the ray's random draws
// (TapeDraws: rows of the caller's tapes; CounterDraws: the counter RNG evaluated here), class by class, so that only the
// draws the ray's class consumes are produced.
template <int NC, int NB, bool EXACT, class Draws>
__device__ __forceinline__ void place_ray(const SampleK& k, int obj, int kf, int iw, int ih, float d, int state, float max_bound,
                                          const Draws& dr, float* zo, float* po) {
    const oo_sample_args& a = k.a;
    const int nc = EXACT ? NC : a.n_c2s, nb = EXACT ? NB : a.n_bins, S = nc + nb;   // EXACT: compile-time loop bounds
    const float eps = a.eps;
    const bool invalid = d <= a.min_bound;
    float zs[NC + NB];
    if (invalid) {
        // stratified_bins(min_bound, max(sampled_depth), S) -- vmap.py:493-498, utils.py:342-379
        float u[NC + NB];
        dr.invalid(u);
        const float range = __fsub_rn(max_bound, a.min_bound);
        const float blen = __fdiv_rn(range, (float)S);
#pragma unroll
        for (int i = 0; i < NC + NB; ++i)
            if (i < S) zs[i] = __fadd_rn(__fadd_rn(__fmul_rn(range, k.lin_s[i]), a.min_bound), __fmul_rn(u[i], blen));
    } else {
        float uc[NC], ub[NB];
        {   // cam -> surface: stratified_bins(min_bound, d - eps, n_c2s) -- vmap.py:506-509
            dr.valid(uc);
            const float range = __fsub_rn(__fsub_rn(d, eps), a.min_bound);
            const float blen = __fdiv_rn(range, (float)nc);
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (i < nc) zs[i] = __fadd_rn(__fadd_rn(__fmul_rn(range, k.lin_c[i]), a.min_bound), __fmul_rn(uc[i], blen));
        }
        if (state == 1) {
            // normal_bins_sampling: N(0, eps/3) draws sorted ascending, clipped to +-eps, + d -- utils.py:382-397
            dr.normal(ub, __fdiv_rn(eps, 3.f));
            sort_regs<NB>(ub, nb);
#pragma unroll
            for (int i = 0; i < NB; ++i)
                if (i < nb) ub[i] = __fadd_rn(d, fminf(fmaxf(ub[i], -eps), eps));
        } else {
            // stratified_bins(d - eps, d + other_eps, n_bins) -- vmap.py:538-542
            dr.other(ub);
            const float lo = __fsub_rn(d, eps), hi = __fadd_rn(d, a.other_eps);
            const float range = __fsub_rn(hi, lo);
            const float blen = __fdiv_rn(range, (float)nb);
#pragma unroll
            for (int i = 0; i < NB; ++i)
                if (i < nb) ub[i] = __fadd_rn(__fadd_rn(__fmul_rn(range, k.lin_b[i]), lo), __fmul_rn(ub[i], blen));
        }
        // zs = [cam->surface bins, surface bins]: the split point nc is a run-time value only in the generic instantiation
        if (EXACT) {
#pragma unroll
            for (int j = 0; j < NB; ++j) zs[NC + j] = ub[j];
        } else {
#pragma unroll
            for (int i = 0; i < NC + NB; ++i) {
                if (i >= nc && i < S) {
#pragma unroll
                    for (int j = 0; j < NB; ++j)
                        if (j == i - nc) zs[i] = ub[j];
                }
            }
        }
    }
    // rays: dir_W = R dir_C, origin = T[:3,3] (utils.py:324-336); points = o + d*z (vmap.py:548-549)
    const float* T = a.store_rgbi != nullptr ? a.store_twc + 16 * (size_t)a.slot_frame[(size_t)obj * a.kf_cap + kf]
                                             : a.t_wc[obj] + 16 * kf;
    const float* dc = a.rays_dir + ((size_t)iw * a.H + ih) * 3;
    const float dx = dc[0], dy = dc[1], dz = dc[2];
    const float wx = T[0] * dx + T[1] * dy + T[2] * dz;
    const float wy = T[4] * dx + T[5] * dy + T[6] * dz;
    const float wz = T[8] * dx + T[9] * dy + T[10] * dz;
    const float ox = T[3], oy = T[7], oz = T[11];
#pragma unroll
    for (int i = 0; i < NC + NB; ++i) {
        if (i < S) {
            zo[i] = zs[i];
            po[3 * i + 0] = __fadd_rn(ox, __fmul_rn(wx, zs[i]));
            po[3 * i + 1] = __fadd_rn(oy, __fmul_rn(wy, zs[i]));
            po[3 * i + 2] = __fadd_rn(oz, __fmul_rn(wz, zs[i]));
        }
    }
}

// draws read from the caller's tapes; row = rank of the ray inside its class (tape_by_rank) or the ray index
template <int NC, int NB>
struct TapeDraws {
    const oo_sample_args& a;
    size_t row_inv, row_val, row_obj, row_oth;
    __device__ __forceinline__ void invalid(float* u) const {
        const int S = a.n_c2s + a.n_bins;
#pragma unroll
        for (int i = 0; i < NC + NB; ++i)
            if (i < S) u[i] = a.r_invalid[row_inv * S + i];
    }
    __device__ __forceinline__ void valid(float* uc) const {
#pragma unroll
        for (int i = 0; i < NC; ++i)
            if (i < a.n_c2s) uc[i] = a.r_valid[row_val * a.n_c2s + i];
    }
    __device__ __forceinline__ void normal(float* ub, float) const {
#pragma unroll
        for (int i = 0; i < NB; ++i)
            if (i < a.n_bins) ub[i] = a.r_normal[row_obj * a.n_bins + i];
    }
    __device__ __forceinline__ void other(float* ub) const {
#pragma unroll
        for (int i = 0; i < NB; ++i)
            if (i < a.n_bins) ub[i] = a.r_other[row_oth * a.n_bins + i];
    }
};

// ---- the counter RNG's ray-blocked stream (rng_mode 1; oo_rng_fill_rows produces the same values as tapes) ----------
// Ray r of object `oid` in frame `frame` owns the words w[4 j + c] = word c of philox(counter = (r, j, oid, 8 frame + 1)):
//   w[0] = u_w, w[1] = u_h, w[2 .. 2 + n_c2s) = cam->surface draws, and from B0 = 4 ceil((2 + n_c2s) / 4) the bin draws:
//   w[B0 + i] uniform (i < S for an invalid-depth ray, i < n_bins for an other-object ray), or for a this-object ray the
//   Box-Muller pairs (w[B0 + 2k], w[B0 + 2k + 1]) -> normals 2k (cos) and 2k + 1 (sin).
// Every ray therefore needs ceil(words / 4) Philox blocks and no word is computed twice; the keyframe draw of frame slot f
// stays element f of stream 8 frame + 0.
__host__ __device__ constexpr int ray_b0(int nc) { return 4 * ((2 + nc + 3) / 4); }

__device__ __forceinline__ uint4 ray_block(const Rng& g, uint32_t ray, uint32_t j) {
    return philox4x32_10(make_uint4(ray, j, g.oid, 8u * g.frame + 1u), g.key);
}
__device__ __forceinline__ float u24(uint32_t w) { return (float)(w >> 8) * TWO_M24; }
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float std, float& n0, float& n1) {   // as k_rng_fill
    const float u0 = ((float)(a >> 8) + 1.f) * TWO_M24, u1 = (float)(b >> 8) * TWO_M24;
    const float rad = sqrtf(-2.f * logf(u0)) * std;
    float sn, cs;
    sincospif(2.f * u1, &sn, &cs);
    n0 = rad * cs;
    n1 = rad * sn;
}

// EXACT: n_c2s == NC and n_bins == NB, every word index is a compile-time constant after unrolling.  Otherwise the bounds
// are upper bounds and each word is fetched on its own (one Philox block per word: correct, slow, not a shipped configuration).
template <int NC, int NB, bool EXACT>
struct CounterDraws {
    static constexpr int NP = (NB + 1) / 2;                                     // Box-Muller pairs
    static constexpr int NWC = 4 * ((NC + 2 + 3) / 4);                          // words of the blocks holding u_w, u_h, r_valid
    static constexpr int NWB = 4 * (((NC + NB > 2 * NP ? NC + NB : 2 * NP) + 3) / 4);   // words of the bin blocks
    const oo_sample_args& a;
    const Rng& g;
    uint32_t ray;
    uint32_t wc[EXACT ? NWC : 1], wb[EXACT ? NWB : 1];
    // EXACT: all blocks of the ray are evaluated once, before the (divergent) class branches of place_ray -- the three
    // classes read the same bin words, so a warp with mixed classes does not evaluate Philox once per class
    // blk0: block 0 of the ray when the caller has it already (the pixel draw), else nullptr
    __device__ __forceinline__ CounterDraws(const oo_sample_args& a_, const Rng& g_, uint32_t ray_, const uint4* blk0 = nullptr)
        : a(a_), g(g_), ray(ray_) {
        if (EXACT) {
#pragma unroll
            for (int j = 0; j < NWC / 4; ++j) {
                const uint4 r = (j == 0 && blk0 != nullptr) ? *blk0 : ray_block(g, ray, (uint32_t)j);
                wc[4 * j] = r.x; wc[4 * j + 1] = r.y; wc[4 * j + 2] = r.z; wc[4 * j + 3] = r.w;
            }
#pragma unroll
            for (int j = 0; j < NWB / 4; ++j) {
                const uint4 r = ray_block(g, ray, (uint32_t)(ray_b0(NC) / 4 + j));
                wb[4 * j] = r.x; wb[4 * j + 1] = r.y; wb[4 * j + 2] = r.z; wb[4 * j + 3] = r.w;
            }
        }
    }
    // generic instantiation: word `idx` of the ray on its own
    __device__ __forceinline__ uint32_t word(int idx) const {
        const uint4 r = ray_block(g, ray, (uint32_t)(idx >> 2));
        const int c = idx & 3;
        return c == 0 ? r.x : c == 1 ? r.y : c == 2 ? r.z : r.w;
    }
    __device__ __forceinline__ void invalid(float* u) const {
        const int nc = EXACT ? NC : a.n_c2s, S = nc + (EXACT ? NB : a.n_bins);
#pragma unroll
        for (int i = 0; i < NC + NB; ++i)
            if (i < S) u[i] = u24(EXACT ? wb[i] : word(ray_b0(nc) + i));
    }
    __device__ __forceinline__ void valid(float* uc) const {
        const int nc = EXACT ? NC : a.n_c2s;
#pragma unroll
        for (int i = 0; i < NC; ++i)
            if (i < nc) uc[i] = u24(EXACT ? wc[i + 2] : word(2 + i));
    }
    __device__ __forceinline__ void normal(float* ub, float std) const {
        const int nc = EXACT ? NC : a.n_c2s, nb = EXACT ? NB : a.n_bins;
#pragma unroll
        for (int p = 0; p < NP; ++p)
            if (2 * p < nb) {
                float n0, n1;
                if (EXACT) box_muller(wb[2 * p], wb[2 * p + 1], std, n0, n1);
                else box_muller(word(ray_b0(nc) + 2 * p), word(ray_b0(nc) + 2 * p + 1), std, n0, n1);
                ub[2 * p] = n0;
                if (2 * p + 1 < NB && 2 * p + 1 < nb) ub[2 * p + 1] = n1;
            }
    }
    __device__ __forceinline__ void other(float* ub) const {
        const int nc = EXACT ? NC : a.n_c2s, nb = EXACT ? NB : a.n_bins;
#pragma unroll
        for (int i = 0; i < NB; ++i)
            if (i < nb) ub[i] = u24(EXACT ? wb[i] : word(ray_b0(nc) + i));
    }
};

__device__ __forceinline__ Rng make_rng(const oo_sample_args& a, int obj) {
    Rng g;
    g.key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    g.oid = a.rng_mode ? (uint32_t)a.obj_ids[obj] : 0u;
    g.frame = a.frame;
    return g;
}

// ---- tape mode: one 1024-thread CTA per object, ranks by block scan --------------------------------------------------
template <int NC, int NB, bool EXACT>
__global__ void __launch_bounds__(NTH, 1) k_sample(const SampleK k) {
    const oo_sample_args& a = k.a;
    const int obj = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int n_rays = a.n_frames * a.n_samples;
    const int cpt = (n_rays + NTH - 1) / NTH;
    const int r_begin = min(tid * cpt, n_rays), r_end = min(r_begin + cpt, n_rays);
    __shared__ int sh_cnt[32][3];
    __shared__ float sh_max[32];
    __shared__ int sh_oob;
    if (tid == 0) sh_oob = 0;
    int n_inv = 0, n_obj = 0, n_oth = 0, oob = 0;
    float dmax = -INFINITY;
    auto pixel_of = [&](int ray) {
        const size_t ui = (size_t)obj * n_rays + ray;
        return pixel_from_uniforms(a, obj, (int)a.kf_ids[(size_t)obj * a.n_frames + ray / a.n_samples], a.u_w[ui], a.u_h[ui]);
    };
    for (int ray = r_begin; ray < r_end; ++ray) {
        float d;
        const int cls = gather_ray(a, obj, ray, n_rays, pixel_of(ray), d, oob);
        dmax = fmaxf(dmax, d);
        n_inv += cls == 0; n_obj += cls == 1; n_oth += cls == 2;
    }
    // block exclusive scan of the three class counters, block max of the sampled depths
    int s_inv = n_inv, s_obj = n_obj, s_oth = n_oth;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t0 = __shfl_up_sync(0xffffffffu, s_inv, o), t1 = __shfl_up_sync(0xffffffffu, s_obj, o),
                  t2 = __shfl_up_sync(0xffffffffu, s_oth, o);
        if (lane >= o) { s_inv += t0; s_obj += t1; s_oth += t2; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    if (lane == 31) { sh_cnt[wv][0] = s_inv; sh_cnt[wv][1] = s_obj; sh_cnt[wv][2] = s_oth; }
    if (lane == 0) sh_max[wv] = dmax;
    if (oob) atomicAdd(&sh_oob, oob);
    __syncthreads();
    if (wv == 0) {
        int c0 = sh_cnt[lane][0], c1 = sh_cnt[lane][1], c2 = sh_cnt[lane][2];
        float m = sh_max[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t0 = __shfl_up_sync(0xffffffffu, c0, o), t1 = __shfl_up_sync(0xffffffffu, c1, o),
                      t2 = __shfl_up_sync(0xffffffffu, c2, o);
            if (lane >= o) { c0 += t0; c1 += t1; c2 += t2; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        sh_cnt[lane][0] = c0 - sh_cnt[lane][0];   // exclusive warp offsets
        sh_cnt[lane][1] = c1 - sh_cnt[lane][1];
        sh_cnt[lane][2] = c2 - sh_cnt[lane][2];
        sh_max[lane] = m;
    }
    __syncthreads();
    int rk_inv = sh_cnt[wv][0] + s_inv - n_inv, rk_obj = sh_cnt[wv][1] + s_obj - n_obj,
        rk_oth = sh_cnt[wv][2] + s_oth - n_oth;
    const float max_bound = sh_max[0];                                             // vmap.py:489
    if (tid == 0 && a.oob_count && sh_oob) atomicAdd(a.oob_count, sh_oob);
    for (int ray = r_begin; ray < r_end; ++ray) {
        const size_t orow = (size_t)obj * n_rays + ray;
        const int S = a.n_c2s + a.n_bins;
        const RayPix p = pixel_of(ray);
        const float d = a.gt_depth[orow];
        const int state = a.labels[orow];
        const size_t base = (size_t)obj * n_rays;
        const bool by_rank = a.tape_by_rank != 0;
        const TapeDraws<NC, NB> dr{a, base + (by_rank ? rk_inv : ray), base + (by_rank ? ray - rk_inv : ray),
                                   base + (by_rank ? rk_obj : ray), base + (by_rank ? rk_oth : ray)};
        place_ray<NC, NB, EXACT>(k, obj, p.kf, p.iw, p.ih, d, state, max_bound, dr, a.z + orow * S, a.pcs + orow * S * 3);
        if (d <= a.min_bound) ++rk_inv;
        else if (state == 1) ++rk_obj;
        else ++rk_oth;
    }
}

// ---- counter-RNG mode: no ranks are needed, so every ray is independent.  ONE pass does everything for the rays with a
// valid depth: keyframe word + block 0 of the ray -> pixel -> gathers -> per-ray outputs -> the remaining blocks -> sample
// placement into the CTA's staging tile (z [256][S], points [256][S][3]), which is a contiguous range of the outputs and
// leaves with 128-bit, fully coalesced stores (a thread writing its own 40 B / 120 B rows directly costs one 32-byte sector
// per 4-byte store).  Only the invalid-depth rays (a few per cent) need the batch-max depth of their object (vmap.py:489,
// quirk 6): they are appended to a per-object list and placed by a small second launch once the maximum (non-negative
// floats order like their bit patterns: atomicMax on the bits) is complete.
template <int NC, int NB, bool EXACT>
__global__ void __launch_bounds__(256) k_sample_main(const SampleK k, int* __restrict__ max_bits, int* __restrict__ inv_cnt,
                                                     int2* __restrict__ inv_list) {
    extern __shared__ __align__(16) float tile[];
    const oo_sample_args& a = k.a;
    const int obj = blockIdx.y, n_rays = a.n_frames * a.n_samples, S = EXACT ? NC + NB : a.n_c2s + a.n_bins;
    const int ray0 = blockIdx.x * blockDim.x, ray = ray0 + threadIdx.x;
    const int n_here = min((int)blockDim.x, n_rays - ray0);
    float* zt = tile;
    float* pt = tile + (size_t)blockDim.x * S;
    const Rng g = make_rng(a, obj);
    float d = 0.f;
    int oob = 0;
    if (ray < n_rays) {
        const int f = ray / a.n_samples, nk = a.n_keyframes[obj];
        int kf;
        if (nk > 2 && f >= a.n_frames - 2) kf = a.latest[2 * obj + (f - (a.n_frames - 2))];       // vmap.py:398-400
        else kf = min((int)(g.uniform(0, (uint64_t)f) * (float)nk), nk - 1);
        const uint4 r0 = ray_block(g, (uint32_t)ray, 0u);
        const RayPix p = pixel_from_uniforms(a, obj, kf, u24(r0.x), u24(r0.y));
        const int cls = gather_ray(a, obj, ray, n_rays, p, d, oob);
        if (cls == 0) {
            const int slot = atomicAdd(inv_cnt + obj, 1);
            inv_list[(size_t)obj * n_rays + slot] = make_int2(ray, p.kf | (p.iw << 5) | (p.ih << 16));
        } else {
            const CounterDraws<NC, NB, EXACT> dr(a, g, (uint32_t)ray, &r0);
            place_ray<NC, NB, EXACT>(k, obj, p.kf, p.iw, p.ih, d, cls, 0.f, dr, zt + threadIdx.x * S, pt + threadIdx.x * S * 3);
        }
    }
    d = fmaxf(d, 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_bits + obj, __float_as_int(d));
    if (oob && a.oob_count) atomicAdd(a.oob_count, oob);
    __syncthreads();
    const size_t row0 = (size_t)obj * n_rays + ray0;
    float* zg = a.z + row0 * S;
    float* pg = a.pcs + row0 * S * 3;
    const int nz = n_here * S, np = 3 * nz;
    if (((reinterpret_cast<uintptr_t>(zg) | reinterpret_cast<uintptr_t>(pg)) & 15) == 0 && ((blockDim.x * S) & 3) == 0) {
        for (int i = threadIdx.x; i < nz / 4; i += blockDim.x) reinterpret_cast<float4*>(zg)[i] = reinterpret_cast<const float4*>(zt)[i];
        for (int i = 4 * (nz / 4) + threadIdx.x; i < nz; i += blockDim.x) zg[i] = zt[i];
        for (int i = threadIdx.x; i < np / 4; i += blockDim.x) reinterpret_cast<float4*>(pg)[i] = reinterpret_cast<const float4*>(pt)[i];
        for (int i = 4 * (np / 4) + threadIdx.x; i < np; i += blockDim.x) pg[i] = pt[i];
    } else {
        for (int i = threadIdx.x; i < nz; i += blockDim.x) zg[i] = zt[i];
        for (int i = threadIdx.x; i < np; i += blockDim.x) pg[i] = pt[i];
    }
}

// the invalid-depth rays of every object: stratified bins over [min_bound, batch max] (vmap.py:493-498), rows written in place
template <int NC, int NB, bool EXACT>
__global__ void __launch_bounds__(128) k_sample_fix(const SampleK k, const int* __restrict__ max_bits, const int* __restrict__ inv_cnt,
                                                    const int2* __restrict__ inv_list) {
    const oo_sample_args& a = k.a;
    const int obj = blockIdx.y, n_rays = a.n_frames * a.n_samples, S = EXACT ? NC + NB : a.n_c2s + a.n_bins;
    const int n = inv_cnt[obj];
    const Rng g = make_rng(a, obj);
    const float max_bound = __int_as_float(max_bits[obj]);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int2 e = inv_list[(size_t)obj * n_rays + idx];
        const size_t o = (size_t)obj * n_rays + e.x;
        const CounterDraws<NC, NB, EXACT> dr(a, g, (uint32_t)e.x);
        place_ray<NC, NB, EXACT>(k, obj, e.y & 31, (e.y >> 5) & 2047, e.y >> 16, a.gt_depth[o], 0, max_bound, dr, a.z + o * S,
                                 a.pcs + o * S * 3);
    }
}

// tapes of the ray-blocked stream: out[o][row][w], w < row_words, = word w of row `row` (uniform, or the Box-Muller normal of
// its pair) -- the values rng_mode 1 draws in-kernel for ray = row
__global__ void k_rng_fill_rows(uint64_t seed, uint32_t frame, const int32_t* __restrict__ obj_ids, int n_rows, int row_words,
                                int kind, float std, float* __restrict__ out) {
    const int o = blockIdx.y, nblk = (row_words + 3) / 4;
    const uint32_t oid = (uint32_t)obj_ids[o];
    const long long n = (long long)n_rows * nblk;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const uint32_t row = (uint32_t)(e / nblk), j = (uint32_t)(e % nblk);
        const uint4 r = philox4x32_10(make_uint4(row, j, oid, frame), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        float v[4];
        if (kind == 0) {
            v[0] = u24(r.x); v[1] = u24(r.y); v[2] = u24(r.z); v[3] = u24(r.w);
        } else {
            box_muller(r.x, r.y, std, v[0], v[1]);
            box_muller(r.z, r.w, std, v[2], v[3]);
        }
        float* dst = out + ((size_t)o * n_rows + row) * row_words + 4 * j;
        for (int i = 0; i < 4; ++i)
            if (4 * (int)j + i < row_words) dst[i] = v[i];
    }
}

__global__ void k_rng_fill(uint64_t seed, uint32_t frame, const int32_t* __restrict__ obj_ids, int64_t per_obj, int kind,
                           float std, float* __restrict__ out) {
    const int o = blockIdx.y;
    const uint32_t oid = (uint32_t)obj_ids[o];
    const int64_t n4 = (per_obj + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), oid, frame),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        float v[4];
        if (kind == 0) {
            v[0] = (float)(r.x >> 8) * 5.9604644775390625e-8f; v[1] = (float)(r.y >> 8) * 5.9604644775390625e-8f;
            v[2] = (float)(r.z >> 8) * 5.9604644775390625e-8f; v[3] = (float)(r.w >> 8) * 5.9604644775390625e-8f;
        } else {
            const float u0 = ((float)(r.x >> 8) + 1.f) * 5.9604644775390625e-8f, u1 = (float)(r.y >> 8) * 5.9604644775390625e-8f;
            const float u2 = ((float)(r.z >> 8) + 1.f) * 5.9604644775390625e-8f, u3 = (float)(r.w >> 8) * 5.9604644775390625e-8f;
            const float ra = sqrtf(-2.f * logf(u0)) * std, rb = sqrtf(-2.f * logf(u2)) * std;
            float s, c;
            sincospif(2.f * u1, &s, &c);
            v[0] = ra * c; v[1] = ra * s;
            sincospif(2.f * u3, &s, &c);
            v[2] = rb * c; v[3] = rb * s;
        }
        float* dst = out + (size_t)o * per_obj + 4 * q;
        for (int i = 0; i < 4; ++i)
            if (4 * q + i < per_obj) dst[i] = v[i];
    }
}

// ---- shared keyframe store: ONE copy of the new frame, whatever the number of objects that see it.
// A pixel becomes {r | g << 8 | b << 16, instance id} (8 bytes: one load in K2 gives colour and pixel state) + depth.
__global__ void __launch_bounds__(256) k_store_frame(const oo_store_args a) {
    const size_t n_pix = (size_t)a.W * a.H;
    int2* dst = reinterpret_cast<int2*>(a.store_rgbi) + (size_t)a.slot * n_pix;
    float* dd = a.store_depth + (size_t)a.slot * n_pix;
    const size_t n4 = n_pix / 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(a.rgb) | reinterpret_cast<uintptr_t>(a.depth) | reinterpret_cast<uintptr_t>(a.inst)) & 15) == 0;
    if (vec) {
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
            const uint32_t* rp = reinterpret_cast<const uint32_t*>(a.rgb) + 3 * q;         // 4 pixels = 12 bytes
            const uint32_t w0 = rp[0], w1 = rp[1], w2 = rp[2];
            const float4 d4 = reinterpret_cast<const float4*>(a.depth)[q];
            const int4 i4 = reinterpret_cast<const int4*>(a.inst)[q];
            const int c0 = (int)(w0 & 0xffffffu), c1 = (int)((w0 >> 24) | ((w1 & 0xffffu) << 8)),
                      c2 = (int)((w1 >> 16) | ((w2 & 0xffu) << 16)), c3 = (int)(w2 >> 8);
            reinterpret_cast<int4*>(dst)[2 * q] = make_int4(c0, i4.x, c1, i4.y);
            reinterpret_cast<int4*>(dst)[2 * q + 1] = make_int4(c2, i4.z, c3, i4.w);
            reinterpret_cast<float4*>(dd)[q] = d4;
        }
    }
    for (size_t pix = (vec ? 4 * n4 : 0) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < n_pix; pix += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)a.rgb[3 * pix] | ((int)a.rgb[3 * pix + 1] << 8) | ((int)a.rgb[3 * pix + 2] << 16);
        dst[pix] = make_int2(c, a.inst[pix]);
        dd[pix] = a.depth[pix];
    }
    if (blockIdx.x == 0 && threadIdx.x < 16) a.store_twc[16 * (size_t)a.slot + threadIdx.x] = a.t_wc[threadIdx.x];
}

}  // namespace

extern "C" int oo_store_frame(const oo_store_args* a, void* stream) {
    OO_REQUIRE(a && a->rgb && a->depth && a->inst && a->t_wc && a->store_rgbi && a->store_depth && a->store_twc,
               "oo_store_frame: null argument");
    OO_REQUIRE(a->W > 0 && a->H > 0 && a->slot >= 0, "oo_store_frame: bad shape / slot");
    const size_t n4 = ((size_t)a->W * a->H + 3) / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_store_frame<<<blocks, 256, 0, (cudaStream_t)stream>>>(*a);
    OO_LAUNCH_CHECK();
    return 0;
}

// ---- part-feature rows of a frame, only where this rank's objects can sample them.  src is the frame's [pw][ph][C] tensor in
// PINNED HOST memory (device-addressable under unified addressing): the kernel reads the cells of each box straight over the
// host link and writes them to the same cells of the resident table slot -- a rank of a sharded run moves the boxes' worth
// of the 66.8 MB, not all of it.  Boxes may overlap (a cell is then written twice with the same values).
namespace {
constexpr int GATHER_MAX_BOXES = 192;
struct GatherBoxes { int n, pw, ph, c4; int box[GATHER_MAX_BOXES][4]; };
__global__ void __launch_bounds__(128) k_gather_part(const float4* __restrict__ src, float4* __restrict__ dst, const GatherBoxes b) {
    const int w0 = b.box[blockIdx.x][0], w1 = b.box[blockIdx.x][1], h0 = b.box[blockIdx.x][2], h1 = b.box[blockIdx.x][3];
    const int nh = h1 - h0 + 1, cells = (w1 - w0 + 1) * nh;
    for (int c = blockIdx.y; c < cells; c += gridDim.y) {
        const size_t cell = (size_t)(w0 + c / nh) * b.ph + (h0 + c % nh);
        for (int q = threadIdx.x; q < b.c4; q += blockDim.x) dst[cell * b.c4 + q] = src[cell * b.c4 + q];
    }
}
}  // namespace

extern "C" int oo_gather_part_rows(const float* src_host, float* dst, int pw, int ph, int n_feat, const int32_t* boxes, int n_boxes,
                                   void* stream) {
    OO_REQUIRE(src_host && dst && boxes, "oo_gather_part_rows: null argument");
    OO_REQUIRE(pw > 0 && ph > 0 && n_feat > 0 && (n_feat & 3) == 0, "oo_gather_part_rows: bad shape");
    OO_REQUIRE(n_boxes >= 1 && n_boxes <= GATHER_MAX_BOXES, "oo_gather_part_rows: 1 .. %d boxes per call", GATHER_MAX_BOXES);
    GatherBoxes b;
    b.n = n_boxes; b.pw = pw; b.ph = ph; b.c4 = n_feat / 4;
    for (int i = 0; i < n_boxes; ++i) {
        for (int k = 0; k < 4; ++k) b.box[i][k] = boxes[4 * i + k];
        OO_REQUIRE(0 <= b.box[i][0] && b.box[i][0] <= b.box[i][1] && b.box[i][1] < pw && 0 <= b.box[i][2] && b.box[i][2] <= b.box[i][3] &&
                       b.box[i][3] < ph, "oo_gather_part_rows: box %d outside the [%d x %d] cell grid", i, pw, ph);
    }
    k_gather_part<<<dim3(n_boxes, 16), 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(src_host),
                                                                      reinterpret_cast<float4*>(dst), b);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_sample_rays(const oo_sample_args* a, void* stream) {
    OO_REQUIRE(a, "oo_sample_rays: null args");
    OO_REQUIRE(a->n_obj > 0 && a->n_frames > 0 && a->n_samples > 0, "oo_sample_rays: empty request");
    const int S = a->n_c2s + a->n_bins;
    OO_REQUIRE(a->n_c2s >= 1 && a->n_bins >= 1 && a->n_c2s <= 16 && a->n_bins <= 16, "oo_sample_rays: need 1 <= n_c2s, n_bins <= 16");
    OO_REQUIRE(a->rays_dir, "oo_sample_rays: null input");
    OO_REQUIRE(a->kf_cap >= 1 && a->kf_cap <= 32, "oo_sample_rays: kf_cap (keyframe_buffer_size) must be in [1, 32]");
    if (a->store_rgbi != nullptr)
        OO_REQUIRE(a->store_depth && a->store_twc && a->slot_frame && a->slot_bbox && a->obj_ids,
                   "oo_sample_rays: the shared keyframe store needs store_depth / store_twc / slot_frame / slot_bbox / obj_ids");
    else
        OO_REQUIRE(a->rgbs && a->depth && a->t_wc && a->bbox, "oo_sample_rays: null keyframe ring table");
    if (a->rng_mode) {
        OO_REQUIRE(a->obj_ids && a->n_keyframes && a->latest, "oo_sample_rays: rng_mode needs obj_ids / n_keyframes / latest");
        OO_REQUIRE(!a->tape_by_rank, "oo_sample_rays: the counter RNG is indexed by ray, not by rank");
    } else {
        OO_REQUIRE(a->kf_ids && a->u_w && a->u_h, "oo_sample_rays: null keyframe / pixel tape");
        OO_REQUIRE(a->r_invalid && a->r_valid && a->r_normal && a->r_other, "oo_sample_rays: null RNG tape");
    }
    OO_REQUIRE(a->gt_rgb && a->gt_depth && a->valid && a->labels && a->pcs && a->z, "oo_sample_rays: null output");
    OO_REQUIRE(a->lin_s_host && a->lin_c2s_host && a->lin_bins_host, "oo_sample_rays: null linspace table");
    OO_REQUIRE(!a->feat_row || (a->part_frame && a->part_down > 0 && a->pw > 0 && a->ph > 0),
               "oo_sample_rays: part features requested without part_frame / part map size");
    SampleK k;
    k.a = *a;
    for (int i = 0; i <= S; ++i) k.lin_s[i] = a->lin_s_host[i];
    for (int i = 0; i <= a->n_c2s; ++i) k.lin_c[i] = a->lin_c2s_host[i];
    for (int i = 0; i <= a->n_bins; ++i) k.lin_b[i] = a->lin_bins_host[i];
    if (a->rng_mode) {
        OO_REQUIRE(a->min_bound >= 0.f, "oo_sample_rays: the parallel path assumes min_bound >= 0");
        OO_REQUIRE(a->W <= 2048 && a->H <= 32768, "oo_sample_rays: the packed pixel needs W <= 2048, H <= 32768");
        const int n_rays = a->n_frames * a->n_samples;
        // caller-owned scratch: per-object batch max and invalid-ray count, then the per-object invalid-ray lists
        const size_t need = (size_t)a->n_obj * (2 + 2 * (size_t)n_rays);
        OO_REQUIRE(a->scratch && (size_t)a->scratch_ints >= need, "oo_sample_rays: rng_mode 1 needs scratch of n_obj * (2 + 2 n_rays) ints");
        int* scratch = a->scratch;
        int* max_bits = scratch;
        int* inv_cnt = scratch + a->n_obj;
        int2* inv_list = reinterpret_cast<int2*>(scratch + 2 * (size_t)a->n_obj);
        OO_CUDA(cudaMemsetAsync(scratch, 0, 2 * (size_t)a->n_obj * sizeof(int), (cudaStream_t)stream));
        const dim3 grid((n_rays + 255) / 256, a->n_obj);
        const size_t tile_bytes = (size_t)256 * S * 4 * sizeof(float);        // z [256][S] + points [256][S][3]; <= 128 KB (S <= 32)
        // one instantiation per launch (exact for the shipped 1 + 9 and 5 + 9 bins, upper bounds otherwise): per-ray arrays in
        // registers and a code size the instruction cache holds
#define OO_LAUNCH_S(NC_, NB_, EX_)                                                                                                     \
        do {                                                                                                                           \
            OO_CUDA(cudaFuncSetAttribute(k_sample_main<NC_, NB_, EX_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes)); \
            k_sample_main<NC_, NB_, EX_><<<grid, 256, tile_bytes, (cudaStream_t)stream>>>(k, max_bits, inv_cnt, inv_list);             \
            OO_LAUNCH_CHECK();                                                                                                         \
            k_sample_fix<NC_, NB_, EX_><<<dim3(8, a->n_obj), 128, 0, (cudaStream_t)stream>>>(k, max_bits, inv_cnt, inv_list);          \
        } while (0)
        if (a->n_c2s == 1 && a->n_bins == 9) OO_LAUNCH_S(1, 9, true);
        else if (a->n_c2s == 5 && a->n_bins == 9) OO_LAUNCH_S(5, 9, true);
        else if (a->n_c2s <= 8 && a->n_bins <= 12) OO_LAUNCH_S(8, 12, false);
        else OO_LAUNCH_S(16, 16, false);
#undef OO_LAUNCH_S
        OO_LAUNCH_CHECK();
        return 0;
    }
    if (a->n_c2s == 1 && a->n_bins == 9) k_sample<1, 9, true><<<a->n_obj, NTH, 0, (cudaStream_t)stream>>>(k);
    else if (a->n_c2s == 5 && a->n_bins == 9) k_sample<5, 9, true><<<a->n_obj, NTH, 0, (cudaStream_t)stream>>>(k);
    else if (a->n_c2s <= 8 && a->n_bins <= 12) k_sample<8, 12, false><<<a->n_obj, NTH, 0, (cudaStream_t)stream>>>(k);
    else k_sample<16, 16, false><<<a->n_obj, NTH, 0, (cudaStream_t)stream>>>(k);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_rng_fill_rows(uint64_t seed, uint32_t frame, const int32_t* obj_ids, int n_obj, int n_rows, int row_words,
                                int kind, float std, float* out, void* stream) {
    OO_REQUIRE(obj_ids && out && n_obj > 0 && n_rows > 0 && row_words > 0 && (kind == 0 || kind == 1), "oo_rng_fill_rows: bad argument");
    const long long n = (long long)n_rows * ((row_words + 3) / 4);
    int gx = (int)((n + 255) / 256);
    if (gx > 148 * 8) gx = 148 * 8;
    k_rng_fill_rows<<<dim3(gx, n_obj), 256, 0, (cudaStream_t)stream>>>(seed, frame, obj_ids, n_rows, row_words, kind, std, out);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_rng_fill(uint64_t seed, uint32_t frame, const int32_t* obj_ids, int n_obj, int64_t per_obj, int kind,
                           float std, float* out, void* stream) {
    OO_REQUIRE(obj_ids && out && n_obj > 0 && per_obj > 0 && (kind == 0 || kind == 1), "oo_rng_fill: bad argument");
    const int64_t n4 = (per_obj + 3) / 4;
    int gx = (int)((n4 + 255) / 256);
    if (gx > 148 * 8) gx = 148 * 8;
    k_rng_fill<<<dim3(gx, n_obj), 256, 0, (cudaStream_t)stream>>>(seed, frame, obj_ids, per_obj, kind, std, out);
    OO_LAUNCH_CHECK();
    return 0;
}
