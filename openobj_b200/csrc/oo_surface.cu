// The stand-alone helpers of the reference surface that the fused kernels (K2 sampling, K5 render) absorb on the hot path:
// utils.ray_box_intersection (objnerf/utils.py:309-319), utils.origin_dirs_W (utils.py:324-336), utils.stratified_bins
// (utils.py:342-379), utils.normal_bins_sampling (utils.py:382-397) and the point placement
// origins + dirs * z of sample_3d_points / sample_points_bbox (vmap.py:548-549, trainer.py:176-177).  Callers that compose
// them by hand (Trainer.sample_points_bbox, sceneObject.sample_3d_points) get the same arithmetic, in the reference's
// operation order, from these small element-wise kernels; random draws are passed in (the host surface draws them with
// torch.rand / normal_ exactly where the reference does).
#include "../../include/openobj_b200.h"
#include "oo_common.cuh"

namespace {

constexpr int SURF_MAX_BINS = 256;

__global__ void k_ray_box(const float* __restrict__ org, const float* __restrict__ dir, float3 bmin, float3 bmax, long long n,
                          float* __restrict__ near_, float* __restrict__ far_, uint8_t* __restrict__ hit) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const float lo[3] = {bmin.x, bmin.y, bmin.z}, hi[3] = {bmax.x, bmax.y, bmax.z};
        float nr = -INFINITY, fr = INFINITY;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float o = org[3 * r + a], d = dir[3 * r + a];
            const float t0 = __fdiv_rn(__fsub_rn(lo[a], o), d), t1 = __fdiv_rn(__fsub_rn(hi[a], o), d);   // utils.py:310-311
            nr = fmaxf(nr, fminf(t0, t1));                                                             // :312,314
            fr = fminf(fr, fmaxf(t0, t1));                                                             // :313,315
        }
        near_[r] = nr;
        far_[r] = fr;
        hit[r] = (nr <= fr && fr > 0.f) ? 1 : 0;                                                       // :316-318
    }
}

// dirs_W[b][i] = R_b dirs_C[b][i], origins[b] = T_b[:3, 3]
__global__ void k_origin_dirs(const float* __restrict__ T, const float* __restrict__ dc, int B, int n, float* __restrict__ org,
                              float* __restrict__ dw) {
    const long long total = (long long)B * n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(e / n);
        const float* t = T + 16 * (size_t)b;
        const float x = dc[3 * e], y = dc[3 * e + 1], z = dc[3 * e + 2];
        dw[3 * e + 0] = t[0] * x + t[1] * y + t[2] * z;
        dw[3 * e + 1] = t[4] * x + t[5] * y + t[6] * z;
        dw[3 * e + 2] = t[8] * x + t[9] * y + t[10] * z;
        if (e % n == 0) {
            org[3 * b + 0] = t[3]; org[3 * b + 1] = t[7]; org[3 * b + 2] = t[11];
        }
    }
}

// z[r][k] = (range_r * lin[k] + min_r) + u[r][k] * (range_r / n_bins): separate roundings as the reference's tensor expression
__global__ void k_stratified(const float* __restrict__ u, const float* __restrict__ mn, const float* __restrict__ mx, float mn_s,
                             float mx_s, const float* __restrict__ lin, long long n, int nb, float* __restrict__ z) {
    const long long total = n * nb;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / nb;
        const int k = (int)(e - r * nb);
        const float lo = mn ? mn[r] : mn_s, hi = mx ? mx[r] : mx_s;
        const float range = __fsub_rn(hi, lo);                                   // utils.py:357
        const float lower = __fadd_rn(__fmul_rn(range, lin[k]), lo);             // :359-360
        const float blen = __fdiv_rn(range, (float)nb);                          // :368
        z[e] = __fadd_rn(lower, __fmul_rn(u[e], blen));                          // :372-376
    }
}

// one thread per ray: ascending sort of its n_bins draws, clip to +-delta, + depth (utils.py:391-393)
__global__ void k_normal_bins(const float* __restrict__ draws, const float* __restrict__ depth, long long n, int nb, float delta,
                              float* __restrict__ z) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        float* row = z + r * nb;
        for (int i = 0; i < nb; ++i) {                 // insertion sort into the output row
            const float x = draws[r * nb + i];
            int j = i - 1;
            while (j >= 0 && row[j] > x) {
                row[j + 1] = row[j];
                --j;
            }
            row[j + 1] = x;
        }
        const float d = depth[r];
        for (int i = 0; i < nb; ++i) row[i] = __fadd_rn(d, fminf(fmaxf(row[i], -delta), delta));
    }
}

// pcs[r][i] = origins[r] + dirs[r] * zz[r][i] - center, zz = z (mid == 0, S points) or the midpoints 0.5 (z[i+1] + z[i])
// (mid == 1, S - 1 points: trainer.py:175-177)
__global__ void k_ray_points(const float* __restrict__ org, const float* __restrict__ dir, const float* __restrict__ z, long long n,
                             int S, int mid, float3 center, float* __restrict__ z_mid, float* __restrict__ pcs) {
    const int So = mid ? S - 1 : S;
    const long long total = n * So;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / So;
        const int i = (int)(e - r * So);
        float zz = z[r * S + i];
        if (mid) {
            zz = __fmul_rn(0.5f, __fadd_rn(z[r * S + i + 1], zz));
            if (z_mid) z_mid[e] = zz;
        }
        pcs[3 * e + 0] = __fsub_rn(__fadd_rn(org[3 * r + 0], __fmul_rn(dir[3 * r + 0], zz)), center.x);
        pcs[3 * e + 1] = __fsub_rn(__fadd_rn(org[3 * r + 1], __fmul_rn(dir[3 * r + 1], zz)), center.y);
        pcs[3 * e + 2] = __fsub_rn(__fadd_rn(org[3 * r + 2], __fmul_rn(dir[3 * r + 2], zz)), center.z);
    }
}

// T[r][i] = occ[r][i] * prod_{j<i} (1 - occ[r][j] + 1e-10), one thread per ray (render_rays.py:32-54)
__global__ void k_termination(const float* __restrict__ occ, long long n, int S, float* __restrict__ T) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        float p = 1.f;
        for (int i = 0; i < S; ++i) {
            const float o = occ[r * S + i];
            T[r * S + i] = __fmul_rn(o, p);
            p = __fmul_rn(p, __fadd_rn(__fsub_rn(1.f, o), 1e-10f));
        }
    }
}

// out[r][c] = sum_i T[r][i] * vals[r][i][c] in sample order (render_rays.py:56-63)
__global__ void k_render_sum(const float* __restrict__ T, const float* __restrict__ vals, long long n, int S, int C,
                             float* __restrict__ out) {
    const long long total = n * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / C;
        const int c = (int)(e - r * C);
        float acc = 0.f;
        for (int i = 0; i < S; ++i) acc = __fadd_rn(acc, __fmul_rn(T[r * S + i], vals[(r * S + i) * C + c]));
        out[e] = acc;
    }
}

int grid_for(long long n) {
    long long b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace

extern "C" int oo_ray_box(const float* origins, const float* dirs, const float* bounds_min, const float* bounds_max, long long n,
                          float* near_out, float* far_out, uint8_t* hit_out, void* stream) {
    OO_REQUIRE(origins && dirs && bounds_min && bounds_max && near_out && far_out && hit_out && n > 0, "oo_ray_box: bad argument");
    const float3 lo = make_float3(bounds_min[0], bounds_min[1], bounds_min[2]);      // HOST 3-vectors
    const float3 hi = make_float3(bounds_max[0], bounds_max[1], bounds_max[2]);
    k_ray_box<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(origins, dirs, lo, hi, n, near_out, far_out, hit_out);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_origin_dirs(const float* t_wc, const float* dirs_c, int n_poses, int n_per_pose, float* origins, float* dirs_w,
                              void* stream) {
    OO_REQUIRE(t_wc && dirs_c && origins && dirs_w && n_poses > 0 && n_per_pose > 0, "oo_origin_dirs: bad argument");
    k_origin_dirs<<<grid_for((long long)n_poses * n_per_pose), 256, 0, (cudaStream_t)stream>>>(t_wc, dirs_c, n_poses, n_per_pose,
                                                                                              origins, dirs_w);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_stratified_bins(const float* u, const float* min_depth, const float* max_depth, float min_scalar,
                                  float max_scalar, const float* lin, long long n_rays, int n_bins, float* z, void* stream) {
    OO_REQUIRE(u && lin && z && n_rays > 0 && n_bins > 0 && n_bins <= SURF_MAX_BINS, "oo_stratified_bins: bad argument");
    k_stratified<<<grid_for(n_rays * n_bins), 256, 0, (cudaStream_t)stream>>>(u, min_depth, max_depth, min_scalar, max_scalar, lin,
                                                                            n_rays, n_bins, z);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_normal_bins(const float* draws, const float* depth, long long n_rays, int n_bins, float delta, float* z,
                              void* stream) {
    OO_REQUIRE(draws && depth && z && n_rays > 0 && n_bins > 0 && n_bins <= SURF_MAX_BINS, "oo_normal_bins: bad argument");
    k_normal_bins<<<grid_for(n_rays), 256, 0, (cudaStream_t)stream>>>(draws, depth, n_rays, n_bins, delta, z);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_ray_points(const float* origins, const float* dirs, const float* z, long long n_rays, int n_samp, int midpoints,
                             const float* center, float* z_mid_out, float* pcs, void* stream) {
    OO_REQUIRE(origins && dirs && z && pcs && n_rays > 0 && n_samp > (midpoints ? 1 : 0), "oo_ray_points: bad argument");
    const float3 c = center ? make_float3(center[0], center[1], center[2]) : make_float3(0.f, 0.f, 0.f);     // HOST 3-vector
    k_ray_points<<<grid_for(n_rays * n_samp), 256, 0, (cudaStream_t)stream>>>(origins, dirs, z, n_rays, n_samp, midpoints, c,
                                                                            z_mid_out, pcs);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_termination(const float* occupancy, long long n_rays, int n_samp, float* termination, void* stream) {
    OO_REQUIRE(occupancy && termination && n_rays > 0 && n_samp > 0, "oo_termination: bad argument");
    k_termination<<<grid_for(n_rays), 256, 0, (cudaStream_t)stream>>>(occupancy, n_rays, n_samp, termination);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_render_sum(const float* termination, const float* vals, long long n_rays, int n_samp, int n_chan, float* out,
                             void* stream) {
    OO_REQUIRE(termination && vals && out && n_rays > 0 && n_samp > 0 && n_chan > 0, "oo_render_sum: bad argument");
    k_render_sum<<<grid_for(n_rays * n_chan), 256, 0, (cudaStream_t)stream>>>(termination, vals, n_rays, n_samp, n_chan, out);
    OO_LAUNCH_CHECK();
    return 0;
}
