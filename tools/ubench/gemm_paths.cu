// Microbenchmark (evidence for DESIGN.md "tensor cores"): the cat_layer forward of one tile,
//   Y[32][100] = relu(W[32][120] X[120][100] + b),  operands resident in shared memory,
// (A) as the production FP32 path (gemm_fwd of oo_tile.h: 4x4 register tiles, 128-bit shared loads), and
// (B) on the legacy tensor path with error-compensated 3xTF32 mma.sync.m16n8k8 (a = a_hi + a_lo, three MMAs).
// One CTA per SM, 256 threads, the GEMM repeated `reps` times; reports cycles per GEMM and the max abs difference.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../openobj_b200/csrc/oo_tile.h"
using namespace oo;

constexpr int K = KP_CAT, WS = WS_CAT, PSB = 120;      // variant B: X row stride 120 (conflict-free fragments, room for 112 points)

__device__ __forceinline__ unsigned tf32_hi(float x) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Y[j][p] for p < 112 (7 m-tiles of 16 points), j < 32 (4 n-tiles of 8); warp w < 7 owns m-tile w
__device__ void gemm_fwd_tf32x3(int tid, const float* __restrict__ W, const float* __restrict__ bias,
                                const float* __restrict__ X, float* __restrict__ Y) {
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    if (warp >= 7) return;
    const int p0 = 16 * warp;
    float d[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const float b0 = bias[8 * n + 2 * t], b1 = bias[8 * n + 2 * t + 1];
        d[n][0] = b0; d[n][1] = b1; d[n][2] = b0; d[n][3] = b1;
    }
#pragma unroll 3
    for (int k0 = 0; k0 < K; k0 += 8) {
        float af[4] = {X[(k0 + t) * PSB + p0 + g], X[(k0 + t) * PSB + p0 + g + 8], X[(k0 + t + 4) * PSB + p0 + g],
                       X[(k0 + t + 4) * PSB + p0 + g + 8]};
        unsigned ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ah[i] = tf32_hi(af[i]); al[i] = tf32_hi(af[i] - __uint_as_float(ah[i])); }
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const float bf[2] = {W[(8 * n + g) * WS + k0 + t], W[(8 * n + g) * WS + k0 + t + 4]};
            unsigned bh[2], bl[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) { bh[i] = tf32_hi(bf[i]); bl[i] = tf32_hi(bf[i] - __uint_as_float(bh[i])); }
            mma_tf32(d[n], al, bh);
            mma_tf32(d[n], ah, bl);
            mma_tf32(d[n], ah, bh);
        }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {       // C fragment: rows g, g+8 (points), cols 2t, 2t+1 (outputs)
        const int j = 8 * n + 2 * t;
        Y[j * PSB + p0 + g] = fmaxf(d[n][0], 0.f);
        Y[(j + 1) * PSB + p0 + g] = fmaxf(d[n][1], 0.f);
        Y[j * PSB + p0 + g + 8] = fmaxf(d[n][2], 0.f);
        Y[(j + 1) * PSB + p0 + g + 8] = fmaxf(d[n][3], 0.f);
    }
}

__global__ void __launch_bounds__(256, 1) k_bench(int variant, int reps, const float* Wg, const float* Xg, float* Yg, long long* cyc) {
    extern __shared__ __align__(16) float sm[];
    float* W = sm;                       // [32][124]
    float* bias = W + 32 * WS;           // [32]
    float* X = bias + 32;                // [120][104]
    float* Y = X + K * PSB;              // [32][104]
    const int tid = threadIdx.x;
    for (int i = tid; i < 32 * WS; i += 256) W[i] = (i % WS) < K ? Wg[(i / WS) * K + (i % WS)] : 0.f;
    for (int i = tid; i < 32; i += 256) bias[i] = 0.01f * i;
    for (int i = tid; i < K * PSB; i += 256) X[i] = (i % PSB) < P ? Xg[(i / PSB) * P + (i % PSB)] : 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (variant == 0) {
            // production path expects row stride PS = 100: use a view with the same data laid out at stride 100
            gemm_fwd<K, WS, true>(tid, W, bias, X, Y);      // X, Y read with stride PS (=100): timing only differs by layout
        } else {
            gemm_fwd_tf32x3(tid, W, bias, X, Y);
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cyc[variant] = (t1 - t0) / reps;
    if (blockIdx.x == 0)
        for (int i = tid; i < 32 * PSB; i += 256) Yg[variant * 32 * PSB + i] = Y[i];
}

int main() {
    const int reps = 2000;
    float *Wg, *Xg, *Yg; long long* cyc;
    cudaMalloc(&Wg, 32 * K * 4); cudaMalloc(&Xg, K * P * 4); cudaMalloc(&Yg, 2 * 32 * PSB * 4); cudaMalloc(&cyc, 16);
    float* h = (float*)malloc(K * P * 4);
    srand(1);
    for (int i = 0; i < 32 * K; ++i) h[i] = (rand() / (float)RAND_MAX - 0.5f) * 0.4f;
    cudaMemcpy(Wg, h, 32 * K * 4, cudaMemcpyHostToDevice);
    for (int i = 0; i < K * P; ++i) h[i] = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    cudaMemcpy(Xg, h, K * P * 4, cudaMemcpyHostToDevice);
    const size_t smem = (32 * WS + 32 + K * PSB + 32 * PSB) * 4;
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int n_sm; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    // variant 0 reads X/Y with stride 100 inside a stride-104 buffer: values differ from variant 1, so parity is checked
    // separately on the host against a double reference for variant 1 and timing only is compared for variant 0.
    for (int v = 0; v < 2; ++v) { k_bench<<<n_sm, 256, smem>>>(v, reps, Wg, Xg, Yg, cyc); cudaDeviceSynchronize(); }
    long long hc[2]; cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
    float* Wh = (float*)malloc(32 * K * 4); float* Xh = (float*)malloc(K * P * 4); float* Yh = (float*)malloc(2 * 32 * PSB * 4);
    cudaMemcpy(Wh, Wg, 32 * K * 4, cudaMemcpyDeviceToHost); cudaMemcpy(Xh, Xg, K * P * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(Yh, Yg, 2 * 32 * PSB * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int j = 0; j < 32; ++j)
        for (int p = 0; p < P; ++p) {
            double acc = 0.01 * j;
            for (int k = 0; k < K; ++k) acc += (double)Wh[j * K + k] * Xh[k * P + p];
            acc = acc > 0 ? acc : 0;
            const double e = fabs(acc - Yh[32 * PSB + j * PSB + p]);
            if (e > maxerr) maxerr = e;
            if (fabs(acc) > maxref) maxref = fabs(acc);
        }
    const double mac = 32.0 * K * P;
    printf("{\"gemm\": \"cat_layer fwd 32x120x100\", \"ffma_cycles\": %lld, \"tf32x3_mma_sync_cycles\": %lld, "
           "\"ffma_mac_per_clk\": %.1f, \"tf32x3_mac_per_clk\": %.1f, \"tf32x3_max_abs_err\": %.3e, \"max_abs_ref\": %.3f, "
           "\"cudaError\": \"%s\"}\n", hc[0], hc[1], mac / hc[0], mac / hc[1], maxerr, maxref, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
