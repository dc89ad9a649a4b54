// Microbenchmark: legacy-path mma.sync.m16n8k8 TF32 and m16n8k16 BF16 throughput per SM on sm_100a vs FFMA.
// Evidence for the "tensor pipe vs FFMA at hidden width 32" decision.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__global__ void k_mma_tf32(int iters, float* out) {
    float d[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f800000u};
    unsigned b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[q][0]), "+f"(d[q][1]), "+f"(d[q][2]), "+f"(d[q][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    if (s == 12345.f) out[0] = s;
}

__global__ void k_mma_bf16(int iters, float* out) {
    float d[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    unsigned a[4] = {0x3f803f80u + threadIdx.x, 0x3f003f00u, 0x3e803e80u, 0x3f803f80u};
    unsigned b[2] = {0x3f803f80u, 0x3f003f00u + threadIdx.x};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[q][0]), "+f"(d[q][1]), "+f"(d[q][2]), "+f"(d[q][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    if (s == 12345.f) out[0] = s;
}

__global__ void k_ffma(int iters, float* out, float b, float c) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.f) out[0] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int n_sm; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, 16);
    const int iters = 20000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int blocks = n_sm, threads = warps * 32;
        float t1 = time_ms([&] { k_mma_tf32<<<blocks, threads>>>(iters, out); });
        float t2 = time_ms([&] { k_mma_bf16<<<blocks, threads>>>(iters, out); });
        float t3 = time_ms([&] { k_ffma<<<blocks, threads>>>(iters / 4, out, 0.999f, 1e-3f); });
        double mac_tf32 = (double)blocks * warps * iters * 4 * 1024.0, mac_bf16 = (double)blocks * warps * iters * 4 * 2048.0;
        double mac_ffma = (double)blocks * threads * (iters / 4) * 16 * 8.0;
        printf("{\"warps_per_sm\": %d, \"tf32_mma_sync_tmacs\": %.2f, \"bf16_mma_sync_tmacs\": %.2f, \"ffma_tmacs\": %.2f, "
               "\"tf32_mac_per_clk_sm\": %.0f, \"bf16_mac_per_clk_sm\": %.0f, \"ffma_mac_per_clk_sm\": %.0f}\n",
               warps, mac_tf32 / t1 / 1e9, mac_bf16 / t2 / 1e9, mac_ffma / t3 / 1e9,
               mac_tf32 / (t1 * 1e-3) / n_sm / (clk * 1e3), mac_bf16 / (t2 * 1e-3) / n_sm / (clk * 1e3),
               mac_ffma / (t3 * 1e-3) / n_sm / (clk * 1e3));
    }
    return 0;
}
