// Microbenchmark: issue interval / dependent latency of legacy-path mma.sync m16n8k8 TF32 on sm_100a as a function of the
// number of independent accumulator chains per warp and of warps per scheduler (one CTA, clock64 around the loop).
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_chain(int iters, long long* cyc, float* out) {
    float d[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f800000u};
    unsigned b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < CH; ++q)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[q][0]), "+f"(d[q][1]), "+f"(d[q][2]), "+f"(d[q][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 12345.f) out[0] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CH>
void run(int warps, long long* cyc, float* out) {
    const int iters = 2000;
    k_chain<CH><<<1, warps * 32>>>(iters, cyc, out);
    cudaDeviceSynchronize();
    k_chain<CH><<<1, warps * 32>>>(iters, cyc, out);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_mma_warp = (double)h / (iters * CH);
    const double mma_per_clk_sm = (double)warps * iters * CH / (double)h;
    printf("{\"chains\": %d, \"warps\": %d, \"cycles_per_mma_per_warp\": %.2f, \"mma_per_clk_sm\": %.3f, \"mac_per_clk_sm\": %.0f}\n", CH, warps,
           per_mma_warp, mma_per_clk_sm, mma_per_clk_sm * 1024);
}

int main() {
    long long* cyc; float* out;
    cudaMalloc(&cyc, 8); cudaMalloc(&out, 16);
    for (int warps : {1, 4, 8, 16}) {
        run<1>(warps, cyc, out);
        run<2>(warps, cyc, out);
        run<4>(warps, cyc, out);
        run<8>(warps, cyc, out);
    }
    return 0;
}
