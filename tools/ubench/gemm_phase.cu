// Cycle count of single GEMM building blocks of the fused tile in isolation (one CTA of 512 threads, data in shared memory).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../openobj_b200/csrc/oo_tile.h"
using namespace oo;

template <int WHICH>
__global__ void __launch_bounds__(NTHREADS, 1) kk(float* out, long long* cyc, int reps) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < SM_TOTAL; i += NTHREADS) sm[i] = 0.001f * (float)((i * 7919) % 1000) - 0.5f;
    __syncthreads();
    float* act = sm + SM_ACT;
    float* w = sm + SM_W;
    float acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (WHICH == 0) gemm_bwd_w<2, H / 8>(acc, tid, act + R_H2 * PS, act + R_H1 * PS);
        if (WHICH == 1) gemm_bwd_w<2, KP_IN / 8>(acc, tid, act + R_H1 * PS, act + R_E1 * PS);
        if (WHICH == 2) gemm_bwd_w<4, KP_HD / 8>(acc, tid, act + R_HC * PS, act + R_H4 * PS);
        if (WHICH == 3) gemm_fwd<H, WS_H, 2, true>(tid, w + W_M1, w + B_M1, act + R_H1 * PS, act + R_H2 * PS);
        if (WHICH == 4) gemm_fwd<KP_IN, WS_IN, 2, true>(tid, w + W_IN, w + B_IN, act + R_E1 * PS, act + R_H1 * PS);
        if (WHICH == 5) gemm_fwd<KP_HD, WS_HD, 4, true>(tid, w + W_CL, w + B_CL, act + R_H4 * PS, act + R_HC * PS);
        if (WHICH == 6) gemm_bwd_data<H, WS_H, H, 8, 0>(tid, w + W_M1, act + R_H2 * PS, nullptr, nullptr, act + R_H1 * PS, H, nullptr, nullptr);
        if (WHICH == 7) gemm_bwd_data<KP_IN, WS_CAT, H, WS_IN, H>(tid, w + W_CAT + H, act + R_H3 * PS, w + W_IN, act + R_H1 * PS, act + R_E1 * PS, 0, nullptr, nullptr);
        if (WHICH == 8) gemm_bwd_data<KP_HD, WS_HD, 2 * H, 8, 0>(tid, w + W_CL, act + R_HC * PS, nullptr, nullptr, act + R_H4 * PS, H, w + W_A, act + (R_MISC + M_DRAW) * PS);
        __syncthreads();
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[tid] = s + sm[tid];
    if (tid == 0) cyc[0] = (t1 - t0) / reps;
}

template <int WHICH>
void run(const char* name, int mmas) {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const size_t smem = (size_t)SM_TOTAL * 4;
    cudaFuncSetAttribute(kk<WHICH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int reps : {1, 20}) {
        kk<WHICH><<<1, NTHREADS, smem>>>(out, cyc, reps);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("{\"gemm\": \"%s\", \"reps\": %d, \"cycles\": %lld, \"mma\": %d, \"tensor_bound_cycles\": %.0f, \"err\": \"%s\"}\n", name, reps, h, mmas,
               mmas / 4.0 * 8.2, cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    run<0>("W m1 (32x32, K=points)", 312);
    run<1>("W in (32x88)", 858);
    run<2>("W heads (64x80)", 1560);
    run<3>("fwd mid1 (K=32)", 312);
    run<4>("fwd in (K=88)", 858);
    run<5>("fwd heads (64 rows, K=80)", 1560);
    run<6>("D h1 (K=32 rows)", 312);
    run<7>("D e1 (88 rows, 2 terms)", 1872);
    run<8>("D heads (80 rows, J=64)", 1560);
    return 0;
}
