// The background model's generic GEMM descriptor (oo_bg.cu) and its two engines:
//   k_gemm      (oo_bg.cu)      mma.sync m16n8k8 TF32 x3, register-split operands           -- the round-1 path, kept for A/B runs
//   k_gemm_tc   (oo_gemm_tc.cu) tcgen05.mma kind::tf32 x3, operands pre-split in shared memory in the canonical core-matrix
//                               layout, accumulators in tensor memory                       -- the default on sm_100a
// Both compute  C(i,j) (+)= epilogue( mult * sum_c A(i,c) B(j,c) + bias[j] )  with fp32-level accuracy (x = hi + lo TF32 halves,
// lo*hi + hi*lo + hi*hi accumulated in fp32).
#pragma once
#include <cuda_runtime.h>

namespace oo {

struct GemmOp {
    const float* A; long long sai, sac;      // A(i,c) = A[i*sai + c*sac]
    const float* B; long long sbj, sbc;      // B(j,c) = B[j*sbj + c*sbc]
    float* C; long long sci, scj;
    int I, J, K;
    const float* bias;          // [J] or null
    float mult, post;           // v = (mult * acc + bias) * post
    int act;                    // 0 none, 1 relu, 2 sigmoid
    const float* mask; long long smi, smj; int mask_cols;   // C(i,j) = 0 where j < mask_cols and mask(i,j) <= 0
    int accumulate;             // C += result
    int split, chunk;           // contraction chunks (chunk is a multiple of the engine's k-tile)
    float* part;                // split > 1: raw partial sums part[z][I][J (+1)]
    float* ones_out;            // non-null: one more virtual column j == J with B(J, c) = 1, written to ones_out[i]
                                // (the bias gradient = column sums rides along with the weight gradient)
};

constexpr int TG_BI = 128, TG_BJ = 256, TG_KC = 16;      // tcgen05 engine: CTA tile 128 x (<= 256), k-chunks of 16

// launches k_gemm_tc for `g` (split / chunk / part already chosen by the caller; chunk % TG_KC == 0).  cluster_z > 1: the
// split CTAs run as thread-block clusters of that size and add their tiles through distributed shared memory before writing:
// part then holds g.split / cluster_z partials (g.split must be a multiple of cluster_z).
int run_gemm_tc(const GemmOp& g, cudaStream_t st, int cluster_z = 1);
constexpr int TG_CLUSTER = 4;
// can the tcgen05 engine stage both operands of `g` (alignment / stride rules in oo_gemm_tc.cu)?
bool gemm_tc_supported(const GemmOp& g);

}  // namespace oo
