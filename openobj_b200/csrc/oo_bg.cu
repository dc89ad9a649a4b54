// a17: the background model (objnerf/train.py:300-315,379-388,447-463; vmap.py:43-47): ONE OccupancyMap of hidden
// width 128 + UniDirsEmbed(scale 5) trained on 1200 rays x 14 samples per step beside the object ensemble.
// M = 16 800 points with K <= 215 are ordinary GEMMs, so this path is layer by layer: a generic GEMM descriptor (oo_gemm.h)
// with fused bias / activation / ReLU-mask epilogues (forward, backward-data, split-contraction backward-weight with a
// fixed-order reduction) on one of two engines -- k_gemm_tc (oo_gemm_tc.cu: tcgen05.mma kind::tf32 x3, accumulators in tensor
// memory; the default wherever the operand layouts can be staged) or k_gemm below (mma.sync 3xTF32; the 1- and 3-wide
// operands of the alpha / colour outputs, and everything with OO_BG_GEMM=mma) -- the encoder forward / backward, and the
// standalone compositing + loss kernels (K3, oo_composite.cu) in between.  Hidden width is a run-time argument (any multiple
// of 4; tests run 32, 64, 128, 256).
#include "../../include/openobj_b200.h"
#include "oo_common.cuh"
#include "oo_gemm.h"
#include "oo_layout.h"

#include <stdlib.h>
#include <string.h>


using namespace oo;

namespace oo {
// oo_bg_clip.cu: the part-feature term of the background step without [points x 512] tensors
int bg_clip_render(const float* alpha, const float* hp, int n_rays, int S, int h, int hs, float* Sx, cudaStream_t st);
int bg_clip_loss(const float* X, const float* gt_feat, const uint8_t* labels, const int* flags, const float* tail, int n_rays, int C,
                 float fs, float* d_x, float* lf, float* terms, float* loss, cudaStream_t st);
int bg_clip_bwd(const float* alpha, const float* hp, const float* dSx, int n_rays, int S, int h, int hs, float* hu, float* d_hp,
                cudaStream_t st);
int bg_clip_scatter(const float* dWb, int C, int h, int hs, float* gW, float* gb, cudaStream_t st);
// oo_composite.cu: oo_loss_bwd with an extra per-sample dL/dT input
int loss_bwd_hu(const float* alpha, const float* color, const float* z, const float* gt_depth, const float* gt_color,
                const uint8_t* labels, const float* pred_feat, const float* gt_feat, int n_obj, int n_rays, int n_samp, int n_feat,
                float cs, float os, float fs, float grad_loss, const int* flags, const float* ray_ws, float* d_alpha, float* d_color,
                float* d_pred_feat, const float* hu_extra, void* stream);
}  // namespace oo

namespace {

// ------------------------------------------------------------------------------------------------
// parameter layout of a model of hidden width h: the reference's named_parameters() order, offsets rounded to 4 floats
// ------------------------------------------------------------------------------------------------
struct BgLayout {
    int off[NT], size[NT], total;
};

BgLayout bg_layout(int h) {
    const int sz[NT] = {h * E1, h, h * h, h, h * (h + E1), h, h * h, h, h, 1, h * (h + E2), h, 3 * h, 3,
                        h * (h + E2), h, C * h, C, NDIR * 3};
    BgLayout L;
    int o = 0;
    for (int i = 0; i < NT; ++i) {
        L.off[i] = o;
        L.size[i] = sz[i];
        o += (sz[i] + 3) & ~3;
    }
    L.total = o;
    return L;
}

enum { T_IN_W, T_IN_B, T_M1_W, T_M1_B, T_CAT_W, T_CAT_B, T_M2_W, T_M2_B, T_A_W, T_A_B, T_CL_W, T_CL_B, T_OC_W, T_OC_B,
       T_CP_W, T_CP_B, T_OCL_W, T_OCL_B, T_PE };

// ------------------------------------------------------------------------------------------------
// generic GEMM at fp32 accuracy:  C(i,j) (+)= epilogue( mult * sum_c A(i,c) B(j,c) + bias[j] ),  A(i,c) = A[i*sai + c*sac] etc.
// CTA tile 128 x 64 x 16, 256 threads = 8 warps of 32 x 32 outputs on the tensor pipe (3 x TF32 mma.sync per product).
// split > 1: grid.z chunks of the contraction write raw partial sums to part[z][I][J]; k_gemm_reduce finishes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int gemm_je(const GemmOp& g) { return g.J + (g.ones_out != nullptr ? 1 : 0); }
__device__ __forceinline__ float* gemm_dst(const GemmOp& g, int i, int j) {
    return j == g.J ? g.ones_out + i : g.C + i * g.sci + j * g.scj;
}

// row strides = 8 mod 32 banks: the mma fragment loads (k = lane % 4, row = lane / 4) hit 32 different banks
constexpr int BI = 128, BJ = 64, BK = 16, LDA_S = BI + 8, LDB_S = BJ + 8;

// mma.sync m16n8k8 TF32 with three-term error compensation (x = hi + lo split in registers when the fragment is loaded;
// lo*hi + hi*lo + hi*hi, fp32 accumulate): fp32-level agreement, the same arithmetic as the fused object tile (oo_tile.h)
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ float gemm_epilogue(const GemmOp& g, int i, int j, float v) {
    v *= g.mult;
    if (g.bias && j < g.J) v += g.bias[j];
    v *= g.post;
    if (g.act == 1) v = fmaxf(v, 0.f);
    else if (g.act == 2) v = 1.f / (1.f + expf(-v));
    if (g.mask && j < g.mask_cols && !(g.mask[i * g.smi + j * g.smj] > 0.f)) v = 0.f;
    return v;
}

// AC / BC: the contraction index is the contiguous one of A / B (compile-time so that the tile loaders have no
// run-time index arithmetic).  The k-tiles arrive through a GEMM_STAGES-deep ring of asynchronous 4-byte copies
// (cp.async with zero fill outside the operand), so GEMM_STAGES - 1 tiles are in flight while one is multiplied:
// with K <= 215 per layer the kernel is a chain of memory round trips unless several of them overlap.
constexpr int GEMM_STAGES = 4;
constexpr int BG_SPLIT_MAX = 320;
constexpr int GEMM_SMEM = GEMM_STAGES * BK * (LDA_S + LDB_S) * (int)sizeof(float);
static_assert(GEMM_SMEM >= BI * (BJ + 8) * (int)sizeof(float), "the epilogue tile reuses the pipeline stages");

__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool ok) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int n = ok ? 4 : 0;                       // src-size 0: nothing is read, the destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Programmatic dependent launch between the GEMM kernels of a step (27 + 9 of its 47 launches): a kernel releases its
// dependents when its main loop is done, so the next kernel's launch latency, block scheduling and index set-up overlap this
// kernel's epilogue and tail; every kernel waits (griddepcontrol.wait = the previous grid has completed and its writes are
// visible) before its first global access.  Kernels launched without the attribute are ordered as usual.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <bool AC, bool BC>
__global__ void __launch_bounds__(256, 2) k_gemm(const GemmOp g) {
    extern __shared__ __align__(16) float gemm_sm[];
    float (*As)[BK][LDA_S] = reinterpret_cast<float (*)[BK][LDA_S]>(gemm_sm);
    float (*Bs)[BK][LDB_S] = reinterpret_cast<float (*)[BK][LDB_S]>(gemm_sm + GEMM_STAGES * BK * LDA_S);
    const int tid = threadIdx.x, i0 = blockIdx.x * BI, j0 = blockIdx.y * BJ;
    const int c_begin = blockIdx.z * g.chunk, c_end = min(g.K, c_begin + g.chunk);
    constexpr int NA = BI * BK / 256, NB_ = BJ * BK / 256;
    // this thread's elements of the A / B tiles: (ii, cc) pairs and their global offsets (without the k-tile offset)
    int a_ii[NA], a_cc[NA], b_jj[NB_], b_cc[NB_];
    int a_off[NA], b_off[NB_];                    // element offsets fit 32 bits (checked in run_gemm)
    bool a_ok[NA], b_ok[NB_];
#pragma unroll
    for (int u = 0; u < NA; ++u) {
        const int e = tid + 256 * u;
        a_ii[u] = AC ? e / BK : e % BI;
        a_cc[u] = AC ? e % BK : e / BI;
        a_ok[u] = i0 + a_ii[u] < g.I;
        a_off[u] = a_ok[u] ? (i0 + a_ii[u]) * (int)g.sai + a_cc[u] * (int)g.sac : 0;
    }
#pragma unroll
    for (int u = 0; u < NB_; ++u) {
        const int e = tid + 256 * u;
        b_jj[u] = BC ? e / BK : e % BJ;
        b_cc[u] = BC ? e % BK : e / BJ;
        const int j = j0 + b_jj[u];
        b_ok[u] = j < g.J;
        b_off[u] = b_ok[u] ? j * (int)g.sbj + b_cc[u] * (int)g.sbc : 0;
    }
    const int sac = (int)g.sac, sbc = (int)g.sbc;
    auto issue = [&](int c0, int st) {              // one commit group per k-tile (empty past the end of the chunk)
        if (c0 < c_end) {
#pragma unroll
            for (int u = 0; u < NA; ++u) {
                const bool ok = a_ok[u] && c0 + a_cc[u] < c_end;
                cp_async4(&As[st][a_cc[u]][a_ii[u]], ok ? g.A + (a_off[u] + c0 * sac) : g.A, ok);
            }
#pragma unroll
            for (int u = 0; u < NB_; ++u) {
                const bool ok = b_ok[u] && c0 + b_cc[u] < c_end;
                cp_async4(&Bs[st][b_cc[u]][b_jj[u]], ok ? g.B + (b_off[u] + c0 * sbc) : g.B, ok);
            }
        }
        cp_async_commit();
    };
    // warp tile 32 x 32 = 2 x 4 mma tiles (4 warps along i, 2 along j); fragment element maps as in oo_tile.h
    const int lane = tid & 31, fg = lane >> 2, ft = lane & 3;
    const int wi = 32 * ((tid >> 5) & 3), wj = 32 * (tid >> 7);
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
    // ones_out: sum_c A(i, c) (the bias gradient = the product with a virtual column of ones) is accumulated from the
    // staged A tiles by the CTAs of j-tile 0, thread = row; zero fill makes out-of-range rows / k contribute nothing
    const bool row_sums = g.ones_out != nullptr && blockIdx.y == 0;
    float rs = 0.f;
    pdl_wait();
#pragma unroll 1
    for (int s = 0; s < GEMM_STAGES - 1; ++s) issue(c_begin + s * BK, s);
    int st = 0;
    for (int c0 = c_begin; c0 < c_end; c0 += BK) {
        cp_async_wait<GEMM_STAGES - 2>();           // this thread's copies of tile c0 have landed ...
        __syncthreads();                            // ... and everybody's; the stage multiplied last iteration is free
        issue(c0 + (GEMM_STAGES - 1) * BK, st == 0 ? GEMM_STAGES - 1 : st - 1);
        if (row_sums && tid < BI) {
#pragma unroll
            for (int k = 0; k < BK; ++k) rs += As[st][k][tid];
        }
#pragma unroll
        for (int k0 = 0; k0 < BK; k0 += 8) {
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float* ap = &As[st][k0 + ft][wi + 16 * mt + fg];
                tf32_split(ap[0], ah[mt][0], al[mt][0]);
                tf32_split(ap[8], ah[mt][1], al[mt][1]);
                tf32_split(ap[4 * LDA_S], ah[mt][2], al[mt][2]);
                tf32_split(ap[4 * LDA_S + 8], ah[mt][3], al[mt][3]);
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float* bp = &Bs[st][k0 + ft][wj + 8 * nt + fg];
                tf32_split(bp[0], bh[nt][0], bl[nt][0]);
                tf32_split(bp[4 * LDB_S], bh[nt][1], bl[nt][1]);
            }
            // term-major order: the eight MMAs of a term are independent, the small terms go first
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[mt][nt], al[mt], bh[nt]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[mt][nt], ah[mt], bl[nt]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
        }
        st = st + 1 == GEMM_STAGES ? 0 : st + 1;
    }
    // Epilogue: accumulators -> shared tile [BI][LDT] -> global.  A thread owns ONE output column (bias, column offset and
    // mask column are thread constants) and walks down the rows four at a time, all loads of the four (ReLU mask, the old
    // value when accumulating) before the first store; a warp writes 128 contiguous bytes of a row.  The body is a short
    // rolled loop on purpose: a fully unrolled per-fragment epilogue was thousands of instructions executed once per warp,
    // and with K <= 215 that instruction stream, not the tensor pipe, set the kernel time.
    pdl_release();
    cp_async_wait<0>();
    __syncthreads();
    constexpr int LDT = BJ + 8;
    float* tile = gemm_sm;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                *reinterpret_cast<float2*>(tile + (wi + 16 * mt + fg + 8 * half) * LDT + wj + 8 * nt + 2 * ft) =
                    make_float2(acc[mt][nt][2 * half], acc[mt][nt][2 * half + 1]);
    __syncthreads();
    const int jl = tid & (BJ - 1), j = j0 + jl, je = gemm_je(g);
    const int n_rows = min(BI, g.I - i0);
    if (row_sums && tid < n_rows) {                             // column J of the partials / ones_out
        const int i = i0 + tid;
        if (g.split > 1) g.part[((size_t)blockIdx.z * g.I + i) * je + g.J] = rs;
        else {
            const float v = gemm_epilogue(g, i, g.J, rs);
            g.ones_out[i] = g.accumulate ? g.ones_out[i] + v : v;
        }
    }
    if (j >= g.J) return;
    const float* tp = tile + jl;
    int il = tid / BJ;                                          // 0..3; rows il, il + 4, ...
    if (g.split > 1) {
        float* dst = g.part + ((size_t)blockIdx.z * g.I + i0 + il) * je + j;
        for (; il < n_rows; il += 4, dst += 4 * (size_t)je) *dst = tp[il * LDT];
        return;
    }
    const bool use_mask = g.mask != nullptr && j < g.mask_cols;
    const float bias = g.bias != nullptr ? g.bias[j] : 0.f, mult = g.mult, post = g.post;
    const int act = g.act, accumulate = g.accumulate;
    const long long dstep = 4 * g.sci, mstep = 4 * g.smi;
    float* dst = g.C + (i0 + il) * g.sci + j * g.scj;
    const float* mp = use_mask ? g.mask + (i0 + il) * g.smi + j * g.smj : nullptr;
#pragma unroll 1
    for (; il < n_rows; il += 16) {
        float mk[4], old[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool ok = il + 4 * u < n_rows;
            mk[u] = (ok && use_mask) ? mp[u * mstep] : 1.f;
            old[u] = (ok && accumulate) ? dst[u * dstep] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (il + 4 * u >= n_rows) break;
            float v = (tp[(il + 4 * u) * LDT] * mult + bias) * post;
            if (act == 1) v = fmaxf(v, 0.f);
            else if (act == 2) v = 1.f / (1.f + expf(-v));
            if (!(mk[u] > 0.f)) v = 0.f;
            dst[u * dstep] = accumulate ? old[u] + v : v;
        }
        dst += 4 * dstep;
        if (use_mask) mp += 4 * mstep;
    }
}

// fixed-order reduction of the split partials: eight lanes per output element, lane q sums z = q, q + 8, ... in order and
// the eight sums are combined by a shuffle tree (deterministic; eight times shorter load chains than one thread per element)
__device__ __forceinline__ void gemm_reduce_body(const GemmOp& g);

__global__ void __launch_bounds__(256) k_gemm_reduce(const GemmOp g) { gemm_reduce_body(g); }

// The weight gradients of a step are only consumed by AdamW at its end, so their split partials stay in separate regions
// and ONE launch finishes all of them (blockIdx.y = which GEMM) instead of a latency-bound launch after every GEMM.
constexpr int REDUCE_BATCH_MAX = 9;
struct ReduceBatch {
    GemmOp op[REDUCE_BATCH_MAX];
    int n;
};
__global__ void __launch_bounds__(256) k_gemm_reduce_batch(const __grid_constant__ ReduceBatch b) { gemm_reduce_body(b.op[blockIdx.y]); }

__device__ __forceinline__ void gemm_reduce_body(const GemmOp& g) {
    const int je = gemm_je(g), n = g.I * je;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, q = threadIdx.x & 7;
    pdl_wait();
    float s = 0.f;
    if (e < n)
        for (int z = q; z < g.split; z += 8) s += g.part[(size_t)z * n + e];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (e >= n || q != 0) return;
    const int i = e / je, j = e - i * je;
    float* dst = gemm_dst(g, i, j);
    const float v = gemm_epilogue(g, i, j, s);
    *dst = g.accumulate ? *dst + v : v;
}

// split > 1 asks for a split contraction (the weight gradients: K = all points); how many chunks is chosen here so that the
// grid is ONE wave of two CTAs per SM: a CTA's time is its number of k-tiles, whatever share of its tile is real output
int run_gemm(GemmOp g, int split, float* part, cudaStream_t st, ReduceBatch* defer = nullptr) {
    static PerDevice n_sm_d;
    if (n_sm_d.cur() == 0) {
        int dev = 0, n = 0;
        OO_CUDA(cudaGetDevice(&dev));
        OO_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        n_sm_d.cur() = (size_t)n;
    }
    const int n_sm = (int)n_sm_d.cur();
    OO_REQUIRE(n_sm <= 160, "oo_bg gemm: the split-partial regions are sized for at most 160 SMs");
    // engine: tcgen05 (default) or the round-1 mma.sync kernel (OO_BG_GEMM=mma, kept for A/B runs and as the documented baseline)
    static const bool tc_on = []() { const char* e = getenv("OO_BG_GEMM"); return !(e != nullptr && strcmp(e, "mma") == 0); }();
    const bool use_tc = tc_on && gemm_tc_supported(g);     // operand layouts the tcgen05 engine cannot stage go to mma.sync
    const int tiles = use_tc ? ((g.I + TG_BI - 1) / TG_BI) * ((g.J + TG_BJ - 1) / TG_BJ) : ((g.I + BI - 1) / BI) * ((g.J + BJ - 1) / BJ);
    const int kq = use_tc ? TG_KC : BK;          // contraction granularity of a chunk
    g.split = 1;
    g.chunk = use_tc ? (g.K + kq - 1) / kq * kq : g.K;
    g.part = part;
    if (split > 1) {
        int want = (use_tc ? n_sm : 2 * n_sm) / tiles;       // one wave: 1 CTA per SM (tcgen05: 198 KB of staging) or 2 (mma.sync)
        if (want > BG_SPLIT_MAX) want = BG_SPLIT_MAX;
        if (want < 1) want = 1;
        g.chunk = ((g.K + want - 1) / want + kq - 1) / kq * kq;
        g.split = (g.K + g.chunk - 1) / g.chunk;
        if (g.split < 2) { g.split = 2; g.chunk = ((g.K + 1) / 2 + kq - 1) / kq * kq; g.split = (g.K + g.chunk - 1) / g.chunk; }
    }
    const int je = g.J + (g.ones_out != nullptr ? 1 : 0);
    OO_REQUIRE((long long)g.I * (g.sai > 0 ? g.sai : 1) + (long long)g.K * (g.sac > 0 ? g.sac : 1) < (1LL << 31) &&
                   (long long)je * (g.sbj > 0 ? g.sbj : 1) + (long long)g.K * (g.sbc > 0 ? g.sbc : 1) < (1LL << 31),
               "oo_bg gemm: operand larger than 2^31 elements");
    OO_REQUIRE((long long)g.I * (g.sci > 0 ? g.sci : 1) + (long long)je * (g.scj > 0 ? g.scj : 1) < (1LL << 31) &&
                   (g.mask == nullptr || (long long)g.I * (g.smi > 0 ? g.smi : 1) + (long long)je * (g.smj > 0 ? g.smj : 1) < (1LL << 31)),
               "oo_bg gemm: output / mask larger than 2^31 elements");
    const bool split_path = g.split > 1;          // partials in `part`: a reduction launch must follow, however few they are
    if (use_tc) {
        // split CTAs of one output tile form clusters of TG_CLUSTER and reduce through distributed shared memory first
        static const int cl_env = []() { const char* e = getenv("OO_GEMM_TC_CLUSTER"); return e ? atoi(e) : TG_CLUSTER; }();
        const int cl = g.split > 1 && cl_env > 1 ? cl_env : 1;
        g.split = (g.split + cl - 1) / cl * cl;              // padding CTAs have an empty contraction range: zero tiles
        if (int rc = run_gemm_tc(g, st, cl)) return rc;
        g.split /= cl;                                       // partials the reduction will read
    } else {
    const dim3 grid((g.I + BI - 1) / BI, (g.J + BJ - 1) / BJ, g.split);
    const bool ac = g.sac == 1, bc = g.sbc == 1;
    static PerDevice attr_set;
    if (!attr_set.cur()) {
        OO_CUDA(cudaFuncSetAttribute(k_gemm<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        OO_CUDA(cudaFuncSetAttribute(k_gemm<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        OO_CUDA(cudaFuncSetAttribute(k_gemm<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        OO_CUDA(cudaFuncSetAttribute(k_gemm<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        attr_set.cur() = 1;
    }
    if (ac && bc) OO_CUDA(launch_pdl(k_gemm<true, true>, grid, dim3(256), (size_t)GEMM_SMEM, st, g));
    else if (ac) OO_CUDA(launch_pdl(k_gemm<true, false>, grid, dim3(256), (size_t)GEMM_SMEM, st, g));
    else if (bc) OO_CUDA(launch_pdl(k_gemm<false, true>, grid, dim3(256), (size_t)GEMM_SMEM, st, g));
    else OO_CUDA(launch_pdl(k_gemm<false, false>, grid, dim3(256), (size_t)GEMM_SMEM, st, g));
    OO_LAUNCH_CHECK();
    }
    if (split_path) {
        if (defer != nullptr) {
            OO_REQUIRE(defer->n < REDUCE_BATCH_MAX, "oo_bg gemm: too many deferred reductions");
            defer->op[defer->n++] = g;
        } else {
            OO_CUDA(launch_pdl(k_gemm_reduce, dim3((g.I * je * 8 + 255) / 256), dim3(256), (size_t)0, st, g));
            OO_LAUNCH_CHECK();
        }
    }
    return 0;
}

int run_reduce_batch(const ReduceBatch& b, cudaStream_t st) {
    if (b.n == 0) return 0;
    int gx = 1;
    for (int i = 0; i < b.n; ++i) {
        const int blocks = (b.op[i].I * (b.op[i].J + (b.op[i].ones_out != nullptr ? 1 : 0)) * 8 + 255) / 256;
        gx = blocks > gx ? blocks : gx;
    }
    OO_CUDA(launch_pdl(k_gemm_reduce_batch, dim3(gx, b.n), dim3(256), (size_t)0, st, b));
    OO_LAUNCH_CHECK();
    return 0;
}

// floats of split partials run_gemm may write for an I x J weight gradient (+ the bias column): split * tiles <= 2 CTAs per SM
long long part_floats(int I, int J) {
    const long long tiles = (long long)((I + BI - 1) / BI) * ((J + BJ - 1) / BJ);
    const long long tiles_tc = (long long)((I + TG_BI - 1) / TG_BI) * ((J + TG_BJ - 1) / TG_BJ);
    long long split = 2 * 160 / tiles;                 // >= what run_gemm picks for any SM count up to 160 (mma.sync engine)
    if (160 / tiles_tc > split) split = 160 / tiles_tc;    // tcgen05 engine: one CTA per SM
    if (split > BG_SPLIT_MAX) split = BG_SPLIT_MAX;
    if (split < 2) split = 2;
    return (split + 1) * (long long)I * (J + 1);
}

// ------------------------------------------------------------------------------------------------
// encoder (embedding.py:46-55): e = [t, sin(pi 2^k B t)], written straight into the three places that consume it:
// in_layer input X1 [M][ld1] (cols 0..86), cat_layer input XC [M][ldc] (cols h..h+86), head input XH [M][ldh]
// (cols h..h+41)
// ------------------------------------------------------------------------------------------------
struct EmbedBufs {
    float *x1, *xc, *xh;     // forward destinations, or the gradient sources in the backward kernel
    int ld1, ldc, ldh, h;
};

__device__ __forceinline__ void emb_store(const EmbedBufs& b, size_t p, int idx, float v) {
    if (idx < E1) {
        b.x1[p * b.ld1 + idx] = v;
        b.xc[p * b.ldc + b.h + idx] = v;
    } else {
        b.xh[p * b.ldh + b.h + idx - E1] = v;
    }
}

__global__ void k_embed_fwd(const float* __restrict__ pcs, const float* __restrict__ Bm, float scale, int n_pts,
                            const EmbedBufs b, float* __restrict__ emb_out) {
    // thread = (point, slot): slots 0..20 = directions, 21..23 = the three scaled coordinates
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t p = e / 24;
    const int d = (int)(e - p * 24);
    if (p >= (size_t)n_pts) return;
    const float t0 = pcs[3 * p] / scale, t1 = pcs[3 * p + 1] / scale, t2 = pcs[3 * p + 2] / scale;
    if (d >= NDIR) {
        const int ch = d - NDIR;
        const float t = ch == 0 ? t0 : ch == 1 ? t1 : t2;
        if (b.x1) emb_store(b, p, ch, t);
        if (emb_out) emb_out[p * EMB + ch] = t;
        return;
    }
    const float proj = Bm[3 * d] * t0 + Bm[3 * d + 1] * t1 + Bm[3 * d + 2] * t2;
    const float arg = proj * PI_F;                  // band k: fl(2^k proj pi_f) == 2^k fl(proj pi_f) exactly
    float band = 1.f;
#pragma unroll
    for (int k = 0; k < NBAND; ++k) {
        const float s = sinf(arg * band);
        if (b.x1) emb_store(b, p, 3 + NDIR * k + d, s);
        if (emb_out) emb_out[p * EMB + 3 + NDIR * k + d] = s;
        band *= 2.f;
    }
}

// OccupancyMap.forward on a caller-supplied embedding [M][129] (the module-level call form fc_occ_map(pe(x)), train.py:449-450):
// the embedding goes to the three places that consume it; its gradient is gathered back from them
__global__ void k_embed_scatter(const float* __restrict__ emb, int n_pts, const EmbedBufs b) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n_pts * EMB) return;
    const size_t p = e / EMB;
    emb_store(b, p, (int)(e - p * EMB), emb[e]);
}
__global__ void k_embed_gather_grad(const EmbedBufs g, int n_pts, float* __restrict__ d_emb) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n_pts * EMB) return;
    const size_t p = e / EMB;
    const int idx = (int)(e - p * EMB);
    d_emb[e] = idx < E1 ? g.x1[p * g.ld1 + idx] + g.xc[p * g.ldc + g.h + idx] : g.xh[p * g.ldh + g.h + idx - E1];
}

// d B[d][ch] = sum_p t[p][ch] * sum_k de[3+21k+d][p] * pi 2^k cos(pi 2^k proj): per-block partials [nblk][63], then a
// fixed-order reduction
constexpr int EB_PTS = 32;        // points per block
__global__ void __launch_bounds__(NDIR * 4) k_embed_bwd(const float* __restrict__ pcs, const float* __restrict__ Bm, float scale,
                                                         int n_pts, const EmbedBufs g, float* __restrict__ partial) {
    // thread = (point lane q of 4, direction d) with d fastest: the 21 directions of a band are 84 contiguous bytes of a row.
    // cos(2^k theta) for the six bands from ONE sincosf by angle doubling (as the fused tile, oo_tile.h phase 30: the reference's
    // argument for band k is exactly 2^k times that of band 0); each thread walks EB_PTS / 4 points
    __shared__ float sh[4][NDIR * 3];
    const int d = threadIdx.x % NDIR, q = threadIdx.x / NDIR;
    const float b0 = Bm[3 * d], b1 = Bm[3 * d + 1], b2 = Bm[3 * d + 2];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    const int p_end = min(n_pts, (int)(blockIdx.x + 1) * EB_PTS);
    for (int p = blockIdx.x * EB_PTS + q; p < p_end; p += 4) {
        const float t0 = pcs[3 * (size_t)p] / scale, t1 = pcs[3 * (size_t)p + 1] / scale, t2 = pcs[3 * (size_t)p + 2] / scale;
        float sn, cs;
        sincosf((b0 * t0 + b1 * t1 + b2 * t2) * PI_F, &sn, &cs);
        float de[NBAND];
#pragma unroll
        for (int k = 0; k < NBAND; ++k) {                       // all loads before the dependent chain
            const int idx = 3 + NDIR * k + d;
            de[k] = idx < E1 ? g.x1[(size_t)p * g.ld1 + idx] + g.xc[(size_t)p * g.ldc + g.h + idx]
                             : g.xh[(size_t)p * g.ldh + g.h + idx - E1];
        }
        float band = PI_F, dp = 0.f;
#pragma unroll
        for (int k = 0; k < NBAND; ++k) {
            dp += de[k] * (cs * band);
            const float s2_ = 2.f * sn * cs, c2_ = (cs - sn) * (cs + sn);
            sn = s2_; cs = c2_;
            band *= 2.f;
        }
        s0 += dp * t0; s1 += dp * t1; s2 += dp * t2;
    }
    sh[q][3 * d] = s0; sh[q][3 * d + 1] = s1; sh[q][3 * d + 2] = s2;
    __syncthreads();
    if (threadIdx.x < NDIR * 3)
        partial[(size_t)blockIdx.x * (NDIR * 3) + threadIdx.x] =
            (sh[0][threadIdx.x] + sh[1][threadIdx.x]) + (sh[2][threadIdx.x] + sh[3][threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_embed_bwd_reduce(const float* __restrict__ partial, int nblk, float* __restrict__ dB) {
    // one warp per output element: lane l sums blocks l, l + 32, ... in order, then a shuffle tree (fixed order)
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (e >= NDIR * 3) return;
    float s = 0.f;
    for (int b = lane; b < nblk; b += 32) s += partial[(size_t)b * (NDIR * 3) + e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) dB[e] = s;
}

// ------------------------------------------------------------------------------------------------
// small elementwise kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_sigmoid_bwd(const float* __restrict__ d, const float* __restrict__ y, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = d[i] * y[i] * (1.f - y[i]);
}

__global__ void k_gt_prepare(const uint8_t* __restrict__ rgb8, const int32_t* __restrict__ feat_row,
                             const float* __restrict__ feat_table, int n_rays, float* __restrict__ rgb,
                             float* __restrict__ feat) {
    // gt colour u8 -> float / 255 (train.py:373); gt part-feature rows gathered from the resident table (train.py:378)
    const int r = blockIdx.x;
    if (threadIdx.x < 3) rgb[3 * r + threadIdx.x] = (float)rgb8[3 * r + threadIdx.x] / 255.f;
    if (feat_row) {
        const float4* src = reinterpret_cast<const float4*>(feat_table + (size_t)feat_row[r] * C);
        float4* dst = reinterpret_cast<float4*>(feat + (size_t)r * C);
        for (int q = threadIdx.x; q < C / 4; q += blockDim.x) dst[q] = src[q];
    }
}

__global__ void k_fill(float* p, float v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// workspace map (floats)
// ------------------------------------------------------------------------------------------------
struct BgWs {
    // forward activations
    float *x1, *xc, *h1, *h3, *xh, *hc, *hp, *alpha, *color, *clip;
    // gradients
    float *d_alpha, *d_color, *d_colpre, *d_clip, *d_hc, *d_hp, *d_xh, *d_h3, *d_xc, *d_h1, *d_x1;
    float *gt_rgb, *gt_feat, *loss_ws, *ones, *adam_scal, *part[9], *emb_part, *grads;   // part[i]: split partials of weight gradient i
    float *wp_in, *wp_cat, *wp_cl, *wp_cp;   // in_layer / cat_layer / color_linear / clip_linear weights with rows padded to ld1 / ldc / ldh
    // factored clip head (oo_bg_clip.cu): [W_ocl | b_ocl | 0] rows of hs = h + 4, per-ray tensors, split partials
    float *wb_ocl, *Sx, *X, *d_x, *dSx, *hu, *lf, *dWb, *part_ds, *part_wb;
    int hs;
    int ld1, ldc, ldh;
    long long total;
};

constexpr int BG_SPLIT = 64;            // "split the contraction" request of the weight-gradient GEMMs (run_gemm picks the count)

BgWs bg_ws_map(float* base, int h, int n_pts, int n_rays) {
    BgWs w;
    w.ld1 = 88; w.ldc = (h + E1 + 3) & ~3; w.ldh = (h + E2 + 3) & ~3;
    long long o = 0;
    auto take = [&](long long n) { float* p = base ? base + o : nullptr; o += (n + 3) & ~3LL; return p; };
    const long long M = n_pts;
    w.x1 = take(M * w.ld1); w.xc = take(M * w.ldc); w.h1 = take(M * h); w.h3 = take(M * h); w.xh = take(M * w.ldh);
    w.hc = take(M * h); w.hp = take(M * h); w.alpha = take(M); w.color = take(M * 3); w.clip = take(M * C);
    w.d_alpha = take(M); w.d_color = take(M * 3); w.d_colpre = take(M * 3); w.d_clip = take(M * C);
    w.d_hc = take(M * h); w.d_hp = take(M * h); w.d_xh = take(M * w.ldh); w.d_h3 = take(M * h); w.d_xc = take(M * w.ldc);
    w.d_h1 = take(M * h); w.d_x1 = take(M * w.ld1);
    w.gt_rgb = take(3LL * n_rays); w.gt_feat = take((long long)C * n_rays);
    w.loss_ws = take((long long)n_rays * oo_loss_ws_per_ray() + 8);
    w.ones = take(4);
    w.adam_scal = take(12);
    {   // one region per weight-gradient GEMM, in the order oo_bg_train_step runs them (they are reduced together at the end)
        const int wi[9] = {C, h, 3, h, 1, h, h, h, h}, wj[9] = {h, h + E2, h, h + E2, h, h, h + E1, h, E1};
        for (int i = 0; i < 9; ++i) w.part[i] = take(part_floats(wi[i], wj[i]));
    }
    w.emb_part = take(((M + EB_PTS - 1) / EB_PTS) * (NDIR * 3));
    w.grads = take(bg_layout(h).total);
    w.wp_in = take((long long)h * w.ld1); w.wp_cat = take((long long)h * w.ldc);
    w.wp_cl = take((long long)h * w.ldh); w.wp_cp = take((long long)h * w.ldh);
    w.hs = h + 4;
    w.wb_ocl = take((long long)C * w.hs); w.Sx = take((long long)n_rays * w.hs); w.X = take((long long)n_rays * C);
    w.d_x = take((long long)n_rays * C); w.dSx = take((long long)n_rays * w.hs); w.hu = take(M); w.lf = take(n_rays);
    w.dWb = take((long long)C * w.hs); w.part_ds = take(part_floats(n_rays, w.hs)); w.part_wb = take(part_floats(C, w.hs));
    w.total = o;
    return w;
}

// The four weight matrices whose rows are not multiples of 4 floats (in 87, cat h + 87, heads h + 42) get a zero-padded,
// 16-byte aligned copy per step: the tcgen05 engine stages aligned rows with 16-byte asynchronous copies (forward: weights are
// the contraction-contiguous B operand) and 128-bit loads (backward-data: row-contiguous B operand).
__global__ void k_pack_weights(const float* __restrict__ th, int h, int o_in, int o_cat, int o_cl, int o_cp, float* __restrict__ p_in,
                               float* __restrict__ p_cat, float* __restrict__ p_cl, float* __restrict__ p_cp, int ld1, int ldc, int ldh,
                               int o_ocl_w, int o_ocl_b, float* __restrict__ p_wb, int hs) {
    const int which = blockIdx.y;
    if (which == 4) {                     // [W_ocl | b_ocl | 0 0 0]: the factored clip head's one operand (oo_bg_clip.cu)
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < C * hs; e += gridDim.x * blockDim.x) {
            const int r = e / hs, c = e - r * hs;
            p_wb[e] = c < h ? th[o_ocl_w + r * h + c] : c == h ? th[o_ocl_b + r] : 0.f;
        }
        return;
    }
    const float* src = th + (which == 0 ? o_in : which == 1 ? o_cat : which == 2 ? o_cl : o_cp);
    float* dst = which == 0 ? p_in : which == 1 ? p_cat : which == 2 ? p_cl : p_cp;
    const int cols = which == 0 ? E1 : which == 1 ? h + E1 : h + E2, ld = which == 0 ? ld1 : which == 1 ? ldc : ldh;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < h * ld; e += gridDim.x * blockDim.x) {
        const int r = e / ld, c = e - r * ld;
        dst[e] = c < cols ? src[r * cols + c] : 0.f;
    }
}

GemmOp op(const float* A, long long sai, long long sac, const float* B, long long sbj, long long sbc, float* Cp, long long sci,
          long long scj, int I, int J, int K) {
    GemmOp g = {};
    g.A = A; g.sai = sai; g.sac = sac; g.B = B; g.sbj = sbj; g.sbc = sbc; g.C = Cp; g.sci = sci; g.scj = scj;
    g.I = I; g.J = J; g.K = K; g.mult = 1.f; g.post = 1.f;
    return g;
}

#define OO_TRY(expr)              \
    do {                          \
        if (int rc__ = (expr)) return rc__; \
    } while (0)

// forward of the whole model into the workspace (model.py:61-103); returns with ws.alpha (x10 applied), ws.color
// (after sigmoid) and, if want_clip, ws.clip filled
int bg_forward(const float* th, const BgLayout& L, int h, const float* pcs, int M, float scale, const BgWs& w, bool want_clip,
               float* emb_out, cudaStream_t st, const float* emb_in = nullptr) {
    const EmbedBufs eb = {w.x1, w.xc, w.xh, w.ld1, w.ldc, w.ldh, h};
    if (emb_in != nullptr) k_embed_scatter<<<(unsigned)(((size_t)M * EMB + 255) / 256), 256, 0, st>>>(emb_in, M, eb);
    else k_embed_fwd<<<(unsigned)(((size_t)M * 24 + 255) / 256), 256, 0, st>>>(pcs, th + L.off[T_PE], scale, M, eb, emb_out);
    OO_LAUNCH_CHECK();
    k_pack_weights<<<dim3(16, 5), 256, 0, st>>>(th, h, L.off[T_IN_W], L.off[T_CAT_W], L.off[T_CL_W], L.off[T_CP_W], w.wp_in, w.wp_cat,
                                               w.wp_cl, w.wp_cp, w.ld1, w.ldc, w.ldh, L.off[T_OCL_W], L.off[T_OCL_B], w.wb_ocl, w.hs);
    OO_LAUNCH_CHECK();
    GemmOp g;
    // fc1 = relu(in_layer(e1))
    g = op(w.x1, w.ld1, 1, w.wp_in, w.ld1, 1, w.h1, h, 1, M, h, E1); g.bias = th + L.off[T_IN_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // fc2 = relu(mid1(fc1)) -> cols 0..h-1 of the cat input
    g = op(w.h1, h, 1, th + L.off[T_M1_W], h, 1, w.xc, w.ldc, 1, M, h, h); g.bias = th + L.off[T_M1_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // fc3 = relu(cat_layer([fc2, e1]))
    g = op(w.xc, w.ldc, 1, w.wp_cat, w.ldc, 1, w.h3, h, 1, M, h, h + E1); g.bias = th + L.off[T_CAT_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // fc4 = relu(mid2(fc3)) -> cols 0..h-1 of the head input
    g = op(w.h3, h, 1, th + L.off[T_M2_W], h, 1, w.xh, w.ldh, 1, M, h, h); g.bias = th + L.off[T_M2_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // alpha = 10 * (out_alpha(fc4))   (model.py:86-88)
    g = op(w.xh, w.ldh, 1, th + L.off[T_A_W], h, 1, w.alpha, 1, 1, M, 1, h); g.bias = th + L.off[T_A_B]; g.post = 10.f;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // color = sigmoid(out_color(relu(color_linear([fc4, e2]))))
    g = op(w.xh, w.ldh, 1, w.wp_cl, w.ldh, 1, w.hc, h, 1, M, h, h + E2); g.bias = th + L.off[T_CL_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    g = op(w.hc, h, 1, th + L.off[T_OC_W], h, 1, w.color, 3, 1, M, 3, h); g.bias = th + L.off[T_OC_B]; g.act = 2;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // clip = out_clip(relu(clip_linear([fc4, e2])))
    g = op(w.xh, w.ldh, 1, w.wp_cp, w.ldh, 1, w.hp, h, 1, M, h, h + E2); g.bias = th + L.off[T_CP_B]; g.act = 1;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    if (want_clip) {
        g = op(w.hp, h, 1, th + L.off[T_OCL_W], h, 1, w.clip, C, 1, M, C, h); g.bias = th + L.off[T_OCL_B];
        OO_TRY(run_gemm(g, 1, nullptr, st));
    }
    return 0;
}


// backward of the whole model given dL/d(alpha, colour, clip) in w.d_alpha / w.d_color / w.d_clip and the forward activations
// of bg_forward in `w`: all 19 gradients into G (flat parameter layout)
int bg_backward(const float* theta, const BgLayout& L, int h, const float* pcs, int M, float scale, const BgWs& w, bool part,
                float* G, cudaStream_t st, float* d_emb_out = nullptr, int factored_rays = 0) {
    const float* th = theta;
    GemmOp g;
    ReduceBatch pending;                         // split partials of the weight gradients, reduced together before AdamW
    pending.n = 0;
    OO_CUDA(cudaMemsetAsync(G, 0, sizeof(float) * L.total, st));
    // ---- clip head
    if (part && factored_rays > 0) {
        // factored clip head (oo_bg_clip.cu): d hp is already in w.d_hp; [dW | db | 0] = d_x^T [S | opac | 0] over the RAYS
        g = op(w.d_x, 1, C, w.Sx, 1, w.hs, w.dWb, w.hs, 1, C, w.hs, factored_rays);
        OO_TRY(run_gemm(g, BG_SPLIT, w.part_wb, st, &pending));
        g = op(w.d_hp, 1, h, w.xh, 1, w.ldh, G + L.off[T_CP_W], h + E2, 1, h, h + E2, M);
        g.ones_out = G + L.off[T_CP_B];
        OO_TRY(run_gemm(g, BG_SPLIT, w.part[1], st, &pending));
    } else if (part) {
        // d out_clip.weight [C][h] = d_clip^T hp ; bias = column sums ; d hp = (d_clip W_ocl) * [hp > 0]
        g = op(w.d_clip, 1, C, w.hp, 1, h, G + L.off[T_OCL_W], h, 1, C, h, M);
        g.ones_out = G + L.off[T_OCL_B];
        OO_TRY(run_gemm(g, BG_SPLIT, w.part[0], st, &pending));
        g = op(w.d_clip, C, 1, th + L.off[T_OCL_W], 1, h, w.d_hp, h, 1, M, h, C);
        g.mask = w.hp; g.smi = h; g.smj = 1; g.mask_cols = h;
        OO_TRY(run_gemm(g, 1, nullptr, st));
        g = op(w.d_hp, 1, h, w.xh, 1, w.ldh, G + L.off[T_CP_W], h + E2, 1, h, h + E2, M);
        g.ones_out = G + L.off[T_CP_B];
        OO_TRY(run_gemm(g, BG_SPLIT, w.part[1], st, &pending));
    }
    // ---- colour head: sigmoid backward, out_color, color_linear
    k_sigmoid_bwd<<<(3 * M + 255) / 256, 256, 0, st>>>(w.d_color, w.color, w.d_colpre, 3 * M);
    OO_LAUNCH_CHECK();
    g = op(w.d_colpre, 1, 3, w.hc, 1, h, G + L.off[T_OC_W], h, 1, 3, h, M);
    g.ones_out = G + L.off[T_OC_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part[2], st, &pending));
    g = op(w.d_colpre, 3, 1, th + L.off[T_OC_W], 1, h, w.d_hc, h, 1, M, h, 3);
    g.mask = w.hc; g.smi = h; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    g = op(w.d_hc, 1, h, w.xh, 1, w.ldh, G + L.off[T_CL_W], h + E2, 1, h, h + E2, M);
    g.ones_out = G + L.off[T_CL_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part[3], st, &pending));
    // ---- alpha head: alpha = 10 * (W_a fc4 + b_a)
    g = op(w.d_alpha, 1, 1, w.xh, 1, w.ldh, G + L.off[T_A_W], h, 1, 1, h, M); g.mult = 10.f;
    g.ones_out = G + L.off[T_A_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part[4], st, &pending));
    // ---- d [fc4, e2] = d_hc W_cl + d_hp W_cp + 10 d_alpha W_a (cols < h).  The ReLU mask of fc4 is a 0/1 factor, so it is
    // applied to every term as it is added: (a + b + c) m == a m + b m + c m exactly, in the same order
    g = op(w.d_hc, h, 1, w.wp_cl, 1, w.ldh, w.d_xh, w.ldh, 1, M, h + E2, h);
    g.mask = w.xh; g.smi = w.ldh; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    if (part) {
        g = op(w.d_hp, h, 1, w.wp_cp, 1, w.ldh, w.d_xh, w.ldh, 1, M, h + E2, h); g.accumulate = 1;
        g.mask = w.xh; g.smi = w.ldh; g.smj = 1; g.mask_cols = h;
        OO_TRY(run_gemm(g, 1, nullptr, st));
    }
    g = op(w.d_alpha, 1, 1, th + L.off[T_A_W], 1, h, w.d_xh, w.ldh, 1, M, h, 1); g.mult = 10.f; g.accumulate = 1;
    g.mask = w.xh; g.smi = w.ldh; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // ---- mid2: d W = d_fc4^T fc3, d fc3 = (d_fc4 W_m2) * [fc3 > 0]
    g = op(w.d_xh, 1, w.ldh, w.h3, 1, h, G + L.off[T_M2_W], h, 1, h, h, M);
    g.ones_out = G + L.off[T_M2_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part[5], st, &pending));
    g = op(w.d_xh, w.ldh, 1, th + L.off[T_M2_W], 1, h, w.d_h3, h, 1, M, h, h);
    g.mask = w.h3; g.smi = h; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // ---- cat_layer: d W = d_fc3^T [fc2, e1], d [fc2, e1] = d_fc3 W_cat with the ReLU mask on the fc2 columns
    g = op(w.d_h3, 1, h, w.xc, 1, w.ldc, G + L.off[T_CAT_W], h + E1, 1, h, h + E1, M);
    g.ones_out = G + L.off[T_CAT_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part[6], st, &pending));
    g = op(w.d_h3, h, 1, w.wp_cat, 1, w.ldc, w.d_xc, w.ldc, 1, M, h + E1, h);
    g.mask = w.xc; g.smi = w.ldc; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // ---- mid1
    g = op(w.d_xc, 1, w.ldc, w.h1, 1, h, G + L.off[T_M1_W], h, 1, h, h, M);
    g.ones_out = G + L.off[T_M1_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part[7], st, &pending));
    g = op(w.d_xc, w.ldc, 1, th + L.off[T_M1_W], 1, h, w.d_h1, h, 1, M, h, h);
    g.mask = w.h1; g.smi = h; g.smj = 1; g.mask_cols = h;
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // ---- in_layer
    g = op(w.d_h1, 1, h, w.x1, 1, w.ld1, G + L.off[T_IN_W], E1, 1, h, E1, M);
    g.ones_out = G + L.off[T_IN_B];
    OO_TRY(run_gemm(g, BG_SPLIT, w.part[8], st, &pending));
    g = op(w.d_h1, h, 1, w.wp_in, 1, w.ld1, w.d_x1, w.ld1, 1, M, E1, h);
    OO_TRY(run_gemm(g, 1, nullptr, st));
    // ---- encoder: B_layer.weight is trainable (embedding.py:43); on a caller-supplied embedding its gradient
    // goes back to the caller instead
    if (pcs == nullptr) {
        const EmbedBufs gb = {w.d_x1, w.d_xc, w.d_xh, w.ld1, w.ldc, w.ldh, h};
        if (d_emb_out != nullptr) {
            k_embed_gather_grad<<<(unsigned)(((size_t)M * EMB + 255) / 256), 256, 0, st>>>(gb, M, d_emb_out);
            OO_LAUNCH_CHECK();
        }
    } else {
        const EmbedBufs gb = {w.d_x1, w.d_xc, w.d_xh, w.ld1, w.ldc, w.ldh, h};
        const int nblk = (M + EB_PTS - 1) / EB_PTS;
        k_embed_bwd<<<nblk, NDIR * 4, 0, st>>>(pcs, th + L.off[T_PE], scale, M, gb, w.emb_part);
        OO_LAUNCH_CHECK();
        k_embed_bwd_reduce<<<(NDIR * 3 + 7) / 8, 256, 0, st>>>(w.emb_part, nblk, G + L.off[T_PE]);
        OO_LAUNCH_CHECK();
    }
    OO_TRY(run_reduce_batch(pending, st));
    if (part && factored_rays > 0)
        OO_TRY(bg_clip_scatter(w.dWb, C, h, w.hs, G + L.off[T_OCL_W], G + L.off[T_OCL_B], st));
    return 0;
}

// ---- torch.optim.AdamW bookkeeping of the background model on the device (train.py:473): which parameter groups
// autograd reaches in this step follows from the zero-mask flags of its own step_batch_loss call (render_rays.py:89-94 with
// N = 1: a term whose mask is empty is a constant zero, so the tensors only it reaches have grad None and are skipped --
// no decay, no step count).  Groups: 0 trunk + alpha + PE, 1 colour head, 2 clip head (as the ensemble, oo_layout.h).
__global__ void k_bg_adam_sched(const int* __restrict__ flags, int part_on, double lr, double b1, double b2, int* __restrict__ adam_t,
                                float* __restrict__ scal) {
    const int g = threadIdx.x;
    if (g >= 3) return;
    const int f = flags[0];
    const bool obj_terms = !(f & OO_FLAG_NO_OBJ), op_term = !(f & OO_FLAG_NO_SEM);
    const bool active = g == 0 ? (obj_terms || op_term) : g == 1 ? obj_terms : (obj_terms && part_on != 0);
    float* s = scal + 4 * g;
    if (active) {
        const int t = adam_t[g] + 1;
        adam_t[g] = t;
        const double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
        s[0] = 1.f; s[1] = (float)(lr / bc1); s[2] = (float)sqrt(bc2); s[3] = (float)t;
    } else {
        s[0] = s[1] = s[2] = s[3] = 0.f;
    }
}

__global__ void k_bg_adamw(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int n,
                           int off_color, int off_clip, int off_pe, const float* __restrict__ scal, float decay, float b1, float b2,
                           float eps) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int grp = i < off_color ? 0 : i < off_clip ? 1 : i < off_pe ? 2 : 0;
        if (scal[4 * grp] == 0.f) continue;
        const float step = scal[4 * grp + 1], bc2s = scal[4 * grp + 2];
        float P = p[i] * decay;
        const float G = g[i];
        const float Mn = m[i] + (G - m[i]) * (1.f - b1);
        const float Vn = v[i] * b2 + ((1.f - b2) * G) * G;
        P = P - step * (Mn / (sqrtf(Vn) / bc2s + eps));
        p[i] = P; m[i] = Mn; v[i] = Vn;
    }
}
}  // namespace

extern "C" int oo_bg_param_count(int hidden) { return hidden > 0 ? bg_layout(hidden).total : -1; }
extern "C" int oo_bg_param_offset(int hidden, int i) { return (hidden > 0 && i >= 0 && i < NT) ? bg_layout(hidden).off[i] : -1; }
extern "C" int oo_bg_param_size(int hidden, int i) { return (hidden > 0 && i >= 0 && i < NT) ? bg_layout(hidden).size[i] : -1; }
extern "C" int64_t oo_bg_ws_floats(int hidden, int n_pts, int n_rays) {
    if (hidden <= 0 || n_pts <= 0 || n_rays <= 0) return -1;
    return bg_ws_map(nullptr, hidden, n_pts, n_rays).total;
}

extern "C" int oo_bg_forward(const float* theta, int hidden, const float* pcs, const float* emb_in, int n_pts, float scale,
                             float* alpha, float* color, float* clip, float* emb_out, float* ws, void* stream) {
    OO_REQUIRE(theta && ws && alpha && color, "oo_bg_forward: null argument");
    OO_REQUIRE((pcs != nullptr) != (emb_in != nullptr), "oo_bg_forward: give exactly one of pcs / emb_in");
    OO_REQUIRE(hidden > 0 && hidden % 4 == 0 && n_pts > 0, "oo_bg_forward: hidden must be a positive multiple of 4");
    cudaStream_t st = (cudaStream_t)stream;
    const BgLayout L = bg_layout(hidden);
    const BgWs w = bg_ws_map(ws, hidden, n_pts, 1);
    OO_TRY(bg_forward(theta, L, hidden, pcs, n_pts, scale, w, clip != nullptr, emb_out, st, emb_in));
    OO_CUDA(cudaMemcpyAsync(alpha, w.alpha, sizeof(float) * n_pts, cudaMemcpyDeviceToDevice, st));
    OO_CUDA(cudaMemcpyAsync(color, w.color, sizeof(float) * 3 * n_pts, cudaMemcpyDeviceToDevice, st));
    if (clip) OO_CUDA(cudaMemcpyAsync(clip, w.clip, sizeof(float) * C * (size_t)n_pts, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int oo_bg_forward_bwd(const float* theta, int hidden, const float* pcs, const float* emb_in, int n_pts, float scale,
                                 const float* d_alpha, const float* d_color, const float* d_clip, float* grads_out,
                                 float* d_emb_out, float* ws, void* stream) {
    OO_REQUIRE(theta && d_alpha && d_color && grads_out && ws, "oo_bg_forward_bwd: null argument");
    OO_REQUIRE((pcs != nullptr) != (emb_in != nullptr), "oo_bg_forward_bwd: give exactly one of pcs / emb_in");
    OO_REQUIRE(hidden > 0 && hidden % 4 == 0 && n_pts > 0, "oo_bg_forward_bwd: hidden must be a positive multiple of 4");
    cudaStream_t st = (cudaStream_t)stream;
    const BgLayout L = bg_layout(hidden);
    BgWs w = bg_ws_map(ws, hidden, n_pts, 1);
    k_fill<<<1, 32, 0, st>>>(w.ones, 1.f, 4);
    OO_LAUNCH_CHECK();
    OO_TRY(bg_forward(theta, L, hidden, pcs, n_pts, scale, w, false, nullptr, st, emb_in));
    // the caller's upstream gradients take the place of the loss kernel's outputs (read-only in bg_backward)
    w.d_alpha = const_cast<float*>(d_alpha);
    w.d_color = const_cast<float*>(d_color);
    w.d_clip = const_cast<float*>(d_clip);
    return bg_backward(theta, L, hidden, pcs, n_pts, scale, w, d_clip != nullptr, grads_out, st, d_emb_out);
}

extern "C" int oo_bg_train_step(float* theta, float* adam_m, float* adam_v, int hidden, const float* pcs, const float* z,
                                const float* gt_depth, const uint8_t* gt_rgb, const uint8_t* labels, const int32_t* feat_row,
                                const float* feat_table, int n_rays, int n_samp, float scale, int* adam_t, float lr,
                                float weight_decay, float beta1, float beta2, float eps, float color_scaling,
                                float opacity_scaling, float feat_scaling, float* ws, float* terms_out, float* loss_out,
                                int* flags_out, float* grads_out, void* stream) {
    OO_REQUIRE(theta && pcs && z && gt_depth && gt_rgb && labels && ws && terms_out && loss_out && flags_out,
               "oo_bg_train_step: null argument");
    OO_REQUIRE(hidden > 0 && hidden % 4 == 0 && n_rays > 0 && n_samp > 0 && n_samp <= 32, "oo_bg_train_step: bad shape");
    OO_REQUIRE(grads_out || (adam_m && adam_v && adam_t), "oo_bg_train_step: optimiser state missing");
    OO_REQUIRE((feat_row == nullptr) == (feat_table == nullptr), "oo_bg_train_step: feat_row and feat_table go together");
    cudaStream_t st = (cudaStream_t)stream;
    const int h = hidden, M = n_rays * n_samp;
    const bool part = feat_row != nullptr;
    const BgLayout L = bg_layout(h);
    const BgWs w = bg_ws_map(ws, h, M, n_rays);
    float* G = grads_out ? grads_out : w.grads;
    k_fill<<<1, 32, 0, st>>>(w.ones, 1.f, 4);
    OO_LAUNCH_CHECK();
    k_gt_prepare<<<n_rays, 128, 0, st>>>(gt_rgb, feat_row, feat_table, n_rays, w.gt_rgb, w.gt_feat);
    OO_LAUNCH_CHECK();
    // part features: the clip head is applied per RAY to S_r = sum_i T_i hp_i (oo_bg_clip.cu) unless OO_BG_CLIP=dense asks for
    // the [points x 512] formulation (kept for A/B runs and for hidden widths the factored kernels do not cover)
    static const bool dense_env = []() { const char* e = getenv("OO_BG_CLIP"); return e != nullptr && strcmp(e, "dense") == 0; }();
    const bool factored = part && !dense_env && h <= 256;
    OO_TRY(bg_forward(theta, L, h, pcs, M, scale, w, part && !factored, nullptr, st));
    if (factored) {
        OO_TRY(bg_clip_render(w.alpha, w.hp, n_rays, n_samp, h, w.hs, w.Sx, st));
        GemmOp g = op(w.Sx, w.hs, 1, w.wb_ocl, w.hs, 1, w.X, C, 1, n_rays, C, w.hs);          // x_r = W S_r + b opac_r
        OO_TRY(run_gemm(g, 1, nullptr, st));
        // ---- depth / colour / opacity terms of loss.step_batch_loss on [1, R, S] (K3 without features), then the feature term
        OO_TRY(oo_loss_fwd(w.alpha, w.color, z, gt_depth, w.gt_rgb, labels, nullptr, nullptr, 1, n_rays, n_samp, 0, color_scaling,
                           opacity_scaling, feat_scaling, terms_out, loss_out, flags_out, w.loss_ws, stream));
        const float* tail = w.loss_ws + (size_t)n_rays * oo_loss_ws_per_ray();
        OO_TRY(bg_clip_loss(w.X, w.gt_feat, labels, flags_out, tail, n_rays, C, feat_scaling, w.d_x, w.lf, terms_out, loss_out, st));
        g = op(w.d_x, C, 1, w.wb_ocl, 1, w.hs, w.dSx, w.hs, 1, n_rays, w.hs, C);               // [dS | d opac] = d_x [W | b]
        OO_TRY(run_gemm(g, BG_SPLIT, w.part_ds, st));
        OO_TRY(bg_clip_bwd(w.alpha, w.hp, w.dSx, n_rays, n_samp, h, w.hs, w.hu, w.d_hp, st));
        OO_TRY(loss_bwd_hu(w.alpha, w.color, z, gt_depth, w.gt_rgb, labels, nullptr, nullptr, 1, n_rays, n_samp, 0, color_scaling,
                           opacity_scaling, feat_scaling, 1.f, flags_out, w.loss_ws, w.d_alpha, w.d_color, nullptr, w.hu, stream));
        OO_TRY(bg_backward(theta, L, h, pcs, M, scale, w, part, G, st, nullptr, n_rays));
    } else {
    // ---- loss.step_batch_loss on [1, R, S] (train.py:452-462) and its gradient w.r.t. alpha / colour / clip (K3)
    OO_TRY(oo_loss_fwd(w.alpha, w.color, z, gt_depth, w.gt_rgb, labels, part ? w.clip : nullptr, part ? w.gt_feat : nullptr, 1,
                       n_rays, n_samp, part ? C : 0, color_scaling, opacity_scaling, feat_scaling, terms_out, loss_out,
                       flags_out, w.loss_ws, stream));
    OO_TRY(oo_loss_bwd(w.alpha, w.color, z, gt_depth, w.gt_rgb, labels, part ? w.clip : nullptr, part ? w.gt_feat : nullptr, 1,
                       n_rays, n_samp, part ? C : 0, color_scaling, opacity_scaling, feat_scaling, 1.f, flags_out, w.loss_ws,
                       w.d_alpha, w.d_color, part ? w.d_clip : nullptr, stream));
    OO_TRY(bg_backward(theta, L, h, pcs, M, scale, w, part, G, st));
    }
    if (grads_out) return 0;
    // ---- torch.optim.AdamW over the flat block, per parameter group: a group autograd does not reach in this step (part
    // features off: the clip head, quirk 8; an empty label mask: see k_bg_adam_sched) is skipped entirely
    k_bg_adam_sched<<<1, 32, 0, st>>>(flags_out, part ? 1 : 0, (double)lr, (double)beta1, (double)beta2, adam_t, w.adam_scal);
    OO_LAUNCH_CHECK();
    k_bg_adamw<<<148 * 4, 256, 0, st>>>(theta, G, adam_m, adam_v, L.total, L.off[T_CL_W], L.off[T_CP_W], L.off[T_PE], w.adam_scal,
                                        (float)(1.0 - (double)lr * (double)weight_decay), beta1, beta2, eps);
    OO_LAUNCH_CHECK();
    return 0;
}
