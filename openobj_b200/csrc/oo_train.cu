// K1 fused per-object encode + MLP + compositing + loss + backward (one persistent CTA per SM),
// K4 fused multi-tensor AdamW over the stacked parameter blocks, and the per-frame bookkeeping kernels.
// Reference path replaced: objnerf/train.py:394-474 (vmap(pe), vmap(fc), loss.step_batch_loss,
// backward, AdamW.step), utils.py:55-62 (combine_state_for_ensemble).
#include "../../include/openobj_b200.h"
#include "oo_common.cuh"
#include "oo_sched.h"
#include "oo_tile.h"

using namespace oo;

namespace {

struct TrainParams {
    const float* theta;
    float* derived;        // [n_obj][DERIVED] out_clip constants of every object (k_gram)
    float* rayrec;         // [n_obj][R][RAYREC] (this step)
    oo_batch b;
    int ray0;              // first ray of this step inside each object's batch (it * rays_per_step)
    int R;                 // rays per object per step
    int tiles_per_obj;
    int n_obj;
    const int* sched;      // cta_tile_begin[n_cta+1], cta_slot_begin[n_cta], ...
    const int* counts;     // [n_obj][2] for this step
    const int* flags;      // [1] for this step
    float* slab;
    float* slot_loss;
    float scale, cs, os, fs;
    long long* phase_cycles;   // debug: [N_TRAIN_PHASES + 8] cycle totals of block 0: phases, block, tiles, staging, flush,
                               // prologue (start -> past the dependency wait), wait (inside griddepcontrol.wait), tail
    long long* block_times;    // debug: [n_cta][3] = {globaltimer at block start, at block end, tiles} of the last launch
};

// Programmatic dependent launch: the two kernels of a step are launched with programmaticStreamSerialization, so the CTAs of
// the next kernel become resident as soon as SMs free up and run their prologue (everything that does not depend on the
// previous kernel's output) before pdl_wait(); a kernel releases its own dependents only AFTER its wait, so the chain
// k_update(i-1) -> k_train(i) -> k_update(i) stays transitively ordered.
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

long long* g_phase_cycles = nullptr;
long long* g_block_times = nullptr;

__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

template <int PH, int END, bool PART>
struct Phases {
    static __device__ __forceinline__ void run(int tid, float* sm, const TileCtx& c, TileAcc& a, long long* cyc) {
        long long t0 = 0;
        if (cyc) t0 = clock64();
        tile_phase<kTrainOrder[PH], PART>(tid, sm, c, a);
        __syncthreads();
        if (cyc && tid == 0) cyc[PH] += clock64() - t0;
        Phases<PH + 1, END, PART>::run(tid, sm, c, a, cyc);
    }
};
template <int END, bool PART>
struct Phases<END, END, PART> {
    static __device__ __forceinline__ void run(int, float*, const TileCtx&, TileAcc&, long long*) {}
};

// G = W_ocl^T W_ocl, wb = W_ocl^T b_ocl, bb = b.b of every object (DESIGN.md "clip head inside K1"); one CTA per object
__global__ void __launch_bounds__(NTHREADS, 1) k_gram(const float* __restrict__ theta, float* __restrict__ derived) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const float* th = theta + (size_t)blockIdx.x * PSTRIDE;
    float* der = derived + (size_t)blockIdx.x * DERIVED;
    gram_stage<-1>(tid, sm, th, der);
    gram_stage<0>(tid, sm, th, der); __syncthreads();
    gram_stage<1>(tid, sm, th, der); __syncthreads();
    gram_stage<2>(tid, sm, th, der);
}

template <bool PART>
__global__ void __launch_bounds__(NTHREADS, 1) k_train(const TrainParams prm) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int n_cta = gridDim.x;
    const int t_begin = prm.sched[blockIdx.x], t_end = prm.sched[blockIdx.x + 1];
    if (t_begin >= t_end) return;
    int slot = prm.sched[n_cta + 1 + blockIdx.x];
    const int flags = prm.flags[0];

    if (prm.block_times != nullptr && tid == 0) {
        prm.block_times[3 * blockIdx.x] = global_ns();
        prm.block_times[3 * blockIdx.x + 2] = t_end - t_begin;
    }
    long long* cyc = (prm.phase_cycles != nullptr && blockIdx.x == 0) ? prm.phase_cycles : nullptr;
    const long long t_start = cyc ? clock64() : 0;
    const long long ns_start = cyc ? global_ns() : 0;
    zero_pad_rows(tid, sm);
    // tensor memory for the weight-gradient accumulators: 16 warps x 64 columns (oo_tile.h, TileAcc); one CTA per SM
    __shared__ uint32_t tm_base_s;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tm_base_s)), "r"((uint32_t)(AC_COLS * NWARPS / 4)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_base = tm_base_s;
    TileAcc acc;
    acc.tm = tm_base + ((uint32_t)(32 * ((tid >> 5) & 3)) << 16) + (uint32_t)(AC_COLS * (tid >> 7));
    acc_zero(acc);
    TileCtx c;
    c.flags = flags;
    c.scale = prm.scale;
    c.cs = prm.cs; c.os = prm.os; c.fs = prm.fs;
    c.feat_table = prm.b.feat_table;

    // per-tile pointers (everything that depends only on the tile index)
    auto tile_ptrs = [&](int t, TileCtx& x) {
        const int obj = t / prm.tiles_per_obj;
        const int r0 = (t - obj * prm.tiles_per_obj) * RT;
        const size_t ray = (size_t)obj * prm.b.rays_per_obj + prm.ray0 + r0;
        x.nrays = min(RT, prm.R - r0);
        x.npts = x.nrays * S;
        x.pcs = prm.b.pcs + ray * (S * 3);
        x.z = prm.b.z + ray * S;
        x.gt_depth = prm.b.gt_depth + ray;
        x.gt_rgb = prm.b.gt_rgb + ray * 3;
        x.labels = prm.b.labels + ray;
        x.feat_row = PART ? prm.b.feat_row + ray : nullptr;
        x.rayrec = prm.rayrec + ((size_t)obj * prm.R + r0) * RAYREC;
    };
    tile_ptrs(t_begin, c);
    float pre = tile_prefetch<PART>(tid, c);          // this thread's input value of the first tile
    // everything above touches only this CTA's shared / tensor memory and per-frame constants; the parameters and the
    // out_clip constants below come from the previous update kernel
    const long long t_w0 = cyc ? clock64() : 0;
    pdl_wait();
    pdl_release();
    if (cyc && tid == 0) {
        const long long now = clock64();
        cyc[N_TRAIN_PHASES + 4] += t_w0 - t_start;
        cyc[N_TRAIN_PHASES + 5] += now - t_w0;
    }
    int cur_obj = -1;
    for (int t = t_begin; t < t_end; ++t) {
        const int obj = t / prm.tiles_per_obj;
        if (obj != cur_obj) {
            const long long ts0 = cyc ? clock64() : 0;
            cur_obj = obj;
            c.theta = prm.theta + (size_t)obj * PSTRIDE;
            c.derived = prm.derived + (size_t)obj * DERIVED;
            c.slab = prm.slab + (size_t)slot * PSTRIDE;
            c.inv1 = 1.f / ((float)prm.counts[2 * obj] + 1e-10f);
            c.invs = 1.f / ((float)prm.counts[2 * obj + 1] + 1e-10f);
            stage_weights(tid, sm, c.theta);
            if (PART) stage_derived(tid, sm, c.derived);
            __syncthreads();
            if (cyc && tid == 0) cyc[N_TRAIN_PHASES + 2] += clock64() - ts0;
        }
        float nxt = 0.f;
        if (t + 1 < t_end) {                           // the next tile's inputs start their trip now
            TileCtx nx;
            tile_ptrs(t + 1, nx);
            nxt = tile_prefetch<PART>(tid, nx);
        }
        tile_ptrs(t, c);
        {
            const long long t0 = cyc ? clock64() : 0;
            tile_phase0_pre<PART>(tid, sm, c, pre);
            __syncthreads();
            if (cyc && tid == 0) cyc[0] += clock64() - t0;
        }
        Phases<1, N_TRAIN_PHASES, PART>::run(tid, sm, c, acc, cyc);
        pre = nxt;

        const bool last_of_obj = (t + 1 == t_end) || ((t + 1) / prm.tiles_per_obj != obj);
        if (last_of_obj) {
            const long long tf0 = cyc ? clock64() : 0;
            float* sl = prm.slot_loss + 4 * slot;
            tile_flush<0, PART>(tid, sm, c.slab, sl, acc); __syncthreads();
            tile_flush<1, PART>(tid, sm, c.slab, sl, acc); __syncthreads();
            acc_zero(acc);
            ++slot;
            cur_obj = -1;
            if (cyc && tid == 0) cyc[N_TRAIN_PHASES + 3] += clock64() - tf0;
        }
    }
    const long long t_tail = cyc ? clock64() : 0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (prm.block_times != nullptr && tid == 0) prm.block_times[3 * blockIdx.x + 1] = global_ns();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"((uint32_t)(AC_COLS * NWARPS / 4))
                     : "memory");
    if (cyc && tid == 0) {
        const long long now = clock64();
        cyc[N_TRAIN_PHASES] += now - t_start;           // whole block
        cyc[N_TRAIN_PHASES + 1] += t_end - t_begin;     // tiles processed
        cyc[N_TRAIN_PHASES + 6] += now - t_tail;
        cyc[N_TRAIN_PHASES + 7] += global_ns() - ns_start;     // the same interval in nanoseconds: cycles / ns = SM clock
    }
}

// ---- per-frame: ray counts per (step, object) and the cross-object zero-mask flags (render_rays.py:88-94)
__global__ void __launch_bounds__(256) k_label_counts(const uint8_t* __restrict__ labels, int n_obj, int rays_per_obj, int R,
                                                      int* __restrict__ counts, int* __restrict__ flags, int* __restrict__ flag_bits) {
    // one block per step; a warp takes objects w, w + 8, ... and counts its R labels with ballots (no block barrier per object:
    // the first version walked the objects one after the other with two __syncthreads_count each, 38 us per frame at N = 60)
    __shared__ int s_f[8];
    const int it = blockIdx.x, lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    int f = 0;
    for (int o = wv; o < n_obj; o += 8) {
        const uint8_t* lab = labels + (size_t)o * rays_per_obj + (size_t)it * R;
        int n1 = 0, ns = 0;
        for (int r0 = 0; r0 < R; r0 += 32) {
            const int r = r0 + lane;
            const int l = r < R ? lab[r] : 2;
            n1 += __popc(__ballot_sync(0xffffffffu, r < R && l == 1));
            ns += __popc(__ballot_sync(0xffffffffu, r < R && l != 2));
        }
        if (lane == 0) {
            counts[((size_t)it * n_obj + o) * 2 + 0] = n1;
            counts[((size_t)it * n_obj + o) * 2 + 1] = ns;
        }
        if (n1 == 0) f |= OO_FLAG_NO_OBJ;
        if (ns == 0) f |= OO_FLAG_NO_SEM;
    }
    if (lane == 0) s_f[wv] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
        f = 0;
        for (int w = 0; w < 8; ++w) f |= s_f[w];
        flags[it] = f;
        if (flag_bits != nullptr) {           // one int per bit: a MAX all-reduce over ranks is then the OR the rule needs
            flag_bits[2 * it] = (f & OO_FLAG_NO_OBJ) ? 1 : 0;
            flag_bits[2 * it + 1] = (f & OO_FLAG_NO_SEM) ? 1 : 0;
        }
    }
}

// ---- per-frame: which parameter groups autograd reaches in each step, their Adam step numbers and bias
// corrections (torch.optim.AdamW: tensors whose grad is None are skipped entirely)
__global__ void k_adam_schedule(int* __restrict__ flags, const int* __restrict__ reduced_bits, int iters, int part_on, double lr,
                                double b1, double b2, int* __restrict__ adam_t, float* __restrict__ scal) {
    // one thread per step; the step number of a group = its counter + number of active steps up to and including `it`
    __shared__ int s_act[3][1024];
    const int it = threadIdx.x;
    bool active[3] = {false, false, false};
    if (it < iters) {
        int f = flags[it];
        if (reduced_bits != nullptr) {        // sharded run: the zero-mask bits OR-ed over all ranks replace the local ones
            f = (reduced_bits[2 * it] ? OO_FLAG_NO_OBJ : 0) | (reduced_bits[2 * it + 1] ? OO_FLAG_NO_SEM : 0);
            flags[it] = f;
        }
        const bool obj_terms = !(f & OO_FLAG_NO_OBJ), op_term = !(f & OO_FLAG_NO_SEM);
        active[0] = obj_terms || op_term;
        active[1] = obj_terms;
        active[2] = obj_terms && part_on != 0;
    }
    for (int g = 0; g < 3; ++g) s_act[g][it] = active[g] ? 1 : 0;
    __syncthreads();
    if (it < iters) {
        for (int g = 0; g < 3; ++g) {
            int t = adam_t[g];
            for (int q = 0; q <= it; ++q) t += s_act[g][q];
            float* s = scal + ((size_t)it * 3 + g) * 4;
            if (active[g]) {
                const double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
                s[0] = 1.f;
                s[1] = (float)(lr / bc1);
                s[2] = (float)sqrt(bc2);
                s[3] = (float)t;
            } else {
                s[0] = s[1] = s[2] = s[3] = 0.f;
            }
        }
    }
    __syncthreads();
    if (it == 0) {
        for (int g = 0; g < 3; ++g) {
            int t = adam_t[g];
            for (int q = 0; q < iters; ++q) t += s_act[g][q];
            adam_t[g] = t;
        }
    }
}

// ---- K4a: out_clip gradient of each object from what K1 left:
//   dW[c][j] = sum_r A_r y_r[c] S_r[j] + sum_k W[c][k] M[k][j] + b[c] m[j],   db[c] = sum_r A_r opac_r y_r[c] + W[c].m + b[c] beta
struct ClipGrad {
    const float* rayrec;       // [n_obj][R][RAYREC]
    const int32_t* feat_row;   // batch feat_row + ray0 (this step), stride rays_per_obj per object
    const float* feat_table;
    int rays_per_obj, R;
};

__global__ void __launch_bounds__(512) k_clipgrad(const float* __restrict__ theta, const float* __restrict__ slab,
                                                  const int* __restrict__ obj_slot, const ClipGrad cg,
                                                  float* __restrict__ clip_grad) {
    // grid = (4 chunks of 128 feature rows, n_obj).  Rays whose coefficient A_r is zero (label != 1, or the zero-mask rule)
    // are compacted away first; then a [128 x n_act] x [n_act x 32] product with A_r y_r staged in shared memory.
    extern __shared__ float sh[];          // [1088] M, m, beta ; [R] active list ; [R][RAYREC] records ; [R][128] y_r[c]
    __shared__ int s_warp[33];
    const int o = blockIdx.y, c_lo = 128 * blockIdx.x, tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int s0 = obj_slot[o], s1 = obj_slot[o + 1];
    const int Rp = (cg.R + 3) & ~3;
    int* act = reinterpret_cast<int*>(sh + DERIVED);
    float* rec = sh + DERIVED + Rp;
    float* ys = rec + (size_t)cg.R * RAYREC;
    const float* src = cg.rayrec + (size_t)o * cg.R * RAYREC;
    // this thread's out_clip row and bias: issued first, consumed last
    const float* th = theta + (size_t)o * PSTRIDE;
    const int cl = tid >> 2, j0 = 8 * (tid & 3), cc = c_lo + cl;
    float wr[H];
#pragma unroll
    for (int k = 0; k < H; k += 4) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(th + OFF_OCL_W + cc * H + k));
        wr[k] = w4.x; wr[k + 1] = w4.y; wr[k + 2] = w4.z; wr[k + 3] = w4.w;
    }
    const float bc = __ldg(th + OFF_OCL_B + cc);
    for (int q = tid; q < 1057; q += 512) {            // slot totals, 4 independent loads per round trip
        float t = 0.f;
        for (int sl = s0; sl < s1; sl += 4) {
            const float* p0 = slab + (size_t)sl * PSTRIDE + SLAB_M + q;
            const float v0 = p0[0];
            const float v1 = sl + 1 < s1 ? p0[PSTRIDE] : 0.f;
            const float v2 = sl + 2 < s1 ? p0[2 * (size_t)PSTRIDE] : 0.f;
            const float v3 = sl + 3 < s1 ? p0[3 * (size_t)PSTRIDE] : 0.f;
            t += v0; t += v1; t += v2; t += v3;
        }
        sh[q] = t;
    }
    // ---- compaction of the active rays (order preserved: deterministic summation order)
    int n_act = 0;
    for (int base = 0; base < cg.R; base += 512) {
        const int r = base + tid;
        const bool a = r < cg.R && src[(size_t)r * RAYREC + REC_A] != 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, a);
        if (lane == 0) s_warp[wv] = __popc(m);
        __syncthreads();
        int off = n_act;
        for (int w = 0; w < wv; ++w) off += s_warp[w];
        if (a) act[off + __popc(m & ((1u << lane) - 1u))] = r;
        int tot = 0;
        for (int w = 0; w < 16; ++w) tot += s_warp[w];
        n_act += tot;
        __syncthreads();
    }
    // ---- records and gt feature rows of the active rays
    for (int q = tid; q < n_act * RAYREC; q += 512) {
        const int i = q / RAYREC, e = q - i * RAYREC;
        rec[q] = src[(size_t)act[i] * RAYREC + e];
    }
    {
        const int cc = tid & 127, n_it = (n_act * 128 + 511) / 512;
        for (int i0 = 0; i0 < n_it; i0 += 6) {
            float v[6];
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int i = ((i0 + u) * 512 + tid) >> 7;
                v[u] = i < n_act ? __ldg(cg.feat_table + (size_t)cg.feat_row[(size_t)o * cg.rays_per_obj + act[i]] * C + c_lo + cc) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int i = ((i0 + u) * 512 + tid) >> 7;
                if (i < n_act) ys[i * 128 + cc] = v[u];
            }
        }
    }
    __syncthreads();
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float gb = 0.f;
#pragma unroll 4
    for (int i = 0; i < n_act; ++i) {
        const float ay = rec[i * RAYREC + REC_A] * ys[i * 128 + cl];
        const float4 Sa = *reinterpret_cast<const float4*>(rec + i * RAYREC + REC_S + j0);
        const float4 Sb = *reinterpret_cast<const float4*>(rec + i * RAYREC + REC_S + j0 + 4);
        g[0] += ay * Sa.x; g[1] += ay * Sa.y; g[2] += ay * Sa.z; g[3] += ay * Sa.w;
        g[4] += ay * Sb.x; g[5] += ay * Sb.y; g[6] += ay * Sb.z; g[7] += ay * Sb.w;
        gb += ay * rec[i * RAYREC + REC_OPAC];
    }
    float wm = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) {
        const float* M = sh + k * H + j0;
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] += wr[k] * M[q];
        wm += wr[k] * sh[1024 + k];
    }
    float* out = clip_grad + (size_t)o * (C * H + C);
#pragma unroll
    for (int q = 0; q < 8; ++q) out[cc * H + j0 + q] = g[q] + bc * sh[1024 + j0 + q];
    if ((tid & 3) == 0) out[C * H + cc] = gb + wm + bc * sh[1056];
}

// ---- K4b: sum the gradient slots of each object (fixed order) and apply AdamW in place; HBM-bound.
// grid = (PSTRIDE/4/256, n_obj); thread = 4 consecutive parameters of one object.
template <bool UPDATE>
__global__ void __launch_bounds__(256) k_adamw(float* __restrict__ theta, float* __restrict__ am, float* __restrict__ av,
                                               const float* __restrict__ slab, const int* __restrict__ obj_slot,
                                               const float* __restrict__ scal, float decay, float b1, float b2, float eps,
                                               const float* __restrict__ clip_grad, const float* __restrict__ slot_loss,
                                               const int* __restrict__ counts, float* __restrict__ loss_terms,
                                               float* __restrict__ grads_out, int* __restrict__ step_flags) {
    const int o = blockIdx.y;
    const int i = 4 * (blockIdx.x * 256 + threadIdx.x);
    const int s0 = obj_slot[o], s1 = obj_slot[o + 1];
    if (blockIdx.x == 0 && threadIdx.x < 4) {
        // per-object loss terms: masked mean = sum / (count + 1e-10)  (render_rays.py:108); a term above 1e5 is where the
        // reference prints "loss explode" and exits (render_rays.py:109-111): bit OO_FLAG_EXPLODE of this step's flags
        float s = 0.f;
        for (int q = s0; q < s1; ++q) s += slot_loss[4 * q + threadIdx.x];
        const int n = counts[2 * o + (threadIdx.x == 2 ? 1 : 0)];
        const float term = s / ((float)n + 1e-10f);
        if (loss_terms != nullptr) loss_terms[4 * o + threadIdx.x] = term;
        if (term > 100000.f) atomicOr(step_flags, OO_FLAG_EXPLODE);
    }
    if (i >= PEND) return;
    const int grp = group_of_offset(i);
    const bool active = scal[grp * 4] != 0.f;
    float4 g = {0.f, 0.f, 0.f, 0.f};
    const size_t idx = (size_t)o * PSTRIDE + i;
    float4 p = {0.f, 0.f, 0.f, 0.f}, m = p, v = p;
    if (UPDATE && active) {                       // issued before the gradient loads: all independent
        p = *reinterpret_cast<const float4*>(theta + idx);
        m = *reinterpret_cast<const float4*>(am + idx);
        v = *reinterpret_cast<const float4*>(av + idx);
    }
    if (active) {
        if (i >= OFF_OCL_W && i < OFF_OCL_B + C) {
            if (clip_grad != nullptr) g = *reinterpret_cast<const float4*>(clip_grad + (size_t)o * (C * H + C) + (i - OFF_OCL_W));
        } else {
            const float4 z4 = {0.f, 0.f, 0.f, 0.f};
            for (int q = s0; q < s1; q += 4) {    // slots summed in order, 4 loads per round trip
                const float* sp = slab + (size_t)q * PSTRIDE + i;
                const float4 v0 = *reinterpret_cast<const float4*>(sp);
                const float4 v1 = q + 1 < s1 ? *reinterpret_cast<const float4*>(sp + PSTRIDE) : z4;
                const float4 v2 = q + 2 < s1 ? *reinterpret_cast<const float4*>(sp + 2 * (size_t)PSTRIDE) : z4;
                const float4 v3 = q + 3 < s1 ? *reinterpret_cast<const float4*>(sp + 3 * (size_t)PSTRIDE) : z4;
                g.x += v0.x; g.y += v0.y; g.z += v0.z; g.w += v0.w;
                g.x += v1.x; g.y += v1.y; g.z += v1.z; g.w += v1.w;
                g.x += v2.x; g.y += v2.y; g.z += v2.z; g.w += v2.w;
                g.x += v3.x; g.y += v3.y; g.z += v3.z; g.w += v3.w;
            }
        }
    }
    if (!UPDATE) {
        *reinterpret_cast<float4*>(grads_out + idx) = g;
        return;
    }
    if (!active) return;
    const float step = scal[grp * 4 + 1], bc2s = scal[grp * 4 + 2];
#define OO_ADAM1(P, M, V, G)                               \
    {                                                      \
        P = P * decay;                                     \
        M = M + (G - M) * (1.f - b1);                      \
        V = V * b2 + ((1.f - b2) * G) * G;                 \
        const float den = sqrtf(V) / bc2s + eps;           \
        P = P - step * (M / den);                          \
    }
    OO_ADAM1(p.x, m.x, v.x, g.x)
    OO_ADAM1(p.y, m.y, v.y, g.y)
    OO_ADAM1(p.z, m.z, v.z, g.z)
    OO_ADAM1(p.w, m.w, v.w, g.w)
#undef OO_ADAM1
    *reinterpret_cast<float4*>(theta + idx) = p;
    *reinterpret_cast<float4*>(am + idx) = m;
    *reinterpret_cast<float4*>(av + idx) = v;
}

__global__ void k_adamw_flat(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, int64_t n, float decay, float b1, float b2, float eps,
                             float step, float bc2s) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float P = p[i] * decay, G = g[i];
        const float M = m[i] + (G - m[i]) * (1.f - b1);
        const float V = v[i] * b2 + ((1.f - b2) * G) * G;
        P = P - step * (M / (sqrtf(V) / bc2s + eps));
        p[i] = P; m[i] = M; v[i] = V;
    }
}

// ---- K4 fused (the path oo_train_step / oo_train_frame take): ONE launch per step does
//   (a) the out_clip gradient assembly of k_clipgrad, applied straight to out_clip.{weight,bias} with AdamW,
//   (b) slab reduction + AdamW of every other tensor (k_adamw), a quarter of the block per CTA,
//   (c) the out_clip constants K1 needs in the NEXT step (k_gram: G = W^T W, wb = W^T b, bb = b.b of the UPDATED layer):
//       each CTA forms the partial product of its 128 rows, the last CTA of an object to arrive sums the four partials
//       in a fixed order (deterministic).
// grid = (4 chunks of 128 out_clip rows, n_obj), 256 threads: thread = a 4 x 4 block of the chunk's [128 x 32] gradient.
constexpr int UPD_CHUNKS = 4, UPD_ROWS = C / UPD_CHUNKS, UPD_THREADS = 256;
constexpr int GPART = 1092;                   // 33 x 33 partial Gram matrix, padded
constexpr int UPD_WN = UPD_ROWS * GS;         // staged updated rows [128][36]
constexpr int UPD_GT = 81, UPD_RG = 3;        // 9 x 9 tiles of 4 x 4 over 36 x 36; three row groups

__device__ __forceinline__ void adam1(float& P, float& M, float& V, float G, float decay, float b1, float b2, float eps,
                                      float step, float bc2s) {
    P = P * decay;
    M = M + (G - M) * (1.f - b1);
    V = V * b2 + ((1.f - b2) * G) * G;
    const float den = sqrtf(V) / bc2s + eps;
    P = P - step * (M / den);
}

__device__ __forceinline__ void adam4(float* theta, float* am, float* av, size_t idx, const float4& g, float decay, float b1,
                                      float b2, float eps, float step, float bc2s) {
    float4 p = *reinterpret_cast<const float4*>(theta + idx);
    float4 m = *reinterpret_cast<const float4*>(am + idx);
    float4 v = *reinterpret_cast<const float4*>(av + idx);
    adam1(p.x, m.x, v.x, g.x, decay, b1, b2, eps, step, bc2s);
    adam1(p.y, m.y, v.y, g.y, decay, b1, b2, eps, step, bc2s);
    adam1(p.z, m.z, v.z, g.z, decay, b1, b2, eps, step, bc2s);
    adam1(p.w, m.w, v.w, g.w, decay, b1, b2, eps, step, bc2s);
    *reinterpret_cast<float4*>(theta + idx) = p;
    *reinterpret_cast<float4*>(am + idx) = m;
    *reinterpret_cast<float4*>(av + idx) = v;
}

constexpr int UPD_YS_MIN = UPD_RG * 1296;   // the three Gram row-group partials alias the gt-feature staging area

inline size_t update_smem_floats(int R, bool part) {
    if (!part) return 0;
    const size_t ys = (size_t)R * UPD_ROWS;
    return DERIVED + 2 * (size_t)((R + 3) & ~3) + (size_t)R * RAYREC + UPD_WN + (ys > UPD_YS_MIN ? ys : UPD_YS_MIN);
}

// The kernel is a chain of dependent memory round trips, so every stage issues all of its loads before the first use:
// stage 1 = parameters / moments / slab slots of this CTA's quarter, the M, m, beta slot sums, the out_clip rows (cp.async)
// and the ray coefficients; stage 2 = feature-table rows of the active rays; stage 3 = the gt-feature gather and the ray
// records (cp.async) together with the moments of the out_clip rows.
template <bool PART>
__global__ void __launch_bounds__(UPD_THREADS, 2) k_update(float* theta, float* am, float* av, const float* __restrict__ slab,
                                                           const int* __restrict__ obj_slot, const float* __restrict__ scal,
                                                           float decay, float b1, float b2, float eps, const ClipGrad cg,
                                                           const float* __restrict__ slot_loss, const int* __restrict__ counts,
                                                           float* __restrict__ loss_terms, float* __restrict__ derived,
                                                           float* gram_part, int* gram_cnt, int* __restrict__ step_flags) {
    extern __shared__ __align__(16) float sh[];   // [1088] M, m, beta | act [Rp] | frow [Rp] | rec [R][36] | wn [128][36] | ys [R][128]
    __shared__ int s_warp[UPD_THREADS / 32];
    __shared__ int s_last;
    const int o = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int s0 = obj_slot[o], s1 = obj_slot[o + 1];
    const int Rp = (cg.R + 3) & ~3;
    int* act = reinterpret_cast<int*>(sh + DERIVED);
    int* frow = act + Rp;
    float* rec = sh + DERIVED + 2 * Rp;
    float* wn = rec + (size_t)cg.R * RAYREC;
    float* ys = wn + UPD_WN;
    float* gp = ys;                         // reused once the gradient is formed
    const int c_lo = UPD_ROWS * chunk;
    float* th = theta + (size_t)o * PSTRIDE;
    const bool clip_on = PART && scal[GROUP_CLIP * 4] != 0.f;
    const float* src = PART ? cg.rayrec + (size_t)o * cg.R * RAYREC : nullptr;

    if (PART) {
        // out_clip rows of this chunk -> wn [128][36] (col 32 = bias, 33..35 = 0); awaited before the gradient loop
        for (int q = tid; q < UPD_ROWS * (H / 4); q += UPD_THREADS) {
            const int row = q >> 3, v4 = q & 7;
            OO_CP_ASYNC16(wn + row * GS + 4 * v4, th + OFF_OCL_W + (c_lo + row) * H + 4 * v4);
        }
        if (tid < UPD_ROWS) *reinterpret_cast<float4*>(wn + tid * GS + H) = float4{th[OFF_OCL_B + c_lo + tid], 0.f, 0.f, 0.f};
    }
    // ---- (b) this CTA's quarter of the tensors whose gradient is the slot sum, and the slot sums of M, m, beta
    {
        constexpr int NV_LO = OFF_OCL_W / 4, NV = NV_LO + (PEND - OFF_PE_B) / 4, PER = (NV + UPD_CHUNKS - 1) / UPD_CHUNKS;
        constexpr int NQ = (PER + UPD_THREADS - 1) / UPD_THREADS, NM = (1057 + UPD_THREADS - 1) / UPD_THREADS;
        const int q1 = min(NV, (chunk + 1) * PER);
        float4 p[NQ], m[NQ], v[NQ], g[NQ];
        int off[NQ], grp[NQ];
        bool on[NQ];
        float tm[NM];
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
            const int q = chunk * PER + tid + UPD_THREADS * u;
            off[u] = q < NV_LO ? 4 * q : OFF_PE_B + 4 * (q - NV_LO);
            grp[u] = group_of_offset(off[u]);
            on[u] = q < q1 && scal[grp[u] * 4] != 0.f;
            g[u] = float4{0.f, 0.f, 0.f, 0.f};
            if (on[u]) {
                const size_t idx = (size_t)o * PSTRIDE + off[u];
                p[u] = *reinterpret_cast<const float4*>(theta + idx);
                m[u] = *reinterpret_cast<const float4*>(am + idx);
                v[u] = *reinterpret_cast<const float4*>(av + idx);
            }
        }
#pragma unroll
        for (int u = 0; u < NM; ++u) tm[u] = 0.f;
        // parameters, moments and the out_clip rows above were last written by the previous update kernel (complete: K1 only
        // released this launch after its own wait); everything below reads what K1 wrote
        pdl_wait();
        pdl_release();
        if (chunk == 0 && tid < 4) {
            // per-object loss terms: masked mean = sum / (count + 1e-10)  (render_rays.py:108); a term above 1e5 is where the
            // reference prints "loss explode" and exits (render_rays.py:109-111): bit OO_FLAG_EXPLODE of this step's flags
            // (k_train of this step has read its zero-mask bits already; the host looks at the flags once per frame)
            float s = 0.f;
            for (int q = s0; q < s1; ++q) s += slot_loss[4 * q + tid];
            const int n = counts[2 * o + (tid == 2 ? 1 : 0)];
            const float term = s / ((float)n + 1e-10f);
            if (loss_terms != nullptr) loss_terms[4 * o + tid] = term;
            if (term > 100000.f) atomicOr(step_flags, OO_FLAG_EXPLODE);
        }
        for (int sl = s0; sl < s1; sl += 2) {            // slots summed in order, two per round trip
            const float* sp = slab + (size_t)sl * PSTRIDE;
            const bool two = sl + 1 < s1;
            float4 x0[NQ], x1[NQ];
            float y0[NM], y1[NM];
#pragma unroll
            for (int u = 0; u < NQ; ++u) {
                x0[u] = x1[u] = float4{0.f, 0.f, 0.f, 0.f};
                if (on[u]) {
                    x0[u] = *reinterpret_cast<const float4*>(sp + off[u]);
                    if (two) x1[u] = *reinterpret_cast<const float4*>(sp + PSTRIDE + off[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < NM; ++u) {
                const int q = tid + UPD_THREADS * u;
                y0[u] = y1[u] = 0.f;
                if (clip_on && q < 1057) {
                    y0[u] = sp[SLAB_M + q];
                    if (two) y1[u] = sp[PSTRIDE + SLAB_M + q];
                }
            }
#pragma unroll
            for (int u = 0; u < NQ; ++u) {
                g[u].x += x0[u].x; g[u].y += x0[u].y; g[u].z += x0[u].z; g[u].w += x0[u].w;
                g[u].x += x1[u].x; g[u].y += x1[u].y; g[u].z += x1[u].z; g[u].w += x1[u].w;
            }
#pragma unroll
            for (int u = 0; u < NM; ++u) { tm[u] += y0[u]; tm[u] += y1[u]; }
        }
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
            if (!on[u]) continue;
            const float step = scal[grp[u] * 4 + 1], bc2s = scal[grp[u] * 4 + 2];
            adam1(p[u].x, m[u].x, v[u].x, g[u].x, decay, b1, b2, eps, step, bc2s);
            adam1(p[u].y, m[u].y, v[u].y, g[u].y, decay, b1, b2, eps, step, bc2s);
            adam1(p[u].z, m[u].z, v[u].z, g[u].z, decay, b1, b2, eps, step, bc2s);
            adam1(p[u].w, m[u].w, v[u].w, g[u].w, decay, b1, b2, eps, step, bc2s);
            const size_t idx = (size_t)o * PSTRIDE + off[u];
            *reinterpret_cast<float4*>(theta + idx) = p[u];
            *reinterpret_cast<float4*>(am + idx) = m[u];
            *reinterpret_cast<float4*>(av + idx) = v[u];
        }
        if (clip_on) {
#pragma unroll
            for (int u = 0; u < NM; ++u)
                if (tid + UPD_THREADS * u < 1057) sh[tid + UPD_THREADS * u] = tm[u];
        }
    }
    if (!PART) return;
    // ---- (a) gradient of out_clip rows [128 chunk, 128 chunk + 128) and their AdamW update
    if (clip_on) {
        // compaction of the rays with a non-zero coefficient A_r (order preserved: deterministic summation order)
        int n_act = 0;
        for (int base = 0; base < cg.R; base += UPD_THREADS) {
            const int r = base + tid;
            const bool a = r < cg.R && src[(size_t)r * RAYREC + REC_A] != 0.f;
            const unsigned mk = __ballot_sync(0xffffffffu, a);
            if (lane == 0) s_warp[wv] = __popc(mk);
            __syncthreads();
            int o2 = n_act, tot = 0;
#pragma unroll
            for (int w = 0; w < UPD_THREADS / 32; ++w) {
                if (w < wv) o2 += s_warp[w];
                tot += s_warp[w];
            }
            if (a) act[o2 + __popc(mk & ((1u << lane) - 1u))] = r;
            n_act += tot;
            __syncthreads();
        }
        for (int i = tid; i < n_act; i += UPD_THREADS) frow[i] = cg.feat_row[(size_t)o * cg.rays_per_obj + act[i]];
        __syncthreads();
        // gt feature rows (this chunk's 128 columns) and records of the active rays: all copies in flight together
        for (int q = tid; q < n_act * (UPD_ROWS / 4); q += UPD_THREADS) {
            const int i = q >> 5, v4 = q & 31;
            OO_CP_ASYNC16(ys + i * UPD_ROWS + 4 * v4, cg.feat_table + (size_t)frow[i] * C + c_lo + 4 * v4);
        }
        for (int q = tid; q < n_act * (RAYREC / 4); q += UPD_THREADS) {
            const int i = q / (RAYREC / 4), e = q - i * (RAYREC / 4);
            OO_CP_ASYNC16(rec + i * RAYREC + 4 * e, src + (size_t)act[i] * RAYREC + 4 * e);
        }
        // compute mapping: thread (tr = tid / 8, tc = tid % 8) owns the 4 x 4 block rows 4 tr.., columns 4 tc.. of the chunk's
        // [128 x 32] gradient (two 128-bit shared loads per 16 FMA).  Its moments travel during the gather.
        const int tr = tid >> 3, tc = tid & 7;
        float4 m4[4], v4r[4];
        float mb[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const size_t idx = (size_t)o * PSTRIDE + OFF_OCL_W + (size_t)(c_lo + 4 * tr + r) * H + 4 * tc;
            m4[r] = *reinterpret_cast<const float4*>(am + idx);
            v4r[r] = *reinterpret_cast<const float4*>(av + idx);
            if (tc == 0) {
                const size_t ib = (size_t)o * PSTRIDE + OFF_OCL_B + c_lo + 4 * tr + r;
                mb[r] = am[ib]; vb[r] = av[ib];
            }
        }
        OO_CP_ASYNC_WAIT();
        __syncthreads();
        // records -> B'_i = A_i [opac_i, S_i]: the gradient is then a plain product  g[row][j] = sum_i y_i[row] B'_i[j]
        for (int q = tid; q < n_act * (H + 1); q += UPD_THREADS) {
            const int i = q / (H + 1), e = q - i * (H + 1);
            const float Ai = rec[i * RAYREC + REC_A];
            if (e < H) rec[i * RAYREC + REC_S + e] *= Ai;
            else rec[i * RAYREC + 2] = Ai * rec[i * RAYREC + REC_OPAC];
        }
        __syncthreads();
        float g[4][4], gb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 4; ++r) g[r][0] = g[r][1] = g[r][2] = g[r][3] = 0.f;
#pragma unroll 4
        for (int i = 0; i < n_act; ++i) {
            const float4 y4 = *reinterpret_cast<const float4*>(ys + i * UPD_ROWS + 4 * tr);
            const float4 b4 = *reinterpret_cast<const float4*>(rec + i * RAYREC + REC_S + 4 * tc);
            const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                g[r][0] += yv[r] * b4.x; g[r][1] += yv[r] * b4.y; g[r][2] += yv[r] * b4.z; g[r][3] += yv[r] * b4.w;
            }
            if (tc == 0) {
                const float bo = rec[i * RAYREC + 2];
#pragma unroll
                for (int r = 0; r < 4; ++r) gb[r] += yv[r] * bo;
            }
        }
        // + W M + b m^T  (and W m + b beta for the bias)
        float wm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
        for (int k4 = 0; k4 < H; k4 += 4) {
            float4 wr4[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) wr4[r] = *reinterpret_cast<const float4*>(wn + (4 * tr + r) * GS + k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 M4 = *reinterpret_cast<const float4*>(sh + (k4 + kk) * H + 4 * tc);
                const float mk = sh[1024 + k4 + kk];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float wk = kk == 0 ? wr4[r].x : kk == 1 ? wr4[r].y : kk == 2 ? wr4[r].z : wr4[r].w;
                    g[r][0] += wk * M4.x; g[r][1] += wk * M4.y; g[r][2] += wk * M4.z; g[r][3] += wk * M4.w;
                    wm[r] += wk * mk;
                }
            }
        }
        float bcr[4];
        const float4 mj = *reinterpret_cast<const float4*>(sh + 1024 + 4 * tc);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            bcr[r] = wn[(4 * tr + r) * GS + H];
            g[r][0] += bcr[r] * mj.x; g[r][1] += bcr[r] * mj.y; g[r][2] += bcr[r] * mj.z; g[r][3] += bcr[r] * mj.w;
            gb[r] += wm[r] + bcr[r] * sh[1056];
        }
        __syncthreads();                     // every thread has read its rows of wn (and ys) before anything is overwritten
        const float step = scal[GROUP_CLIP * 4 + 1], bc2s = scal[GROUP_CLIP * 4 + 2];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = 4 * tr + r;
            const size_t idx = (size_t)o * PSTRIDE + OFF_OCL_W + (size_t)(c_lo + row) * H + 4 * tc;
            float4 pw = *reinterpret_cast<const float4*>(wn + row * GS + 4 * tc);
            adam1(pw.x, m4[r].x, v4r[r].x, g[r][0], decay, b1, b2, eps, step, bc2s);
            adam1(pw.y, m4[r].y, v4r[r].y, g[r][1], decay, b1, b2, eps, step, bc2s);
            adam1(pw.z, m4[r].z, v4r[r].z, g[r][2], decay, b1, b2, eps, step, bc2s);
            adam1(pw.w, m4[r].w, v4r[r].w, g[r][3], decay, b1, b2, eps, step, bc2s);
            *reinterpret_cast<float4*>(wn + row * GS + 4 * tc) = pw;
            *reinterpret_cast<float4*>(theta + idx) = pw;
            *reinterpret_cast<float4*>(am + idx) = m4[r];
            *reinterpret_cast<float4*>(av + idx) = v4r[r];
            if (tc == 0) {
                const size_t ib = (size_t)o * PSTRIDE + OFF_OCL_B + c_lo + row;
                float bc = bcr[r];
                adam1(bc, mb[r], vb[r], gb[r], decay, b1, b2, eps, step, bc2s);
                wn[row * GS + H] = bc;
                theta[ib] = bc; am[ib] = mb[r]; av[ib] = vb[r];
            }
        }
    } else {
        OO_CP_ASYNC_WAIT();
    }
    __syncthreads();
    // ---- (c) partial [W | b]^T [W | b] of the (updated) rows of this chunk
    if (tid < UPD_GT * UPD_RG) {
        const int t = tid % UPD_GT, rg = tid / UPD_GT;
        const int k4 = 4 * (t / 9), j4 = 4 * (t % 9);
        constexpr int RPG = (UPD_ROWS + UPD_RG - 1) / UPD_RG;
        const int r1 = min(UPD_ROWS, (rg + 1) * RPG);
        float acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x) acc[x][0] = acc[x][1] = acc[x][2] = acc[x][3] = 0.f;
#pragma unroll 4
        for (int r = rg * RPG; r < r1; ++r) {
            const float4 a4 = *reinterpret_cast<const float4*>(wn + r * GS + k4), b4 = *reinterpret_cast<const float4*>(wn + r * GS + j4);
            acc[0][0] += a4.x * b4.x; acc[0][1] += a4.x * b4.y; acc[0][2] += a4.x * b4.z; acc[0][3] += a4.x * b4.w;
            acc[1][0] += a4.y * b4.x; acc[1][1] += a4.y * b4.y; acc[1][2] += a4.y * b4.z; acc[1][3] += a4.y * b4.w;
            acc[2][0] += a4.z * b4.x; acc[2][1] += a4.z * b4.y; acc[2][2] += a4.z * b4.z; acc[2][3] += a4.z * b4.w;
            acc[3][0] += a4.w * b4.x; acc[3][1] += a4.w * b4.y; acc[3][2] += a4.w * b4.z; acc[3][3] += a4.w * b4.w;
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
            *reinterpret_cast<float4*>(gp + rg * 1296 + (k4 + x) * 36 + j4) = float4{acc[x][0], acc[x][1], acc[x][2], acc[x][3]};
    }
    __syncthreads();
    float* mine = gram_part + ((size_t)o * UPD_CHUNKS + chunk) * GPART;
    for (int q = tid; q < 33 * 33; q += UPD_THREADS) {
        const int k = q / 33, j = q - 33 * k;
        __stcg(mine + q, gp[k * 36 + j] + gp[1296 + k * 36 + j] + gp[2 * 1296 + k * 36 + j]);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(gram_cnt + o, 1) == UPD_CHUNKS - 1 ? 1 : 0;
    __syncthreads();
    if (s_last) {
        __threadfence();
        const float* gp0 = gram_part + (size_t)o * UPD_CHUNKS * GPART;
        float* der = derived + (size_t)o * DERIVED;
        for (int q = tid; q < 33 * 33; q += UPD_THREADS) {
            const int k = q / 33, j = q - 33 * k;
            const float v = (__ldcg(gp0 + q) + __ldcg(gp0 + GPART + q)) + (__ldcg(gp0 + 2 * GPART + q) + __ldcg(gp0 + 3 * GPART + q));
            if (k < H && j < H) der[DER_G + k * H + j] = v;
            else if (k < H) der[DER_WB + k] = v;            // column 32: W^T b
            else if (j == H) der[DER_BB] = v;               // b . b
        }
        if (tid == 0) gram_cnt[o] = 0;
    }
}

int check_train_args(int n_obj, const oo_batch* b, int rays_per_step, const oo_train_ws* ws) {
    OO_REQUIRE(n_obj > 0 && b && ws, "oo_train: null argument / n_obj <= 0");
    OO_REQUIRE(rays_per_step > 0 && b->rays_per_obj >= rays_per_step, "oo_train: bad rays_per_step");
    OO_REQUIRE(ws->slab && ws->slot_loss && ws->sched && ws->counts && ws->flags && ws->adam_scal && ws->derived && ws->rayrec && ws->clip_grad && ws->gram_part && ws->gram_cnt,
               "oo_train: workspace not allocated");
    return 0;
}

int launch_k1(const float* theta, int n_obj, const oo_batch* b, int it, int R, float scale, const oo_train_ws* ws,
              int n_sm, cudaStream_t st, bool gram = true) {
    Schedule s;                      // only the counts are needed here; the tables are already on the device
    s.tiles_per_obj = tiles_per_object(R);
    const long long T = (long long)n_obj * s.tiles_per_obj;
    const int n_cta = (int)(T < n_sm ? T : n_sm);
    TrainParams prm;
    prm.theta = theta;
    prm.derived = ws->derived;
    prm.rayrec = ws->rayrec;
    prm.b = *b;
    prm.ray0 = it * R;
    prm.R = R;
    prm.tiles_per_obj = s.tiles_per_obj;
    prm.n_obj = n_obj;
    prm.sched = ws->sched;
    prm.counts = ws->counts + (size_t)it * n_obj * 2;
    prm.flags = ws->flags + it;
    prm.slab = ws->slab;
    prm.slot_loss = ws->slot_loss;
    prm.scale = scale;
    prm.phase_cycles = g_phase_cycles;
    prm.block_times = g_block_times;
    prm.cs = 5.f; prm.os = 10.f; prm.fs = 5.f;   // loss.py:6 defaults (the JSON values are never read)
    const size_t smem = (size_t)SM_TOTAL * sizeof(float);
    static PerDevice attr_set;
    if (!attr_set.cur()) {
        OO_CUDA(cudaFuncSetAttribute(k_train<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OO_CUDA(cudaFuncSetAttribute(k_train<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set.cur() = 1;
    }
    if (b->feat_row != nullptr && gram) {
        const size_t gsmem = (size_t)(SM_GPART + NGG * 36 * 36) * sizeof(float);
        static PerDevice gattr;
        if (!gattr.cur()) {
            OO_CUDA(cudaFuncSetAttribute(k_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
            gattr.cur() = 1;
        }
        k_gram<<<n_obj, NTHREADS, gsmem, st>>>(theta, ws->derived);
        OO_LAUNCH_CHECK();
    }
    if (b->feat_row != nullptr) OO_CUDA(launch_pdl(k_train<true>, dim3(n_cta), dim3(NTHREADS), smem, st, prm));
    else OO_CUDA(launch_pdl(k_train<false>, dim3(n_cta), dim3(NTHREADS), smem, st, prm));
    OO_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// debug hook (not part of the ABI header): per-phase cycle counters of block 0
extern "C" int oo_debug_block_times(long long* dev_ptr) {     // [n_cta][3]; nullptr switches it off
    g_block_times = dev_ptr;
    return 3;
}

extern "C" int oo_debug_phase_cycles(long long* dev_ptr) {
    g_phase_cycles = dev_ptr;
    return N_TRAIN_PHASES + 8;
}

extern "C" int oo_train_ws_sizes(int n_obj, int rays_per_step, int iters, int n_sm, int* n_cta, int* n_slots,
                                 int64_t* slab_floats, int64_t* sched_ints_out) {
    OO_REQUIRE(n_obj > 0 && rays_per_step > 0 && n_sm > 0 && iters > 0, "oo_train_ws_sizes: bad argument");
    const Schedule s = build_schedule(n_obj, rays_per_step, n_sm);
    if (n_cta) *n_cta = s.n_cta;
    if (n_slots) *n_slots = s.n_slots;
    if (slab_floats) *slab_floats = (int64_t)s.n_slots * PSTRIDE;
    if (sched_ints_out) *sched_ints_out = (int64_t)s.data.size();
    return 0;
}

extern "C" int oo_train_schedule(int n_obj, int rays_per_step, int n_sm, oo_train_ws* ws, void* stream) {
    OO_REQUIRE(ws && ws->sched, "oo_train_schedule: workspace not allocated");
    const Schedule s = build_schedule(n_obj, rays_per_step, n_sm);
    OO_CUDA(cudaMemcpyAsync(ws->sched, s.data.data(), s.data.size() * sizeof(int), cudaMemcpyHostToDevice,
                            (cudaStream_t)stream));
    OO_CUDA(cudaStreamSynchronize((cudaStream_t)stream));   // `s` dies at return
    return 0;
}

extern "C" int oo_label_counts(const uint8_t* labels, int n_obj, int rays_per_obj, int rays_per_step, int iters,
                               int* counts, int* flags, int* flag_bits, void* stream) {
    OO_REQUIRE(counts && flags && iters > 0 && n_obj >= 0, "oo_label_counts: null argument");
    OO_REQUIRE(n_obj == 0 || labels, "oo_label_counts: null labels");
    OO_REQUIRE(n_obj == 0 || (long long)iters * rays_per_step <= rays_per_obj, "oo_label_counts: iters*rays_per_step > rays_per_obj");
    k_label_counts<<<iters, 256, 0, (cudaStream_t)stream>>>(labels, n_obj, rays_per_obj, rays_per_step, counts, flags, flag_bits);
    OO_LAUNCH_CHECK();
    return 0;
}

extern "C" int oo_adam_schedule(int* flags, const int* reduced_bits, int iters, int part_on, float lr, float beta1, float beta2,
                                int* adam_t, float* adam_scal, void* stream) {
    OO_REQUIRE(flags && adam_t && adam_scal, "oo_adam_schedule: null argument");
    OO_REQUIRE(iters >= 1 && iters <= 1024, "oo_adam_schedule: iters must be in [1, 1024]");
    k_adam_schedule<<<1, 1024, 0, (cudaStream_t)stream>>>(flags, reduced_bits, iters, part_on, (double)lr, (double)beta1,
                                                        (double)beta2, adam_t, adam_scal);
    OO_LAUNCH_CHECK();
    return 0;
}

static int launch_k4(float* theta, float* am, float* av, int n_obj, const oo_batch* b, int it, int R, float lr, float wd,
                     float b1, float b2, float eps, oo_train_ws* ws, float* loss_terms, float* grads_out, int n_sm,
                     cudaStream_t st) {
    const long long T = (long long)n_obj * tiles_per_object(R);
    const int n_cta = (int)(T < n_sm ? T : n_sm);
    const int* obj_slot = ws->sched + 2 * n_cta + 1;
    const dim3 grid(PSTRIDE / 4 / 256, n_obj);
    const float* scal = ws->adam_scal + (size_t)it * 12;
    const int* counts = ws->counts + (size_t)it * n_obj * 2;
    const bool part = b && b->feat_row;
    if (!grads_out) {
        // the training path: one fused launch (out_clip gradient + AdamW of everything + next step's out_clip constants)
        ClipGrad cg = {};
        cg.R = R;
        if (part) {
            cg.rayrec = ws->rayrec;
            cg.feat_row = b->feat_row + (size_t)it * R;
            cg.feat_table = b->feat_table;
            cg.rays_per_obj = b->rays_per_obj;
        }
        const size_t smem = update_smem_floats(R, part) * sizeof(float);
        OO_REQUIRE(smem <= 112 * 1024, "oo_train: rays_per_step too large for the update kernel's staging buffer");
        static PerDevice attr_smem_d;
        size_t& attr_smem = attr_smem_d.cur();
        if (smem > attr_smem || attr_smem == 0) {
            OO_CUDA(cudaFuncSetAttribute(k_update<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // same L1 / shared-memory split as k_train (which needs the maximum): no SM reconfiguration between the two
            // kernels of a step
            OO_CUDA(cudaFuncSetAttribute(k_update<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            OO_CUDA(cudaFuncSetAttribute(k_update<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            attr_smem = smem > 0 ? smem : 1;
        }
        const float decay = (float)(1.0 - (double)lr * (double)wd);
        const dim3 ugrid(UPD_CHUNKS, n_obj);
        const float* cslab = ws->slab;
        const float* cslot = ws->slot_loss;
        float* nullf = nullptr;
        int* nulli = nullptr;
        if (part)
            OO_CUDA(launch_pdl(k_update<true>, ugrid, dim3(UPD_THREADS), smem, st, theta, am, av, cslab, obj_slot, scal, decay, b1,
                               b2, eps, cg, cslot, counts, loss_terms, ws->derived, ws->gram_part, ws->gram_cnt, ws->flags + it));
        else
            OO_CUDA(launch_pdl(k_update<false>, ugrid, dim3(UPD_THREADS), (size_t)0, st, theta, am, av, cslab, obj_slot, scal, decay,
                               b1, b2, eps, cg, cslot, counts, loss_terms, nullf, nullf, nulli, ws->flags + it));
        OO_LAUNCH_CHECK();
        return 0;
    }
    if (part) {
        ClipGrad cg;
        cg.rayrec = ws->rayrec;
        cg.feat_row = b->feat_row + (size_t)it * R;
        cg.feat_table = b->feat_table;
        cg.rays_per_obj = b->rays_per_obj;
        cg.R = R;
        const size_t smem = (size_t)(DERIVED + ((R + 3) & ~3) + (size_t)R * RAYREC + (size_t)R * 128) * sizeof(float);
        OO_REQUIRE(smem <= 200 * 1024, "oo_train: rays_per_step too large for the out_clip gradient kernel's staging buffer");
        static PerDevice attr_smem_d;
        size_t& attr_smem = attr_smem_d.cur();
        if (smem > attr_smem) {
            OO_CUDA(cudaFuncSetAttribute(k_clipgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_smem = smem;
        }
        k_clipgrad<<<dim3(C / 128, n_obj), 512, smem, st>>>(theta, ws->slab, obj_slot, cg, ws->clip_grad);
        OO_LAUNCH_CHECK();
    }
    const float* cgrad = part ? ws->clip_grad : nullptr;
    if (grads_out) {
        k_adamw<false><<<grid, 256, 0, st>>>(theta, nullptr, nullptr, ws->slab, obj_slot, scal, 0.f, 0.f, 0.f, 0.f, cgrad,
                                             ws->slot_loss, counts, loss_terms, grads_out, ws->flags + it);
    } else {
        const float decay = (float)(1.0 - (double)lr * (double)wd);
        k_adamw<true><<<grid, 256, 0, st>>>(theta, am, av, ws->slab, obj_slot, scal, decay, b1, b2, eps, cgrad, ws->slot_loss,
                                            counts, loss_terms, nullptr, ws->flags + it);
    }
    OO_LAUNCH_CHECK();
    return 0;
}

static int train_step_impl(float* theta, float* am, float* av, int n_obj, const oo_batch* b, int it, int R, float scale,
                           float lr, float wd, float b1, float b2, float eps, oo_train_ws* ws, float* loss_terms,
                           float* grads_out, int n_sm, cudaStream_t st, bool gram = true) {
    if (int rc = check_train_args(n_obj, b, R, ws)) return rc;
    OO_REQUIRE((long long)(it + 1) * R <= b->rays_per_obj, "oo_train: step %d exceeds the pre-sampled batch", it);
    if (int rc = launch_k1(theta, n_obj, b, it, R, scale, ws, n_sm, st, gram)) return rc;
    return launch_k4(theta, am, av, n_obj, b, it, R, lr, wd, b1, b2, eps, ws, loss_terms, grads_out, n_sm, st);
}

extern "C" int oo_train_k1(const float* theta, int n_obj, const oo_batch* batch, int it, int rays_per_step, float scale,
                           oo_train_ws* ws, int refresh_derived, int n_sm, void* stream) {
    OO_REQUIRE(theta, "oo_train_k1: null theta");
    if (int rc = check_train_args(n_obj, batch, rays_per_step, ws)) return rc;
    OO_REQUIRE((long long)(it + 1) * rays_per_step <= batch->rays_per_obj, "oo_train_k1: step %d exceeds the batch", it);
    return launch_k1(theta, n_obj, batch, it, rays_per_step, scale, ws, n_sm, (cudaStream_t)stream, refresh_derived != 0);
}

extern "C" int oo_train_k4(float* theta, float* adam_m, float* adam_v, int n_obj, const oo_batch* batch, int it,
                           int rays_per_step, float lr, float weight_decay, float beta1, float beta2, float eps,
                           oo_train_ws* ws, float* loss_terms, int n_sm, void* stream) {
    OO_REQUIRE(theta && adam_m && adam_v && batch && ws && ws->slab, "oo_train_k4: null argument");
    return launch_k4(theta, adam_m, adam_v, n_obj, batch, it, rays_per_step, lr, weight_decay, beta1, beta2, eps, ws,
                     loss_terms, nullptr, n_sm, (cudaStream_t)stream);
}

extern "C" int oo_train_step(float* theta, float* adam_m, float* adam_v, int n_obj, const oo_batch* batch, int it,
                             int rays_per_step, float scale, float lr, float weight_decay, float beta1, float beta2,
                             float eps, oo_train_ws* ws, float* loss_terms, int n_sm, void* stream) {
    OO_REQUIRE(theta && adam_m && adam_v, "oo_train_step: null parameter buffers");
    return train_step_impl(theta, adam_m, adam_v, n_obj, batch, it, rays_per_step, scale, lr, weight_decay, beta1, beta2,
                           eps, ws, loss_terms, nullptr, n_sm, (cudaStream_t)stream);
}

extern "C" int oo_train_grads(const float* theta, int n_obj, const oo_batch* batch, int it, int rays_per_step, float scale,
                              oo_train_ws* ws, float* grads_out, float* loss_terms, int n_sm, void* stream) {
    OO_REQUIRE(theta && grads_out, "oo_train_grads: null argument");
    return train_step_impl(const_cast<float*>(theta), nullptr, nullptr, n_obj, batch, it, rays_per_step, scale, 0.f, 0.f,
                           0.f, 0.f, 0.f, ws, loss_terms, grads_out, n_sm, (cudaStream_t)stream);
}

extern "C" int oo_train_frame(float* theta, float* adam_m, float* adam_v, int n_obj, const oo_batch* batch, int iters,
                              int rays_per_step, float scale, float lr, float weight_decay, float beta1, float beta2,
                              float eps, oo_train_ws* ws, float* loss_terms, int n_sm, void* stream) {
    OO_REQUIRE(theta && adam_m && adam_v, "oo_train_frame: null parameter buffers");
    for (int it = 0; it < iters; ++it) {
        float* lt = loss_terms ? loss_terms + (size_t)it * n_obj * 4 : nullptr;
        // the out_clip constants are computed from theta at the first step (theta may have been written from outside
        // between frames); afterwards every update kernel leaves them ready for the next step
        if (int rc = train_step_impl(theta, adam_m, adam_v, n_obj, batch, it, rays_per_step, scale, lr, weight_decay,
                                     beta1, beta2, eps, ws, lt, nullptr, n_sm, (cudaStream_t)stream, it == 0))
            return rc;
    }
    return 0;
}

extern "C" int oo_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, int step, float lr,
                             float weight_decay, float beta1, float beta2, float eps, void* stream) {
    OO_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "oo_adamw_flat: bad argument");
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    k_adamw_flat<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, (float)(1.0 - (double)lr * weight_decay), beta1,
                                                           beta2, eps, (float)(lr / bc1), (float)sqrt(bc2));
    OO_LAUNCH_CHECK();
    return 0;
}
