// Background model GEMMs on the 5th-generation tensor cores (a17: objnerf/train.py:300-315,447-463 -- the hidden-128
// OccupancyMap's nine layers forward, backward-data and backward-weight, M = 16 800 points, K <= 215).
//
//   C(i,j) (+)= epilogue( mult * sum_c A(i,c) B(j,c) + bias[j] )           (GemmOp, oo_gemm.h)
//
// tcgen05.mma kind::tf32, M = 128 rows (the 128 lanes of tensor memory), N = up to 256 columns per CTA, fp32 accumulators in
// TENSOR MEMORY.  fp32-level accuracy from TF32 inputs by three-term error compensation exactly as in the fused object tile
// (oo_tile.h) and the mma.sync engine it replaces: x = hi + lo, a.b ~ lo_a hi_b + hi_a lo_b + hi_a hi_b -> three MMAs per
// 8-wide k-step into the same accumulator, small terms first.
//
// tcgen05 reads both operands from shared memory and ignores the low 13 mantissa bits, so every element is split ONCE, when
// its k-chunk (16 contraction indices) is staged into the canonical K-major core-matrix layout the shared-memory descriptor
// addresses ([k / 4][row][4 floats]: 8 rows x 16 bytes per core matrix, SBO = 128 B between 8-row groups, LBO = rows x 16 B +
// a 16 / 32 B pad between 16-byte k-chunks; the pad keeps a quarter-warp's 16-byte accesses on eight bank groups).  An operand
// arrives in one of two ways (neither uses 4-byte cp.async, which the mma.sync engine stages everything with):
//   V  contraction-contiguous with 16-byte aligned rows (activations, gradients, packed weights): 16-byte cp.async straight
//      into the `hi` block through a 4-stage ring (2 chunks in flight ahead of the one being split); the thread that copied a
//      piece later masks hi in place and writes lo beside it -- no block barrier in the main loop;
//   R  row-contiguous (the transposed operands of the weight gradients, the weights of backward-data): 128-bit loads along
//      the rows, a 4 x 4 register transpose, hi / lo stored directly; the next chunk's loads are in flight meanwhile.
// Warp-specialised: 16 producer warps + one MMA warp.  Producers arrive (one arrival per warp) on the stage's `full` mbarrier;
// the MMA thread waits, issues ONE generic->async proxy fence and the 6 MMAs of the chunk, and tcgen05.commit -> the stage's
// `empty` mbarrier tells the producers when the stage may be refilled.  (A proxy fence in the producers compiles to
// MEMBAR.ALL.CTA and waits for their in-flight copies: measured, it made every chunk a full memory round trip.)
// Epilogue: accumulator rows come back with tcgen05.ld (thread = row) into a shared tile, then leave row by row as contiguous
// 128-bit stores with bias / x10 / ReLU / sigmoid / ReLU-mask / accumulate applied (instantiated per combination), or as raw
// split partials for the weight gradients (k_gemm_reduce_batch in oo_bg.cu finishes them in a fixed order).
// Measured on B200 (tools/bg_probe.py, same run): background step 0.815 ms (mma.sync engine) -> 0.635 ms.
#include "oo_gemm.h"

#include <stdint.h>
#include <stdlib.h>

#include "oo_common.cuh"

namespace oo {
namespace {

constexpr int TG_THREADS = 512, TG_STAGES = 4, TG_Q = TG_KC / 4;      // 4 sixteen-byte k-chunks per stage, 4 stages
constexpr int TG_PRE = TG_STAGES - 2;                                  // V chunks in flight ahead of the one being split
// (measured: a deeper ring -- 6 stages fit when N <= 128 -- is slower: the loop is not load-latency bound, and run-time stage
// arithmetic in the loop costs more than it hides)
constexpr int TG_BLOCK = TG_THREADS + 32;                              // 16 producer warps + the MMA-issuing warp
static_assert(TG_KC == 16 && TG_BI == 128 && TG_THREADS == 512, "thread maps below");

__host__ __device__ constexpr int stage_floats(int nj) { return 2 * TG_Q * (TG_BI * 4 + 8) + 2 * TG_Q * (nj * 4 + 8); }   // upper bound (pads)

// instruction descriptor of tcgen05.mma kind::tf32: D fp32 (bits 4-5 = 1), A and B TF32 (bits 7-9 / 10-12 = 2), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TG_BI >> 4) << 24);
}
// K-major, no swizzle: start >> 4 | LBO (next 16-byte k-chunk) >> 4 at bit 16 | SBO (128 B, next 8 rows) >> 4 at bit 32 |
// descriptor version 1 (sm_100) at bit 46
__device__ __forceinline__ uint64_t smem_desc(uint32_t byte_addr, uint32_t lbo_bytes) {
    return (uint64_t)((byte_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t id, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(id), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a descriptor mistake must end as a launch failure, not as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 2000000000LL) __trap();          // ~1 s
    }
}
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}
// in place: hi = x with the low 13 mantissa bits cleared (what the tensor core would read anyway), lo = x - hi exactly
__device__ __forceinline__ float4 split4(float4& x) {
    float4 l;
    split(x.x, x.x, l.x); split(x.y, x.y, l.y); split(x.z, x.z, l.z); split(x.w, x.w, l.w);
    return l;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(float* dst, const float* src, int bytes) {       // src-size < 16 zero-fills the rest
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- R operand: thread (rq = tid / 4, kq = tid % 4) owns rows 4 rq .. 4 rq + 3 and contraction indices c0 + 4 kq .. + 3:
// four 128-bit loads along the rows (a quarter-warp reads 2 x 64 contiguous bytes of each of 4 contraction indices)
struct RFrag { float4 v[4]; };
__device__ __forceinline__ void load_r(RFrag& f, bool mine, const float* __restrict__ base, long long scol, int row, int r_end,
                                       int c, int c_end) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        f.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mine && row < r_end && c + u < c_end) {
            float4 x = __ldg(reinterpret_cast<const float4*>(base + (long long)(c + u) * scol + row));
            if (row + 1 >= r_end) x.y = 0.f;          // the allocation's row pad is in bounds (checked on the host) but not data
            if (row + 2 >= r_end) x.z = 0.f;
            if (row + 3 >= r_end) x.w = 0.f;
            f.v[u] = x;
        }
    }
}
// transpose to [row][4 k], split, store hi / lo; bank group of a store = (kq + 4 rq + u) % 8: conflict-free per quarter-warp
__device__ __forceinline__ void store_r(const RFrag& f, float* __restrict__ hi_blk, float* __restrict__ lo_blk, int lbo_f, int rq,
                                        int kq, float* rsum) {
    const float rows[4][4] = {{f.v[0].x, f.v[1].x, f.v[2].x, f.v[3].x}, {f.v[0].y, f.v[1].y, f.v[2].y, f.v[3].y},
                              {f.v[0].z, f.v[1].z, f.v[2].z, f.v[3].z}, {f.v[0].w, f.v[1].w, f.v[2].w, f.v[3].w}};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        float4 h = make_float4(rows[u][0], rows[u][1], rows[u][2], rows[u][3]);
        if (rsum != nullptr) rsum[u] += (h.x + h.y) + (h.z + h.w);
        const float4 l = split4(h);
        const int e = kq * lbo_f + (4 * rq + u) * 4;
        *reinterpret_cast<float4*>(hi_blk + e) = h;
        *reinterpret_cast<float4*>(lo_blk + e) = l;
    }
}

__device__ __forceinline__ float epilogue1(const GemmOp& g, float acc, float bias, float mk, float old) {
    float v = (acc * g.mult + bias) * g.post;
    if (g.act == 1) v = fmaxf(v, 0.f);
    else if (g.act == 2) v = 1.f / (1.f + expf(-v));
    if (!(mk > 0.f)) v = 0.f;
    return g.accumulate ? old + v : v;
}

// Epilogue, part 2, vector path: full column quads as 128-bit accesses (lane = quad, up to two per lane), RB rows per batch
// with every load of the batch issued before the first store (mask and old values come from global memory: one round trip per
// batch, not per row); the last n_cols % 4 columns of a row go one by one.  ACT 0 none / 1 ReLU / 2 sigmoid / 3 = read g.act.
template <int ACT, bool MASK, bool ACC>
__device__ __forceinline__ float ep_val(const GemmOp& g, float acc, float bias, float mk, float old, float mult, float post) {
    float v = (acc * mult + bias) * post;
    const int act = ACT == 3 ? g.act : ACT;
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 2) v = 1.f / (1.f + expf(-v));
    if (MASK && !(mk > 0.f)) v = 0.f;
    return ACC ? old + v : v;
}
template <int ACT, bool MASK, bool ACC>
__device__ __noinline__ void ep2_vec(const GemmOp& g, const float* __restrict__ sm, int ldt, int warp, int lane, int i0, int j0,
                                     int n_rows, int n_cols, int dbg) {
    const bool has_mask = MASK && g.mask != nullptr, acc_on = ACC && g.accumulate;
    const float mult = g.mult, post = g.post;
    const int nq = n_cols >> 2, sci = (int)g.sci, smi = (int)g.smi, mask_cols = g.mask_cols;
    float* __restrict__ Cb = g.C + (long long)i0 * g.sci + j0;
    const float* __restrict__ Mb = has_mask ? g.mask + (long long)i0 * g.smi + j0 : nullptr;
    float4 b4[2];
#pragma unroll
    for (int qq = 0; qq < 2; ++qq) {
        const int j = j0 + 4 * (lane + 32 * qq);
        b4[qq] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.bias != nullptr && lane + 32 * qq < nq) b4[qq] = make_float4(__ldg(g.bias + j), __ldg(g.bias + j + 1), __ldg(g.bias + j + 2), __ldg(g.bias + j + 3));
    }
    constexpr int NW = TG_THREADS / 32, RB = 2;
    const int nqq = nq > 32 ? 2 : 1;
    for (int rb = warp; rb < n_rows; rb += NW * RB) {
        float4 a4[RB][2], m4[RB][2], o4[RB][2];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int r = rb + NW * u;
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
                const int c = 4 * (lane + 32 * qq);
                m4[u][qq] = make_float4(1.f, 1.f, 1.f, 1.f);
                o4[u][qq] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qq < nqq && r < n_rows && lane + 32 * qq < nq) {
                    a4[u][qq] = *reinterpret_cast<const float4*>(sm + r * ldt + c);
                    if (has_mask && j0 + c < mask_cols) m4[u][qq] = *reinterpret_cast<const float4*>(Mb + r * smi + c);
                    if (acc_on) o4[u][qq] = *reinterpret_cast<const float4*>(Cb + r * sci + c);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int r = rb + NW * u;
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
                const int c = 4 * (lane + 32 * qq);
                if (qq < nqq && r < n_rows && lane + 32 * qq < nq) {
                    float4 o;
                    o.x = ep_val<ACT, MASK, ACC>(g, a4[u][qq].x, b4[qq].x, m4[u][qq].x, o4[u][qq].x, mult, post);
                    o.y = ep_val<ACT, MASK, ACC>(g, a4[u][qq].y, b4[qq].y, m4[u][qq].y, o4[u][qq].y, mult, post);
                    o.z = ep_val<ACT, MASK, ACC>(g, a4[u][qq].z, b4[qq].z, m4[u][qq].z, o4[u][qq].z, mult, post);
                    o.w = ep_val<ACT, MASK, ACC>(g, a4[u][qq].w, b4[qq].w, m4[u][qq].w, o4[u][qq].w, mult, post);
                    if (dbg != 3) *reinterpret_cast<float4*>(Cb + r * sci + c) = o;
                }
            }
            if (r < n_rows && lane < (n_cols & 3)) {
                const int c = 4 * nq + lane, j = j0 + c;
                float* dst = Cb + r * sci + c;
                const float mk = (has_mask && j < mask_cols) ? Mb[r * smi + c] : 1.f;
                *dst = ep_val<ACT, MASK, ACC>(g, sm[r * ldt + c], g.bias != nullptr ? __ldg(g.bias + j) : 0.f, mk, acc_on ? *dst : 0.f, mult, post);
            }
        }
    }
}

template <bool AV, bool BV>
__global__ void __launch_bounds__(TG_BLOCK, 1) k_gemm_tc(const GemmOp g, int tm_cols, int cl, int dbg) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint32_t tm_base_s;
    __shared__ __align__(8) uint64_t bars[2 * TG_STAGES];             // full[s] (producers -> MMA warp), empty[s] (MMAs done)
    __shared__ float rs_sm[TG_BI];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i0 = blockIdx.x * TG_BI, j0 = blockIdx.y * TG_BJ;
    const int nj = min(TG_BJ, (g.J - j0 + 15) & ~15);                 // MMA N of this CTA (multiple of 16)
    const int c_begin = blockIdx.z * g.chunk, c_end = min(g.K, c_begin + g.chunk);
    const int n_chunks = c_end > c_begin ? (c_end - c_begin + TG_KC - 1) / TG_KC : 0;
    // k-chunk strides: rows x 16 B plus a pad that keeps a quarter-warp's 16-byte accesses on eight bank groups
    // (V pieces: (row, q) = (e / 4, e % 4) -> 32 B; R pieces: (rq, kq) -> 16 B)
    const int lboa = TG_BI * 4 + (AV ? 8 : 4), lbob = nj * 4 + (BV ? 8 : 4);
    const int stg_f = 2 * TG_Q * lboa + 2 * TG_Q * lbob;

    const uint32_t full0 = (uint32_t)__cvta_generic_to_shared(&bars[0]), empty0 = full0 + 8u * TG_STAGES;
    if (tid == 0) {
        for (int s = 0; s < TG_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8u * s), "r"(TG_THREADS / 32) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty0 + 8u * s) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tm_base_s)), "r"((uint32_t)tm_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tm_base_s;
    const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sm);
    const bool row_sums = g.ones_out != nullptr && blockIdx.y == 0;
    // R fragments: A items (rq, kq) on threads 0..127, B items on threads 128..128 + nj - 1 (disjoint: staged in parallel)
    const int kq = tid & 3, rq = tid >> 2, rqb = (tid - TG_BI) >> 2;
    const bool a_mine = tid < TG_BI, b_mine = tid >= TG_BI && rqb < nj / 4;
    float rs4[4] = {0.f, 0.f, 0.f, 0.f};

    pdl_wait();
    if (warp == TG_THREADS / 32) {
        // ================= MMA issuer: one thread; waits for a full stage, multiplies it, commit -> the stage's `empty`
        if (lane == 0 && dbg != 1) {
            const uint32_t lba = (uint32_t)lboa * 4u, lbb = (uint32_t)lbob * 4u, id = idesc(nj);
            for (int n = 0; n < n_chunks; ++n) {
                const int st = n % TG_STAGES, c0 = c_begin + n * TG_KC;
                mbar_wait(full0 + 8u * st, (uint32_t)((n / TG_STAGES) & 1));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> tensor-core reads
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ah = sm_base + (uint32_t)(st * stg_f) * 4u, al = ah + (uint32_t)(TG_Q * lboa) * 4u;
                const uint32_t bh = al + (uint32_t)(TG_Q * lboa) * 4u, bl = bh + (uint32_t)(TG_Q * lbob) * 4u;
                const int ksteps = min(TG_KC / 8, (c_end - c0 + 7) / 8);
                for (int s = 0; s < ksteps; ++s) {
                    const uint64_t dah = smem_desc(ah + 2u * s * lba, lba), dal = smem_desc(al + 2u * s * lba, lba);
                    const uint64_t dbh = smem_desc(bh + 2u * s * lbb, lbb), dbl = smem_desc(bl + 2u * s * lbb, lbb);
                    if (dbg != 2) {
                        umma_ss(tm, dal, dbh, id, (n == 0 && s == 0) ? 0u : 1u);      // small terms first
                        umma_ss(tm, dah, dbl, id, 1u);
                    }
                    umma_ss(tm, dah, dbh, id, (dbg == 2 && n == 0 && s == 0) ? 0u : 1u);
                }
                commit(empty0 + 8u * st);
            }
        }
    } else {
        // ================= producers (512 threads): every thread stages, splits and publishes ITS OWN pieces, so the main
        // loop has no block barrier: cp.async.wait_group for its copies, then one arrive per warp on the stage's `full` barrier
        // V pieces of this thread: A (row, q) = (tid / 4, tid % 4); B (row, q) = (e / 4, e % 4), e = tid + 512 u (u < nj / 128)
        constexpr int NPA = TG_BI * TG_Q / TG_THREADS, NPB = TG_BJ * TG_Q / TG_THREADS;      // 1, 2
        const float* a_src[NPA]; int a_off[NPA]; bool a_ok[NPA];
        const float* b_src[NPB]; int b_off[NPB]; bool b_ok[NPB], b_has[NPB];
        if (AV) {
#pragma unroll
            for (int u = 0; u < NPA; ++u) {
                const int e = tid + TG_THREADS * u, r = e >> 2, q = e & 3;
                a_ok[u] = i0 + r < g.I;
                a_src[u] = g.A + (long long)(a_ok[u] ? i0 + r : 0) * g.sai + 4 * q;
                a_off[u] = q * lboa + r * 4;
            }
        }
        if (BV) {
#pragma unroll
            for (int u = 0; u < NPB; ++u) {
                const int e = tid + TG_THREADS * u, r = e >> 2, q = e & 3;
                b_has[u] = r < nj;
                b_ok[u] = b_has[u] && j0 + r < g.J;
                b_src[u] = g.B + (long long)(b_ok[u] ? j0 + r : 0) * g.sbj + 4 * q;
                b_off[u] = q * lbob + r * 4;
            }
        }
        auto issue = [&](int n) {                  // V operands of chunk n; one commit group per chunk (empty past the end)
            if (n < n_chunks) {
                float* a_hi = sm + (n % TG_STAGES) * stg_f;
                float* b_hi = a_hi + 2 * TG_Q * lboa;
                const int c0 = c_begin + n * TG_KC;
                if (AV) {
#pragma unroll
                    for (int u = 0; u < NPA; ++u) {
                        int bytes = a_ok[u] ? 4 * (c_end - c0 - 4 * ((tid + TG_THREADS * u) & 3)) : 0;
                        bytes = bytes < 0 ? 0 : bytes > 16 ? 16 : bytes;
                        cp_async16(a_hi + a_off[u], bytes > 0 ? a_src[u] + c0 : g.A, bytes);
                    }
                }
                if (BV) {
#pragma unroll
                    for (int u = 0; u < NPB; ++u) {
                        if (!b_has[u]) continue;
                        int bytes = b_ok[u] ? 4 * (c_end - c0 - 4 * ((tid + TG_THREADS * u) & 3)) : 0;
                        bytes = bytes < 0 ? 0 : bytes > 16 ? 16 : bytes;
                        cp_async16(b_hi + b_off[u], bytes > 0 ? b_src[u] + c0 : g.B, bytes);
                    }
                }
            }
            cp_async_commit();
        };
        RFrag fa, fb;
#pragma unroll 1
        for (int s = 0; s < TG_PRE; ++s) issue(s);
        if (!AV) load_r(fa, a_mine, g.A, g.sac, i0 + 4 * rq, g.I, c_begin + 4 * kq, c_end);
        if (!BV) load_r(fb, b_mine, g.B, g.sbc, j0 + 4 * rqb, g.J, c_begin + 4 * kq, c_end);
#pragma unroll 1
        for (int n = 0; n < n_chunks; ++n) {
            const int st = n % TG_STAGES, c0 = c_begin + n * TG_KC;
            float* a_hi = sm + st * stg_f;
            float* a_lo = a_hi + TG_Q * lboa;
            float* b_hi = a_lo + TG_Q * lboa;
            float* b_lo = b_hi + TG_Q * lbob;
            // ---- R operands: this chunk from registers (the stage is free: its previous MMAs were awaited at iteration
            // n - STAGES + 1, below), then the next chunk's loads start their trip
            if (!AV) {
                if (a_mine) store_r(fa, a_hi, a_lo, lboa, rq, kq, row_sums ? rs4 : nullptr);
                load_r(fa, a_mine && n + 1 < n_chunks, g.A, g.sac, i0 + 4 * rq, g.I, c0 + TG_KC + 4 * kq, c_end);
            }
            if (!BV) {
                if (b_mine) store_r(fb, b_hi, b_lo, lbob, rqb, kq, nullptr);
                load_r(fb, b_mine && n + 1 < n_chunks, g.B, g.sbc, j0 + 4 * rqb, g.J, c0 + TG_KC + 4 * kq, c_end);
            }
            // ---- V operands: this thread's copies of chunk n have landed; hi masked in place, lo = x - hi beside it
            if (AV || BV) {
                cp_async_wait<TG_PRE - 1>();
                float4 xa[NPA], xb[NPB];
                if (AV) {
#pragma unroll
                    for (int u = 0; u < NPA; ++u) xa[u] = *reinterpret_cast<const float4*>(a_hi + a_off[u]);
                }
                if (BV) {
#pragma unroll
                    for (int u = 0; u < NPB; ++u)
                        if (b_has[u]) xb[u] = *reinterpret_cast<const float4*>(b_hi + b_off[u]);
                }
                if (AV) {
#pragma unroll
                    for (int u = 0; u < NPA; ++u) {
                        if (row_sums) rs4[u] += (xa[u].x + xa[u].y) + (xa[u].z + xa[u].w);
                        const float4 l = split4(xa[u]);
                        if (dbg != 6) *reinterpret_cast<float4*>(a_hi + a_off[u]) = xa[u];
                        *reinterpret_cast<float4*>(a_lo + a_off[u]) = l;
                    }
                }
                if (BV) {
#pragma unroll
                    for (int u = 0; u < NPB; ++u) {
                        if (!b_has[u]) continue;
                        const float4 l = split4(xb[u]);
                        if (dbg != 6) *reinterpret_cast<float4*>(b_hi + b_off[u]) = xb[u];
                        *reinterpret_cast<float4*>(b_lo + b_off[u]) = l;
                    }
                }
            }
            // publish: the arrive (release) orders this thread's shared-memory stores before the MMA warp's wait (acquire);
            // the generic -> async proxy fence is issued ONCE, by the MMA thread, after that wait.  (A fence here compiles to
            // MEMBAR.ALL.CTA, which also waits for this thread's in-flight cp.async / prefetch loads: measured, it turned
            // every iteration into a full memory round trip.)
            // one arrival per warp (512 arrivals on one barrier are 512 serialised shared-memory atomics per chunk)
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + 8u * st) : "memory");
            // chunk n + PRE goes into the stage chunk n + PRE - STAGES was multiplied from: those MMAs were issued two
            // iterations ago (use (n + PRE - STAGES) / STAGES of that stage's `empty` barrier), so this wait is normally free
            if (dbg != 1 && n + TG_PRE >= TG_STAGES && n + TG_PRE < n_chunks)
                mbar_wait(empty0 + 8u * ((n + TG_PRE) % TG_STAGES), (uint32_t)(((n + TG_PRE - TG_STAGES) / TG_STAGES) & 1));
            issue(n + TG_PRE);
        }
    }
    pdl_release();
    if (n_chunks > 0 && dbg != 1) {
        const int last = n_chunks - 1;
        mbar_wait(empty0 + 8u * (last % TG_STAGES), (uint32_t)((last / TG_STAGES) & 1));     // a commit covers every MMA issued before it
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- row sums of A (the bias gradients) -> rs_sm[row]
    const int je = g.J + (g.ones_out != nullptr ? 1 : 0);
    if (row_sums) {
        if (warp < TG_THREADS / 32) {
            if (AV) {
                // piece u of thread tid is (row (tid + 256 u) / 4, q tid % 4): the four q threads of a row are neighbouring lanes
#pragma unroll
                for (int u = 0; u < TG_BI * TG_Q / TG_THREADS; ++u) {
                    rs4[u] += __shfl_xor_sync(0xffffffffu, rs4[u], 1);
                    rs4[u] += __shfl_xor_sync(0xffffffffu, rs4[u], 2);
                    if ((tid & 3) == 0) rs_sm[(tid + TG_THREADS * u) >> 2] = rs4[u];
                }
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {             // the four kq threads of a row quad are neighbouring lanes
                    rs4[u] += __shfl_xor_sync(0xffffffffu, rs4[u], 1);
                    rs4[u] += __shfl_xor_sync(0xffffffffu, rs4[u], 2);
                }
                if (a_mine && kq == 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) rs_sm[4 * rq + u] = rs4[u];
                }
            }
        }
        __syncthreads();
        if (tid < TG_BI && i0 + tid < g.I) {
            const int i = i0 + tid;
            if (g.split > 1) { if (cl <= 1) g.part[((size_t)blockIdx.z * g.I + i) * je + g.J] = rs_sm[tid]; }
            else g.ones_out[i] = epilogue1(g, rs_sm[tid], 0.f, 1.f, g.accumulate ? g.ones_out[i] : 0.f);
        }
    }

    // ---- epilogue, part 1: accumulator rows -> shared tile T [128][nj + 4] (the stages are free: every MMA has completed).
    // Thread = row (lane quarter warp % 4); warp w takes the 16-column groups w / 4, w / 4 + 4, ...
    const int ldt = nj + 4;                        // ldt % 32 == 4 or 20: eight float4 stores of neighbouring rows, eight bank groups
    if (warp < TG_THREADS / 32) {
        const int row = 32 * (warp & 3) + lane;
        const uint32_t tm_lane = tm + ((uint32_t)(32 * (warp & 3)) << 16);
        for (int c0 = 16 * (warp >> 2); c0 < nj; c0 += 16 * (TG_THREADS / 128)) {
            uint32_t r[16];
            if (n_chunks > 0 && dbg != 1) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(tm_lane + (uint32_t)c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int u = 0; u < 16; ++u) r[u] = 0u;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(sm + row * ldt + c0 + 4 * q) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"((uint32_t)tm_cols) : "memory");

    // ---- split partials, cluster form: the `cl` CTAs of a thread-block cluster hold consecutive contraction chunks of the SAME
    // output tile; they add their tiles through distributed shared memory in a fixed order (rank 0, 1, ..) before anything goes
    // to global memory -- CTA rank q sums rows [128 q / cl, 128 (q + 1) / cl) -- so the fixed-order reduction launch that
    // follows reads cl x fewer partials (132 -> 33 per weight gradient)
    if (g.split > 1 && cl > 1) {
        uint32_t rank;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        const int rows_per = TG_BI / cl, r_lo = (int)rank * rows_per;
        const int n_cols_c = min(nj, g.J - j0);
        const uint32_t t_base = sm_base, rs_base = (uint32_t)__cvta_generic_to_shared(&rs_sm[0]);
        const size_t zc = blockIdx.z / cl;
        // items = (row, 4-column group): the cl remote 128-bit loads of an item are issued together, summed in rank order
        const int nq4 = (n_cols_c + 3) >> 2, n_items = rows_per * nq4;
        for (int e = tid; e < n_items; e += TG_BLOCK) {
            const int r = r_lo + e / nq4, c = 4 * (e % nq4);
            if (i0 + r >= g.I) continue;
            float4 x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q < cl) {
                    uint32_t ra;
                    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(t_base + (uint32_t)(r * ldt + c) * 4u), "r"(q));
                    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[q].x), "=f"(x[q].y), "=f"(x[q].z), "=f"(x[q].w) : "r"(ra));
                }
            }
            float4 v = x[0];
#pragma unroll
            for (int q = 1; q < 8; ++q)
                if (q < cl) { v.x += x[q].x; v.y += x[q].y; v.z += x[q].z; v.w += x[q].w; }
            float* dst = g.part + (zc * g.I + i0 + r) * je + j0 + c;
            dst[0] = v.x;
            if (c + 1 < n_cols_c) dst[1] = v.y;
            if (c + 2 < n_cols_c) dst[2] = v.z;
            if (c + 3 < n_cols_c) dst[3] = v.w;
        }
        if (row_sums && tid < rows_per && i0 + r_lo + tid < g.I) {
            const int r = r_lo + tid;
            float v = 0.f;
            for (int q = 0; q < cl; ++q) {
                uint32_t ra;
                float x;
                asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(rs_base + (uint32_t)r * 4u), "r"(q));
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(x) : "r"(ra));
                v += x;
            }
            g.part[(zc * g.I + i0 + r) * je + g.J] = v;
        }
        // nobody leaves while its tile may still be read by the others
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        return;
    }

    // ---- epilogue, part 2: a warp per row, lanes along the columns (contiguous in C when scj == 1)
    if (warp >= TG_THREADS / 32) return;
    const int n_rows = min(TG_BI, g.I - i0), n_cols = min(nj, g.J - j0);
    if (g.split > 1) {
        for (int r = warp; r < n_rows; r += TG_THREADS / 32) {
            float* dst = g.part + ((size_t)blockIdx.z * g.I + i0 + r) * je + j0;
            for (int c = lane; c < n_cols; c += 32) dst[c] = sm[r * ldt + c];
        }
    }
    const bool c_vec = g.scj == 1 && (g.sci & 3) == 0 && ((size_t)g.C & 15) == 0;
    const bool m_vec = g.mask == nullptr || (g.smj == 1 && (g.smi & 3) == 0 && ((size_t)g.mask & 15) == 0 && (g.mask_cols & 3) == 0);
    if (g.split > 1) {
    } else if (c_vec && m_vec) {
        // uniform switches resolved once: the row loop below is instantiated per (activation, mask, accumulate)
        const int key = g.act * 4 + (g.mask != nullptr ? 2 : 0) + (g.accumulate ? 1 : 0);
        switch (key) {
            case 0: ep2_vec<0, false, false>(g, sm, ldt, warp, lane, i0, j0, n_rows, n_cols, dbg); break;
            case 1: ep2_vec<0, false, true>(g, sm, ldt, warp, lane, i0, j0, n_rows, n_cols, dbg); break;
            case 2: ep2_vec<0, true, false>(g, sm, ldt, warp, lane, i0, j0, n_rows, n_cols, dbg); break;
            case 3: ep2_vec<0, true, true>(g, sm, ldt, warp, lane, i0, j0, n_rows, n_cols, dbg); break;
            case 4: ep2_vec<1, false, false>(g, sm, ldt, warp, lane, i0, j0, n_rows, n_cols, dbg); break;
            case 8: ep2_vec<2, false, false>(g, sm, ldt, warp, lane, i0, j0, n_rows, n_cols, dbg); break;
            default: ep2_vec<3, true, true>(g, sm, ldt, warp, lane, i0, j0, n_rows, n_cols, dbg); break;     // generic
        }
    } else {
        for (int r = warp; r < n_rows; r += TG_THREADS / 32) {
            const int i = i0 + r;
            for (int c = lane; c < n_cols; c += 32) {
                const int j = j0 + c;
                float* dst = g.C + (long long)i * g.sci + (long long)j * g.scj;
                const float mk = (g.mask != nullptr && j < g.mask_cols) ? g.mask[(long long)i * g.smi + (long long)j * g.smj] : 1.f;
                const float old = g.accumulate ? *dst : 0.f;
                *dst = epilogue1(g, sm[r * ldt + c], g.bias != nullptr ? __ldg(g.bias + j) : 0.f, mk, old);
            }
        }
    }
}

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_z, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cluster_z > 1) {                        // thread-block cluster along the split dimension (distributed shared memory)
        at[1].id = cudaLaunchAttributeClusterDimension;
        at[1].val.clusterDim.x = 1; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = (unsigned)cluster_z;
        cfg.numAttrs = 2;
    }
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

bool aligned16(const void* p) { return ((size_t)p & 15) == 0; }
bool op_v(const float* p, long long srow, long long scol) { return scol == 1 && (srow & 3) == 0 && aligned16(p); }
bool op_r(const float* p, long long srow, long long scol, int rows) {
    return srow == 1 && (scol & 3) == 0 && aligned16(p) && (((long long)rows + 3) & ~3LL) <= scol;
}

}  // namespace

// Which way each operand can be staged (see the header of this file); false = this GEMM goes to the mma.sync engine.
// V: contraction-contiguous, rows 16-byte aligned.  R: row-contiguous, contraction stride a multiple of 4 floats that covers
// the row quads the 128-bit loads touch.
bool gemm_tc_supported(const GemmOp& g) {
    return (op_v(g.A, g.sai, g.sac) || op_r(g.A, g.sai, g.sac, g.I)) && (op_v(g.B, g.sbj, g.sbc) || op_r(g.B, g.sbj, g.sbc, g.J));
}

int run_gemm_tc(const GemmOp& g, cudaStream_t st, int cluster_z) {
    OO_REQUIRE(g.chunk > 0 && g.chunk % TG_KC == 0, "oo_bg gemm (tcgen05): chunk must be a multiple of %d", TG_KC);
    OO_REQUIRE(gemm_tc_supported(g), "oo_bg gemm (tcgen05): operand layout not supported");
    const int nj_max = g.J >= TG_BJ ? TG_BJ : (g.J + 15) & ~15;
    int tm_cols = 32;
    while (tm_cols < nj_max) tm_cols *= 2;
    size_t smem = (size_t)TG_STAGES * stage_floats(nj_max) * sizeof(float);
    const size_t tile = (size_t)TG_BI * (nj_max + 4) * sizeof(float);
    if (tile > smem) smem = tile;
    static PerDevice attr_set;
    if (!attr_set.cur()) {
        const int mx = (int)((size_t)TG_STAGES * stage_floats(TG_BJ) * sizeof(float));
        // one shared-memory carve-out for every GEMM of the chain: no SM reconfiguration between consecutive launches
        OO_CUDA(cudaFuncSetAttribute(k_gemm_tc<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        OO_CUDA(cudaFuncSetAttribute(k_gemm_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        OO_CUDA(cudaFuncSetAttribute(k_gemm_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        OO_CUDA(cudaFuncSetAttribute(k_gemm_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        OO_CUDA(cudaFuncSetAttribute(k_gemm_tc<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        OO_CUDA(cudaFuncSetAttribute(k_gemm_tc<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        OO_CUDA(cudaFuncSetAttribute(k_gemm_tc<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        OO_CUDA(cudaFuncSetAttribute(k_gemm_tc<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set.cur() = 1;
    }
    static const int dbg = []() { const char* e = getenv("OO_GEMM_TC_DEBUG"); return e ? atoi(e) : 0; }();   // profiling experiments
    OO_REQUIRE(cluster_z >= 1 && cluster_z <= 8 && g.split % cluster_z == 0 && TG_BI % cluster_z == 0, "oo_bg gemm (tcgen05): split %d is not a multiple of the cluster size %d", g.split, cluster_z);
    const dim3 grid((g.I + TG_BI - 1) / TG_BI, (g.J + TG_BJ - 1) / TG_BJ, g.split);
    const bool av = op_v(g.A, g.sai, g.sac), bv = op_v(g.B, g.sbj, g.sbc);
    if (av && bv) OO_CUDA(launch_pdl(k_gemm_tc<true, true>, grid, dim3(TG_BLOCK), smem, st, cluster_z, g, tm_cols, cluster_z, dbg));
    else if (av) OO_CUDA(launch_pdl(k_gemm_tc<true, false>, grid, dim3(TG_BLOCK), smem, st, cluster_z, g, tm_cols, cluster_z, dbg));
    else if (bv) OO_CUDA(launch_pdl(k_gemm_tc<false, true>, grid, dim3(TG_BLOCK), smem, st, cluster_z, g, tm_cols, cluster_z, dbg));
    else OO_CUDA(launch_pdl(k_gemm_tc<false, false>, grid, dim3(TG_BLOCK), smem, st, cluster_z, g, tm_cols, cluster_z, dbg));
    OO_LAUNCH_CHECK();
    return 0;
}

}  // namespace oo
