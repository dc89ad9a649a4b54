// CPU emulator of the fused training tile (TEST INFRASTRUCTURE ONLY -- never linked into the product).
//
// It compiles openobj_b200/csrc/oo_tile.h for the host and runs every phase for tid = 0..255 in
// sequence, with the same static schedule, slot/slab bookkeeping and slab reduction as
// openobj_b200/csrc/oo_train.cu.  The CPU test-suite compares its gradients with the oracle, which checks
// the tile's index algebra (thread -> tile maps, shared-memory map, flush layout) without a GPU.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../openobj_b200/csrc/oo_sched.h"
#include "../../openobj_b200/csrc/oo_tile.h"

using namespace oo;

namespace {

template <int PH, int END, bool PART>
struct Run {
    static void go(float* sm, const TileCtx& c, std::vector<TileAcc>& acc) {
        for (int tid = 0; tid < NTHREADS; ++tid) tile_phase<kTrainOrder[PH], PART>(tid, sm, c, acc[tid]);
        Run<PH + 1, END, PART>::go(sm, c, acc);
    }
};
template <int END, bool PART>
struct Run<END, END, PART> {
    static void go(float*, const TileCtx&, std::vector<TileAcc>&) {}
};

template <bool PART>
void emulate(const float* theta, const float* derived, int n_obj, const float* pcs, const float* z, const float* gt_depth,
             const uint8_t* gt_rgb, const uint8_t* labels, const int32_t* feat_row, const float* feat_table,
             int rays_per_obj, int it, int R, float scale, const int* counts, int flags, int n_sm,
             float* grads_out, float* loss_terms) {
    const Schedule s = build_schedule(n_obj, R, n_sm);
    std::vector<float> slab((size_t)s.n_slots * PSTRIDE, 0.f), slot_loss((size_t)s.n_slots * 4, 0.f);
    std::vector<float> rayrec((size_t)n_obj * R * RAYREC, 0.f);
    std::vector<float> smv(SM_TOTAL + 4);
    for (int cta = 0; cta < s.n_cta; ++cta) {
        // poison shared memory so that reads of never-written locations show up as NaN in the outputs
        for (auto& v : smv) v = __builtin_nanf("");
        float* sm = smv.data();
        const int t_begin = s.cta_tile()[cta], t_end = s.cta_tile()[cta + 1];
        if (t_begin >= t_end) continue;
        int slot = s.cta_slot()[cta];
        for (int tid = 0; tid < NTHREADS; ++tid) zero_pad_rows(tid, sm);
        std::vector<TileAcc> acc(NTHREADS);
        std::vector<float> der(DERIVED, __builtin_nanf(""));
        for (auto& a : acc) acc_zero(a);
        TileCtx c;
        c.flags = flags; c.scale = scale; c.cs = 5.f; c.os = 10.f; c.fs = 5.f; c.feat_table = feat_table;
        int cur_obj = -1;
        for (int t = t_begin; t < t_end; ++t) {
            const int obj = t / s.tiles_per_obj;
            const int r0 = (t - obj * s.tiles_per_obj) * RT;
            if (obj != cur_obj) {
                cur_obj = obj;
                c.theta = theta + (size_t)obj * PSTRIDE;
                c.derived = der.data();
                c.slab = slab.data() + (size_t)slot * PSTRIDE;
                c.inv1 = 1.f / ((float)counts[2 * obj] + 1e-10f);
                c.invs = 1.f / ((float)counts[2 * obj + 1] + 1e-10f);
                if (PART)
                    for (int tid = 0; tid < NTHREADS; ++tid) gram_stage<-1>(tid, sm, c.theta, nullptr);
                for (int tid = 0; tid < NTHREADS; ++tid) stage_weights(tid, sm, c.theta);
                if (PART) {
                    for (int tid = 0; tid < NTHREADS; ++tid) gram_stage<0>(tid, sm, c.theta, der.data());
                    for (int tid = 0; tid < NTHREADS; ++tid) gram_stage<1>(tid, sm, c.theta, der.data());
                    for (int tid = 0; tid < NTHREADS; ++tid) gram_stage<2>(tid, sm, c.theta, der.data());
                    for (int tid = 0; tid < NTHREADS; ++tid) zero_pad_rows(tid, sm);
                    for (int tid = 0; tid < NTHREADS; ++tid) stage_derived(tid, sm, der.data());
                }
            }
            const size_t ray = (size_t)obj * rays_per_obj + (size_t)it * R + r0;
            c.nrays = (R - r0) < RT ? (R - r0) : RT;
            c.npts = c.nrays * S;
            c.pcs = pcs + ray * (S * 3);
            c.z = z + ray * S;
            c.gt_depth = gt_depth + ray;
            c.gt_rgb = gt_rgb + ray * 3;
            c.labels = labels + ray;
            c.feat_row = PART ? feat_row + ray : nullptr;
            c.rayrec = rayrec.data() + ((size_t)obj * R + r0) * RAYREC;
            Run<0, N_TRAIN_PHASES, PART>::go(sm, c, acc);
            const bool last = (t + 1 == t_end) || ((t + 1) / s.tiles_per_obj != obj);
            if (last) {
                float* sl = slot_loss.data() + 4 * slot;
                for (int tid = 0; tid < NTHREADS; ++tid) tile_flush<0, PART>(tid, sm, c.slab, sl, acc[tid]);
                for (int tid = 0; tid < NTHREADS; ++tid) tile_flush<1, PART>(tid, sm, c.slab, sl, acc[tid]);
                for (auto& a : acc) acc_zero(a);
                ++slot;
                cur_obj = -1;
            }
        }
    }
    // slab reduction in the optimiser kernel's order (k_adamw<false>)
    const bool obj_terms = !(flags & 2), op_term = !(flags & 4);
    const bool active[3] = {obj_terms || op_term, obj_terms, obj_terms && PART};
    for (int o = 0; o < n_obj; ++o) {
        const int s0 = s.obj_slot()[o], s1 = s.obj_slot()[o + 1];
        // totals of M, m, beta over the object's slots (k_adamw staging)
        std::vector<float> mt(1057, 0.f);
        for (int q = 0; q < 1057; ++q)
            for (int sl = s0; sl < s1; ++sl) mt[q] += slab[(size_t)sl * PSTRIDE + SLAB_M + q];
        const float* th = theta + (size_t)o * PSTRIDE;
        for (int i = 0; i < PSTRIDE; ++i) {
            float g = 0.f;
            const bool in_w = i >= OFF_OCL_W && i < OFF_OCL_B, in_b = i >= OFF_OCL_B && i < OFF_OCL_B + C;
            if (i < PEND && active[group_of_offset(i)]) {
                if (in_w || in_b) {
                    if (PART) {
                        const int cc = in_w ? (i - OFF_OCL_W) / H : i - OFF_OCL_B, j = in_w ? (i - OFF_OCL_W) % H : 0;
                        for (int r = 0; r < R; ++r) {
                            const float* rec = rayrec.data() + ((size_t)o * R + r) * RAYREC;
                            if (rec[REC_A] == 0.f) continue;
                            const float y = feat_table[(size_t)feat_row[(size_t)o * rays_per_obj + (size_t)it * R + r] * C + cc];
                            g += in_w ? rec[REC_A] * y * rec[REC_S + j] : rec[REC_A] * rec[REC_OPAC] * y;
                        }
                        if (in_w) {
                            for (int k = 0; k < H; ++k) g += th[OFF_OCL_W + cc * H + k] * mt[k * H + j];
                            g += th[OFF_OCL_B + cc] * mt[1024 + j];
                        } else {
                            float t = th[OFF_OCL_B + cc] * mt[1056];
                            for (int k = 0; k < H; ++k) t += th[OFF_OCL_W + cc * H + k] * mt[1024 + k];
                            g += t;
                        }
                    }
                } else {
                    for (int q = s0; q < s1; ++q) g += slab[(size_t)q * PSTRIDE + i];
                }
            }
            grads_out[(size_t)o * PSTRIDE + i] = g;
        }
        for (int k = 0; k < 4; ++k) {
            float sum = 0.f;
            for (int q = s0; q < s1; ++q) sum += slot_loss[4 * q + k];
            loss_terms[4 * o + k] = sum / ((float)counts[2 * o + (k == 2 ? 1 : 0)] + 1e-10f);
        }
    }
}

}  // namespace

extern "C" int emu_train_grads(const float* theta, int n_obj, const float* pcs, const float* z, const float* gt_depth,
                               const uint8_t* gt_rgb, const uint8_t* labels, const int32_t* feat_row,
                               const float* feat_table, int rays_per_obj, int it, int R, float scale,
                               const int* counts, int flags, int n_sm, float* grads_out, float* loss_terms) {
    std::vector<float> wt(1);
    if (feat_row)
        emulate<true>(theta, wt.data(), n_obj, pcs, z, gt_depth, gt_rgb, labels, feat_row, feat_table, rays_per_obj, it, R,
                      scale, counts, flags, n_sm, grads_out, loss_terms);
    else
        emulate<false>(theta, wt.data(), n_obj, pcs, z, gt_depth, gt_rgb, labels, nullptr, nullptr, rays_per_obj, it, R,
                       scale, counts, flags, n_sm, grads_out, loss_terms);
    return 0;
}

extern "C" int emu_schedule(int n_obj, int R, int n_sm, int* out, int cap) {
    const Schedule s = build_schedule(n_obj, R, n_sm);
    if ((int)s.data.size() + 3 > cap) return -1;
    out[0] = s.n_cta; out[1] = s.n_slots; out[2] = s.tiles_per_obj;
    memcpy(out + 3, s.data.data(), s.data.size() * sizeof(int));
    return (int)s.data.size() + 3;
}
