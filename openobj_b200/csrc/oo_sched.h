// Static work schedule of the fused training step: which tiles (10-ray groups of one object) each
// persistent CTA processes and which gradient slot every (CTA, object) run writes.  Host only; shared by
// the C-ABI (oo_train.cu) and the CPU tile emulator used in the tests.
#pragma once
#include <vector>

#include "oo_layout.h"

namespace oo {

struct Schedule {
    int n_cta = 0, n_slots = 0, tiles_per_obj = 0, n_obj = 0;
    // data = [cta_tile_begin (n_cta+1)] [cta_slot_begin (n_cta)] [obj_slot_begin (n_obj+1)]
    std::vector<int> data;
    const int* cta_tile() const { return data.data(); }
    const int* cta_slot() const { return data.data() + n_cta + 1; }
    const int* obj_slot() const { return data.data() + 2 * n_cta + 1; }
};

inline int tiles_per_object(int rays_per_step) { return (rays_per_step + RT - 1) / RT; }

// Tiles are numbered object-major.  CTA c gets the contiguous range [c*T/n, (c+1)*T/n): every CTA differs from
// the mean by < 1 tile.  A slot is a maximal run of one object's tiles inside one CTA; slots of an object are
// contiguous, so the optimiser kernel sums slots [obj_slot[o], obj_slot[o+1]) in a fixed order (deterministic).
inline Schedule build_schedule(int n_obj, int rays_per_step, int n_sm) {
    Schedule s;
    s.n_obj = n_obj;
    s.tiles_per_obj = tiles_per_object(rays_per_step);
    const long long T = (long long)n_obj * s.tiles_per_obj;
    s.n_cta = (int)(T < n_sm ? T : n_sm);
    if (s.n_cta < 1) s.n_cta = 1;
    s.data.assign(sched_ints(s.n_cta, n_obj), 0);
    int* ct = s.data.data();
    int* cs = ct + s.n_cta + 1;
    int* os = cs + s.n_cta;
    for (int c = 0; c <= s.n_cta; ++c) ct[c] = (int)((T * c) / s.n_cta);
    int slot = 0;
    std::vector<int> per_obj(n_obj, 0);
    for (int c = 0; c < s.n_cta; ++c) {
        cs[c] = slot;
        if (ct[c] >= ct[c + 1]) continue;
        const int o0 = ct[c] / s.tiles_per_obj, o1 = (ct[c + 1] - 1) / s.tiles_per_obj;
        for (int o = o0; o <= o1; ++o) per_obj[o]++;
        slot += o1 - o0 + 1;
    }
    s.n_slots = slot;
    os[0] = 0;
    for (int o = 0; o < n_obj; ++o) os[o + 1] = os[o] + per_obj[o];
    return s;
}

}  // namespace oo
