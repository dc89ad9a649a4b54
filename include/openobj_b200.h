/*
 * openobj_b200 -- C ABI of the B200-native (sm_100a) replacement for OpenObj's
 * vectorised per-object NeRF training hot path (reference: BIT-DYN/OpenObj, objnerf/).
 *
 * The reference has no FFI layer: its boundary is the Python call surface used by
 * objnerf/train.py.  The Python modules under openobj_b200/ keep those names and
 * signatures and call the entry points below through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the library borrows pointers for the duration of a call and allocates nothing
 *     the caller can see; all launches are asynchronous on `stream` (a cudaStream_t
 *     passed as void*);
 *   - return value: 0 = OK, negative = error (message via oo_last_error());
 *     no C++ exception or abort() crosses this boundary;
 *   - "theta" is the stacked parameter block [n_obj][OO_PSTRIDE] (float32); tensor i of
 *     the reference's named_parameters() order lives at oo_param_offset(i) inside an
 *     object's block, row-major exactly as the reference stores it, so the Python side
 *     exposes the 19 stacked tensors of utils.update_vmap (objnerf/utils.py:55-62) as
 *     strided views of one buffer.
 *   - frames are [W][H]-major like the reference (objnerf/dataset.py:100-106).
 */
#ifndef OPENOBJ_B200_H
#define OPENOBJ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OO_ABI_VERSION 1
#define OO_N_TENSORS 19        /* 18 OccupancyMap tensors (model.py:31-56) + UniDirsEmbed.B_layer.weight (embedding.py:39) */
#define OO_PCOUNT 30659        /* trainable floats per object */
#define OO_PSTRIDE 30720       /* floats between consecutive objects' blocks (16-byte aligned tensors, 128-byte aligned blocks) */
#define OO_HIDDEN 32           /* hidden_feature_size (configs/Replica/room_0.json:53) */
#define OO_CLIP 512            /* clip_point_feature_size (room_0.json:55) */
#define OO_EMB 129             /* 3 + 21*6 (embedding.py:47-53) */
#define OO_EMB1 87             /* trainer.py:20 */
#define OO_NSAMP 10            /* n_bins_cam2surface + n_bins (room_0.json:31-32) */
#define OO_TILE_RAYS 10        /* rays one CTA tile processes */
#define OO_DERIVED_FLOATS 1088
#define OO_RAYREC_FLOATS 36
#define OO_GRAM_PART_FLOATS 1092

/* flags written by oo_label_counts / oo_loss_* (reference: render_rays.py:89-94,109-111) */
#define OO_FLAG_EXPLODE 1      /* some per-object loss term > 1e5: the reference prints and exit(-1)s */
#define OO_FLAG_NO_OBJ 2       /* some object has no label==1 ray: depth/colour/feature terms are 0 for ALL objects */
#define OO_FLAG_NO_SEM 4       /* some object has no label!=2 ray: opacity term is 0 for ALL objects */

int oo_version(void);
const char* oo_last_error(void);

/* offset (floats) and element count of tensor i (0..18) inside an object's block. */
int oo_param_offset(int i);
int oo_param_size(int i);

/* ---- a2+a3: vmap(pe_model) -> vmap(fc_model) forward (objnerf/train.py:424-425,
 *      embedding.py:46-55, model.py:61-103).  pcs [n_obj][n_pts][3]; outputs alpha [n_obj][n_pts]
 *      (already x10), color [n_obj][n_pts][3] (after sigmoid), clip [n_obj][n_pts][512] or NULL.
 *      emb_out [n_obj][n_pts][129] or NULL.  Exactly one of pcs / emb_in is non-NULL: with emb_in
 *      [n_obj][n_pts][129] the encoder is skipped (OccupancyMap.forward on a given embedding). */
int oo_forward(const float* theta, int n_obj, const float* pcs, const float* emb_in, int n_pts, float scale,
               float* alpha, float* color, float* clip, float* emb_out, void* stream);
/* alpha == color == clip == NULL with pcs and emb_out set: the encoder alone (vmap(pe_model), train.py:424). */

/* ---- a10 for hand-composed losses: what autograd's backward (train.py:472) does for
 *      emb = vmap(pe_model)(pe_param, pe_buffer, pcs); alpha, color, clip = vmap(fc_model)(fc_param, fc_buffer, emb)
 *      (train.py:424-425) given dL/d(alpha, color, clip) from the caller.
 *      oo_forward_bwd: emb_in [n_obj][n_pts][129]; d_alpha [n_obj][n_pts] (w.r.t. the x10 output), d_color [n_obj][n_pts][3],
 *      d_clip [n_obj][n_pts][512] or NULL.  grads_out [n_obj][OO_PSTRIDE]: gradients of the 18 OccupancyMap tensors in
 *      theta layout (fully written; the B_layer block is zero).  d_emb_out [n_obj][n_pts][129] or NULL: gradient w.r.t. the
 *      embedding (feeds oo_embed_bwd).  ws: oo_forward_bwd_ws_floats(n_obj, n_pts) floats of caller-owned scratch.
 *      oo_embed_bwd: backward of UniDirsEmbed.forward (embedding.py:46-55) for the trainable direction matrix:
 *      d_B [n_obj][21][3] from pcs [n_obj][n_pts][3] and d_emb; ws: oo_embed_bwd_ws_floats floats. */
int64_t oo_forward_bwd_ws_floats(int n_obj, int n_pts);
int oo_forward_bwd(const float* theta, int n_obj, const float* emb_in, int n_pts, const float* d_alpha, const float* d_color,
                   const float* d_clip, float* grads_out, float* d_emb_out, float* ws, void* stream);
int64_t oo_embed_bwd_ws_floats(int n_obj, int n_pts);
int oo_embed_bwd(const float* theta, int n_obj, const float* pcs, int n_pts, float scale, const float* d_emb, float* d_B,
                 float* ws, void* stream);

/* ---- f3: Trainer.eval_points (objnerf/trainer.py:104-128): occupancy = sigmoid(alpha)
 *      (render_rays.py:6-14), colour and (optionally) the 512-wide part feature of ONE model at n_pts free points
 *      pts [n_pts][3]; occ [n_pts], color [n_pts][3], clip [n_pts][512] or NULL.  One launch for the whole query
 *      (the reference walks it in chunks of 300 000 points, trainer.py:104).
 *      oo_make_grid: the query grid of Trainer.meshing (trainer.py:46-66) = render_rays.make_3D_grid
 *      (render_rays.py:119-146) with scale and a 3x4 transform, minus obj_center: pts_out [dim^3][3], index
 *      (i, j, k) -> (i*dim + j)*dim + k as torch.meshgrid(indexing='ij').  t = the dim linspace values (device). */
typedef struct {
    int dim;
    const float* t;          /* [dim] on the device: torch.linspace(occ_range[0], occ_range[1], dim) */
    float scale[3];
    float transform[12];     /* rows 0..2 of the 4x4 transform, row-major: [R | trans] */
    float center[3];         /* obj_center subtracted last (trainer.py:64) */
} oo_grid;
int oo_make_grid(const oo_grid* g, float* pts_out, void* stream);
/* render_rays.occupancy_activation (render_rays.py:6-14): sigmoid(alpha), or 1 - exp(-alpha * distances) when
 * distances != NULL; n elements. */
int oo_occupancy_activation(const float* alpha, const float* distances, long long n, float* occ, void* stream);
int oo_eval_points(const float* theta, const float* pts, long long n_pts, float pe_scale, float* occ, float* color,
                   float* clip, void* stream);

/* Blackwell-native variant of oo_eval_points without the part-feature output: tcgen05.mma kind::tf32 with three-term error
 * compensation (fp32-level accuracy), accumulators and the chained activations in tensor memory, the object's pre-split
 * weights resident in shared memory; a tile = 128 points = the 128 TMEM lanes (csrc/oo_forward_tc.cu).  occ = sigmoid(alpha)
 * and / or alpha (the x10 output) [n_pts], either may be NULL; color [n_pts][3].  err_flag: device int[1], zeroed by the
 * caller, set to 1 if a tensor-core completion did not arrive within the (bounded) wait. */
int oo_eval_points_tc(const float* theta, const float* pts, long long n_pts, float pe_scale, float* occ, float* alpha,
                      float* color, int* err_flag, void* stream);

/* ---- a4-a9: loss.step_batch_loss forward and backward as one pair of HBM-bound kernels
 *      (objnerf/loss.py:5-103, render_rays.py:6-117).  alpha [N][R][S], color [N][R][S][3],
 *      z [N][R][S], gt_depth [N][R], gt_color [N][R][3] float in [0,1], labels [N][R] u8,
 *      pred_feat [N][R][S][C] / gt_feat [N][R][C] or both NULL.
 *      terms_out [N][4] = per-object depth, colour, opacity, feature terms (after the zero-mask rule);
 *      loss_out [1] = sum_obj (d + cs*c + os*o + fs*f); flags_out [1] int.
 *      Backward writes d_alpha, d_color, d_pred_feat (same shapes; d_pred_feat NULL iff pred_feat NULL)
 *      for upstream gradient grad_loss (host scalar). */
int oo_loss_fwd(const float* alpha, const float* color, const float* z, const float* gt_depth,
                const float* gt_color, const uint8_t* labels, const float* pred_feat, const float* gt_feat,
                int n_obj, int n_rays, int n_samp, int n_feat,
                float color_scaling, float opacity_scaling, float feat_scaling,
                float* terms_out, float* loss_out, int* flags_out, float* ray_ws, void* stream);
int oo_loss_bwd(const float* alpha, const float* color, const float* z, const float* gt_depth,
                const float* gt_color, const uint8_t* labels, const float* pred_feat, const float* gt_feat,
                int n_obj, int n_rays, int n_samp, int n_feat,
                float color_scaling, float opacity_scaling, float feat_scaling, float grad_loss,
                const int* flags, const float* ray_ws,
                float* d_alpha, float* d_color, float* d_pred_feat, void* stream);
/* floats of workspace per (object, ray) that oo_loss_fwd leaves for oo_loss_bwd;
 * ray_ws must hold n_obj*n_rays*oo_loss_ws_per_ray() + 8*n_obj floats */
int oo_loss_ws_per_ray(void);

/* ---- fused training path (a1-a11 in one step: train.py:394-474) ------------------------- */

typedef struct oo_batch {
    const float*   pcs;        /* [n_obj][rays_per_obj][S][3]  sampled points        (train.py:371) */
    const float*   z;          /* [n_obj][rays_per_obj][S]     sampled depths        (train.py:376) */
    const float*   gt_depth;   /* [n_obj][rays_per_obj]                               (train.py:372) */
    const uint8_t* gt_rgb;     /* [n_obj][rays_per_obj][3] u8; /255 happens in-kernel (train.py:373) */
    const uint8_t* labels;     /* [n_obj][rays_per_obj] 0 other / 1 this / 2 unknown  (train.py:375) */
    const int32_t* feat_row;   /* [n_obj][rays_per_obj] row of `feat_table` holding the ray's gt part feature, or NULL = part_mode off */
    const float*   feat_table; /* [rows][512] (global_partfeat flattened, train.py:183-188; or the materialised Batch_N_gt_partfeat) */
    int rays_per_obj;          /* iters_per_frame * n_per_optim = 12000 */
} oo_batch;

typedef struct oo_train_ws {   /* caller-allocated scratch; sizes from oo_train_ws_sizes() */
    float* slab;               /* [n_slots][OO_PSTRIDE] per-(CTA,object) gradient partials */
    float* slot_loss;          /* [n_slots][4] */
    float* derived;            /* [n_obj][OO_DERIVED_FLOATS] out_clip constants (W^T W, W^T b, b.b) of every object, refreshed by every step */
    float* clip_grad;          /* [n_obj][512*32 + 512] out_clip.weight / .bias gradient assembled by K4a for K4b */
    float* rayrec;             /* [n_obj][rays_per_step][OO_RAYREC_FLOATS] per-ray records K1 leaves for K4 (out_clip gradient) */
    int*   sched;              /* device copy of the static schedule */
    int*   counts;             /* [iters][n_obj][2] label==1 / label!=2 ray counts per step */
    int*   flags;              /* [iters] OO_FLAG_* per step (OR over objects; all-reduce across ranks when sharded) */
    float* adam_scal;          /* [iters][3][4]: per step and parameter group {active, lr/bc1, 1/sqrt(bc2), 0} */
    int*   adam_t;             /* [3] persistent Adam step counters per group (trunk+alpha+PE, colour head, clip head) */
    float* gram_part;          /* [n_obj][4][OO_GRAM_PART_FLOATS] partial out_clip Gram matrices of the fused update kernel */
    int*   gram_cnt;           /* [n_obj] arrival counters of the update kernel, zero-initialised by the caller (self-resetting) */
} oo_train_ws;

/* number of CTAs the fused step launches and the scratch sizes (in elements) it needs. */
int oo_train_ws_sizes(int n_obj, int rays_per_step, int iters, int n_sm,
                      int* n_cta, int* n_slots, int64_t* slab_floats, int64_t* sched_ints);
/* build the static ray-range schedule on the host and copy it to ws->sched (synchronous, once per ensemble rebuild). */
int oo_train_schedule(int n_obj, int rays_per_step, int n_sm, oo_train_ws* ws, void* stream);

/* per frame: ray counts + zero-mask flags for every step (render_rays.py:88-94 evaluated up front),
 * then the Adam step/bias-correction schedule (torch.optim.AdamW bookkeeping, train.py:473).
 * Sharded runs: oo_label_counts also writes flag_bits [iters][2] (one int per zero-mask bit), the caller MAX-all-reduces that
 * array over the ranks (= the OR the cross-object rule needs) and hands it to oo_adam_schedule as reduced_bits, which
 * rewrites flags[it] from it.  n_obj == 0 (a rank that owns no object yet) writes all-clear flags. */
int oo_label_counts(const uint8_t* labels, int n_obj, int rays_per_obj, int rays_per_step, int iters,
                    int* counts, int* flags, int* flag_bits, void* stream);
int oo_adam_schedule(int* flags, const int* reduced_bits, int iters, int part_on, float lr, float beta1, float beta2,
                     int* adam_t, float* adam_scal, void* stream);

/* one optimisation step `it` of the whole ensemble: encode + MLP + compositing + loss + backward (K1),
 * then fused multi-tensor AdamW over the stacked blocks (K4).  loss_terms [n_obj][4] (this step) or NULL. */
int oo_train_step(float* theta, float* adam_m, float* adam_v, int n_obj, const oo_batch* batch, int it,
                  int rays_per_step, float scale, float lr, float weight_decay, float beta1, float beta2, float eps,
                  oo_train_ws* ws, float* loss_terms, int n_sm, void* stream);
/* gradients only (no parameter update): grads_out [n_obj][OO_PSTRIDE] in theta layout. For parity tests and
 * for the autograd-facing Python surface. */
int oo_train_grads(const float* theta, int n_obj, const oo_batch* batch, int it, int rays_per_step, float scale,
                   oo_train_ws* ws, float* grads_out, float* loss_terms, int n_sm, void* stream);
/* `iters` consecutive steps (train.py:394 loop); loss_terms [iters][n_obj][4] or NULL. */
int oo_train_frame(float* theta, float* adam_m, float* adam_v, int n_obj, const oo_batch* batch, int iters,
                   int rays_per_step, float scale, float lr, float weight_decay, float beta1, float beta2, float eps,
                   oo_train_ws* ws, float* loss_terms, int n_sm, void* stream);

/* the two halves of oo_train_step, separately launchable (bench.py brackets K1 with CUDA events for the roofline;
 * ncu captures use them too): K1 = fused encode/MLP/composite/loss/backward into ws->slab; K4 = one fused launch:
 * out_clip gradient assembly + slab reduction + AdamW + the out_clip constants (ws->derived) of the NEXT step.
 * refresh_derived != 0 makes K1 recompute ws->derived from theta first (needed whenever theta was not last written by K4). */
int oo_train_k1(const float* theta, int n_obj, const oo_batch* batch, int it, int rays_per_step, float scale,
                oo_train_ws* ws, int refresh_derived, int n_sm, void* stream);
int oo_train_k4(float* theta, float* adam_m, float* adam_v, int n_obj, const oo_batch* batch, int it, int rays_per_step,
                float lr, float weight_decay, float beta1, float beta2, float eps,
                oo_train_ws* ws, float* loss_terms, int n_sm, void* stream);

/* a11 standalone: torch.optim.AdamW over a flat float buffer. step is 1-based. */
int oo_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, int step,
                  float lr, float weight_decay, float beta1, float beta2, float eps, void* stream);

/* ---- a17: the background model (objnerf/train.py:300-315,379-388,447-463; vmap.py:43-47): ONE OccupancyMap of
 *      hidden width `hidden` (128 in room_0.json) + UniDirsEmbed(scale = bg_scale 5), n_per_optim_bg = 1200 rays x
 *      (n_bins_cam2surface_bg 5 + n_bins 9) = 14 samples per step.  M = 16 800 points with K <= 215 are plain GEMMs:
 *      layer-by-layer FP32 kernels + the K3 compositing/loss pair.  Parameters live in ONE flat block in
 *      named_parameters() order (18 fc tensors, then B_layer.weight), offsets rounded to 4 floats. */
int oo_bg_param_count(int hidden);
int oo_bg_param_offset(int hidden, int i);
int oo_bg_param_size(int hidden, int i);
/* floats of caller-allocated scratch for n_pts points / n_rays rays */
int64_t oo_bg_ws_floats(int hidden, int n_pts, int n_rays);
/* pe -> fc_occ_map forward (train.py:449-450): alpha [n_pts] (x10 applied), color [n_pts][3], clip [n_pts][512] or NULL,
 * emb_out [n_pts][129] or NULL.  Exactly one of pcs [n_pts][3] / emb_in [n_pts][129] is non-NULL: with emb_in the encoder is
 * skipped (OccupancyMap.forward on a given embedding, the module-level call form). */
int oo_bg_forward(const float* theta, int hidden, const float* pcs, const float* emb_in, int n_pts, float scale, float* alpha,
                  float* color, float* clip, float* emb_out, float* ws, void* stream);
/* backward of oo_bg_forward for upstream gradients d_alpha [n_pts] (w.r.t. the x10 output), d_color [n_pts][3], d_clip
 * [n_pts][512] or NULL (what autograd does for `fc_occ_map(pe(x))` of the hidden-128 model, train.py:449-450,472):
 * grads_out = the flat gradient block (parameter layout; with emb_in the B_layer block stays zero and d_emb_out
 * [n_pts][129], if set, receives the gradient w.r.t. the embedding). */
int oo_bg_forward_bwd(const float* theta, int hidden, const float* pcs, const float* emb_in, int n_pts, float scale,
                      const float* d_alpha, const float* d_color, const float* d_clip, float* grads_out, float* d_emb_out,
                      float* ws, void* stream);
/* one optimisation step on rays [n_rays] x samples [n_samp] (train.py:447-474 for the background): forward,
 * step_batch_loss on [1,R,S], backward, AdamW.  adam_t: DEVICE int[3], the Adam step counters of the three parameter groups
 * (trunk + alpha + PE / colour head / clip head), zero-initialised by the caller and advanced by the call: a group whose
 * tensors autograd does not reach in this step is skipped entirely, exactly like torch.optim.AdamW skips grad-None tensors --
 * part features off (the clip head; quirk 8) or an empty label mask in this step's step_batch_loss (render_rays.py:89-94
 * with N = 1: no label-1 ray -> colour and clip heads, neither mask -> everything).  grads_out != NULL: write the flat
 * gradient (parameter layout) and do NOT update.  terms_out [4], loss_out [1], flags_out [1] as oo_loss_fwd. */
int oo_bg_train_step(float* theta, float* adam_m, float* adam_v, int hidden, const float* pcs, const float* z,
                     const float* gt_depth, const uint8_t* gt_rgb, const uint8_t* labels, const int32_t* feat_row,
                     const float* feat_table, int n_rays, int n_samp, float scale, int* adam_t, float lr,
                     float weight_decay, float beta1, float beta2, float eps, float color_scaling, float opacity_scaling,
                     float feat_scaling, float* ws, float* terms_out, float* loss_out, int* flags_out, float* grads_out,
                     void* stream);

/* ---- a13-a15: sceneObject.get_training_samples + sample_3d_points for ALL objects in one launch
 *      (objnerf/vmap.py:386-554, utils.py:324-397).  Keyframe rings stay in the reference's layout and
 *      are addressed through per-object pointer tables. */
typedef struct oo_sample_args {
    int n_obj, n_frames, n_samples;      /* per object: n_frames keyframe draws x n_samples pixels (500 x 24) */
    int W, H, n_c2s, n_bins;             /* frame size; 1 + 9 samples per ray */
    float eps, other_eps, min_bound;     /* surface_eps 0.1, other_eps 0.05, min_depth 0 */
    int part_down;                       /* 5; 0 = part_mode off */
    int pw, ph;                          /* part map size (W/part_down, H/part_down) */
    const uint8_t* const* rgbs;          /* [n_obj] -> u8 [KF][W][H][4]  (vmap.py:97-102) */
    const float* const*   depth;         /* [n_obj] -> f32 [KF][W][H]    (vmap.py:131-135) */
    const float* const*   t_wc;          /* [n_obj] -> f32 [KF][4][4]    (vmap.py:139-142) */
    const float* const*   bbox;          /* [n_obj] -> f32 [KF][4] = w_lo,w_hi,h_lo,h_hi (vmap.py:84-89) */
    const int32_t*        part_frame;    /* [n_obj][KF_MAX=20] (use_frame/stride).long()  (vmap.py:438-440) */
    const float*          rays_dir;      /* [W][H][3] cameraInfo.rays_dir_cache (vmap.py:701-720) */
    /* RNG tape; rows of the class tapes are consumed by rank within the class */
    const int64_t* kf_ids;               /* [n_obj][n_frames] */
    const float* u_w;                    /* [n_obj][n_frames*n_samples] */
    const float* u_h;
    const float* r_invalid;              /* [n_obj][n_rays][S]      */
    const float* r_valid;                /* [n_obj][n_rays][n_c2s]  */
    const float* r_normal;               /* [n_obj][n_rays][n_bins] normal_(0, eps/3) draws, unsorted */
    const float* r_other;                /* [n_obj][n_rays][n_bins] */
    int tape_by_rank;                    /* 1: reference order (row j -> j-th ray of the class); 0: row = ray index */
    /* torch.linspace(0,1,n+1) tables used by utils.stratified_bins (utils.py:349) for n = S, n_c2s, n_bins.
     * HOST pointers: ATen's vectorised linspace is not a closed formula, so the caller supplies what torch gives. */
    const float* lin_s_host; const float* lin_c2s_host; const float* lin_bins_host;
    /* rng_mode 1: no tapes; every draw comes from the counter RNG (Philox4x32-10 keyed by seed) evaluated in-kernel.
     * Keyframe draw of frame slot f: element f of oo_rng_fill stream 8*frame + 0; keyframe id = min(int(u * n_keyframes),
     * n_keyframes-1), the last two forced to `latest` when n_keyframes > 2.  Everything else of ray r comes from the
     * ray-blocked stream of oo_rng_fill_rows(frame stream 8*frame + 1, row = r): word 0 = u_w, 1 = u_h, 2 .. 2+n_c2s =
     * r_valid, and from B0 = 4*ceil((2 + n_c2s)/4): r_invalid (S words) / r_other (n_bins words) / r_normal (Box-Muller
     * pairs, std eps/3) -- so the result equals tape mode (row = ray index) with tapes cut from oo_rng_fill_rows. */
    int rng_mode; uint64_t seed; uint32_t frame;
    const int32_t* obj_ids;              /* [n_obj] RNG key per object (its instance id: shard independent) */
    const int32_t* n_keyframes;          /* [n_obj] */
    const int32_t* latest;               /* [n_obj][2] lastest_kf_queue[-2:] (vmap.py:398) */
    /* outputs, all [n_obj][n_rays...] */
    uint8_t* gt_rgb; float* gt_depth; uint8_t* valid; uint8_t* labels;
    float* pcs; float* z; int32_t* feat_row; int64_t* pix;   /* pix [n_obj][n_rays][3] = kf, w, h */
    int* oob_count;                      /* [1] rays whose pixel index had to be clamped (quirk 11) */
    int kf_cap;                          /* keyframe_buffer_size: row length of part_frame / slot_frame / slot_bbox (<= 32) */
    /* shared keyframe store (oo_store_frame): when store_rgbi != NULL the per-object ring tables above
     * (rgbs / depth / t_wc / bbox) are not read.  Ring slot s of object o refers to store frame slot_frame[o][s]; the
     * per-object pixel state (train.py:203-205) is derived from the instance id: 1 where inst == obj_ids[o], 2 where
     * inst == -1, else 0 -- so obj_ids is required in this mode, also with tapes. */
    const int32_t* store_rgbi;           /* [F][W][H][2]: word 0 = r | g << 8 | b << 16, word 1 = instance id */
    const float*   store_depth;          /* [F][W][H] */
    const float*   store_twc;            /* [F][16] camera-to-world of each stored frame */
    const int32_t* slot_frame;           /* [n_obj][kf_cap] */
    const float*   slot_bbox;            /* [n_obj][kf_cap][4] = w_lo,w_hi,h_lo,h_hi (vmap.py:84-89) */
    /* rng_mode 1: caller-owned scratch of >= n_obj * (2 + 2 * n_frames * n_samples) int32 (per-object batch maximum,
     * invalid-depth ray lists); the library allocates nothing. */
    int32_t* scratch; int64_t scratch_ints;
} oo_sample_args;
int oo_sample_rays(const oo_sample_args* a, void* stream);
/* ---- f2: the shared keyframe store.  One new frame -> store slot `slot`, ONE launch and 12 bytes per
 *      pixel whatever the number of objects that see the frame (replaces the per-object ring copies of
 *      sceneObject.append_keyframe, vmap.py:186-240, and the per-object state maps of train.py:203-205). */
typedef struct oo_store_args {
    int W, H, slot;
    const uint8_t* rgb;                  /* [W][H][3] */
    const float*   depth;                /* [W][H] */
    const int32_t* inst;                 /* [W][H] instance map (-1 unknown, 0 background, k > 0 object) */
    const float*   t_wc;                 /* DEVICE float[16]: camera-to-world of this frame */
    int32_t* store_rgbi; float* store_depth; float* store_twc;
} oo_store_args;
int oo_store_frame(const oo_store_args* a, void* stream);
/* Part-feature cells of a frame, only inside the given boxes (train.py:183-188 copies the whole [W/5][H/5][512] tensor; a rank of
 * a sharded run only ever gathers rows inside its own objects' boxes, vmap.py:437-452).  src_host: the frame's tensor in PINNED
 * host memory (read by the kernel over the host link), dst: the same-shaped slot of the resident table on the device,
 * boxes: HOST int32 [n_boxes][4] = w_lo, w_hi, h_lo, h_hi in cells, inclusive; at most 192 boxes per call. */
int oo_gather_part_rows(const float* src_host, float* dst, int pw, int ph, int n_feat, const int32_t* boxes, int n_boxes,
                        void* stream);
/* counter-based uniform / normal tapes keyed by (seed, frame, object id, element) -- shard independent. */
int oo_rng_fill(uint64_t seed, uint32_t frame, const int32_t* obj_ids, int n_obj, int64_t per_obj,
                int kind /*0 uniform [0,1), 1 normal(0,std)*/, float std, float* out, void* stream);

/* ray-blocked variant: out [n_obj][n_rows][row_words]; word w of row r = word (w % 4) of
 * philox(counter = (r, w / 4, object id, frame), key = seed) as a uniform, or (kind 1) the Box-Muller normal of its
 * pair (w even: cos, w odd: sin) -- what oo_sample_rays draws in rng_mode 1 for ray = row. */
int oo_rng_fill_rows(uint64_t seed, uint32_t frame, const int32_t* obj_ids, int n_obj, int n_rows, int row_words,
                     int kind /*0 uniform [0,1), 1 normal(0,std)*/, float std, float* out, void* stream);

/* ---- the stand-alone helpers of the reference surface (utils.{stratified_bins, normal_bins_sampling,
 *      origin_dirs_W, ray_box_intersection}, Trainer.sample_points_bbox, sceneObject.sample_3d_points).  On the hot path
 *      they are fused into oo_sample_rays / oo_render_object; these element-wise kernels give hand-composed callers the
 *      same arithmetic in the reference's operation order.  Random draws are inputs (the host surface draws them where the
 *      reference calls torch.rand / normal_).  All pointers are device pointers unless marked HOST.
 *      oo_ray_box: utils.ray_box_intersection (utils.py:309-319); origins / dirs [n][3]; bounds HOST float[3];
 *                  near / far [n], hit u8 [n] = near <= far && far > 0.
 *      oo_origin_dirs: utils.origin_dirs_W (utils.py:324-336); t_wc [B][4][4], dirs_c [B][n][3] -> origins [B][3], dirs_w.
 *      oo_stratified_bins: utils.stratified_bins (utils.py:342-379) given u [n][n_bins] in [0,1) and
 *                  lin = torch.linspace(0, 1, n_bins + 1) (device); min / max per ray, or NULL -> the scalar.
 *      oo_normal_bins: utils.normal_bins_sampling (utils.py:382-397) given the normal_(0, delta/3) draws [n][n_bins]:
 *                  rows sorted ascending, clipped to +-delta, + depth[n].
 *      oo_ray_points: origins + dirs * z - center (vmap.py:548-551); midpoints = 1: z is replaced by
 *                  0.5 (z[i+1] + z[i]) first (n_samp - 1 points per ray; trainer.py:175-177), written to z_mid_out if set;
 *                  center HOST float[3] or NULL. */
int oo_ray_box(const float* origins, const float* dirs, const float* bounds_min, const float* bounds_max, long long n,
               float* near_out, float* far_out, uint8_t* hit_out, void* stream);
int oo_origin_dirs(const float* t_wc, const float* dirs_c, int n_poses, int n_per_pose, float* origins, float* dirs_w,
                   void* stream);
int oo_stratified_bins(const float* u, const float* min_depth, const float* max_depth, float min_scalar, float max_scalar,
                       const float* lin, long long n_rays, int n_bins, float* z, void* stream);
int oo_normal_bins(const float* draws, const float* depth, long long n_rays, int n_bins, float delta, float* z, void* stream);
int oo_ray_points(const float* origins, const float* dirs, const float* z, long long n_rays, int n_samp, int midpoints,
                  const float* center, float* z_mid_out, float* pcs, void* stream);
/* render_rays.occupancy_to_termination (render_rays.py:32-54): T_i = occ_i * prod_{j<i} (1 - occ_j + 1e-10) along the
 * n_samp samples of each ray; render_rays.render (render_rays.py:56-63): out[r][c] = sum_i T[r][i] vals[r][i][c]
 * (n_chan = 1 for depth / variance, 3 colour, 512 part features).  Forward only: the training path differentiates through
 * the fused kernels (oo_train_step, oo_loss_fwd / _bwd). */
int oo_termination(const float* occupancy, long long n_rays, int n_samp, float* termination, void* stream);
int oo_render_sum(const float* termination, const float* vals, long long n_rays, int n_samp, int n_chan, float* out,
                  void* stream);

/* ---- a19: render_2D_syn for one object over all W*H pixels (vmap.py:604-685, trainer.py:130-198)
 *      and the sequential depth-test merge (train.py:577-594). */
typedef struct oo_render_args {
    int W, H, n_bins;                    /* 150 bins -> 149 midpoints */
    float scale;
    const float* theta1;                 /* this object's parameter block */
    const float* T_wc;                   /* [4][4] f32 */
    const float* T_oc;                   /* [4][4] f32 = inv(T_WO) @ T_WC (trainer.py:157-160, computed by the caller) */
    const float* half_extent;            /* [3] */
    const float* rays_dir;               /* [W][H][3] */
    const float* jitter;                 /* [W*H][n_bins] uniform draws */
    int jitter_by_rank;                  /* 1: row j -> j-th hit ray (reference order) */
    const float* lin_host;               /* HOST: torch.linspace(0,1,n_bins+1) (utils.py:349) */
    uint8_t* mask; float* depth; uint8_t* rgb; float* feat;  /* [W][H], [W][H], [W][H][3], [W][H][512] or NULL */
    float* opacity;                      /* [W][H] or NULL */
    int* n_hit;                          /* [1] */
    /* compact per-hit records for the winner-only feature path (both or neither): ray_rec [W*H][OO_RENDER_REC] = {S[32] =
     * sum_i T_i hp_i (clip_linear activations composited), opacity, 0, 0, 0} of hit j (rank among the rays that intersect the
     * box), hit_pix [W*H] = pixel index of hit j.  Use with feat == NULL, then oo_winner_features after the merge. */
    float* ray_rec; int32_t* hit_pix;
    /* feat == NULL runs on the tcgen05 / TMEM kernel (csrc/oo_forward_tc.cu) unless force_mma_sync != 0 (the mma.sync tile
     * kernel, kept for the dense-feature output and as the comparison arm); tc_err: device int[1] or NULL, set to 1 if a
     * tensor-core completion did not arrive within the bounded wait. */
    int force_mma_sync; int* tc_err;
} oo_render_args;
#define OO_RENDER_REC 36
int oo_render_object(const oo_render_args* a, void* stream);
/* ---- a19 for ALL objects a rank owns, one call per frame (BASELINE config 5): the hit lists of every object are built in three
 *      parallel passes (slab test per (object, pixel); pixels in increasing order inside an object's list, the reference's rank
 *      order, trainer.py:166-176), then ONE launch of the tcgen05 / TMEM kernel walks the pooled hits: a CTA takes an equal share
 *      of the pool and re-stages the weights when its share crosses into the next object.  Jitter rows are taken by pixel.
 *      Outputs: dense per-object maps mask / depth / rgb [n_obj][W*H](x3), ZERO-INITIALISED BY THE CALLER (only hit pixels are
 *      written); obj_start [n_obj + 1] (device) = prefix of the hit counts; hit_pix [pool_rows] pooled pixel ids; ray_rec
 *      [pool_rows][OO_RENDER_REC] pooled per-hit records (or NULL).  If obj_start[n_obj] > pool_rows the excess hits are not
 *      rendered: the caller checks and re-runs with a larger pool.  scratch: >= n_obj * ceil(W*H / 256) + 1 int32. */
typedef struct oo_render_frame_args {
    int W, H, n_bins, n_obj;
    float scale;
    const float* const* theta;           /* DEVICE array of n_obj device pointers to parameter blocks */
    const float* T_wc;                   /* [16] */
    const float* T_oc;                   /* [n_obj][16] = inv(T_WO) @ T_WC per object (trainer.py:157-160) */
    const float* half_extent;            /* [n_obj][3] */
    const float* rays_dir;               /* [W][H][3] */
    const float* jitter;                 /* [W*H][n_bins], rows by pixel */
    const float* lin;                    /* DEVICE: torch.linspace(0, 1, n_bins + 1) */
    uint8_t* mask; float* depth; uint8_t* rgb;
    int32_t* obj_start; int32_t* hit_pix; float* ray_rec; int64_t pool_rows;
    int32_t* scratch; int64_t scratch_ints;
    int* tc_err;
} oo_render_frame_args;
int oo_render_frame(const oo_render_frame_args* f, void* stream);
/* winner-only features over the pooled hits of oo_render_frame: k_of [n_obj] = global ensemble index of each local object */
int oo_winner_features_frame(const float* const* theta, const float* ray_rec, const int32_t* hit_pix, const int* obj_start,
                             int n_obj, const int32_t* k_of, const int32_t* winner, int64_t pool_rows, int64_t cap_rows,
                             float* rows, int32_t* row_pix, int* n_rows, void* stream);
/* the depth-test merge over per-object pointers (device arrays of n_obj pointers to [n_pix] / [n_pix][3] maps, in global
 * insertion order): what a sharded run uses on the all-gathered buffer without re-packing it */
int oo_zmerge_ptr(const uint8_t* const* masks, const float* const* depths, const uint8_t* const* rgbs, const uint8_t* is_bg,
                  int n_obj, int64_t n_pix, float* depth_out, uint8_t* rgb_out, int32_t* winner_out, void* stream);
/* part features of the pixels object k_global won (winner [n_pix] from the merge): rows [<= cap_rows][512] = W_ocl S +
 * b_ocl opacity, row_pix = their pixels, n_rows [1] = running count (zeroed by the caller; rows beyond cap_rows are counted
 * but not written).  The 512-wide layer runs once per WON pixel instead of once per hit of every object
 * (vmap.py:660-670 with render_part, train.py:577-594). */
int oo_winner_features(const float* theta1, const float* ray_rec, const int32_t* hit_pix, const int* n_hit, const int32_t* winner,
                       int k_global, int64_t cap_rows, float* rows, int32_t* row_pix, int* n_rows, void* stream);
int oo_zmerge(const uint8_t* masks, const float* depths, const uint8_t* rgbs, const uint8_t* is_bg, int n_obj,
              int64_t n_pix, float* depth_out, uint8_t* rgb_out, int32_t* winner_out, void* stream);

/* FP32 FFMA-pipe peak probe used as the roofline denominator of the fused step. */
int oo_fma_peak(int n_sm, int iters, float* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif
