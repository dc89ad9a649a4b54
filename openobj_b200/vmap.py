"""Mirror of objnerf/vmap.py: `sceneObject` (keyframe rings + sampling + eval render + checkpoints) and
`cameraInfo`, with the reference's constructor / method signatures.  Sampling and rendering run in the CUDA
kernels K2 / K5; the keyframe bookkeeping (vmap.py:166-257) is host logic restated here."""
import copy
import ctypes
import os
import random

import numpy as np
import torch

from . import layout, sampler, trainer
from ._lib import RenderArgs, check, lib, ptr, stream


class KeyframeRing:
    """Which ring slot a new frame lands in, and which keyframes survive (vmap.py:166-257).

    Keyframe every `keyframe_step` appended frames (float step, e.g. 2.5 -> every 5th), otherwise the newest slot
    is overwritten; once buffer_size-1 keyframes exist the frame goes to `kf_pointer`, and when that frame is a
    keyframe a random non-latest entry is pruned with Python's `random.choice` (vmap.py:256)."""

    def __init__(self, first_frame_id, buffer_size, keyframe_step):
        self.buffer_size, self.keyframe_step = buffer_size, keyframe_step
        self.n_keyframes = 1
        self.kf_pointer = None
        self.kf_buffer_full = False
        self.frame_cnt = 0
        self.latest = []
        self.slot_of = {first_frame_id: 0}        # insertion-ordered frame id -> slot (the reference's bidict)

    def _rekey(self, slot, frame_id):
        for k in [k for k, v in self.slot_of.items() if v == slot]:
            del self.slot_of[k]
        self.slot_of[frame_id] = slot

    def push(self, frame_id):
        """Returns the slot to write the frame into."""
        is_kf = (self.frame_cnt % self.keyframe_step == 0) or self.n_keyframes == 1
        if self.n_keyframes == self.buffer_size - 1:
            self.kf_buffer_full = True
            if self.kf_pointer is None:
                self.kf_pointer = self.n_keyframes
            slot = self.kf_pointer
            self._rekey(slot, frame_id)
            if is_kf:
                self.latest.append(slot)
                _, self.kf_pointer = random.choice(list(self.slot_of.items())[:-2])
        elif not is_kf:
            slot = self.n_keyframes - 1
            self._rekey(slot, frame_id)
        else:
            slot = self.n_keyframes
            self.slot_of[frame_id] = slot
            self.latest.append(slot)
            self.n_keyframes += 1
        self.frame_cnt += 1
        self.latest = self.latest[-2:]
        return slot


class RingBank:
    """The keyframe policy of MANY objects as arrays (scene.Scene): the same decisions as one KeyframeRing per object
    (vmap.py:166-257), taken for all objects a frame shows in a handful of numpy operations instead of a Python loop -- the
    per-frame host work ahead of the first kernel is what a short training window pays for.  The insertion-ordered
    frame-id -> slot dictionary of the reference (`kf_id_dict`, whose ORDER the pruning draw depends on) is kept as two arrays
    per object: the frame id held by each slot and a stamp of when the slot was last (re)keyed; the dictionary's order is the
    order of the stamps.  Rings whose buffer is full (pruning with Python's `random.choice`) are stepped one by one, in the
    order given, so the draws are consumed exactly as by per-object rings.  `ring(i)` is a KeyframeRing-compatible view."""

    def __init__(self, cap, buffer_size):
        self.cap, self.K = int(cap), int(buffer_size)
        n, K = self.cap, self.K
        self.frame_cnt = np.zeros(n, dtype=np.int64)
        self.n_kf = np.zeros(n, dtype=np.int32)
        self.step = np.ones(n, dtype=np.float64)
        self.full = np.zeros(n, dtype=bool)
        self.kf_ptr = np.full(n, -1, dtype=np.int32)           # -1 = None
        self.lat = np.zeros((n, 2), dtype=np.int32)
        self.lat_len = np.zeros(n, dtype=np.int32)
        self.fid = np.full((n, K), -1, dtype=np.int64)         # frame id held by each slot
        self.stamp = np.full((n, K), -1, dtype=np.int64)       # when the slot was last (re)keyed; -1 = never
        self.use_frame = np.zeros((n, K))                      # sceneObject.use_frame rows (float64 like the reference)
        self.clock = 0

    def add(self, i, first_frame_id, keyframe_step):
        self.frame_cnt[i], self.n_kf[i], self.step[i], self.full[i], self.kf_ptr[i] = 0, 1, keyframe_step, False, -1
        self.lat_len[i] = 0
        self.fid[i], self.stamp[i] = -1, -1
        self.fid[i, 0], self.stamp[i, 0] = first_frame_id, self._tick()
        return BankRing(self, i)

    def _tick(self):
        self.clock += 1
        return self.clock

    def _append_latest(self, i, slot):
        if self.lat_len[i] < 2:
            self.lat[i, self.lat_len[i]] = slot
            self.lat_len[i] += 1
        else:
            self.lat[i, 0], self.lat[i, 1] = self.lat[i, 1], slot

    def ordered_items(self, i):
        """[(frame id, slot)] in the reference dictionary's order."""
        used = np.nonzero(self.stamp[i] >= 0)[0]
        order = used[np.argsort(self.stamp[i, used], kind="stable")]
        return [(int(self.fid[i, s]), int(s)) for s in order]

    def push_one(self, i, frame_id):
        """KeyframeRing.push for ring i (the general path: full buffers prune with random.choice)."""
        K = self.K
        nk, cnt = int(self.n_kf[i]), int(self.frame_cnt[i])
        is_kf = (cnt % float(self.step[i]) == 0) or nk == 1
        if nk == K - 1:
            self.full[i] = True
            if self.kf_ptr[i] < 0:
                self.kf_ptr[i] = nk
            slot = int(self.kf_ptr[i])
            self.fid[i, slot], self.stamp[i, slot] = frame_id, self._tick()
            if is_kf:
                self._append_latest(i, slot)
                _, ptr_ = random.choice(self.ordered_items(i)[:-2])
                self.kf_ptr[i] = ptr_
        elif not is_kf:
            slot = nk - 1
            self.fid[i, slot], self.stamp[i, slot] = frame_id, self._tick()
        else:
            slot = nk
            self.fid[i, slot], self.stamp[i, slot] = frame_id, self._tick()
            self._append_latest(i, slot)
            self.n_kf[i] = nk + 1
        self.frame_cnt[i] = cnt + 1
        return slot

    def push_many(self, ii, frame_id):
        """One new frame for the rings `ii` (int array, in the order the reference would visit them) -> their slots."""
        ii = np.asarray(ii, dtype=np.int64)
        K = self.K
        nk, cnt = self.n_kf[ii].astype(np.int64), self.frame_cnt[ii]
        is_kf = (np.mod(cnt.astype(np.float64), self.step[ii]) == 0) | (nk == 1)
        fullm = nk == K - 1
        t = self._tick()
        # ---- rings with room: a keyframe takes the next slot, any other frame overwrites the newest one
        slots = np.where(is_kf, nk, nk - 1)
        # ---- full rings: the frame goes to the slot the last pruning draw freed (the first time: the one spare slot)
        if fullm.any():
            f = ii[fullm]
            self.full[f] = True
            first = self.kf_ptr[f] < 0
            self.kf_ptr[f[first]] = K - 1
            slots[fullm] = self.kf_ptr[f]
        self.fid[ii, slots], self.stamp[ii, slots] = frame_id, t
        ik, sk = ii[is_kf], slots[is_kf]
        if ik.size:                                           # keyframes enter the latest-two queue
            two = self.lat_len[ik] >= 2
            i2, s2 = ik[two], sk[two]
            self.lat[i2, 0] = self.lat[i2, 1]
            self.lat[i2, 1] = s2
            i1, s1 = ik[~two], sk[~two]
            self.lat[i1, self.lat_len[i1]] = s1
            self.lat_len[i1] += 1
            grow = is_kf & ~fullm
            self.n_kf[ii[grow]] += 1
        prune = is_kf & fullm
        if prune.any():
            # random.choice(list(kf_id_dict.items())[:-2]) (vmap.py:256): every slot of a full ring is keyed, the dictionary's
            # order is the stamps' order; choice(seq) is seq[randbelow(len(seq))], so a range of the same length consumes the
            # generator identically -- one draw per ring, in the order given
            rows = ii[prune]
            order = np.argsort(self.stamp[rows], axis=1, kind="stable")[:, :K - 2]
            n = range(K - 2)
            js = np.fromiter((random.choice(n) for _ in range(rows.size)), dtype=np.int64, count=rows.size)
            self.kf_ptr[rows] = order[np.arange(rows.size), js]
        self.frame_cnt[ii] += 1
        return slots


class BankRing:
    """KeyframeRing's attributes and push() for ring i of a RingBank."""

    def __init__(self, bank, i):
        self.bank, self.i = bank, int(i)
        self.buffer_size = bank.K

    keyframe_step = property(lambda self: float(self.bank.step[self.i]))
    n_keyframes = property(lambda self: int(self.bank.n_kf[self.i]))
    kf_pointer = property(lambda self: None if self.bank.kf_ptr[self.i] < 0 else int(self.bank.kf_ptr[self.i]))
    kf_buffer_full = property(lambda self: bool(self.bank.full[self.i]))
    frame_cnt = property(lambda self: int(self.bank.frame_cnt[self.i]))
    latest = property(lambda self: [int(v) for v in self.bank.lat[self.i, :self.bank.lat_len[self.i]]])
    slot_of = property(lambda self: dict(self.bank.ordered_items(self.i)))

    def push(self, frame_id):
        return self.bank.push_one(self.i, frame_id)


class sceneObject:
    def __init__(self, cfg, obj_id, rgb, depth, mask, bbox_2d, t_wc, live_frame_id, clip_feat=None, caption_feat=None,
                 shared=False, bank=None, bank_index=None):
        """`shared=True` (used by scene.Scene): the object owns NO pixel rings -- its keyframes are slots of the scene's shared
        frame store (framestore.FrameStore) and its pixel state is derived from the stored instance map;
        only the keyframe policy (self.ring) and the semantic features live here."""
        assert rgb.shape[:2] == depth.shape and tuple(bbox_2d.shape) == (4,) and tuple(t_wc.shape) == (4, 4)
        assert shared or rgb.shape[:2] == mask.shape
        defer_write = shared
        self.do_bg, self.obj_id = cfg.do_bg, obj_id
        self.data_device, self.training_device = cfg.data_device, cfg.training_device
        self.part_mode, self.stride = cfg.part_mode, cfg.stride
        bg = self.do_bg and obj_id == 0
        self.obj_scale = cfg.bg_scale if bg else cfg.obj_scale
        self.hidden_feature_size = cfg.hidden_feature_size_bg if bg else cfg.hidden_feature_size
        self.n_bins_cam2surface = cfg.n_bins_cam2surface_bg if bg else cfg.n_bins_cam2surface
        self.keyframe_step = cfg.keyframe_step_bg if bg else cfg.keyframe_step
        self.frames_width, self.frames_height = rgb.shape[0], rgb.shape[1]
        self.min_bound, self.max_bound = cfg.min_depth, cfg.max_depth
        self.n_bins, self.n_unidir_funcs = cfg.n_bins, cfg.n_unidir_funcs
        self.surface_eps, self.stop_eps = cfg.surface_eps, cfg.stop_eps
        self.keyframe_buffer_size = cfg.keyframe_buffer_size
        # bank: the scene keeps every object's keyframe policy in one RingBank (arrays); this object's ring is a view of it
        self.ring = (bank.add(bank_index, live_frame_id, self.keyframe_step) if bank is not None
                     else KeyframeRing(live_frame_id, self.keyframe_buffer_size, self.keyframe_step))
        self._bank_row = bank.use_frame[bank_index] if bank is not None else None
        self.feat_cnt, self.clip_feat, self.caption_feat = 1, clip_feat, caption_feat
        self.eps_fine_vis, self.n_bins_fine_vis = cfg.eps_fine_vis, cfg.n_bins_fine_vis
        dev, K, W, H = self.data_device, self.keyframe_buffer_size, self.frames_width, self.frames_height
        self.rgb_idx, self.state_idx = slice(0, 3), slice(3, 4)
        self.shared = bool(shared)
        if shared:
            self.bbox = self.rgbs_batch = self.depth_batch = self.t_wc_batch = None
        else:
            self.bbox = torch.empty(K, 4, device=dev)                       # [w_lo, w_hi, h_lo, h_hi] per slot
            self.rgbs_batch = torch.empty(K, W, H, 4, dtype=torch.uint8, device=dev)   # rgb + pixel state
            self.depth_batch = torch.empty(K, W, H, dtype=torch.float32, device=dev)
            self.t_wc_batch = torch.empty(K, 4, 4, dtype=torch.float32, device=dev)
        if self.part_mode:
            self.part_down = cfg.part_down
            self.use_frame = self._bank_row if self._bank_row is not None else np.zeros(K)
            self.use_frame[:] = 0
        self.other_obj, self.this_obj, self.unknown_obj = 0, 1, 2
        self.semantic_id = None
        if defer_write:
            if self.part_mode:
                self.use_frame[0] = live_frame_id
        else:
            self._write_slot(0, rgb, depth, mask, bbox_2d, t_wc, live_frame_id)
        tcfg = copy.deepcopy(cfg)
        tcfg.obj_id, tcfg.hidden_feature_size, tcfg.obj_scale = obj_id, self.hidden_feature_size, self.obj_scale
        self.trainer = trainer.Trainer(tcfg)
        self.bbox_final, self.serialized_bbox, self.bbox3d, self.bbox3dour, self.pc = False, None, None, None, []
        self.obj_center = torch.tensor(0.0)

    # reference attribute names, backed by the ring
    n_keyframes = property(lambda self: self.ring.n_keyframes)
    kf_pointer = property(lambda self: self.ring.kf_pointer)
    kf_buffer_full = property(lambda self: self.ring.kf_buffer_full)
    frame_cnt = property(lambda self: self.ring.frame_cnt)
    lastest_kf_queue = property(lambda self: self.ring.latest)
    kf_id_dict = property(lambda self: self.ring.slot_of)

    def _write_slot(self, s, rgb, depth, mask, bbox_2d, t_wc, frame_id):
        self.rgbs_batch[s, :, :, self.rgb_idx] = rgb
        self.rgbs_batch[s, :, :, self.state_idx] = mask[..., None]
        self.depth_batch[s] = depth
        self.t_wc_batch[s] = t_wc
        self.bbox[s] = bbox_2d
        if self.part_mode:
            self.use_frame[s] = frame_id

    def append_keyframe(self, rgb, depth, mask, bbox_2d, t_wc, frame_id=1, clip_feat=None, caption_feat=None):
        assert rgb.dtype == torch.uint8 and mask.dtype == torch.uint8 and depth.dtype == torch.float32
        assert self.n_keyframes <= self.keyframe_buffer_size - 1
        self._write_slot(self.ring.push(frame_id), rgb, depth, mask, bbox_2d, t_wc, frame_id)
        if clip_feat is not None:
            self.clip_feat = np.vstack((self.clip_feat, clip_feat))
            self.caption_feat = np.vstack((self.caption_feat, caption_feat))
            self.feat_cnt += 1

    def push_slot(self, frame_id, clip_feat=None, caption_feat=None):
        """Keyframe policy + semantic-feature accumulation of append_keyframe (vmap.py:166-257) without any pixel copy:
        returns the ring slot the new frame takes (scene.Scene points that slot at the shared frame store)."""
        s = self.ring.push(frame_id)
        if self.part_mode:
            self.use_frame[s] = frame_id
        if clip_feat is not None and self.clip_feat is not None:
            self.clip_feat = np.vstack((self.clip_feat, clip_feat))        # vmap.py:241-246
            self.caption_feat = np.vstack((self.caption_feat, caption_feat))
            self.feat_cnt += 1
        return s

    def add_semantic(self, clip_feat, caption_feat):
        """The semantic-feature accumulation of append_keyframe (vmap.py:241-246) on its own."""
        if clip_feat is not None and self.clip_feat is not None:
            self.clip_feat = np.vstack((self.clip_feat, clip_feat))
            self.caption_feat = np.vstack((self.caption_feat, caption_feat))
            self.feat_cnt += 1

    def prune_keyframe(self):
        return random.choice(list(self.ring.slot_of.items())[:-2])

    def part_frame_row(self):
        """(use_frame / stride).long() per slot (vmap.py:438-440, float64 on the host)."""
        return torch.from_numpy((self.use_frame / self.stride).astype(np.int64)).to(torch.int32)

    def get_training_samples(self, n_frames, n_samples, cached_rays_dir, global_partfeat, tapes=None):
        """vmap.py:386-454 -> (gt_rgb, gt_depth, valid_mask, labels, pcs, z, partfeat), one launch of K2.
        Random draws come from torch's generator on the data device in the reference's order unless `tapes`
        (a sampler.SampleTapes with n_obj = 1) is given."""
        if self.shared:
            raise RuntimeError("this sceneObject's keyframes live in a Scene's shared frame store: sample through Scene.sample()")
        dev = self.rgbs_batch.device
        n_rays, S = n_frames * n_samples, self.n_bins_cam2surface + self.n_bins
        if tapes is None:
            nk = self.n_keyframes
            draws = torch.randint(0, nk, (n_frames - 2 if nk > 2 else n_frames,), dtype=torch.long, device=dev)
            kf = sampler.latest_kf_ids(draws, nk, self.lastest_kf_queue)
            u_w = torch.rand(n_frames, n_samples, device=dev)
            u_h = torch.rand(n_frames, n_samples, device=dev)
            tapes = sampler.SampleTapes(kf[None].contiguous(), u_w.view(1, -1), u_h.view(1, -1),
                                        torch.rand(1, n_rays, S, device=dev),
                                        torch.rand(1, n_rays, self.n_bins_cam2surface, device=dev),
                                        torch.empty(1, n_rays, self.n_bins, device=dev).normal_(0., self.surface_eps / 3.),
                                        torch.rand(1, n_rays, self.n_bins, device=dev), by_rank=True)
        part = self.part_mode and global_partfeat is not None
        out = sampler.sample([self.rgbs_batch], [self.depth_batch], [self.t_wc_batch], [self.bbox],
                             self.part_frame_row()[None].to(dev).contiguous() if part else None, cached_rays_dir, tapes,
                             n_frames, n_samples, self.n_bins_cam2surface, self.n_bins, self.surface_eps, self.stop_eps,
                             self.min_bound, self.part_down if part else 0,
                             tuple(global_partfeat.shape[1:3]) if part else (0, 0), kf_cap=self.keyframe_buffer_size)
        pf = None
        if part:
            pf = global_partfeat.reshape(-1, global_partfeat.shape[-1])[out.feat_row[0].long()].view(n_frames, n_samples, -1)
        return (out.gt_rgb[0].view(n_frames, n_samples, 3), out.gt_depth[0].view(n_frames, n_samples),
                out.valid[0].bool(), out.labels[0], out.pcs[0].view(n_frames, n_samples, S, 3),
                out.z[0].view(n_frames, n_samples, S), pf)

    def sample_3d_points(self, sampled_rgbs, sampled_depth, origins, dirs_w, sampled_partfeat=None, draws=None):
        """vmap.py:456-554 on caller-supplied rays (get_training_samples does all of this in one oo_sample_rays launch; this
        is the reference's stand-alone entry, composed from the same utils helpers in the same order, so it consumes
        torch's generator exactly like the reference: rand(invalid, S), rand(valid, N), normal_(this-object, M),
        rand(other, M)).  sampled_rgbs [F,P,4] u8 (rgb + pixel state), sampled_depth [F,P], origins [F,3], dirs_w [F,P,3].
        `draws` = (invalid, valid, normal, other) tensors replacing those four generator calls (tests).
        Returns (rgb, depth, valid_depth_mask, obj_labels, input_pcs [F,P,S,3], sampled_z [F,P,S], sampled_partfeat)."""
        from . import utils
        dr = draws if draws is not None else (None, None, None, None)
        nc, nb, eps = self.n_bins_cam2surface, self.n_bins, self.surface_eps
        dev = sampled_depth.device
        F_, P_ = sampled_rgbs.shape[0], sampled_rgbs.shape[1]
        sampled_z = torch.zeros(F_ * P_, nc + nb, dtype=torch.float32, device=dev)
        d = sampled_depth.reshape(-1).float()
        invalid = d <= self.min_bound
        max_bound = torch.max(sampled_depth)
        n_inv = int(invalid.count_nonzero())
        if n_inv:
            sampled_z[invalid, :] = utils.stratified_bins(self.min_bound, max_bound.reshape(1), nc + nb, n_inv, device=dev, draws=dr[0])
        valid = ~invalid
        n_val = int(valid.count_nonzero())
        if n_val:
            sampled_z[valid, :nc] = utils.stratified_bins(self.min_bound, d[valid] - eps, nc, n_val, device=dev, draws=dr[1])
            state = sampled_rgbs[..., -1].reshape(-1)
            obj = (state == self.this_obj) & valid
            n_obj = int(obj.count_nonzero())
            if n_obj:
                sampled_z[obj, nc:] = utils.normal_bins_sampling(d[obj], nb, n_obj, delta=eps, device=dev, draws=dr[2])
            oth = (state != self.this_obj) & valid
            n_oth = int(oth.count_nonzero())
            if n_oth:
                sampled_z[oth, nc:] = utils.stratified_bins(d[oth] - eps, d[oth] + self.stop_eps, nb, n_oth, device=dev, draws=dr[3])
        org = origins[:, None, :].expand(F_, P_, 3).reshape(-1, 3)
        center = None if float(torch.as_tensor(self.obj_center).abs().sum()) == 0.0 else torch.as_tensor(self.obj_center).expand(3)
        pcs, _ = utils.ray_points(org, dirs_w.reshape(-1, 3), sampled_z, center=center)
        return (sampled_rgbs[..., :3], sampled_depth, valid, sampled_rgbs[..., -1].reshape(-1), pcs.view(F_, P_, nc + nb, 3),
                sampled_z.view(F_, P_, nc + nb), sampled_partfeat)

    def set_semantic(self, semantic_id):
        self.semantic_id = semantic_id            # vmap.py:284-285

    def get_bound(self, intrinsic_open3d=None, final=False):
        if self.bbox_final or self.bbox3dour is not None:
            return self.bbox3d, self.bbox3dour
        raise NotImplementedError("3-D bounds come from open3d/trimesh in the reference (vmap.py:285-379); outside the "
                                  "accelerated path -- set .bbox3dour (center, R, extent) from the host pipeline")

    def _render(self, T_WC, cached_rays_dir, jitter=None, render_part=False, out=None, want_rec=False, theta=None,
                jitter_by_pixel=False, tensor_core=True):
        """One K5 launch over every pixel.  out = (mask u8 [W,H], depth f32 [W,H], rgb u8 [W,H,3]) views to write into (e.g. a
        rank's all-gather send buffer) or None; want_rec: compact per-hit records for the winner-only feature path instead of
        a dense [W,H,512] map.  Returns dict(mask, depth, rgb, feat, rec, hit_pix, n_hit)."""
        _, bb = self.get_bound(None, final=True)
        dev = torch.device(self.training_device)
        W, H = cached_rays_dir.shape[:2]
        T_wc = torch.as_tensor(np.asarray(T_WC), dtype=torch.float32)
        T_wo = torch.eye(4)
        T_wo[:3, :3] = torch.as_tensor(np.asarray(bb.R), dtype=torch.float32)
        T_wo[:3, 3] = torch.as_tensor(np.asarray(bb.center), dtype=torch.float32)
        T_oc = torch.inverse(T_wo) @ T_wc                                  # trainer.py:157-160 (4x4 host algebra)
        half = torch.as_tensor(np.asarray(bb.extent), dtype=torch.float32) / 2.0
        n_bins = 150
        by_rank = jitter is not None and not jitter_by_pixel        # an explicit tape holds the reference's rows: by hit rank
        if jitter is None:
            jitter = torch.rand(W * H, n_bins, device=dev)
        lin = sampler.torch_linspace01(n_bins)
        theta = self.trainer.packed(dev) if theta is None else theta
        f32 = dict(dtype=torch.float32, device=dev)
        if out is None:
            mask = torch.empty(W, H, dtype=torch.uint8, device=dev)
            depth = torch.empty(W, H, **f32)
            rgb = torch.empty(W, H, 3, dtype=torch.uint8, device=dev)
        else:
            mask, depth, rgb = out
        feat = torch.empty(W, H, layout.CLIP, **f32) if (render_part and not want_rec) else None
        rec = torch.empty(W * H, 36, **f32) if want_rec else None
        hit_pix = torch.empty(W * H, dtype=torch.int32, device=dev) if want_rec else None
        n_hit = torch.zeros(1, dtype=torch.int32, device=dev)
        keep = [T_wc.to(dev), T_oc.to(dev).contiguous(), half.to(dev), cached_rays_dir.to(dev).contiguous(),
                jitter.to(dev).contiguous()]
        a = RenderArgs()
        a.W, a.H, a.n_bins, a.scale = W, H, n_bins, float(self.trainer.obj_scale)
        a.theta1, a.T_wc, a.T_oc, a.half_extent, a.rays_dir, a.jitter = [ptr(t) for t in [theta] + keep]
        a.jitter_by_rank = int(by_rank)
        a.lin_host = ctypes.c_void_p(lin.data_ptr())
        a.mask, a.depth, a.rgb, a.feat, a.opacity, a.n_hit = ptr(mask), ptr(depth), ptr(rgb), ptr(feat), None, ptr(n_hit)
        a.ray_rec, a.hit_pix = ptr(rec), ptr(hit_pix)
        from . import ops
        err = ops._TC_ERR.get(dev)
        if err is None:
            err = ops._TC_ERR[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
        a.force_mma_sync, a.tc_err = int(not tensor_core), ptr(err)
        with torch.cuda.device(dev):
            check(lib().oo_render_object(ctypes.byref(a), stream()), "oo_render_object")
        return dict(mask=mask, depth=depth, rgb=rgb, feat=feat, rec=rec, hit_pix=hit_pix, n_hit=n_hit, theta=theta)

    def render_2D_syn(self, T_WC, intrinsic_open3d, cached_rays_dir, T_WO=None, chunk_size=1000, do_fine=True,
                      obj_mask=None, render_part=False, jitter=None, dense=False):
        """vmap.py:604-685 over every pixel, as one K5 launch.  Returns (obj_mask [W,H] bool, depth[n], rgb[n,3] u8,
        feat[n,512] or None) compressed by the mask like the reference; `dense=True` returns device maps instead.
        `jitter` [>= n_hit, 150] replaces the reference's torch.rand(n_hit, 150) (rows by hit rank)."""
        r = self._render(T_WC, cached_rays_dir, jitter=jitter, render_part=render_part)
        mask, depth, rgb, feat, n_hit = r["mask"], r["depth"], r["rgb"], r["feat"], r["n_hit"]
        dev = mask.device
        if dense:
            return mask.bool(), depth, rgb, feat
        if int(n_hit.item()) <= 1:
            return None, None, None
        m = mask.bool()
        if obj_mask is not None:
            m = m & torch.as_tensor(obj_mask, device=dev)
        return (m.cpu().numpy(), depth[m].cpu().numpy(), rgb[m].cpu().numpy(),
                feat[m].cpu().numpy() if render_part else None)

    def save_checkpoints(self, path, epoch):
        """Same dict keys as vmap.py:556-576 (the viewer, visualization/gen_map_vis.py, reads these)."""
        # the modules' tensors may be views of a whole ensemble block [N, 30720]: clone them, or torch.save would write the
        # entire storage (every other object's weights) into each file
        clone = lambda sd: type(sd)((k, v.detach().clone()) for k, v in sd.items())      # noqa: E731
        torch.save({"epoch": epoch, "FC_state_dict": clone(self.trainer.fc_occ_map.state_dict()),
                    "PE_state_dict": clone(self.trainer.pe.state_dict()), "obj_id": self.obj_id, "bbox": self.bbox3dour,
                    "obj_scale": self.trainer.obj_scale, "clip_feat": self.clip_feat, "caption_feat": self.caption_feat,
                    "semantic_id": self.semantic_id}, os.path.join(path, "obj_" + str(self.obj_id) + ".pth"))

    def load_checkpoints(self, ckpt_file):
        if not os.path.exists(ckpt_file):
            print("ckpt not exist ", ckpt_file)
            return
        ck = torch.load(ckpt_file, weights_only=False)
        self.trainer.fc_occ_map.load_state_dict(ck["FC_state_dict"])
        self.trainer.pe.load_state_dict(ck["PE_state_dict"])
        self.obj_id, self.bbox3dour, self.trainer.obj_scale = ck["obj_id"], ck["bbox"], ck["obj_scale"]
        self.trainer.fc_occ_map.to(self.training_device)
        self.trainer.pe.to(self.training_device)
        if "clip_feat" not in ck:
            return False
        self.clip_feat, self.caption_feat, self.semantic_id = ck["clip_feat"], ck["caption_feat"], ck["semantic_id"]
        self.bbox_final = True
        return True


class cameraInfo:
    """vmap.py:687-720.  Ray directions are built on the CPU (true division) and then moved, so that they equal the
    reference's CPU values bit for bit (a CUDA build of the reference multiplies by 1/fx instead)."""

    def __init__(self, cfg):
        self.device = cfg.data_device
        self.width, self.height = cfg.W, cfg.H
        self.fx, self.fy, self.cx, self.cy = cfg.fx, cfg.fy, cfg.cx, cfg.cy
        self.rays_dir_cache = self.get_rays_dirs()

    def get_rays_dirs(self, depth_type="z"):
        if depth_type != "z":
            raise Exception("Get camera rays directions with euclidean depth not yet implemented")
        dirs = torch.ones((self.width, self.height, 3))
        dirs[:, :, 0] = ((torch.arange(end=self.width) - self.cx) / self.fx)[:, None]
        dirs[:, :, 1] = ((torch.arange(end=self.height) - self.cy) / self.fy)
        return dirs.to(self.device)
