"""Per-frame driver of the accelerated path: the body of objnerf/train.py:158-485 for the vmap strategy, with the
Python object loop replaced by three batched launches per frame (append is still per object in round 1):

    add_frame   (train.py:164-256)  H2D, per-object pixel state, keyframe rings, new objects -> ensemble rebuild
    sample      (train.py:300-388)  K2 for all local objects in one launch, no [N,12000,512] feature copy
    train       (train.py:394-474)  100 x (K1 + K4)
    write-back  (train.py:478-485)  not needed: every object's nn.Parameters are views of the ensemble buffer

Objects are sharded by ensemble index k (order of first appearance): rank = k % world; ranks share
nothing but the per-step zero-mask flags, OR-reduced once per frame."""
import torch

from . import layout, sampler, vmap
from .background import BackgroundModel
from .ensemble import Ensemble, FrameBatch


class Scene:
    def __init__(self, cfg, rank=0, world=1, seed=0, max_frames=64, n_sm=None, flag_allreduce=None):
        self.cfg, self.rank, self.world, self.seed = cfg, rank, world, seed
        self.device = torch.device(cfg.training_device)
        self.cam = vmap.cameraInfo(cfg)
        self.obj_dict = {}            # local objects, insertion order == local ensemble index
        self.global_index = {}        # obj id -> global ensemble index k (all ranks agree)
        self.ens = None
        self.n_sm = n_sm
        self.flag_allreduce = flag_allreduce
        self.frames_seen = 0
        self.part_mode = cfg.part_mode
        self.part_table = None
        if self.part_mode:
            self.pw, self.ph = cfg.W // cfg.part_down, cfg.H // cfg.part_down
            self.part_table = torch.empty(max_frames, self.pw, self.ph, cfg.clip_point_feature_size,
                                          dtype=torch.float32, device=self.device)
        self._stale = False
        self.batch = None
        # the separate background model (train.py:236-242): one hidden-128 model, not part of the vmap ensemble; it lives
        # on the last rank (the ensemble's round-robin starts at rank 0)
        self._copy_stream = None
        self.scene_bg, self.bg, self.bg_batch, self.bg_tables = None, None, None, None
        self.bg_rank = world - 1

    # ---- host -> device staging of the NEXT frame on a copy stream (what a DataLoader with pin_memory + non_blocking
    # copies gives the reference, train.py:158-188): the 76 MB of a Replica frame travel while the current frame trains
    def stage_frame(self, sample):
        """Start the H2D copies of `sample` (pinned host tensors) on the copy stream; returns a dict to pass to
        add_frame.  Tensors already on the device pass through."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        out = dict(sample)
        with torch.cuda.stream(self._copy_stream):
            for k in ("image", "depth", "obj", "T", "part_feat"):
                v = sample.get(k)
                if torch.is_tensor(v) and v.device != self.device:
                    out[k] = v.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        out["_staged"] = ev
        return out

    # ---- train.py:164-256 ---------------------------------------------------------------------------------
    def add_frame(self, sample):
        cfg, dev = self.cfg, self.device
        nb = dict(non_blocking=True)
        if "_staged" in sample:
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(sample["_staged"])
            for k in ("image", "depth", "obj", "T", "part_feat"):     # allocated on the copy stream, consumed on this one
                if torch.is_tensor(sample.get(k)) and sample[k].is_cuda:
                    sample[k].record_stream(cur)
        rgb, depth = sample["image"].to(dev, **nb), sample["depth"].to(dev, **nb)
        inst = sample["obj"].to(dev, **nb)
        twc = sample["T"].to(dev, **nb)
        frame_id = sample.get("frame_id", self.frames_seen)
        if self.part_mode:
            if self.frames_seen >= self.part_table.shape[0]:
                raise RuntimeError("part-feature table full: raise max_frames")
            self.part_table[self.frames_seen].copy_(sample["part_feat"], non_blocking=True)   # train.py:183-188
        twc32 = twc.to(torch.float32)
        objs, slots, bboxes = [], [], []
        for obj_id in sorted(int(k) for k in sample["bbox_dict"].keys()):                 # torch.unique order (train.py:191)
            if obj_id == -1:
                continue
            if cfg.do_bg and obj_id == 0:
                # the separate background model is not part of the vmap ensemble (train.py:236-242)
                if self.rank != self.bg_rank:
                    continue
                bbox = sample["bbox_dict"][obj_id]
                if self.scene_bg is None:
                    self.scene_bg = vmap.sceneObject(cfg, 0, rgb, depth, None, bbox, twc, frame_id, defer_write=True)
                    self.bg = BackgroundModel(hidden=cfg.hidden_feature_size_bg, device=self.device,
                                              rays_per_step=cfg.n_per_optim_bg,
                                              n_samp=self.scene_bg.n_bins_cam2surface + self.scene_bg.n_bins,
                                              lr=cfg.learning_rate, weight_decay=cfg.weight_decay, scale=cfg.bg_scale)
                    self.bg.adopt(self.scene_bg.trainer.fc_occ_map, self.scene_bg.trainer.pe)
                    self.bg_tables = sampler.RingTables([self.scene_bg], self.device)
                    slot = 0
                else:
                    slot = self.scene_bg.push_slot(frame_id)
                objs.append(self.scene_bg); slots.append(slot); bboxes.append(bbox)
                continue
            if obj_id not in self.global_index:
                if len(self.global_index) >= cfg.max_n_models * self.world:
                    continue           # "models full" (train.py:231-233)
                self.global_index[obj_id] = len(self.global_index)
            if self.global_index[obj_id] % self.world != self.rank:
                continue
            bbox = sample["bbox_dict"][obj_id]
            if obj_id in self.obj_dict:
                o = self.obj_dict[obj_id]
                slot = o.push_slot(frame_id)
            else:
                o = vmap.sceneObject(cfg, obj_id, rgb, depth, None, bbox, twc, frame_id, defer_write=True)
                self.obj_dict[obj_id] = o
                slot = 0
                self._stale = True
            objs.append(o); slots.append(slot); bboxes.append(bbox)
        # pixel state (train.py:203-205) + ring writes of every visible object in ONE launch
        sampler.append_frame(rgb, depth, inst, twc32, objs, slots, bboxes)
        self.frames_seen += 1
        if self._stale:
            self._rebuild_ensemble()

    # ---- utils.update_vmap (train.py:272-276): restack, Adam state restarts ---------------------------------
    def _rebuild_ensemble(self):
        objs = list(self.obj_dict.values())
        new = Ensemble(len(objs), device=self.device, rays_per_step=self.cfg.n_per_optim,
                       iters_per_frame=self.cfg.n_iter_per_frame, lr=self.cfg.learning_rate,
                       weight_decay=self.cfg.weight_decay, scale=self.cfg.obj_scale, n_sm=self.n_sm)
        views = new.stacked()
        with torch.no_grad():
            for k, o in enumerate(objs):
                ps = list(o.trainer.fc_occ_map.parameters()) + [o.trainer.pe.B_layer.weight]
                for v, p in zip(views, ps):
                    v[k].copy_(p.detach())
                    p.data = v[k]      # the module now aliases the ensemble buffer: write-back (train.py:478-485) is free
        new.params_changed()
        new.reset_optimizer()
        self.ens = new
        self.tables = sampler.RingTables(objs, self.device)
        self._stale = False

    # ---- train.py:300-388 -----------------------------------------------------------------------------------
    def sample(self):
        cfg = self.cfg
        objs = list(self.obj_dict.values())
        n_frames = cfg.n_iter_per_frame * cfg.win_size
        n_samples = cfg.n_samples_per_frame
        o0 = objs[0]
        rng = sampler.counter_rng(objs, self.seed, self.frames_seen, self.device)
        part_frame = None
        if self.part_mode:
            import numpy as np
            pf = np.stack([(o.use_frame / o.stride).astype(np.int64) for o in objs]).astype(np.int32)   # vmap.py:438-440
            part_frame = torch.from_numpy(pf).to(self.device, non_blocking=True)
        out = sampler.sample(None, None, None, None, part_frame, self.cam.rays_dir_cache, rng, n_frames, n_samples,
                             o0.n_bins_cam2surface, o0.n_bins, o0.surface_eps, o0.stop_eps, o0.min_bound,
                             cfg.part_down if self.part_mode else 0, (self.pw, self.ph) if self.part_mode else (0, 0),
                             out=getattr(self, "sample_out", None) if getattr(self, "_out_n", -1) == len(objs) else None,
                             tables=self.tables)
        self._out_n = len(objs)
        table = self.part_table.view(-1, self.part_table.shape[-1]) if self.part_mode else None
        self.batch = FrameBatch(out.pcs, out.z, out.gt_depth, out.gt_rgb, out.labels, out.feat_row, table)
        self.sample_out = out
        if self.scene_bg is not None:
            # train.py:300-315: the background draws n_iter_per_frame * win_size_bg keyframes x n_samples_per_frame_bg
            # pixels with 5 + 9 samples per ray
            b = self.scene_bg
            rng_bg = sampler.counter_rng([b], self.seed, self.frames_seen, self.device)
            pf_bg = None
            if self.part_mode:
                import numpy as np
                pf_bg = torch.from_numpy((b.use_frame / b.stride).astype(np.int64).astype(np.int32)[None]).to(self.device, non_blocking=True)
            ob = sampler.sample(None, None, None, None, pf_bg, self.cam.rays_dir_cache, rng_bg,
                                cfg.n_iter_per_frame * cfg.win_size_bg, cfg.n_samples_per_frame_bg, b.n_bins_cam2surface,
                                b.n_bins, b.surface_eps, b.stop_eps, b.min_bound, cfg.part_down if self.part_mode else 0,
                                (self.pw, self.ph) if self.part_mode else (0, 0), out=getattr(self, "bg_sample_out", None),
                                tables=self.bg_tables)
            self.bg_sample_out = ob
            self.bg_batch = FrameBatch(ob.pcs, ob.z, ob.gt_depth, ob.gt_rgb, ob.labels, ob.feat_row, table)
        return self.batch

    # ---- train.py:394-474 -----------------------------------------------------------------------------------
    def train(self, iters=None, loss_terms=None, bg_loss=None):
        """loss_terms [iters, N, 4] (optional) receives the ensemble's per-object terms; bg_loss [iters] the background's
        scalar loss of each step (the reference adds it to the same scalar, train.py:463: the two problems share no
        tensor, so they are trained as two independent launch sequences)."""
        self.ens.train_frame(self.batch, iters=iters, loss_terms=loss_terms, flag_allreduce=self.flag_allreduce)
        if self.bg is not None and self.bg_batch is not None:
            self.bg.train_frame(self.bg_batch, iters=iters or self.cfg.n_iter_per_frame, loss_out=bg_loss)

    def step_frame(self, sample, iters=None, loss_terms=None):
        self.add_frame(sample)
        self.sample()
        self.train(iters, loss_terms)
