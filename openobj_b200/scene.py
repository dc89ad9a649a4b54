"""Per-frame driver of the accelerated path: the body of objnerf/train.py:158-485 for the vmap strategy.

    add_frame   (train.py:164-276)  one launch stores the frame ONCE in the shared keyframe store (framestore.py); the
                                    keyframe policy of every visible local object updates a slot table (host integers);
                                    new objects -> ensemble rebuild; one small H2D copy carries all per-frame tables
    sample      (train.py:300-388)  K2 for all local objects in one launch, no [N,12000,512] feature copy
    train       (train.py:394-474)  iters x (K1 + K4)
    write-back  (train.py:478-485)  not needed: every object's nn.Parameters are views of the ensemble buffer

Objects are sharded by ensemble index k (order of first appearance): rank = k % world.  Ranks share two things
only: (1) the per-step zero-mask bits, OR-reduced once per frame; (2) the fact that ANY new object restarts Adam's moments
and step counts for every object (the reference restacks all models into fresh tensors on update_vmap, train.py:272-276) --
every rank sees every frame's object ids, so each applies that reset locally without communication."""
import ctypes

import numpy as np
import torch

from . import sampler, vmap
from .background import BackgroundModel
from .dist import ShardBook
from .ensemble import Ensemble, FrameBatch
from .framestore import FrameStore
from ._lib import check, lib, ptr


class _Tables:
    """Per-object integer / float tables K2 reads (slot -> store frame, slot bbox, slot part-feature frame, keyframe count,
    latest two keyframes, object id) + the new frame's pose: ONE pinned block, ONE async H2D copy per frame.  A ring of
    staging blocks lets the host run ahead of the device."""
    N_STAGE = 4

    def __init__(self, cap, kf, device):
        self.cap, self.kf, self.device = cap, kf, device
        per = kf * 6 + 4                              # slot_frame kf, bbox 4 kf, part_frame kf, n_kf 1, latest 2, obj_id 1
        self.words = cap * per + 16
        self.master = np.zeros(self.words, dtype=np.int32)
        o = 0
        def take(n, shape, dtype=np.int32):
            nonlocal o
            v = self.master[o:o + n].view(dtype).reshape(shape)
            o += n
            return v
        self.slot_frame = take(cap * kf, (cap, kf)); self.slot_frame[:] = -1
        self.slot_bbox = take(cap * kf * 4, (cap, kf, 4), np.float32)
        self.part_frame = take(cap * kf, (cap, kf))
        self.n_kf = take(cap, (cap,))
        self.latest = take(cap * 2, (cap, 2))
        self.obj_id = take(cap, (cap,))
        self.t_wc = take(16, (16,), np.float32)
        self.stage = [torch.zeros(self.words, dtype=torch.int32).pin_memory() for _ in range(self.N_STAGE)]
        self.stage_np = [s.numpy() for s in self.stage]
        self.events = [None] * self.N_STAGE
        self.k = 0
        self.dev = torch.zeros(self.words, dtype=torch.int32, device=device)
        d, o = self.dev, 0
        def dtake(n, shape, dtype=torch.int32):
            nonlocal o
            v = d[o:o + n].view(dtype).view(shape)
            o += n
            return v
        self.d_slot_frame = dtake(cap * kf, (cap, kf))
        self.d_slot_bbox = dtake(cap * kf * 4, (cap, kf, 4), torch.float32)
        self.d_part_frame = dtake(cap * kf, (cap, kf))
        self.d_n_kf = dtake(cap, (cap,))
        self.d_latest = dtake(cap * 2, (cap, 2))
        self.d_obj_id = dtake(cap, (cap,))
        self.d_t_wc = dtake(16, (16,), torch.float32)

    def upload(self):
        k = self.k
        self.k = (k + 1) % self.N_STAGE
        if self.events[k] is not None:
            self.events[k].synchronize()           # the copy that last read this staging block has run
        self.stage_np[k][:] = self.master
        self.dev.copy_(self.stage[k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev


class Scene:
    def __init__(self, cfg, rank=0, world=1, seed=0, max_frames=64, n_sm=None, flag_allreduce=None, store_capacity=32,
                 init_seed=None):
        """init_seed: None = new objects draw their initial weights from torch's global generator in creation order, exactly
        like the reference (trainer.py:43, model.py:4-6).  A sharded run (world > 1) defaults to init_seed = seed: every
        object's initial weights are then drawn from a generator keyed by (init_seed, object id), so they do not depend on
        which rank creates the object or on how many ranks there are."""
        self.cfg, self.rank, self.world, self.seed = cfg, rank, world, seed
        self.init_seed = init_seed if init_seed is not None else (seed if world > 1 else None)
        self.device = torch.device(cfg.training_device)
        self.cam = vmap.cameraInfo(cfg)
        self.obj_dict = {}            # local objects, insertion order == local ensemble index
        self.book = ShardBook(rank, world, cfg.max_n_models)
        self.global_index = self.book.global_index        # obj id -> global ensemble index k (all ranks agree)
        self.ens = None
        self.n_sm = n_sm
        self.flag_allreduce = flag_allreduce
        self.frames_seen = 0
        self.part_mode = cfg.part_mode
        self.part_table = None
        self.pw = self.ph = 0
        if self.part_mode:
            self.pw, self.ph = cfg.W // cfg.part_down, cfg.H // cfg.part_down
            self.part_table = torch.empty(max_frames, self.pw, self.ph, cfg.clip_point_feature_size,
                                          dtype=torch.float32, device=self.device)
        self.kf = cfg.keyframe_buffer_size
        self.store = FrameStore(cfg.W, cfg.H, self.device, capacity=store_capacity)
        self.tab = _Tables(max((int(cfg.max_n_models) + world - 1) // world, 1), self.kf, self.device)
        self.bank = vmap.RingBank(self.tab.cap, self.kf)      # keyframe policy of every local object, as arrays
        self._objs = []
        self.tab_bg = None
        self._stale = False
        self.batch = None
        self.sample_out = None
        self._copy_stream = None
        self._sorted_ids = (None, None, None)
        self._vis_cache = (None, None, None, None)
        self._bbox_cache = {}
        self._sample_cache = {}
        self._staged_pending, self._part_event = 0, None
        self.h2d_bytes = 0                # bytes stage_frame has moved host -> device so far
        self._host_keep = []
        self.gather_fraction = 0.6        # stage_frame gathers part-feature rows per box when the boxes cover less than this
        self._empty_bits = None
        # the separate background model (train.py:236-242): one hidden-128 model, not part of the vmap ensemble; it lives
        # on the last rank (the ensemble's round-robin starts at rank 0)
        self.scene_bg, self.bg, self.bg_batch, self.bg_sample_out = None, None, None, None
        self.bg_rank = world - 1

    # ---- host -> device staging of the NEXT frame on a copy stream (what a DataLoader with pin_memory + non_blocking
    # copies gives the reference, train.py:158-188): the 76 MB of a Replica frame travel while the current frame trains
    def stage_frame(self, sample):
        """Start the H2D copies of `sample` (pinned host tensors) on the copy stream; returns a dict to pass to
        add_frame.  Tensors already on the device pass through.  The small tensors ingestion and sampling need (image, depth,
        instance map: 9.8 MB at Replica size) go first and get their own event; the part features (66.8 MB), which only the
        training kernels read, follow STRAIGHT into their slot of the resident part-feature table with a second event, so
        the frame store write, K2 and the per-frame bookkeeping run while they are still on the wire."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        out = dict(sample)
        with torch.cuda.stream(self._copy_stream):
            for k in ("image", "depth", "obj"):
                v = sample.get(k)
                if torch.is_tensor(v) and v.device != self.device:
                    out[k] = v.to(self.device, non_blocking=True)
                    self.h2d_bytes += v.numel() * v.element_size()
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
            out["_staged"] = ev
            pf = sample.get("part_feat")
            if self.part_mode and torch.is_tensor(pf):
                slot = self.frames_seen + self._staged_pending
                if slot >= self.part_table.shape[0]:
                    raise RuntimeError("part-feature table full: raise max_frames")
                # nothing reads this slot yet: rows of the table are only reached through keyframes already ingested
                boxes = self._part_boxes(sample["bbox_dict"], pf) if self._staged_pending == 0 else None
                if boxes is not None:
                    # a rank of a sharded run only ever gathers feature rows inside its own objects' boxes (vmap.py:437-452):
                    # the kernel reads those cells straight from the pinned host tensor instead of moving all 66.8 MB
                    check(lib().oo_gather_part_rows(ctypes.c_void_p(pf.data_ptr()), ptr(self.part_table[slot]), self.pw, self.ph,
                                                    self.part_table.shape[-1], ctypes.c_void_p(boxes.ctypes.data), int(boxes.shape[0]),
                                                    ctypes.c_void_p(self._copy_stream.cuda_stream)), "oo_gather_part_rows")
                    self.h2d_bytes += int(((boxes[:, 1] - boxes[:, 0] + 1) * (boxes[:, 3] - boxes[:, 2] + 1)).sum()) * self.part_table.shape[-1] * 4
                else:
                    self.part_table[slot].copy_(pf, non_blocking=True)       # train.py:183-188, without a staging copy
                    self.h2d_bytes += pf.numel() * pf.element_size()
                ev2 = torch.cuda.Event()
                ev2.record(self._copy_stream)
                out["_part_slot"], out["_part_event"] = slot, ev2
                # the gather kernel reads the host tensor itself: it must outlive the kernel, whatever the caller does with it
                self._host_keep = [(e, t) for e, t in self._host_keep if not e.query()] + [(ev2, pf)]
                out.pop("part_feat", None)
        self._staged_pending += 1
        return out

    def _part_boxes(self, bd, pf):
        """int32 [n, 4] cell boxes (w_lo, w_hi, h_lo, h_hi, inclusive) of the part-feature rows this rank can sample from the
        frame with boxes `bd`, or None when moving the whole tensor is the better (or the only correct) choice: the background
        model samples everywhere; objects first seen in this frame are assigned with ShardBook's rule (ascending id, next
        ensemble index) without touching the book; K2 clamps pixel / 5 to the grid (oo_sample.cu gather_ray)."""
        if not (pf.device.type == "cpu" and pf.is_pinned() and pf.is_contiguous() and pf.dtype == torch.float32):
            return None
        if self.cfg.do_bg and self.rank == self.bg_rank:
            return None
        gi, k_next, mine = self.book.global_index, len(self.book.global_index), []
        for i in self._ids_of(bd):
            if i == -1 or (self.cfg.do_bg and i == 0):
                continue
            k = gi.get(i)
            if k is None:
                if k_next >= self.book.cap:
                    continue
                k, k_next = k_next, k_next + 1
            if k % self.world == self.rank:
                mine.append(i)
        if not mine or len(mine) > 192:
            return None
        bb = np.stack([self._bbox_np(i, bd[i]) for i in mine]).astype(np.float64)
        pd = float(self.cfg.part_down)
        box = np.empty((len(mine), 4), dtype=np.int32)
        box[:, 0] = np.clip(np.floor(bb[:, 0] / pd), 0, self.pw - 1)
        box[:, 1] = np.clip(np.floor(bb[:, 1] / pd), 0, self.pw - 1)
        box[:, 2] = np.clip(np.floor(bb[:, 2] / pd), 0, self.ph - 1)
        box[:, 3] = np.clip(np.floor(bb[:, 3] / pd), 0, self.ph - 1)
        cells = int(((box[:, 1] - box[:, 0] + 1) * (box[:, 3] - box[:, 2] + 1)).sum())
        if cells > self.gather_fraction * self.pw * self.ph:
            return None                   # the boxes cover most of the frame (single-GPU runs): one DMA of everything is cheaper
        return np.ascontiguousarray(box)

    def _wait_part_features(self):
        """Training kernels gather rows of the part-feature table: the newest frame's rows must have landed."""
        if self._part_event is not None:
            torch.cuda.current_stream(self.device).wait_event(self._part_event)
            self._part_event = None

    def _ids_of(self, bbox_dict):
        """Sorted instance ids of a frame (torch.unique order, train.py:191); cached while the same dict object comes back with
        the same keys (a dict mutated in place -- other ids, same length -- is noticed: its key tuple is compared)."""
        keys = tuple(bbox_dict.keys())
        if self._sorted_ids[0] is not bbox_dict or self._sorted_ids[2] != keys:
            self._sorted_ids = (bbox_dict, sorted(int(k) for k in keys), keys)
            self._vis_cache = (None, None, None, None)
        return self._sorted_ids[1]

    def _place_all(self, tab, placed, store_slot, frame_id):
        """Ring slot `slot` of local object i now holds store frame `store_slot`, for every (i, object, slot, bbox) of this
        frame at once (numpy fancy indexing: the per-object Python work is the keyframe policy alone)."""
        ii = np.fromiter((p[0] for p in placed), dtype=np.int64, count=len(placed))
        ss = np.fromiter((p[2] for p in placed), dtype=np.int64, count=len(placed))
        self.store.release_many(tab.slot_frame[ii, ss].astype(np.int64))
        self.store.acquire(store_slot, len(placed))
        tab.slot_frame[ii, ss] = store_slot
        tab.slot_bbox[ii, ss] = np.stack([self._bbox_np(p[1].obj_id, p[3]) for p in placed])
        if self.part_mode:
            tab.part_frame[ii, ss] = int(frame_id / placed[0][1].stride)      # (use_frame / stride).long(), vmap.py:438-440
        tab.n_kf[ii] = np.fromiter((p[1].ring.n_keyframes for p in placed), dtype=np.int32, count=len(placed))
        for i, o, _, _ in placed:
            lat = o.ring.latest
            if len(lat) >= 2:
                tab.latest[i, 0], tab.latest[i, 1] = lat[-2], lat[-1]

    def _place_bank(self, tab, ii, ss, boxes, store_slot, frame_id):
        """_place_all for objects whose rings live in self.bank: no per-object Python at all."""
        self.store.release_many(tab.slot_frame[ii, ss].astype(np.int64))
        self.store.acquire(store_slot, int(ii.size))
        tab.slot_frame[ii, ss] = store_slot
        tab.slot_bbox[ii, ss] = boxes
        if self.part_mode:
            self.bank.use_frame[ii, ss] = frame_id
            tab.part_frame[ii, ss] = int(frame_id / self.cfg.stride)      # (use_frame / stride).long(), vmap.py:438-440
        tab.n_kf[ii] = self.bank.n_kf[ii]
        two = self.bank.lat_len[ii] >= 2
        tab.latest[ii[two]] = self.bank.lat[ii[two]]

    def _bbox_np(self, obj_id, bbox):
        """float32 [4] of a frame's 2-D box; a box object that comes back unchanged (same tensor) is converted once."""
        c = self._bbox_cache.get(obj_id)
        if c is None or c[0] is not bbox:
            c = self._bbox_cache[obj_id] = (bbox, np.asarray(bbox, dtype=np.float32))
        return c[1]

    # ---- train.py:164-276 ---------------------------------------------------------------------------------
    def add_frame(self, sample):
        cfg, dev = self.cfg, self.device
        nb = dict(non_blocking=True)
        if "_staged" in sample:
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(sample["_staged"])
            for k in ("image", "depth", "obj", "part_feat"):     # allocated on the copy stream, consumed on this one
                if torch.is_tensor(sample.get(k)) and sample[k].is_cuda:
                    sample[k].record_stream(cur)
            self._staged_pending -= 1
        rgb, depth = sample["image"].to(dev, **nb), sample["depth"].to(dev, **nb)
        inst = sample["obj"].to(dev, **nb)
        frame_id = sample.get("frame_id", self.frames_seen)
        if self.part_mode:
            if self.frames_seen >= self.part_table.shape[0]:
                raise RuntimeError("part-feature table full: raise max_frames")
            if "_part_slot" in sample:           # stage_frame sent the features straight to their slot
                assert sample["_part_slot"] == self.frames_seen, "staged frames must be added in the order they were staged"
                self._part_event = sample["_part_event"]
            else:
                self._wait_part_features()
                self.part_table[self.frames_seen].copy_(sample["part_feat"], non_blocking=True)   # train.py:183-188
        twc = sample["T"]
        twc32 = (twc.detach().cpu() if torch.is_tensor(twc) else torch.as_tensor(np.asarray(twc))).to(torch.float32).reshape(16)
        bd = sample["bbox_dict"]
        obj_clip, obj_cap = sample.get("obj_clip"), sample.get("obj_cap")
        placed = []                   # new objects and the background: (tab, local index, object, ring slot, bbox)
        old_idx, old_ids = [], []     # local objects seen before: their keyframe policy runs as arrays (vmap.RingBank)
        new_global = False
        ids = self._ids_of(bd)
        # frames usually come with the same objects as the one before: once every id of this dict is known, only the ids this
        # rank owns (and the background) are walked, and their local indices / boxes are cached with them
        key = (id(bd), len(bd), len(self.book.global_index))
        cached = self._vis_cache[0] == key and self._vis_cache[1] is bd
        if cached:
            ids = self._vis_cache[2]
        for obj_id in ids:
            if obj_id == -1:
                continue
            if cfg.do_bg and obj_id == 0:
                # the separate background model is not part of the vmap ensemble (train.py:236-242)
                if self.rank != self.bg_rank:
                    continue
                if self.scene_bg is None:
                    if self.book.full():
                        continue       # "models full" is tested before the background branch (train.py:231-236)
                    self.scene_bg = vmap.sceneObject(cfg, 0, rgb, depth, None, bd[obj_id], twc32.view(4, 4), frame_id, shared=True,
                                                     clip_feat=_first(obj_clip, obj_id), caption_feat=_get(obj_cap, obj_id))
                    self.bg = BackgroundModel(hidden=cfg.hidden_feature_size_bg, device=self.device,
                                              rays_per_step=cfg.n_per_optim_bg,
                                              n_samp=self.scene_bg.n_bins_cam2surface + self.scene_bg.n_bins,
                                              lr=cfg.learning_rate, weight_decay=cfg.weight_decay, scale=cfg.bg_scale)
                    self.bg.adopt(self.scene_bg.trainer.fc_occ_map, self.scene_bg.trainer.pe)
                    self.tab_bg = _Tables(1, self.kf, self.device)
                    self.tab_bg.obj_id[0] = 0
                    slot = 0
                else:
                    slot = self.scene_bg.push_slot(frame_id, _first(obj_clip, obj_id), _get(obj_cap, obj_id))
                placed.append((self.tab_bg, 0, self.scene_bg, slot, bd[obj_id]))
                continue
            if cached and self._vis_cache[3] is not None:
                break                  # every other cached id is a known local object: taken from the cache below
            seen = self.book.see(obj_id)
            if seen is None:
                continue               # "models full" (train.py:231-233): the cap is global, whatever the number of ranks
            k, i, new = seen
            new_global |= new
            if i is None:
                continue               # another rank's object
            if new:
                o = self._new_object(obj_id, rgb, depth, bd[obj_id], twc32, frame_id, obj_clip, obj_cap, i)
                assert i == len(self.obj_dict)
                self.obj_dict[obj_id] = o
                self.tab.obj_id[i] = obj_id
                self._stale = True
                placed.append((self.tab, i, o, 0, bd[obj_id]))
            else:
                old_idx.append(i)
                old_ids.append(obj_id)
        self._objs = list(self.obj_dict.values())
        box_ids = None
        if cached and self._vis_cache[3] is not None:
            old_arr, old_ids, old_box, box_ids = self._vis_cache[3]
            now_ids = tuple(id(bd[i]) for i in old_ids)
            if now_ids != box_ids:         # same dict, boxes replaced in place: convert them again
                old_box, box_ids = np.stack([self._bbox_np(i, bd[i]) for i in old_ids]), now_ids
        else:
            old_arr = np.asarray(old_idx, dtype=np.int64)
            old_box = (np.stack([self._bbox_np(i, bd[i]) for i in old_ids]) if old_ids else np.zeros((0, 4), np.float32))
        if not new_global:
            mine = [i for i in self._ids_of(bd) if i in self.book.local_index or (cfg.do_bg and i == 0)]
            # the background id (0) sorts first, so the walk above handles it before it reaches the cached local objects
            fast = not placed or all(p[0] is self.tab_bg for p in placed)
            if fast and box_ids is None:
                box_ids = tuple(id(bd[i]) for i in old_ids)
            self._vis_cache = ((id(bd), len(bd), len(self.book.global_index)), bd, mine,
                               (old_arr, old_ids, old_box, box_ids) if fast else None)
        # ---- keyframe policy of the known local objects (vmap.py:166-257), all at once
        old_slots = self.bank.push_many(old_arr, frame_id) if old_arr.size else np.zeros(0, np.int64)
        if old_arr.size and (obj_clip is not None or obj_cap is not None):
            for i, obj_id in zip(old_arr.tolist(), old_ids):           # semantic features of append_keyframe (vmap.py:241-246)
                self._objs[i].add_semantic(_first(obj_clip, obj_id), _get(obj_cap, obj_id))
        # ---- the frame itself: ONE copy in the shared store (12 bytes per pixel whatever the number of objects)
        self.tab.t_wc[:] = twc32.numpy()
        g = self.store.alloc() if (placed or old_arr.size) else -1
        if old_arr.size:
            self._place_bank(self.tab, old_arr, old_slots, old_box, g, frame_id)
        for tab in (self.tab, self.tab_bg):
            mine = [(i, o, slot, bbox) for t, i, o, slot, bbox in placed if t is tab]
            if mine:
                self._place_all(tab, mine, g, frame_id)
        self.tab.upload()                                          # slot tables + the pose the store kernel reads
        if self.tab_bg is not None:
            self.tab_bg.upload()
        if g >= 0:
            self.store.write(g, rgb, depth, inst, self.tab.d_t_wc)
        self.frames_seen += 1
        if self._stale:
            self._rebuild_ensemble()
        elif new_global and self.ens is not None:
            self.ens.reset_optimizer()     # another rank's new object: the reference restacks everything (quirk 7)

    def _new_object(self, obj_id, rgb, depth, bbox, twc32, frame_id, obj_clip, obj_cap, index):
        mk = lambda: vmap.sceneObject(self.cfg, obj_id, rgb, depth, None, bbox, twc32.view(4, 4), frame_id, shared=True,   # noqa: E731
                                      clip_feat=_first(obj_clip, obj_id), caption_feat=_get(obj_cap, obj_id),
                                      bank=self.bank, bank_index=index)
        if self.init_seed is None:
            return mk()
        with torch.random.fork_rng(devices=[self.device]):
            torch.manual_seed((int(self.init_seed) * 1000003 + int(obj_id)) % (2 ** 63 - 1))
            return mk()

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
        new.reset_optimizer()
        if self.ens is not None:
            self.ens.check_explode(wait=True)
        self.ens = new
        self._stale = False

    # ---- train.py:300-388 -----------------------------------------------------------------------------------
    def _sample(self, tab, n, obj0, n_frames, n_samples, out):
        cfg = self.cfg
        cache = self._sample_cache.setdefault(id(tab), {})
        rng = sampler.CounterRng(self.seed, self.frames_seen, tab.d_obj_id[:n], tab.d_n_kf[:n], tab.d_latest[:n])
        return sampler.sample(None, None, None, None, tab.d_part_frame[:n] if self.part_mode else None, self.cam.rays_dir_cache,
                              rng, n_frames, n_samples, obj0.n_bins_cam2surface, obj0.n_bins, obj0.surface_eps, obj0.stop_eps,
                              obj0.min_bound, cfg.part_down if self.part_mode else 0, (self.pw, self.ph), out=out,
                              store=self.store, slot_frame=tab.d_slot_frame[:n], slot_bbox=tab.d_slot_bbox[:n], kf_cap=self.kf,
                              cache=cache)

    def sample(self):
        cfg = self.cfg
        objs = self._objs if self.obj_dict else []
        table = self.part_table.view(-1, self.part_table.shape[-1]) if self.part_mode else None
        if objs:
            out = self.sample_out if (self.sample_out is not None and self.sample_out.labels.shape[0] == len(objs)) else None
            out = self._sample(self.tab, len(objs), objs[0], cfg.n_iter_per_frame * cfg.win_size, cfg.n_samples_per_frame, out)
            self.sample_out = out
            self.batch = FrameBatch(out.pcs, out.z, out.gt_depth, out.gt_rgb, out.labels, out.feat_row, table)
        else:
            self.batch = None                       # this rank owns no object (yet)
        if self.scene_bg is not None:
            # train.py:300-315: the background draws n_iter_per_frame * win_size_bg keyframes x n_samples_per_frame_bg
            # pixels with 5 + 9 samples per ray
            ob = self._sample(self.tab_bg, 1, self.scene_bg, cfg.n_iter_per_frame * cfg.win_size_bg, cfg.n_samples_per_frame_bg,
                              self.bg_sample_out)
            self.bg_sample_out = ob
            self.bg_batch = FrameBatch(ob.pcs, ob.z, ob.gt_depth, ob.gt_rgb, ob.labels, ob.feat_row, table)
        return self.batch

    # ---- train.py:394-474 -----------------------------------------------------------------------------------
    def train(self, iters=None, loss_terms=None, bg_loss=None):
        """loss_terms [iters, N, 4] (optional) receives the ensemble's per-object terms; bg_loss [iters] the background's
        scalar loss of each step (the reference adds it to the same scalar, train.py:463: the two problems share no
        tensor, so they are trained as two independent launch sequences)."""
        iters = int(iters or self.cfg.n_iter_per_frame)
        self._wait_part_features()
        if self.batch is not None:
            self.ens.train_frame(self.batch, iters=iters, loss_terms=loss_terms, flag_allreduce=self.flag_allreduce)
        elif self.flag_allreduce is not None:
            # no local object: still join this frame's all-reduce of the zero-mask bits, with nothing to report
            if self._empty_bits is None or self._empty_bits.shape[0] < iters:
                self._empty_bits = torch.zeros(max(iters, self.cfg.n_iter_per_frame), 2, dtype=torch.int32, device=self.device)
            self._empty_bits.zero_()
            self.flag_allreduce(self._empty_bits[:iters])
        if self.bg is not None and self.bg_batch is not None:
            self.bg.train_frame(self.bg_batch, iters=iters, loss_out=bg_loss)

    def step_frame(self, sample, iters=None, loss_terms=None):
        self.add_frame(sample)
        self.sample()
        self.train(iters, loss_terms)

    def finish(self):
        """End of the run: look at the last frame's explode flag."""
        if self.ens is not None:
            self.ens.check_explode(wait=True)


def _first(d, k):
    """obj_clip[obj_id][0] (train.py:221), tolerant of frames that carry no semantic features."""
    if d is None or k not in d:
        return None
    return d[k][0]


def _get(d, k):
    if d is None or k not in d:
        return None
    return d[k]
