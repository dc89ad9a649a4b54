"""Stacked per-object ensemble living in one device buffer + the fused training step.

Replaces, for the vmap training strategy of the reference (objnerf/train.py):
  * utils.update_vmap / functorch.combine_state_for_ensemble (utils.py:55-62): `Ensemble.load_stacked`
    and `Ensemble.stacked()` -- the 19 stacked tensors are strided views of ONE buffer theta[N, PSTRIDE];
  * the per-iteration body train.py:394-474 (vmap(pe), vmap(fc), loss.step_batch_loss, backward,
    AdamW.step, zero_grad): `Ensemble.train_frame` / `train_step`, two kernel launches per iteration
    (K1 fused forward/backward, K4 fused out_clip-gradient assembly + AdamW + next-step constants).
All arithmetic happens in libopenobj_b200.so; this file only owns buffers and bookkeeping.
"""
import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib, layout
from ._lib import Batch, TrainWs, check, lib, ptr, stream


@dataclass
class FrameBatch:
    """Pre-sampled training data of one frame for all objects (train.py:371-388).

    pcs [N,RAYS,S,3] f32, z [N,RAYS,S] f32, gt_depth [N,RAYS] f32, gt_rgb [N,RAYS,3] u8,
    labels [N,RAYS] u8 (0 other / 1 this / 2 unknown), and for part_mode either
    feat_row [N,RAYS] int32 rows of feat_table [rows,512], or None (part features off)."""
    pcs: torch.Tensor
    z: torch.Tensor
    gt_depth: torch.Tensor
    gt_rgb: torch.Tensor
    labels: torch.Tensor
    feat_row: Optional[torch.Tensor] = None
    feat_table: Optional[torch.Tensor] = None

    @staticmethod
    def from_dense(pcs, z, gt_depth, gt_rgb_u8, labels, gt_feat=None):
        """The reference materialises Batch_N_gt_partfeat [N,RAYS,512] (train.py:378): use it as the table."""
        rows = table = None
        if gt_feat is not None:
            n, r = labels.shape
            table = gt_feat.reshape(n * r, gt_feat.shape[-1]).contiguous()
            rows = torch.arange(n * r, dtype=torch.int32, device=labels.device).reshape(n, r)
        return FrameBatch(pcs.contiguous(), z.contiguous(), gt_depth.contiguous(), gt_rgb_u8.contiguous(),
                          labels.contiguous(), rows, table)

    @property
    def rays_per_obj(self):
        return self.labels.shape[1]

    def to_c(self):
        assert self.gt_rgb.dtype == torch.uint8 and self.labels.dtype == torch.uint8
        assert self.pcs.dtype == torch.float32 and self.z.dtype == torch.float32
        if self.feat_row is not None:
            assert self.feat_row.dtype == torch.int32 and self.feat_table.dtype == torch.float32
            assert self.feat_table.shape[-1] == layout.CLIP
        b = Batch(ptr(self.pcs), ptr(self.z), ptr(self.gt_depth), ptr(self.gt_rgb), ptr(self.labels),
                  ptr(self.feat_row), ptr(self.feat_table), int(self.rays_per_obj))
        return b


class Ensemble:
    def __init__(self, n_obj, device="cuda:0", rays_per_step=120, iters_per_frame=100, lr=1e-3, weight_decay=0.013,
                 betas=(0.9, 0.999), eps=1e-8, scale=2.0, n_sm=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.OOError("Ensemble needs a CUDA device (no CPU fallback)")
        self.L = lib()
        self.n_obj, self.R, self.iters = int(n_obj), int(rays_per_step), int(iters_per_frame)
        self.lr, self.wd, self.betas, self.eps, self.scale = lr, weight_decay, betas, eps, scale
        self.n_sm = int(n_sm or _lib.n_sm(self.device))
        with torch.cuda.device(self.device):
            f32 = dict(dtype=torch.float32, device=self.device)
            i32 = dict(dtype=torch.int32, device=self.device)
            self.theta = torch.zeros(self.n_obj, layout.PSTRIDE, **f32)
            self.m = torch.zeros_like(self.theta)
            self.v = torch.zeros_like(self.theta)
            n_cta, n_slots = ctypes.c_int(), ctypes.c_int()
            slab_f, sched_i = ctypes.c_int64(), ctypes.c_int64()
            check(self.L.oo_train_ws_sizes(self.n_obj, self.R, self.iters, self.n_sm, ctypes.byref(n_cta),
                                           ctypes.byref(n_slots), ctypes.byref(slab_f), ctypes.byref(sched_i)),
                  "oo_train_ws_sizes")
            self.n_cta, self.n_slots = n_cta.value, n_slots.value
            self._slab = torch.zeros(slab_f.value, **f32)
            self._slot_loss = torch.zeros(self.n_slots * 4, **f32)
            self._derived = torch.zeros(self.n_obj, 1088, **f32)                   # OO_DERIVED_FLOATS per object (k_gram)
            self._clip_grad = torch.zeros(self.n_obj, 512 * 32 + 512, **f32)
            self._rayrec = torch.zeros(self.n_obj, self.R, 36, **f32)      # OO_RAYREC_FLOATS
            self._sched = torch.zeros(sched_i.value, **i32)
            self.counts = torch.zeros(self.iters, self.n_obj, 2, **i32)
            self.flags = torch.zeros(self.iters, **i32)
            self.flag_bits = torch.zeros(self.iters, 2, **i32)      # one int per zero-mask bit: what a sharded run all-reduces
            self._explode_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._explode_dev = torch.zeros(1, **i32)
            self._explode_event = None
            self._adam_scal = torch.zeros(self.iters, 3, 4, **f32)
            self.adam_t = torch.zeros(3, **i32)
            self._gram_part = torch.zeros(self.n_obj, 4, 1092, **f32)             # OO_GRAM_PART_FLOATS
            self._gram_cnt = torch.zeros(self.n_obj, **i32)
            self.ws = TrainWs(ptr(self._slab), ptr(self._slot_loss), ptr(self._derived), ptr(self._clip_grad), ptr(self._rayrec), ptr(self._sched),
                              ptr(self.counts), ptr(self.flags), ptr(self._adam_scal), ptr(self.adam_t),
                              ptr(self._gram_part), ptr(self._gram_cnt))
            check(self.L.oo_train_schedule(self.n_obj, self.R, self.n_sm, ctypes.byref(self.ws), stream()),
                  "oo_train_schedule")

    # ---- parameters -------------------------------------------------------------------------
    def stacked(self):
        """The 19 stacked tensors [N,...] of utils.update_vmap, as views of theta."""
        return layout.views(self.theta)

    def load_stacked(self, tensors):
        for v, t in zip(self.stacked(), tensors):
            v.copy_(t.to(self.device))

    def reset_optimizer(self):
        """Adam moments and step counters restart whenever the ensemble is rebuilt."""
        self.m.zero_()
        self.v.zero_()
        self.adam_t.zero_()

    # ---- per frame --------------------------------------------------------------------------
    def _iters(self, iters):
        iters = int(iters or self.iters)
        if not 1 <= iters <= self.iters:
            raise ValueError("iters=%d outside [1, iters_per_frame=%d]: the per-step tables are sized by the constructor" % (iters, self.iters))
        return iters

    def prepare_frame(self, batch: FrameBatch, iters=None, flag_allreduce=None):
        """Ray counts + zero-mask flags for every step, then the Adam schedule.  `flag_allreduce(bits)` lets a sharded run
        OR the zero-mask bits across ranks (the one cross-object coupling, render_rays.py:89-94): it receives the int32
        tensor [iters, 2] (one int per bit) and MAX-all-reduces it in place."""
        iters = self._iters(iters)
        part_on = batch.feat_row is not None
        self.check_explode()                      # of the previous frame: one look per frame
        mark = getattr(self, "mark", None) or (lambda name: None)
        with torch.cuda.device(self.device):
            bits = self.flag_bits[:iters] if flag_allreduce is not None else None
            check(self.L.oo_label_counts(ptr(batch.labels), self.n_obj, batch.rays_per_obj, self.R, iters,
                                         ptr(self.counts), ptr(self.flags), ptr(bits), stream()), "oo_label_counts")
            mark("label_counts")
            if flag_allreduce is not None:
                flag_allreduce(bits)
                mark("flag_allreduce")
            check(self.L.oo_adam_schedule(ptr(self.flags), ptr(bits), iters, int(part_on), self.lr, self.betas[0], self.betas[1],
                                          ptr(self.adam_t), ptr(self._adam_scal), stream()), "oo_adam_schedule")
            mark("adam_schedule")

    # ---- the reference's `loss > 1e5 -> print, exit(-1)` guard (render_rays.py:109-111): the update kernel raises bit
    # OO_FLAG_EXPLODE of a step's flags; the host looks once per frame, without stalling the stream
    def _post_explode(self, iters):
        with torch.cuda.device(self.device):
            torch.amax(self.flags[:iters] & 1, dim=0, keepdim=True, out=self._explode_dev)
            self._explode_host.copy_(self._explode_dev, non_blocking=True)
            self._explode_event = torch.cuda.Event()
            self._explode_event.record()

    def check_explode(self, wait=False):
        """Raises FloatingPointError if a per-object loss term of an already finished frame exceeded 1e5."""
        ev = self._explode_event
        if ev is None:
            return
        if wait:
            ev.synchronize()
        if ev.query():
            self._explode_event = None
            if int(self._explode_host[0]) != 0:
                raise FloatingPointError("loss explode (a per-object loss term > 1e5, render_rays.py:109-111)")

    def grads(self, batch: FrameBatch, it=0):
        """Gradients of the step loss w.r.t. theta (no update) + per-object loss terms [N,4]."""
        g = torch.empty_like(self.theta)
        terms = torch.empty(self.n_obj, 4, dtype=torch.float32, device=self.device)
        b = batch.to_c()
        with torch.cuda.device(self.device):
            check(self.L.oo_train_grads(ptr(self.theta), self.n_obj, ctypes.byref(b), it, self.R, self.scale,
                                        ctypes.byref(self.ws), ptr(g), ptr(terms), self.n_sm, stream()), "oo_train_grads")
        return g, terms

    def train_step(self, batch: FrameBatch, it, loss_terms=None):
        b = batch.to_c()
        with torch.cuda.device(self.device):
            check(self.L.oo_train_step(ptr(self.theta), ptr(self.m), ptr(self.v), self.n_obj, ctypes.byref(b), it, self.R,
                                       self.scale, self.lr, self.wd, self.betas[0], self.betas[1], self.eps,
                                       ctypes.byref(self.ws), ptr(loss_terms), self.n_sm, stream()), "oo_train_step")

    def train_frame(self, batch: FrameBatch, iters=None, loss_terms=None, prepare=True, flag_allreduce=None):
        """`iters` optimisation steps over the pre-sampled batch (train.py:394-474).
        loss_terms: optional [iters,N,4] output (depth, colour, opacity, feature per object)."""
        iters = self._iters(iters)
        if prepare:
            self.prepare_frame(batch, iters, flag_allreduce)
        b = batch.to_c()
        with torch.cuda.device(self.device):
            check(self.L.oo_train_frame(ptr(self.theta), ptr(self.m), ptr(self.v), self.n_obj, ctypes.byref(b), iters,
                                        self.R, self.scale, self.lr, self.wd, self.betas[0], self.betas[1], self.eps,
                                        ctypes.byref(self.ws), ptr(loss_terms), self.n_sm, stream()), "oo_train_frame")
        self._post_explode(iters)

    def k1(self, batch_c, it, refresh_derived=True):
        """K1 alone (fused encode/MLP/composite/loss/backward into the slabs).  refresh_derived=False skips the
        per-object out_clip constants kernel: valid when the previous launch on theta was `k4` (which leaves them ready)."""
        check(self.L.oo_train_k1(ptr(self.theta), self.n_obj, ctypes.byref(batch_c), it, self.R, self.scale,
                                 ctypes.byref(self.ws), int(bool(refresh_derived)), self.n_sm, stream()), "oo_train_k1")

    def k4(self, batch_c, it, loss_terms=None):
        """K4 alone (slab reduction + out_clip gradient assembly + AdamW + derived-constant refresh)."""
        check(self.L.oo_train_k4(ptr(self.theta), ptr(self.m), ptr(self.v), self.n_obj, ctypes.byref(batch_c), it, self.R,
                                 self.lr, self.wd,
                                 self.betas[0], self.betas[1], self.eps, ctypes.byref(self.ws), ptr(loss_terms), self.n_sm,
                                 stream()), "oo_train_k4")

    @staticmethod
    def total_loss(terms, color_scaling=5.0, opacity_scaling=10.0, feat_scaling=5.0):
        """loss.py:79,99,101: sum over objects of d + 5 c + 10 o + 5 f."""
        return (terms[..., 0] + color_scaling * terms[..., 1] + opacity_scaling * terms[..., 2]
                + feat_scaling * terms[..., 3]).sum(-1)
