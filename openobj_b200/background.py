"""The background model (objnerf/train.py:300-315,379-388,447-463; vmap.py:43-47): ONE OccupancyMap of hidden width
128 + UniDirsEmbed(scale 5), trained beside the object ensemble on n_per_optim_bg = 1200 rays x 14 samples per step.
Its loss is added to the ensemble's scalar (`batch_loss += bg_loss`) and it shares the AdamW hyper-parameters; since
no tensor is shared, the sum only matters for reporting: the background trains as an independent problem.

Everything arithmetic happens in libopenobj_b200.so (csrc/oo_bg.cu); this class owns the flat parameter block, the
Adam moments and the scratch buffer."""
import ctypes

import torch

from . import _lib
from ._lib import check, lib, ptr, stream

N_TENSORS = 19
_SHAPES = lambda h: [(h, 87), (h,), (h, h), (h,), (h, h + 87), (h,), (h, h), (h,), (1, h), (1,), (h, h + 42), (h,),   # noqa: E731
                     (3, h), (3,), (h, h + 42), (h,), (512, h), (512,), (21, 3)]


class BackgroundModel:
    def __init__(self, hidden=128, device="cuda:0", rays_per_step=1200, n_samp=14, lr=1e-3, weight_decay=0.013,
                 betas=(0.9, 0.999), eps=1e-8, scale=5.0, color_scaling=5.0, opacity_scaling=10.0, feat_scaling=5.0):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.OOError("BackgroundModel needs a CUDA device (no CPU fallback)")
        self.L = lib()
        self.hidden, self.R, self.S = int(hidden), int(rays_per_step), int(n_samp)
        self.lr, self.wd, self.betas, self.eps, self.scale = lr, weight_decay, betas, eps, scale
        self.cs, self.os, self.fs = color_scaling, opacity_scaling, feat_scaling
        n = self.L.oo_bg_param_count(self.hidden)
        if n <= 0:
            raise _lib.OOError("bad hidden width %r" % hidden)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.theta = torch.zeros(n, **f32)
        self.m = torch.zeros(n, **f32)
        self.v = torch.zeros(n, **f32)
        self.adam_t = torch.zeros(3, dtype=torch.int32, device=self.device)   # per-group Adam step counters (device)
        self._ws = None
        self._ws_key = None
        self.terms = torch.zeros(4, **f32)
        self.loss = torch.zeros(1, **f32)
        self.flags = torch.zeros(1, dtype=torch.int32, device=self.device)

    # ---- parameters -------------------------------------------------------------------------
    def views(self, buf=None):
        """The 19 tensors (named_parameters() order of fc_occ_map, then pe.B_layer.weight) as views of the flat block."""
        buf = self.theta if buf is None else buf
        out = []
        for i, shp in enumerate(_SHAPES(self.hidden)):
            off, size = self.L.oo_bg_param_offset(self.hidden, i), self.L.oo_bg_param_size(self.hidden, i)
            out.append(buf[off:off + size].view(shp))
        return out

    def load(self, tensors):
        with torch.no_grad():
            for v, t in zip(self.views(), tensors):
                if t is not None:
                    v.copy_(t.detach().reshape(v.shape).to(self.device))

    def adopt(self, fc_occ_map, pe):
        """Copy a reference-surface model in and make its nn.Parameters alias the block (write-back is then free)."""
        ps = list(fc_occ_map.parameters()) + [pe.B_layer.weight]
        with torch.no_grad():
            for v, p in zip(self.views(), ps):
                v.copy_(p.detach().to(self.device))
                p.data = v

    def reset_optimizer(self):
        self.m.zero_(); self.v.zero_(); self.adam_t.zero_()

    def _scratch(self, n_pts, n_rays):
        key = (n_pts, n_rays)
        if self._ws_key != key:
            n = self.L.oo_bg_ws_floats(self.hidden, n_pts, n_rays)
            self._ws = torch.empty(n, dtype=torch.float32, device=self.device)
            self._ws_key = key
        return self._ws

    # ---- forward (train.py:449-450; also the eval / meshing query path) ------------------------
    def forward(self, pcs=None, want_clip=True, want_emb=False, emb=None):
        """pcs [...,3] -> encoder + MLP, or emb [...,129] -> the MLP alone (OccupancyMap.forward on a given embedding)."""
        src = pcs if emb is None else emb
        lead = list(src.shape[:-1])
        x = src.reshape(-1, src.shape[-1]).contiguous().float()
        n = x.shape[0]
        f32 = dict(dtype=torch.float32, device=self.device)
        alpha, color = torch.empty(n, **f32), torch.empty(n, 3, **f32)
        clip = torch.empty(n, 512, **f32) if want_clip else None
        emb_o = torch.empty(n, 129, **f32) if want_emb else None
        ws = self._scratch(n, 1)
        with torch.cuda.device(self.device):
            check(self.L.oo_bg_forward(ptr(self.theta), self.hidden, ptr(x) if emb is None else None,
                                       ptr(x) if emb is not None else None, n, float(self.scale), ptr(alpha), ptr(color),
                                       ptr(clip), ptr(emb_o), ptr(ws), stream()), "oo_bg_forward")
        return (alpha.view(lead + [1]), color.view(lead + [3]), None if clip is None else clip.view(lead + [512]),
                None if emb_o is None else emb_o.view(lead + [129]))

    def forward_bwd(self, pcs, d_alpha, d_color, d_clip=None, emb=None, want_d_emb=False):
        """Flat gradient block of all 19 tensors for upstream gradients of forward()'s outputs (what autograd does for
        `fc_occ_map(pe(x))`, train.py:449-450,472).  With `emb` (the MLP on a given embedding) also returns d_emb if asked."""
        src = pcs if emb is None else emb
        x = src.reshape(-1, src.shape[-1]).contiguous().float()
        n = x.shape[0]
        g = torch.empty_like(self.theta)
        ws = self._scratch(n, 1)
        da, dc = d_alpha.reshape(n).contiguous().float(), d_color.reshape(n, 3).contiguous().float()
        dcl = None if d_clip is None else d_clip.reshape(n, 512).contiguous().float()
        d_emb = torch.empty(n, 129, dtype=torch.float32, device=self.device) if (want_d_emb and emb is not None) else None
        with torch.cuda.device(self.device):
            check(self.L.oo_bg_forward_bwd(ptr(self.theta), self.hidden, ptr(x) if emb is None else None,
                                           ptr(x) if emb is not None else None, n, float(self.scale), ptr(da), ptr(dc), ptr(dcl),
                                           ptr(g), ptr(d_emb), ptr(ws), stream()), "oo_bg_forward_bwd")
        return (g, d_emb) if emb is not None else g

    # ---- training (train.py:447-474) ------------------------------------------------------------
    def _step(self, pcs, z, gt_depth, gt_rgb8, labels, feat_row, feat_table, grads_out):
        r, s = z.shape
        ws = self._scratch(r * s, r)
        b1, b2 = self.betas
        with torch.cuda.device(self.device):
            check(self.L.oo_bg_train_step(ptr(self.theta), ptr(self.m), ptr(self.v), self.hidden, ptr(pcs), ptr(z),
                                          ptr(gt_depth), ptr(gt_rgb8), ptr(labels), ptr(feat_row), ptr(feat_table), r, s,
                                          float(self.scale), ptr(self.adam_t), self.lr, self.wd, b1, b2, self.eps,
                                          self.cs, self.os, self.fs, ptr(ws), ptr(self.terms), ptr(self.loss),
                                          ptr(self.flags), ptr(grads_out), stream()), "oo_bg_train_step")

    def grads(self, pcs, z, gt_depth, gt_rgb8, labels, feat_row=None, feat_table=None):
        """Flat gradient of the step loss (no update), loss [1], terms [4]."""
        g = torch.zeros_like(self.theta)
        self._step(pcs.contiguous(), z.contiguous(), gt_depth.contiguous(), gt_rgb8.contiguous(), labels.contiguous(),
                   feat_row, feat_table, g)
        return g, self.loss.clone(), self.terms.clone()

    def train_step(self, pcs, z, gt_depth, gt_rgb8, labels, feat_row=None, feat_table=None):
        """pcs [R,S,3], z [R,S], gt_depth [R], gt_rgb8 [R,3] u8, labels [R] u8, feat_row [R] int32 rows of feat_table."""
        self._step(pcs, z, gt_depth, gt_rgb8, labels, feat_row, feat_table, None)

    def train_frame(self, batch, iters=100, loss_out=None):
        """`iters` steps over a pre-sampled frame batch (ensemble.FrameBatch with N = 1): step `it` uses rays
        [it*R, (it+1)*R) (train.py:447 bg_data_idx)."""
        R = self.R
        for it in range(iters):
            sl = slice(it * R, (it + 1) * R)
            self.train_step(batch.pcs[0, sl], batch.z[0, sl], batch.gt_depth[0, sl], batch.gt_rgb[0, sl], batch.labels[0, sl],
                            None if batch.feat_row is None else batch.feat_row[0, sl], batch.feat_table)
            if loss_out is not None:
                loss_out[it].copy_(self.loss[0], non_blocking=True)
