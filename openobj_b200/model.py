"""Mirror of objnerf/model.py: `OccupancyMap` keeps the reference's constructor, sub-module names (hence state-dict
keys: in_layer.0.weight, mid1.0.0.weight, ...) and forward signature; forward() runs the CUDA tile on the given
embedding and is differentiable w.r.t. the module's parameters and the embedding (oo_forward_bwd), so
`alpha, color, clip = fc_occ_map(pe(x)); loss.backward(); optimiser.step()` (train.py:449-474) works on it.  Hidden width 32
(every object model) runs the fused tile; other widths (the hidden-128 background model) the layer-by-layer kernels.
The fast path for training many objects is `openobj_b200.ensemble.Ensemble` (fused forward + loss + backward + AdamW)."""
import torch

from . import layout, ops


def init_weights(m, init_fn=torch.nn.init.xavier_normal_):
    """model.py:4-6: xavier-normal weights for every nn.Linear, biases keep PyTorch's default."""
    if type(m) == torch.nn.Linear:
        init_fn(m.weight)


def fc_block(in_f, out_f):
    return torch.nn.Sequential(torch.nn.Linear(in_f, out_f), torch.nn.ReLU(out_f))


class OccupancyMap(torch.nn.Module):
    def __init__(self, emb_size1, emb_size2, hidden_size=256, do_color=True, do_clip=True, clip_size=512,
                 hidden_layers_block=1):
        super().__init__()
        self.do_color, self.do_clip = do_color, do_clip
        self.embedding_size1, self.embedding_size2 = emb_size1, emb_size2
        self.hidden_size, self.clip_size = hidden_size, clip_size
        self.in_layer = fc_block(emb_size1, hidden_size)
        self.mid1 = torch.nn.Sequential(*[fc_block(hidden_size, hidden_size) for _ in range(hidden_layers_block)])
        self.cat_layer = fc_block(hidden_size + emb_size1, hidden_size)
        self.mid2 = torch.nn.Sequential(*[fc_block(hidden_size, hidden_size) for _ in range(hidden_layers_block)])
        self.out_alpha = torch.nn.Linear(hidden_size, 1)
        if do_color:
            self.color_linear = fc_block(emb_size2 + hidden_size, hidden_size)
            self.out_color = torch.nn.Linear(hidden_size, 3)
        if do_clip:
            self.clip_linear = fc_block(emb_size2 + hidden_size, hidden_size)
            self.out_clip = torch.nn.Linear(hidden_size, clip_size)
        self.sigmoid = torch.sigmoid

    def _supported(self):
        return (self.hidden_size == layout.HIDDEN and self.clip_size == layout.CLIP and self.embedding_size1 == layout.E1
                and self.embedding_size2 == layout.E2 and self.do_color and self.do_clip and len(self.mid1) == 1)

    def packed(self, device=None):
        """This module's parameters as a theta block [1, PSTRIDE] (PE directions left zero)."""
        ps = [p.detach()[None] for p in self.parameters()]
        theta = torch.zeros(1, layout.PSTRIDE, dtype=torch.float32, device=device or ps[0].device)
        for v, p in zip(layout.views(theta)[:18], ps):
            v.copy_(p)
        return theta

    def forward(self, x, noise_std=None, do_alpha=True, do_color=True, do_cat=True, do_clip=True):
        """model.py:61-103 on an embedding x [...,129]; returns (alpha [...,1], color [...,3], clip [...,512])."""
        if noise_std is not None or not do_cat:
            raise NotImplementedError("CUDA OccupancyMap: noise_std / do_cat=False are never used by the reference's callers")
        if not (self.clip_size == layout.CLIP and self.embedding_size1 == layout.E1 and self.embedding_size2 == layout.E2
                and self.do_color and self.do_clip and len(self.mid1) == 1 and len(self.mid2) == 1):
            raise NotImplementedError("CUDA OccupancyMap supports clip=512, e1/e2=87/42, one hidden block (every shipped config)")
        if not x.is_cuda:
            raise RuntimeError("openobj_b200.model.OccupancyMap needs CUDA tensors (no CPU fallback)")
        params = list(self.parameters())
        want_clip = bool(do_clip)
        if self.hidden_size == layout.HIDDEN:
            stacked = [p[None] for p in params]
            theta = ops._as_theta(stacked, 0)
            out = ops.fc_autograd(x[None], theta, want_clip, stacked)
            a, c = out[0][0], out[1][0]
            f = out[2][0] if want_clip else None
        else:
            from .background import BackgroundModel
            if getattr(self, "_wide", None) is None or self._wide.device != x.device:
                object.__setattr__(self, "_wide", BackgroundModel(hidden=self.hidden_size, device=x.device))
            out = ops.wide_autograd(x, self._wide, want_clip, params)
            a, c = out[0], out[1]
            f = out[2] if want_clip else None
        return (a if do_alpha else None, c if do_color else None, f)
