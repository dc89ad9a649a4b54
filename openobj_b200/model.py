"""Mirror of objnerf/model.py: `OccupancyMap` keeps the reference's constructor, sub-module names (hence state-dict
keys: in_layer.0.weight, mid1.0.0.weight, ...) and forward signature; forward() runs the CUDA tile on the given
embedding.  Module-level forward is inference-only (no autograd through the kernel); training goes through
`openobj_b200.ensemble.Ensemble` (fused forward + loss + backward + AdamW)."""
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
        if not self._supported() or noise_std is not None or not do_cat:
            raise NotImplementedError("CUDA OccupancyMap supports hidden=32, clip=512, e1/e2=87/42, do_cat=True, "
                                      "noise_std=None (the configuration every shipped object model uses)")
        with torch.no_grad():
            a, c, f, _ = ops.forward(self.packed(x.device), emb=x.detach()[None], want_clip=bool(do_clip))
        return (a[0] if do_alpha else None, c[0] if do_color else None, f[0] if do_clip else None)
