"""Mirror of objnerf/embedding.py: `UniDirsEmbed` with the same constructor, buffers and state-dict keys
(`scale`, `B_layer.weight`; `frequency_bands` non-persistent).  forward() runs the encoder phase of the CUDA tile."""
import torch

from . import layout, ops

_DIRS = [
    0.8506508, 0, 0.5257311, 0.809017, 0.5, 0.309017, 0.5257311, 0.8506508, 0, 1, 0, 0, 0.809017, 0.5, -0.309017,
    0.8506508, 0, -0.5257311, 0.309017, 0.809017, -0.5, 0, 0.5257311, -0.8506508, 0.5, 0.309017, -0.809017, 0, 1, 0,
    -0.5257311, 0.8506508, 0, -0.309017, 0.809017, -0.5, 0, 0.5257311, 0.8506508, -0.309017, 0.809017, 0.5,
    0.309017, 0.809017, 0.5, 0.5, 0.309017, 0.809017, 0.5, -0.309017, 0.809017, 0, 0, 1, -0.5, 0.309017, 0.809017,
    -0.809017, 0.5, 0.309017, -0.809017, 0.5, -0.309017]


class UniDirsEmbed(torch.nn.Module):
    """Icosahedral-direction positional encoding (reference: embedding.py:4-55)."""

    def __init__(self, min_deg=0, max_deg=2, scale=2.):
        super().__init__()
        self.min_deg, self.max_deg = min_deg, max_deg
        self.n_freqs = max_deg - min_deg + 1
        self.tensor_scale = torch.tensor(scale, requires_grad=False)
        self.B_layer = torch.nn.Linear(3, 21, bias=False)
        self.B_layer.weight.data = torch.tensor(_DIRS).reshape(-1, 3)
        self.register_buffer("frequency_bands", 2.0 ** torch.linspace(min_deg, max_deg, self.n_freqs), persistent=False)
        self.register_buffer("scale", self.tensor_scale, persistent=True)

    def forward(self, x):
        if self.min_deg != 0 or self.n_freqs != 6:
            raise NotImplementedError("the CUDA encoder is built for n_unidir_funcs=5 (6 bands), as every shipped config uses")
        theta = torch.zeros(1, layout.PSTRIDE, dtype=torch.float32, device=x.device)
        layout.views(theta)[18].copy_(self.B_layer.weight.detach()[None])
        with torch.no_grad():
            _, _, _, emb = ops.forward(theta, pcs=x.detach()[None], scale=float(self.scale), want_clip=False, want_emb=True)
        return emb[0]
