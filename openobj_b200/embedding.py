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
        """embedding.py:46-55: x [...,3] -> [...,129]; differentiable w.r.t. B_layer.weight (oo_embed_bwd)."""
        if self.min_deg != 0 or self.n_freqs != 6:
            raise NotImplementedError("the CUDA encoder is built for n_unidir_funcs=5 (6 bands), as every shipped config uses")
        if not x.is_cuda:
            raise RuntimeError("openobj_b200.embedding.UniDirsEmbed needs CUDA tensors (no CPU fallback)")
        B = self.B_layer.weight[None]
        theta = ops._as_theta([B], 18)
        return ops.embed_autograd(x[None], theta, self._scale_value(), B)[0]

    def _scale_value(self):
        """The `scale` buffer as a Python float (it is persistent: a checkpoint may have replaced it); read back from the
        device only when the buffer changed."""
        key = (self.scale.data_ptr(), self.scale._version)
        if getattr(self, "_scale_key", None) != key:
            object.__setattr__(self, "_scale_key", key)
            object.__setattr__(self, "_scale_float", float(self.scale))
        return self._scale_float
