"""Parameter-block layout of the C-ABI (mirror of csrc/oo_layout.h; checked against
oo_param_offset()/oo_param_size() in tests/test_abi.py).

The reference stacks N modules into 19 tensors ``[N, ...]`` with
``functorch.combine_state_for_ensemble`` (objnerf/utils.py:55-62).  Here ONE float32 buffer
``theta[N, PSTRIDE]`` holds all of them; ``views(theta)`` returns the same 19 stacked tensors as
strided views, in ``named_parameters()`` order (18 OccupancyMap tensors, then B_layer.weight).
"""
import torch

HIDDEN, CLIP, E1, E2, NDIR = 32, 512, 87, 42, 21
PSTRIDE = 30720
PCOUNT = 30659
NAMES = [
    "in_layer.0.weight", "in_layer.0.bias", "mid1.0.0.weight", "mid1.0.0.bias",
    "cat_layer.0.weight", "cat_layer.0.bias", "mid2.0.0.weight", "mid2.0.0.bias",
    "out_alpha.weight", "out_alpha.bias", "color_linear.0.weight", "color_linear.0.bias",
    "out_color.weight", "out_color.bias", "clip_linear.0.weight", "clip_linear.0.bias",
    "out_clip.weight", "out_clip.bias", "B_layer.weight",
]
SHAPES = [(32, 87), (32,), (32, 32), (32,), (32, 119), (32,), (32, 32), (32,), (1, 32), (1,),
          (32, 74), (32,), (3, 32), (3,), (32, 74), (32,), (512, 32), (512,), (21, 3)]
OFFSETS = [0, 2784, 2816, 3840, 3872, 7680, 7712, 8736, 8768, 8800, 8804, 11172, 11204, 11300,
           11304, 13672, 13704, 30088, 30600]
# AdamW groups by which loss terms reach a tensor (trunk+alpha+PE / colour head / clip head)
GROUP = [0] * 10 + [1] * 4 + [2] * 4 + [0]


def numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


def views(theta):
    """19 stacked tensors [N, *shape] aliasing `theta` [N, PSTRIDE]."""
    assert theta.dim() == 2 and theta.shape[1] == PSTRIDE
    out = []
    for off, shp in zip(OFFSETS, SHAPES):
        out.append(theta[:, off:off + numel(shp)].view((theta.shape[0],) + tuple(shp)))
    return out


def pack(tensors, device=None):
    """Copy 19 stacked tensors [N, ...] (reference order) into a fresh theta [N, PSTRIDE]."""
    n = tensors[0].shape[0]
    theta = torch.zeros(n, PSTRIDE, dtype=torch.float32, device=device or tensors[0].device)
    for v, t in zip(views(theta), tensors):
        v.copy_(t)
    return theta


def unpack(theta):
    return [v.clone() for v in views(theta)]
