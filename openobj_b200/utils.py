"""Mirror of the hot-path entries of objnerf/utils.py: update_vmap (+ the vmap call form train.py:424-425 uses),
performance_measure, BoundingBox, enlarge_bbox, and the stand-alone sampling helpers ray_box_intersection, origin_dirs_W,
stratified_bins, normal_bins_sampling (on the hot path these are fused into oo_sample_rays / oo_render_object; the
functions below keep the reference's names and signatures for callers that compose them by hand, each backed by an
element-wise kernel of libopenobj_b200.so)."""
import ctypes
from time import perf_counter_ns

import numpy as np
import torch

from . import layout, ops


class performance_measure:
    """utils.py:13-27 with a CUDA synchronise on both sides (the reference's timer only sees launches)."""

    def __init__(self, name, sync=True):
        self.name, self.sync = name, sync

    def __enter__(self):
        if self.sync and torch.cuda.is_available():
            torch.cuda.synchronize()
        self.start_time = perf_counter_ns()

    def __exit__(self, type, value, tb):
        if self.sync and torch.cuda.is_available():
            torch.cuda.synchronize()
        self.exec_time = perf_counter_ns() - self.start_time
        print(f"{self.name} excution time: {(self.exec_time)/1000000:.2f} ms")


class BoundingBox:
    def __init__(self):
        self.extent = self.R = self.center = self.points3d = None


def enlarge_bbox(bbox, scale, w, h):
    """utils.py:64-88: grow a 2-D box by `scale` and clip to the frame; None if it degenerates."""
    assert scale >= 0
    x0, y0, x1, y1 = bbox
    mx, my = int(0.5 * scale * (x1 - x0)), int(0.5 * scale * (y1 - y0))
    if mx == 0 or my == 0:
        return None
    return [int(np.clip(x0 - mx, 0, w - 1)), int(np.clip(y0 - my, 0, h - 1)),
            int(np.clip(x1 + mx, 0, w - 1)), int(np.clip(y1 + my, 0, h - 1))]


class EnsembleFn:
    """What update_vmap returns as `fmodel` (functorch's FunctionalModuleWithBuffers in the reference): `vmap(fmodel)` maps it
    over the stacked dimension in one kernel launch.  Holds the stacked parameter block the views in `params` alias."""

    def __init__(self, kind, theta, scale):
        self.kind, self.theta, self.scale = kind, theta, scale


def update_vmap(models, optimiser=None):
    """utils.py:55-62: stack N modules into `(fmodel, params, buffers)`; params require grad and join `optimiser` as a NEW
    parameter group (so Adam's moments and step counts restart, quirk 7).  Called once with the OccupancyMap list and once
    with the UniDirsEmbed list (train.py:274-275).  The stacked tensors are strided views of one block theta[N, PSTRIDE]
    per call; `vmap(fmodel)(params, buffers, x)` is differentiable (oo_forward / oo_forward_bwd / oo_embed_bwd), so the
    reference's loop body -- step_batch_loss, backward, optimiser.step -- runs unchanged on top of it."""
    n = len(models)
    dev = next(models[0].parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("openobj_b200.utils.update_vmap: the models must live on a CUDA device (no CPU fallback)")
    is_pe = hasattr(models[0], "B_layer")
    theta = torch.zeros(n, layout.PSTRIDE, dtype=torch.float32, device=dev)
    views = layout.views(theta)
    with torch.no_grad():
        if is_pe:
            views[18].copy_(torch.stack([m.B_layer.weight.detach() for m in models]))
            params = (views[18],)
            buffers = (torch.stack([m.frequency_bands for m in models]), torch.stack([m.scale for m in models]))
            scale = float(models[0].tensor_scale)
        else:
            for m in models:
                if not m._supported():
                    raise NotImplementedError("update_vmap stacks the shipped object model (hidden 32, clip 512); other widths "
                                              "are single models (OccupancyMap.forward)")
            for i in range(18):
                views[i].copy_(torch.stack([list(m.parameters())[i].detach() for m in models]))
            params, buffers, scale = tuple(views[:18]), (), None
    [p.requires_grad_() for p in params]
    if optimiser is not None:
        optimiser.add_param_group({"params": params})          # utils.py:61
    return EnsembleFn("pe" if is_pe else "fc", theta, scale), params, buffers


def vmap(fmodel):
    """Call form of train.py:424-425: vmap(pe_model)(pe_param, pe_buffer, pcs) -> embedding [N,...,129];
    vmap(fc_model)(fc_param, fc_buffer, embedding) -> (alpha, color, clip).  Differentiable w.r.t. the stacked parameters
    (and the embedding)."""

    def run(params, buffers, x):
        params = tuple(params)
        if fmodel.kind == "pe":
            theta = ops._as_theta(params, 18, fmodel.theta)
            return ops.embed_autograd(x, theta, fmodel.scale, params[0])
        theta = ops._as_theta(params, 0, fmodel.theta)
        return ops.fc_autograd(x, theta, True, params)
    return run


# ---- stand-alone sampling helpers (utils.py:309-397) -------------------------------------------------------------------
def _cuda_f32(t, device=None):
    t = torch.as_tensor(t)
    if device is not None:
        t = t.to(device)
    if not t.is_cuda:
        raise RuntimeError("openobj_b200.utils: this helper runs on the GPU and needs CUDA tensors (no CPU fallback)")
    return t.contiguous().float()


def _host3(v):
    return (ctypes.c_float * 3)(*[float(x) for x in torch.as_tensor(v).reshape(-1).tolist()])


def ray_box_intersection(origins, directions, bounds_min, bounds_max):
    """utils.py:309-319: slab test of n rays against an axis-aligned box; returns (near [n], far [n], hit [n] bool)."""
    from ._lib import check, lib, ptr, stream
    o, d = _cuda_f32(origins).reshape(-1, 3), _cuda_f32(directions).reshape(-1, 3)
    n = o.shape[0]
    near, far = torch.empty(n, device=o.device), torch.empty(n, device=o.device)
    hit = torch.empty(n, dtype=torch.uint8, device=o.device)
    with torch.cuda.device(o.device):
        check(lib().oo_ray_box(ptr(o), ptr(d), _host3(bounds_min), _host3(bounds_max), n, ptr(near), ptr(far), ptr(hit), stream()),
              "oo_ray_box")
    return near, far, hit.bool()


def origin_dirs_W(T_WC, dirs_C):
    """utils.py:324-336: rays to world coordinates.  T_WC [B,4,4]; dirs_C [B,3] or [B,n,3] -> (origins [B,3], dirs_W)."""
    from ._lib import check, lib, ptr, stream
    assert T_WC.shape[0] == dirs_C.shape[0] and T_WC.shape[1:] == (4, 4)
    T, dc = _cuda_f32(T_WC), _cuda_f32(dirs_C)
    B = T.shape[0]
    n = 1 if dc.dim() == 2 else int(dc.numel() // (3 * B))
    org, dw = torch.empty(B, 3, device=T.device), torch.empty_like(dc)
    with torch.cuda.device(T.device):
        check(lib().oo_origin_dirs(ptr(T), ptr(dc), B, n, ptr(org), ptr(dw), stream()), "oo_origin_dirs")
    return org, dw


def stratified_bins(min_depth, max_depth, n_bins, n_rays, type=torch.float32, device="cuda:0", draws=None):
    """utils.py:342-379: one uniform draw per bin between per-ray (tensor) or shared (number) depth bounds.
    `draws` [n_rays, n_bins] replaces the torch.rand call (tests feed the reference's recorded draws)."""
    from ._lib import check, lib, ptr, stream
    n_rays, n_bins = int(n_rays), int(n_bins)
    dev = torch.device(device)
    lin = torch.linspace(0, 1, n_bins + 1, dtype=torch.float32, device=dev)                       # utils.py:349 (knot table)
    u = torch.rand(n_rays, n_bins, device=dev, dtype=torch.float32) if draws is None else _cuda_f32(draws, dev).reshape(n_rays, n_bins)
    mn = _cuda_f32(min_depth, dev).reshape(-1) if torch.is_tensor(min_depth) else None
    mx = _cuda_f32(max_depth, dev).reshape(-1) if torch.is_tensor(max_depth) else None
    if mn is not None and mn.numel() == 1:
        mn = mn.expand(n_rays).contiguous()
    if mx is not None and mx.numel() == 1:
        mx = mx.expand(n_rays).contiguous()
    assert (mn is None or mn.numel() == n_rays) and (mx is None or mx.numel() == n_rays)
    z = torch.empty(n_rays, n_bins, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().oo_stratified_bins(ptr(u), ptr(mn), ptr(mx), 0.0 if mn is not None else float(min_depth),
                                       0.0 if mx is not None else float(max_depth), ptr(lin), n_rays, n_bins, ptr(z), stream()),
              "oo_stratified_bins")
    return z


def normal_bins_sampling(depth, n_bins, n_rays, delta, device="cuda:0", draws=None):
    """utils.py:382-397: N(0, delta/3) draws sorted ascending, clipped to +-delta, around `depth` [n_rays]."""
    from ._lib import check, lib, ptr, stream
    n_rays, n_bins = int(n_rays), int(n_bins)
    dev = torch.device(device)
    if draws is None:
        draws = torch.empty(n_rays, n_bins, dtype=torch.float32, device=dev).normal_(mean=0., std=delta / 3.)
    g, d = _cuda_f32(draws, dev).reshape(n_rays, n_bins), _cuda_f32(depth, dev).reshape(n_rays)
    z = torch.empty(n_rays, n_bins, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().oo_normal_bins(ptr(g), ptr(d), n_rays, n_bins, float(delta), ptr(z), stream()), "oo_normal_bins")
    return z


def ray_points(origins, dirs, z, midpoints=False, center=None):
    """origins + dirs * z (- center) for rays [n,3] and depths [n,S]; midpoints=True first replaces z by the bin midpoints
    (trainer.py:175-177).  Returns (points [n,S',3], z' [n,S'])."""
    from ._lib import check, lib, ptr, stream
    o, d, zz = _cuda_f32(origins).reshape(-1, 3), _cuda_f32(dirs).reshape(-1, 3), _cuda_f32(z)
    n, S = zz.shape
    So = S - 1 if midpoints else S
    pcs = torch.empty(n, So, 3, device=zz.device)
    zm = torch.empty(n, So, device=zz.device) if midpoints else None
    with torch.cuda.device(zz.device):
        check(lib().oo_ray_points(ptr(o), ptr(d), ptr(zz), n, S, int(bool(midpoints)), None if center is None else _host3(center),
                                  ptr(zm), ptr(pcs), stream()), "oo_ray_points")
    return pcs, (zm if midpoints else zz)
