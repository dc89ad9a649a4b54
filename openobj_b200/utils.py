"""Mirror of the hot-path entries of objnerf/utils.py: update_vmap (+ the vmap call form train.py:424-425 uses),
performance_measure, BoundingBox, enlarge_bbox."""
from time import perf_counter_ns

import numpy as np
import torch

from . import layout, ops
from .ensemble import Ensemble


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
    """What update_vmap returns as `fmodel`: calling convention fmodel(params, buffers, x) on ONE object's slice is
    what functorch would vmap; `vmap(fmodel)` below maps it over the stacked dimension in one kernel launch."""

    def __init__(self, ensemble, kind):
        self.ensemble, self.kind = ensemble, kind


_current = {}


def update_vmap(models, optimiser=None, rays_per_step=120, iters_per_frame=100):
    """utils.py:55-62: stack N modules.  Called once with the OccupancyMap list and once with the UniDirsEmbed list
    (train.py:274-275); both calls address the same Ensemble (one theta buffer).  Adam state restarts (quirk 7)."""
    n = len(models)
    dev = next(models[0].parameters()).device
    is_pe = hasattr(models[0], "B_layer")
    ens = _current.get("ens")
    if ens is None or ens.n_obj != n or (not is_pe and _current.get("fc_done")) or (is_pe and _current.get("pe_done")):
        scale = float(models[0].scale) if is_pe else 2.0
        ens = Ensemble(n, device=dev, rays_per_step=rays_per_step, iters_per_frame=iters_per_frame, scale=scale)
        _current.update(ens=ens, fc_done=False, pe_done=False)
    views = ens.stacked()
    with torch.no_grad():
        if is_pe:
            views[18].copy_(torch.stack([m.B_layer.weight.detach() for m in models]))
            ens.scale = float(models[0].scale)
            _current["pe_done"] = True
            params = (views[18],)
            buffers = (torch.stack([m.frequency_bands for m in models]), torch.stack([m.scale for m in models]))
        else:
            for i in range(18):
                views[i].copy_(torch.stack([list(m.parameters())[i].detach() for m in models]))
            _current["fc_done"] = True
            params, buffers = tuple(views[:18]), ()
    ens.params_changed()
    ens.reset_optimizer()
    return EnsembleFn(ens, "pe" if is_pe else "fc"), params, buffers


def vmap(fmodel):
    """Call form of train.py:424-425: vmap(pe_model)(pe_param, pe_buffer, pcs) -> embedding [N,...,129];
    vmap(fc_model)(fc_param, fc_buffer, embedding) -> (alpha, color, clip)."""
    ens = fmodel.ensemble

    def run(params, buffers, x):
        if fmodel.kind == "pe":
            return ops.forward(ens.theta, pcs=x, scale=ens.scale, want_clip=False, want_emb=True)[3]
        a, c, f, _ = ops.forward(ens.theta, emb=x, scale=ens.scale)
        return a, c, f
    return run
