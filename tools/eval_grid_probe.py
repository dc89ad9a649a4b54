"""Trainer.eval_grid probe: grid_dim^3 query points of one object model through oo_make_grid +
oo_eval_points; device time, points/s and algorithmic TFLOP/s (11 199 MAC per point without the 512-wide part head)."""
import json
import sys
import types

import numpy as np
import torch

sys.path.insert(0, ".")
from openobj_b200 import cfg as C, ops, trainer as T

dev = "cuda:0"
dims = [int(a) for a in sys.argv[1:]] or [128, 256]
cfg = C.room0_config()
cfg.obj_id = 1
cfg.training_device = dev
torch.manual_seed(0)
tr = T.Trainer(cfg)
bound = types.SimpleNamespace(R=np.eye(3), center=np.array([0.2, 0.1, 2.0]), extent=np.array([1.5, 1.0, 0.8]))
for dim in dims:
    for _ in range(2):
        tr.eval_grid(bound, None, grid_dim=dim)
    scale, trf = tr.grid_transform(bound)
    theta = tr.packed(dev)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    grid = ops.make_grid(dim, scale, trf, device=dev)
    e[1].record()
    occ, color, _ = ops.eval_points(theta, grid, scale=2.0, want_clip=False)
    e[2].record()
    torch.cuda.synchronize()
    n = dim ** 3
    t_grid, t_eval = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    print(json.dumps({"grid_dim": dim, "points": n, "make_grid_ms": t_grid, "eval_ms": t_eval,
                      "points_per_s": n / (t_eval * 1e-3), "tflops_alg": n * 11199 * 2 / (t_eval * 1e-3) / 1e12,
                      "grid_GBps": n * 12 / (t_grid * 1e-3) / 1e9, "occupied_frac": float((occ > 0.5).float().mean())}))
