"""Full-resolution (1200x680) eval render timing of one object: rays/s and points/s of K5."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from openobj_b200 import cfg as C, utils as U, vmap as V
dev = "cuda:0"
cfg = C.room0_config()
W, H = cfg.W, cfg.H
cam = V.cameraInfo(cfg)
o = V.sceneObject(cfg, 1, torch.zeros(W, H, 3, dtype=torch.uint8, device=dev), torch.ones(W, H, device=dev),
                  torch.ones(W, H, dtype=torch.uint8, device=dev), torch.tensor([0, W - 1, 0, H - 1]), torch.eye(4), 0)
with torch.no_grad():
    o.trainer.fc_occ_map.out_alpha.bias.fill_(0.5)
bb = U.BoundingBox()
bb.R, bb.center, bb.extent = np.eye(3), np.array([0.0, 0.0, 2.5]), np.array([2.0, 1.5, 1.5])
o.bbox3dour = bb
T = np.eye(4)
res = {}
for feat in (False, True):
    o.render_2D_syn(T, None, cam.rays_dir_cache, render_part=feat, dense=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m, d, c, f = o.render_2D_syn(T, None, cam.rays_dir_cache, render_part=feat, dense=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    n_hit = int((o.render_2D_syn(T, None, cam.rays_dir_cache, render_part=False, dense=True)[0] | True).sum())
    res["feat" if feat else "nofeat"] = dict(ms=ms, visible=int(m.sum()))
# hit count via the kernel's own counter is inside render_2D_syn; recompute hit rays from geometry
dirs = cam.rays_dir_cache.reshape(-1, 3).cpu()
he = torch.tensor(bb.extent).float() / 2
tmin = (-he - torch.tensor([0, 0, -2.5])) / dirs; tmax = (he - torch.tensor([0, 0, -2.5])) / dirs
near = torch.minimum(tmin, tmax).amax(1); far = torch.maximum(tmin, tmax).amin(1)
n_hit = int(((near <= far) & (far > 0)).sum())
for k, v in res.items():
    v.update(hit_rays=n_hit, rays_per_s=n_hit / (v["ms"] * 1e-3), points_per_s=n_hit * 149 / (v["ms"] * 1e-3),
             fwd_tflops=n_hit * 149 * 2 * (13567 - 0) / (v["ms"] * 1e-3) / 1e12)
print(json.dumps(res))
