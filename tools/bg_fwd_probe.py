"""Background-model FORWARD probe (9 GEMMs + encoder, hidden 128, 16 800 points): us per forward, for GEMM-engine experiments."""
import json
import sys

import torch

sys.path.insert(0, ".")
from openobj_b200.background import BackgroundModel

dev = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
g = torch.Generator(device=dev).manual_seed(3)
R, S, H = 1200, 14, 128
m = BackgroundModel(hidden=H, device=dev, rays_per_step=R, n_samp=S)
for v in m.views():
    v.copy_(torch.randn(v.shape, generator=g, device=dev) * (1.0 / max(v.shape[-1], 1)) ** 0.5 if v.dim() == 2 else torch.zeros(v.shape, device=dev))
pts = torch.randn(R * S, 3, generator=g, device=dev)
for _ in range(3):
    m.forward(pts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    m.forward(pts)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"us_per_forward": 1e3 * e0.elapsed_time(e1) / n}))
