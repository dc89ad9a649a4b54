"""Background-model step probe (hidden 128, 1200 rays x 14 samples): ms per step; run under ncu for the launch list."""
import json
import sys

import torch

sys.path.insert(0, ".")
from openobj_b200.background import BackgroundModel

dev = "cuda:0"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
part = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
g = torch.Generator(device=dev).manual_seed(3)
R, S, H = 1200, 14, 128
m = BackgroundModel(hidden=H, device=dev, rays_per_step=R, n_samp=S)
for v in m.views():
    v.copy_(torch.randn(v.shape, generator=g, device=dev) * (1.0 / max(v.shape[-1], 1)) ** 0.5 if v.dim() == 2 else torch.zeros(v.shape, device=dev))
z = torch.sort(0.5 + 5.0 * torch.rand(R, S, generator=g, device=dev), dim=-1).values
d = torch.nn.functional.normalize(torch.randn(R, 1, 3, generator=g, device=dev), dim=-1)
pcs = (d * z[..., None]).contiguous()
rgb = torch.randint(0, 256, (R, 3), generator=g, device=dev, dtype=torch.uint8)
lab = torch.randint(0, 3, (R,), generator=g, device=dev, dtype=torch.uint8)
tab = torch.randn(20000, 512, generator=g, device=dev) if part else None
row = torch.randint(0, 20000, (R,), generator=g, device=dev, dtype=torch.int32) if part else None
gd = z[:, 8].contiguous()
for _ in range(3):
    m.train_step(pcs, z, gd, rgb, lab, row, tab)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    m.train_step(pcs, z, gd, rgb, lab, row, tab)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"ms_per_step": e0.elapsed_time(e1) / steps, "loss": float(m.loss)}))
