"""K2 probe: the bench scene (60 objects, Replica frame size), Scene.sample() timed alone; run under ncu for k_sample_a/_b."""
import json
import sys

import torch

sys.path.insert(0, ".")
from openobj_b200 import cfg as C
from openobj_b200.scene import Scene
from openobj_b200.synthetic import SyntheticScene

n_obj = int(sys.argv[1]) if len(sys.argv) > 1 else 60
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
cfg = C.room0_config()
cfg.training_device = cfg.data_device = str(dev)
cfg.part_mode = True
cfg.do_bg = False
synth = SyntheticScene(n_obj, W=cfg.W, H=cfg.H, part_mode=True, seed=0, pin=True, n_distinct=2)
scene = Scene(cfg, rank=0, world=1, seed=1234, max_frames=16)
for f in range(6):
    scene.add_frame(synth.frame(f))
for _ in range(3):
    scene.sample()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    scene.sample()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
n_rays = n_obj * 12000
print(json.dumps({"objects": n_obj, "rays": n_rays, "sample_ms_median": ts[len(ts) // 2], "sample_ms_min": ts[0],
                  "GBps_at_184B_per_ray": n_rays * 184 / (ts[len(ts) // 2] * 1e-3) / 1e9}))
