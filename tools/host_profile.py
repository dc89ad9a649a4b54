"""Host-side profile of the per-frame path (Scene.add_frame / sample / train enqueue) at the bench's shape."""
import cProfile, pstats, sys, io
import torch
sys.path.insert(0, ".")
from openobj_b200 import cfg as C
from openobj_b200.scene import Scene
from openobj_b200.synthetic import SyntheticScene
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = C.room0_config(); cfg.do_bg = False; cfg.max_n_models = 60 * world
synth = SyntheticScene(60 * world, W=cfg.W, H=cfg.H, part_mode=True, seed=0, n_distinct=2)
sc = Scene(cfg, rank=0, world=world, seed=1, max_frames=80, flag_allreduce=(lambda b: None) if world > 1 else None)
dev = torch.device("cuda:0")
frames = [{k: (v.to(dev) if torch.is_tensor(v) and k != "T" else v) for k, v in synth.frame(f).items()} for f in range(4)]
f = 0
for _ in range(6):
    s = dict(frames[f % 4]); s["frame_id"] = 10 * f; sc.add_frame(s); sc.sample(); sc.train(iters=3); f += 1
torch.cuda.synchronize()
pr = cProfile.Profile()
for name, fn in (("add_frame", lambda s: sc.add_frame(s)), ("sample", lambda s: sc.sample()), ("train20", lambda s: sc.train(iters=20))):
    pass
pr.enable()
for _ in range(40):
    s = dict(frames[f % 4]); s["frame_id"] = 10 * f
    sc.add_frame(s); sc.sample(); sc.train(iters=20); f += 1
    torch.cuda.synchronize()
pr.disable()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(45)
print(st.getvalue()[:9000])
