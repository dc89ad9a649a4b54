"""CI-size run of every training-path kernel for compute-sanitizer (memcheck / racecheck / synccheck):
k_train + k_update (+ k_gram, k_label_counts, k_adam_schedule), k_clipgrad + k_adamw (gradients-only entry), k_store_frame,
k_sample_main + k_sample_fix, the loss pair, the background model's k_gemm chain, k_fwdbwd + k_clip_dhp + k_clip_dw +
k_reduce_slots + k_embed_bwd_ens (the autograd surface).

    compute-sanitizer --tool racecheck python tools/sanitize_ci.py
"""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import openobj_oracle as oc
from openobj_b200 import cfg as C, ops, layout
from openobj_b200.background import BackgroundModel
from openobj_b200.ensemble import Ensemble, FrameBatch
from openobj_b200.scene import Scene
from openobj_b200.synthetic import SyntheticScene

dev = "cuda:0"
g = torch.Generator().manual_seed(0)
N, R, I, S = 3, 16, 2, 10
z = torch.sort(0.5 + 3.0 * torch.rand(N, R * I, S, generator=g), dim=-1).values
pcs = torch.randn(N, R * I, 1, 3, generator=g) * 0.2 + torch.nn.functional.normalize(torch.randn(N, R * I, 1, 3, generator=g), dim=-1) * z[..., None]
rgb8 = torch.randint(0, 256, (N, R * I, 3), generator=g, dtype=torch.uint8)
labels = torch.randint(0, 3, (N, R * I), generator=g, dtype=torch.uint8)
labels[:, 0], labels[:, 1] = 1, 0
feat = torch.randn(N, R * I, 512, generator=g)
fc, B = oc.init_params(N, generator=g)
for part in (True, False):
    ens = Ensemble(N, rays_per_step=R, iters_per_frame=I)
    ens.load_stacked(fc + [B])
    b = FrameBatch.from_dense(pcs.to(dev), z.to(dev), z[..., 5].contiguous().to(dev), rgb8.to(dev), labels.to(dev), feat.to(dev) if part else None)
    lt = torch.zeros(I, N, 4, device=dev)
    ens.train_frame(b, loss_terms=lt)
    ens.grads(b, 1)
    torch.cuda.synchronize()
    print("ensemble part=%d loss %.4f" % (part, float(ens.total_loss(lt[-1].cpu()))))
# autograd surface
theta = layout.pack(fc + [B]).to(dev)
params = [v.requires_grad_() for v in layout.views(theta)]
emb = ops.embed_autograd(pcs[:, :R].to(dev), theta, 2.0, params[18])
a, c, f = ops.fc_autograd(emb, theta, True, params[:18])
l, _, _ = ops.step_loss(a, c, z[:, :R, 5].to(dev), rgb8[:, :R].float().to(dev) / 255., labels[:, :R].to(dev), z[:, :R].to(dev), f, feat[:, :R].to(dev))
l.backward()
torch.cuda.synchronize()
print("autograd surface loss %.4f grad norm %.4f" % (float(l), float(params[0].grad.norm())))
# background model
bg = BackgroundModel(hidden=128, device=dev, rays_per_step=24, n_samp=14)
fcb, Bb = oc.init_params(1, hidden=128, generator=g)
bg.load([p[0] for p in fcb] + [Bb[0]])
zb = torch.sort(0.5 + 5.0 * torch.rand(24, 14, generator=g), dim=-1).values
pb = (torch.nn.functional.normalize(torch.randn(24, 1, 3, generator=g), dim=-1) * zb[..., None]).contiguous()
bg.train_step(pb.to(dev), zb.to(dev), zb[:, 8].contiguous().to(dev), rgb8[0, :24].contiguous().to(dev), labels[0, :24].contiguous().to(dev),
              torch.arange(24, dtype=torch.int32, device=dev), feat[0, :24].contiguous().to(dev))
torch.cuda.synchronize()
print("background loss %.4f" % float(bg.loss))
# scene: shared store + sampler + two frames of training
cfg = C.room0_config(w=100, h=60)
cfg.do_bg = True
cfg.n_iter_per_frame = 2
synth = SyntheticScene(4, W=100, H=60, part_mode=True, seed=2, n_distinct=1, with_bg=True)
sc = Scene(cfg, seed=5, max_frames=4)
for fidx in range(3):
    sc.step_frame(synth.frame(fidx))
sc.finish()
torch.cuda.synchronize()
print("scene ok: %d objects, %d frames alive in the store" % (len(sc.obj_dict), sc.store.frames_alive()))
