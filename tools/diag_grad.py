import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import openobj_oracle as oc
from openobj_b200 import layout
from openobj_b200.ensemble import Ensemble, FrameBatch
from test_train_gpu import synth_batch
N, R, I = 8, 120, 3
pcs, z, gt_depth, rgb8, labels, gt_feat = synth_batch(N, R * I, seed=N, feat=True)
fc, B = oc.init_params(N, generator=torch.Generator().manual_seed(100 + N))
fc[8] *= 0.3
fc[9] *= 0.3
ens = Ensemble(N, rays_per_step=R, iters_per_frame=I)
ens.load_stacked(fc + [B])
dev = "cuda:0"
batch = FrameBatch.from_dense(pcs.to(dev), z.to(dev), gt_depth.to(dev), rgb8.to(dev), labels.to(dev), gt_feat.to(dev))
ens.prepare_frame(batch)
g, terms = ens.grads(batch, 1)
sl = slice(R, 2 * R)
rt, rg = oc.train_step_grads(fc, B, pcs[:, sl], z[:, sl], gt_depth[:, sl], rgb8[:, sl] / 255., labels[:, sl], gt_feat[:, sl])
d = lambda t: t.double()
rt64, rg64 = oc.train_step_grads([d(p) for p in fc], d(B), d(pcs[:, sl]), d(z[:, sl]), d(gt_depth[:, sl]),
                                 d(rgb8[:, sl]) / 255., labels[:, sl], d(gt_feat[:, sl]))
print("terms gpu", terms.cpu()[:2], "\nref", rt.depth[:2], rt.color[:2], rt.opacity[:2], rt.feat[:2])
for name, gg, r, r64 in zip(layout.NAMES, layout.views(g.cpu()), rg, rg64):
    sc = float(r64.abs().max())
    e_gpu = (gg.double() - r64)
    e_cpu = (r.double() - r64)
    per_obj = e_gpu.reshape(N, -1).abs().max(1).values / sc
    print("%-22s scale %.2e | gpu-vs-f64 max %.2e l2rel %.2e | cpu32-vs-f64 max %.2e l2rel %.2e | per-obj %s" % (
        name, sc, float(e_gpu.abs().max()) / sc, float(e_gpu.norm() / r64.norm()),
        float(e_cpu.abs().max()) / sc, float(e_cpu.norm() / r64.norm()),
        " ".join("%.0e" % v for v in per_obj.tolist())))
