"""tcgen05 / TMEM forward (oo_eval_points_tc) against the mma.sync tile path (oo_eval_points) and the oracle; timing at 256^3."""
import json, sys, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import openobj_oracle as oc
from openobj_b200 import layout, ops
dev = "cuda:0"
g = torch.Generator().manual_seed(3)
fc, B = oc.init_params(1, generator=g)
theta = layout.pack(fc + [B]).to(dev)
n = 50653
pts = (torch.rand(n, 3, generator=g) * 2 - 1) * 1.5
r_occ, r_col, _ = oc.eval_points([p.double() for p in fc], B.double(), pts.double(), scale=2.0)
occ_tc, col_tc = ops.eval_points_tc(theta, pts.to(dev), 2.0)
torch.cuda.synchronize()
ops.check_tc(torch.device(dev))
occ_mm, col_mm, _ = ops.eval_points(theta, pts.to(dev), 2.0, want_clip=False, tensor_core=False)
alpha_tc, _ = ops.eval_points_tc(theta, pts.to(dev), 2.0, want_alpha=True)
r_alpha = oc.ensemble_forward([p.double() for p in fc], B.double(), pts[None].double())[0][0, :, 0]
res = {
    "n_points": n,
    "occ_max_abs_err_tc": float((occ_tc.cpu().double() - r_occ).abs().max()), "occ_max_abs_err_mma_sync": float((occ_mm.cpu().double() - r_occ).abs().max()),
    "color_max_abs_err_tc": float((col_tc.cpu().double() - r_col).abs().max()), "color_max_abs_err_mma_sync": float((col_mm.cpu().double() - r_col).abs().max()),
    "alpha_max_abs_err_tc": float((alpha_tc.cpu().double() - r_alpha).abs().max()), "alpha_max_abs": float(r_alpha.abs().max()),
}
print(json.dumps(res))
dim = 256
big = (torch.rand(dim ** 3, 3, device=dev) * 2 - 1) * 1.5
for name, fn in (("tcgen05", lambda: ops.eval_points_tc(theta, big, 2.0)), ("mma_sync", lambda: ops.eval_points(theta, big, 2.0, want_clip=False, tensor_core=False))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res[name + "_ms_256cube"] = ms
    res[name + "_gpoints_per_s"] = dim ** 3 / ms / 1e6
    res[name + "_algorithmic_tflops"] = 2 * 11199 * dim ** 3 / (ms * 1e-3) / 1e12
ops.check_tc(torch.device(dev))
print(json.dumps(res))
json.dump(res, open("gpurun_out/tc_probe.json", "w"), indent=1)
