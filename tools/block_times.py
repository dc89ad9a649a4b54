"""Per-CTA start / end timestamps (globaltimer) of one k_train launch: where the launch's wall time goes."""
import ctypes, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
from openobj_b200 import _lib
from openobj_b200.ensemble import Ensemble, FrameBatch
import openobj_oracle as oc
N, R, I, S, dev = 60, 120, 100, 10, "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
RAYS = R * I
z = torch.sort(0.5 + 3.0 * torch.rand(N, RAYS, S, generator=g, device=dev), dim=-1).values
pcs = (torch.randn(N, RAYS, 1, 3, generator=g, device=dev) * 0.2 + torch.nn.functional.normalize(torch.randn(N, RAYS, 1, 3, generator=g, device=dev), dim=-1) * z[..., None]).contiguous()
rgb8 = torch.randint(0, 256, (N, RAYS, 3), generator=g, device=dev, dtype=torch.uint8)
labels = torch.randint(0, 3, (N, RAYS), generator=g, device=dev, dtype=torch.uint8)
table = torch.randn(100000, 512, generator=g, device=dev)
rows = torch.randint(0, 100000, (N, RAYS), generator=g, device=dev, dtype=torch.int32)
batch = FrameBatch(pcs, z, z[..., 6].contiguous(), rgb8, labels, rows, table)
fc, B = oc.init_params(N, generator=torch.Generator().manual_seed(1))
ens = Ensemble(N, rays_per_step=R, iters_per_frame=I)
ens.load_stacked(fc + [B])
ens.train_frame(batch)
torch.cuda.synchronize()
L = _lib.lib()._cdll
L.oo_debug_block_times.argtypes = [ctypes.c_void_p]
bt = torch.zeros(ens.n_cta, 3, dtype=torch.int64, device=dev)
L.oo_debug_block_times(ctypes.c_void_p(bt.data_ptr()))
bc = batch.to_c()
ens.prepare_frame(batch)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(5):
    e0.record(); ens.k1(bc, it, refresh_derived=(it == 0)); e1.record(); ens.k4(bc, it)
torch.cuda.synchronize()
L.oo_debug_block_times(None)
b = bt.cpu()
t0 = int(b[:, 0].min())
start, end, tiles = (b[:, 0] - t0).float() / 1e3, (b[:, 1] - t0).float() / 1e3, b[:, 2]
dur = end - start
print("event ms %.4f | first start 0, last start %.1f us, last end %.1f us" % (e0.elapsed_time(e1), float(start.max()), float(end.max())))
for n in sorted(set(tiles.tolist())):
    m = tiles == n
    print("tiles=%d: %3d CTAs, duration us min %.1f avg %.1f max %.1f, end avg %.1f max %.1f" % (n, int(m.sum()), float(dur[m].min()), float(dur[m].mean()), float(dur[m].max()), float(end[m].mean()), float(end[m].max())))
