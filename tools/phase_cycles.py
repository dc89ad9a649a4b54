"""Per-phase cycle table of k_train (block 0), via the debug hook oo_debug_phase_cycles."""
import ctypes, json, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
from openobj_b200 import _lib
from openobj_b200.ensemble import Ensemble, FrameBatch
import openobj_oracle as oc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
feat = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
R, I, S, dev = 120, 100, 10, "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
RAYS = R * I
z = torch.sort(0.5 + 3.0 * torch.rand(N, RAYS, S, generator=g, device=dev), dim=-1).values
pcs = (torch.randn(N, RAYS, 1, 3, generator=g, device=dev) * 0.2 + torch.nn.functional.normalize(torch.randn(N, RAYS, 1, 3, generator=g, device=dev), dim=-1) * z[..., None]).contiguous()
rgb8 = torch.randint(0, 256, (N, RAYS, 3), generator=g, device=dev, dtype=torch.uint8)
labels = torch.randint(0, 3, (N, RAYS), generator=g, device=dev, dtype=torch.uint8)
table = torch.randn(100000, 512, generator=g, device=dev) if feat else None
rows = torch.randint(0, 100000, (N, RAYS), generator=g, device=dev, dtype=torch.int32) if feat else None
batch = FrameBatch(pcs, z, z[..., 6].contiguous(), rgb8, labels, rows, table)
fc, B = oc.init_params(N, generator=torch.Generator().manual_seed(1))
ens = Ensemble(N, rays_per_step=R, iters_per_frame=I)
ens.load_stacked(fc + [B])
ens.train_frame(batch)
torch.cuda.synchronize()
L = _lib.lib()._cdll
L.oo_debug_phase_cycles.argtypes = [ctypes.c_void_p]
n = L.oo_debug_phase_cycles(None)
cyc = torch.zeros(n, dtype=torch.int64, device=dev)
L.oo_debug_phase_cycles(ctypes.c_void_p(cyc.data_ptr()))
separate = len(sys.argv) > 4 and sys.argv[4] == "separate"
if separate:       # K1 and K4 launched apart with an event between them: no programmatic overlap, block 0's total is its own time
    bc = batch.to_c()
    ens.prepare_frame(batch)
    ev = torch.cuda.Event()
    for it in range(I):
        ens.k1(bc, it, refresh_derived=(it == 0)); ev.record(); ens.k4(bc, it); ev.record()
else:
    ens.train_frame(batch)
torch.cuda.synchronize()
L.oo_debug_phase_cycles(None)
c = cyc.cpu().tolist()
NP = n - 8
tiles = c[NP + 1]
label = {0: "load", 1: "PE fwd", 2: "in", 3: "mid1", 4: "cat", 5: "mid2", 6: "heads", 7: "out", 33: "termination", 8: "ray sums+loss",
         9: "S", 10: "v=W^T y part", 11: "v reduce", 12: "G S", 32: "cos/A/B", 13: "U+rec+M", 16: "g per point",
         17: "bwd composite", 18: "dWoc+dhp", 19: "dhc", 20: "W heads", 21: "D heads", 22: "W m2", 23: "D h3", 24: "W cat",
         25: "D h2", 26: "W m1", 27: "D h1", 28: "W in", 29: "D e1", 30: "PE bwd", 31: "bias",
         40: "out + v=W^T y", 41: "ray chain", 42: "M + dWoc + dhp + dhc"}
order = [0, 1, 2, 3, 4, 5, 6, 40, 41, 42, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31]     # device order (oo_tile.h kTrainOrder)
if len(sys.argv) > 3:
    order = [int(x) for x in sys.argv[3].split(",")]
names = ["%d %s" % (o, label.get(o, "?")) for o in order][:NP]
tot = sum(c[:NP])
print("tiles", tiles, "block cycles/tile %.0f, phases sum/tile %.0f, staging/tile %.0f, flush/tile %.0f" % (c[NP] / tiles, tot / tiles, c[NP + 2] / tiles, c[NP + 3] / tiles))
print("per launch (block 0): block %.0f = phases %.0f + staging %.0f + flush %.0f + prologue %.0f + dependency wait %.0f + tail %.0f + other %.0f cycles"
      % (c[NP] / I, tot / I, c[NP + 2] / I, c[NP + 3] / I, c[NP + 4] / I, c[NP + 5] / I, c[NP + 6] / I,
         (c[NP] - tot - c[NP + 2] - c[NP + 3] - c[NP + 4] - c[NP + 5] - c[NP + 6]) / I))
print("block 0: %.1f us per launch by globaltimer -> SM clock during the kernel %.0f MHz" % (c[NP + 7] / I / 1e3, 1e3 * c[NP] / max(c[NP + 7], 1)))
for nme, v in zip(names, c[:NP]):
    print("%-14s %8.0f cyc/tile  %5.1f%%" % (nme, v / tiles, 100.0 * v / tot))
