"""Launch-gap probe: one frame (100 steps) of the fused step launched directly vs replayed from a CUDA graph."""
import json
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
from openobj_b200.ensemble import Ensemble, FrameBatch
import openobj_oracle as oc

N, R, I, S, dev = 60, 120, 100, 10, "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
RAYS = R * I
z = torch.sort(0.5 + 3.0 * torch.rand(N, RAYS, S, generator=g, device=dev), dim=-1).values
d = torch.nn.functional.normalize(torch.randn(N, RAYS, 1, 3, generator=g, device=dev), dim=-1)
pcs = (torch.randn(N, RAYS, 1, 3, generator=g, device=dev) * 0.2 + d * z[..., None]).contiguous()
rgb8 = torch.randint(0, 256, (N, RAYS, 3), generator=g, device=dev, dtype=torch.uint8)
labels = torch.randint(0, 3, (N, RAYS), generator=g, device=dev, dtype=torch.uint8)
table = torch.randn(100000, 512, generator=g, device=dev)
rows = torch.randint(0, 100000, (N, RAYS), generator=g, device=dev, dtype=torch.int32)
batch = FrameBatch(pcs, z, z[..., 6].contiguous(), rgb8, labels, rows, table)
fc, B = oc.init_params(N, generator=torch.Generator().manual_seed(1))
ens = Ensemble(N, rays_per_step=R, iters_per_frame=I)
ens.load_stacked(fc + [B])
lt = torch.zeros(I, N, 4, device=dev)
ens.train_frame(batch, loss_terms=lt)
torch.cuda.synchronize()


def timed(fn, n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * I)


direct = timed(lambda: ens.train_frame(batch, loss_terms=lt, prepare=False))
st = torch.cuda.Stream()
st.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(st):
    ens.train_frame(batch, loss_terms=lt, prepare=False)
torch.cuda.current_stream().wait_stream(st)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    ens.train_frame(batch, loss_terms=lt, prepare=False)
graph = timed(gr.replay)
print(json.dumps({"ms_per_step_direct": direct, "ms_per_step_graph": graph}))
