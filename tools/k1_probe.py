"""Quick device-side probe of the fused step (not the bench contract): rays/s at N objects + FFMA peak."""
import ctypes
import json
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
from openobj_b200 import _lib, layout
from openobj_b200.ensemble import Ensemble, FrameBatch


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    feat = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    R, I, S = 120, 100, 10
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    RAYS = R * I
    z = torch.sort(0.5 + 3.0 * torch.rand(N, RAYS, S, generator=g, device=dev), dim=-1).values
    o = torch.randn(N, RAYS, 1, 3, generator=g, device=dev) * 0.2
    d = torch.randn(N, RAYS, 1, 3, generator=g, device=dev)
    d = d / d.norm(dim=-1, keepdim=True)
    pcs = (o + d * z[..., None]).contiguous()
    gt_depth = z[..., 6].contiguous()
    rgb8 = torch.randint(0, 256, (N, RAYS, 3), generator=g, device=dev, dtype=torch.uint8)
    labels = torch.randint(0, 3, (N, RAYS), generator=g, device=dev, dtype=torch.uint8)
    table = torch.randn(200 * 240 * 136 // 8, 512, generator=g, device=dev) if feat else None
    rows = torch.randint(0, table.shape[0], (N, RAYS), generator=g, device=dev, dtype=torch.int32) if feat else None
    batch = FrameBatch(pcs, z, gt_depth, rgb8, labels, rows, table)
    import openobj_oracle as oc
    fc, B = oc.init_params(N, generator=torch.Generator().manual_seed(1))
    ens = Ensemble(N, rays_per_step=R, iters_per_frame=I)
    ens.load_stacked(fc + [B])
    lt = torch.zeros(I, N, 4, device=dev)
    ens.train_frame(batch, loss_terms=lt)       # warm-up frame
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in range(frames):
        ens.train_frame(batch, loss_terms=lt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = frames * I
    rays = N * R * steps
    # FFMA peak
    L = _lib.lib()
    sink = torch.zeros(4, device=dev)
    n_sm = _lib.n_sm()
    iters = 4000
    _lib.check(L.oo_fma_peak(n_sm, 100, _lib.ptr(sink), _lib.stream()))
    torch.cuda.synchronize()
    e0.record()
    _lib.check(L.oo_fma_peak(n_sm, iters, _lib.ptr(sink), _lib.stream()))
    e1.record()
    torch.cuda.synchronize()
    pk_ms = e0.elapsed_time(e1)
    flops = n_sm * 4 * 256 * iters * 16 * 8 * 2
    macs_per_ray = 10 * (63 + 2784 + 1024 + 3808 + 1024 + 32 + 2368 + 96 + (2368 if feat else 0)) + (16384 if feat else 0)
    out = dict(N=N, feat=feat, ms_per_step=ms / steps, rays_per_s=rays / (ms * 1e-3), n_cta=ens.n_cta, n_slots=ens.n_slots,
               fma_peak_tflops=flops / (pk_ms * 1e-3) / 1e12,
               k1_tflops=rays * macs_per_ray * 2 * 3 / (ms * 1e-3) / 1e12,
               last_loss=float(Ensemble.total_loss(lt[-1]).item()), first_loss=float(Ensemble.total_loss(lt[0]).item()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
