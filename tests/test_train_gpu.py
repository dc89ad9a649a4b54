"""GPU parity of the fused training step (K1 + K4) against the oracle, through the C ABI."""
import os

import numpy as np
import pytest
import torch

import openobj_oracle as oc
from openobj_b200 import layout

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PTOL = dict(rtol=1e-3, atol=2e-4)   # see tests/test_oracle_golden.py


def load(name):
    d = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def make_batch(ms, labels=None, feat=True, dev="cuda:0"):
    from openobj_b200.ensemble import FrameBatch
    lab = ms["labels"] if labels is None else labels
    return FrameBatch.from_dense(ms["pcs"].to(dev), ms["z"].to(dev), ms["gt_depth"].to(dev), ms["gt_rgb8"].to(dev),
                                 lab.to(dev), ms["gt_feat"].to(dev) if feat else None)


def check_grads(got_theta_grads, ref_grads, tol=1e-4, ref_grads_alt=None):
    """Per object and tensor: max |error| <= tol * max |reference|.

    Two evaluations of the same algorithm serve as references: the oracle in float64 and (ref_grads_alt) in float32.
    Neither alone is a fair judge: the fp32 autograd of the oracle is ~5e-4 (clip head up to 5e-2) away from fp64 on
    saturated inputs, while fp64 takes the other side of a ReLU / |.| kink about once per 3 object-steps, which changes
    a row of a weight gradient by ~1/sqrt(1200) of its size (profiles/r1b_grad_diag.txt: on such an object the kernel
    and the fp32 oracle agree to 1e-6 and both differ from fp64 by 6e-2).  An object passes if the kernel is within
    `tol` of EITHER reference; all objects but max(1, 5 %) must pass."""
    alt = ref_grads_alt or [None] * len(ref_grads)
    for name, g, r, ra in zip(layout.NAMES, layout.views(got_theta_grads.cpu()), ref_grads, alt):
        r = torch.zeros_like(g) if r is None else r.float()
        n = g.shape[0]
        scale = r.reshape(n, -1).abs().max(1).values
        err = (g - r).reshape(n, -1).abs().max(1).values
        if ra is not None:
            err = torch.minimum(err, (g - ra.float()).reshape(n, -1).abs().max(1).values)
        ok = (err <= tol * scale + 1e-7)
        n_out = int((~ok).sum())
        assert n_out <= max(1, n // 20) and bool((err <= 0.2 * scale + 1e-7).all()), (name, (err / (scale + 1e-12)).tolist())


def grads64(fc, B, pcs, z, gt_depth, rgb01, labels, gt_feat):
    d = lambda t: None if t is None else t.double()
    return oc.train_step_grads([d(p) for p in fc], d(B), d(pcs), d(z), d(gt_depth), d(rgb01), labels, d(gt_feat))


@pytest.mark.parametrize("mode", ["on", "off", "zm"])
@pytest.mark.parametrize("n_sm", [None, 2])
def test_golden_step_grads_and_loss(mode, n_sm):
    from openobj_b200.ensemble import Ensemble
    ms = load("model_step.npz")
    fc = [ms["fc%02d" % i] for i in range(18)]
    ens = Ensemble(3, rays_per_step=16, iters_per_frame=1, n_sm=n_sm)
    ens.load_stacked(fc + [ms["peB"]])
    labels = ms["labels_zm"] if mode == "zm" else ms["labels"]
    batch = make_batch(ms, labels, feat=(mode != "off"))
    ens.prepare_frame(batch)
    g, terms = ens.grads(batch, 0)
    torch.cuda.synchronize()
    total = float(ens.total_loss(terms.cpu()))
    assert abs(total - float(ms["loss_" + mode])) <= 1e-4 * abs(float(ms["loss_" + mode])) + 1e-6   # rel 1e-4 on losses
    ref = [ms["g_%s%02d" % (mode, i)] if ("g_%s%02d" % (mode, i)) in ms else None for i in range(19)]
    check_grads(g, ref, tol=2e-4)     # reference's own fp32 autograd (frozen golden)
    rt, rg = grads64(fc, ms["peB"], ms["pcs"], ms["z"], ms["gt_depth"], ms["gt_rgb8"] / 255., labels,
                     ms["gt_feat"] if mode != "off" else None)
    check_grads(g, rg)                # same algorithm in float64
    assert int(ens.flags[0]) == (2 if mode == "zm" else 0)


def params_close(v, ref, steps, lr=1e-3):
    """Parameters after k AdamW steps: rel 1e-3 / abs 2e-4 (PTOL) on all but <= 0.1 % of the elements of a tensor.
    Adam normalises every element's step to ~lr whatever the size of its gradient, so an element whose gradient is a
    near-complete cancellation (|g| at rounding level relative to its terms) gets a step whose SIGN depends on the
    summation order -- fp32 FMA chains, 3xTF32 tensor-core fragments and the reference's own CPU/CUDA kernels all differ
    there.  Such elements are bounded by the step size instead: |dp| <= 2 lr per step."""
    d = (v - ref).abs()
    bad = d > (PTOL["atol"] + PTOL["rtol"] * ref.abs())
    assert int(bad.sum()) <= max(1, int(1e-3 * ref.numel())), (int(bad.sum()), ref.numel(), float(d.max()))
    assert float(d.max()) <= 2 * lr * steps, float(d.max())


def test_golden_adamw_3_plus_2_steps():
    """3 steps with part features, then 2 without: the clip head must stay untouched (no decay) in the last two."""
    from openobj_b200.ensemble import Ensemble
    ms = load("model_step.npz")
    fc = [ms["fc%02d" % i] for i in range(18)]
    ens = Ensemble(3, rays_per_step=16, iters_per_frame=1)
    ens.load_stacked(fc + [ms["peB"]])
    b_on, b_off = make_batch(ms, feat=True), make_batch(ms, feat=False)
    losses = []
    for it in range(5):
        b = b_on if it < 3 else b_off
        lt = torch.zeros(1, 3, 4, device="cuda:0")
        ens.train_frame(b, iters=1, loss_terms=lt)
        losses.append(float(ens.total_loss(lt[0].cpu())))
        if it == 2:
            np.testing.assert_allclose(losses, ms["losses_3"].numpy(), rtol=1e-4)
            for v, i in zip(ens.stacked(), range(19)):
                params_close(v.cpu(), ms["p3_%02d" % i], 3)
    for v, i in zip(ens.stacked(), range(19)):
        params_close(v.cpu(), ms["p5_%02d" % i], 5)
    assert ens.adam_t.cpu().tolist() == [5, 5, 3]


def synth_batch(N, RAYS, seed, dev="cuda:0", feat=True, S=10):
    g = torch.Generator().manual_seed(seed)
    z = torch.sort(0.5 + 3.0 * torch.rand(N, RAYS, S, generator=g), dim=-1).values
    o = torch.randn(N, RAYS, 1, 3, generator=g) * 0.2
    d = torch.randn(N, RAYS, 1, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    pcs = (o + d * z[..., None]).float()
    gt_depth = (z[..., 6] + 0.05 * torch.randn(N, RAYS, generator=g)).float()
    rgb8 = torch.randint(0, 256, (N, RAYS, 3), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 3, (N, RAYS), generator=g, dtype=torch.uint8)
    gt_feat = torch.randn(N, RAYS, 512, generator=g) if feat else None
    return pcs, z, gt_depth, rgb8, labels, gt_feat


@pytest.mark.parametrize("N,feat", [(8, True), (5, False), (61, True)])
def test_room0_shape_steps_match_oracle(N, feat):
    """room_0 shape: R=120 rays x S=10 per object and step; 3 steps of a 3-step frame vs the oracle's
    autograd + AdamW restatement; losses rel 1e-4, parameters after 3 steps rel 1e-3."""
    from openobj_b200.ensemble import Ensemble, FrameBatch
    R, I = 120, 3
    pcs, z, gt_depth, rgb8, labels, gt_feat = synth_batch(N, R * I, seed=N, feat=feat)
    fc, B = oc.init_params(N, generator=torch.Generator().manual_seed(100 + N))
    # keep sigmoid(alpha) away from saturation: the reference forms the free probability as 1 - occ in fp32
    # (render_rays.py:38), so with |alpha| ~ 15 one ulp of occ moves T, var and the depth weight by ~10 % and no two
    # fp32 implementations (nor the reference on CPU vs CUDA) agree to 1e-4; see DESIGN.md "conditioning".
    fc[8] *= 0.3
    fc[9] *= 0.3
    ens = Ensemble(N, rays_per_step=R, iters_per_frame=I)
    ens.load_stacked(fc + [B])
    dev = "cuda:0"
    batch = FrameBatch.from_dense(pcs.to(dev), z.to(dev), gt_depth.to(dev), rgb8.to(dev), labels.to(dev),
                                  gt_feat.to(dev) if feat else None)
    # gradients of step 1 (second slice) before any update
    ens.prepare_frame(batch)
    g, terms = ens.grads(batch, 1)
    sl = slice(R, 2 * R)
    rt, rg = grads64(fc, B, pcs[:, sl], z[:, sl], gt_depth[:, sl], rgb8[:, sl] / 255., labels[:, sl],
                     gt_feat[:, sl] if feat else None)
    ref_t = torch.stack([rt.depth, rt.color, rt.opacity, rt.feat], 1).float()
    torch.testing.assert_close(terms.cpu(), ref_t, rtol=1e-4, atol=1e-6)
    _, rg32 = oc.train_step_grads(fc, B, pcs[:, sl], z[:, sl], gt_depth[:, sl], rgb8[:, sl] / 255., labels[:, sl],
                                  gt_feat[:, sl] if feat else None)
    check_grads(g, rg, ref_grads_alt=rg32)
    # three optimisation steps
    ens.reset_optimizer()
    lt = torch.zeros(I, N, 4, device=dev)
    ens.train_frame(batch, loss_terms=lt)
    # reference trajectory in float64 (the oracle's fp32 autograd is too noisy on the clip head, see check_grads)
    d = lambda t: None if t is None else t.double()
    P = [d(p) for p in fc] + [d(B)]
    M = [torch.zeros_like(p) for p in P]
    V = [torch.zeros_like(p) for p in P]
    steps = [0] * 19
    for it in range(I):
        sl = slice(it * R, (it + 1) * R)
        rt, rg = oc.train_step_grads(P[:18], P[18], d(pcs[:, sl]), d(z[:, sl]), d(gt_depth[:, sl]), d(rgb8[:, sl]) / 255.,
                                     labels[:, sl], d(gt_feat[:, sl]) if feat else None)
        got, ref = float(ens.total_loss(lt[it].cpu())), float(rt.total.detach())
        assert abs(got - ref) <= 1e-4 * abs(ref) + 1e-6, (it, got, ref)
        for i, gr in enumerate(rg):
            if gr is None:
                continue
            steps[i] += 1
            oc.adamw_step(P[i], gr, M[i], V[i], steps[i])
    for name, v, p in zip(layout.NAMES, ens.stacked(), P):
        # Adam's normalised step turns a ReLU / sign flip of one evaluation into an O(lr) move of a few elements:
        # 99.9 % of the elements within PTOL, all within 2 x 3 lr (+ slack)
        diff = (v.cpu().double() - p).abs()
        bad = diff > (PTOL["atol"] + PTOL["rtol"] * p.abs())
        assert float(bad.double().mean()) <= 1e-3 and float(diff.max()) <= 7e-3, (name, float(bad.double().mean()), float(diff.max()))


def test_no_cpu_fallback():
    from openobj_b200 import _lib
    from openobj_b200.ensemble import Ensemble
    with pytest.raises(_lib.OOError):
        Ensemble(2, device="cpu")
    with pytest.raises(_lib.OOError):
        _lib.ptr(torch.zeros(4))
