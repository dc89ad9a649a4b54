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


def _grads_with_relu_flips(fc1, B1, pcs1, z1, gt_depth1, rgb1, labels1, gt_feat1, inv_counts, flags, flips):
    """float64 gradients of ONE object's loss with the ReLU decision of the listed (layer, point, unit) entries inverted.
    inv_counts / flags: the per-object normalisers and the cross-object zero-mask bits of the full batch (the object is
    evaluated alone here)."""
    fc = [p.detach().clone().requires_grad_(True) for p in fc1]
    B = B1.detach().clone().requires_grad_(True)
    (W_in, b_in, W_m1, b_m1, W_cat, b_cat, W_m2, b_m2, W_a, b_a, W_cl, b_cl, W_oc, b_oc, W_cp, b_cp, W_ocp, b_ocp) = fc
    emb = oc.pe_forward(pcs1, B, 2.0)
    x = emb.reshape(1, -1, emb.shape[-1])
    xa, xb = x[..., :87], x[..., 87:]
    pres = []

    def act(pre, li):
        pres.append(pre.detach())
        m = pre.detach() > 0
        for (l, p, j) in flips:
            if l == li:
                m[0, p, j] = ~m[0, p, j]
        return pre * m

    h1 = act(oc._lin(xa, W_in, b_in), 0)
    h2 = act(oc._lin(h1, W_m1, b_m1), 1)
    h3 = act(oc._lin(torch.cat((h2, xa), -1), W_cat, b_cat), 2)
    h4 = act(oc._lin(h3, W_m2, b_m2), 3)
    alpha = oc._lin(h4, W_a, b_a) * 10.0
    hx = torch.cat((h4, xb), -1)
    color = torch.sigmoid(oc._lin(act(oc._lin(hx, W_cl, b_cl), 4), W_oc, b_oc))
    clip = oc._lin(act(oc._lin(hx, W_cp, b_cp), 5), W_ocp, b_ocp)
    R, S = z1.shape[1], z1.shape[2]
    # the object's share of loss.step_batch_loss with the batch's normalisers (sums are per object: loss.py:101)
    lab = labels1[0]
    m1, ms = (lab == 1).double(), (lab != 2).double()
    occ, T = oc.termination(alpha.reshape(1, R, S))
    d = (T * z1).sum(-1)[0]
    var = (T * (z1 - d[None, :, None]) ** 2).sum(-1)[0].detach()
    col = (T[..., None] * color.reshape(1, R, S, 3)).sum(-2)[0]
    opac = T.sum(-1)[0]
    loss = 0.0
    if not flags & 2:
        loss = loss + ((d - gt_depth1[0]).abs() * m1 / (var.sqrt() + 1e-4)).sum() * inv_counts[0]
        loss = loss + 5.0 * ((col - rgb1[0]).abs().sum(-1) * m1).sum() * inv_counts[0]
        if gt_feat1 is not None:
            rf = (T[..., None] * clip.reshape(1, R, S, -1)).sum(-2)[0]
            loss = loss + 5.0 * ((1.0 - oc.cosine(rf, gt_feat1[0])) * m1).sum() * inv_counts[0]
    if not flags & 4:
        loss = loss + 10.0 * ((opac - (lab != 0).double()).abs() * ms).sum() * inv_counts[1]
    grads = torch.autograd.grad(loss, fc + [B], allow_unused=True)
    return list(grads), pres


def relu_flip_explainer(fc, B, pcs, z, gt_depth, rgb01, labels, gt_feat, rel=2e-5, max_candidates=6):
    """For check_grads: explain(k, got) -> True iff the kernel's gradients of object k equal the float64 gradients once the
    ReLU decision of ONE or TWO hidden units whose pre-activation is within rounding of zero (|pre| < rel * rms of its layer)
    is inverted.  That is the only legitimate way two correct fp32 evaluations can differ by more than rounding here: a ReLU
    whose argument is at rounding level switches a whole row contribution (1 / 1200 of the points) on or off."""
    d = lambda t: None if t is None else t.double()
    n1 = (labels == 1).sum(1).double()
    ns = (labels != 2).sum(1).double()
    flags = (2 if bool((n1 == 0).any()) else 0) | (4 if bool((ns == 0).any()) else 0)

    def explain(k, got_k, tol):
        sl = slice(k, k + 1)
        args = ([d(p[sl]) for p in fc], d(B[sl]), d(pcs[sl]), d(z[sl]), d(gt_depth[sl]), d(rgb01[sl]), labels[sl],
                d(gt_feat[sl]) if gt_feat is not None else None, (1.0 / (n1[k] + 1e-10), 1.0 / (ns[k] + 1e-10)), flags)
        _, pres = _grads_with_relu_flips(*args, flips=[])
        cand = []
        for li, pre in enumerate(pres):
            rms = float(pre.pow(2).mean().sqrt())
            idx = (pre.abs() < rel * rms).nonzero()
            cand += [(li, int(p), int(j)) for _, p, j in idx.tolist()]
        cand = cand[:max_candidates]
        trials = [[c] for c in cand] + [[a, b] for i, a in enumerate(cand) for b in cand[i + 1:]]
        for flips in trials:
            g, _ = _grads_with_relu_flips(*args, flips=flips)
            ok = True
            for gk, rk in zip(got_k, g):
                rk = torch.zeros_like(gk) if rk is None else rk[0].float()
                if float((gk - rk).abs().max()) > tol * float(rk.abs().max()) + 1e-7:
                    ok = False
                    break
            if ok:
                return True
        return False
    return explain


def check_grads(got_theta_grads, ref_grads, tol=1e-4, ref_grads_alt=None, explain=None):
    """Per object and tensor: max |error| <= tol * max |reference|.

    Two evaluations of the same algorithm serve as references: the oracle in float64 and (ref_grads_alt) the oracle in
    float32, i.e. the reference's own arithmetic.  The fp32 evaluation is itself some distance away from float64 (sums over
    1200 points; ~1e-4 on the trunk, more on the clip head), and that distance -- measured here, per tensor, as the worst
    object's |fp32 - f64| -- is the conditioning of the problem, not an error of either implementation.  An object passes if
    the kernel is within `tol` of EITHER reference, or within max(tol, 2 x that measured fp32-vs-f64 distance) of float64.
    An object that passes none of these must be EXPLAINED: `explain` (relu_flip_explainer) has to reproduce the kernel's
    gradients of that object -- all 19 tensors, to `tol` -- from the float64 evaluation by inverting the ReLU decision of
    hidden units whose pre-activation is at rounding level.  No unexplained deviation is accepted, whatever its size."""
    alt = ref_grads_alt or [None] * len(ref_grads)
    views = layout.views(got_theta_grads.cpu())
    n = views[0].shape[0]
    failing = {}
    for name, g, r, ra in zip(layout.NAMES, views, ref_grads, alt):
        r = torch.zeros_like(g) if r is None else r.float()
        scale = r.reshape(n, -1).abs().max(1).values
        err = (g - r).reshape(n, -1).abs().max(1).values
        ok = err <= tol * scale + 1e-7
        if ra is not None:
            ra = ra.float()
            noise = float(((ra - r).reshape(n, -1).abs().max(1).values / (scale + 1e-12)).max())
            ok |= err <= max(tol, 2.0 * noise) * scale + 1e-7
            ok |= (g - ra).reshape(n, -1).abs().max(1).values <= tol * scale + 1e-7
        for k in (~ok).nonzero().flatten().tolist():
            failing.setdefault(k, []).append((name, float(err[k] / (scale[k] + 1e-12))))
    for k, what in failing.items():
        assert explain is not None and explain(k, [v[k] for v in views], tol), ("unexplained gradient deviation", k, what)
    return sorted(failing)


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
    gt_feat = ms["gt_feat"] if mode != "off" else None
    ex = relu_flip_explainer(fc, ms["peB"], ms["pcs"], ms["z"], ms["gt_depth"], ms["gt_rgb8"] / 255., labels, gt_feat)
    rt, rg = grads64(fc, ms["peB"], ms["pcs"], ms["z"], ms["gt_depth"], ms["gt_rgb8"] / 255., labels, gt_feat)
    # the reference's own fp32 autograd (frozen golden, 2e-4: its noise against float64) or the same algorithm in float64
    check_grads(g, rg, tol=2e-4, ref_grads_alt=ref, explain=ex)
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
    # Conditioning, measured on the B200 (gpurun_out/pytest_r2c.log): with the reference's init as it is (mean |alpha| 4, max
    # 35) some objects put all termination weight on one sample, var = sum T (z - d)^2 is then pure fp32 cancellation noise and
    # the depth weight 1 / (sqrt(var) + 1e-4) (render_rays.py:95-100) swings between 1e3 and 1e4: the depth term of such an
    # object differs by 10-80 % between ANY two fp32 evaluations (kernel vs oracle-f64 here; the oracle's own fp32 vs f64 in
    # tests/test_parity_r2_gpu.py::test_reference_init_and_saturated_alpha_measured_tolerance, which holds the kernel to that
    # measured conditioning on un-softened and saturated weights).  A fixed rel-1e-4 comparison is only meaningful where the
    # reference's arithmetic is well conditioned, so this test keeps sigmoid(alpha) out of saturation.
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
    check_grads(g, rg, ref_grads_alt=rg32, explain=relu_flip_explainer(
        fc, B, pcs[:, sl], z[:, sl], gt_depth[:, sl], rgb8[:, sl] / 255., labels[:, sl], gt_feat[:, sl] if feat else None))
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
