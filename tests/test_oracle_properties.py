"""Properties of the CPU oracle that do not need a fixture: its pieces against torch's own implementations of the same
published operations (AdamW, cosine_similarity, cumprod compositing) and the reference's cross-object rules."""
import pytest
import torch

import openobj_oracle as oc


def test_adamw_step_equals_torch_optim_adamw():
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(37, 19, generator=g)
    grads = [torch.randn(37, 19, generator=g) * s for s in (1.0, 1e-3, 1e-6, 0.5, 2.0)]
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=0.013)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for t, gr in enumerate(grads, 1):
        ref.grad = gr.clone()
        opt.step()
        oc.adamw_step(p, gr, m, v, t)
        torch.testing.assert_close(p, ref.detach(), rtol=1e-6, atol=1e-7)


def test_cosine_matches_torch_including_zero_vectors():
    g = torch.Generator().manual_seed(2)
    x, y = torch.randn(50, 512, generator=g), torch.randn(50, 512, generator=g)
    x[3] = 0.0
    y[7] = 0.0
    x[9] = 1e-12
    torch.testing.assert_close(oc.cosine(x, y), torch.nn.functional.cosine_similarity(x, y, dim=-1), rtol=1e-6, atol=1e-7)


def test_termination_is_a_sub_probability_distribution():
    g = torch.Generator().manual_seed(3)
    alpha = torch.randn(200, 10, generator=g) * 6
    occ, T = oc.termination(alpha)
    assert bool((T >= 0).all()) and bool((T.sum(-1) <= 1.0 + 1e-5).all())
    # sum_i T_i = 1 - prod_i (1 - occ_i) up to the 1e-10 the reference adds to every factor
    torch.testing.assert_close(T.sum(-1), 1.0 - torch.prod(1.0 - occ, dim=-1), rtol=1e-4, atol=1e-5)
    assert torch.equal(T[:, 0], occ[:, 0])


def _batch(n=3, r=20, s=10, c=16, seed=4):
    g = torch.Generator().manual_seed(seed)
    alpha, color = torch.randn(n, r, s, generator=g), torch.rand(n, r, s, 3, generator=g)
    z = torch.sort(torch.rand(n, r, s, generator=g) * 3 + 0.5, dim=-1).values
    gd, gc = z[..., 5].clone(), torch.rand(n, r, 3, generator=g)
    lab = torch.randint(0, 3, (n, r), generator=g, dtype=torch.uint8)
    lab[:, 0], lab[:, 1] = 1, 0
    pf, gf = torch.randn(n, r, s, c, generator=g), torch.randn(n, r, c, generator=g)
    return alpha, color, gd, gc, lab, z, gf, pf


def test_zero_mask_rule_couples_objects():
    """render_rays.py:89-94: one object without a label-1 ray zeroes the depth / colour / feature terms of ALL objects (flag
    bit 1); the opacity term (mask label != 2) is unaffected."""
    alpha, color, gd, gc, lab, z, gf, pf = _batch()
    full = oc.step_loss(alpha, color, gd, gc, lab, z, gf, pf)
    assert full.flags == 0 and bool((full.depth > 0).all()) and bool((full.feat > 0).all())
    lab2 = lab.clone()
    lab2[1][lab2[1] == 1] = 0
    hit = oc.step_loss(alpha, color, gd, gc, lab2, z, gf, pf)
    assert hit.flags & 2
    assert float(hit.depth.abs().sum()) == 0.0 and float(hit.color.abs().sum()) == 0.0 and float(hit.feat.abs().sum()) == 0.0
    assert bool((hit.opacity > 0).all())
    torch.testing.assert_close(hit.total, (hit.opacity * 10.0).sum())


def test_loss_is_a_sum_over_independent_objects():
    """loss.py:101: with no empty mask, the scalar is the sum of per-object losses -- evaluating objects separately gives the
    same terms (this is what makes sharding by object exact)."""
    alpha, color, gd, gc, lab, z, gf, pf = _batch(n=4, seed=5)
    full = oc.step_loss(alpha, color, gd, gc, lab, z, gf, pf)
    parts = [oc.step_loss(alpha[i:i + 1], color[i:i + 1], gd[i:i + 1], gc[i:i + 1], lab[i:i + 1], z[i:i + 1], gf[i:i + 1], pf[i:i + 1])
             for i in range(4)]
    torch.testing.assert_close(full.total, sum(p.total for p in parts), rtol=1e-6, atol=1e-6)
    for name in ("depth", "color", "opacity", "feat"):
        torch.testing.assert_close(getattr(full, name), torch.cat([getattr(p, name) for p in parts]), rtol=1e-6, atol=1e-7)


def test_stratified_and_normal_bins_structure():
    g = torch.Generator().manual_seed(6)
    u = torch.rand(64, 9, generator=g)
    lo, hi = torch.rand(64, generator=g), 1.0 + torch.rand(64, generator=g)
    z = oc.stratified(lo, hi, 9, u)
    edges = lo[:, None] + (hi - lo)[:, None] * torch.linspace(0, 1, 10)[None]
    assert bool((z >= edges[:, :-1] - 1e-6).all()) and bool((z <= edges[:, 1:] + 1e-6).all())
    d = 1.0 + torch.rand(64, generator=g)
    nb = oc.normal_bins(d, torch.randn(64, 9, generator=g) * 0.1, 0.1)
    assert bool((nb[:, 1:] >= nb[:, :-1]).all()) and float((nb - d[:, None]).abs().max()) <= 0.1 + 1e-6
