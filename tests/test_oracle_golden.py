"""Pin oracle/openobj_oracle.py against the golden vectors frozen from the unmodified
reference by oracle/make_golden.py (CPU, no GPU needed)."""
import json
import os

import numpy as np
import pytest
import torch

import openobj_oracle as oc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    d = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def params(d):
    return [d["fc%02d" % i] for i in range(18)], d["peB"]


def close(a, b, rtol=1e-4, atol=1e-6):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


# parameters after k AdamW steps: SURVEY 8(d) asks rel 1e-3.  Adam's normalised update
# lr*m/(sqrt(v)+eps) turns round-off in a near-zero gradient (|g| ~ eps) into an O(lr) step, so a
# handful of elements fed by almost-dead ReLU units move by a fraction of lr=1e-3: atol 2e-4.
PTOL = dict(rtol=1e-3, atol=2e-4)


def close_params(a, b, n_steps, lr=1e-3):
    """Parameters after n AdamW steps: PTOL for all but a few elements.  An element whose gradient is pure round-off
    (dead ReLU inputs, |g| ~ 1e-9) gets Adam's normalised step +-lr in a direction that no two fp32 evaluations agree
    on; it can be off by at most n*lr.  With 100k parameters (hidden 128) a handful of such elements always exist."""
    err = (a - b).abs()
    bad = err > PTOL["atol"] + PTOL["rtol"] * b.abs()
    assert float(bad.float().mean()) <= 1e-3, float(bad.float().mean())
    assert float(err.max()) <= 1.05 * n_steps * lr + PTOL["atol"], float(err.max())


@pytest.fixture(scope="module")
def ms():
    return load("model_step.npz")


def test_pe_and_mlp_forward(ms):
    fc, B = params(ms)
    emb = oc.pe_forward(ms["pcs"], B, 2.0)
    close(emb, ms["emb"], rtol=1e-5, atol=2e-5)
    a, c, f = oc.mlp_forward(fc, ms["emb"])
    close(a, ms["alpha"], rtol=1e-5, atol=1e-5)
    close(c, ms["color"], rtol=1e-5, atol=1e-6)
    close(f, ms["clip"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("mode", ["on", "off", "zm"])
def test_loss_and_grads(ms, mode):
    fc, B = params(ms)
    rgb = ms["gt_rgb8"] / 255.
    labels = ms["labels_zm"] if mode == "zm" else ms["labels"]
    gt_feat = None if mode == "off" else ms["gt_feat"]
    terms, grads = oc.train_step_grads(fc, B, ms["pcs"], ms["z"], ms["gt_depth"], rgb, labels, gt_feat)
    close(terms.total, ms["loss_" + mode], rtol=1e-5, atol=1e-6)
    if mode == "zm":
        assert terms.flags & 2
        assert float(terms.depth.abs().sum() + terms.color.abs().sum() + terms.feat.abs().sum()) == 0.0
    none_idx = set(ms["g_off_none"].tolist()) if mode == "off" else set()
    for i, g in enumerate(grads):
        if i in none_idx:
            assert g is None
            continue
        ref = ms["g_%s%02d" % (mode, i)]
        got = g if g is not None else torch.zeros_like(ref)
        close(got, ref, rtol=1e-4, atol=1e-7)
    if mode == "off":
        assert none_idx == {14, 15, 16, 17}


def test_adamw_three_plus_two_steps(ms):
    fc, B = params(ms)
    P = [p.clone() for p in fc] + [B.clone()]
    M = [torch.zeros_like(p) for p in P]
    V = [torch.zeros_like(p) for p in P]
    rgb = ms["gt_rgb8"] / 255.
    losses = []
    steps = [0] * len(P)
    for it in range(5):
        gt_feat = ms["gt_feat"] if it < 3 else None
        terms, grads = oc.train_step_grads(P[:18], P[18], ms["pcs"], ms["z"], ms["gt_depth"], rgb,
                                           ms["labels"], gt_feat)
        losses.append(terms.total)
        for i, g in enumerate(grads):
            if g is None:          # AdamW skips tensors without grad entirely (no decay)
                continue
            steps[i] += 1
            oc.adamw_step(P[i], g, M[i], V[i], steps[i])
        if it == 2:
            close(torch.stack(losses), ms["losses_3"], rtol=1e-5, atol=1e-6)
            for i in range(19):
                close(P[i], ms["p3_%02d" % i], **PTOL)
    for i in range(19):
        close(P[i], ms["p5_%02d" % i], **PTOL)


@pytest.mark.parametrize("name", ["sample_obj.npz", "sample_bg.npz"])
def test_sampling_bit_exact(name):
    d = load(name)
    tape = oc.SampleTape(d["kf_ids"], d["u_w"], d["u_h"], d["r_invalid"], d["r_valid"],
                         d["r_normal"], d["r_other"])
    out = oc.sample_object(d["rgbs_batch"], d["depth_batch"], d["t_wc_batch"], d["bbox"], d["rays_dir"],
                           tape, n_c2s=int(d["n_c2s"]), n_bins=int(d["n_bins"]),
                           use_frame=d["use_frame"], stride=int(d["stride"]), part_down=int(d["part_down"]))
    assert torch.equal(out["rgb"], d["gt_rgb"])
    assert torch.equal(out["depth"], d["gt_depth"])
    assert torch.equal(out["valid"], d["valid"])
    assert torch.equal(out["labels"], d["labels"])
    assert torch.equal(out["z"], d["z"])
    assert torch.equal(out["pcs"], d["pcs"])
    pf = d["global_partfeat"][out["pf"], out["pw"], out["ph"]]
    assert torch.equal(pf, d["partfeat"])


def test_render_object():
    d = load("render_obj.npz")
    fc, B = params(d)
    r = oc.render_object(fc, B, d["T_wc"].float(), d["rays_dir"], d["obb_R"].float(), d["obb_center"].float(),
                         d["obb_extent"].float(), d["jitter"], scale=2.0)
    assert torch.equal(r["mask"], d["mask"])
    m = r["mask"]
    close(r["depth"][m], d["depth"], rtol=1e-5, atol=1e-6)
    diff = (r["rgb"][m].int() - d["color"].int()).abs()
    assert int(diff.max()) <= 1            # u8 truncation: +-1 LSB
    assert float((diff > 0).float().mean()) < 0.01
    close(r["feat"][m], d["feat"], rtol=1e-4, atol=1e-5)


def test_zmerge_rule():
    W, H = 4, 3
    m0 = torch.ones(W, H, dtype=torch.bool)
    d0 = torch.full((W, H), 2.0)
    m1 = torch.zeros(W, H, dtype=torch.bool); m1[:2] = True
    d1 = torch.full((W, H), 1.0)
    m2 = torch.ones(W, H, dtype=torch.bool)
    d2 = torch.full((W, H), 1.5)
    c = [torch.full((W, H, 3), v, dtype=torch.uint8) for v in (10, 20, 30)]
    # object 0 is a "bg id": paints but does not write depth
    dbuf, cbuf, win, _ = oc.zmerge([m0, m1, m2], [d0, d1, d2], c, [True, False, False])
    assert win[:2].eq(1).all() and win[2:].eq(2).all()
    assert dbuf[:2].eq(1.0).all() and dbuf[2:].eq(1.5).all()
    assert cbuf[0, 0, 0] == 20 and cbuf[3, 0, 0] == 30


def test_keyframe_policy_fixture_sane():
    t = json.load(open(os.path.join(GOLDEN, "keyframe_policy.json")))
    assert t["keyframe_step"] == 2.5 and t["buffer"] == 20
    assert t["trace"][-1]["n_keyframes"] == 19


# ---- background model (hidden 128, S = 14, N = 1): the oracle is generic over the hidden width --------------------
@pytest.fixture(scope="module")
def bg():
    return load("bg_step.npz")


def bg_params(d):
    return [d["p%02d" % i][None] for i in range(18)], d["p18"][None]


@pytest.mark.parametrize("mode", ["on", "off"])
def test_bg_loss_and_grads(bg, mode):
    fc, B = bg_params(bg)
    scale = float(bg["scale"])
    assert int(bg["hidden"]) == 128 and scale == 5.0
    emb = oc.pe_forward(bg["pcs"][None], B, scale)
    close(emb[0, :4], bg["emb"], rtol=1e-5, atol=2e-5)
    gt_feat = bg["gt_feat"][None] if mode == "on" else None
    terms, grads = oc.train_step_grads(fc, B, bg["pcs"][None], bg["z"][None], bg["gt_depth"][None],
                                       (bg["gt_rgb8"] / 255.)[None], bg["labels"][None], gt_feat, scale=scale)
    close(terms.total, bg["loss_" + mode], rtol=1e-5, atol=1e-6)
    none_idx = set(bg["g_off_none"].tolist()) if mode == "off" else set()
    for i, g in enumerate(grads):
        if i in none_idx:
            assert g is None
            continue
        ref = bg["g_%s%02d" % (mode, i)]
        scale_ = float(ref.abs().max()) + 1e-12
        assert float((g[0] - ref).abs().max()) <= 2e-4 * scale_ + 1e-7, (i, float((g[0] - ref).abs().max()), scale_)


def test_bg_three_adamw_steps(bg):
    fc, B = bg_params(bg)
    P = [p.clone() for p in fc] + [B.clone()]
    M = [torch.zeros_like(p) for p in P]
    V = [torch.zeros_like(p) for p in P]
    t = [0] * 19
    for it in range(3):
        gt_feat = bg["gt_feat"][None] if it < 2 else None
        terms, grads = oc.train_step_grads(P[:18], P[18], bg["pcs"][None], bg["z"][None], bg["gt_depth"][None],
                                           (bg["gt_rgb8"] / 255.)[None], bg["labels"][None], gt_feat, scale=5.0)
        close(terms.total, bg["losses_3"][it], rtol=1e-4, atol=1e-6)
        for i, g in enumerate(grads):
            if g is not None:
                t[i] += 1
                oc.adamw_step(P[i], g, M[i], V[i], t[i])
    for i in range(19):
        close_params(P[i][0], bg["q3_%02d" % i], 3)


def test_bg_three_step_loss_conditioning(bg):
    """How well-conditioned are the golden 3-step losses?  Gradient noise of 1e-7 / 1e-6 (relative to each tensor's largest
    entry; fp32 round-off level) is injected before AdamW: step 0 does not move, step 1 stays within 1e-4, step 2 moves by
    more than the 1e-4 loss tolerance -- Adam's first updates are -lr * sign(g), so entries whose sign is round-off flip.
    This is why tests/test_bg_gpu.py bounds the step-2 loss by 1e-3."""
    def run(noise, seed):
        gen = torch.Generator().manual_seed(seed)
        fc, B = bg_params(bg)
        P = [p.clone() for p in fc] + [B.clone()]
        M = [torch.zeros_like(p) for p in P]
        V = [torch.zeros_like(p) for p in P]
        t = [0] * 19
        out = []
        for it in range(3):
            gt_feat = bg["gt_feat"][None] if it < 2 else None
            terms, grads = oc.train_step_grads(P[:18], P[18], bg["pcs"][None], bg["z"][None], bg["gt_depth"][None],
                                               (bg["gt_rgb8"] / 255.)[None], bg["labels"][None], gt_feat, scale=5.0)
            out.append(float(terms.total.detach()))
            for i, g in enumerate(grads):
                if g is not None:
                    t[i] += 1
                    g = g + noise * g.abs().max() * torch.randn(g.shape, generator=gen) * (g != 0)
                    oc.adamw_step(P[i], g, M[i], V[i], t[i])
        return out
    ref = [float(x) for x in bg["losses_3"]]
    worst = 0.0
    for noise in (1e-7, 1e-6):
        for seed in (0, 1):
            l = run(noise, seed)
            assert l[0] == pytest.approx(ref[0], rel=1e-6)
            assert abs(l[1] - ref[1]) <= 1e-4 * ref[1]
            assert abs(l[2] - ref[2]) <= 1e-3 * ref[2]
            worst = max(worst, abs(l[2] - ref[2]) / ref[2])
    assert worst > 1e-4      # the step-2 loss is NOT determined to 1e-4 by fp32 arithmetic


# ---- evaluation on the meshing grid: eval_grid.npz frozen from the reference's Trainer.eval_points ----
@pytest.mark.parametrize("tag", ["obj", "bg"])
def test_eval_grid_oracle(tag):
    d = load("eval_grid.npz")
    fc = [d["%s_fc%02d" % (tag, i)] for i in range(18)]
    dim = int(d[tag + "_dim"])
    pts = oc.meshing_grid(d["obb_R"], d["obb_center"], d["obb_extent"], float(d[tag + "_bound_extent"]), dim, d["obj_center"])
    assert torch.equal(pts, d[tag + "_grid"])                   # the same torch expressions: bit-identical
    occ, color, clip = oc.eval_points(fc, d[tag + "_peB"], pts, scale=float(d[tag + "_scale"]))
    close(occ, d[tag + "_occ"], rtol=1e-5, atol=1e-6)
    close(color, d[tag + "_color"], rtol=1e-5, atol=1e-6)
    close(clip[::7], d[tag + "_clip"], rtol=1e-4, atol=1e-5)
    o = d[tag + "_occ"]
    assert float((o > 0.5).float().mean()) > 0.1 and float((o < 0.5).float().mean()) > 0.01     # occupied and free points


# ---- stand-alone surface helpers: surface.npz frozen from the reference's utils / Trainer / sceneObject ----
def test_surface_helpers_oracle():
    d = load("surface.npz")
    assert torch.equal(oc.stratified(d["sb_min"], d["sb_max"], 7, d["sb_u"]), d["sb_z"])
    assert torch.equal(oc.stratified(0.0, 3.5, 10, d["sbs_u"]), d["sbs_z"])
    assert torch.equal(oc.normal_bins(d["nb_depth"], d["nb_draws"], 0.1), d["nb_z"])
    o1, w1 = oc.origin_dirs_w(d["od_T"], d["od_d1"])
    o2, w2 = oc.origin_dirs_w(d["od_T"], d["od_d2"])
    close(w1, d["od_w1"], rtol=1e-6, atol=1e-6); close(w2, d["od_w2"], rtol=1e-6, atol=1e-6)
    assert torch.equal(o1, d["od_o1"]) and torch.equal(o2, d["od_o2"])
    near, far, hit = oc.ray_box(d["rb_o"], d["rb_d"], d["rb_min"], d["rb_max"])
    assert torch.equal(near, d["rb_near"]) and torch.equal(far, d["rb_far"]) and torch.equal(hit, d["rb_hit"])
    # sample_points_bbox: depths between the clipped entry and exit + 0.2, bin midpoints, points
    zc = oc.stratified(d["spb_near"], d["spb_far"], 150, d["spb_u"])
    assert torch.equal(zc, d["spb_zcat"])
    zm = 0.5 * (zc[..., 1:] + zc[..., :-1])
    assert torch.equal(zm, d["spb_z"])
    close(d["spb_origins"][:, None, :] + d["spb_dirsW"][:, None, :] * zm[:, :, None], d["spb_pcs"], rtol=0, atol=0)
