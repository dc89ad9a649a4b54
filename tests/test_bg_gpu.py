"""Background model (train.py:447-474 with scene_bg): the layer-by-layer CUDA path against the golden
vectors frozen from the reference and against the oracle at the reference's full step size (1200 rays x 14 samples)."""
import os

import numpy as np
import pytest
import torch

import openobj_oracle as oc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def load(name):
    d = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def make(bg, **kw):
    from openobj_b200.background import BackgroundModel
    model = BackgroundModel(hidden=128, device=DEV, rays_per_step=bg["z"].shape[0], n_samp=bg["z"].shape[1], **kw)
    model.load([bg["p%02d" % i] for i in range(19)])
    return model


def dev_batch(bg, part):
    pcs, z = bg["pcs"].to(DEV).contiguous(), bg["z"].to(DEV).contiguous()
    gd, rgb, lab = bg["gt_depth"].to(DEV), bg["gt_rgb8"].to(DEV).contiguous(), bg["labels"].to(DEV)
    rows = table = None
    if part:
        table = bg["gt_feat"].to(DEV).contiguous()
        rows = torch.arange(table.shape[0], dtype=torch.int32, device=DEV)
    return pcs, z, gd, rgb, lab, rows, table


def test_bg_forward_matches_reference():
    bg = load("bg_step.npz")
    model = make(bg)
    a, c, f, e = model.forward(bg["pcs"].to(DEV), want_clip=True, want_emb=True)
    torch.testing.assert_close(e.cpu()[:4], bg["emb"], rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(a.cpu(), bg["alpha"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(c.cpu(), bg["color"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(f.cpu()[:2], bg["clip"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("mode", ["on", "off"])
def test_bg_loss_and_grads_match_reference(mode):
    bg = load("bg_step.npz")
    model = make(bg)
    g, loss, terms = model.grads(*dev_batch(bg, mode == "on"))
    ref = float(bg["loss_" + mode])
    assert abs(float(loss) - ref) <= 1e-4 * abs(ref) + 1e-6, (float(loss), ref)      # losses: rel 1e-4 (north_star)
    none_idx = set(bg["g_off_none"].tolist()) if mode == "off" else set()
    for i, v in enumerate(model.views(g)):
        if i in none_idx:
            assert float(v.abs().max()) == 0.0
            continue
        r = bg["g_%s%02d" % (mode, i)]
        err, sc = float((v.cpu() - r).abs().max()), float(r.abs().max()) + 1e-12
        assert err <= 2e-4 * sc + 1e-7, (i, err, sc)       # same bound the oracle itself is held to (test_oracle_golden)


def test_bg_three_steps_match_reference():
    bg = load("bg_step.npz")
    model = make(bg)
    losses = []
    for it in range(3):
        model.train_step(*dev_batch(bg, it < 2))
        losses.append(float(model.loss))
    # losses: rel 1e-4 (north_star) at steps 0 and 1.  Step 2 is evaluated on parameters that went through two AdamW
    # updates, and Adam's first updates are -lr * sign(g): with 24 rays, gradient noise of ONE fp32 ulp (1e-7 relative)
    # already moves the step-2 loss by 1.4e-4 and 1e-6 by up to 2e-4 (tests/test_oracle_golden.py::
    # test_bg_three_step_loss_conditioning measures this on the oracle), so the step-2 bound is the conditioning, 1e-3.
    for it in range(3):
        ref = float(bg["losses_3"][it])
        assert abs(losses[it] - ref) <= (1e-4 if it < 2 else 1e-3) * abs(ref) + 1e-6, (it, losses[it], ref)
    for i, v in enumerate(model.views()):
        # PTOL for all but <= 1 % of the elements (with only 24 rays many gradient entries are pure round-off, and those get Adam's +-lr step in an arbitrary
        # direction; see tests/test_oracle_golden.py::close_params), never more than 3 lr
        err, ref = (v.cpu() - bg["q3_%02d" % i]).abs(), bg["q3_%02d" % i]
        bad = err > 2e-4 + 1e-3 * ref.abs()
        assert float(bad.float().mean()) <= 1e-2 and float(err.max()) <= 3.4e-3, (i, float(bad.float().mean()), float(err.max()))


def test_bg_full_size_step_against_oracle():
    """room_0.json sizes: 1200 rays x 14 samples, hidden 128, part features on; oracle evaluated in float64."""
    g = torch.Generator().manual_seed(5)
    R, S = 1200, 14
    fc, B = oc.init_params(1, hidden=128, generator=g)
    fc[8] *= 0.3
    fc[9] *= 0.3
    z = torch.sort(0.5 + 5.0 * torch.rand(R, S, generator=g), dim=-1).values
    d = torch.nn.functional.normalize(torch.randn(R, 1, 3, generator=g), dim=-1)
    pcs = (torch.randn(R, 1, 3, generator=g) * 0.3 + d * z[..., None]).float()
    gt_depth = z[:, 8].clone()
    rgb8 = torch.randint(0, 256, (R, 3), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 3, (R,), generator=g, dtype=torch.uint8)
    table = torch.randn(300, 512, generator=g)
    rows = torch.randint(0, 300, (R,), generator=g, dtype=torch.int32)
    from openobj_b200.background import BackgroundModel
    model = BackgroundModel(hidden=128, device=DEV)
    model.load([p[0] for p in fc] + [B[0]])
    gr, loss, terms = model.grads(pcs.to(DEV), z.to(DEV), gt_depth.to(DEV), rgb8.to(DEV), labels.to(DEV), rows.to(DEV),
                                  table.to(DEV))
    f64 = [p.double() for p in fc]
    t64, g64 = oc.train_step_grads(f64, B.double(), pcs[None].double(), z[None].double(), gt_depth[None].double(),
                                   (rgb8 / 255.)[None].double(), labels[None], table[rows.long()][None].double(), scale=5.0)
    ref = float(t64.total)
    assert abs(float(loss) - ref) <= 1e-4 * abs(ref) + 1e-6, (float(loss), ref)
    for i, v in enumerate(model.views(gr)):
        r = g64[i][0].float()
        err, sc = float((v.cpu() - r).abs().max()), float(r.abs().max()) + 1e-12
        assert err <= 1e-4 * sc + 1e-7, (i, err, sc)


def test_scene_with_background_model():
    """Scene with do_bg: the background object gets its own ring, is sampled with 5 + 9 bins over win_size_bg keyframes
    (train.py:300-315) and trained beside the ensemble; its loss goes down over two frames."""
    from openobj_b200 import cfg as C
    from openobj_b200.scene import Scene
    from openobj_b200.synthetic import SyntheticScene
    torch.manual_seed(0)                  # the reference-surface modules draw their initial weights from the global RNG
    cfg = C.room0_config()
    cfg.do_bg = True
    cfg.W, cfg.H = 200, 120
    cfg.fx = cfg.fy = 100.0
    cfg.cx, cfg.cy = 99.5, 59.5
    cfg.n_iter_per_frame = 20
    synth = SyntheticScene(4, W=cfg.W, H=cfg.H, part_mode=True, seed=2, n_distinct=1, with_bg=True)
    sc = Scene(cfg, seed=5, max_frames=6)
    first = last = None
    for f in range(3):
        sc.add_frame(synth.frame(f))
        sc.sample()
        bg_loss = torch.zeros(cfg.n_iter_per_frame, device=DEV)
        sc.train(bg_loss=bg_loss)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(bg_loss).all())
        first = float(bg_loss.mean()) if first is None else first       # mean over the frame's steps (each step sees other rays)
        last = float(bg_loss.mean())
    assert sc.bg is not None and sc.bg.hidden == 128 and sc.bg.adam_t.cpu().tolist() == [3 * cfg.n_iter_per_frame] * 3
    b = sc.bg_batch
    assert b.z.shape == (1, cfg.n_iter_per_frame * cfg.n_per_optim_bg, 14)
    assert 0 not in sc.obj_dict and len(sc.obj_dict) == 4
    # the module the reference's checkpoint code reads (vmap.py:556-576) aliases the trained block
    p0 = next(sc.scene_bg.trainer.fc_occ_map.parameters())
    assert p0.data_ptr() == sc.bg.views()[0].data_ptr()
    assert last < first, (first, last)


@pytest.mark.parametrize("hidden,R,S,part", [(64, 37, 14, True), (256, 53, 14, True), (32, 41, 10, False), (128, 129, 6, True)])
def test_bg_other_widths_and_ragged_sizes_against_oracle(hidden, R, S, part):
    """The layer-by-layer path is generic in the hidden width (model.OccupancyMap routes every width but 32 to it; the
    reference's default is 256) and in the point count: ragged sizes put partial 128-row tiles, column tails (87 / h + 87 / h +
    42 wide weights), split contractions that end inside a 16-wide chunk and both GEMM engines (tcgen05 where the operand
    layouts allow, mma.sync for the 1- and 3-wide ones) on the path.  Loss and all 19 gradients against the float64 oracle."""
    g = torch.Generator().manual_seed(100 + hidden + R)
    fc, B = oc.init_params(1, hidden=hidden, generator=g)
    fc[8] *= 0.3
    fc[9] *= 0.3
    z = torch.sort(0.5 + 5.0 * torch.rand(R, S, generator=g), dim=-1).values
    d = torch.nn.functional.normalize(torch.randn(R, 1, 3, generator=g), dim=-1)
    pcs = (torch.randn(R, 1, 3, generator=g) * 0.3 + d * z[..., None]).float()
    gt_depth = z[:, S // 2].clone()
    rgb8 = torch.randint(0, 256, (R, 3), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 3, (R,), generator=g, dtype=torch.uint8)
    labels[0], labels[1] = 1, 0
    table = torch.randn(64, 512, generator=g) if part else None
    rows = torch.randint(0, 64, (R,), generator=g, dtype=torch.int32) if part else None
    from openobj_b200.background import BackgroundModel
    model = BackgroundModel(hidden=hidden, device=DEV, rays_per_step=R, n_samp=S)
    model.load([p[0] for p in fc] + [B[0]])
    gr, loss, terms = model.grads(pcs.to(DEV), z.to(DEV), gt_depth.to(DEV), rgb8.to(DEV), labels.to(DEV),
                                  rows.to(DEV) if part else None, table.to(DEV) if part else None)
    f64 = [p.double() for p in fc]
    feat = table[rows.long()][None].double() if part else None
    t64, g64 = oc.train_step_grads(f64, B.double(), pcs[None].double(), z[None].double(), gt_depth[None].double(),
                                   (rgb8 / 255.)[None].double(), labels[None], feat, scale=5.0)
    ref = float(t64.total)
    assert abs(float(loss) - ref) <= 1e-4 * abs(ref) + 1e-6, (float(loss), ref)
    for i, v in enumerate(model.views(gr)):
        if g64[i] is None:
            assert float(v.abs().max()) == 0.0, i            # part features off: the clip head gets no gradient
            continue
        r = g64[i][0].float()
        err, sc = float((v.cpu() - r).abs().max()), float(r.abs().max()) + 1e-12
        assert err <= 1e-4 * sc + 1e-7, (hidden, i, err, sc)
