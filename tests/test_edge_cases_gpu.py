"""Edge cases and size-independent properties on the GPU: ragged ray counts, a single object, many
objects (ScanNet-shape config 4: part features off), both zero-mask flags, sampler invariants at full frame size."""
import numpy as np
import pytest
import torch

import openobj_oracle as oc
from openobj_b200 import layout

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batch(N, RAYS, seed, feat):
    from test_train_gpu import synth_batch
    return synth_batch(N, RAYS, seed=seed, feat=feat)


def _ensemble(N, R, I, seed):
    from openobj_b200.ensemble import Ensemble
    fc, B = oc.init_params(N, generator=torch.Generator().manual_seed(seed))
    fc[8] *= 0.3       # out of saturation: see the conditioning note in tests/test_train_gpu.py::test_room0_shape_steps_match_oracle
    fc[9] *= 0.3
    ens = Ensemble(N, rays_per_step=R, iters_per_frame=I)
    ens.load_stacked(fc + [B])
    return ens, fc, B


@pytest.mark.parametrize("N,R,feat", [(1, 120, True), (3, 7, True), (2, 33, False), (5, 101, True), (200, 120, False)])
def test_ragged_and_extreme_shapes(N, R, feat):
    """rays per step not a multiple of the 10-ray tile, one object, 200 objects (more tiles than CTAs): losses rel 1e-4
    against the oracle; gradients against its float64 / float32 evaluations."""
    from openobj_b200.ensemble import FrameBatch
    from test_train_gpu import check_grads, grads64, relu_flip_explainer
    pcs, z, gt_depth, rgb8, labels, gt_feat = _batch(N, R, seed=N * 1000 + R, feat=feat)
    labels[:, 0], labels[:, -1] = 1, 0
    ens, fc, B = _ensemble(N, R, 1, seed=N + R)
    batch = FrameBatch.from_dense(pcs.to(DEV), z.to(DEV), gt_depth.to(DEV), rgb8.to(DEV), labels.to(DEV),
                                  gt_feat.to(DEV) if feat else None)
    ens.prepare_frame(batch)
    g, terms = ens.grads(batch, 0)
    rt, rg = grads64(fc, B, pcs, z, gt_depth, rgb8 / 255., labels, gt_feat)
    ref_t = torch.stack([rt.depth, rt.color, rt.opacity, rt.feat], 1).float()
    torch.testing.assert_close(terms.cpu(), ref_t, rtol=1e-4, atol=1e-6)
    _, rg32 = oc.train_step_grads(fc, B, pcs, z, gt_depth, rgb8 / 255., labels, gt_feat)
    check_grads(g, rg, ref_grads_alt=rg32,
                explain=relu_flip_explainer(fc, B, pcs, z, gt_depth, rgb8 / 255., labels, gt_feat))
    if not feat:       # clip head: no gradient at all (quirk 8)
        for i in (14, 15, 16, 17):
            assert float(layout.views(g)[i].abs().max()) == 0.0


def test_both_zero_mask_flags_skip_every_group():
    """Some object without label==1 rays AND some object without label!=2 rays: every loss term is zero for all objects
    (render_rays.py:89-94) and no parameter moves (no gradient reaches any tensor)."""
    from openobj_b200.ensemble import FrameBatch
    N, R = 3, 20
    pcs, z, gt_depth, rgb8, labels, gt_feat = _batch(N, R, seed=5, feat=True)
    labels[0] = 2            # object 0: only unknown pixels
    labels[1] = 0
    ens, fc, B = _ensemble(N, R, 1, seed=9)
    batch = FrameBatch.from_dense(pcs.to(DEV), z.to(DEV), gt_depth.to(DEV), rgb8.to(DEV), labels.to(DEV), gt_feat.to(DEV))
    before = ens.theta.clone()
    lt = torch.ones(1, N, 4, device=DEV)
    ens.train_frame(batch, loss_terms=lt)
    torch.cuda.synchronize()
    assert int(ens.flags[0]) == 6 and float(lt.abs().max()) == 0.0
    assert torch.equal(before, ens.theta) and ens.adam_t.cpu().tolist() == [0, 0, 0]
    terms = oc.step_loss(*oc.ensemble_forward(fc, B, pcs)[:2], gt_depth, rgb8 / 255., labels, z, gt_feat,
                         oc.ensemble_forward(fc, B, pcs)[2])
    assert terms.flags & 6 == 6 and float(terms.total) == 0.0


def test_partial_frame_and_step_offsets():
    """train_frame(iters < iters_per_frame) consumes exactly the first slices; Adam step counters advance by iters."""
    from openobj_b200.ensemble import FrameBatch
    N, R, I = 4, 30, 5
    pcs, z, gt_depth, rgb8, labels, gt_feat = _batch(N, R * I, seed=11, feat=True)
    ens, fc, B = _ensemble(N, R, I, seed=12)
    batch = FrameBatch.from_dense(pcs.to(DEV), z.to(DEV), gt_depth.to(DEV), rgb8.to(DEV), labels.to(DEV), gt_feat.to(DEV))
    lt = torch.zeros(I, N, 4, device=DEV)
    ens.train_frame(batch, iters=2, loss_terms=lt)
    assert ens.adam_t.cpu().tolist() == [2, 2, 2]
    assert float(lt[2:].abs().max()) == 0.0 and float(lt[:2].abs().min()) > 0.0
    sl = slice(0, R)
    rt, _ = oc.train_step_grads(fc, B, pcs[:, sl], z[:, sl], gt_depth[:, sl], rgb8[:, sl] / 255., labels[:, sl], gt_feat[:, sl])
    got = float(ens.total_loss(lt[0].cpu()))
    assert abs(got - float(rt.total.detach())) <= 1e-4 * abs(float(rt.total.detach()))
    with pytest.raises(Exception):
        ens.train_frame(batch, iters=I + 1)        # more steps than the pre-sampled batch holds


def test_sampler_invariants_full_frame_size():
    """Replica frame size, 60 objects, counter RNG: pixel indices inside each object's bbox, labels match the ring's state
    plane, z inside its class interval, normal bins sorted and within +-eps of the depth, pcs = o + d z."""
    from openobj_b200 import cfg as C
    from openobj_b200.scene import Scene
    from openobj_b200.synthetic import SyntheticScene
    cfg = C.room0_config()
    cfg.do_bg = False
    n_obj = 60
    synth = SyntheticScene(n_obj, W=cfg.W, H=cfg.H, part_mode=True, seed=2, n_distinct=1)
    sc = Scene(cfg, seed=5, max_frames=4)
    for f in range(3):
        sc.add_frame(synth.frame(f))
    b = sc.sample()
    out = sc.sample_out
    torch.cuda.synchronize()
    assert int(out.oob.item()) == 0
    objs = list(sc.obj_dict.values())
    n_rays = out.labels.shape[1]
    assert n_rays == cfg.n_iter_per_frame * cfg.n_per_optim == 12000
    eps, oeps = cfg.surface_eps, cfg.stop_eps
    z, d, lab = out.z, out.gt_depth, out.labels
    valid = out.valid.bool()
    assert torch.equal(valid, d > 0)
    zmax = d.max(dim=1, keepdim=True).values
    inv = ~valid
    assert bool((z[inv] >= 0).all()) and bool((z[inv] <= zmax.expand_as(d)[inv][:, None] + 1e-6).all())
    assert bool((z[valid][:, 0] >= 0).all()) and bool((z[valid][:, 0] <= (d[valid] - eps) + 1e-6).all())
    is_obj = valid & (lab == 1)
    zo = z[is_obj][:, 1:]
    assert bool((zo[:, 1:] >= zo[:, :-1]).all())                      # sorted normal bins (utils.py:391)
    assert bool(((zo - d[is_obj][:, None]).abs() <= eps + 1e-6).all())
    oth = valid & (lab != 1)
    zt = z[oth][:, 1:]
    assert bool((zt >= (d[oth] - eps)[:, None] - 1e-6).all()) and bool((zt <= (d[oth] + oeps)[:, None] + 1e-6).all())
    assert int((lab == 1).sum()) > 0 and int((lab == 2).sum()) > 0 and int(lab.max()) <= 2
    # part-feature rows inside the table; every object samples its two latest keyframes in the last two frame draws
    assert int(out.feat_row.min()) >= 0 and int(out.feat_row.max()) < sc.part_table.shape[0] * sc.pw * sc.ph
    # linearity of the points in z: pcs = origin + dir * z  =>  second differences along the ray vanish relative to |dz|
    p = out.pcs
    dz = (z[..., 1:] - z[..., :-1])[..., None]
    dirs = (p[..., 1:, :] - p[..., :-1, :]) / dz.clamp_min(1e-6)
    ok = dz[..., 0] > 1e-3
    spread = (dirs - dirs[..., :1, :]).abs().amax(-1)
    assert float(spread[ok].max()) < 2e-3


def test_scene_scannet_shape_part_features_off():
    """BASELINE config 4 shape: 640x480 frames, part_mode off (the shipped ScanNet JSONs): the clip head gets no gradient,
    so its parameters and moments stay untouched while everything else trains; staged (copy-stream) frames give the
    same result as frames handed over directly."""
    from openobj_b200 import cfg as C
    from openobj_b200.scene import Scene
    from openobj_b200.synthetic import SyntheticScene
    from openobj_b200 import layout as Lo
    res = []
    for staged in (False, True):
        torch.manual_seed(11)
        cfg = C.room0_config()
        cfg.do_bg = False
        cfg.part_mode = False
        cfg.W, cfg.H = 640, 480
        cfg.fx = cfg.fy = 320.0
        cfg.cx, cfg.cy = 319.5, 239.5
        cfg.n_iter_per_frame = 10
        synth = SyntheticScene(8, W=cfg.W, H=cfg.H, part_mode=False, seed=4, n_distinct=1, pin=True)
        sc = Scene(cfg, seed=21, max_frames=4)
        lt = torch.zeros(cfg.n_iter_per_frame, 8, 4, device=DEV)
        for f in range(2):
            fr = synth.frame(f)
            sc.add_frame(sc.stage_frame(fr) if staged else fr)
            sc.sample()
            if f == 0:
                clip0 = [Lo.views(sc.ens.theta)[i].clone() for i in (14, 15, 16, 17)]
            sc.train(loss_terms=lt)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(lt).all()) and float(lt[..., 3].abs().max()) == 0.0      # no feature term
        for i, c0 in zip((14, 15, 16, 17), clip0):
            assert torch.equal(Lo.views(sc.ens.theta)[i], c0)                                # clip head untouched (quirk 8)
            assert float(Lo.views(sc.ens.m)[i].abs().max()) == 0.0
        assert sc.ens.adam_t.cpu().tolist()[2] == 0 and sc.ens.adam_t.cpu().tolist()[0] == 2 * cfg.n_iter_per_frame
        res.append((sc.ens.theta.clone(), lt.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


def test_staged_frames_with_part_features_equal_direct_frames():
    """Scene.stage_frame sends the part features of the NEXT frame straight into their slot of the resident table on the copy
    stream (their own event, awaited by the training kernels only): pipelined staging must give bit-identical parameters and
    loss terms to frames handed over directly, with part features on."""
    from openobj_b200 import cfg as C
    from openobj_b200.scene import Scene
    from openobj_b200.synthetic import SyntheticScene
    res = []
    for staged in (False, True):
        torch.manual_seed(5)
        cfg = C.room0_config()
        cfg.do_bg = False
        cfg.part_mode = True
        cfg.W, cfg.H = 320, 240
        cfg.fx = cfg.fy = 160.0
        cfg.cx, cfg.cy = 159.5, 119.5
        cfg.n_iter_per_frame = 6
        synth = SyntheticScene(5, W=cfg.W, H=cfg.H, part_mode=True, seed=9, n_distinct=2, pin=True)
        sc = Scene(cfg, seed=3, max_frames=6)
        lt = torch.zeros(cfg.n_iter_per_frame, 5, 4, device=DEV)
        terms = []
        nxt = sc.stage_frame(synth.frame(0)) if staged else synth.frame(0)
        for f in range(4):
            sc.add_frame(nxt)
            if f + 1 < 4:                      # the next frame's copies are in flight while this one samples and trains
                nxt = sc.stage_frame(synth.frame(f + 1)) if staged else synth.frame(f + 1)
            sc.sample()
            sc.train(loss_terms=lt)
            terms.append(lt.clone())
        torch.cuda.synchronize()
        assert bool(torch.isfinite(torch.stack(terms)).all()) and float(torch.stack(terms)[..., 3].abs().max()) > 0.0
        res.append((sc.ens.theta.clone(), torch.stack(terms), sc.part_table[:4].clone()))
    assert torch.equal(res[0][2], res[1][2])
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


@pytest.mark.parametrize("rank", [0, 1])
def test_sharded_staging_gathers_only_local_part_rows(rank):
    """A rank of a sharded run stages the part features of a frame by reading ONLY the cells inside its own objects' boxes from
    the pinned host tensor (oo_gather_part_rows), objects first seen in that frame included: parameters and loss terms must be
    bit-identical to the same rank fed whole frames directly."""
    from openobj_b200 import cfg as C
    from openobj_b200.scene import Scene
    from openobj_b200.synthetic import SyntheticScene
    res = []
    for staged in (False, True):
        torch.manual_seed(5)
        cfg = C.room0_config()
        cfg.do_bg = False
        cfg.part_mode = True
        cfg.W, cfg.H = 320, 240
        cfg.fx = cfg.fy = 160.0
        cfg.cx, cfg.cy = 159.5, 119.5
        cfg.n_iter_per_frame = 5
        cfg.max_n_models = 6
        synth = SyntheticScene(6, W=cfg.W, H=cfg.H, part_mode=True, seed=9, n_distinct=2, pin=True)
        sc = Scene(cfg, rank=rank, world=2, seed=3, max_frames=6)
        sc.gather_fraction = 2.0                    # always gather (the default only does when the boxes cover < 60 % of the frame)
        n_loc = 3
        lt = torch.zeros(cfg.n_iter_per_frame, n_loc, 4, device=DEV)
        terms = []
        nxt = sc.stage_frame(synth.frame(0)) if staged else synth.frame(0)
        for f in range(3):
            sc.add_frame(nxt)
            if f + 1 < 3:
                nxt = sc.stage_frame(synth.frame(f + 1)) if staged else synth.frame(f + 1)
            sc.sample()
            sc.train(loss_terms=lt)
            terms.append(lt.clone())
        torch.cuda.synchronize()
        assert len(sc.obj_dict) == n_loc and bool(torch.isfinite(torch.stack(terms)).all())
        assert float(torch.stack(terms)[..., 3].abs().max()) > 0.0
        if staged:
            full = 3 * (synth.frame_bytes() - 128)
            assert 0 < sc.h2d_bytes < full            # fewer bytes than three whole frames crossed the host link
        res.append((sc.ens.theta.clone(), torch.stack(terms)))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
