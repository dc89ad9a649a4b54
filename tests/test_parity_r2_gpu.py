"""Round-2 parity tests (VERDICT r1 "next round" item 1): the fused path at BASELINE config-2 size against the oracle over a
whole frame, un-softened and saturated weights with a measured tolerance, the explode flag of the fused path, the background
model's zero-mask AdamW skip, sharded Scene == single-rank Scene (N = 100 part features on, N = 200 at 640x480 part features
off, more ranks than objects), checkpoints written through Scene."""
import os

import numpy as np
import pytest
import torch

import openobj_oracle as oc
from openobj_b200 import layout

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PTOL = dict(rtol=1e-3, atol=2e-4)


def params_close(v, ref, steps, lr=1e-3, frac=1e-3):
    """Parameters after k AdamW steps: rel 1e-3 / abs 2e-4 on all but `frac` of a tensor's elements, every element within
    2 lr per step.  Adam normalises each element's step to ~lr whatever the size of its gradient, so an element whose gradient
    is a near-complete cancellation (|g| at rounding level) steps in a direction that depends on the summation order -- fp32
    FMA chains, 3xTF32 fragments and the reference's own CPU / CUDA kernels all differ there (tests/test_train_gpu.py)."""
    d = (v - ref).abs()
    bad = d > (PTOL["atol"] + PTOL["rtol"] * ref.abs())
    assert int(bad.sum()) <= max(1, int(frac * ref.numel())), (int(bad.sum()), ref.numel(), float(d.max()))
    assert float(d.max()) <= 2 * lr * steps + 1e-6, float(d.max())


def _oracle_frame(fc, B, batch, iters, R, part, dtype=torch.float64):
    """The oracle's trajectory over one pre-sampled frame (train.py:394-474): per-step LossTerms and final parameters."""
    c = lambda t: t.detach().cpu()
    cast = lambda t: t.to(dtype)
    pcs, z, gd = cast(c(batch.pcs)), cast(c(batch.z)), cast(c(batch.gt_depth))
    rgb, lab = cast(c(batch.gt_rgb)) / 255., c(batch.labels)
    rows = c(batch.feat_row).long() if part else None
    table = c(batch.feat_table) if part else None
    P = [cast(p.clone()) for p in fc] + [cast(B.clone())]
    M = [torch.zeros_like(p) for p in P]
    V = [torch.zeros_like(p) for p in P]
    steps = [0] * 19
    terms = []
    for it in range(iters):
        sl = slice(it * R, (it + 1) * R)
        gf = cast(table[rows[:, sl]]) if part else None
        t, g = oc.train_step_grads(P[:18], P[18], pcs[:, sl], z[:, sl], gd[:, sl], rgb[:, sl], lab[:, sl], gf)
        terms.append(t)
        for i, gr in enumerate(g):
            if gr is not None:
                steps[i] += 1
                oc.adamw_step(P[i], gr, M[i], V[i], steps[i])
    return terms, P


def test_full_frame_trajectory_config2_size():
    """BASELINE config 2: 60 objects, 1200 x 680 frames, part features on, ONE whole frame = 100 optimisation steps through
    Scene (shared store -> K2 counter RNG -> 100 x (K1 + K4)), the reference's weight init.  The oracle (float64)
    trains on the same sampled rays; for six objects the samples themselves are re-derived by the oracle's sampler from the
    same counter stream (bit-exact).  Per-step loss rel 1e-4 over the first 10 steps; over all 100 steps and for
    the parameters after 100 steps the bound is the divergence measured in the same run between the oracle in float32 (the
    reference's arithmetic) and the oracle in float64, factor 2."""
    from openobj_b200 import cfg as C, sampler
    from openobj_b200.scene import Scene
    from openobj_b200.synthetic import SyntheticScene
    torch.manual_seed(7)
    cfg = C.room0_config()
    cfg.do_bg = False
    n_obj, n_fill = 60, 3
    synth = SyntheticScene(n_obj, W=cfg.W, H=cfg.H, part_mode=True, seed=0, n_distinct=2)
    sc = Scene(cfg, seed=1234, max_frames=n_fill + 1)
    for f in range(n_fill):
        sc.add_frame(synth.frame(f))
    with torch.no_grad():           # out of saturation (conditioning note in tests/test_train_gpu.py): the objects keep the
        for i in (8, 9):            # reference's init (model.init_weights) except for a 0.3 x gain on the out_alpha layer
            sc.ens.stacked()[i].mul_(0.3)
    sc.sample()
    batch = sc.batch
    R, I = cfg.n_per_optim, cfg.n_iter_per_frame
    assert tuple(batch.labels.shape) == (n_obj, R * I) and sc.store.frames_alive() <= n_fill
    fc = [v.detach().cpu().clone() for v in sc.ens.stacked()[:18]]
    B = sc.ens.stacked()[18].detach().cpu().clone()
    # ---- the samples: oracle sampler on rings rebuilt from the frames, for six objects
    n_frames, n_samples = I * cfg.win_size, cfg.n_samples_per_frame
    frames = [synth.frame(f) for f in range(n_fill)]
    rays_dir = sc.cam.rays_dir_cache.cpu()
    for i in (0, 7, 19, 33, 48, 59):
        o = sc._objs[i]
        tapes = sampler.device_tapes([o], n_frames, n_samples, 1, 9, o.surface_eps, sc.seed, sc.frames_seen, DEV)
        nk = o.n_keyframes
        rgbs = torch.zeros(nk, cfg.W, cfg.H, 4, dtype=torch.uint8)
        dep = torch.zeros(nk, cfg.W, cfg.H)
        twc = torch.zeros(nk, 4, 4)
        bbox = torch.zeros(nk, 4)
        for fid, slot in o.ring.slot_of.items():
            fr = frames[fid // 10]
            rgbs[slot, ..., :3] = fr["image"]
            rgbs[slot, ..., 3] = (fr["obj"] == o.obj_id).to(torch.uint8) + 2 * (fr["obj"] == -1).to(torch.uint8)
            dep[slot], twc[slot], bbox[slot] = fr["depth"], fr["T"].float(), fr["bbox_dict"][o.obj_id].float()
        val, lab = sc.sample_out.valid[i].cpu().bool(), batch.labels[i].cpu()
        tp = oc.SampleTape(kf_ids=tapes.kf_ids[0].cpu(), u_w=tapes.u_w[0].cpu().view(n_frames, n_samples),
                           u_h=tapes.u_h[0].cpu().view(n_frames, n_samples), r_invalid=tapes.r_invalid[0].cpu()[~val],
                           r_valid=tapes.r_valid[0].cpu()[val], r_normal=tapes.r_normal[0].cpu()[val & (lab == 1)],
                           r_other=tapes.r_other[0].cpu()[val & (lab != 1)])
        ref = oc.sample_object(rgbs, dep, twc, bbox, rays_dir, tp)
        assert torch.equal(lab, ref["labels"]) and torch.equal(val, ref["valid"])
        assert torch.equal(batch.gt_rgb[i].cpu(), ref["rgb"].reshape(-1, 3)) and torch.equal(batch.gt_depth[i].cpu(), ref["depth"].reshape(-1))
        assert torch.equal(batch.z[i].cpu(), ref["z"].reshape(-1, 10))
        torch.testing.assert_close(batch.pcs[i].cpu(), ref["pcs"].reshape(-1, 10, 3), rtol=1e-6, atol=1e-6)
    # ---- the frame's 100 steps
    lt = torch.zeros(I, n_obj, 4, device=DEV)
    sc.train(loss_terms=lt)
    sc.finish()
    torch.cuda.synchronize()
    terms, P = _oracle_frame(fc, B, batch, I, R, part=True)
    terms32, P32 = _oracle_frame(fc, B, batch, I, R, part=True, dtype=torch.float32)      # the reference's own arithmetic
    got = sc.ens.total_loss(lt.cpu().double())
    ref = torch.stack([t.total.detach() for t in terms])
    own = torch.stack([t.total.detach().double() for t in terms32])
    rel = ((got - ref).abs() / ref.abs())
    rel32 = ((own - ref).abs() / ref.abs())
    print("loss rel error vs float64, kernel / oracle-fp32: steps 0-9 %.1e / %.1e, 10-49 %.1e / %.1e, 50-99 %.1e / %.1e"
          % (rel[:10].max(), rel32[:10].max(), rel[10:50].max(), rel32[10:50].max(), rel[50:].max(), rel32[50:].max()))
    assert float(rel[:10].max()) <= 1e-4, rel[:10].tolist()
    # Later steps: two correct fp32 evaluations of this recurrence drift apart (Adam turns a rounding-level difference of a
    # gradient element into a +-lr step), so the bound is the drift MEASURED on the reference's own arithmetic: the oracle in
    # float32 against the oracle in float64 over the same window of steps, factor 3.  Measured on the B200
    # (gpurun_out/pytest_r2e.log): kernel / oracle-fp32 = 1.3e-5 / 8.2e-6 (steps 0-9), 9.9e-4 / 5.9e-4 (10-49),
    # 4.1e-3 / 4.2e-3 (50-99).
    for lo, hi in ((0, 10), (10, 50), (50, I)):
        assert float(rel[lo:hi].max()) <= max(1e-4, 3 * float(rel32[lo:hi].max())), (lo, hi, float(rel[lo:hi].max()), float(rel32[lo:hi].max()))
    lr = cfg.learning_rate
    for name, v, p, p32 in zip(layout.NAMES, sc.ens.stacked(), P, P32):
        tolv = PTOL["atol"] + PTOL["rtol"] * p.abs()
        dk, do = (v.cpu().double() - p).abs(), (p32.double() - p).abs()
        fk, fo = float((dk > tolv).double().mean()), float((do > tolv).double().mean())
        print("%-24s outside rel 1e-3 / abs 2e-4 after 100 steps: kernel %.4f (max %.4f), oracle-fp32 %.4f (max %.4f)"
              % (name, fk, float(dk.max()), fo, float(do.max())))
        assert fk <= max(1e-3, 3 * fo) and float(dk.max()) <= max(20 * lr, 3 * float(do.max())), (name, fk, fo, float(dk.max()), float(do.max()))


def _synth(N, RAYS, seed, feat=True, S=10):
    g = torch.Generator().manual_seed(seed)
    z = torch.sort(0.5 + 3.0 * torch.rand(N, RAYS, S, generator=g), dim=-1).values
    o = torch.randn(N, RAYS, 1, 3, generator=g) * 0.2
    d = torch.nn.functional.normalize(torch.randn(N, RAYS, 1, 3, generator=g), dim=-1)
    pcs = (o + d * z[..., None]).float()
    gt_depth = (z[..., 6] + 0.05 * torch.randn(N, RAYS, generator=g)).float()
    rgb8 = torch.randint(0, 256, (N, RAYS, 3), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 3, (N, RAYS), generator=g, dtype=torch.uint8)
    labels[:, 0], labels[:, 1] = 1, 0
    gt_feat = torch.randn(N, RAYS, 512, generator=g) if feat else None
    return pcs, z, gt_depth, rgb8, labels, gt_feat


@pytest.mark.parametrize("alpha_gain", [1.0, 3.0])
def test_reference_init_and_saturated_alpha_measured_tolerance(alpha_gain):
    """Reference-init weights as they are (gain 1: mean |alpha| 4, max 36 on this batch) and a trained-looking, saturated
    out_alpha layer (gain 3: mean |alpha| 12, max 107: sigmoid(alpha) within an ulp of 0 or 1 on most samples).  The reference
    forms 1 - occ in fp32 (render_rays.py:38), so on saturated rays its OWN fp32 evaluation moves against exact arithmetic --
    measured on this batch with the oracle: depth term 4e-6 (gain 1) and 0.28 (gain 3, through the depth weight
    1 / (sqrt(var) + 1e-4)), gradients 6e-6 and 1.5e-4.  The stated tolerance is therefore measured in the same run: per loss
    term (and per gradient tensor) the kernel must be within max(1e-4, 2 x the worst deviation over the objects of the oracle's
    fp32 evaluation from its float64 evaluation), relative to the float64 value."""
    from openobj_b200.ensemble import Ensemble, FrameBatch
    N, R = 8, 120
    pcs, z, gt_depth, rgb8, labels, gt_feat = _synth(N, R, seed=31)
    fc, B = oc.init_params(N, generator=torch.Generator().manual_seed(77))
    fc[8] = fc[8] * alpha_gain
    fc[9] = fc[9] * alpha_gain
    d = lambda t: t.double()
    t64, g64 = oc.train_step_grads([d(p) for p in fc], d(B), d(pcs), d(z), d(gt_depth), d(rgb8) / 255., labels, d(gt_feat))
    t32, g32 = oc.train_step_grads(fc, B, pcs, z, gt_depth, rgb8 / 255., labels, gt_feat)
    alpha64 = oc.ensemble_forward([d(p) for p in fc], d(B), d(pcs))[0]
    amax = float(alpha64.abs().max())
    assert amax > (10.0 if alpha_gain > 1 else 2.0), amax
    ens = Ensemble(N, rays_per_step=R, iters_per_frame=1)
    ens.load_stacked(fc + [B])
    batch = FrameBatch.from_dense(pcs.to(DEV), z.to(DEV), gt_depth.to(DEV), rgb8.to(DEV), labels.to(DEV), gt_feat.to(DEV))
    ens.prepare_frame(batch)
    g, terms = ens.grads(batch, 0)
    torch.cuda.synchronize()
    ref_t = torch.stack([t64.depth, t64.color, t64.opacity, t64.feat], 1)
    own_t = torch.stack([t32.depth, t32.color, t32.opacity, t32.feat], 1).double()
    err_k = (terms.cpu().double() - ref_t).abs() / (ref_t.abs() + 1e-12)
    err_o = (own_t - ref_t).abs() / (ref_t.abs() + 1e-12)
    bound = torch.clamp(2 * err_o.max(0).values, min=1e-4)
    assert bool((err_k.max(0).values <= bound).all()), (err_k.max(0).values.tolist(), err_o.max(0).values.tolist())
    worst = 0.0
    for name, gk, r64, r32 in zip(layout.NAMES, layout.views(g.cpu()), g64, g32):
        sc = r64.reshape(N, -1).abs().max(1).values
        ek = (gk.double() - r64).reshape(N, -1).abs().max(1).values
        eo = (r32.double() - r64).reshape(N, -1).abs().max(1).values
        ok = (ek / sc) <= max(1e-4, 2 * float((eo / sc).max()))
        assert bool(ok.all()), (name, (ek / sc).tolist(), (eo / sc).tolist())
        worst = max(worst, float((ek / sc).max()))
    print("alpha gain %.0f: max|alpha| %.1f, kernel-vs-f64 worst gradient error %.2e" % (alpha_gain, amax, worst))


def test_explode_flag_from_the_fused_path():
    """render_rays.py:109-111: a per-object loss term above 1e5 makes the reference print 'loss explode' and exit(-1).  The
    fused path raises OO_FLAG_EXPLODE in that step's flags (k_update) and the host looks once per frame."""
    from openobj_b200.ensemble import Ensemble, FrameBatch
    N, R, I = 4, 30, 3
    pcs, z, gt_depth, rgb8, labels, gt_feat = _synth(N, R * I, seed=3)
    fc, B = oc.init_params(N, generator=torch.Generator().manual_seed(1))
    gd = gt_depth.clone()
    gd[2, R:2 * R] = 3.0e6                     # object 2, step 1: |depth - gt| / (sqrt(var) + 1e-4) averages far above 1e5
    ens = Ensemble(N, rays_per_step=R, iters_per_frame=I)
    ens.load_stacked(fc + [B])
    b = FrameBatch.from_dense(pcs.to(DEV), z.to(DEV), gd.to(DEV), rgb8.to(DEV), labels.to(DEV), gt_feat.to(DEV))
    lt = torch.zeros(I, N, 4, device=DEV)
    ens.train_frame(b, loss_terms=lt)
    torch.cuda.synchronize()
    assert ens.flags.cpu().tolist() == [0, 1, 0] and float(lt[1, 2, 0]) > 1e5
    ref = oc.step_loss(*oc.ensemble_forward(fc, B, pcs[:, R:2 * R])[:2], gd[:, R:2 * R], rgb8[:, R:2 * R] / 255., labels[:, R:2 * R], z[:, R:2 * R])
    assert ref.flags & 1                       # the oracle's restatement of the same guard fires on the same step
    with pytest.raises(FloatingPointError):
        ens.check_explode(wait=True)
    # a clean frame leaves the flag down
    b2 = FrameBatch.from_dense(pcs.to(DEV), z.to(DEV), gt_depth.to(DEV), rgb8.to(DEV), labels.to(DEV), gt_feat.to(DEV))
    ens.train_frame(b2, loss_terms=lt)
    ens.check_explode(wait=True)
    assert ens.flags.cpu().tolist() == [0, 0, 0]


def test_background_zero_mask_skips_parameter_groups():
    """train.py:455-463 with the background's step_batch_loss on [1, R, S]: no label-1 ray => depth / colour / feature terms
    are constant zeros (render_rays.py:89-94), the colour and clip heads have grad None and torch.optim.AdamW leaves them
    alone (no decay, no step); the trunk still trains through the opacity term.  With neither mask nothing moves."""
    from openobj_b200.background import BackgroundModel
    R, S, h = 48, 14, 128
    g = torch.Generator().manual_seed(5)
    bg = BackgroundModel(hidden=h, device=DEV, rays_per_step=R, n_samp=S)
    fc, B = oc.init_params(1, hidden=h, generator=g)
    bg.load([p[0] for p in fc] + [B[0]])
    z = torch.sort(0.5 + 5.0 * torch.rand(R, S, generator=g), dim=-1).values
    dirs = torch.nn.functional.normalize(torch.randn(R, 1, 3, generator=g), dim=-1)
    pcs = (dirs * z[..., None]).contiguous()
    gd = z[:, 8].contiguous()
    rgb8 = torch.randint(0, 256, (R, 3), generator=g, dtype=torch.uint8)
    feat = torch.randn(R, 512, generator=g)
    rows = torch.arange(R, dtype=torch.int32)
    lab_no1 = (torch.randint(0, 2, (R,), generator=g) * 2).to(torch.uint8)        # labels 0 / 2 only
    before = [v.clone() for v in bg.views()]
    bg.train_step(pcs.to(DEV), z.to(DEV), gd.to(DEV), rgb8.to(DEV), lab_no1.to(DEV), rows.to(DEV), feat.to(DEV))
    torch.cuda.synchronize()
    assert int(bg.flags) == 2 and bg.adam_t.cpu().tolist() == [1, 0, 0]
    t, gr = oc.train_step_grads([p.clone() for p in fc], B.clone(), pcs[None], z[None], gd[None], rgb8[None] / 255.,
                                lab_no1[None], feat[None], scale=5.0)
    assert [i for i, x in enumerate(gr) if x is None] == list(range(10, 18))       # autograd reaches neither head
    for i, (v, v0) in enumerate(zip(bg.views(), before)):
        if 10 <= i < 18:
            assert torch.equal(v, v0), layout.NAMES[i]
        else:
            p = (fc + [B])[i][0].clone()
            oc.adamw_step(p, gr[i][0], torch.zeros_like(p), torch.zeros_like(p), 1)
            params_close(v.cpu(), p, 1)
    lab_none = torch.full((R,), 2, dtype=torch.uint8)
    before = [v.clone() for v in bg.views()]
    bg.train_step(pcs.to(DEV), z.to(DEV), gd.to(DEV), rgb8.to(DEV), lab_none.to(DEV), rows.to(DEV), feat.to(DEV))
    torch.cuda.synchronize()
    assert int(bg.flags) == 6 and bg.adam_t.cpu().tolist() == [1, 0, 0]
    assert all(torch.equal(v, v0) for v, v0 in zip(bg.views(), before))


def _run_scene(cfg, synth, frames, iters, rank=0, world=1, bits_other=None, seed=11):
    """Frames through a Scene; with world > 1 the flag all-reduce is emulated: bits_other[f] = the other ranks' zero-mask
    bits of frame f.  Returns (scene, per-frame bits of this rank, per-frame loss terms)."""
    from openobj_b200._lib import check, ptr, stream
    from openobj_b200.scene import Scene
    torch.manual_seed(seed)               # the objects' initial weights come from torch's generator (Trainer.load_network)
    fr = {"f": 0}
    calls = []

    def allreduce(bits):
        calls.append(bits.clone())
        if bits_other is not None:
            bits.copy_(torch.maximum(bits, bits_other[fr["f"]][:bits.shape[0]].to(bits.device)))

    sc = Scene(cfg, rank=rank, world=world, seed=99, init_seed=99, max_frames=frames + 1,
               flag_allreduce=allreduce if world > 1 else None)
    losses = []
    for f in range(frames):
        fr["f"] = f
        sc.add_frame(synth.frame(f))
        sc.sample()
        n = len(sc.obj_dict)
        lt = torch.zeros(iters, max(n, 1), 4, device=DEV)
        sc.train(iters=iters, loss_terms=lt if n else None)
        losses.append(lt)
    sc.finish()
    torch.cuda.synchronize()
    return sc, calls, losses


@pytest.mark.parametrize("n_obj,shape,part", [(100, (1200, 680), True), (200, (640, 480), False)])
def test_sharded_scene_equals_single_rank(n_obj, shape, part):
    """BASELINE configs 3 and 4 through Scene: N = 100 (Replica shape, CLIP + part heads) and N = 200 (ScanNet shape 640 x 480,
    part features off, n_models raised to 200 as SURVEY 8d notes).  Two ranks (emulated one after the other on this GPU, the
    flag all-reduce emulated with the other rank's bits) must give every object the parameters the single-rank run gives it:
    object -> rank assignment k mod 2 bit-exact, samples bit-exact (counter RNG keyed by object id), parameters equal up to
    the summation order of the gradient slots (the tile schedule depends on how many objects a rank holds)."""
    from openobj_b200 import cfg as C
    from openobj_b200.synthetic import SyntheticScene
    W, H = shape
    cfg = C.room0_config()
    cfg.do_bg = False
    cfg.part_mode = part
    cfg.W, cfg.H = W, H
    cfg.fx = cfg.fy = 0.5 * W
    cfg.cx, cfg.cy = 0.5 * W - 0.5, 0.5 * H - 0.5
    cfg.max_n_models = n_obj
    cfg.n_iter_per_frame = 4
    frames, iters = 2, 4
    synth = SyntheticScene(n_obj, W=W, H=H, part_mode=part, seed=6, n_distinct=2)
    full, _, loss_full = _run_scene(cfg, synth, frames, iters)
    assert len(full.obj_dict) == n_obj
    # pass 1: every rank alone, to learn its own bits; pass 2: with the other rank's bits merged in
    parts = []
    own = [[c for c in _run_scene(cfg, synth, frames, iters, rank=r, world=2)[1]] for r in range(2)]
    for r in range(2):
        sc, calls, _ = _run_scene(cfg, synth, frames, iters, rank=r, world=2, bits_other=own[1 - r])
        assert len(calls) == frames
        parts.append(sc)
    ids = list(full.obj_dict.keys())
    for r in range(2):
        assert list(parts[r].obj_dict.keys()) == ids[r::2]                         # k mod G, bit-exact
        assert parts[r].global_index == full.global_index
        idx = torch.arange(r, n_obj, 2, device=DEV)
        for name in ("labels", "z", "gt_depth", "gt_rgb", "pcs"):
            assert torch.equal(getattr(parts[r].batch, name), getattr(full.batch, name)[idx]), name
        assert parts[r].ens.adam_t.cpu().tolist() == full.ens.adam_t.cpu().tolist()
        a, b = parts[r].ens.theta, full.ens.theta[idx]
        diff = (a - b).abs()
        # 8 AdamW steps: an element whose gradient is a near-complete cancellation may step the other way (+- lr per step)
        assert float(diff.max()) <= 8e-3 and float((diff > 1e-5 + 1e-4 * b.abs()).float().mean()) <= 2e-3, float(diff.max())
    assert full.store.frames_alive() <= frames


def test_more_ranks_than_objects():
    """a rank that owns no object must neither crash nor leave the others waiting in the collective: it
    skips K2 / K1 / K4 and still joins the per-frame all-reduce; when its first object arrives it starts training; and a new
    object on ANY rank restarts Adam on every rank (train.py:272-276)."""
    from openobj_b200 import cfg as C
    from openobj_b200.synthetic import SyntheticScene
    cfg = C.room0_config(w=100, h=60)
    cfg.do_bg = False
    cfg.n_iter_per_frame = 3
    synth2 = SyntheticScene(2, W=100, H=60, part_mode=True, seed=1, n_distinct=1)
    for rank in range(4):
        sc, calls, _ = _run_scene(cfg, synth2, frames=3, iters=3, rank=rank, world=4)
        assert len(calls) == 3 and all(tuple(c.shape) == (3, 2) for c in calls)
        assert len(sc.obj_dict) == (1 if rank < 2 else 0)
        assert (sc.ens is None) == (rank >= 2) and sc.global_index == {1: 0, 2: 1}
    # objects 1, 2 from frame 0, objects 3..6 from frame 2: rank 0 gets object 5 late, and the arrival of 3, 4, 6 (other
    # ranks' objects) restarts its optimiser too
    from openobj_b200.scene import Scene
    synth6 = SyntheticScene(6, W=100, H=60, part_mode=True, seed=1, n_distinct=1)
    torch.manual_seed(0)
    sc = Scene(cfg, rank=0, world=4, seed=3, max_frames=8, flag_allreduce=lambda b: None)
    def frame(f, keep):
        s = dict(synth6.frame(f))
        s["bbox_dict"] = {k: v for k, v in s["bbox_dict"].items() if k in keep}
        inst = s["obj"].clone()
        inst[(inst > 0) & ~torch.isin(inst, torch.tensor(sorted(keep), dtype=inst.dtype))] = 0
        s["obj"] = inst
        return s
    for f in range(2):
        sc.step_frame(frame(f, {1, 2}), iters=3)
    assert sc.ens.adam_t.cpu().tolist() == [6, 6, 6] and list(sc.obj_dict) == [1]
    sc.step_frame(frame(2, {1, 2, 3, 4}), iters=3)                  # objects 3, 4 belong to ranks 2, 3: reset, then 3 steps
    assert sc.ens.adam_t.cpu().tolist() == [3, 3, 3] and list(sc.obj_dict) == [1]
    sc.step_frame(frame(3, {1, 2, 3, 4, 5, 6}), iters=3)            # object 5 -> k = 4 -> rank 0: ensemble rebuilt
    assert sc.ens.n_obj == 2 and list(sc.obj_dict) == [1, 5] and sc.ens.adam_t.cpu().tolist() == [3, 3, 3]
    sc.finish()


def test_checkpoints_through_scene(tmp_path):
    """objects trained through Scene keep accumulating clip / caption features (vmap.py:241-246) and save small,
    self-contained checkpoints with the reference's keys (vmap.py:556-576) that load back into the reference-surface modules."""
    from openobj_b200 import cfg as C, vmap as V
    from openobj_b200.scene import Scene
    from openobj_b200.synthetic import SyntheticScene
    cfg = C.room0_config(w=100, h=60)
    cfg.do_bg = False
    cfg.n_iter_per_frame = 2
    synth = SyntheticScene(5, W=100, H=60, part_mode=True, seed=3, n_distinct=1)
    sc = Scene(cfg, seed=1, max_frames=6)
    for f in range(4):
        s = dict(synth.frame(f))
        s["obj_clip"] = {k: np.full((1, 8), 10.0 * k + f, dtype=np.float32) for k in s["bbox_dict"]}
        s["obj_cap"] = {k: np.full((4,), 100.0 * k + f, dtype=np.float32) for k in s["bbox_dict"]}
        sc.step_frame(s, iters=2)
    torch.cuda.synchronize()
    o = sc.obj_dict[3]
    assert o.clip_feat.shape == (4, 8) and o.clip_feat[:, 0].tolist() == [30.0, 31.0, 32.0, 33.0] and o.feat_cnt == 4
    assert o.caption_feat.shape == (4, 4) and o.caption_feat[:, 0].tolist() == [300.0, 301.0, 302.0, 303.0]
    o.save_checkpoints(str(tmp_path), epoch=4)
    path = os.path.join(str(tmp_path), "obj_3.pth")
    assert os.path.getsize(path) < 400_000                                      # ~130 KB of weights, not the [N, 30720] block
    ck = torch.load(path, weights_only=False)
    assert list(ck["FC_state_dict"].keys()) == layout.NAMES[:18] and ck["clip_feat"].shape == (4, 8)
    fresh = V.sceneObject(cfg, 3, torch.zeros(100, 60, 3, dtype=torch.uint8, device=DEV), torch.zeros(100, 60, device=DEV),
                          torch.zeros(100, 60, dtype=torch.uint8, device=DEV), torch.zeros(4), torch.eye(4), 0)
    assert fresh.load_checkpoints(path) is True
    for a, b in zip(fresh.trainer.fc_occ_map.parameters(), o.trainer.fc_occ_map.parameters()):
        assert torch.equal(a.detach(), b.detach())
