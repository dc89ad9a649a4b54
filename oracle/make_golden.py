"""Freeze golden vectors from the UNMODIFIED reference (run here, CPU) into tests/golden/.

    python oracle/make_golden.py            # needs /root/reference
    python oracle/make_golden.py --check    # regenerate in memory and compare with the committed files (bit for bit)

TEST INFRASTRUCTURE ONLY.  The reference ships no tests or fixtures (SURVEY.md
section 4), so these files are the pins for oracle/openobj_oracle.py and, through it, for
the CUDA path.  Everything is seeded; the script is committed next to its outputs.
"""
from __future__ import annotations

import json
import os
import random
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
warnings.simplefilter("ignore")


class Tape:
    """Record every torch.rand / torch.randint / Tensor.normal_ result, in call order."""

    def __init__(self):
        self.calls = []

    def __enter__(self):
        self._rand, self._randint, self._normal = torch.rand, torch.randint, torch.Tensor.normal_
        tape = self

        def rand(*a, **k):
            r = tape._rand(*a, **k)
            tape.calls.append(("rand", r.clone()))
            return r

        def randint(*a, **k):
            r = tape._randint(*a, **k)
            tape.calls.append(("randint", r.clone()))
            return r

        def normal_(self_, *a, **k):
            r = tape._normal(self_, *a, **k)
            tape.calls.append(("normal", r.clone()))
            return r

        torch.rand, torch.randint, torch.Tensor.normal_ = rand, randint, normal_
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randint, torch.Tensor.normal_ = self._rand, self._randint, self._normal


def small_cfg(W=100, H=60, **kw):
    c = rh.make_cfg()
    c.W, c.H, c.width, c.height = W, H, W, H
    c.fx = c.fy = 50.0
    c.cx, c.cy = (W - 1) / 2.0, (H - 1) / 2.0
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def synth_frame(g, W, H, obj_id=1, frame_id=0, with_part=True):
    rgb = torch.randint(0, 256, (W, H, 3), generator=g, dtype=torch.uint8)
    ww, hh = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="ij")
    depth = (2.0 + 0.01 * ww + 0.02 * hh + 0.3 * torch.rand(W, H, generator=g)).float()
    depth[torch.rand(W, H, generator=g) < 0.05] = 0.0                  # invalid depth
    inst = torch.zeros(W, H, dtype=torch.int32)
    inst[20:70, 10:45] = obj_id
    inst[18:20, 8:47] = -1
    inst[70:72, 8:47] = -1
    inst[40:50, 20:30] = 7                                             # another object inside the bbox
    state = torch.zeros(W, H, dtype=torch.uint8)
    state[inst == obj_id] = 1
    state[inst == -1] = 2
    bbox = torch.tensor([12, 78, 4, 52])                               # [w_lo,w_hi,h_lo,h_hi] int64
    ang = 0.05 * frame_id
    T = torch.eye(4, dtype=torch.float64)
    T[0, 0], T[0, 2], T[2, 0], T[2, 2] = np.cos(ang), np.sin(ang), -np.sin(ang), np.cos(ang)
    T[:3, 3] = torch.tensor([0.1 * frame_id, -0.05 * frame_id, 0.02 * frame_id], dtype=torch.float64)
    part = None
    if with_part:
        part = torch.randint(-8, 9, (W // 5, H // 5, 512), generator=g).float() / 4.0
        part[torch.rand(W // 5, H // 5, generator=g) < 0.1] = 0.0      # SAM gaps
    return rgb, depth, state, bbox, T, part


def stacked_from_modules(m, fcs, pes):
    opt = torch.optim.AdamW([torch.autograd.Variable(torch.tensor(0))], lr=1e-3, weight_decay=0.013)
    fc_model, fc_param, fc_buffer = m["utils"].update_vmap(fcs, opt)
    pe_model, pe_param, pe_buffer = m["utils"].update_vmap(pes, opt)
    return opt, (fc_model, fc_param, fc_buffer), (pe_model, pe_param, pe_buffer)


def gen_model_step(m):
    """forward / loss / backward / AdamW through the reference's vmap path (train.py:394-474)."""
    from functorch import vmap
    torch.manual_seed(1234)
    N, R, S = 3, 16, 10
    cfg = small_cfg()
    cfg.obj_id = 1
    trainers = []
    for k in range(N):
        c = small_cfg()
        c.obj_id = k + 1
        trainers.append(m["trainer"].Trainer(c))
        # perturb PE directions a little so the stacked B differs per object
        trainers[-1].pe.B_layer.weight.data += 0.01 * torch.randn(21, 3)
    fcs = [t.fc_occ_map for t in trainers]
    pes = [t.pe for t in trainers]
    opt, (fc_model, fc_param, fc_buffer), (pe_model, pe_param, pe_buffer) = stacked_from_modules(m, fcs, pes)
    g = torch.Generator().manual_seed(99)
    z = torch.sort(0.5 + 3.0 * torch.rand(N, R, S, generator=g), dim=-1).values
    o = torch.randn(N, R, 1, 3, generator=g) * 0.2
    d = torch.randn(N, R, 1, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    pcs = (o + d * z[..., None]).float()
    gt_depth = (z[..., 6] + 0.05 * torch.randn(N, R, generator=g)).float()
    gt_depth[0, 3] = 0.0
    gt_rgb8 = torch.randint(0, 256, (N, R, 3), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 3, (N, R), generator=g, dtype=torch.uint8)
    labels[:, 0], labels[:, 1], labels[:, 2] = 1, 0, 2
    gt_feat = torch.randint(-8, 9, (N, R, 512), generator=g).float() / 4.0
    gt_feat[1, 5] = 0.0
    mask_depth = gt_depth > 0
    out = dict(pcs=pcs, z=z, gt_depth=gt_depth, gt_rgb8=gt_rgb8, labels=labels, gt_feat=gt_feat)
    for i, p in enumerate(fc_param):
        out["fc%02d" % i] = p.detach().clone()
    out["peB"] = pe_param[0].detach().clone()

    def fwd():
        emb = vmap(pe_model)(pe_param, pe_buffer, pcs)
        a, c, f = vmap(fc_model)(fc_param, fc_buffer, emb)
        return emb, a, c, f

    emb, a, c, f = fwd()
    out.update(emb=emb.detach(), alpha=a.detach(), color=c.detach(), clip=f.detach())
    gt_rgb = gt_rgb8 / 255.
    # ---- part features ON (room_0 default)
    loss_on, _ = m["loss"].step_batch_loss(a, c, gt_depth, gt_rgb, labels, mask_depth, z,
                                           gt_partfeat=gt_feat, pred_partfeat=f)
    loss_on.backward()
    out["loss_on"] = loss_on.detach()
    for i, p in enumerate(list(fc_param) + list(pe_param)):
        out["g_on%02d" % i] = p.grad.detach().clone()
    opt.zero_grad(set_to_none=True)
    # ---- part features OFF: clip head gets no grad (quirk 8)
    emb, a, c, f = fwd()
    loss_off, _ = m["loss"].step_batch_loss(a, c, gt_depth, gt_rgb, labels, mask_depth, z)
    loss_off.backward()
    out["loss_off"] = loss_off.detach()
    none_idx = []
    for i, p in enumerate(list(fc_param) + list(pe_param)):
        if p.grad is None:
            none_idx.append(i)
        else:
            out["g_off%02d" % i] = p.grad.detach().clone()
    out["g_off_none"] = torch.tensor(none_idx)
    opt.zero_grad(set_to_none=True)
    # ---- cross-object zero-mask rule: object 2 has no label==1 ray (quirk 1)
    lab2 = labels.clone()
    lab2[2][lab2[2] == 1] = 0
    emb, a, c, f = fwd()
    loss_zm, _ = m["loss"].step_batch_loss(a, c, gt_depth, gt_rgb, lab2, mask_depth, z,
                                           gt_partfeat=gt_feat, pred_partfeat=f)
    loss_zm.backward()
    out["labels_zm"] = lab2
    out["loss_zm"] = loss_zm.detach()
    for i, p in enumerate(list(fc_param) + list(pe_param)):
        out["g_zm%02d" % i] = (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
    opt.zero_grad(set_to_none=True)
    # ---- 3 optimiser steps, part ON (train.py:472-474)
    losses = []
    for it in range(3):
        emb, a, c, f = fwd()
        l, _ = m["loss"].step_batch_loss(a, c, gt_depth, gt_rgb, labels, mask_depth, z,
                                         gt_partfeat=gt_feat, pred_partfeat=f)
        l.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(l.detach())
    out["losses_3"] = torch.stack(losses)
    for i, p in enumerate(list(fc_param) + list(pe_param)):
        out["p3_%02d" % i] = p.detach().clone()
    # ---- 2 more steps with part OFF: clip head untouched (no decay either)
    for it in range(2):
        emb, a, c, f = fwd()
        l, _ = m["loss"].step_batch_loss(a, c, gt_depth, gt_rgb, labels, mask_depth, z)
        l.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
    for i, p in enumerate(list(fc_param) + list(pe_param)):
        out["p5_%02d" % i] = p.detach().clone()
    return out


def gen_sampling(m, bg=False):
    """sceneObject keyframe ring + get_training_samples with the RNG tape recorded."""
    torch.manual_seed(7 if not bg else 8)
    random.seed(3)
    g = torch.Generator().manual_seed(5 if not bg else 6)
    cfg = small_cfg()
    W, H = cfg.W, cfg.H
    obj_id = 0 if bg else 1
    cam = m["vmap"].cameraInfo(cfg)
    frames = [synth_frame(g, W, H, obj_id=1, frame_id=f) for f in range(8)]
    rgb, depth, state, bbox, T, part = frames[0]
    if bg:
        bbox = torch.tensor([0, W, 0, H])
    obj = m["vmap"].sceneObject(cfg, obj_id, rgb, depth, state, bbox, T, 0)
    gpf = [part]
    for f in range(1, 8):
        rgb, depth, state, bbox, T, part = frames[f]
        if bg:
            bbox = torch.tensor([0, W, 0, H])
        obj.append_keyframe(rgb, depth, state, bbox, T, f * cfg.stride)
        gpf.append(part)
    global_partfeat = torch.stack(gpf)
    nkf = obj.n_keyframes
    n_frames, n_samples = (20, 24) if not bg else (20, 30)
    with Tape() as tape:
        r = obj.get_training_samples(n_frames, n_samples, cam.rays_dir_cache, global_partfeat)
    gt_rgb, gt_depth, valid, labels, pcs, z, pf = r
    out = dict(rgbs_batch=obj.rgbs_batch[:nkf].clone(), depth_batch=obj.depth_batch[:nkf].clone(),
               t_wc_batch=obj.t_wc_batch[:nkf].clone(), bbox=obj.bbox[:nkf].clone(),
               use_frame=torch.from_numpy(obj.use_frame[:nkf].copy()),
               rays_dir=cam.rays_dir_cache.clone(), global_partfeat=global_partfeat,
               n_keyframes=torch.tensor(nkf), latest=torch.tensor(obj.lastest_kf_queue),
               n_c2s=torch.tensor(obj.n_bins_cam2surface), n_bins=torch.tensor(obj.n_bins),
               stride=torch.tensor(cfg.stride), part_down=torch.tensor(cfg.part_down),
               intr=torch.tensor([cfg.fx, cfg.fy, cfg.cx, cfg.cy], dtype=torch.float64),
               gt_rgb=gt_rgb, gt_depth=gt_depth, valid=valid, labels=labels, pcs=pcs, z=z, partfeat=pf)
    # tape layout: randint, rand(w), rand(h), then class tapes that exist
    kinds = [k for k, _ in tape.calls]
    vals = [v for _, v in tape.calls]
    assert kinds[:3] == ["randint", "rand", "rand"], kinds
    kf_draw = vals[0]
    if nkf > 2:
        kf_ids = torch.cat([kf_draw, torch.tensor(obj.lastest_kf_queue[-2:])])
    else:
        kf_ids = kf_draw
    out.update(kf_ids=kf_ids, u_w=vals[1], u_h=vals[2])
    S = obj.n_bins_cam2surface + obj.n_bins
    rest = list(zip(kinds[3:], vals[3:]))
    n_inv = int((gt_depth.reshape(-1) <= 0).sum())
    n_val = gt_depth.numel() - n_inv
    lab = labels.reshape(-1)
    n_obj = int(((lab == 1) & valid).sum())
    n_oth = int(((lab != 1) & valid).sum())
    it = iter(rest)
    out["r_invalid"] = next(it)[1] if n_inv else torch.zeros(0, S)
    out["r_valid"] = next(it)[1] if n_val else torch.zeros(0, obj.n_bins_cam2surface)
    if n_obj:
        k, v = next(it)
        assert k == "normal"
        out["r_normal"] = v
    else:
        out["r_normal"] = torch.zeros(0, obj.n_bins)
    out["r_other"] = next(it)[1] if n_oth else torch.zeros(0, obj.n_bins)
    assert out["r_invalid"].shape[0] == n_inv and out["r_valid"].shape[0] == n_val
    assert out["r_normal"].shape[0] == n_obj and out["r_other"].shape[0] == n_oth
    return out


def gen_keyframe_policy(m):
    """append_keyframe / prune_keyframe bookkeeping over 40 frames (vmap.py:166-257)."""
    random.seed(11)
    cfg = small_cfg(W=10, H=10)
    g = torch.Generator().manual_seed(1)
    W = H = 10
    rgb = torch.zeros(W, H, 3, dtype=torch.uint8)
    depth = torch.ones(W, H)
    state = torch.ones(W, H, dtype=torch.uint8)
    bbox = torch.tensor([0, 9, 0, 9])
    T = torch.eye(4, dtype=torch.float64)
    trace = []
    import io
    import contextlib
    obj = m["vmap"].sceneObject(cfg, 1, rgb, depth, state, bbox, T, 0)

    def snap(fid):
        trace.append(dict(frame=fid, n_keyframes=obj.n_keyframes, kf_pointer=obj.kf_pointer,
                          latest=list(obj.lastest_kf_queue), frame_cnt=obj.frame_cnt,
                          kf_id_dict=[[int(k), int(v)] for k, v in obj.kf_id_dict.items()],
                          use_frame=[float(x) for x in obj.use_frame],
                          slot_mark=[int(obj.rgbs_batch[s, 0, 0, 0]) for s in range(cfg.keyframe_buffer_size)]))

    obj.rgbs_batch[:] = 255
    obj.rgbs_batch[0, 0, 0, 0] = 0
    snap(0)
    for f in range(1, 140):
        rgb = torch.full((W, H, 3), f % 250, dtype=torch.uint8)
        with contextlib.redirect_stdout(io.StringIO()):
            obj.append_keyframe(rgb, depth, state, bbox, T, f * 10)
        snap(f * 10)
    return dict(keyframe_step=cfg.keyframe_step, buffer=cfg.keyframe_buffer_size, seed=11, trace=trace)


def gen_render(m):
    """render_2D_syn for one object over every pixel, jitter recorded (vmap.py:604-685)."""
    torch.manual_seed(21)
    cfg = small_cfg(W=40, H=30)
    cfg.fx = cfg.fy = 30.0
    cfg.cx, cfg.cy = 19.5, 14.5
    W, H = cfg.W, cfg.H
    g = torch.Generator().manual_seed(22)
    rgb, depth, state, bbox, T, part = synth_frame(g, 100, 60, with_part=False)
    rgb, depth, state = rgb[:W, :H].contiguous(), depth[:W, :H].contiguous(), state[:W, :H].contiguous()
    obj = m["vmap"].sceneObject(cfg, 1, rgb, depth, state, torch.tensor([0, W - 1, 0, H - 1]), T, 0)
    # make the network opaque enough that some pixels pass the 0.9 opacity test
    with torch.no_grad():
        obj.trainer.fc_occ_map.out_alpha.bias.fill_(0.6)
    cam = m["vmap"].cameraInfo(cfg)
    bb = m["utils"].BoundingBox()
    ang = 0.3
    bb.R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    bb.center = np.array([0.1, -0.05, 2.0])
    bb.extent = np.array([1.2, 0.9, 1.5])
    obj.bbox_final = True
    obj.bbox3dour = bb
    obj.bbox3d = None
    T_wc = np.eye(4)
    T_wc[:3, 3] = [0.05, 0.02, -0.1]
    mask_in = np.ones([W, H], dtype=bool)
    with Tape() as tape:
        res = obj.render_2D_syn(T_wc, None, cam.rays_dir_cache, chunk_size=300, obj_mask=mask_in, render_part=True)
    obj_mask, rdepth, rcolor, rfeat = res
    assert [k for k, _ in tape.calls] == ["rand"]
    out = dict(T_wc=torch.from_numpy(T_wc), rays_dir=cam.rays_dir_cache.clone(),
               obb_R=torch.from_numpy(bb.R), obb_center=torch.from_numpy(bb.center),
               obb_extent=torch.from_numpy(bb.extent), jitter=tape.calls[0][1],
               mask=torch.from_numpy(obj_mask), depth=torch.from_numpy(rdepth),
               color=torch.from_numpy(rcolor), feat=torch.from_numpy(rfeat),
               intr=torch.tensor([cfg.fx, cfg.fy, cfg.cx, cfg.cy], dtype=torch.float64))
    for i, p in enumerate(obj.trainer.fc_occ_map.parameters()):
        out["fc%02d" % i] = p.detach().clone()[None]
    out["peB"] = obj.trainer.pe.B_layer.weight.detach().clone()[None]
    return out


def gen_bg_step(m):
    """The background model's step (train.py:447-474 with scene_bg): Trainer with hidden_feature_size_bg / bg_scale
    (vmap.py:43-47), non-vmap forward, step_batch_loss on [1, R, S = 5 + 9], backward, AdamW."""
    torch.manual_seed(4321)
    R, S = 24, 14
    base = small_cfg()
    c = small_cfg(hidden_feature_size=base.hidden_feature_size_bg, obj_scale=base.bg_scale)
    c.obj_id = 0
    tr = m["trainer"].Trainer(c)
    tr.pe.B_layer.weight.data += 0.01 * torch.randn(21, 3)
    params = list(tr.fc_occ_map.parameters()) + [tr.pe.B_layer.weight]
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.013)
    g = torch.Generator().manual_seed(77)
    z = torch.sort(0.5 + 5.0 * torch.rand(R, S, generator=g), dim=-1).values
    o = torch.randn(R, 1, 3, generator=g) * 0.3
    d = torch.randn(R, 1, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    pcs = (o + d * z[..., None]).float()
    gt_depth = (z[..., 8] + 0.05 * torch.randn(R, generator=g)).float()
    gt_depth[3] = 0.0
    gt_rgb8 = torch.randint(0, 256, (R, 3), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 3, (R,), generator=g, dtype=torch.uint8)
    labels[0], labels[1], labels[2] = 1, 0, 2
    gt_feat = torch.randint(-8, 9, (R, 512), generator=g).float() / 4.0
    gt_feat[5] = 0.0
    mask_depth = gt_depth > 0
    gt_rgb = gt_rgb8 / 255.
    out = dict(pcs=pcs, z=z, gt_depth=gt_depth, gt_rgb8=gt_rgb8, labels=labels, gt_feat=gt_feat,
               hidden=torch.tensor(c.hidden_feature_size), scale=torch.tensor(float(c.obj_scale)))
    for i, p in enumerate(params):
        out["p%02d" % i] = p.detach().clone()

    def step_loss(part):
        emb = tr.pe(pcs)
        a, col, f = tr.fc_occ_map(emb)
        if part:
            l, _ = m["loss"].step_batch_loss(a[None], col[None], gt_depth[None], gt_rgb[None], labels[None], mask_depth[None],
                                             z[None], gt_partfeat=gt_feat[None], pred_partfeat=f[None])
        else:
            l, _ = m["loss"].step_batch_loss(a[None], col[None], gt_depth[None], gt_rgb[None], labels[None], mask_depth[None],
                                             z[None])
        return emb, a, col, f, l

    emb, a, col, f, l = step_loss(True)
    out.update(emb=emb.detach()[:4], alpha=a.detach(), color=col.detach(), clip=f.detach()[:2], loss_on=l.detach())
    l.backward()
    for i, p in enumerate(params):
        out["g_on%02d" % i] = p.grad.detach().clone()
    opt.zero_grad(set_to_none=True)
    _, _, _, _, l = step_loss(False)
    out["loss_off"] = l.detach()
    l.backward()
    none_idx = []
    for i, p in enumerate(params):
        if p.grad is None:
            none_idx.append(i)
        else:
            out["g_off%02d" % i] = p.grad.detach().clone()
    out["g_off_none"] = torch.tensor(none_idx)
    opt.zero_grad(set_to_none=True)
    losses = []
    for it in range(3):                                  # 2 steps part ON, then 1 step part OFF (clip head untouched)
        _, _, _, _, l = step_loss(it < 2)
        l.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(l.detach())
    out["losses_3"] = torch.stack(losses)
    for i, p in enumerate(params):
        out["q3_%02d" % i] = p.detach().clone()
    return out


def gen_eval_grid(m):
    """Trainer.eval_points on the query grid of Trainer.meshing (trainer.py:46-69,104-128): make_3D_grid with the
    oriented box's scale / transform minus obj_center, then pe -> fc_occ_map -> occupancy_activation in chunks, for an
    object model (hidden 32, scale 2, bound_extent 0.9) and the background model (hidden 128, scale 5, 0.995)."""
    torch.manual_seed(99)
    out = {}
    base = small_cfg()
    ang = 0.4
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    center, extent = np.array([0.3, -0.2, 2.5]), np.array([1.6, 1.1, 0.8])
    obj_center = torch.tensor([0.05, -0.02, 0.1])
    out.update(obb_R=torch.from_numpy(R), obb_center=torch.from_numpy(center), obb_extent=torch.from_numpy(extent),
               obj_center=obj_center)
    for tag, cfg, dim in (("obj", small_cfg(), 12),
                          ("bg", small_cfg(hidden_feature_size=base.hidden_feature_size_bg, obj_scale=base.bg_scale), 7)):
        cfg.obj_id = 1 if tag == "obj" else 0
        tr = m["trainer"].Trainer(cfg)
        tr.pe.B_layer.weight.data += 0.01 * torch.randn(21, 3)
        with torch.no_grad():
            tr.fc_occ_map.out_alpha.bias.fill_(0.02)
        # the lines of Trainer.meshing that build the query (trainer.py:50-64)
        scene_scale_np = extent / (2.0 * tr.bound_extent)
        scene_scale = torch.from_numpy(scene_scale_np).float()
        transform_np = np.eye(4, dtype=np.float32)
        transform_np[:3, 3] = center
        transform_np[:3, :3] = R
        grid_pc = m["render_rays"].make_3D_grid(occ_range=[-1., 1.], dim=dim, device="cpu", scale=scene_scale,
                                                transform=torch.from_numpy(transform_np)).view(-1, 3)
        grid_pc -= obj_center
        occ, color, clip = tr.eval_points(grid_pc, chunk_size=500)
        out[tag + "_dim"] = torch.tensor(dim)
        out[tag + "_bound_extent"] = torch.tensor(tr.bound_extent)
        out[tag + "_scale"] = torch.tensor(float(cfg.obj_scale))
        out[tag + "_grid"] = grid_pc.clone()
        out[tag + "_occ"], out[tag + "_color"] = occ.clone(), color.clone()
        out[tag + "_clip"] = clip[::7].clone()                      # every 7th point keeps the file small
        for i, p in enumerate(tr.fc_occ_map.parameters()):
            out["%s_fc%02d" % (tag, i)] = p.detach().clone()[None]
        out[tag + "_peB"] = tr.pe.B_layer.weight.detach().clone()[None]
    return out


def gen_checkpoint(m):
    """A per-object checkpoint written by the reference's sceneObject.save_checkpoints (vmap.py:556-576) and what the
    reference's Trainer.eval_points returns for it at a few points: the drop-in must load the file and reproduce them."""
    torch.manual_seed(31)
    cfg = small_cfg(W=40, H=30)
    g = torch.Generator().manual_seed(32)
    rgb, depth, state, bbox, T, part = synth_frame(g, 100, 60, with_part=False)
    W, H = cfg.W, cfg.H
    obj = m["vmap"].sceneObject(cfg, 7, rgb[:W, :H].contiguous(), depth[:W, :H].contiguous(), state[:W, :H].contiguous(),
                                torch.tensor([0, W - 1, 0, H - 1]), T, 0)
    obj.trainer.pe.B_layer.weight.data += 0.01 * torch.randn(21, 3)
    obj.clip_feat = torch.randn(512)
    obj.caption_feat = torch.randn(384)
    obj.semantic_id = 3
    obj.bbox3dour = None                       # an open3d / utils.BoundingBox object in a real run: not needed here
    obj.save_checkpoints(OUT, 42)              # -> tests/golden/obj_7.pth
    pts = torch.randn(64, 3, generator=g) * 0.8
    occ, color, clip = obj.trainer.eval_points(pts)
    return dict(points=pts, occ=occ, color=color, clip=clip[:8], clip_feat=obj.clip_feat, caption_feat=obj.caption_feat)


def gen_surface(m):
    """The stand-alone helpers of the surface: utils.{stratified_bins, normal_bins_sampling, origin_dirs_W,
    ray_box_intersection}, Trainer.sample_points_bbox, sceneObject.sample_3d_points -- run on the CPU reference with every
    random draw recorded."""
    torch.manual_seed(55)
    U = m["utils"]
    g = torch.Generator().manual_seed(56)
    out = {}
    # stratified_bins: per-ray bounds, scalar bounds
    mn, mx = torch.rand(33, generator=g), 1.0 + 3.0 * torch.rand(33, generator=g)
    with Tape() as t:
        z = U.stratified_bins(mn, mx, 7, 33, device="cpu")
    out.update(sb_min=mn, sb_max=mx, sb_u=t.calls[0][1], sb_z=z)
    with Tape() as t:
        z = U.stratified_bins(0.0, 3.5, 10, 20, device="cpu")
    out.update(sbs_u=t.calls[0][1], sbs_z=z)
    # normal_bins_sampling
    dep = 1.0 + 2.0 * torch.rand(25, generator=g)
    with Tape() as t:
        z = U.normal_bins_sampling(dep, 9, 25, 0.1, device="cpu")
    out.update(nb_depth=dep, nb_draws=t.calls[0][1], nb_z=z)
    # origin_dirs_W, both shapes
    def rigid(k):
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        T = torch.eye(4)
        T[:3, :3], T[:3, 3] = q, torch.randn(3, generator=g)
        return T
    T = torch.stack([rigid(k) for k in range(4)])
    d1, d2 = torch.randn(4, 3, generator=g), torch.randn(4, 17, 3, generator=g)
    o1, w1 = U.origin_dirs_W(T, d1)
    o2, w2 = U.origin_dirs_W(T, d2)
    out.update(od_T=T, od_d1=d1, od_d2=d2, od_o1=o1, od_w1=w1, od_o2=o2, od_w2=w2)
    # ray_box_intersection
    ro, rd = torch.randn(60, 3, generator=g) * 2.0, torch.randn(60, 3, generator=g)
    bmin, bmax = torch.tensor([-0.7, -0.5, -0.9]), torch.tensor([0.8, 0.6, 0.4])
    near, far, hit = U.ray_box_intersection(ro, rd, bmin, bmax)
    out.update(rb_o=ro, rb_d=rd, rb_min=bmin, rb_max=bmax, rb_near=near, rb_far=far, rb_hit=hit)
    # Trainer.sample_points_bbox
    cfg = small_cfg(W=16, H=12)
    cfg.fx = cfg.fy = 14.0
    cfg.cx, cfg.cy = 7.5, 5.5
    cfg.obj_id = 1
    tr = m["trainer"].Trainer(cfg)
    cam = m["vmap"].cameraInfo(cfg)
    pose = torch.eye(4)
    pose[:3, 3] = torch.tensor([0.05, -0.02, -0.3])
    tr.dirs_C_gt = cam.rays_dir_cache.reshape(-1, 3).clone()               # one pose per ray, as render_2D_syn sets them
    tr.T_WC_gt = pose.unsqueeze(0).repeat_interleave(tr.dirs_C_gt.shape[0], dim=0)   # (vmap.py:618-622)
    bb = U.BoundingBox()
    ang = 0.25
    bb.R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    bb.center, bb.extent = np.array([0.1, 0.0, 2.0]), np.array([1.0, 0.8, 1.2])
    with Tape() as t:
        hit, near, far = tr.sample_points_bbox(bb, do_eval=True)
    assert [k for k, _ in t.calls] == ["rand"]
    out.update(spb_T=tr.T_WC_gt, spb_dirs=tr.dirs_C_gt, spb_R=torch.from_numpy(bb.R), spb_center=torch.from_numpy(bb.center),
               spb_extent=torch.from_numpy(bb.extent), spb_u=t.calls[0][1], spb_hit=hit, spb_near=near, spb_far=far,
               spb_zcat=tr.z_vals_cat, spb_z=tr.z_vals, spb_pcs=tr.input_pcs, spb_dirsW=tr.dirs_W, spb_origins=tr.origins)
    # sceneObject.sample_3d_points
    cfg2 = small_cfg(W=40, H=30)
    rgb, depth, state, bbox, Tp, part = synth_frame(g, 100, 60, with_part=False)
    obj = m["vmap"].sceneObject(cfg2, 1, rgb[:40, :30].contiguous(), depth[:40, :30].contiguous(), state[:40, :30].contiguous(),
                                torch.tensor([0, 39, 0, 29]), Tp, 0)
    F_, P_ = 3, 8
    srgb = torch.randint(0, 256, (F_, P_, 4), generator=g, dtype=torch.uint8)
    srgb[..., 3] = torch.randint(0, 3, (F_, P_), generator=g, dtype=torch.uint8)
    sdep = (1.0 + 2.0 * torch.rand(F_, P_, generator=g)).float()
    sdep[0, 1] = 0.0
    sdep[2, 5] = 0.0
    org, dw = torch.randn(F_, 3, generator=g) * 0.2, torch.randn(F_, P_, 3, generator=g)
    with Tape() as t:
        res = obj.sample_3d_points(srgb, sdep, org, dw)
    assert [k for k, _ in t.calls] == ["rand", "rand", "normal", "rand"]
    out.update(s3_rgbs=srgb, s3_depth=sdep, s3_origins=org, s3_dirs=dw, s3_u_inv=t.calls[0][1], s3_u_val=t.calls[1][1],
               s3_n_obj=t.calls[2][1], s3_u_oth=t.calls[3][1], s3_valid=res[2], s3_labels=res[3], s3_pcs=res[4], s3_z=res[5])
    return out


def save(name, d):
    arrs = {k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrs)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def check(m):
    """--check: regenerate every fixture in memory and compare it with the committed file (bit for bit)."""
    gens = [("bg_step.npz", gen_bg_step), ("model_step.npz", gen_model_step), ("sample_obj.npz", lambda mm: gen_sampling(mm, bg=False)),
            ("sample_bg.npz", lambda mm: gen_sampling(mm, bg=True)), ("render_obj.npz", gen_render), ("eval_grid.npz", gen_eval_grid),
            ("surface.npz", gen_surface)]
    bad = 0
    for name, fn in gens:
        new = {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in fn(m).items()}
        old = np.load(os.path.join(OUT, name))
        same = sorted(new) == sorted(old.files) and all(
            new[k].shape == old[k].shape and np.array_equal(new[k], old[k], equal_nan=True) for k in new)
        print("%-18s %s" % (name, "identical" if same else "DIFFERS"))
        bad += not same
    return bad


def main():
    os.makedirs(OUT, exist_ok=True)
    m = rh.load()
    if "--check" in sys.argv:
        sys.exit(check(m))
    if "--only-eval" in sys.argv:               # added after the other files were frozen: do not touch them
        save("eval_grid.npz", gen_eval_grid(m))
        return
    if "--only-surface" in sys.argv:
        save("surface.npz", gen_surface(m))
        return
    if "--only-ckpt" in sys.argv:
        save("ckpt_expect.npz", gen_checkpoint(m))
        return
    if "--only-bg" in sys.argv:                 # added after the other files were frozen: do not touch them
        save("bg_step.npz", gen_bg_step(m))
        return
    save("bg_step.npz", gen_bg_step(m))
    save("model_step.npz", gen_model_step(m))
    save("sample_obj.npz", gen_sampling(m, bg=False))
    save("sample_bg.npz", gen_sampling(m, bg=True))
    save("render_obj.npz", gen_render(m))
    save("eval_grid.npz", gen_eval_grid(m))
    save("ckpt_expect.npz", gen_checkpoint(m))
    save("surface.npz", gen_surface(m))
    with open(os.path.join(OUT, "keyframe_policy.json"), "w") as f:
        json.dump(gen_keyframe_policy(m), f)
    print("done; torch", torch.__version__)


if __name__ == "__main__":
    main()
