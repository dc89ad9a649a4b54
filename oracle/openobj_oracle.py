"""CPU restatement ("port") of OpenObj's per-object NeRF training hot path.

TEST INFRASTRUCTURE ONLY -- this file is the *checker*, never the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it.  Nothing under ``openobj_b200/`` does.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the unmodified reference
modules (imported from /root/reference through ``oracle/ref_harness.py``) on
seeded synthetic inputs and freezes inputs + outputs into ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those
files, and ``tests/test_oracle_vs_reference.py`` re-checks live against the
reference whenever /root/reference is present.  The reference itself ships no
tests or golden vectors, so these are the only pins.

All arithmetic is float32 torch on CPU, written functionally over *stacked*
per-object tensors ``[N, ...]`` (the reference stacks nn.Modules with functorch).
Each function cites the reference lines it follows (paths relative to
/root/reference/objnerf).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

# --------------------------------------------------------------------------
# parameter layout: the 18 OccupancyMap tensors + the PE direction matrix, in
# named_parameters() order (model.py:31-56, embedding.py:39).
# --------------------------------------------------------------------------

N_DIRS = 21


def fc_shapes(hidden=32, clip=512, n_bands=6):
    e = 3 + N_DIRS * n_bands          # 129 (trainer.py:20-21 with n_unidir_funcs=5)
    e1 = N_DIRS * 4 + 3               # 87  (trainer.py:20)
    e2 = e - e1                       # 42  (trainer.py:21)
    h = hidden
    return [
        ("in_layer.0.weight", (h, e1)), ("in_layer.0.bias", (h,)),
        ("mid1.0.0.weight", (h, h)), ("mid1.0.0.bias", (h,)),
        ("cat_layer.0.weight", (h, h + e1)), ("cat_layer.0.bias", (h,)),
        ("mid2.0.0.weight", (h, h)), ("mid2.0.0.bias", (h,)),
        ("out_alpha.weight", (1, h)), ("out_alpha.bias", (1,)),
        ("color_linear.0.weight", (h, h + e2)), ("color_linear.0.bias", (h,)),
        ("out_color.weight", (3, h)), ("out_color.bias", (3,)),
        ("clip_linear.0.weight", (h, h + e2)), ("clip_linear.0.bias", (h,)),
        ("out_clip.weight", (clip, h)), ("out_clip.bias", (clip,)),
    ]


ICOSA_DIRS = np.array([
    # the 21 fixed projection directions the reference initialises B_layer with
    # (embedding.py:15-37); values are data, quoted to the printed precision.
    0.8506508, 0, 0.5257311, 0.809017, 0.5, 0.309017, 0.5257311, 0.8506508, 0,
    1, 0, 0, 0.809017, 0.5, -0.309017, 0.8506508, 0, -0.5257311,
    0.309017, 0.809017, -0.5, 0, 0.5257311, -0.8506508, 0.5, 0.309017, -0.809017,
    0, 1, 0, -0.5257311, 0.8506508, 0, -0.309017, 0.809017, -0.5,
    0, 0.5257311, 0.8506508, -0.309017, 0.809017, 0.5, 0.309017, 0.809017, 0.5,
    0.5, 0.309017, 0.809017, 0.5, -0.309017, 0.809017, 0, 0, 1,
    -0.5, 0.309017, 0.809017, -0.809017, 0.5, 0.309017, -0.809017, 0.5, -0.309017,
], dtype=np.float32).reshape(21, 3)


def init_params(n_obj, hidden=32, clip=512, n_bands=6, generator=None):
    """Random init with the reference's distributions (model.py:4-6 xavier-normal
    weights; nn.Linear default U(-1/sqrt(fan_in), 1/sqrt(fan_in)) biases;
    embedding.py:39-40 fixed directions).  Returns (fc list of [N,...], B [N,21,3]).
    The *values* differ from a reference run (different RNG consumption order);
    parity tests copy tensors from the reference instead."""
    g = generator
    fc = []
    shapes = fc_shapes(hidden, clip, n_bands)
    fan_in_of = {}
    for name, shp in shapes:
        if name.endswith("weight"):
            fan_out, fan_in = shp
            std = math.sqrt(2.0 / (fan_in + fan_out))
            fc.append(torch.randn((n_obj,) + shp, generator=g) * std)
            fan_in_of[name[:-6]] = fan_in
        else:
            bound = 1.0 / math.sqrt(fan_in_of[name[:-4]])
            fc.append((torch.rand((n_obj,) + shp, generator=g) * 2 - 1) * bound)
    B = torch.from_numpy(ICOSA_DIRS).clone()[None].repeat(n_obj, 1, 1)
    return fc, B


# --------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------

def pe_forward(x, B, scale, n_bands=6):
    """UniDirsEmbed.forward (embedding.py:46-55) over stacked objects.
    x [N,...,3]; B [N,21,3]; scale python float or [N] tensor.
    Output [N,...,3+21*n_bands], frequency-major after the 3 scaled coords."""
    N = x.shape[0]
    if torch.is_tensor(scale):
        sc = scale.reshape([N] + [1] * (x.dim() - 1)).to(x.dtype)
    else:
        sc = torch.tensor(float(scale), dtype=x.dtype)
    t = x / sc                                              # :47
    flat = t.reshape(N, -1, 3)
    proj = torch.bmm(flat, B.transpose(1, 2))               # :48  nn.Linear(3,21,bias=False)
    bands = (2.0 ** torch.linspace(0, n_bands - 1, n_bands)).to(x.dtype)   # :42
    pb = proj[:, :, None, :] * bands[None, None, :, None]   # :49
    xb = pb.reshape(N, flat.shape[1], -1)                   # :50
    emb = torch.sin(xb * np.pi)                             # :52
    out = torch.cat([flat, emb], dim=-1)                    # :53
    return out.reshape(list(x.shape[:-1]) + [out.shape[-1]])


def _lin(x, W, b):
    # x [N,M,K], W [N,J,K], b [N,J]
    return torch.baddbmm(b[:, None, :], x, W.transpose(1, 2))


def mlp_forward(fc, emb, e1=87):
    """OccupancyMap.forward (model.py:61-103) with the default flags.
    fc = 18 stacked tensors; emb [N,...,E].  Returns alpha [...,1], color [...,3], clip [...,C]."""
    (W_in, b_in, W_m1, b_m1, W_cat, b_cat, W_m2, b_m2, W_a, b_a,
     W_cl, b_cl, W_oc, b_oc, W_cp, b_cp, W_ocp, b_ocp) = fc
    N = emb.shape[0]
    lead = list(emb.shape[:-1])
    x = emb.reshape(N, -1, emb.shape[-1])
    xa, xb = x[..., :e1], x[..., e1:]
    fc1 = torch.relu(_lin(xa, W_in, b_in))                         # :69
    fc2 = torch.relu(_lin(fc1, W_m1, b_m1))                        # :70
    fc3 = torch.relu(_lin(torch.cat((fc2, xa), -1), W_cat, b_cat))  # :73-74
    fc4 = torch.relu(_lin(fc3, W_m2, b_m2))                        # :77
    alpha = _lin(fc4, W_a, b_a) * 10.0                             # :81,88
    hx = torch.cat((fc4, xb), -1)
    color = torch.sigmoid(_lin(torch.relu(_lin(hx, W_cl, b_cl)), W_oc, b_oc))   # :94-96
    clip = _lin(torch.relu(_lin(hx, W_cp, b_cp)), W_ocp, b_ocp)    # :100-101
    return (alpha.reshape(lead + [1]), color.reshape(lead + [3]),
            clip.reshape(lead + [clip.shape[-1]]))


def termination(alpha):
    """occupancy_activation + occupancy_to_termination(is_batch) (render_rays.py:6-14, 32-54).
    alpha [...,S] -> (occ, T) with T_i = occ_i * prod_{j<i} (1 - occ_j + 1e-10)."""
    occ = torch.sigmoid(alpha)
    free = (1.0 - occ + 1e-10)[..., :-1]
    free = torch.cat([torch.ones_like(occ[..., :1]), free], dim=-1)
    return occ, occ * torch.cumprod(free, dim=-1)


def cosine(x, y, eps=1e-8):
    """F.cosine_similarity(dim=-1) as executed by torch>=1.12 (render_rays.py:75):
    x.y / (max(|x|,eps) * max(|y|,eps))."""
    nx = torch.linalg.vector_norm(x, dim=-1).clamp_min(eps)
    ny = torch.linalg.vector_norm(y, dim=-1).clamp_min(eps)
    return (x * y).sum(-1) / (nx * ny)


@dataclass
class LossTerms:
    total: torch.Tensor          # scalar: sum over objects (loss.py:101)
    depth: torch.Tensor          # [N] per-object, after the cross-object zero-mask rule
    color: torch.Tensor
    opacity: torch.Tensor
    feat: torch.Tensor           # zeros when part features are off
    flags: int                   # bit0 explode (>1e5), bit1 some object has no label==1 ray, bit2 no label!=2 ray
    render_depth: torch.Tensor
    render_color: torch.Tensor
    render_opacity: torch.Tensor


def _reduce(loss_mat, mask, weight=None):
    """reduce_batch_loss (render_rays.py:85-117), avg=True, mask given.
    Returns ([N] losses, zero_mask_hit, explode)."""
    mask_num = mask.sum(-1)
    if bool((mask_num == 0).any()):                      # :89-94 quirk 1: zero for ALL objects, no grad
        return torch.zeros(loss_mat.shape[0], dtype=loss_mat.dtype), True, False
    lw = loss_mat * weight if weight is not None else loss_mat
    out = lw.sum(-1) / (mask_num + 1e-10)                # :108
    return out, False, bool((out > 100000).any())        # :109-111 (reference exits)


def step_loss(alpha, color, gt_depth, gt_color, labels, z, gt_feat=None, pred_feat=None,
              color_scaling=5.0, opacity_scaling=10.0, feat_scaling=5.0):
    """loss.step_batch_loss (loss.py:5-103).  alpha [N,R,S(,1)], color [N,R,S,3],
    gt_depth [N,R], gt_color [N,R,3] in [0,1], labels [N,R] (0 other/1 this/2 unknown),
    z [N,R,S], gt_feat [N,R,C], pred_feat [N,R,S,C].  mask_depth is accepted by the
    reference but unused (quirk 2) so it is not a parameter here."""
    if alpha.dim() == 4:
        alpha = alpha.squeeze(-1)
    mask_obj = labels != 0                                # :16
    mask_sem = labels != 2                                # :20
    m_both = (mask_obj & mask_sem)                        # == (labels == 1)
    occ, T = termination(alpha)                           # :27-29
    d = (T * z).sum(-1)                                   # :31
    var = (T * (z - d[..., None]) ** 2).sum(-1).detach()  # :32-33
    col = (T[..., None] * color).sum(-2)                  # :34
    opac = T.sum(-1)                                      # :35
    flags = 0
    w = 1.0 / (torch.sqrt(var) + 1e-4)                    # render_rays.py:95-100
    l_d, z1, ex = _reduce((d - gt_depth).abs() * m_both, m_both, w)           # :41-49
    flags |= (2 if z1 else 0) | (1 if ex else 0)
    l_c, z1, ex = _reduce((col - gt_color).abs().sum(-1) * m_both, m_both)    # :53-63
    flags |= (2 if z1 else 0) | (1 if ex else 0)
    l_o, z2, ex = _reduce((opac - mask_obj.float()).abs() * mask_sem, mask_sem)  # :71-75
    flags |= (4 if z2 else 0) | (1 if ex else 0)
    per_obj = l_d + l_c * color_scaling + l_o * opacity_scaling              # :79
    l_f = torch.zeros_like(l_d)
    if gt_feat is not None:
        rf = (T[..., None] * pred_feat).sum(-2)                               # :82
        l_f, z1, ex = _reduce((1.0 - cosine(rf, gt_feat)) * m_both, m_both)   # :87-92
        flags |= (2 if z1 else 0) | (1 if ex else 0)
        per_obj = per_obj + l_f * feat_scaling                                # :99
    return LossTerms(per_obj.sum(), l_d, l_c, l_o, l_f, flags, d, col, opac)


# --------------------------------------------------------------------------
# evaluation at free points / on the meshing grid (trainer.py:46-69,104-128; render_rays.py:119-146)
# --------------------------------------------------------------------------

def make_3d_grid(dim, scale=None, transform=None, occ_range=(-1.0, 1.0)):
    """render_rays.make_3D_grid (render_rays.py:119-146): meshgrid of linspace knots, * scale, rows of the transform
    applied as (R_r * grid).sum(-1), + translation.  Returns [dim, dim, dim, 3]."""
    t = torch.linspace(occ_range[0], occ_range[1], steps=dim)                   # :123
    g = torch.stack(torch.meshgrid(t, t, t, indexing="ij"), dim=3)             # :124-129
    if scale is not None:
        g = g * scale                                                           # :131-132
    if transform is not None:
        rows = [(transform[r, :3][None, None, None] * g).sum(-1, keepdim=True) for r in range(3)]   # :134-140
        g = torch.cat(rows, dim=-1) + transform[:3, 3][None, None, None]        # :141-144
    return g


def meshing_grid(obb_R, obb_center, obb_extent, bound_extent, dim, obj_center=None):
    """The query points of Trainer.meshing (trainer.py:50-64): scale = extent / (2 * bound_extent) in float64 -> float32,
    float32 4x4 [R | center], make_3D_grid, minus obj_center.  Returns [dim^3, 3]."""
    scale = (obb_extent.double() / (2.0 * float(bound_extent))).float()
    tr = torch.eye(4, dtype=torch.float32)
    tr[:3, 3] = obb_center.float()
    tr[:3, :3] = obb_R.float()
    pts = make_3d_grid(dim, scale, tr).reshape(-1, 3)
    return pts if obj_center is None else pts - obj_center


def eval_points(fc1, B1, points, scale=2.0):
    """Trainer.eval_points (trainer.py:104-128) for one model: fc1 = 18 tensors with a leading 1, B1 [1,21,3],
    points [n,3].  Returns occ [n] = sigmoid(alpha) (render_rays.py:6-14), color [n,3], clip [n,C]."""
    alpha, color, clip = ensemble_forward(fc1, B1, points[None], scale)
    return torch.sigmoid(alpha[0, :, 0]), color[0], clip[0]


# --------------------------------------------------------------------------
# one optimisation step of the ensemble (train.py:394-474) and AdamW
# --------------------------------------------------------------------------

def ensemble_forward(fc, B, pcs, scale=2.0, n_bands=6):
    """vmap(pe_model) then vmap(fc_model) (train.py:424-425)."""
    emb = pe_forward(pcs, B, scale, n_bands)
    return mlp_forward(fc, emb)


def train_step_grads(fc, B, pcs, z, gt_depth, gt_rgb01, labels, gt_feat=None, scale=2.0):
    """forward + loss + backward for one iteration.  Returns (LossTerms, grads) where grads
    is a list of 19 tensors (18 fc + B); an entry is None when the reference's autograd
    would leave ``.grad`` None (clip head with part features off: SURVEY A.4 / quirk 8)."""
    leaves = [p.detach().clone().requires_grad_(True) for p in list(fc) + [B]]
    alpha, color, clip = ensemble_forward(leaves[:18], leaves[18], pcs, scale)
    terms = step_loss(alpha, color, gt_depth, gt_rgb01, labels, z,
                      gt_feat=gt_feat, pred_feat=clip if gt_feat is not None else None)
    if terms.total.requires_grad:
        grads = torch.autograd.grad(terms.total, leaves, allow_unused=True)
    else:
        grads = [None] * len(leaves)
    return terms, list(grads)


def adamw_step(p, g, m, v, step, lr=1e-3, wd=0.013, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.AdamW single-tensor update as executed (train.py:78,473).
    In place on p, m, v.  `step` is the 1-based step count *after* increment."""
    p.mul_(1.0 - lr * wd)
    m.lerp_(g, 1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


# --------------------------------------------------------------------------
# sampling (vmap.py:386-554, utils.py:324-397) driven by an explicit RNG tape
# --------------------------------------------------------------------------

@dataclass
class SampleTape:
    """Every random draw get_training_samples makes for ONE object, in reference order
    (SURVEY A.5).  Rows of the class tapes are consumed by rank within the class mask."""
    kf_ids: torch.Tensor      # int64 [n_frames]  (randint draws with the forced latest two appended)
    u_w: torch.Tensor         # f32 [n_frames, n_samples]
    u_h: torch.Tensor         # f32 [n_frames, n_samples]
    r_invalid: torch.Tensor   # f32 [>=n_invalid, S]       rand
    r_valid: torch.Tensor     # f32 [>=n_valid, n_c2s]     rand
    r_normal: torch.Tensor    # f32 [>=n_obj, n_bins]      normal_(0, eps/3), unsorted
    r_other: torch.Tensor     # f32 [>=n_other, n_bins]    rand


def stratified(min_d, max_d, n_bins, u):
    """utils.stratified_bins (utils.py:342-379) with the uniform draws `u` [n, n_bins] supplied."""
    n = u.shape[0]
    lim = torch.linspace(0, 1, n_bins + 1, dtype=torch.float32)
    if not torch.is_tensor(min_d):
        min_d = torch.ones(n, dtype=torch.float32) * min_d
    if not torch.is_tensor(max_d):
        max_d = torch.ones(n, dtype=torch.float32) * max_d
    rng = max_d - min_d
    lower = (rng[..., None] * lim + min_d[..., None])[:, :-1]
    return lower + u * (rng / n_bins)[..., None]


def normal_bins(depth, draws, delta):
    """utils.normal_bins_sampling (utils.py:382-397) with the normal_(0, delta/3) draws [n, n_bins] supplied."""
    return depth[:, None] + torch.clip(draws.sort().values, -delta, delta)


def origin_dirs_w(T_wc, dirs_c):
    """utils.origin_dirs_W (utils.py:324-336)."""
    if dirs_c.shape[1] == 3 and dirs_c.dim() == 2:
        dirs_w = torch.matmul(T_wc[:, :3, :3], dirs_c.unsqueeze(-1)).squeeze(-1)
    else:
        dirs_w = (T_wc[:, None, :3, :3] @ dirs_c[..., None]).squeeze(-1)
    return T_wc[:, :3, -1], dirs_w


def sample_object(rgbs, depth, t_wc, bbox, rays_dir, tape: SampleTape, n_c2s=1, n_bins=9,
                  eps=0.1, other_eps=0.05, min_bound=0.0, use_frame=None, stride=10, part_down=5):
    """sceneObject.get_training_samples + sample_3d_points for one object.
    rgbs u8 [KF,W,H,4] (rgb + state), depth f32 [KF,W,H], t_wc f32 [KF,4,4],
    bbox f32 [KF,4] = [w_lo,w_hi,h_lo,h_hi], rays_dir f32 [W,H,3].
    Returns dict with kf, iw, ih (int64), rgb u8 [F,S,3], depth [F,S], valid [F*S] bool,
    labels u8 [F*S], pcs [F,S,Z,3], z [F,S,Z], and part index triplet (pf, pw, ph) if use_frame."""
    kf = tape.kf_ids.reshape(-1, 1)                                           # vmap.py:411
    # separate mul / add (no FMA), then truncation -- vmap.py:418-422
    iwf = tape.u_w * (bbox[kf, 1] - bbox[kf, 0]) + bbox[kf, 0]
    ihf = tape.u_h * (bbox[kf, 3] - bbox[kf, 2]) + bbox[kf, 2]
    iw, ih = iwf.long(), ihf.long()
    s_rgbs = rgbs[kf, iw, ih]                                                 # :424
    s_depth = depth[kf, iw, ih]                                               # :425
    dirs_c = rays_dir[iw, ih]                                                 # :428
    T = t_wc[kf[:, 0]]                                                        # :431
    dirs_w = (T[:, None, :3, :3] @ dirs_c[..., None]).squeeze(-1)             # utils.py:334
    origins = T[:, :3, -1]                                                    # utils.py:335
    out = {}
    if use_frame is not None:                                                 # :437-452
        uf = torch.as_tensor(use_frame, dtype=torch.float64)
        out["pf"] = (uf[kf] / stride).long().expand_as(iw).contiguous()
        out["pw"] = torch.floor(iwf / part_down).long()
        out["ph"] = torch.floor(ihf / part_down).long()
    S = n_c2s + n_bins
    n_rays = iw.numel()
    z = torch.zeros(n_rays, S, dtype=torch.float32)                           # :479-483
    dflat = s_depth.reshape(-1)
    state = s_rgbs[..., -1].reshape(-1)
    invalid = dflat <= min_bound                                              # :485
    max_bound = s_depth.max()                                                 # :489
    n_inv = int(invalid.sum())
    if n_inv:                                                                 # :493-498
        z[invalid] = stratified(min_bound, max_bound, S, tape.r_invalid[:n_inv])
    valid = ~invalid
    n_val = int(valid.sum())
    if n_val:
        z[valid, :n_c2s] = stratified(min_bound, dflat[valid] - eps, n_c2s, tape.r_valid[:n_val])   # :506-509
        m_obj = (state == 1) & valid                                          # :512
        n_o = int(m_obj.sum())
        if n_o:                                                               # :516-529, utils.py:382-397
            bins = tape.r_normal[:n_o].sort(dim=-1).values
            bins = torch.clip(bins, -eps, eps)
            z[m_obj, n_c2s:] = dflat[m_obj][:, None] + bins
        m_oth = (state != 1) & valid                                          # :536
        n_t = int(m_oth.sum())
        if n_t:                                                               # :538-542
            z[m_oth, n_c2s:] = stratified(dflat[m_oth] - eps, dflat[m_oth] + other_eps, n_bins,
                                          tape.r_other[:n_t])
    z = z.view(iw.shape[0], iw.shape[1], S)
    pcs = origins[:, None, None, :] + dirs_w[:, :, None, :] * z[..., None]    # :548-549 (obj_center = 0)
    out.update(kf=kf[:, 0].clone(), iw=iw, ih=ih, rgb=s_rgbs[..., :3], depth=s_depth,
               valid=valid, labels=state.clone(), pcs=pcs, z=z)
    return out


# --------------------------------------------------------------------------
# eval: ray / OBB slab test, 150-bin stratified midpoints, compositing, z-merge
# (trainer.py:130-198, utils.py:309-319, vmap.py:604-685, train.py:577-594)
# --------------------------------------------------------------------------

def ray_box(origins, dirs, bmin, bmax):
    """utils.ray_box_intersection (utils.py:309-319)."""
    tmin = (bmin - origins) / dirs
    tmax = (bmax - origins) / dirs
    t1 = torch.min(tmin, tmax)
    t2 = torch.max(tmin, tmax)
    near = torch.amax(t1, dim=1)
    far = torch.amin(t2, dim=1)
    return near, far, (near <= far) & (far > 0)


def render_object(fc1, B1, T_wc, rays_dir, obb_R, obb_center, obb_extent, jitter, scale=2.0,
                  n_bins=150, render_feat=True):
    """render_2D_syn for one object over ALL pixels (vmap.py:604-685 with obj_mask=None).
    fc1/B1: that object's tensors with leading dim 1.  T_wc [4,4] f32; rays_dir [W,H,3];
    jitter: uniform draws [W*H, n_bins] (the reference draws only for hit rays, by rank:
    row j of its rand() goes to the j-th hit ray; pass `jitter` already in that rank order).
    Returns dict(mask [W,H] bool, depth [W,H] f32, rgb [W,H,3] u8, feat [W,H,C] or None,
    hit, near, far, opacity) -- dense maps with zeros where the mask is False."""
    W, H = rays_dir.shape[:2]
    dirs_c = rays_dir.reshape(-1, 3)
    R = T_wc[:3, :3]
    dirs_w = (R[None] @ dirs_c[..., None]).squeeze(-1)
    origin = T_wc[:3, 3]
    T_wo = torch.eye(4)
    T_wo[:3, :3] = obb_R
    T_wo[:3, 3] = obb_center
    T_oc = torch.inverse(T_wo) @ T_wc                                         # trainer.py:157-160
    dirs_o = (T_oc[None, :3, :3] @ dirs_c[..., None]).squeeze(-1)
    org_o = T_oc[:3, 3][None].expand_as(dirs_o)
    near, far, hit = ray_box(org_o, dirs_o, -obb_extent / 2.0, obb_extent / 2.0)  # :164-167
    near = torch.clip(near, 0).float()
    far = far.float() + 0.2                                                   # :169
    n_hit = int(hit.sum())
    out = dict(hit=hit.view(W, H), near=near.view(W, H), far=far.view(W, H))
    mask = torch.zeros(W * H, dtype=torch.bool)
    depth = torch.zeros(W * H)
    rgb = torch.zeros(W * H, 3, dtype=torch.uint8)
    C = fc1[16].shape[1]
    feat = torch.zeros(W * H, C) if render_feat else None
    opac_all = torch.zeros(W * H)
    if n_hit > 1:                                                             # :167 `<= 1` -> miss
        zc = stratified(near[hit], far[hit], n_bins, jitter[:n_hit])          # :174-176
        zm = 0.5 * (zc[..., 1:] + zc[..., :-1])                               # :177
        pts = origin[None, None, :] + dirs_w[hit][:, None, :] * zm[:, :, None]  # :178
        with torch.no_grad():
            a, c, f = ensemble_forward(fc1, B1, pts[None], scale)
        _, T = termination(a[0, ..., 0])                                      # vmap.py:662-663 (non-batch form)
        opac = T.sum(-1)
        d = (T * zm).sum(-1)
        col = (T[..., None] * c[0]).sum(-2)
        col8 = (col.numpy() * 255).astype(np.uint8)                           # :671 truncation
        bad = (d < near[hit]) | (d > far[hit]) | (opac < 0.9)                 # :665,672
        idx = torch.nonzero(hit).squeeze(-1)
        keep = idx[~bad]
        mask[keep] = True
        depth[keep] = d[~bad]
        rgb[keep] = torch.from_numpy(col8)[~bad]
        opac_all[idx] = opac
        if render_feat:
            rf = (T[..., None] * f[0]).sum(-2)                                # :677
            feat[keep] = rf[~bad]
    out.update(mask=mask.view(W, H), depth=depth.view(W, H), rgb=rgb.view(W, H, 3),
               feat=None if feat is None else feat.view(W, H, C), opacity=opac_all.view(W, H))
    return out


def zmerge(masks, depths, rgbs, is_bg, feats=None):
    """Sequential depth-test merge across objects in insertion order (train.py:577-594):
    an object's pixel wins if its depth < current depth buffer (strict `>` test at :582; the
    buffer starts at 100.0, :562); ids in cfg.bg_id paint colour but never write depth (:593-594)."""
    K = len(masks)
    W, H = masks[0].shape
    dbuf = torch.full((W, H), 100.0)
    cbuf = torch.zeros(W, H, 3, dtype=torch.uint8)
    win = torch.full((W, H), -1, dtype=torch.int32)
    for k in range(K):
        upd = masks[k] & (depths[k] < dbuf)
        cbuf[upd] = rgbs[k][upd]
        win[upd] = k
        if not is_bg[k]:
            dbuf[upd] = depths[k][upd]
    fbuf = None
    if feats is not None:
        fbuf = torch.zeros(W, H, feats[0].shape[-1])
        for k in range(K):
            sel = win == k
            fbuf[sel] = feats[k][sel]
    return dbuf, cbuf, win, fbuf
