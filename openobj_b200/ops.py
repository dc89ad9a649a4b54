"""Thin functional wrappers over the C ABI (tensors in, tensors out).  Everything here launches CUDA kernels
from libopenobj_b200.so; nothing falls back to PyTorch arithmetic."""
import ctypes

import torch

from . import _lib, layout
from ._lib import check, lib, ptr, stream


def _dev(t):
    if not t.is_cuda:
        raise _lib.OOError("openobj_b200 kernels need CUDA tensors; got a %s tensor (no CPU fallback)" % t.device)
    return torch.cuda.device(t.device)


def forward(theta, pcs=None, emb=None, scale=2.0, want_clip=True, want_emb=False):
    """vmap(pe) -> vmap(fc) forward.  theta [N,PSTRIDE]; pcs [N,...,3] or emb [N,...,129].
    Returns alpha [N,...,1] (x10 applied), color [N,...,3], clip [N,...,512] or None, emb or None."""
    src = pcs if pcs is not None else emb
    _dev(theta)
    n = theta.shape[0]
    lead = list(src.shape[:-1])
    m = 1
    for s in lead[1:]:
        m *= s
    src = src.contiguous().float()
    f32 = dict(dtype=torch.float32, device=theta.device)
    alpha = torch.empty(n, m, **f32)
    color = torch.empty(n, m, 3, **f32)
    clip = torch.empty(n, m, layout.CLIP, **f32) if want_clip else None
    emb_o = torch.empty(n, m, 129, **f32) if want_emb else None
    with _dev(theta):
        check(lib().oo_forward(ptr(theta), n, ptr(src) if pcs is not None else None,
                               ptr(src) if pcs is None else None, m, float(scale), ptr(alpha), ptr(color), ptr(clip),
                               ptr(emb_o), stream()), "oo_forward")
    return (alpha.view(lead + [1]), color.view(lead + [3]),
            None if clip is None else clip.view(lead + [layout.CLIP]),
            None if emb_o is None else emb_o.view(lead + [129]))


def embed(theta, pcs, scale=2.0):
    """vmap(pe_model) on its own: theta [N,PSTRIDE] (only the B_layer block is read), pcs [N,...,3] -> emb [N,...,129]."""
    _dev(theta)
    n = theta.shape[0]
    lead = list(pcs.shape[:-1])
    x = pcs.reshape(n, -1, 3).contiguous().float()
    m = x.shape[1]
    emb = torch.empty(n, m, 129, dtype=torch.float32, device=theta.device)
    with _dev(theta):
        check(lib().oo_forward(ptr(theta), n, ptr(x), None, m, float(scale), None, None, None, ptr(emb), stream()), "oo_forward")
    return emb.view(lead + [129])


# ---- the autograd surface of the reference's call form (train.py:424-425, backward at :472) ----------------------------
def _as_theta(params, first, fallback_theta=None):
    """A [N,PSTRIDE] block holding `params` (stacked tensors first .. first+len-1 of the layout).  When they already are
    the strided views of one such block (what update_vmap hands out) that block is used in place; otherwise they are
    packed into a fresh one."""
    n = params[0].shape[0]
    base = params[0].data_ptr() - 4 * layout.OFFSETS[first]
    aliased = all(p.is_cuda and p.dtype == torch.float32 and p.data_ptr() == base + 4 * layout.OFFSETS[first + i]
                  and (n == 1 or p.stride(0) == layout.PSTRIDE) and p[0].is_contiguous() for i, p in enumerate(params))
    if aliased and fallback_theta is not None and fallback_theta.data_ptr() == base and fallback_theta.shape[0] == n:
        return fallback_theta
    theta = torch.zeros(n, layout.PSTRIDE, dtype=torch.float32, device=params[0].device)
    with torch.no_grad():
        for v, p in zip(layout.views(theta)[first:first + len(params)], params):
            v.copy_(p)
    return theta


class _EmbedFn(torch.autograd.Function):
    """emb = vmap(pe_model)(pe_param, pe_buffer, pcs) with the gradient of the trainable direction matrix."""

    @staticmethod
    def forward(ctx, pcs, theta, scale, B):
        ctx.set_materialize_grads(False)
        emb = embed(theta, pcs, scale)
        ctx.save_for_backward(pcs, theta)
        ctx.scale = float(scale)
        return emb

    @staticmethod
    def backward(ctx, d_emb):
        if d_emb is None:
            return None, None, None, None
        pcs, theta = ctx.saved_tensors
        n = theta.shape[0]
        x = pcs.reshape(n, -1, 3).contiguous().float()
        m = x.shape[1]
        de = d_emb.reshape(n, m, 129).contiguous().float()
        L = lib()
        d_B = torch.empty(n, 21, 3, dtype=torch.float32, device=theta.device)
        ws = torch.empty(L.oo_embed_bwd_ws_floats(n, m), dtype=torch.float32, device=theta.device)
        with _dev(theta):
            check(L.oo_embed_bwd(ptr(theta), n, ptr(x), m, ctx.scale, ptr(de), ptr(d_B), ptr(ws), stream()), "oo_embed_bwd")
        return None, None, None, d_B


class _FcFn(torch.autograd.Function):
    """alpha, color, clip = vmap(fc_model)(fc_param, fc_buffer, emb) with gradients of the 18 stacked tensors and of emb."""

    @staticmethod
    def forward(ctx, emb, theta, want_clip, *params):
        ctx.set_materialize_grads(False)          # an output no loss term uses must leave its tensors with grad None
        a, c, f, _ = forward(theta, emb=emb, want_clip=want_clip)
        ctx.save_for_backward(emb, theta)
        ctx.want_clip = want_clip
        if f is None:
            return a, c
        return a, c, f

    @staticmethod
    def backward(ctx, d_alpha, d_color, d_clip=None):
        emb, theta = ctx.saved_tensors
        n = theta.shape[0]
        e = emb.reshape(n, -1, 129).contiguous().float()
        m = e.shape[1]
        dev = theta.device
        da = (torch.zeros(n, m, device=dev) if d_alpha is None else d_alpha.reshape(n, m).contiguous().float())
        dc = (torch.zeros(n, m, 3, device=dev) if d_color is None else d_color.reshape(n, m, 3).contiguous().float())
        dcl = None if (d_clip is None or not ctx.want_clip) else d_clip.reshape(n, m, layout.CLIP).contiguous().float()
        L = lib()
        grads = torch.empty(n, layout.PSTRIDE, dtype=torch.float32, device=dev)
        d_emb = torch.empty(n, m, 129, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        ws = torch.empty(L.oo_forward_bwd_ws_floats(n, m), dtype=torch.float32, device=dev)
        with _dev(theta):
            check(L.oo_forward_bwd(ptr(theta), n, ptr(e), m, ptr(da), ptr(dc), ptr(dcl), ptr(grads), ptr(d_emb), ptr(ws),
                                   stream()), "oo_forward_bwd")
        gv = layout.views(grads)[:18]
        # autograd leaves tensors no loss term reaches with grad None (the optimiser then skips them: quirk 8)
        if d_color is None:
            gv[10:14] = [None] * 4
        if dcl is None:
            gv[14:18] = [None] * 4
        if d_alpha is None and d_color is None and dcl is None:
            gv = [None] * 18
        return (None if d_emb is None else d_emb.view(emb.shape), None, None) + tuple(gv)


def embed_autograd(pcs, theta, scale, B):
    return _EmbedFn.apply(pcs, theta, scale, B)


def fc_autograd(emb, theta, want_clip, params):
    return _FcFn.apply(emb, theta, want_clip, *params)


class _WideFn(torch.autograd.Function):
    """OccupancyMap.forward of a model of another hidden width (the hidden-128 background model called as
    `fc_occ_map(pe(x))`, train.py:449-450) through the layer-by-layer kernels of oo_bg."""

    @staticmethod
    def forward(ctx, emb, model, want_clip, *params):
        ctx.set_materialize_grads(False)
        model.load(list(params) + [None])
        a, c, f, _ = model.forward(emb=emb, want_clip=want_clip)
        ctx.model, ctx.want_clip = model, want_clip
        ctx.save_for_backward(emb, *params)
        if f is None:
            return a, c
        return a, c, f

    @staticmethod
    def backward(ctx, d_alpha, d_color, d_clip=None):
        emb = ctx.saved_tensors[0]
        params = ctx.saved_tensors[1:]
        model = ctx.model
        model.load(list(params) + [None])              # the scratch model may have served another call in between
        n = emb.reshape(-1, 129).shape[0]
        dev = emb.device
        da = torch.zeros(n, device=dev) if d_alpha is None else d_alpha
        dc = torch.zeros(n, 3, device=dev) if d_color is None else d_color
        dcl = d_clip if ctx.want_clip else None
        g, d_emb = model.forward_bwd(None, da, dc, dcl, emb=emb, want_d_emb=ctx.needs_input_grad[0])
        gv = [v.clone() for v in model.views(g)[:18]]
        if d_color is None:
            gv[10:14] = [None] * 4
        if dcl is None:
            gv[14:18] = [None] * 4
        if d_alpha is None and d_color is None and dcl is None:
            gv = [None] * 18
        return (None if d_emb is None else d_emb.view(emb.shape), None, None) + tuple(gv)


def wide_autograd(emb, model, want_clip, params):
    return _WideFn.apply(emb, model, want_clip, *params)


_TC_ERR = {}


def eval_points_tc(theta, points, scale=2.0, want_alpha=False):
    """The tcgen05 / TMEM forward (oo_eval_points_tc): occ [n] (or alpha when want_alpha), color [n,3].  The kernel's error
    flag (a tensor-core completion that never arrived) is checked by check_tc()."""
    _dev(theta)
    pts = points.reshape(-1, 3).contiguous().float()
    n = pts.shape[0]
    f32 = dict(dtype=torch.float32, device=theta.device)
    out, color = torch.empty(n, **f32), torch.empty(n, 3, **f32)
    err = _TC_ERR.get(theta.device)
    if err is None:
        err = _TC_ERR[theta.device] = torch.zeros(1, dtype=torch.int32, device=theta.device)
    with _dev(theta):
        check(lib().oo_eval_points_tc(ptr(theta), ptr(pts), n, float(scale), None if want_alpha else ptr(out),
                                      ptr(out) if want_alpha else None, ptr(color), ptr(err), stream()), "oo_eval_points_tc")
    return out, color


def check_tc(device):
    """Raises if a tcgen05 kernel reported a missed tensor-core completion (synchronises)."""
    err = _TC_ERR.get(torch.device(device))
    if err is not None and int(err.item()) != 0:
        raise _lib.OOError("oo_eval_points_tc: a tcgen05.mma completion did not arrive (results are invalid)")


def eval_points(theta, points, scale=2.0, want_clip=True, tensor_core=True):
    """Trainer.eval_points for ONE hidden-32 model (trainer.py:104-128): theta [1,PSTRIDE] (or [PSTRIDE]), points [n,3].
    Returns occ [n] = sigmoid(alpha), color [n,3], clip [n,512] or None -- one launch for the whole query.  Without the part
    feature (what Trainer.meshing uses) the query runs on the tcgen05 / TMEM kernel."""
    _dev(theta)
    if not want_clip and tensor_core:
        occ, color = eval_points_tc(theta, points, scale)
        return occ, color, None
    pts = points.reshape(-1, 3).contiguous().float()
    n = pts.shape[0]
    f32 = dict(dtype=torch.float32, device=theta.device)
    occ, color = torch.empty(n, **f32), torch.empty(n, 3, **f32)
    clip = torch.empty(n, layout.CLIP, **f32) if want_clip else None
    with _dev(theta):
        check(lib().oo_eval_points(ptr(theta), ptr(pts), n, float(scale), ptr(occ), ptr(color), ptr(clip), stream()),
              "oo_eval_points")
    return occ, color, clip


def make_grid(dim, scale, transform, center=None, occ_range=(-1., 1.), device="cuda:0"):
    """The query grid of Trainer.meshing (trainer.py:46-66) = render_rays.make_3D_grid with scale [3] and transform [4,4],
    minus `center`; returns [dim^3, 3] on the device ((i, j, k) -> (i*dim + j)*dim + k)."""
    dev = torch.device(device)
    t = torch.linspace(occ_range[0], occ_range[1], steps=dim, device=dev)       # the dim knot values (plumbing)
    g = _lib.Grid()
    g.dim, g.t = int(dim), ptr(t)
    sc = [1.0, 1.0, 1.0] if scale is None else [float(v) for v in torch.as_tensor(scale).reshape(-1).tolist()]
    if len(sc) == 1:
        sc = sc * 3
    tr = torch.eye(4) if transform is None else torch.as_tensor(transform).detach().cpu().float()
    ce = [0.0, 0.0, 0.0] if center is None else [float(v) for v in torch.as_tensor(center).reshape(-1).tolist()]
    g.scale[:] = sc
    g.transform[:] = [float(v) for v in tr[:3, :4].reshape(-1).tolist()]
    g.center[:] = ce
    out = torch.empty(dim ** 3, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().oo_make_grid(ctypes.byref(g), ptr(out), stream()), "oo_make_grid")
    return out


def occupancy_activation(alpha, distances=None):
    """render_rays.occupancy_activation (render_rays.py:6-14) on the device."""
    _dev(alpha)
    a = alpha.contiguous().float()
    d = None if distances is None else distances.expand_as(a).contiguous().float()
    occ = torch.empty_like(a)
    with _dev(a):
        check(lib().oo_occupancy_activation(ptr(a), ptr(d), a.numel(), ptr(occ), stream()), "oo_occupancy_activation")
    return occ


class _StepLoss(torch.autograd.Function):
    """loss.step_batch_loss as one fused forward / backward kernel pair (K3)."""

    @staticmethod
    def forward(ctx, alpha, color, gt_depth, gt_color, labels, z, pred_feat, gt_feat, cs, os_, fs):
        n, r, s = z.shape
        dev = z.device
        alpha_c = alpha.reshape(n, r, s).contiguous().float()
        color_c = color.reshape(n, r, s, 3).contiguous().float()
        gt_depth, gt_color, z = gt_depth.contiguous().float(), gt_color.contiguous().float(), z.contiguous().float()
        labels = labels.contiguous().to(torch.uint8)
        cf = 0
        if pred_feat is not None:
            pred_feat, gt_feat = pred_feat.contiguous().float(), gt_feat.contiguous().float()
            cf = pred_feat.shape[-1]
        L = lib()
        ws = torch.empty(n * r * L.oo_loss_ws_per_ray() + 8 * n, dtype=torch.float32, device=dev)
        terms = torch.empty(n, 4, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        with _dev(z):
            check(L.oo_loss_fwd(ptr(alpha_c), ptr(color_c), ptr(z), ptr(gt_depth), ptr(gt_color), ptr(labels),
                                ptr(pred_feat), ptr(gt_feat), n, r, s, cf, cs, os_, fs, ptr(terms), ptr(loss), ptr(flags),
                                ptr(ws), stream()), "oo_loss_fwd")
        ctx.save_for_backward(alpha_c, color_c, z, gt_depth, gt_color, labels, pred_feat, gt_feat, flags, ws)
        ctx.dims = (n, r, s, cf, cs, os_, fs, alpha.shape, color.shape)
        ctx.mark_non_differentiable(terms, flags)
        return loss.reshape(()), terms, flags

    @staticmethod
    def backward(ctx, g_loss, _g_terms, _g_flags):
        alpha_c, color_c, z, gt_depth, gt_color, labels, pred_feat, gt_feat, flags, ws = ctx.saved_tensors
        n, r, s, cf, cs, os_, fs, a_shape, c_shape = ctx.dims
        # reduce_batch_loss returns constant zeros for a term with an empty mask (render_rays.py:89-94): the inputs only such
        # terms reach get no gradient at all (None), which is what lets the optimiser skip their tensors.  The reference pays
        # the same host sync here (`.any()` in a Python `if`).
        f = int(flags.item())
        no_obj, no_sem = bool(f & 2), bool(f & 4)
        if no_obj and no_sem:
            return (None,) * 11
        d_alpha = torch.empty_like(alpha_c)
        d_color = torch.empty_like(color_c)
        d_pred = torch.empty_like(pred_feat) if pred_feat is not None else None
        with _dev(z):
            check(lib().oo_loss_bwd(ptr(alpha_c), ptr(color_c), ptr(z), ptr(gt_depth), ptr(gt_color), ptr(labels),
                                    ptr(pred_feat), ptr(gt_feat), n, r, s, cf, cs, os_, fs, float(g_loss), ptr(flags),
                                    ptr(ws), ptr(d_alpha), ptr(d_color), ptr(d_pred), stream()), "oo_loss_bwd")
        return (d_alpha.view(a_shape), None if no_obj else d_color.view(c_shape), None, None, None, None,
                None if no_obj else d_pred, None, None, None, None)


def step_loss(alpha, color, gt_depth, gt_color, labels, z, pred_feat=None, gt_feat=None,
              color_scaling=5.0, opacity_scaling=10.0, feat_scaling=5.0):
    """Returns (loss scalar with autograd, per-object terms [N,4], flags int32[1])."""
    return _StepLoss.apply(alpha, color, gt_depth, gt_color, labels, z, pred_feat, gt_feat,
                           float(color_scaling), float(opacity_scaling), float(feat_scaling))


def adamw_flat(p, g, m, v, step, lr=1e-3, weight_decay=0.013, betas=(0.9, 0.999), eps=1e-8):
    with _dev(p):
        check(lib().oo_adamw_flat(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), int(step), lr, weight_decay, betas[0], betas[1],
                                  eps, stream()), "oo_adamw_flat")


def rng_fill(out, seed, frame, obj_ids, kind="uniform", std=1.0):
    """out [n_obj, per_obj] f32 <- Philox4x32-10 keyed by (seed, frame, object id, element)."""
    n, per = out.shape
    with _dev(out):
        check(lib().oo_rng_fill(int(seed), int(frame), ptr(obj_ids), n, per, 0 if kind == "uniform" else 1, float(std),
                                ptr(out), stream()), "oo_rng_fill")
    return out


def rng_fill_rows(n_obj_rows_words, seed, frame, obj_ids, kind="uniform", std=1.0):
    """Ray-blocked counter tapes [n_obj, n_rows, row_words] (oo_rng_fill_rows): what K2 draws in-kernel for ray = row."""
    n, rows, words = n_obj_rows_words
    out = torch.empty(n, rows, words, dtype=torch.float32, device=obj_ids.device)
    with _dev(out):
        check(lib().oo_rng_fill_rows(int(seed), int(frame), ptr(obj_ids), n, rows, words, 0 if kind == "uniform" else 1,
                                     float(std), ptr(out), stream()), "oo_rng_fill_rows")
    return out


def zmerge(masks, depths, rgbs, is_bg):
    """Sequential strict depth test across objects in order (train.py:577-594).
    masks [K,W,H] u8/bool, depths [K,W,H] f32, rgbs [K,W,H,3] u8, is_bg [K]."""
    k = masks.shape[0]
    shp = masks.shape[1:]
    npix = masks[0].numel()
    dev = masks.device
    masks = masks.contiguous().to(torch.uint8)
    depth_out = torch.empty(shp, dtype=torch.float32, device=dev)
    rgb_out = torch.empty(tuple(shp) + (3,), dtype=torch.uint8, device=dev)
    win = torch.empty(shp, dtype=torch.int32, device=dev)
    bg = torch.as_tensor(is_bg, dtype=torch.uint8).to(dev).contiguous()
    with _dev(masks):
        check(lib().oo_zmerge(ptr(masks), ptr(depths.contiguous()), ptr(rgbs.contiguous()), ptr(bg), k, npix,
                              ptr(depth_out), ptr(rgb_out), ptr(win), stream()), "oo_zmerge")
    return depth_out, rgb_out, win


def fma_peak_tflops(iters=4000):
    """Measured FP32 FFMA throughput of this GPU (roofline denominator of the fused step)."""
    L = lib()
    sink = torch.zeros(4, device="cuda")
    n_sm = _lib.n_sm()
    check(L.oo_fma_peak(n_sm, 200, ptr(sink), stream()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 0.0
    for _ in range(3):
        e0.record()
        check(L.oo_fma_peak(n_sm, iters, ptr(sink), stream()))
        e1.record()
        torch.cuda.synchronize()
        flops = n_sm * 4 * 256 * iters * 16 * 8 * 2
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best
