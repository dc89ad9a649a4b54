"""Mirror of objnerf/render_rays.py.  The compositing / loss arithmetic of the training path runs in the fused
kernels (`loss.step_batch_loss` -> K3, `Ensemble` -> K1); the small helpers below keep the reference's names for
callers that compose them by hand (train.py reaches them only through loss.step_batch_loss).  For CUDA tensors that
do not require grad, occupancy_activation / occupancy_to_termination / render run as kernels of libopenobj_b200.so
(oo_occupancy_activation, oo_termination, oo_render_sum); the tensor expressions with the same semantics remain for
hand-composed autograd graphs, which the fused path does not need."""
import torch


def _kernel_ok(*ts):
    """CUDA float32 tensors outside any autograd graph: the forward-only kernels apply."""
    return all(torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and not t.requires_grad for t in ts)


def occupancy_activation(alpha, distances=None):
    if _kernel_ok(alpha) and (distances is None or _kernel_ok(distances)):
        from . import ops
        return ops.occupancy_activation(alpha, distances)
    if distances is not None:
        return 1.0 - torch.exp(-alpha * distances)
    return torch.sigmoid(alpha)


def alpha_to_occupancy(depths, dirs, alpha, add_last=False):
    d = depths[..., 1:] - depths[..., :-1]
    if add_last:
        d = torch.cat([d, torch.full((depths.shape[0], 1), 0.1, device=depths.device, dtype=depths.dtype)], dim=-1)
    d = d * torch.norm(dirs, dim=-1)[:, None]
    return occupancy_activation(alpha, distances=d.to(alpha.device))


def occupancy_to_termination(occupancy, is_batch=False):
    if _kernel_ok(occupancy) and occupancy.dim() >= 1 and occupancy.numel() > 0:
        from ._lib import check, lib, ptr, stream
        occ = occupancy.contiguous()
        out = torch.empty_like(occ)
        with torch.cuda.device(occ.device):
            check(lib().oo_termination(ptr(occ), occ.numel() // occ.shape[-1], occ.shape[-1], ptr(out), stream()), "oo_termination")
        return out
    free = 1. - occupancy + 1e-10
    free = torch.cat([torch.ones_like(occupancy[..., :1]), free[..., :-1]], dim=-1)
    return occupancy * torch.cumprod(free, dim=-1)


def render(termination, vals, dim=-1):
    """render_rays.py:56-63.  The two call forms of loss.py: (T [..,S], vals [..,S], dim=-1) and
    (T [..,S,1], vals [..,S,C], dim=-2)."""
    if _kernel_ok(termination, vals) and termination.numel() > 0:
        form1 = dim in (-1, vals.dim() - 1) and termination.shape == vals.shape
        form2 = (dim in (-2, vals.dim() - 2) and termination.dim() == vals.dim() and termination.shape[-1] == 1
                 and termination.shape[:-1] == vals.shape[:-1])
        if form1 or form2:
            from ._lib import check, lib, ptr, stream
            S = vals.shape[-1] if form1 else vals.shape[-2]
            C = 1 if form1 else vals.shape[-1]
            lead = list(vals.shape[:-1]) if form1 else list(vals.shape[:-2])
            T, v = termination.contiguous(), vals.contiguous()
            out = torch.empty(lead if form1 else lead + [C], dtype=torch.float32, device=v.device)
            with torch.cuda.device(v.device):
                check(lib().oo_render_sum(ptr(T), ptr(v), v.numel() // (S * C), S, C, ptr(out), stream()), "oo_render_sum")
            return out
    return (termination * vals).sum(dim=dim)


def render_loss(render, gt, loss="L1", normalise=False):
    if loss == "L2":
        m = (render - gt) ** 2
    elif loss == "L1":
        m = torch.abs(render - gt)
    elif loss == "cos":
        m = 1 - torch.nn.functional.cosine_similarity(render, gt, dim=-1)
    else:
        raise ValueError("loss type {} not implemented!".format(loss))
    return m / gt if normalise else m


def reduce_batch_loss(loss_mat, var=None, avg=True, mask=None, loss_type="L1"):
    mask_num = torch.sum(mask, dim=-1)
    if (mask_num == 0).any():
        z = torch.zeros_like(loss_mat)
        return torch.mean(z, dim=-1) if avg else z
    if var is not None:
        info = 1.0 / (var + 1e-4) if loss_type == "L2" else 1.0 / (torch.sqrt(var) + 1e-4)
        loss_mat = loss_mat * info
    if not avg:
        return loss_mat
    if mask is None:
        return torch.mean(loss_mat, dim=-1).sum()
    out = torch.sum(loss_mat, dim=-1) / (mask_num + 1e-10)
    if (out > 100000).any():
        raise FloatingPointError("loss explode")     # the reference prints and exit(-1)s (render_rays.py:109-111)
    return out


def make_3D_grid(occ_range=(-1., 1.), dim=256, device="cuda:0", transform=None, scale=None):
    t = torch.linspace(occ_range[0], occ_range[1], steps=dim, device=device)
    g = torch.stack(torch.meshgrid(t, t, t, indexing="ij"), dim=3)
    if scale is not None:
        g = g * scale
    if transform is not None:
        g = (g[..., None, :] * transform[None, None, None, :3, :3]).sum(-1) + transform[None, None, None, :3, 3]
    return g
