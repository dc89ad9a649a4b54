"""Full-frame evaluation: every object rendered over all pixels (K5), then the reference's sequential
depth-test merge (train.py:577-594) with K6.  Multi-GPU: each rank renders the objects it owns, the dense depth / rgb /
mask tiles are all-gathered over NCCL, the merge is replicated on every rank in GLOBAL insertion order (ensemble index
k; rank = k mod G), and only the winning object's 512-d feature per pixel is exchanged (sum of disjoint maps)."""
import torch
import torch.distributed as dist

from . import ops


def global_order(n_local_per_rank, world):
    """Ensemble indices in gather order -> permutation that sorts them by k.  Rank r holds k = r, r+G, r+2G, ...;
    all_gather concatenates rank-major, so gathered position (r, i) is object k = i*G + r."""
    ks = []
    for r, n in enumerate(n_local_per_rank):
        ks += [i * world + r for i in range(n)]
    order = sorted(range(len(ks)), key=lambda q: ks[q])
    return ks, order


def render_frame(objects, T_wc, rays_dir, is_bg=None, render_feat=False, group=None):
    """objects: this rank's sceneObjects in local insertion order (each with .bbox3dour set).  Returns
    (depth [W,H] f32, rgb [W,H,3] u8, winner [W,H] int32 = global ensemble index or -1, feat [W,H,512] or None)."""
    dev = rays_dir.device
    W, H = rays_dir.shape[:2]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    masks, depths, rgbs, feats = [], [], [], []
    for o in objects:
        m, d, c, f = o.render_2D_syn(T_wc, None, rays_dir, render_part=render_feat, dense=True)
        masks.append(m.to(torch.uint8)); depths.append(d); rgbs.append(c); feats.append(f)
    n_local = len(objects)
    counts = [n_local]
    if world > 1:
        t = torch.tensor([n_local], device=dev)
        allc = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allc, t, group=group)
        counts = [int(x.item()) for x in allc]
    n_max = max(counts) if counts else 0

    def stack(lst, shape, dtype):
        out = torch.zeros((n_max,) + shape, dtype=dtype, device=dev)
        if lst:
            out[:len(lst)] = torch.stack(lst)
        return out

    M, D, C = stack(masks, (W, H), torch.uint8), stack(depths, (W, H), torch.float32), stack(rgbs, (W, H, 3), torch.uint8)
    if world > 1:
        gm = [torch.empty_like(M) for _ in range(world)]
        gd = [torch.empty_like(D) for _ in range(world)]
        gc = [torch.empty_like(C) for _ in range(world)]
        dist.all_gather(gm, M, group=group); dist.all_gather(gd, D, group=group); dist.all_gather(gc, C, group=group)
        M = torch.cat([g[:n] for g, n in zip(gm, counts)]); D = torch.cat([g[:n] for g, n in zip(gd, counts)])
        C = torch.cat([g[:n] for g, n in zip(gc, counts)])
    else:
        M, D, C = M[:n_local], D[:n_local], C[:n_local]
    ks, order = global_order(counts, world)
    idx = torch.tensor(order, device=dev, dtype=torch.long)
    M, D, C = M[idx].contiguous(), D[idx].contiguous(), C[idx].contiguous()
    k_sorted = [ks[q] for q in order]
    bg = [False] * len(k_sorted) if is_bg is None else [bool(is_bg.get(k, False)) if isinstance(is_bg, dict) else bool(is_bg[k]) for k in k_sorted]
    depth, rgb, win_pos = ops.zmerge(M, D, C, bg)
    lut = torch.tensor(k_sorted + [-1], device=dev, dtype=torch.int32)
    winner = lut[win_pos.long()]                       # win_pos == -1 indexes the sentinel
    feat = None
    if render_feat:
        feat = torch.zeros(W, H, 512, device=dev)
        for i, f in enumerate(feats):
            k = i * world + rank
            sel = winner == k
            feat[sel] = f[sel]
        if world > 1:
            dist.all_reduce(feat, group=group)         # disjoint supports: the sum is the winner's feature
    return depth, rgb, winner, feat
