"""Full-frame evaluation (BASELINE config 5): every object rendered over all pixels (K5), then the
reference's sequential depth-test merge (train.py:577-594) with K6, and the 512-d part feature of the WINNING object per pixel.

Multi-GPU: each rank renders the objects it owns straight into ONE send buffer (depth f32 | rgb u8x3 | mask u8 per object),
a single `all_gather_into_tensor` moves 8 bytes per pixel and object, the merge is replicated on every rank in GLOBAL
insertion order (ensemble index k; rank = k mod G) directly on the gathered buffer through per-object pointers (no re-packing),
and features travel winner-only: K5 leaves 36 floats per hit ray (S = sum_i T_i hp_i, opacity), the 512-wide out_clip layer is
applied after the merge and only to the pixels an object won (oo_winner_features), and the owner's compact rows
(pixel, 512 floats) are all-gathered -- never a dense [W, H, 512] reduction."""
import torch
import torch.distributed as dist

from . import layout
from ._lib import check, lib, ptr, stream


def global_order(n_local_per_rank, world):
    """Ensemble indices in gather order -> permutation that sorts them by k.  Rank r holds k = r, r+G, r+2G, ...;
    all_gather concatenates rank-major, so gathered position (r, i) is object k = i*G + r."""
    ks = []
    for r, n in enumerate(n_local_per_rank):
        ks += [i * world + r for i in range(n)]
    order = sorted(range(len(ks)), key=lambda q: ks[q])
    return ks, order


def _send_layout(n_max, npix):
    """Byte offsets of the depth / rgb / mask regions of one rank's chunk (depth first: 4-byte aligned)."""
    d_off, c_off = 0, n_max * npix * 4
    m_off = c_off + n_max * npix * 3
    total = (m_off + n_max * npix + 15) // 16 * 16
    return d_off, c_off, m_off, total


def render_frame(objects, T_wc, rays_dir, is_bg=None, render_feat=False, group=None, stats=None):
    """objects: this rank's sceneObjects in local insertion order (each with .bbox3dour set).  Returns
    (depth [W,H] f32, rgb [W,H,3] u8, winner [W,H] int32 = global ensemble index or -1, feat [W,H,512] or None).
    stats (dict, optional) receives the bytes this rank received over the interconnect."""
    dev = rays_dir.device
    W, H = rays_dir.shape[:2]
    npix = W * H
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_local = len(objects)
    counts = [n_local]
    if world > 1:
        t = torch.tensor([n_local], device=dev)
        allc = torch.empty(world, dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(allc, t, group=group)
        counts = [int(x) for x in allc.tolist()]
    n_max = max(max(counts), 1)
    d_off, c_off, m_off, chunk = _send_layout(n_max, npix)
    # ---- K5: every local object renders straight into this rank's chunk of the gather buffer
    gathered = torch.zeros(world, chunk, dtype=torch.uint8, device=dev)
    mine = gathered[rank]
    depth_v = mine[d_off:d_off + n_max * npix * 4].view(torch.float32).view(n_max, W, H)
    rgb_v = mine[c_off:c_off + n_max * npix * 3].view(n_max, W, H, 3)
    mask_v = mine[m_off:m_off + n_max * npix].view(n_max, W, H)
    jitter = torch.rand(W * H, 150, device=dev)                 # trainer.py:174-176 (one draw per frame, rows by pixel)
    recs = []
    for i, o in enumerate(objects):
        recs.append(o._render(T_wc, rays_dir, jitter=jitter, jitter_by_pixel=True, out=(mask_v[i], depth_v[i], rgb_v[i]),
                              want_rec=render_feat))
    if world > 1:
        dist.all_gather_into_tensor(gathered.view(-1), mine.clone(), group=group)
    # ---- K6 on the gathered buffer, objects addressed through pointers in global insertion order
    ks, order = global_order(counts, world)
    k_sorted = [ks[q] for q in order]
    base = gathered.data_ptr()
    slots = []
    for r, n in enumerate(counts):
        slots += [(r, i) for i in range(n)]
    pm, pd, pc = [], [], []
    for q in order:
        r, i = slots[q]
        pd.append(base + r * chunk + d_off + i * npix * 4)
        pc.append(base + r * chunk + c_off + i * npix * 3)
        pm.append(base + r * chunk + m_off + i * npix)
    K = len(order)
    tabs = torch.tensor([pm, pd, pc], dtype=torch.int64).to(dev) if K else torch.zeros(3, 1, dtype=torch.int64, device=dev)
    bg = [False] * K if is_bg is None else [bool(is_bg.get(k, False)) if isinstance(is_bg, dict) else bool(is_bg[k]) for k in k_sorted]
    bg_t = torch.tensor(bg if K else [0], dtype=torch.uint8).to(dev)
    depth = torch.empty(W, H, dtype=torch.float32, device=dev)
    rgb = torch.empty(W, H, 3, dtype=torch.uint8, device=dev)
    win_pos = torch.empty(W, H, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().oo_zmerge_ptr(ptr(tabs[0]), ptr(tabs[1]), ptr(tabs[2]), ptr(bg_t), K, npix, ptr(depth), ptr(rgb), ptr(win_pos),
                                  stream()), "oo_zmerge_ptr")
    lut = torch.tensor(k_sorted + [-1], device=dev, dtype=torch.int32)
    winner = lut[win_pos.long()]                       # win_pos == -1 indexes the sentinel
    if stats is not None:
        stats["allgather_bytes_received"] = (world - 1) * chunk
    feat = None
    if render_feat:
        # ---- winner-only features: out_clip applied to the pixels each LOCAL object won, rows compacted, then exchanged
        cap = npix
        rows = torch.empty(cap, layout.CLIP, dtype=torch.float32, device=dev)
        rpix = torch.empty(cap, dtype=torch.int32, device=dev)
        n_rows = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            for i, r in enumerate(recs):
                check(lib().oo_winner_features(ptr(r["theta"]), ptr(r["rec"]), ptr(r["hit_pix"]), ptr(r["n_hit"]), ptr(winner),
                                               i * world + rank, cap, ptr(rows), ptr(rpix), ptr(n_rows), stream()),
                      "oo_winner_features")
        feat = torch.zeros(npix, layout.CLIP, dtype=torch.float32, device=dev)
        if world == 1:
            n = int(n_rows.item())
            feat.index_copy_(0, rpix[:n].long(), rows[:n])
        else:
            alln = torch.empty(world, dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(alln, n_rows, group=group)
            ns = [int(x) for x in alln.tolist()]
            n_pad = max(max(ns), 1)
            g_rows = torch.empty(world, n_pad, layout.CLIP, dtype=torch.float32, device=dev)
            g_pix = torch.empty(world, n_pad, dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(g_rows.view(-1), rows[:n_pad].contiguous().view(-1), group=group)
            dist.all_gather_into_tensor(g_pix.view(-1), rpix[:n_pad].contiguous(), group=group)
            for r, n in enumerate(ns):
                if n:
                    feat.index_copy_(0, g_pix[r, :n].long(), g_rows[r, :n])
            if stats is not None:
                stats["feature_bytes_received"] = (world - 1) * n_pad * (layout.CLIP * 4 + 4)
                stats["feature_rows_won"] = ns
        feat = feat.view(W, H, layout.CLIP)
    return depth, rgb, winner, feat
