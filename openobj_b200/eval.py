"""Full-frame evaluation (BASELINE config 5): every object rendered over all pixels (K5), then the
reference's sequential depth-test merge (train.py:577-594) with K6, and the 512-d part feature of the WINNING object per pixel.

Multi-GPU: each rank renders the objects it owns straight into ONE send buffer (depth f32 | rgb u8x3 | mask u8 per object),
a single `all_gather_into_tensor` moves 8 bytes per pixel and object, the merge is replicated on every rank in GLOBAL
insertion order (ensemble index k; rank = k mod G) directly on the gathered buffer through per-object pointers (no re-packing),
and features travel winner-only: K5 leaves 36 floats per hit ray (S = sum_i T_i hp_i, opacity), the 512-wide out_clip layer is
applied after the merge and only to the pixels an object won (oo_winner_features), and the owner's compact rows
(pixel, 512 floats) are all-gathered -- never a dense [W, H, 512] reduction."""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import layout, sampler
from ._lib import RenderFrameArgs, check, lib, ptr, stream

_LIN = {}
_POOL = {}


def _object_tables(objects, T_wc, dev):
    """Per-object tables of oo_render_frame: parameter-block pointers, T_OC = inv(T_WO) @ T_WC (trainer.py:153-160, float32
    host algebra like the reference), half extents."""
    Twc = torch.as_tensor(np.asarray(T_wc), dtype=torch.float32)
    thetas, tocs, halves = [], [], []
    for o in objects:
        _, bb = o.get_bound(None, final=True)
        key = id(bb)
        if getattr(o, "_tow_key", None) != key:
            T_wo = torch.eye(4)
            T_wo[:3, :3] = torch.as_tensor(np.asarray(bb.R), dtype=torch.float32)
            T_wo[:3, 3] = torch.as_tensor(np.asarray(bb.center), dtype=torch.float32)
            o._tow_key, o._tow = key, torch.inverse(T_wo)
            o._half = torch.as_tensor(np.asarray(bb.extent), dtype=torch.float32) / 2.0
        tocs.append(o._tow @ Twc)
        halves.append(o._half)
        thetas.append(o.trainer.packed_cached(dev))
    ptrs = torch.tensor([t.data_ptr() for t in thetas], dtype=torch.int64)
    pack = torch.cat([torch.stack(tocs).reshape(-1), torch.stack(halves).reshape(-1), Twc.reshape(-1)]).to(dev, non_blocking=True)
    n = len(objects)
    return thetas, ptrs.to(dev, non_blocking=True), pack[:16 * n], pack[16 * n:19 * n], pack[19 * n:]


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
    n_bins = 150
    jitter = torch.rand(W * H, n_bins, device=dev)              # trainer.py:174-176 (one draw per frame, rows by pixel)
    fr = None
    if n_local:
        # ONE call renders every local object: parallel hit lists, then one launch of the tcgen05 kernel over the pooled hits
        thetas, th_ptrs, T_oc, half, T_wc_d = _object_tables(objects, T_wc, dev)
        lin = _LIN.get(dev)
        if lin is None:
            lin = _LIN[dev] = sampler.torch_linspace01(n_bins).to(dev)
        n_blk = (npix + 255) // 256
        pool_rows = _POOL.get((dev, npix), 2 * npix)
        while True:
            obj_start = torch.zeros(n_local + 1, dtype=torch.int32, device=dev)
            hit_pix = torch.empty(pool_rows, dtype=torch.int32, device=dev)
            ray_rec = torch.empty(pool_rows, 36, dtype=torch.float32, device=dev) if render_feat else None
            scratch = torch.empty(n_local * n_blk + 1, dtype=torch.int32, device=dev)
            from . import ops
            err = ops._TC_ERR.get(dev)
            if err is None:
                err = ops._TC_ERR[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
            a = RenderFrameArgs()
            a.W, a.H, a.n_bins, a.n_obj, a.scale = W, H, n_bins, n_local, float(objects[0].trainer.obj_scale)
            a.theta, a.T_wc, a.T_oc, a.half_extent = ptr(th_ptrs), ptr(T_wc_d), ptr(T_oc), ptr(half)
            a.rays_dir, a.jitter, a.lin = ptr(rays_dir.contiguous()), ptr(jitter), ptr(lin)
            a.mask, a.depth, a.rgb = ptr(mask_v), ptr(depth_v), ptr(rgb_v)
            a.obj_start, a.hit_pix, a.ray_rec, a.pool_rows = ptr(obj_start), ptr(hit_pix), ptr(ray_rec), pool_rows
            a.scratch, a.scratch_ints, a.tc_err = ptr(scratch), scratch.numel(), ptr(err)
            timed = stats is not None and stats.get("time_kernels")
            if timed:
                ek0, ek1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ek0.record()
            with torch.cuda.device(dev):
                check(lib().oo_render_frame(ctypes.byref(a), stream()), "oo_render_frame")
            if timed:
                ek1.record()
            total = int(obj_start[n_local].item())
            if timed:
                stats["k5_ms"] = ek0.elapsed_time(ek1)
            if total <= pool_rows:
                break
            pool_rows = _POOL[(dev, npix)] = int(1.25 * total) + 1024          # rare: many overlapping boxes; render again
            mine.zero_()
        ops.check_tc(dev)
        fr = dict(thetas=thetas, th_ptrs=th_ptrs, obj_start=obj_start, hit_pix=hit_pix, ray_rec=ray_rec, pool_rows=pool_rows,
                  total=total)
        if stats is not None:
            stats["hit_rays_local"] = total
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
        if fr is not None:
            k_of = torch.arange(rank, rank + world * n_local, world, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                check(lib().oo_winner_features_frame(ptr(fr["th_ptrs"]), ptr(fr["ray_rec"]), ptr(fr["hit_pix"]), ptr(fr["obj_start"]),
                                                     n_local, ptr(k_of), ptr(winner), fr["pool_rows"], cap, ptr(rows), ptr(rpix),
                                                     ptr(n_rows), stream()), "oo_winner_features_frame")
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
