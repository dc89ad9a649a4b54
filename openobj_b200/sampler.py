"""Host side of K2: batches sceneObject.get_training_samples + sample_3d_points of MANY objects into one launch
(reference loop: objnerf/train.py:317-332 calling vmap.py:386-554 per object)."""
import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from ._lib import SampleArgs, check, lib, ptr, stream

_LIN = {}
_SCRATCH = {}


def _lin_tables(S, n_c2s, n_bins):
    key = (S, n_c2s, n_bins)
    if key not in _LIN:
        _LIN[key] = [torch_linspace01(S), torch_linspace01(n_c2s), torch_linspace01(n_bins)]
    return _LIN[key]


def _scratch(dev, n_ints):
    """Caller-owned scratch of the counter-RNG path (per-object batch maximum + invalid-depth ray lists), one per device and
    stream; the library allocates nothing."""
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    t = _SCRATCH.get(key)
    if t is None or t.numel() < n_ints:
        t = torch.empty(max(int(n_ints), 1 << 20), dtype=torch.int32, device=dev)
        _SCRATCH[key] = t
    return t


def torch_linspace01(n):
    """torch.linspace(0, 1, n+1) on the CPU -- what utils.stratified_bins uses (utils.py:349)."""
    return torch.linspace(0, 1, n + 1, dtype=torch.float32).contiguous()


@dataclass
class SampleTapes:
    """RNG tape for n_obj objects.  by_rank=True: class-tape row j belongs to the j-th ray of that
    class (the reference's consumption order); False: row = ray index (shard-independent counter RNG)."""
    kf_ids: torch.Tensor       # int64 [n_obj, n_frames]
    u_w: torch.Tensor          # f32 [n_obj, n_rays]
    u_h: torch.Tensor
    r_invalid: torch.Tensor    # f32 [n_obj, n_rays, S]
    r_valid: torch.Tensor      # f32 [n_obj, n_rays, n_c2s]
    r_normal: torch.Tensor     # f32 [n_obj, n_rays, n_bins]
    r_other: torch.Tensor      # f32 [n_obj, n_rays, n_bins]
    by_rank: bool = False


@dataclass
class SampleOut:
    gt_rgb: torch.Tensor       # u8 [n_obj, n_rays, 3]
    gt_depth: torch.Tensor     # f32 [n_obj, n_rays]
    valid: torch.Tensor        # u8 [n_obj, n_rays]
    labels: torch.Tensor       # u8 [n_obj, n_rays]
    pcs: torch.Tensor          # f32 [n_obj, n_rays, S, 3]
    z: torch.Tensor            # f32 [n_obj, n_rays, S]
    feat_row: Optional[torch.Tensor]   # int32 [n_obj, n_rays] rows of global_partfeat.view(-1, C)
    pix: Optional[torch.Tensor]        # int64 [n_obj, n_rays, 3] (kf, w, h)
    oob: torch.Tensor          # int32 [1]


def _ptr_table(tensors, device):
    return torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device=device)


def latest_kf_ids(draws, n_keyframes, latest):
    """vmap.py:390-410: the last two of the n_frames keyframe ids are forced to the two latest keyframes."""
    if n_keyframes > 2:
        return torch.cat([draws, torch.as_tensor(latest[-2:], dtype=torch.int64, device=draws.device)])
    return draws


def device_tapes(objects, n_frames, n_samples, n_c2s, n_bins, eps, seed, frame, device):
    """Counter-based tapes (Philox keyed by seed/frame/object id) for a list of sceneObject."""
    n = len(objects)
    n_rays = n_frames * n_samples
    S = n_c2s + n_bins
    ids = torch.tensor([o.obj_id for o in objects], dtype=torch.int32, device=device)
    f32 = dict(dtype=torch.float32, device=device)
    u_kf = ops.rng_fill(torch.empty(n, n_frames, **f32), seed, 8 * frame + 0, ids)
    nkf = torch.tensor([o.n_keyframes for o in objects], dtype=torch.float32, device=device)[:, None]
    kf = torch.minimum((u_kf * nkf).int().long(), (nkf - 1).long())
    for i, o in enumerate(objects):              # forced latest two keyframes (host bookkeeping, vmap.py:398-400)
        if o.n_keyframes > 2:
            kf[i, -2:] = torch.as_tensor(o.lastest_kf_queue[-2:], device=device)
    # every other draw of ray r: the ray-blocked stream 8*frame + 1 (include/openobj_b200.h, oo_sample_args.rng_mode)
    b0 = 4 * ((2 + n_c2s + 3) // 4)
    words = b0 + 4 * ((max(S, 2 * ((n_bins + 1) // 2)) + 3) // 4)
    U = ops.rng_fill_rows((n, n_rays, words), seed, 8 * frame + 1, ids)
    Nn = ops.rng_fill_rows((n, n_rays, words), seed, 8 * frame + 1, ids, "normal", eps / 3.)
    return SampleTapes(
        kf_ids=kf.contiguous(),
        u_w=U[..., 0].contiguous(), u_h=U[..., 1].contiguous(),
        r_invalid=U[..., b0:b0 + S].contiguous(),
        r_valid=U[..., 2:2 + n_c2s].contiguous(),
        r_normal=Nn[..., b0:b0 + n_bins].contiguous(),
        r_other=U[..., b0:b0 + n_bins].contiguous(),
        by_rank=False)


@dataclass
class CounterRng:
    """In-kernel counter RNG (no tapes): same values as device_tapes() for the same (seed, frame, object ids)."""
    seed: int
    frame: int
    obj_ids: torch.Tensor      # int32 [n_obj]
    n_keyframes: torch.Tensor  # int32 [n_obj]
    latest: torch.Tensor       # int32 [n_obj, 2]


class RingTables:
    """Device pointer tables of the objects' keyframe rings (rebuilt only when the object set changes)."""

    def __init__(self, objects, device):
        self.n = len(objects)
        self.rgbs = _ptr_table([o.rgbs_batch for o in objects], device)
        self.depth = _ptr_table([o.depth_batch for o in objects], device)
        self.t_wc = _ptr_table([o.t_wc_batch for o in objects], device)
        self.bbox = _ptr_table([o.bbox for o in objects], device)


def sample(rgbs, depth, t_wc, bbox, part_frame, rays_dir, tapes, n_frames, n_samples, n_c2s=1, n_bins=9,
           eps=0.1, other_eps=0.05, min_bound=0.0, part_down=0, part_hw=(0, 0), want_pix=False, out: SampleOut = None,
           tables: RingTables = None, store=None, slot_frame=None, slot_bbox=None, kf_cap=20, obj_ids=None, cache=None):
    """Keyframes come either from private per-object rings -- rgbs/depth/t_wc/bbox: lists (one per object) of the ring tensors
    in the reference layout (u8 [KF,W,H,4], f32 [KF,W,H], f32 [KF,4,4], f32 [KF,4]), or `tables` -- or from a shared
    `store` (framestore.FrameStore) with slot_frame int32 [n_obj,kf_cap] and slot_bbox f32 [n_obj,kf_cap,4] on the device.
    part_frame int32 [n_obj,kf_cap] or None; tapes: SampleTapes (explicit draws) or CounterRng (draws generated in the
    kernel); obj_ids int32 [n_obj] (device): needed with a store and tapes (the pixel state is derived from the id)."""
    n = slot_frame.shape[0] if store is not None else (tables.n if tables is not None else len(rgbs))
    dev = rays_dir.device
    # `cache` (a dict the caller keeps, scene.Scene): the filled argument block of the previous frame is reused when nothing but
    # the frame counter changed -- same output buffers, store, tables, stream -- so the per-frame host work ahead of the
    # launch is one struct field and the call
    if cache is not None and store is not None and isinstance(tapes, CounterRng) and out is not None:
        st = stream()
        key = (id(out), n, n_frames, n_samples, n_c2s, n_bins, store.rgbi.data_ptr(), store.depth.data_ptr(), slot_frame.data_ptr(),
               slot_bbox.data_ptr(), None if part_frame is None else part_frame.data_ptr(), tapes.obj_ids.data_ptr(), int(tapes.seed),
               st.value, float(eps), float(other_eps), float(min_bound))
        if cache.get("key") == key:
            a = cache["args"]
            a.frame = int(tapes.frame)
            with torch.cuda.device(dev):
                check(lib().oo_sample_rays(ctypes.byref(a), st), "oo_sample_rays")
            return out
    else:
        key = None
    W, H = rays_dir.shape[:2]
    n_rays = n_frames * n_samples
    S = n_c2s + n_bins
    if out is None:
        u8 = dict(dtype=torch.uint8, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        out = SampleOut(torch.empty(n, n_rays, 3, **u8), torch.empty(n, n_rays, **f32), torch.empty(n, n_rays, **u8),
                        torch.empty(n, n_rays, **u8), torch.empty(n, n_rays, S, 3, **f32), torch.empty(n, n_rays, S, **f32),
                        torch.empty(n, n_rays, dtype=torch.int32, device=dev) if part_frame is not None else None,
                        torch.empty(n, n_rays, 3, dtype=torch.int64, device=dev) if want_pix else None,
                        torch.zeros(1, dtype=torch.int32, device=dev))
    lin = _lin_tables(S, n_c2s, n_bins)
    a = SampleArgs()
    a.n_obj, a.n_frames, a.n_samples = n, n_frames, n_samples
    a.W, a.H, a.n_c2s, a.n_bins = W, H, n_c2s, n_bins
    a.eps, a.other_eps, a.min_bound = eps, other_eps, min_bound
    a.part_down, a.pw, a.ph = int(part_down), int(part_hw[0]), int(part_hw[1])
    a.kf_cap = int(kf_cap)
    if store is not None:
        assert slot_frame.shape == (n, kf_cap) and slot_bbox.shape == (n, kf_cap, 4)
        a.store_rgbi, a.store_depth, a.store_twc = ptr(store.rgbi), ptr(store.depth), ptr(store.t_wc)
        a.slot_frame, a.slot_bbox = ptr(slot_frame), ptr(slot_bbox)
    else:
        tabs = ([tables.rgbs, tables.depth, tables.t_wc, tables.bbox] if tables is not None
                else [_ptr_table(x, dev) for x in (rgbs, depth, t_wc, bbox)])
        a.rgbs, a.depth, a.t_wc, a.bbox = [ptr(t) for t in tabs]
    assert part_frame is None or tuple(part_frame.shape) == (n, kf_cap), "part_frame must be [n_obj, kf_cap]"
    a.part_frame = ptr(part_frame)
    a.rays_dir = ptr(rays_dir)
    if isinstance(tapes, CounterRng):
        a.rng_mode, a.seed, a.frame = 1, int(tapes.seed), int(tapes.frame)
        a.obj_ids, a.n_keyframes, a.latest = ptr(tapes.obj_ids), ptr(tapes.n_keyframes), ptr(tapes.latest)
        a.tape_by_rank = 0
        need = n * (2 + 2 * n_rays)
        scr = _scratch(dev, need)
        a.scratch, a.scratch_ints = ptr(scr), scr.numel()
    else:
        a.rng_mode = 0
        a.kf_ids, a.u_w, a.u_h = ptr(tapes.kf_ids), ptr(tapes.u_w), ptr(tapes.u_h)
        a.r_invalid, a.r_valid, a.r_normal, a.r_other = (ptr(tapes.r_invalid), ptr(tapes.r_valid), ptr(tapes.r_normal),
                                                         ptr(tapes.r_other))
        a.tape_by_rank = int(tapes.by_rank)
        a.obj_ids = ptr(obj_ids)
    a.lin_s_host, a.lin_c2s_host, a.lin_bins_host = [ctypes.c_void_p(t.data_ptr()) for t in lin]
    a.gt_rgb, a.gt_depth, a.valid, a.labels = ptr(out.gt_rgb), ptr(out.gt_depth), ptr(out.valid), ptr(out.labels)
    a.pcs, a.z, a.feat_row, a.pix, a.oob_count = ptr(out.pcs), ptr(out.z), ptr(out.feat_row), ptr(out.pix), ptr(out.oob)
    with torch.cuda.device(dev):
        check(lib().oo_sample_rays(ctypes.byref(a), stream()), "oo_sample_rays")
    if key is not None:
        cache["key"], cache["args"] = key, a
        cache["keep"] = (out, lin, rays_dir, tapes, scr)       # the block holds raw pointers into these
    return out


def counter_rng(objects, seed, frame, device):
    ids = torch.tensor([o.obj_id for o in objects], dtype=torch.int32)
    nkf = torch.tensor([o.n_keyframes for o in objects], dtype=torch.int32)
    latest = torch.tensor([list(o.lastest_kf_queue)[-2:] if len(o.lastest_kf_queue) >= 2 else [0, 0]
                           for o in objects], dtype=torch.int32)
    pack = torch.cat([ids, nkf, latest.reshape(-1)]).to(device, non_blocking=True)     # one H2D copy
    n = len(objects)
    return CounterRng(seed, frame, pack[:n], pack[n:2 * n], pack[2 * n:].view(n, 2))
