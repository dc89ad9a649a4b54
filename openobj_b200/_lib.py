"""ctypes binding of libopenobj_b200.so (include/openobj_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, an exception is
raised.  Device memory, streams and torch.distributed are torch's; the arithmetic is the library's.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_uint64, c_void_p

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libopenobj_b200.so")


class OOError(RuntimeError):
    pass


class Batch(Structure):
    _fields_ = [("pcs", c_void_p), ("z", c_void_p), ("gt_depth", c_void_p), ("gt_rgb", c_void_p),
                ("labels", c_void_p), ("feat_row", c_void_p), ("feat_table", c_void_p), ("rays_per_obj", c_int)]


class TrainWs(Structure):
    _fields_ = [("slab", c_void_p), ("slot_loss", c_void_p), ("derived", c_void_p), ("clip_grad", c_void_p), ("rayrec", c_void_p), ("sched", c_void_p),
                ("counts", c_void_p), ("flags", c_void_p), ("adam_scal", c_void_p), ("adam_t", c_void_p),
                ("gram_part", c_void_p), ("gram_cnt", c_void_p)]


class SampleArgs(Structure):
    _fields_ = [("n_obj", c_int), ("n_frames", c_int), ("n_samples", c_int),
                ("W", c_int), ("H", c_int), ("n_c2s", c_int), ("n_bins", c_int),
                ("eps", c_float), ("other_eps", c_float), ("min_bound", c_float),
                ("part_down", c_int), ("pw", c_int), ("ph", c_int),
                ("rgbs", c_void_p), ("depth", c_void_p), ("t_wc", c_void_p), ("bbox", c_void_p),
                ("part_frame", c_void_p), ("rays_dir", c_void_p),
                ("kf_ids", c_void_p), ("u_w", c_void_p), ("u_h", c_void_p),
                ("r_invalid", c_void_p), ("r_valid", c_void_p), ("r_normal", c_void_p), ("r_other", c_void_p),
                ("tape_by_rank", c_int),
                ("lin_s_host", c_void_p), ("lin_c2s_host", c_void_p), ("lin_bins_host", c_void_p),
                ("rng_mode", c_int), ("seed", c_uint64), ("frame", c_uint32),
                ("obj_ids", c_void_p), ("n_keyframes", c_void_p), ("latest", c_void_p),
                ("gt_rgb", c_void_p), ("gt_depth", c_void_p), ("valid", c_void_p), ("labels", c_void_p),
                ("pcs", c_void_p), ("z", c_void_p), ("feat_row", c_void_p), ("pix", c_void_p),
                ("oob_count", c_void_p), ("kf_cap", c_int),
                ("store_rgbi", c_void_p), ("store_depth", c_void_p), ("store_twc", c_void_p),
                ("slot_frame", c_void_p), ("slot_bbox", c_void_p), ("scratch", c_void_p), ("scratch_ints", c_int64)]


class StoreArgs(Structure):
    _fields_ = [("W", c_int), ("H", c_int), ("slot", c_int), ("rgb", c_void_p), ("depth", c_void_p), ("inst", c_void_p),
                ("t_wc", c_void_p), ("store_rgbi", c_void_p), ("store_depth", c_void_p), ("store_twc", c_void_p)]


class Grid(Structure):
    """oo_grid (include/openobj_b200.h)."""
    _fields_ = [("dim", c_int), ("t", c_void_p), ("scale", c_float * 3), ("transform", c_float * 12),
                ("center", c_float * 3)]


class RenderArgs(Structure):
    _fields_ = [("W", c_int), ("H", c_int), ("n_bins", c_int), ("scale", c_float),
                ("theta1", c_void_p), ("T_wc", c_void_p), ("T_oc", c_void_p), ("half_extent", c_void_p),
                ("rays_dir", c_void_p), ("jitter", c_void_p), ("jitter_by_rank", c_int), ("lin_host", c_void_p),
                ("mask", c_void_p), ("depth", c_void_p), ("rgb", c_void_p), ("feat", c_void_p),
                ("opacity", c_void_p), ("n_hit", c_void_p), ("ray_rec", c_void_p), ("hit_pix", c_void_p),
                ("force_mma_sync", c_int), ("tc_err", c_void_p)]


class RenderFrameArgs(Structure):
    _fields_ = [("W", c_int), ("H", c_int), ("n_bins", c_int), ("n_obj", c_int), ("scale", c_float),
                ("theta", c_void_p), ("T_wc", c_void_p), ("T_oc", c_void_p), ("half_extent", c_void_p), ("rays_dir", c_void_p),
                ("jitter", c_void_p), ("lin", c_void_p), ("mask", c_void_p), ("depth", c_void_p), ("rgb", c_void_p),
                ("obj_start", c_void_p), ("hit_pix", c_void_p), ("ray_rec", c_void_p), ("pool_rows", c_int64),
                ("scratch", c_void_p), ("scratch_ints", c_int64), ("tc_err", c_void_p)]


_SIGS = {
    "oo_version": ([], c_int),
    "oo_last_error": ([], c_char_p),
    "oo_param_offset": ([c_int], c_int),
    "oo_param_size": ([c_int], c_int),
    "oo_forward": ([c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "oo_loss_fwd": ([c_void_p] * 8 + [c_int] * 4 + [c_float] * 3 + [c_void_p] * 5, c_int),
    "oo_loss_bwd": ([c_void_p] * 8 + [c_int] * 4 + [c_float] * 4 + [c_void_p] * 6, c_int),
    "oo_loss_ws_per_ray": ([], c_int),
    "oo_train_ws_sizes": ([c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int64),
                           POINTER(c_int64)], c_int),
    "oo_train_schedule": ([c_int, c_int, c_int, POINTER(TrainWs), c_void_p], c_int),
    "oo_label_counts": ([c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "oo_adam_schedule": ([c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p], c_int),
    "oo_train_step": ([c_void_p, c_void_p, c_void_p, c_int, POINTER(Batch), c_int, c_int, c_float, c_float, c_float,
                       c_float, c_float, c_float, POINTER(TrainWs), c_void_p, c_int, c_void_p], c_int),
    "oo_train_grads": ([c_void_p, c_int, POINTER(Batch), c_int, c_int, c_float, POINTER(TrainWs), c_void_p, c_void_p,
                        c_int, c_void_p], c_int),
    "oo_train_frame": ([c_void_p, c_void_p, c_void_p, c_int, POINTER(Batch), c_int, c_int, c_float, c_float, c_float,
                        c_float, c_float, c_float, POINTER(TrainWs), c_void_p, c_int, c_void_p], c_int),
    "oo_train_k1": ([c_void_p, c_int, POINTER(Batch), c_int, c_int, c_float, POINTER(TrainWs), c_int, c_int, c_void_p], c_int),
    "oo_train_k4": ([c_void_p, c_void_p, c_void_p, c_int, POINTER(Batch), c_int, c_int, c_float, c_float, c_float, c_float, c_float,
                     POINTER(TrainWs), c_void_p, c_int, c_void_p], c_int),
    "oo_adamw_flat": ([c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_float, c_float,
                       c_float, c_void_p], c_int),
    "oo_sample_rays": ([POINTER(SampleArgs), c_void_p], c_int),
    "oo_store_frame": ([POINTER(StoreArgs), c_void_p], c_int),
    "oo_gather_part_rows": ([c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p], c_int),
    "oo_rng_fill": ([c_uint64, c_uint32, c_void_p, c_int, c_int64, c_int, c_float, c_void_p, c_void_p], c_int),
    "oo_rng_fill_rows": ([c_uint64, c_uint32, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p], c_int),
    "oo_render_object": ([POINTER(RenderArgs), c_void_p], c_int),
    "oo_zmerge": ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p],
                  c_int),
    "oo_render_frame": ([POINTER(RenderFrameArgs), c_void_p], c_int),
    "oo_winner_features_frame": ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_void_p], c_int),
    "oo_zmerge_ptr": ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "oo_winner_features": ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                            c_void_p], c_int),
    "oo_make_grid": ([POINTER(Grid), c_void_p, c_void_p], c_int),
    "oo_eval_points": ([c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "oo_eval_points_tc": ([c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "oo_occupancy_activation": ([c_void_p, c_void_p, c_int64, c_void_p, c_void_p], c_int),
    "oo_ray_box": ([c_void_p, c_void_p, POINTER(c_float), POINTER(c_float), c_int64, c_void_p, c_void_p, c_void_p, c_void_p],
                   c_int),
    "oo_origin_dirs": ([c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p], c_int),
    "oo_stratified_bins": ([c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_int64, c_int, c_void_p, c_void_p], c_int),
    "oo_normal_bins": ([c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p], c_int),
    "oo_ray_points": ([c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, POINTER(c_float), c_void_p, c_void_p, c_void_p],
                      c_int),
    "oo_termination": ([c_void_p, c_int64, c_int, c_void_p, c_void_p], c_int),
    "oo_render_sum": ([c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p], c_int),
    "oo_fma_peak": ([c_int, c_int, c_void_p, c_void_p], c_int),
    "oo_bg_param_count": ([c_int], c_int),
    "oo_bg_param_offset": ([c_int, c_int], c_int),
    "oo_bg_param_size": ([c_int, c_int], c_int),
    "oo_bg_ws_floats": ([c_int, c_int, c_int], c_int64),
    "oo_bg_forward": ([c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                       c_void_p], c_int),
    "oo_bg_forward_bwd": ([c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p], c_int),
    "oo_forward_bwd_ws_floats": ([c_int, c_int], c_int64),
    "oo_forward_bwd": ([c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                        c_void_p], c_int),
    "oo_embed_bwd_ws_floats": ([c_int, c_int], c_int64),
    "oo_embed_bwd": ([c_void_p, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "oo_bg_train_step": ([c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_void_p, c_int, c_int, c_float, c_void_p, c_float, c_float, c_float, c_float, c_float, c_float,
                          c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
}

EXPORTS = tuple(_SIGS)
_lib = None


class _Lib:
    """Binds each C entry point on first use; a symbol missing from the .so raises OOError (never a fallback)."""

    def __init__(self, cdll):
        self._cdll = cdll

    def __getattr__(self, name):
        if name not in _SIGS:
            raise AttributeError(name)
        try:
            fn = getattr(self._cdll, name)
        except AttributeError:
            raise OOError("libopenobj_b200.so does not export %s (stale build? run python -m openobj_b200.build)" % name)
        fn.argtypes, fn.restype = _SIGS[name]
        setattr(self, name, fn)
        return fn


def lib():
    """The loaded library; raises if it has not been built (python -m openobj_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OOError("%s is missing: build it with `python -m openobj_b200.build` "
                          "(there is no CPU or PyTorch fallback)" % LIB_PATH)
        _lib = _Lib(ctypes.CDLL(LIB_PATH))
        if _lib.oo_version() != 1:
            raise OOError("libopenobj_b200.so ABI version mismatch")
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().oo_last_error()
        raise OOError("%s failed (%d): %s" % (what or "openobj_b200 call", rc, msg.decode() if msg else "?"))


def ptr(t):
    """Raw device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise OOError("openobj_b200 kernels need CUDA tensors; got a %s tensor (no CPU fallback)" % t.device)
    if not t.is_contiguous():
        raise OOError("openobj_b200 kernels need contiguous tensors")
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def n_sm(device=None):
    return torch.cuda.get_device_properties(device or torch.cuda.current_device()).multi_processor_count
