"""Seeded synthetic RGB-D + instance-mask (+ part-feature) frames with the dict layout of objnerf/dataset.py:181-186
(SURVEY 8d): image u8 [W,H,3], depth f32 [W,H] metres (0 = invalid), obj int32 [W,H] (-1 unknown, 0 background,
k>0 instance), T float64 [4,4], bbox_dict[id] = int64 [w_lo,w_hi,h_lo,h_hi] enlarged by 0.2 (utils.enlarge_bbox),
part_feat f32 [W/5,H/5,512].  Host tensors (optionally pinned), like a DataLoader would deliver them."""
import math

import numpy as np
import torch

from .utils import enlarge_bbox


def object_layout(n_obj, W, H, rng):
    """n_obj axis-aligned rectangles (>= 40x40 when the frame allows) on a jittered grid covering ~70 % of the frame."""
    cols = int(math.ceil(math.sqrt(n_obj * W / H)))
    rows = int(math.ceil(n_obj / cols))
    cw, ch = W // cols, H // rows
    boxes = []
    for k in range(n_obj):
        cx, cy = (k % cols) * cw, (k // cols) * ch
        bw = max(min(cw - 8, 40), int(cw * rng.uniform(0.75, 0.9)))
        bh = max(min(ch - 8, 40), int(ch * rng.uniform(0.75, 0.9)))
        x0 = cx + rng.integers(2, max(3, cw - bw - 2))
        y0 = cy + rng.integers(2, max(3, ch - bh - 2))
        boxes.append((int(x0), int(y0), int(min(x0 + bw, W - 3)), int(min(y0 + bh, H - 3))))
    return boxes


class SyntheticScene:
    """Frame generator: the same N objects seen from a smoothly moving camera."""

    def __init__(self, n_obj, W=1200, H=680, part_mode=True, part_down=5, clip=512, seed=0, first_id=1, pin=False,
                 n_distinct=4, with_bg=False):
        self.n_obj, self.W, self.H, self.part_mode, self.part_down, self.clip = n_obj, W, H, part_mode, part_down, clip
        self.rng = np.random.default_rng(seed)
        self.ids = list(range(first_id, first_id + n_obj))
        self.boxes = object_layout(n_obj, W, H, self.rng)
        self.pin = pin
        inst = np.zeros((W, H), np.int32)
        for oid, (x0, y0, x1, y1) in zip(self.ids, self.boxes):
            inst[max(x0 - 4, 0):x1 + 4, max(y0 - 4, 0):y1 + 4] = -1          # unknown ring around each object
        for oid, (x0, y0, x1, y1) in zip(self.ids, self.boxes):
            inst[x0:x1, y0:y1] = oid
        self.inst = inst
        self.bbox = {}
        for oid, (x0, y0, x1, y1) in zip(self.ids, self.boxes):
            e = enlarge_bbox([x0, y0, x1 - 1, y1 - 1], 0.2, w=W, h=H)       # note the reference's (w=shape[1], h=shape[0]) order
            self.bbox[oid] = torch.tensor([e[0], e[2], e[1], e[3]], dtype=torch.int64)
        if with_bg:
            self.bbox[0] = torch.tensor([0, W - 1, 0, H - 1], dtype=torch.int64)   # id 0 = background: the whole frame
        # a few distinct payloads are cycled: content does not change the work per frame
        ww, hh = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="ij")
        self._payload = []
        for j in range(n_distinct):
            image = self.rng.integers(0, 256, (W, H, 3), dtype=np.uint8)
            depth = (2.5 + 1.5 * np.sin(ww / W * 3 + j) + 1.2 * np.cos(hh / H * 2 + 0.5 * j)
                     + 0.05 * self.rng.standard_normal((W, H)).astype(np.float32)).astype(np.float32)
            depth = np.clip(depth, 0.5, 6.0)
            depth[self.rng.random((W, H)) < 0.02] = 0.0
            part = None
            if part_mode:
                part = self.rng.standard_normal((W // part_down, H // part_down, clip), dtype=np.float32)
                part[self.rng.random((W // part_down, H // part_down)) < 0.1] = 0.0
            self._payload.append(tuple(None if a is None else self._mk(a) for a in (image, depth, part)))
        self.inst_t = self._mk(self.inst)

    def _mk(self, a):
        t = torch.from_numpy(a)
        return t.pin_memory() if self.pin else t       # pinned once, like a DataLoader with pin_memory=True

    def pose(self, f):
        ang = 0.02 * f
        T = np.eye(4)
        T[0, 0], T[0, 2], T[2, 0], T[2, 2] = math.cos(ang), math.sin(ang), -math.sin(ang), math.cos(ang)
        T[:3, 3] = [0.02 * f, 0.01 * math.sin(0.3 * f), -0.015 * f]
        return torch.from_numpy(T)

    def frame(self, f, stride=10):
        image, depth, part = self._payload[f % len(self._payload)]
        s = {"image": image, "depth": depth, "T": self.pose(f), "obj": self.inst_t,
             "bbox_dict": self.bbox, "frame_id": f * stride}
        if self.part_mode:
            s["part_feat"] = part
        return s

    def frame_bytes(self):
        image, depth, part = self._payload[0]
        nb = lambda t: t.numel() * t.element_size()
        return nb(image) + nb(depth) + self.inst.nbytes + 128 + (nb(part) if part is not None else 0)
