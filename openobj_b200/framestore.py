"""Shared keyframe store: ONE device copy of every frame that some object still holds as a keyframe,
instead of the reference's private ring per object (`rgbs_batch[20,W,H,4]`, `depth_batch[20,W,H]`: 130 MB per object at
Replica size, objnerf/vmap.py:84-143).  An object's ring slot is an index into this store (`slot_frame[o][s]`); the
per-object pixel state of train.py:203-205 is derived in the sampling kernel from the stored instance map.  Memory is
O(frames alive), not O(objects): a frame is dropped when the last ring slot referring to it is overwritten."""
import ctypes

import numpy as np
import torch

from ._lib import StoreArgs, check, lib, ptr, stream


class FrameStore:
    def __init__(self, W, H, device, capacity=32):
        self.W, self.H, self.device = int(W), int(H), torch.device(device)
        self.capacity = 0
        self.rgbi = self.depth = self.t_wc = None
        self.refs, self.free = np.zeros(0, dtype=np.int64), []
        self._grow(int(capacity))

    # ---- storage ---------------------------------------------------------------------------------------------------------
    def _grow(self, capacity):
        dev, W, H = self.device, self.W, self.H
        rgbi = torch.empty(capacity, W, H, 2, dtype=torch.int32, device=dev)     # word 0 = r | g << 8 | b << 16, word 1 = instance id
        depth = torch.empty(capacity, W, H, dtype=torch.float32, device=dev)
        t_wc = torch.empty(capacity, 16, dtype=torch.float32, device=dev)
        if self.capacity:
            rgbi[:self.capacity].copy_(self.rgbi)
            depth[:self.capacity].copy_(self.depth)
            t_wc[:self.capacity].copy_(self.t_wc)
        self.free = list(range(capacity - 1, self.capacity - 1, -1)) + self.free
        self.refs = np.concatenate([self.refs, np.zeros(capacity - self.capacity, dtype=np.int64)])
        self.rgbi, self.depth, self.t_wc, self.capacity = rgbi, depth, t_wc, capacity

    def bytes_per_frame(self):
        return self.W * self.H * 12 + 64

    def frames_alive(self):
        return self.capacity - len(self.free)

    # ---- one new frame -> one store slot (one launch, 12 bytes per pixel) -----------------------------------------------------
    def alloc(self):
        """A free slot (the store grows when none is left).  The caller takes references with acquire(); a slot nobody
        acquired goes back with drop_if_unreferenced()."""
        if not self.free:
            self._grow(2 * self.capacity)
        return self.free.pop()

    def write(self, slot, rgb, depth, inst, t_wc_dev):
        """rgb u8 [W,H,3], depth f32 [W,H], inst int32 [W,H], t_wc_dev f32 [16] -- all on the device -> store slot `slot`."""
        a = StoreArgs()
        a.W, a.H, a.slot = self.W, self.H, slot
        a.rgb, a.depth, a.inst, a.t_wc = ptr(rgb.contiguous()), ptr(depth.contiguous()), ptr(inst.contiguous()), ptr(t_wc_dev)
        a.store_rgbi, a.store_depth, a.store_twc = ptr(self.rgbi), ptr(self.depth), ptr(self.t_wc)
        with torch.cuda.device(self.device):
            check(lib().oo_store_frame(ctypes.byref(a), stream()), "oo_store_frame")

    def acquire(self, slot, n=1):
        self.refs[slot] += n

    def release_many(self, slots):
        """Drop one reference per entry of `slots` (int array, -1 = nothing)."""
        slots = slots[slots >= 0]
        if slots.size == 0:
            return
        np.subtract.at(self.refs, slots, 1)
        for s in np.unique(slots).tolist():
            if self.refs[s] <= 0:
                self.refs[s] = 0
                self.free.append(s)

    def release(self, slot):
        """Drop one reference; a slot nobody refers to goes back to the free list.  (Stream order makes the reuse safe: the
        kernel that overwrites it is enqueued after every kernel that read it.)"""
        if slot < 0:
            return
        self.refs[slot] -= 1
        if self.refs[slot] <= 0:
            self.refs[slot] = 0
            self.free.append(slot)

    def drop_if_unreferenced(self, slot):
        if self.refs[slot] == 0 and slot not in self.free:
            self.free.append(slot)
