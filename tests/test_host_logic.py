"""CPU tests of host-side logic: keyframe ring policy vs the reference trace, C-ABI exports, layout."""
import ctypes
import json
import os
import random
import re

import numpy as np
import pytest
import torch

from openobj_b200 import layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_keyframe_ring_matches_reference_trace():
    from openobj_b200.vmap import KeyframeRing
    t = json.load(open(os.path.join(ROOT, "tests", "golden", "keyframe_policy.json")))
    random.seed(t["seed"])
    ring = KeyframeRing(0, t["buffer"], t["keyframe_step"])
    slots = {0: 0}
    for rec in t["trace"][1:]:
        s = ring.push(rec["frame"])
        slots[s] = rec["frame"] // 10 % 250
        assert ring.n_keyframes == rec["n_keyframes"] and ring.kf_pointer == rec["kf_pointer"]
        assert ring.latest == rec["latest"] and ring.frame_cnt == rec["frame_cnt"]
        assert [[k, v] for k, v in ring.slot_of.items()] == rec["kf_id_dict"]
        for slot, mark in slots.items():
            assert rec["slot_mark"][slot] == mark


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "openobj_b200.h")).read()
    declared = set(re.findall(r"\b(oo_[a-z0-9_]+)\s*\(", hdr))
    from openobj_b200 import _lib
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    so = _lib.LIB_PATH
    if not os.path.exists(so):
        from openobj_b200 import build
        build.build()
    L = ctypes.CDLL(so)
    for name in declared:
        assert hasattr(L, name), name
    L.oo_version.restype = ctypes.c_int
    assert L.oo_version() == 1
    L.oo_param_offset.restype = ctypes.c_int
    L.oo_param_size.restype = ctypes.c_int
    for i, (off, shp) in enumerate(zip(layout.OFFSETS, layout.SHAPES)):
        assert L.oo_param_offset(i) == off and L.oo_param_size(i) == layout.numel(shp)
    assert L.oo_param_offset(19) == -1


def test_layout_views_roundtrip_and_reference_shapes():
    import openobj_oracle as oc
    fc, B = oc.init_params(4, generator=torch.Generator().manual_seed(0))
    theta = layout.pack(fc + [B])
    for v, t, (name, shp) in zip(layout.views(theta), fc + [B], oc.fc_shapes() + [("B_layer.weight", (21, 3))]):
        assert tuple(v.shape[1:]) == tuple(shp) and torch.equal(v, t)
    assert sum(layout.numel(s) for s in layout.SHAPES) == layout.PCOUNT


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from openobj_b200 import _lib, ops
    from openobj_b200.ensemble import Ensemble
    with pytest.raises(_lib.OOError):
        Ensemble(2, device="cpu")
    with pytest.raises(_lib.OOError):
        ops.forward(torch.zeros(1, layout.PSTRIDE), pcs=torch.zeros(1, 4, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "openobj_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "openobj_oracle" not in src and "ref_harness" not in src and "import philox" not in src, f


def test_reference_checkpoint_keys_are_the_drop_in_state_dict():
    """tests/golden/obj_7.pth was written by the reference's sceneObject.save_checkpoints (vmap.py:556-576): its state-dict
    keys and shapes are exactly what openobj_b200.model.OccupancyMap / embedding.UniDirsEmbed expose (no GPU needed)."""
    from openobj_b200 import embedding, model
    ck = torch.load(os.path.join(os.path.dirname(__file__), "golden", "obj_7.pth"), weights_only=False)
    assert list(ck.keys()) == ["epoch", "FC_state_dict", "PE_state_dict", "obj_id", "bbox", "obj_scale", "clip_feat",
                               "caption_feat", "semantic_id"]
    fc = model.OccupancyMap(87, 42, hidden_size=32, clip_size=512)
    pe = embedding.UniDirsEmbed(max_deg=5, scale=2.0)
    assert list(ck["FC_state_dict"].keys()) == list(fc.state_dict().keys()) == list(layout.NAMES[:18])
    assert {k: tuple(v.shape) for k, v in ck["FC_state_dict"].items()} == {k: tuple(v.shape) for k, v in fc.state_dict().items()}
    assert list(ck["PE_state_dict"].keys()) == list(pe.state_dict().keys())
    fc.load_state_dict(ck["FC_state_dict"])
    pe.load_state_dict(ck["PE_state_dict"])


def test_ring_bank_matches_reference_trace():
    """The array form of the keyframe policy (vmap.RingBank, what scene.Scene runs) against the reference's 140-frame trace."""
    from openobj_b200.vmap import RingBank
    t = json.load(open(os.path.join(ROOT, "tests", "golden", "keyframe_policy.json")))
    random.seed(t["seed"])
    bank = RingBank(3, t["buffer"])
    ring = bank.add(1, 0, t["keyframe_step"])
    for n, rec in enumerate(t["trace"][1:]):
        s = int(bank.push_many(np.array([1]), rec["frame"])[0]) if n % 2 else ring.push(rec["frame"])
        assert ring.n_keyframes == rec["n_keyframes"] and ring.kf_pointer == rec["kf_pointer"]
        assert ring.latest == rec["latest"] and ring.frame_cnt == rec["frame_cnt"]
        assert [[k, v] for k, v in ring.slot_of.items()] == rec["kf_id_dict"]
        assert 0 <= s < t["buffer"]


def test_ring_bank_equals_per_object_rings():
    """RingBank.push_many == one KeyframeRing per object on random schedules: objects appearing at different frames, dropping
    out of view, different keyframe steps, small buffers so that most rings are full and prune with random.choice -- the
    draws must be consumed in the same order."""
    from openobj_b200.vmap import KeyframeRing, RingBank
    for seed in range(24):
        K = [4, 6, 20][seed % 3]
        n_obj, n_frames = 7, 150
        steps = [[2.5, 5.0, 1.0, 2.0][(seed + o) % 4] for o in range(n_obj)]
        rs = np.random.RandomState(seed)
        first, present = rs.randint(0, 10, n_obj), rs.rand(n_frames, n_obj) < 0.85
        state = lambda r, o, s: (o, s, r.n_keyframes, r.kf_pointer, r.kf_buffer_full, r.frame_cnt, list(r.latest), list(r.slot_of.items()))   # noqa: E731
        random.seed(seed)
        rings, ref = {}, []
        for f in range(n_frames):
            for o in range(n_obj):
                if f < first[o] or not present[f, o]:
                    continue
                if o not in rings:
                    rings[o], s = KeyframeRing(f * 10, K, steps[o]), 0
                else:
                    s = rings[o].push(f * 10)
                ref.append(state(rings[o], o, s))
        random.seed(seed)
        bank, have, out = RingBank(n_obj, K), {}, []
        for f in range(n_frames):
            vis = [o for o in range(n_obj) if f >= first[o] and present[f, o]]
            old = [o for o in vis if o in have]
            slots = dict(zip(old, bank.push_many(np.array(old), f * 10).tolist())) if old else {}
            for o in vis:
                if o not in have:
                    have[o], slots[o] = bank.add(o, f * 10, steps[o]), 0
            out += [state(have[o], o, slots[o]) for o in vis]
        assert ref == out, (seed, K)
