"""CPU check of the fused training tile's index algebra: openobj_b200/csrc/oo_tile.h is compiled for the host
(tests/emu/emu_train.cpp, every phase run for tid 0..255) and its gradients / loss terms are compared with the
oracle on the golden step.  No GPU involved; the emulator is test infrastructure, not a fallback."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import openobj_oracle as oc
from openobj_b200 import layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "tests", "emu", "emu_train.cpp")
    so = os.path.join(ROOT, "tests", "emu", "libemu_train.so")
    deps = [src] + [os.path.join(ROOT, "openobj_b200", "csrc", f) for f in ("oo_tile.h", "oo_layout.h", "oo_sched.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def load(name):
    d = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def run_emu(emu, theta, pcs, z, gt_depth, rgb8, labels, gt_feat, n_sm, it=0, R=None):
    N, RAYS = labels.shape
    R = R or RAYS
    lab = labels[:, it * R:(it + 1) * R]
    counts = torch.stack([(lab == 1).sum(1), (lab != 2).sum(1)], 1).to(torch.int32).contiguous()
    flags = (2 if bool((counts[:, 0] == 0).any()) else 0) | (4 if bool((counts[:, 1] == 0).any()) else 0)
    grads = torch.zeros(N, layout.PSTRIDE)
    terms = torch.zeros(N, 4)
    if gt_feat is not None:
        table = gt_feat.reshape(-1, 512).contiguous()
        rows = torch.arange(N * RAYS, dtype=torch.int32).reshape(N, RAYS).contiguous()
        fr, ft = ptr(rows), ptr(table)
    else:
        fr, ft = None, None
    pcs, z, gt_depth, rgb8, labels = [t.contiguous() for t in (pcs, z, gt_depth, rgb8, labels)]
    rc = emu.emu_train_grads(ptr(theta), N, ptr(pcs), ptr(z), ptr(gt_depth), ptr(rgb8), ptr(labels), fr, ft,
                             RAYS, it, R, ctypes.c_float(2.0), ptr(counts), flags, n_sm, ptr(grads), ptr(terms))
    assert rc == 0
    return grads, terms, flags


@pytest.mark.parametrize("mode", ["on", "off", "zm"])
@pytest.mark.parametrize("n_sm", [148, 4, 1])
def test_emulated_tile_matches_oracle(emu, mode, n_sm):
    ms = load("model_step.npz")
    fc = [ms["fc%02d" % i] for i in range(18)]
    B = ms["peB"]
    theta = layout.pack(fc + [B])
    labels = ms["labels_zm"] if mode == "zm" else ms["labels"]
    gt_feat = None if mode == "off" else ms["gt_feat"]
    grads, terms, flags = run_emu(emu, theta, ms["pcs"], ms["z"], ms["gt_depth"], ms["gt_rgb8"], labels, gt_feat, n_sm)
    assert not torch.isnan(grads).any() and not torch.isnan(terms).any()
    ref_terms, ref_grads = oc.train_step_grads(fc, B, ms["pcs"], ms["z"], ms["gt_depth"], ms["gt_rgb8"] / 255.,
                                               labels, gt_feat)
    assert flags == (ref_terms.flags & 6)
    ref_t = torch.stack([ref_terms.depth, ref_terms.color, ref_terms.opacity, ref_terms.feat], 1)
    torch.testing.assert_close(terms, ref_t, rtol=1e-4, atol=1e-6)
    total = (terms[:, 0] + 5 * terms[:, 1] + 10 * terms[:, 2] + 5 * terms[:, 3]).sum()
    torch.testing.assert_close(total, ms["loss_" + mode], rtol=1e-4, atol=1e-6)
    got = layout.views(grads)
    for i, (g, r) in enumerate(zip(got, ref_grads)):
        r = torch.zeros_like(g) if r is None else r
        scale = float(r.abs().max()) + 1e-12
        err = float((g - r).abs().max())
        assert err <= 2e-4 * scale + 1e-7, (i, layout.NAMES[i], err, scale)
    # pad floats of the block must stay zero
    pad = torch.ones(layout.PSTRIDE, dtype=torch.bool)
    for off, shp in zip(layout.OFFSETS, layout.SHAPES):
        pad[off:off + layout.numel(shp)] = False
    assert float(grads[:, pad].abs().max()) == 0.0


def test_emulated_tile_multi_step_slices(emu):
    """rays_per_obj = 2 steps x 8 rays: step 1 must read the second slice (train.py:396-404)."""
    ms = load("model_step.npz")
    fc = [ms["fc%02d" % i] for i in range(18)]
    B = ms["peB"]
    theta = layout.pack(fc + [B])
    grads, terms, _ = run_emu(emu, theta, ms["pcs"], ms["z"], ms["gt_depth"], ms["gt_rgb8"], ms["labels"],
                              ms["gt_feat"], 148, it=1, R=8)
    sl = slice(8, 16)
    ref_terms, ref_grads = oc.train_step_grads(fc, B, ms["pcs"][:, sl], ms["z"][:, sl], ms["gt_depth"][:, sl],
                                               ms["gt_rgb8"][:, sl] / 255., ms["labels"][:, sl], ms["gt_feat"][:, sl])
    ref_t = torch.stack([ref_terms.depth, ref_terms.color, ref_terms.opacity, ref_terms.feat], 1)
    torch.testing.assert_close(terms, ref_t, rtol=1e-4, atol=1e-6)
    for g, r in zip(layout.views(grads), ref_grads):
        assert float((g - r).abs().max()) <= 2e-4 * float(r.abs().max()) + 1e-7


def test_schedule_properties(emu):
    buf = (ctypes.c_int * 4096)()
    for n_obj, R, n_sm in [(60, 120, 148), (8, 120, 148), (13, 120, 148), (200, 120, 148), (3, 16, 4), (1, 5, 148)]:
        n = emu.emu_schedule(n_obj, R, n_sm, buf, 4096)
        assert n > 0
        n_cta, n_slots, tpo = buf[0], buf[1], buf[2]
        data = list(buf[3:n])
        ct, cs, osl = data[:n_cta + 1], data[n_cta + 1:2 * n_cta + 1], data[2 * n_cta + 1:]
        T = n_obj * tpo
        assert tpo == -(-R // 10) and ct[0] == 0 and ct[-1] == T and n_cta == min(n_sm, T)
        sizes = [b - a for a, b in zip(ct[:-1], ct[1:])]
        assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
        # slots: one per (cta, object) run, contiguous per object
        runs = []
        for c in range(n_cta):
            objs = sorted({t // tpo for t in range(ct[c], ct[c + 1])})
            assert cs[c] == len(runs)
            runs += objs
        assert len(runs) == n_slots and runs == sorted(runs)
        assert osl[0] == 0 and osl[-1] == n_slots
        for o in range(n_obj):
            assert runs[osl[o]:osl[o + 1]] == [o] * (osl[o + 1] - osl[o]) and osl[o + 1] > osl[o]
