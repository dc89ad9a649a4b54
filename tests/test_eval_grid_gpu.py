"""Trainer.eval_points / the query grid of Trainer.meshing (objnerf/trainer.py:46-128,
render_rays.py:119-146) on the GPU, against golden vectors frozen from the reference (tests/golden/eval_grid.npz,
oracle/make_golden.py::gen_eval_grid) and against the oracle at a larger, ragged size."""
import os
import types

import numpy as np
import pytest
import torch

import openobj_oracle as oc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def load(name):
    d = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def make_trainer(d, tag):
    from openobj_b200 import cfg as C, trainer as T
    cfg = C.room0_config(w=40, h=30)
    cfg.obj_id = 1 if tag == "obj" else 0
    if tag == "bg":
        cfg.hidden_feature_size, cfg.obj_scale = cfg.hidden_feature_size_bg, cfg.bg_scale
    cfg.training_device = DEV
    tr = T.Trainer(cfg)
    with torch.no_grad():
        for p, i in zip(tr.fc_occ_map.parameters(), range(18)):
            p.copy_(d["%s_fc%02d" % (tag, i)][0].to(DEV))
        tr.pe.B_layer.weight.copy_(d[tag + "_peB"][0].to(DEV))
    assert tr.bound_extent == pytest.approx(float(d[tag + "_bound_extent"]))
    return tr


def bound_of(d):
    return types.SimpleNamespace(R=d["obb_R"].numpy(), center=d["obb_center"].numpy(), extent=d["obb_extent"].numpy())


@pytest.mark.parametrize("tag", ["obj", "bg"])
def test_eval_grid_matches_reference_golden(tag):
    """obj: hidden 32 through the fused forward tile (one launch); bg: hidden 128 through the layer-by-layer path."""
    d = load("eval_grid.npz")
    tr = make_trainer(d, tag)
    dim = int(d[tag + "_dim"])
    grid, occ, color, clip = tr.eval_grid(bound_of(d), d["obj_center"], grid_dim=dim, want_clip=True)
    torch.testing.assert_close(grid.cpu(), d[tag + "_grid"], rtol=1e-6, atol=1e-6)
    # alpha is held to abs 1e-4 (x10 output, test_forward_matches_reference_golden); sigmoid' <= 1/4
    torch.testing.assert_close(occ.cpu(), d[tag + "_occ"], rtol=1e-4, atol=2.5e-5)
    torch.testing.assert_close(color.cpu(), d[tag + "_color"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(clip.cpu()[::7], d[tag + "_clip"], rtol=1e-4, atol=1e-4)
    # eval_points on the reference's own grid tensor gives the same numbers (chunk_size is accepted and irrelevant)
    occ2, color2, _ = tr.eval_points(d[tag + "_grid"].to(DEV), chunk_size=500, want_clip=False)
    torch.testing.assert_close(occ2.cpu(), d[tag + "_occ"], rtol=1e-4, atol=2.5e-5)
    torch.testing.assert_close(color2.cpu(), d[tag + "_color"], rtol=1e-4, atol=1e-5)


def test_eval_grid_large_ragged_vs_oracle():
    """37^3 = 50 653 points (not a multiple of the 100-point tile): grid bit-identical to the oracle's torch expressions up
    to 1 ulp, occupancy / colour against the oracle, and consistency with the per-point forward entry."""
    from openobj_b200 import ops
    d = load("eval_grid.npz")
    tr = make_trainer(d, "obj")
    dim = 37
    grid, occ, color, clip = tr.eval_grid(bound_of(d), d["obj_center"], grid_dim=dim, want_clip=False)
    assert clip is None and grid.shape == (dim ** 3, 3)
    ref_grid = oc.meshing_grid(d["obb_R"], d["obb_center"], d["obb_extent"], tr.bound_extent, dim, d["obj_center"])
    torch.testing.assert_close(grid.cpu(), ref_grid, rtol=1e-6, atol=1e-6)
    fc = [d["obj_fc%02d" % i] for i in range(18)]
    r_occ, r_color, _ = oc.eval_points(fc, d["obj_peB"], ref_grid, scale=2.0)
    torch.testing.assert_close(occ.cpu(), r_occ, rtol=1e-4, atol=2.5e-5)
    torch.testing.assert_close(color.cpu(), r_color, rtol=1e-4, atol=1e-5)
    a, c, _, _ = ops.forward(tr.packed(DEV), pcs=grid[None], scale=2.0, want_clip=False)
    # eval_grid without the part feature runs on the tcgen05 / TMEM kernel, ops.forward on the mma.sync tile: same algorithm,
    # same three-term TF32 compensation, different accumulation order
    torch.testing.assert_close(c[0], color, rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(occ, ops.occupancy_activation(a[0, :, 0]), rtol=1e-5, atol=5e-6)
    o_mm, c_mm, _ = ops.eval_points(tr.packed(DEV), grid, scale=2.0, want_clip=False, tensor_core=False)
    assert torch.equal(c_mm, c[0]) and torch.equal(o_mm, ops.occupancy_activation(a[0, :, 0]))     # same tile phases, same order
    ops.check_tc(torch.device(DEV))
    # volume layout: (i, j, k) -> (i*dim + j)*dim + k, the reference's .view(dim, dim, dim)
    vol = occ.view(dim, dim, dim)
    i, j, k = 3, 17, 30
    assert float(vol[i, j, k]) == float(occ[(i * dim + j) * dim + k])


def test_eval_points_none_when_nothing_is_occupied_and_cpu_points_raise():
    d = load("eval_grid.npz")
    tr = make_trainer(d, "obj")
    with torch.no_grad():
        tr.fc_occ_map.out_alpha.weight.zero_()
        tr.fc_occ_map.out_alpha.bias.fill_(-100.0)      # alpha = -1000: sigmoid underflows to exactly 0 (trainer.py:124-126)
    assert tr.eval_points(torch.randn(250, 3, device=DEV)) is None
    with pytest.raises(RuntimeError):
        tr.eval_points(torch.randn(10, 3))


def test_occupancy_activation_with_distances():
    from openobj_b200 import ops, render_rays
    g = torch.Generator().manual_seed(5)
    a = torch.randn(1000, generator=g) * 4
    dist = torch.rand(1000, generator=g)
    torch.testing.assert_close(ops.occupancy_activation(a.to(DEV)).cpu(), torch.sigmoid(a), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(ops.occupancy_activation(a.to(DEV), dist.to(DEV)).cpu(),
                               render_rays.occupancy_activation(a, dist), rtol=1e-5, atol=1e-6)


def test_reference_checkpoint_loads_and_round_trips(tmp_path):
    """tests/golden/obj_7.pth was written by the REFERENCE's sceneObject.save_checkpoints
    (vmap.py:556-576).  Our sceneObject.load_checkpoints reads it, eval_points reproduces what the reference's
    Trainer.eval_points returned for it, and save_checkpoints writes the same keys / tensors back bit for bit."""
    from openobj_b200 import cfg as C, vmap as V
    exp = load("ckpt_expect.npz")
    cfg = C.room0_config(w=40, h=30)
    cfg.training_device = cfg.data_device = DEV
    W, H = cfg.W, cfg.H
    obj = V.sceneObject(cfg, 1, torch.zeros(W, H, 3, dtype=torch.uint8, device=DEV), torch.ones(W, H, device=DEV),
                        torch.ones(W, H, dtype=torch.uint8, device=DEV), torch.tensor([0., W - 1, 0., H - 1]),
                        torch.eye(4, device=DEV), 0)
    src = os.path.join(GOLDEN, "obj_7.pth")
    assert obj.load_checkpoints(src) is True
    assert obj.obj_id == 7 and obj.semantic_id == 3 and obj.bbox_final and obj.trainer.obj_scale == 2.0
    torch.testing.assert_close(obj.clip_feat, exp["clip_feat"], rtol=0, atol=0)
    occ, color, clip = obj.trainer.eval_points(exp["points"].to(DEV))
    torch.testing.assert_close(occ.cpu(), exp["occ"], rtol=1e-4, atol=2.5e-5)
    torch.testing.assert_close(color.cpu(), exp["color"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(clip.cpu()[:8], exp["clip"], rtol=1e-4, atol=1e-4)
    obj.save_checkpoints(str(tmp_path), 42)
    a = torch.load(src, weights_only=False)
    b = torch.load(os.path.join(str(tmp_path), "obj_7.pth"), weights_only=False)
    assert list(a.keys()) == list(b.keys()) and b["epoch"] == 42
    for part in ("FC_state_dict", "PE_state_dict"):
        assert list(a[part].keys()) == list(b[part].keys())
        for k in a[part]:
            assert torch.equal(a[part][k], b[part][k].cpu()), (part, k)
    assert obj.load_checkpoints(os.path.join(str(tmp_path), "missing.pth")) is None     # "ckpt not exist"
