"""the stand-alone helpers of the reference surface -- utils.{stratified_bins, normal_bins_sampling,
origin_dirs_W, ray_box_intersection}, Trainer.sample_points_bbox, sceneObject.sample_3d_points -- on the GPU, against
golden vectors frozen from the reference on the CPU with every random draw recorded (tests/golden/surface.npz,
oracle/make_golden.py::gen_surface).  Tolerances: rel 1e-5 / abs 1e-6 (torch.linspace on the device may differ from the
CPU table in the last bit; the 3-term dot products may be contracted differently); sorted / clipped / compared values exact."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"
TOL = dict(rtol=1e-5, atol=1e-6)


def load(name):
    d = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def test_utils_helpers_match_reference():
    from openobj_b200 import utils as U
    d = load("surface.npz")
    z = U.stratified_bins(d["sb_min"].to(DEV), d["sb_max"].to(DEV), 7, 33, device=DEV, draws=d["sb_u"])
    torch.testing.assert_close(z.cpu(), d["sb_z"], **TOL)
    z = U.stratified_bins(0.0, 3.5, 10, 20, device=DEV, draws=d["sbs_u"])
    torch.testing.assert_close(z.cpu(), d["sbs_z"], **TOL)
    z = U.stratified_bins(0.0, 3.5, 10, 20, device=DEV)                                   # own draws: bin structure
    edges = torch.linspace(0, 3.5, 11, device=DEV)
    assert bool((z >= edges[:-1] - 1e-6).all()) and bool((z <= edges[1:] + 1e-6).all())
    z = U.normal_bins_sampling(d["nb_depth"].to(DEV), 9, 25, 0.1, device=DEV, draws=d["nb_draws"])
    assert torch.equal(z.cpu(), d["nb_z"])
    z = U.normal_bins_sampling(d["nb_depth"].to(DEV), 9, 25, 0.1, device=DEV)             # own draws: sorted, within +-delta
    assert bool((z[:, 1:] >= z[:, :-1]).all()) and float((z - d["nb_depth"].to(DEV)[:, None]).abs().max()) <= 0.1 + 1e-6
    o1, w1 = U.origin_dirs_W(d["od_T"].to(DEV), d["od_d1"].to(DEV))
    o2, w2 = U.origin_dirs_W(d["od_T"].to(DEV), d["od_d2"].to(DEV))
    assert torch.equal(o1.cpu(), d["od_o1"]) and torch.equal(o2.cpu(), d["od_o2"])
    torch.testing.assert_close(w1.cpu(), d["od_w1"], **TOL)
    torch.testing.assert_close(w2.cpu(), d["od_w2"], **TOL)
    near, far, hit = U.ray_box_intersection(d["rb_o"].to(DEV), d["rb_d"].to(DEV), d["rb_min"], d["rb_max"])
    assert torch.equal(near.cpu(), d["rb_near"]) and torch.equal(far.cpu(), d["rb_far"]) and torch.equal(hit.cpu(), d["rb_hit"])
    with pytest.raises(RuntimeError):
        U.ray_box_intersection(d["rb_o"], d["rb_d"], d["rb_min"], d["rb_max"])            # CPU tensors: no fallback


def test_trainer_sample_points_bbox_matches_reference():
    from openobj_b200 import cfg as C, trainer as T
    d = load("surface.npz")
    cfg = C.room0_config(w=16, h=12)
    cfg.obj_id = 1
    cfg.training_device = DEV
    tr = T.Trainer(cfg)
    tr.T_WC_gt, tr.dirs_C_gt = d["spb_T"].to(DEV), d["spb_dirs"].to(DEV)
    bb = types.SimpleNamespace(R=d["spb_R"].numpy(), center=d["spb_center"].numpy(), extent=d["spb_extent"].numpy())
    hit, near, far = tr.sample_points_bbox(bb, do_eval=True, draws=d["spb_u"])
    assert torch.equal(hit.cpu(), d["spb_hit"])
    torch.testing.assert_close(near.cpu(), d["spb_near"], **TOL)
    torch.testing.assert_close(far.cpu(), d["spb_far"], **TOL)
    torch.testing.assert_close(tr.z_vals_cat.cpu(), d["spb_zcat"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(tr.z_vals.cpu(), d["spb_z"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(tr.input_pcs.cpu(), d["spb_pcs"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(tr.dirs_W.cpu(), d["spb_dirsW"], **TOL)
    assert tr.z_vals.shape == (54, 149) and tr.input_pcs.shape == (54, 149, 3)
    # a box behind the camera: at most one hit -> (None, None, None) (trainer.py:165-166)
    far_box = types.SimpleNamespace(R=np.eye(3), center=np.array([0.0, 0.0, -50.0]), extent=np.array([0.1, 0.1, 0.1]))
    assert tr.sample_points_bbox(far_box) == (None, None, None)


def test_scene_object_sample_3d_points_matches_reference():
    from openobj_b200 import cfg as C, vmap as V
    d = load("surface.npz")
    cfg = C.room0_config(w=40, h=30)
    cfg.training_device = cfg.data_device = DEV
    W, H = cfg.W, cfg.H
    obj = V.sceneObject(cfg, 1, torch.zeros(W, H, 3, dtype=torch.uint8, device=DEV), torch.ones(W, H, device=DEV),
                        torch.ones(W, H, dtype=torch.uint8, device=DEV), torch.tensor([0., W - 1, 0., H - 1]),
                        torch.eye(4, device=DEV), 0)
    draws = tuple(d[k] for k in ("s3_u_inv", "s3_u_val", "s3_n_obj", "s3_u_oth"))
    rgb, dep, valid, labels, pcs, z, pf = obj.sample_3d_points(d["s3_rgbs"].to(DEV), d["s3_depth"].to(DEV), d["s3_origins"].to(DEV),
                                                               d["s3_dirs"].to(DEV), draws=draws)
    assert torch.equal(valid.cpu(), d["s3_valid"]) and torch.equal(labels.cpu(), d["s3_labels"]) and pf is None
    assert torch.equal(rgb.cpu(), d["s3_rgbs"][..., :3])
    torch.testing.assert_close(z.cpu(), d["s3_z"], **TOL)
    torch.testing.assert_close(pcs.cpu(), d["s3_pcs"], rtol=1e-5, atol=1e-5)
    # without supplied draws the generator is consumed in the reference's order and the structure holds
    torch.manual_seed(3)
    _, _, _, _, pcs2, z2, _ = obj.sample_3d_points(d["s3_rgbs"].to(DEV), d["s3_depth"].to(DEV), d["s3_origins"].to(DEV),
                                                   d["s3_dirs"].to(DEV))
    assert z2.shape == (3, 8, 10) and pcs2.shape == (3, 8, 10, 3) and bool(torch.isfinite(pcs2).all())
    lab1 = (d["s3_rgbs"][..., 3] == 1) & (d["s3_depth"] > 0)
    near_surface = (z2.cpu()[lab1][:, 1:] - d["s3_depth"][lab1][:, None]).abs()
    assert float(near_surface.max()) <= 0.1 + 1e-6


def test_render_rays_helpers_kernels_vs_tensor_expressions():
    """render_rays.occupancy_activation / occupancy_to_termination / render: CUDA tensors outside autograd run as kernels
    (oo_occupancy_activation, oo_termination, oo_render_sum); the same functions on CPU tensors are the reference's tensor
    expressions (render_rays.py:6-63).  Both call forms of loss.py (:31-35,82) are covered; a tensor that requires grad
    keeps the differentiable expression."""
    from openobj_b200 import render_rays as R
    g = torch.Generator().manual_seed(11)
    alpha = torch.randn(6, 50, 10, generator=g) * 3
    z = torch.sort(torch.rand(6, 50, 10, generator=g) * 4 + 0.3, dim=-1).values
    color = torch.rand(6, 50, 10, 3, generator=g)
    feat = torch.randn(6, 50, 10, 64, generator=g)
    occ_c, occ_g = R.occupancy_activation(alpha), R.occupancy_activation(alpha.to(DEV))
    torch.testing.assert_close(occ_g.cpu(), occ_c, rtol=1e-6, atol=1e-7)
    T_c, T_g = R.occupancy_to_termination(occ_c, is_batch=True), R.occupancy_to_termination(occ_c.to(DEV), is_batch=True)
    torch.testing.assert_close(T_g.cpu(), T_c, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(R.render(T_c.to(DEV), z.to(DEV)).cpu(), R.render(T_c, z), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(R.render(T_c.to(DEV)[..., None], color.to(DEV), dim=-2).cpu(), R.render(T_c[..., None], color, dim=-2),
                               rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(R.render(T_c.to(DEV)[..., None], feat.to(DEV), dim=-2).cpu(), R.render(T_c[..., None], feat, dim=-2),
                               rtol=1e-5, atol=1e-5)
    a = alpha.to(DEV).requires_grad_(True)
    d = R.render(R.occupancy_to_termination(R.occupancy_activation(a)), z.to(DEV)).sum()
    d.backward()
    assert a.grad is not None and bool(torch.isfinite(a.grad).all())
