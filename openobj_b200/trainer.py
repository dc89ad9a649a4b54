"""Mirror of objnerf/trainer.py:11-44: `Trainer(cfg)` owns one UniDirsEmbed + one OccupancyMap for an object."""
import torch

from . import embedding, model


class Trainer:
    def __init__(self, cfg):
        self.obj_id = cfg.obj_id
        self.device = cfg.training_device
        self.hidden_feature_size = cfg.hidden_feature_size
        self.clip_point_feature_size = cfg.clip_point_feature_size
        self.obj_scale = cfg.obj_scale
        self.n_unidir_funcs = cfg.n_unidir_funcs
        self.emb_size1 = 21 * (3 + 1) + 3
        self.emb_size2 = 21 * (5 + 1) + 3 - self.emb_size1
        self.load_network()
        self.bound_extent = 0.995 if self.obj_id == 0 else 0.9
        self.W_vis, self.H_vis = cfg.W, cfg.H
        self.T_WC_gt = self.dirs_C_gt = self.input_pcs = None

    def load_network(self):
        self.fc_occ_map = model.OccupancyMap(self.emb_size1, self.emb_size2, hidden_size=self.hidden_feature_size,
                                             clip_size=self.clip_point_feature_size)
        self.fc_occ_map.apply(model.init_weights).to(self.device)
        self.pe = embedding.UniDirsEmbed(max_deg=self.n_unidir_funcs, scale=self.obj_scale).to(self.device)

    def packed(self, device=None):
        """theta block [1, PSTRIDE] with both the MLP and the PE directions."""
        from . import layout
        theta = self.fc_occ_map.packed(device)
        layout.views(theta)[18].copy_(self.pe.B_layer.weight.detach()[None])
        return theta

    def packed_cached(self, device=None):
        """packed(), re-used while no parameter has been written (tensor version counters): full-frame evaluation renders the
        same trained weights for every pose."""
        ps = list(self.fc_occ_map.parameters()) + [self.pe.B_layer.weight]
        key = (str(device), tuple((p.data_ptr(), p._version) for p in ps))
        if getattr(self, "_packed_key", None) != key:
            self._packed_key, self._packed = key, self.packed(device)
        return self._packed

    # ---- evaluation at free points / on the meshing grid (trainer.py:46-128) ----------------
    def eval_points(self, points, chunk_size=300000, want_clip=True):
        """trainer.py:104-128: occupancy, colour and part feature at `points` [n,3]; None when nothing is occupied.
        Hidden width 32 (every object model): ONE launch of the fused forward tile for the whole query (chunk_size is kept
        for the signature; the reference needs it to bound its activations).  Other widths (the hidden-128 background
        model): the layer-by-layer path in chunks."""
        from . import layout, ops
        if not points.is_cuda:
            raise RuntimeError("eval_points needs CUDA points (no CPU fallback)")
        pts = points.reshape(-1, 3)
        if self.hidden_feature_size == layout.HIDDEN:
            occ, color, clip = ops.eval_points(self.packed(pts.device), pts, scale=self.obj_scale, want_clip=want_clip)
        else:
            bg = self._wide_model(pts.device)
            occ, color, clip = [], [], []
            for k in range(0, pts.shape[0], int(chunk_size)):
                a, c, f, _ = bg.forward(pts[k:k + int(chunk_size)], want_clip=want_clip)
                occ.append(ops.occupancy_activation(a[:, 0]))
                color.append(c)
                clip.append(f)
            occ, color = torch.cat(occ), torch.cat(color)
            clip = torch.cat(clip) if want_clip else None
        if float(occ.max()) == 0:
            print("no occ")
            return None
        if self.hidden_feature_size == layout.HIDDEN and not want_clip:
            ops.check_tc(pts.device)
        return occ, color, clip

    def sample_points_bbox(self, bbox, do_eval=True, draws=None):
        """trainer.py:130-198: rays of `self.T_WC_gt` [B,4,4] / `self.dirs_C_gt` [B,n,3] (or [B,3]) against the oriented box
        `bbox` (.R, .center, .extent); stratified depths between entry and exit (+0.2), bin midpoints, sample points.
        Sets .dirs_W .origins .z_vals_cat .z_vals .input_pcs for the rays that hit and returns (hit_mask, near, far), or
        (None, None, None) when at most one ray hits.  `draws` replaces the torch.rand of stratified_bins."""
        import numpy as np
        from . import utils
        dev = self.dirs_C_gt.device
        T_wc_all = self.T_WC_gt.to(dev).float()
        origins, dirs_W = utils.origin_dirs_W(T_wc_all, self.dirs_C_gt)
        B = T_wc_all.shape[0]
        dirs_W = dirs_W.reshape(-1, 3)
        origins = origins[:, None, :].expand(B, dirs_W.shape[0] // B, 3).reshape(-1, 3)
        n_bins = 150 if do_eval else (60 if self.obj_id == 0 else 20)
        T_WO = torch.eye(4)                                              # 4x4 host algebra, as the reference (trainer.py:153-160)
        T_WO[:3, :3] = torch.as_tensor(np.asarray(bbox.R), dtype=torch.float32)
        T_WO[:3, 3] = torch.as_tensor(np.asarray(bbox.center), dtype=torch.float32)
        T_OC = torch.inverse(T_WO) @ self.T_WC_gt[0].cpu().to(T_WO.dtype)
        T_OC_gt = T_OC.float().unsqueeze(0).repeat_interleave(B, dim=0).to(dev)
        origins_r, dirs_r = utils.origin_dirs_W(T_OC_gt, self.dirs_C_gt)
        dirs_r = dirs_r.reshape(-1, 3)
        origins_r = origins_r[:, None, :].expand(B, dirs_r.shape[0] // B, 3).reshape(-1, 3)
        half = np.asarray(bbox.extent, dtype=np.float64).reshape(-1).astype(np.float32) / 2.0
        near, far, hit = utils.ray_box_intersection(origins_r, dirs_r, -half, half)
        if int(hit.sum()) <= 1:
            return None, None, None
        near_h = torch.clamp(near[hit], min=0.0)                         # trainer.py:168 (selection / clamp of the bounds)
        far_h = far[hit] + 0.2                                           # :169 if cam inside bound extend ray a bit
        n_rays = int(near_h.shape[0])
        self.dirs_W, self.origins = dirs_W[hit], origins[hit]
        self.z_vals_cat = utils.stratified_bins(near_h, far_h, n_bins, n_rays, device=dev, draws=draws)
        self.input_pcs, self.z_vals = utils.ray_points(self.origins, self.dirs_W, self.z_vals_cat, midpoints=True)
        return hit, near_h, far_h

    def _wide_model(self, device):
        from .background import BackgroundModel
        bg = BackgroundModel(hidden=self.hidden_feature_size, device=device, scale=self.obj_scale)
        bg.load([p.detach() for p in self.fc_occ_map.parameters()] + [self.pe.B_layer.weight.detach()])
        return bg

    def grid_transform(self, bound):
        """scene_scale and the 4x4 grid -> world transform of trainer.py:50-59 (float32, as the reference builds them)."""
        import numpy as np
        scene_scale = np.asarray(bound.extent) / (2.0 * self.bound_extent)
        tr = np.eye(4, dtype=np.float32)
        tr[:3, 3] = np.asarray(bound.center)
        tr[:3, :3] = np.asarray(bound.R)
        return torch.from_numpy(scene_scale).float(), torch.from_numpy(tr)

    def eval_grid(self, bound, obj_center=None, grid_dim=256, want_clip=False, device=None):
        """The compute part of Trainer.meshing (trainer.py:46-69): grid_dim^3 query points inside the oriented box `bound`
        (.R, .center, .extent), the model evaluated at all of them.  Returns (grid_pc [n,3], occ [n], color [n,3], clip) or
        None when nothing is occupied.  The reference discards the grid's part features, hence want_clip=False."""
        from . import ops
        dev = torch.device(device or self.device)
        scale, tr = self.grid_transform(bound)
        grid_pc = ops.make_grid(grid_dim, scale, tr, center=obj_center, device=dev)
        ret = self.eval_points(grid_pc, want_clip=want_clip)
        if ret is None:
            return None
        return (grid_pc,) + tuple(ret)

    def meshing(self, bound, obj_center, grid_dim=256, save_pcd=True, save_mesh=True, if_color=False, if_part=False):
        """trainer.py:46-102.  The grid evaluation runs on the GPU (eval_grid); turning the occupancy volume into an open3d
        point cloud / a marching-cubes mesh needs open3d, skimage and trimesh, which are host-side geometry libraries
        outside the accelerated path: without them use eval_grid() directly."""
        ret = self.eval_grid(bound, obj_center, grid_dim)
        if ret is None:
            return None, None
        grid_pc, occ, colors, _ = ret
        try:
            import open3d as o3d
        except ImportError as e:
            raise ImportError("Trainer.meshing: open3d is not installed; Trainer.eval_grid() returns the occupancy volume "
                              "and colours this step would convert") from e
        if save_pcd:
            keep = occ > 0.5
            pcd = o3d.geometry.PointCloud()
            pcd.points = o3d.utility.Vector3dVector(grid_pc[keep].cpu().numpy())
            pcd.colors = o3d.utility.Vector3dVector(colors[keep].cpu().numpy())
            pcd.voxel_down_sample(voxel_size=0.02)
            return pcd, None, None
        raise NotImplementedError("marching-cubes mesh export (skimage + trimesh) is outside the accelerated path; "
                                  "Trainer.eval_grid() returns the occupancy volume it starts from")
