"""Mirror of objnerf/cfg.py: the same JSON schema -> the same flat attribute names (cfg.py:8-114)."""
import json
import os

import numpy as np

_FIELDS = [  # (attribute, section, key, transform)
    ("start", "trainer", "start", None), ("stride", "trainer", "stride", None), ("do_bg", "trainer", "do_bg", bool),
    ("training_device", "trainer", "train_device", None), ("data_device", "trainer", "data_device", None),
    ("max_n_models", "trainer", "n_models", None), ("live_mode", "dataset", "live", bool),
    ("keep_live_time", "dataset", "keep_alive", None), ("imap_mode", "trainer", "imap_mode", None),
    ("training_strategy", "trainer", "training_strategy", None), ("dataset_format", "dataset", "format", None),
    ("dataset_dir", "dataset", "path", None), ("mh", "camera", "mh", None), ("mw", "camera", "mw", None),
    ("height", "camera", "h", None), ("width", "camera", "w", None), ("win_size", "model", "window_size", None),
    ("n_iter_per_frame", "render", "iters_per_frame", None), ("n_per_optim", "render", "n_per_optim", None),
    ("win_size_bg", "model", "window_size_bg", None), ("n_per_optim_bg", "render", "n_per_optim_bg", None),
    ("keyframe_buffer_size", "model", "keyframe_buffer_size", None), ("obj_scale", "model", "obj_scale", None),
    ("bg_scale", "model", "bg_scale", None), ("hidden_feature_size", "model", "hidden_feature_size", None),
    ("hidden_feature_size_bg", "model", "hidden_feature_size_bg", None),
    ("clip_point_feature_size", "model", "clip_point_feature_size", None),
    ("n_bins_cam2surface", "render", "n_bins_cam2surface", None),
    ("n_bins_cam2surface_bg", "render", "n_bins_cam2surface_bg", None), ("n_bins", "render", "n_bins", None),
    ("n_unidir_funcs", "model", "n_unidir_funcs", None), ("surface_eps", "model", "surface_eps", None),
    ("stop_eps", "model", "other_eps", None), ("if_vis", "vis", "if_vis", bool), ("if_ckpt", "vis", "if_ckpt", bool),
    ("if_render", "vis", "if_render", bool), ("if_obj", "vis", "if_obj", bool), ("save_pcd", "vis", "save_pcd", bool),
    ("save_mesh", "vis", "save_mesh", bool), ("vis_device", "vis", "vis_device", None), ("bg_id", "vis", "bg_id", None),
    ("n_vis_iter", "vis", "n_vis_iter", None), ("eps_fine_vis", "vis", "eps_fine_vis", None),
    ("n_bins_fine_vis", "vis", "n_bins_fine_vis", None), ("live_voxel_size", "vis", "live_voxel_size", None),
    ("grid_dim", "vis", "grid_dim", None),
]


class Config:
    def __init__(self, config_file=None, config=None):
        if config is None:
            with open(config_file) as f:
                config = json.load(f)
        for attr, sec, key, fn in _FIELDS:
            v = config[sec][key]
            setattr(self, attr, fn(v) if fn else v)
        self.obj_id = -1
        self.depth_scale = 1 / config["trainer"]["scale"]
        self.min_depth, self.max_depth = config["render"]["depth_range"]
        self.H = self.height - 2 * self.mh
        self.W = self.width - 2 * self.mw
        cam = config["camera"]
        if "fx" in cam:
            self.fx, self.fy = cam["fx"], cam["fy"]
            self.cx, self.cy = cam["cx"] - self.mw, cam["cy"] - self.mh
        else:  # ScanNet: intrinsics come from the dataset directory (cfg.py:45-50)
            K = np.loadtxt(os.path.join(self.dataset_dir, "intrinsic/intrinsic_depth.txt"))
            self.fx, self.fy, self.cx, self.cy = K[0, 0], K[1, 1], K[0, 2] - self.mw, K[1, 2] - self.mh
        if "distortion" in cam:
            self.distortion_array = np.array(cam["distortion"])
        elif "k1" in cam:
            self.distortion_array = np.array([cam[k] for k in ("k1", "k2", "p1", "p2", "k3", "k4", "k5", "k6")])
        else:
            self.distortion_array = None
        self.part_mode = bool(config["trainer"].get("part_mode", 0))
        if "part_mode" in config["trainer"]:
            self.part_down = config["trainer"]["part_down"]
        self.n_samples_per_frame = self.n_per_optim // self.win_size
        self.n_samples_per_frame_bg = self.n_per_optim_bg // self.win_size_bg
        # float on purpose: 25 / 10 = 2.5, so "is keyframe" fires on every 5th appended frame
        self.keyframe_step = config["model"]["keyframe_step"] / self.stride
        self.keyframe_step_bg = config["model"]["keyframe_step_bg"] / self.stride
        self.learning_rate = config["optimizer"]["args"]["lr"]
        self.weight_decay = config["optimizer"]["args"]["weight_decay"]


ROOM0 = {
    "dataset": {"live": 0, "path": "", "format": "Replica", "keep_alive": 20},
    "optimizer": {"args": {"lr": 0.001, "weight_decay": 0.013, "pose_lr": 0.001}},
    "trainer": {"part_mode": 1, "part_down": 5, "imap_mode": 0, "start": 0, "stride": 10, "do_bg": 1, "n_models": 100,
                "train_device": "cuda:0", "data_device": "cuda:0", "training_strategy": "vmap", "epochs": 1000000,
                "scale": 1000.0},
    "render": {"depth_range": [0.0, 8.0], "n_bins": 9, "n_bins_cam2surface": 1, "n_bins_cam2surface_bg": 5,
               "iters_per_frame": 100, "n_per_optim": 120, "n_per_optim_bg": 1200},
    "model": {"n_unidir_funcs": 5, "obj_scale": 2.0, "bg_scale": 5.0, "color_scaling": 5.0, "opacity_scaling": 10.0,
              "gt_scene": 1, "surface_eps": 0.1, "other_eps": 0.05, "keyframe_buffer_size": 20, "keyframe_step": 25,
              "keyframe_step_bg": 50, "window_size": 5, "window_size_bg": 10, "hidden_layers_block": 1,
              "hidden_feature_size": 32, "hidden_feature_size_bg": 128, "clip_point_feature_size": 512},
    "camera": {"w": 1200, "h": 680, "fx": 600.0, "fy": 600.0, "cx": 599.5, "cy": 339.5, "mw": 0, "mh": 0},
    "vis": {"if_vis": 0, "if_ckpt": 1, "if_render": 0, "if_obj": 0, "save_pcd": 0, "save_mesh": 1, "vis_device": "cuda:0",
            "bg_id": [0, 2, 3], "n_vis_iter": 9999, "eps_fine_vis": 0.1, "n_bins_fine_vis": 10, "im_vis_reduce": 10,
            "grid_dim": 128, "live_vis": 1, "live_voxel_size": 0.005},
}


def room0_config(**camera_overrides):
    """The hyper-parameters of configs/Replica/room_0.json (values are data), optionally with another frame size."""
    import copy
    c = copy.deepcopy(ROOM0)
    c["camera"].update(camera_overrides)
    return Config(config=c)
