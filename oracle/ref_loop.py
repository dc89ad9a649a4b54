"""The reference's own training loop, driven on synthetic frames: UNMODIFIED reference modules (oracle/ref_harness.py) +
a restatement of the parts of objnerf/train.py that cannot be imported (train.py is a script with top-level imports of CLIP /
SBERT): frame ingestion train.py:164-276, sampling + stacking :300-388, the iteration body :394-474, write-back :478-485.

MEASUREMENT INFRASTRUCTURE ONLY (bench.py's `--impl reference` arm and `cpu_baseline` / `gpu_eager_reference` legs); nothing
under openobj_b200/ imports it.  Every arithmetic operation below is executed by the reference's code or by torch, as in
the reference; this file only sequences the calls like train.py does.
"""
from __future__ import annotations

import time

import torch

import ref_harness


class ReferenceLoop:
    def __init__(self, n_obj, device="cpu", part_mode=True, W=1200, H=680, seed=0, threads=None, max_n_models=None):
        from functorch import vmap
        self.vmap = vmap
        self.m = ref_harness.load()
        self.device = torch.device(device)
        if threads:
            torch.set_num_threads(int(threads))
        cfg = ref_harness.make_cfg(device=str(device))
        cfg.part_mode = bool(part_mode)
        cfg.do_bg = False
        cfg.W, cfg.H = W, H
        cfg.max_n_models = max_n_models or max(n_obj, cfg.max_n_models)
        self.cfg = cfg
        torch.manual_seed(seed)
        self.cam = self.m["vmap"].cameraInfo(cfg)
        self.obj_dict, self.fc_models, self.pe_models = {}, [], []
        self.optimiser = torch.optim.AdamW([torch.autograd.Variable(torch.tensor(0))], betas=(0.9, 0.999),
                                           weight_decay=cfg.weight_decay)                          # train.py:78
        self.global_partfeat = None
        self.update_vmap_model = False
        self.frames = 0

    # ---- train.py:164-276 ------------------------------------------------------------------------------------------------
    def add_frame(self, sample):
        cfg, m = self.cfg, self.m
        rgb = sample["image"].to(cfg.data_device)
        depth = sample["depth"].to(cfg.data_device)
        twc = sample["T"].to(cfg.data_device)
        bbox_dict = sample["bbox_dict"]
        live_frame_id = sample.get("frame_id", self.frames)
        if cfg.part_mode:
            part_feat = sample["part_feat"].to(cfg.data_device)
            self.global_partfeat = part_feat.unsqueeze(0) if self.global_partfeat is None else \
                torch.cat((self.global_partfeat, part_feat.unsqueeze(0)), dim=0)
        inst = sample["obj"].to(cfg.data_device)
        for obj_id in torch.unique(inst):
            if obj_id == -1:
                continue
            obj_id = int(obj_id)
            if obj_id == 0:
                continue               # synthetic frames carry no background object (do_bg off in this harness)
            state = torch.zeros_like(inst, dtype=torch.uint8, device=cfg.data_device)
            state[inst == obj_id] = 1
            state[inst == -1] = 2
            bbox = bbox_dict[obj_id]
            if obj_id in self.obj_dict:
                self.obj_dict[obj_id].append_keyframe(rgb, depth, state, bbox, twc, live_frame_id)
            else:
                if len(self.obj_dict) >= cfg.max_n_models:
                    continue
                o = m["vmap"].sceneObject(cfg, obj_id, rgb, depth, state, bbox, twc, live_frame_id)
                self.obj_dict[obj_id] = o
                self.optimiser.add_param_group({"params": o.trainer.fc_occ_map.parameters(), "lr": cfg.learning_rate,
                                                "weight_decay": cfg.weight_decay})
                self.optimiser.add_param_group({"params": o.trainer.pe.parameters(), "lr": cfg.learning_rate,
                                                "weight_decay": cfg.weight_decay})
                self.update_vmap_model = True
                self.fc_models.append(o.trainer.fc_occ_map)
                self.pe_models.append(o.trainer.pe)
        if self.update_vmap_model:
            self.fc_model, self.fc_param, self.fc_buffer = m["utils"].update_vmap(self.fc_models, self.optimiser)
            self.pe_model, self.pe_param, self.pe_buffer = m["utils"].update_vmap(self.pe_models, self.optimiser)
            self.update_vmap_model = False
        self.frames += 1

    # ---- train.py:300-388 ------------------------------------------------------------------------------------------------
    def sample(self):
        cfg = self.cfg
        B = {k: [] for k in ("depth", "rgb", "dmask", "omask", "pcs", "z", "feat")}
        for obj_id, obj_k in self.obj_dict.items():
            gt_rgb, gt_depth, valid_depth_mask, obj_mask, input_pcs, sampled_z, gt_partfeat = obj_k.get_training_samples(
                cfg.n_iter_per_frame * cfg.win_size, cfg.n_samples_per_frame, self.cam.rays_dir_cache, self.global_partfeat)
            B["depth"].append(gt_depth.reshape([gt_depth.shape[0] * gt_depth.shape[1]]))
            B["rgb"].append(gt_rgb.reshape([gt_rgb.shape[0] * gt_rgb.shape[1], gt_rgb.shape[2]]))
            B["dmask"].append(valid_depth_mask)
            B["omask"].append(obj_mask)
            B["pcs"].append(input_pcs.reshape([input_pcs.shape[0] * input_pcs.shape[1], input_pcs.shape[2], input_pcs.shape[3]]))
            B["z"].append(sampled_z.reshape([sampled_z.shape[0] * sampled_z.shape[1], sampled_z.shape[2]]))
            if cfg.part_mode:
                B["feat"].append(gt_partfeat.reshape([gt_partfeat.shape[0] * gt_partfeat.shape[1], gt_partfeat.shape[2]]))
        dev = cfg.training_device
        self.Batch_N_input_pcs = torch.stack(B["pcs"]).to(dev)
        self.Batch_N_gt_depth = torch.stack(B["depth"]).to(dev)
        self.Batch_N_gt_rgb = torch.stack(B["rgb"]).to(dev) / 255.
        self.Batch_N_depth_mask = torch.stack(B["dmask"]).to(dev)
        self.Batch_N_obj_mask = torch.stack(B["omask"]).to(dev)
        self.Batch_N_sampled_z = torch.stack(B["z"]).to(dev)
        if cfg.part_mode:
            self.Batch_N_gt_partfeat = torch.stack(B["feat"]).to(dev)

    # ---- train.py:394-474 ------------------------------------------------------------------------------------------------
    def train(self, iters):
        cfg, loss, vmap = self.cfg, self.m["loss"], self.vmap
        n = cfg.n_per_optim
        last = None
        for iter_step in range(iters):
            data_idx = slice(iter_step * n, (iter_step + 1) * n)
            batch_input_pcs = self.Batch_N_input_pcs[:, data_idx, ...]
            batch_gt_depth = self.Batch_N_gt_depth[:, data_idx, ...]
            batch_gt_rgb = self.Batch_N_gt_rgb[:, data_idx, ...]
            batch_depth_mask = self.Batch_N_depth_mask[:, data_idx, ...]
            batch_obj_mask = self.Batch_N_obj_mask[:, data_idx, ...]
            batch_sampled_z = self.Batch_N_sampled_z[:, data_idx, ...]
            batch_embedding = vmap(self.pe_model)(self.pe_param, self.pe_buffer, batch_input_pcs)
            batch_alpha, batch_color, batch_clip = vmap(self.fc_model)(self.fc_param, self.fc_buffer, batch_embedding)
            if not cfg.part_mode:
                batch_loss, _ = loss.step_batch_loss(batch_alpha, batch_color, batch_gt_depth.detach(), batch_gt_rgb.detach(),
                                                     batch_obj_mask.detach(), batch_depth_mask.detach(), batch_sampled_z.detach())
            else:
                batch_gt_partfeat = self.Batch_N_gt_partfeat[:, data_idx, ...]
                batch_loss, _ = loss.step_batch_loss(batch_alpha, batch_color, batch_gt_depth.detach(), batch_gt_rgb.detach(),
                                                     batch_obj_mask.detach(), batch_depth_mask.detach(), batch_sampled_z.detach(),
                                                     gt_partfeat=batch_gt_partfeat.detach(), pred_partfeat=batch_clip)
            batch_loss.backward()
            self.optimiser.step()
            self.optimiser.zero_grad(set_to_none=True)
            last = batch_loss
        return last

    # ---- train.py:478-485 ------------------------------------------------------------------------------------------------
    def write_back(self):
        with torch.no_grad():
            for model_id, (obj_id, obj_k) in enumerate(self.obj_dict.items()):
                for i, param in enumerate(obj_k.trainer.fc_occ_map.parameters()):
                    param.copy_(self.fc_param[i][model_id])
                for i, param in enumerate(obj_k.trainer.pe.parameters()):
                    param.copy_(self.pe_param[i][model_id])

    def frame(self, sample, iters):
        """One frame of train.py's main loop; returns the last step's loss (a tensor)."""
        self.add_frame(sample)
        self.sample()
        last = self.train(iters)
        self.write_back()
        return last


def timed_run(n_obj, device, part_mode, steps, warmup, synth, fill_frames=2, iters_per_frame=100, threads=None, seed=0,
              max_seconds=150.0):
    """warmup + `steps` optimisation steps of the reference loop, frames of `iters_per_frame` steps (per-frame ingestion and
    sampling inside the timed region, like bench.py's own arm).  Returns a dict with rays/s and what was run."""
    loop = ReferenceLoop(n_obj, device=device, part_mode=part_mode, W=synth.W, H=synth.H, seed=seed, threads=threads,
                         max_n_models=n_obj)
    cuda = torch.device(device).type == "cuda"
    sync = (lambda: torch.cuda.synchronize()) if cuda else (lambda: None)
    f = 0
    for _ in range(fill_frames):
        loop.add_frame(synth.frame(f)); f += 1
    done = 0
    while done < warmup:
        it = min(iters_per_frame, warmup - done)
        loop.frame(synth.frame(f), it); f += 1
        done += it
    sync()
    t0 = time.perf_counter()
    done, last = 0, None
    while done < steps:
        it = min(iters_per_frame, steps - done)
        last = loop.frame(synth.frame(f), it); f += 1
        done += it
        if time.perf_counter() - t0 > max_seconds:
            break
    loss_value = float(last.detach()) if last is not None else float("nan")      # D2H read of the result
    sync()
    dt = time.perf_counter() - t0
    n = len(loop.obj_dict)
    return dict(rays_per_s=n * loop.cfg.n_per_optim * done / dt, steps_done=done, seconds=dt, n_obj=n,
                ms_per_step=1e3 * dt / max(done, 1), cores=torch.get_num_threads(), loss=loss_value)
