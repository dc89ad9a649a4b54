#!/usr/bin/env python
"""Benchmark of the accelerated OpenObj hot path (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--objects 60] [--part 1]

A *step* is one optimisation iteration of the whole ensemble (N objects x 120 rays x 10 samples: encode, MLP,
compositing, losses, backward, AdamW -- objnerf/train.py:394-474).  Every 100 steps a new synthetic RGB-D + mask
(+ part-feature) frame of Replica shape (1200x680) is appended to the keyframe rings and all objects are re-sampled
(train.py:164-388); that per-frame work is inside the timed region (sampling amortised per frame, as BASELINE.json's
metric says).  Workload = BASELINE.json configs[1]: "Replica room_0 shape, ~60 objects, 1 B200"; with --gpus G each
rank holds 60 objects (weak scaling, objects sharded by ensemble index, no gradient collectives).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  NCCL prints its version banner with printf on file descriptor 1 (seen on the
# 2-GPU box even with NCCL_DEBUG_FILE set), so descriptor 1 is pointed at stderr for the whole process and the JSON
# line goes through a duplicate of the original stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

ITERS = 100           # iters_per_frame (room_0.json:34)
R = 120               # n_per_optim (room_0.json:35)
S = 10
MAC_PER_POINT = 63 + 2784 + 1024 + 3808 + 1024 + 32 + 2368 + 96      # PE, in, mid1, cat, mid2, alpha, color_linear, out_color
MAC_CLIP_POINT = 2368                                                   # clip_linear (per point)
MAC_CLIP_RAY = 16384                                                    # out_clip applied once per ray


def flop_per_ray(part):
    mac = S * (MAC_PER_POINT + (MAC_CLIP_POINT if part else 0)) + (MAC_CLIP_RAY if part else 0)
    return 2 * 3 * mac        # FLOP = 2 MAC; forward + dX + dW


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe) through NVML from a
    background thread (a 10 ms period: the timed region of a default run lasts well under a second, less than the
    start-up time of an `nvidia-smi -lms` child process).  Falls back to one-shot nvidia-smi queries when NVML is
    not importable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index=0, period_s=0.01):
        self.index, self.period, self.rows, self.max_mhz = index, period_s, [], None
        self.stop = threading.Event()
        self.thread = None
        self.err = None

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if self.index < len(ids) and ids[self.index].strip().isdigit():
                return int(ids[self.index])
        return self.index

    def _loop_nvml(self, nv, h):
        while not self.stop.is_set():
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksEventReasons(h),
                                  nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
            except Exception as e:          # noqa: BLE001 -- keep sampling, remember why a sample was lost
                self.err = repr(e)
            self.stop.wait(self.period)

    def _loop_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self._visible_index()), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                r = [x.strip() for x in out.strip().split(",")]
                mask = 0
                for (name, bit), col in zip(self.REASONS, (3, 4, 5, 6)):
                    if len(r) > col and r[col].lower().startswith("active"):
                        mask |= bit
                self.rows.append((int(r[0]), mask, float(r[2])))
                self.max_mhz = int(r[1])
            except Exception as e:          # noqa: BLE001
                self.err = repr(e)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._loop_nvml, args=(nv, h), daemon=True)
        except Exception as e:              # noqa: BLE001
            self.err = repr(e)
            self.thread = threading.Thread(target=self._loop_smi, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no clock samples: %s" % self.err]}
        mask = 0
        for r in self.rows:
            mask |= r[1]
        reasons = [name for name, bit in self.REASONS if mask & bit]
        busy = [v for v in sm if v >= 0.6 * sm[-1]] or sm
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                "power_w_max": max(r[2] for r in self.rows)}


def cuda_timer():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/openobj_oracle.py == the reference's algorithm in torch-CPU) on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_run(n_obj, part, steps, warmup, time_budget_s):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import openobj_oracle as oc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    fc, B = oc.init_params(n_obj, generator=g)
    P = [p.clone() for p in fc] + [B.clone()]
    M = [torch.zeros_like(p) for p in P]
    V = [torch.zeros_like(p) for p in P]
    z = torch.sort(0.5 + 3.0 * torch.rand(n_obj, R, S, generator=g), dim=-1).values
    o = torch.randn(n_obj, R, 1, 3, generator=g) * 0.2
    d = torch.nn.functional.normalize(torch.randn(n_obj, R, 1, 3, generator=g), dim=-1)
    pcs = o + d * z[..., None]
    gt_depth = z[..., 6].clone()
    rgb = torch.rand(n_obj, R, 3, generator=g)
    labels = torch.randint(0, 3, (n_obj, R), generator=g, dtype=torch.uint8)
    labels[:, 0], labels[:, 1] = 1, 0
    gt_feat = torch.randn(n_obj, R, 512, generator=g) if part else None

    def one_step(t):
        terms, grads = oc.train_step_grads(P[:18], P[18], pcs, z, gt_depth, rgb, labels, gt_feat)
        for i, gr in enumerate(grads):
            if gr is not None:
                oc.adamw_step(P[i], gr, M[i], V[i], t)
        return float(terms.total.detach())

    for w in range(warmup):
        one_step(w + 1)
    t0 = time.perf_counter()
    done = 0
    for s in range(steps):
        one_step(warmup + s + 1)
        done += 1
        if time.perf_counter() - t0 > time_budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(rays_per_s=n_obj * R * done / dt, steps_done=done, seconds=dt, cores=cores, ms_per_step=1e3 * dt / done)


def reference_loop_run(n_obj, part, steps, warmup, W, H, device="cpu", max_seconds=150.0):
    """The reference's own loop (UNMODIFIED reference modules from oracle/_ref or /root/reference through
    oracle/ref_loop.py: frame ingestion, per-object sampling, vmap forward, step_batch_loss, backward, AdamW, write-back) on
    synthetic frames of the bench's shape.  Returns None when the reference modules are not available (then the oracle port
    is the CPU arm)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness
    if not ref_harness.available():
        return None
    import warnings
    warnings.filterwarnings("ignore")
    import ref_loop
    from openobj_b200.synthetic import SyntheticScene
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    synth = SyntheticScene(n_obj, W=W, H=H, part_mode=bool(part), seed=0, n_distinct=2)
    r = ref_loop.timed_run(n_obj, device, bool(part), steps, warmup, synth, fill_frames=2, iters_per_frame=ITERS,
                           max_seconds=max_seconds)
    r["kind"] = "reference"
    return r


def config_shape(args):
    """(W, H, part_mode, total objects or None = args.objects per GPU) of BASELINE.json's configs."""
    if args.config == 3:
        return 1200, 680, 1, 100
    if args.config == 4:
        return 640, 480, 0, 200
    return 1200, 680, int(args.part), None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    W, H, part, total = config_shape(args)
    n_obj = total or args.objects
    warm = min(args.warmup, 3)
    r = reference_loop_run(n_obj, part, args.steps, warm, W, H, device=args.device)
    if r is None:
        n_port = 8
        r = cpu_reference_run(n_port, bool(part), args.steps, warm, time_budget_s=150.0)
        r.update(kind="port", n_obj=n_port)
        sample = ("oracle port (torch-CPU restatement of the reference step: vmap forward, step_batch_loss, autograd backward, "
                  "AdamW; no sampling) on %d of the %d objects per step, %d steps (time-bounded); the reference modules "
                  "(oracle/_ref) were not found" % (n_port, n_obj, r["steps_done"]))
    else:
        sample = ("UNMODIFIED reference modules (objnerf/{vmap,utils,trainer,model,embedding,render_rays,loss,cfg}.py via "
                  "oracle/ref_loop.py) on %s: all %d objects, %dx%d frames, part_mode=%d; per frame: ingestion into the "
                  "per-object keyframe rings + per-object sampling + %d steps of vmap forward / step_batch_loss / backward / "
                  "AdamW + write-back; %d steps in %.1f s" % (args.device, r["n_obj"], W, H, part, min(ITERS, args.steps),
                                                              r["steps_done"], r["seconds"]))
    line = {
        "metric": "training rays/sec for N-object ensemble", "value": r["rays_per_s"], "unit": "rays/s", "n_gpus": args.gpus,
        "steps": r["steps_done"], "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak" if total is None else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": "BASELINE configs[%d]: %dx%d frames, %d objects x 120 rays x 10 samples per step, part_mode=%d, "
                               "%d steps per frame; per-frame ingestion + sampling inside the timed region"
                               % (args.config - 1, W, H, n_obj, part, ITERS),
                   "device": args.device},
        "cpu_baseline": {"value": r["rays_per_s"], "unit": "rays/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["rays_per_s"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from openobj_b200 import _lib, cfg as C, dist as D, ops
    from openobj_b200.ensemble import Ensemble
    from openobj_b200.scene import Scene
    from openobj_b200.synthetic import SyntheticScene

    rank, world, local = D.init_from_env()
    if not torch.cuda.is_available():
        raise _lib.OOError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    W_, H_, part_, total_ = config_shape(args)
    cfg = C.room0_config()
    cfg.training_device = cfg.data_device = str(dev)
    cfg.part_mode = bool(part_)
    cfg.do_bg = bool(args.bg)               # --bg 1: the reference's real loop (room_0.json do_bg: 1) -- the hidden-128 background
                                            # model trains beside the ensemble every step (train.py:447-463); informative
    if (W_, H_) != (cfg.W, cfg.H):              # ScanNet shape (configs/ScanNet/scene0011_01.json:56-57)
        cfg.W, cfg.H = W_, H_
        cfg.fx = cfg.fy = 0.5 * W_
        cfg.cx, cfg.cy = 0.5 * W_ - 0.5, 0.5 * H_ - 0.5
    n_total = total_ if total_ is not None else args.objects * world
    cfg.max_n_models = n_total                  # trainer.n_models: the global cap on objects (train.py:231-233)
    steps, warmup = args.steps, max(args.warmup, 3)
    frames_w = (warmup + ITERS - 1) // ITERS
    frames_t = (steps + ITERS - 1) // ITERS
    fill = args.fill_frames
    n_frames_total = fill + 2 * (frames_w + frames_t) + 12
    # every rank sees the same frames; object ids are spread so that rank r owns ids with (index % world == r)
    synth = SyntheticScene(n_total, W=cfg.W, H=cfg.H, part_mode=cfg.part_mode, seed=0, pin=True, n_distinct=2, with_bg=cfg.do_bg)
    scene = Scene(cfg, rank=rank, world=world, seed=1234, max_frames=n_frames_total, flag_allreduce=D.make_flag_allreduce())

    def iters_of(frame_idx, n_frames, total):
        return min(ITERS, total - frame_idx * ITERS)

    # ---- fill the keyframe rings (untimed), then move frame payloads to the device for the device-resident pass
    f = 0
    for _ in range(fill):
        scene.add_frame(synth.frame(f)); f += 1
    scene.sample()
    scene.train(iters=3)
    torch.cuda.synchronize()

    def to_dev(s):
        return {k: (v.to(dev) if torch.is_tensor(v) and k != "T" else v) for k, v in s.items()}

    n_obj = len(scene.obj_dict)
    lt = torch.zeros(ITERS, n_obj, 4, device=dev)
    launches = {"n": 0}

    def run_frames(n_frames, total_steps, host, frames=None):
        """host=False: `frames` are device-resident frame dicts prepared before the timed region (`value`).
        host=True: frames come from pinned host memory; the H2D copy of frame j+1 is issued on Scene's copy stream while
        frame j trains (what pin_memory + non_blocking gives the reference's DataLoader), and the last step's loss is
        read back every frame (`e2e`)."""
        nonlocal f
        staged = scene.stage_frame(synth.frame(f)) if host else None
        for j in range(n_frames):
            it = iters_of(j, n_frames, total_steps)
            t0 = time.perf_counter()
            if host:
                cur = staged
                scene.add_frame(cur)
                if j + 1 < n_frames:
                    staged = scene.stage_frame(synth.frame(f + 1))
            else:
                scene.add_frame(frames[j])
            t1 = time.perf_counter()
            scene.sample()
            t2 = time.perf_counter()
            scene.train(iters=it, loss_terms=lt)
            if args.trace:
                t3 = time.perf_counter()
                print("[trace] rank %d frame %d host ms: add_frame %.3f sample %.3f train(enqueue %d steps) %.3f"
                      % (rank, j, 1e3 * (t1 - t0), 1e3 * (t2 - t1), it, 1e3 * (t3 - t2)), file=sys.stderr)
            if host:
                _ = lt[it - 1].sum().item()        # D2H read of the step result inside the timed region
            f += 1
            # per step: k_train + k_update; per frame: k_gram (part features on), label counts + adam schedule, frame store,
            # sampler (two passes)
            launches["n"] += 2 * it + (1 if args.part else 0) + 2 + 1 + 2
            if args.bg and scene.bg is not None:
                launches["n"] += 47 * it + 2          # the background step (41-47 launches) + its sampling passes

    def resident(n_frames):
        """device copies of the next n_frames frames, made BEFORE the timed region"""
        fr = [to_dev(synth.frame(f + j)) for j in range(n_frames)]
        torch.cuda.synchronize()
        return fr

    # ---- device-resident pass: `value`
    run_frames(frames_w, warmup, host=False, frames=resident(frames_w))
    dev_frames = resident(frames_t)
    torch.cuda.synchronize(); D.barrier()
    e0, e1 = cuda_timer()
    launches["n"] = 0
    marks = []
    if args.trace:           # device-side timeline of the per-frame work (events on the stream, read after the pass)
        def mark(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))
        scene.ens.mark = mark
    # the clock sampler (NVML initialisation, a thread start: milliseconds, different on every rank) is brought up BEFORE the
    # barrier: whatever sits between the barrier and e0 becomes start skew between the ranks, and the first rank to reach the
    # per-frame all-reduce then waits for the last one inside its own timed region
    clk = ClockSampler(local)
    clk.__enter__()
    time.sleep(0.02)
    torch.cuda.synchronize(); D.barrier()
    e0.record()
    run_frames(frames_t, steps, host=False, frames=dev_frames)
    e1.record()
    torch.cuda.synchronize()
    clk.__exit__()
    if args.trace:
        scene.ens.mark = None
        print("[trace] rank %d device ms since start: %s | end %.3f" % (
            rank, " ".join("%s %.3f" % (n, e0.elapsed_time(ev)) for n, ev in marks[:6]), e0.elapsed_time(e1)), file=sys.stderr)
    del dev_frames
    D.barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1), dev)
    n_launch = launches["n"]
    # ---- end-to-end pass: host (pinned) frame -> H2D -> append -> sample -> train -> D2H loss
    run_frames(1, min(warmup, ITERS), host=True)
    torch.cuda.synchronize(); D.barrier()
    h2d0 = scene.h2d_bytes
    e0.record()
    run_frames(frames_t, steps, host=True)
    e1.record()
    torch.cuda.synchronize(); D.barrier()
    print("[bench] rank %d: value pass %.1f ms, e2e pass %.1f ms (this rank)" % (rank, ms, e0.elapsed_time(e1)), file=sys.stderr)
    ms_e2e = D.max_over_ranks(e0.elapsed_time(e1), dev)
    n_all = int(D.sum_over_ranks(n_obj, dev))
    rays = n_all * R * steps
    # bytes Scene.stage_frame actually moved during the timed e2e pass (a sharded rank gathers only its objects' part-feature
    # rows) + the poses; per step of this window
    h2d_per_step = (scene.h2d_bytes - h2d0 + 128 * frames_t) / steps
    d2h_per_step = 4.0 * frames_t / steps

    out = None
    if rank == 0:
        # ---- roofline of the dominant kernel (K1): CUDA events around every K1 launch of one more frame
        ens = scene.ens
        bc = scene.batch.to_c()
        ens.prepare_frame(scene.batch)
        evs = [cuda_timer() for _ in range(ITERS)]
        for it in range(ITERS):
            evs[it][0].record()
            ens.k1(bc, it, refresh_derived=(it == 0))
            evs[it][1].record()
            ens.k4(bc, it)
        torch.cuda.synchronize()
        k1_ms = sorted(a.elapsed_time(b) for a, b in evs)
        k1_avg = sum(k1_ms) / len(k1_ms)
        peak = ops.fma_peak_tflops()
        flop_launch = flop_per_ray(cfg.part_mode) * n_obj * R
        achieved = flop_launch / (k1_avg * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json"))).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        # K4 (AdamW) is the HBM-bound kernel of the step: p, m, v read+write and the gradient slabs read
        evs4 = [cuda_timer() for _ in range(ITERS)]
        for it in range(ITERS):
            ens.k1(bc, it, refresh_derived=(it == 0))
            evs4[it][0].record()
            ens.k4(bc, it)
            evs4[it][1].record()
        torch.cuda.synchronize()
        k4_avg = sum(a.elapsed_time(b) for a, b in evs4) / ITERS
        k4_bytes = n_obj * 30659 * 24 + ens.n_slots * 30659 * 4 + (n_obj * 16384 * 4 if cfg.part_mode else 0)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # per-frame HBM kernels: K2 sampling (all objects, one launch) and the keyframe append
        ev_s = [cuda_timer() for _ in range(5)]
        ev_a = [cuda_timer() for _ in range(5)]
        for j in range(5):
            s_dev = to_dev(synth.frame(f + j))
            torch.cuda.synchronize()
            # a ~1 ms spin kernel ahead of each measurement lets the host enqueue the whole call before the GPU reaches it
            # (as in the real loop, where the host runs ahead during the 100 training steps): the events then bracket
            # device time (small H2D table copies + kernels), not Python
            torch.cuda._sleep(2_000_000)
            ev_a[j][0].record(); scene.add_frame(s_dev); ev_a[j][1].record()
            torch.cuda._sleep(2_000_000)
            ev_s[j][0].record(); scene.sample(); ev_s[j][1].record()
        torch.cuda.synchronize()
        t_append = sorted(a.elapsed_time(b) for a, b in ev_a)[2]
        t_sample = sorted(a.elapsed_time(b) for a, b in ev_s)[2]
        rays_frame = n_obj * ITERS * R
        sample_bytes = rays_frame * (180 + 4)
        # shared keyframe store: the frame is read (11 B/pixel) and written (12 B/pixel) ONCE, + the part-feature map copy
        append_bytes = cfg.W * cfg.H * (11 + 12) + (synth.frame_bytes() - cfg.W * cfg.H * 11) * 2
        # ---- K3: standalone loss.step_batch_loss forward + backward at the ensemble's per-step shape [N,120,10(,512)]
        gk = torch.Generator(device=dev).manual_seed(4)
        k3_alpha = torch.randn(n_obj, R, S, generator=gk, device=dev)
        k3_color = torch.rand(n_obj, R, S, 3, generator=gk, device=dev)
        k3_z = torch.sort(0.5 + 3.0 * torch.rand(n_obj, R, S, generator=gk, device=dev), dim=-1).values
        k3_feat = torch.randn(n_obj, R, S, 512, generator=gk, device=dev) if cfg.part_mode else None
        k3_gtf = torch.randn(n_obj, R, 512, generator=gk, device=dev) if cfg.part_mode else None
        k3_rgb = torch.rand(n_obj, R, 3, generator=gk, device=dev)
        k3_lab = torch.randint(0, 3, (n_obj, R), generator=gk, device=dev, dtype=torch.uint8)
        k3_lab[:, 0], k3_lab[:, 1] = 1, 0
        Lk = _lib.lib()
        k3_ws = torch.empty(n_obj * R * Lk.oo_loss_ws_per_ray() + 8 * n_obj, device=dev)
        k3_terms, k3_loss = torch.empty(n_obj, 4, device=dev), torch.empty(1, device=dev)
        k3_flags = torch.zeros(1, dtype=torch.int32, device=dev)
        k3_da, k3_dc = torch.empty_like(k3_alpha), torch.empty_like(k3_color)
        k3_df = torch.empty_like(k3_feat) if cfg.part_mode else None
        P_ = _lib.ptr

        def k3_once():
            _lib.check(Lk.oo_loss_fwd(P_(k3_alpha), P_(k3_color), P_(k3_z), P_(k3_z[..., 5].contiguous()), P_(k3_rgb), P_(k3_lab),
                                      P_(k3_feat), P_(k3_gtf), n_obj, R, S, 512 if cfg.part_mode else 0, 5.0, 10.0, 5.0,
                                      P_(k3_terms), P_(k3_loss), P_(k3_flags), P_(k3_ws), _lib.stream()), "oo_loss_fwd")
            _lib.check(Lk.oo_loss_bwd(P_(k3_alpha), P_(k3_color), P_(k3_z), P_(k3_z[..., 5].contiguous()), P_(k3_rgb), P_(k3_lab),
                                      P_(k3_feat), P_(k3_gtf), n_obj, R, S, 512 if cfg.part_mode else 0, 5.0, 10.0, 5.0, 1.0,
                                      P_(k3_flags), P_(k3_ws), P_(k3_da), P_(k3_dc), P_(k3_df), _lib.stream()), "oo_loss_bwd")

        for _ in range(3):
            k3_once()
        ek0, ek1 = cuda_timer()
        ek0.record()
        for _ in range(10):
            k3_once()
        ek1.record()
        torch.cuda.synchronize()
        k3_ms = ek0.elapsed_time(ek1) / 10
        # algorithmic bytes per ray: forward reads alpha 40 + colour 120 + z 40 + gt 21 (+ pred 20 480 + gt feature
        # 2 048); backward re-reads the 221 B of saved inputs (+ the gt feature) and writes d_alpha 40 + d_colour 120 (+ d_pred
        # 20 480).  The backward does NOT re-read pred_feat (the forward leaves x, pred.x, pred.y in the workspace).
        k3_bytes = n_obj * R * (221 * 2 + S * 16 + ((S * 512 * 4) * 2 + 512 * 4 * 2 if cfg.part_mode else 0))
        # ---- the separate background model (outside the N-object metric): hidden 128, 1200 rays x 14 samples
        from openobj_b200.background import BackgroundModel
        gb = torch.Generator(device=dev).manual_seed(3)
        Rb, Sb = cfg.n_per_optim_bg, cfg.n_bins_cam2surface_bg + cfg.n_bins
        bgm = BackgroundModel(hidden=cfg.hidden_feature_size_bg, device=dev, rays_per_step=Rb, n_samp=Sb, scale=cfg.bg_scale)
        for v in bgm.views():
            v.copy_(torch.randn(v.shape, generator=gb, device=dev) * (1.0 / max(v.shape[-1], 1)) ** 0.5 if v.dim() == 2 else
                    torch.zeros(v.shape, device=dev))
        zb = torch.sort(0.5 + 5.0 * torch.rand(Rb, Sb, generator=gb, device=dev), dim=-1).values
        db = torch.nn.functional.normalize(torch.randn(Rb, 1, 3, generator=gb, device=dev), dim=-1)
        pb = (db * zb[..., None]).contiguous()
        rgbb = torch.randint(0, 256, (Rb, 3), generator=gb, device=dev, dtype=torch.uint8)
        labb = torch.randint(0, 3, (Rb,), generator=gb, device=dev, dtype=torch.uint8)
        tabb = scene.part_table.view(-1, 512) if cfg.part_mode else None
        rowb = torch.randint(0, tabb.shape[0], (Rb,), generator=gb, device=dev, dtype=torch.int32) if cfg.part_mode else None
        for _ in range(3):
            bgm.train_step(pb, zb, zb[:, 8].contiguous(), rgbb, labb, rowb, tabb)
        eb0, eb1 = cuda_timer()
        eb0.record()
        for _ in range(20):
            bgm.train_step(pb, zb, zb[:, 8].contiguous(), rgbb, labb, rowb, tabb)
        eb1.record()
        torch.cuda.synchronize()
        bg_ms = eb0.elapsed_time(eb1) / 20
        hb = cfg.hidden_feature_size_bg
        bg_mac_pt = 63 + hb * 87 + hb * hb + hb * (hb + 87) + hb * hb + hb + hb * (hb + 42) + 3 * hb + (
            (hb * (hb + 42) + 512 * hb) if cfg.part_mode else 0)
        bg_flop = 2 * 3 * bg_mac_pt * Rb * Sb
        # ---- the reference beside it: (i) its own CPU path on this box's host cores, a bounded sample (one frame: ingestion +
        # sampling + 20 steps, all objects); (ii) informative, the real bar of SURVEY 8d: the same unmodified modules as eager
        # PyTorch + functorch.vmap on this same GPU
        cpu = eager = None
        if not args.no_cpu:
            cpu = reference_loop_run(n_obj, cfg.part_mode, 20, 3, cfg.W, cfg.H, device="cpu", max_seconds=60.0)
            if cpu is None:
                cpu = cpu_reference_run(8, cfg.part_mode, steps=10 ** 6, warmup=1, time_budget_s=15.0)
                cpu.update(kind="port", n_obj=8)
            else:
                torch.cuda.empty_cache()
                eager = reference_loop_run(n_obj, cfg.part_mode, 40, 5, cfg.W, cfg.H, device=str(dev), max_seconds=60.0)
                torch.cuda.empty_cache()
        out = {
            "metric": "training rays/sec for N-object ensemble", "value": rays / (ms * 1e-3), "unit": "rays/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak" if total_ is None else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[%d]: %dx%d frames, %d objects in total (%d on rank 0) x 120 rays x 10 "
                                   "samples per step, part_mode=%d, 100 steps per frame; per-frame ingestion (shared keyframe "
                                   "store) + sampling of all objects inside the timed region%s"
                                   % (args.config - 1, cfg.W, cfg.H, n_all, n_obj, int(cfg.part_mode),
                                      "; do_bg=1: the background model (hidden 128, 1200 rays x 14 samples) also trains every step, "
                                      "on the last rank, its rays NOT counted in `value`" if cfg.do_bg else ""),
                       "objects_total": n_all, "objects_per_gpu": n_obj, "rays_per_step_per_object": R, "iters_per_frame": ITERS,
                       "l2": "inputs larger than L2: %.0f MB sampled batch per frame + part-feature table" %
                             (n_obj * ITERS * R * 181 / 1e6),
                       "parallelism": "objects sharded by ensemble index, rank = k mod %d; no gradient collectives" % world},
            "e2e": {"value": rays / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d_per_step,
                    "d2h_bytes_per_step": d2h_per_step, "ms_per_step": ms_e2e / steps},
            "gpu_launches": n_launch,
            "clocks": clk.summary(),
            "roofline": {"kernel": "k_train (K1 fused encode+MLP+composite+loss+backward)", "bound": "fma",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "FP32 FFMA throughput measured live by oo_fma_peak on this GPU (MEASURED_PEAKS.json "
                                        "holds only HBM and bf16 tensor peaks; K1 reproduces fp32 arithmetic: see DESIGN.md 5)",
                         "k1_ms_avg": k1_avg, "k1_ms_min": k1_ms[0], "flop_per_launch": flop_launch,
                         "arithmetic": "mma.sync m16n8k8 TF32 x3 (hi/lo error compensation, fp32-level parity): 3 tensor MACs "
                                       "per algorithmic MAC",
                         "alt": {"bound": "tensor", "peak": peaks.get("bf16_tflops", 1590.0), "unit": "TFLOP/s",
                                 "frac": achieved / peaks.get("bf16_tflops", 1590.0),
                                 "peak_source": ("MEASURED_PEAKS.json bf16_tflops (cuBLAS bf16 burst)" if "bf16_tflops" in peaks
                                                 else "fallback 1.59 PFLOP/s"),
                                 "note": "for context only: the path must reproduce fp32 arithmetic, it is not a bf16 GEMM"}},
            "roofline_sampling": {"kernel": "k_sample_main + k_sample_fix (K2, all objects; device time of Scene.sample incl. its small H2D table copies, median of 5 frames)",
                                  "bound": "hbm", "achieved": sample_bytes / (t_sample * 1e-3) / 1e9, "peak": hbm_peak,
                                  "unit": "GB/s", "frac": sample_bytes / (t_sample * 1e-3) / 1e9 / hbm_peak, "ms": t_sample,
                                  "bytes_per_launch": sample_bytes, "rays_per_s": rays_frame / (t_sample * 1e-3)},
            "roofline_append": {"kernel": "k_store_frame (one copy of the frame in the shared keyframe store) + part-feature copy + "
                                          "slot-table upload (per frame, device time of Scene.add_frame)", "bound": "hbm",
                                "achieved": append_bytes / (t_append * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": append_bytes / (t_append * 1e-3) / 1e9 / hbm_peak, "ms": t_append},
            "roofline_hbm": {"kernel": "K4 = k_clipgrad (out_clip gradient assembly) + k_adamw (slab reduction + AdamW)", "bound": "hbm",
                             "achieved": k4_bytes / (k4_avg * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": k4_bytes / (k4_avg * 1e-3) / 1e9 / hbm_peak, "k4_ms_avg": k4_avg, "bytes_per_launch": k4_bytes,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"},
        }
        out["roofline_composite"] = {"kernel": "K3 = oo_loss_fwd + oo_loss_bwd (standalone step_batch_loss at [N=%d,120,10%s]; pred_feat "
                                               "%.0f MB + its gradient > L2)" % (n_obj, ",512" if cfg.part_mode else "",
                                                                               n_obj * R * S * 2048 / 1e6),
                                     "bound": "hbm", "achieved": k3_bytes / (k3_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": k3_bytes / (k3_ms * 1e-3) / 1e9 / hbm_peak, "ms": k3_ms, "bytes_per_launch": k3_bytes}
        out["background"] = {"what": "separate background model (train.py:447-463): hidden %d, %d rays x %d samples per step, "
                                     "layer-by-layer path, GEMMs as 3xTF32 tcgen05.mma (accumulators in tensor memory, fp32-level accuracy), clip head "
                                     "applied per ray; NOT part of `value`" % (hb, Rb, Sb),
                             "ms_per_step": bg_ms, "rays_per_s": Rb / (bg_ms * 1e-3), "flop_per_step": bg_flop,
                             "achieved_tflops": bg_flop / (bg_ms * 1e-3) / 1e12, "frac_of_fma_peak": bg_flop / (bg_ms * 1e-3) / 1e12 / peak}
        if cpu is not None:
            what = ("UNMODIFIED reference modules (oracle/ref_loop.py) on the host cores: all %d objects, one frame = ingestion "
                    "into the per-object rings + per-object sampling + %d steps + write-back, %.1f s"
                    % (cpu["n_obj"], cpu["steps_done"], cpu["seconds"])) if cpu["kind"] == "reference" else (
                    "oracle port (torch-CPU restatement of the reference step, no sampling) on 8 of the %d objects, %d steps in "
                    "%.1f s" % (n_obj, cpu["steps_done"], cpu["seconds"]))
            out["cpu_baseline"] = {"value": cpu["rays_per_s"], "unit": "rays/s", "cores": cpu["cores"], "kind": cpu["kind"],
                                   "sample": what}
        if eager is not None:
            out["gpu_eager_reference"] = {
                "what": "informative (SURVEY 8d-ii, 'the real bar'): the same unmodified reference modules as eager PyTorch + "
                        "functorch.vmap on this same B200 (data_device = training_device = cuda), all %d objects, per-frame "
                        "ingestion + sampling + steps; NOT the driver's reference arm" % eager["n_obj"],
                "value": eager["rays_per_s"], "unit": "rays/s", "ms_per_step": eager["ms_per_step"], "steps": eager["steps_done"],
                "speedup_of_value": rays / (ms * 1e-3) / eager["rays_per_s"] if world == 1 else None}
        print(json.dumps(out), file=_JSON_OUT, flush=True)
    D.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------------------
# BASELINE config 5: render-only evaluation (full-frame RGB / depth / part-feature maps of every object, all-gather, z-merge)
# ------------------------------------------------------------------------------------------------------------
def run_eval(args):
    import numpy as np
    import torch
    from openobj_b200 import _lib, cfg as C, dist as D, eval as E, utils as U, vmap as V
    from openobj_b200.synthetic import object_layout
    rank, world, local = D.init_from_env()
    if not torch.cuda.is_available():
        raise _lib.OOError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = C.room0_config()
    cfg.training_device = cfg.data_device = str(dev)
    W, H = cfg.W, cfg.H
    n_total = args.objects
    cam = V.cameraInfo(cfg)
    boxes = object_layout(n_total, W, H, np.random.default_rng(0))
    torch.manual_seed(0)
    blank = (torch.zeros(W, H, 3, dtype=torch.uint8, device=dev), torch.ones(W, H, device=dev))
    mine = []
    for k, (x0, y0, x1, y1) in enumerate(boxes):
        if k % world != rank:
            torch.manual_seed(1000 + k)
            continue
        torch.manual_seed(1000 + k)
        o = V.sceneObject(cfg, k + 1, blank[0], blank[1], None, torch.tensor([x0, x1, y0, y1]), torch.eye(4), 0, shared=True)
        with torch.no_grad():
            o.trainer.fc_occ_map.out_alpha.bias.fill_(0.8)          # opaque enough for the opacity >= 0.9 test (vmap.py:665)
        z = 2.0 + 0.02 * (k % 50)
        bb = U.BoundingBox()
        bb.R = np.eye(3)
        bb.center = np.array([((x0 + x1) / 2 - cfg.cx) / cfg.fx * z, ((y0 + y1) / 2 - cfg.cy) / cfg.fy * z, z])
        bb.extent = np.array([(x1 - x0) / cfg.fx * z, (y1 - y0) / cfg.fy * z, 0.6])
        o.bbox3dour = bb
        mine.append(o)
    steps, warmup = args.steps, max(min(args.warmup, 10), 3)

    def pose(f):
        T = np.eye(4)
        T[:3, 3] = [0.01 * np.sin(0.3 * f), 0.01 * np.cos(0.2 * f), -0.002 * (f % 10)]
        return T

    stats = {}
    hits = torch.zeros(1, dtype=torch.int64, device=dev)
    for f in range(warmup):
        E.render_frame(mine, pose(f), cam.rays_dir_cache, is_bg={0: True}, render_feat=True, stats=stats)
    torch.cuda.synchronize(); D.barrier()
    e0, e1 = cuda_timer()
    clk = ClockSampler(local)
    clk.__enter__()
    time.sleep(0.02)
    torch.cuda.synchronize(); D.barrier()
    e0.record()
    for f in range(steps):
        depth, rgb, winner, feat = E.render_frame(mine, pose(warmup + f), cam.rays_dir_cache, is_bg={0: True}, render_feat=True,
                                                  stats=stats)
    e1.record()
    torch.cuda.synchronize()
    clk.__exit__()
    D.barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1), dev)
    # per-kernel view: K5 (hit lists + the tcgen05 render launch) of this rank's objects, one frame, CUDA events inside render_frame
    st2 = {"time_kernels": True}
    E.render_frame(mine, pose(0), cam.rays_dir_cache, is_bg={0: True}, render_feat=True, stats=st2)
    k5_ms = st2.get("k5_ms", float("nan"))
    n_hit_local = st2.get("hit_rays_local", 0)
    n_hit_all = int(D.sum_over_ranks(n_hit_local, dev))
    covered = float((winner >= 0).float().mean())
    if rank == 0:
        mac_pt = MAC_PER_POINT - 63 + 63 + MAC_CLIP_POINT                  # PE + trunk + colour head + clip_linear per sample point
        k5_flop = 2.0 * mac_pt * 149 * n_hit_local
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        from openobj_b200 import ops
        peak = ops.fma_peak_tflops()
        out = {
            "metric": "full-frame eval frames/sec (every object rendered over all pixels: depth, RGB, 512-d part feature; "
                      "all-gather; depth-test merge)", "value": steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: render-only eval, %dx%d frames, %d objects in total (%d on rank 0), 149 "
                                   "samples per ray, render_part on; one step = one pose" % (W, H, n_total, len(mine)),
                       "objects_total": n_total, "hit_rays_per_frame": n_hit_all, "pixels_covered": covered,
                       "parallelism": "objects sharded by ensemble index, rank = k mod %d; one all_gather_into_tensor of 8 B per "
                                      "pixel and object + winner-only feature rows" % world},
            "object_rays_per_s": n_hit_all * steps / (ms * 1e-3),
            "clocks": clk.summary(),
            "interconnect": stats,
            "roofline": {"kernel": "oo_render_frame = k_hit_count / _scan / _fill + k_forward_tc<render> (K5 on tcgen05 / TMEM: 149 samples "
                                   "per hit ray, 3xTF32 tcgen05.mma, compositing), this rank's %d objects of one frame, one launch"
                                   % len(mine), "bound": "fma", "achieved": k5_flop / (k5_ms * 1e-3) / 1e12,
                         "peak": peak, "unit": "TFLOP/s", "frac": k5_flop / (k5_ms * 1e-3) / 1e12 / peak, "traffic": None,
                         "ms": k5_ms, "flop": k5_flop},
            "gpu_launches": steps * (4 + 1 + 1),       # hit count / scan / fill, render, merge, winner features
        }
        print(json.dumps(out), file=_JSON_OUT, flush=True)
    D.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--objects", type=int, default=60, help="objects per GPU")
    ap.add_argument("--part", type=int, default=1, help="part-level feature head on (room_0.json part_mode)")
    ap.add_argument("--fill-frames", type=int, default=20, help="untimed frames that fill the keyframe rings (SURVEY 8d)")
    ap.add_argument("--bg", type=int, default=0, help="1 = also train the separate background model every step (room_0.json do_bg)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline / GPU-eager reference legs")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json configs (1-based): 2 = room_0 shape, --objects per GPU (weak scaling; the headline); "
                         "3 = 100 objects in total, CLIP + part heads (strong scaling); 4 = ScanNet shape 640x480, 200 objects, "
                         "part features off (strong scaling); 5 = render-only eval: --objects objects in total, one step = one "
                         "pose (full-frame depth / RGB / part-feature maps, all-gather, z-merge)")
    ap.add_argument("--trace", action="store_true", help="per-rank host timings of every frame (stderr)")
    ap.add_argument("--device", default="cpu", help="--impl reference only: cpu (the driver's arm) or cuda:0 (eager reference)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.config == 5:
        return run_eval(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
