"""Object sharding across the GPUs of one box: one process per GPU, objects assigned round-robin by
ensemble index, NO gradient collectives.  The only data-path exchange is the OR of the per-step zero-mask flags
(render_rays.py:89-94 couples objects through `(mask_num == 0).any()`), one tiny all-reduce per frame."""
import os
import sys

import torch
import torch.distributed as dist


def owner_rank(ensemble_index, world):
    """gpu(k) = k mod G: deterministic, keeps the load balanced as objects appear over time."""
    return ensemble_index % world


def bind_to_gpu_cpus(local):
    """Pin this process to the CPUs NVML reports as local to GPU `local` (its NUMA node), BEFORE the CUDA context and the
    pinned host buffers exist: with one process per GPU the per-frame host->device copies (76 MB at Replica size) then come
    from memory on the GPU's own socket.  Returns the CPU list, or None when NVML gives no usable answer (then nothing is
    changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local
        if vis:
            ent = vis.split(",")[local].strip()
            if ent.isdigit():
                idx = int(ent)
            else:
                return None
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def init_from_env(backend=None):
    """(rank, world, local_rank); initialises torch.distributed when launched under torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and torch.cuda.is_available() and os.environ.get("OO_NO_CPU_BIND") != "1":
        cpus = bind_to_gpu_cpus(local)
        if cpus is not None:
            print("[openobj_b200.dist] rank %d bound to %d CPUs local to GPU %d" % (rank, len(cpus), local), file=sys.stderr)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank, world_size=world)
    return rank, world, local


def make_flag_allreduce(group=None):
    """OR-reduce OO_FLAG_* bit masks across ranks.  NCCL has no bitwise reduction: MAX over the two bits separately."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None

    def allreduce(flags):
        bits = torch.stack([(flags >> 1) & 1, (flags >> 2) & 1], dim=1).contiguous()
        dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=group)
        flags.copy_((bits[:, 0] << 1) | (bits[:, 1] << 2))
    return allreduce


def max_over_ranks(value_ms, device):
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
