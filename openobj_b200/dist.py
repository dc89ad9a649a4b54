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


class ShardBook:
    """Which objects exist and who owns them: every rank sees every frame's instance ids (train.py:191) and runs this same
    bookkeeping, so all ranks agree on the ensemble index k of an object (order of first appearance, ties by ascending id),
    on its owner k mod G, on the global "models full" cap (train.py:231-233) and on WHEN a new object appeared anywhere --
    the event at which the reference restacks all models and Adam's state restarts for everybody (train.py:272-276)."""

    def __init__(self, rank=0, world=1, cap=100):
        self.rank, self.world, self.cap = int(rank), int(world), int(cap)
        self.global_index = {}            # obj id -> k
        self.local_index = {}             # obj id -> position among this rank's objects

    def full(self):
        return len(self.global_index) >= self.cap

    def see(self, obj_id):
        """-> (k, local index or None, new) ; None when the object is dropped because the model table is full."""
        k = self.global_index.get(obj_id)
        new = k is None
        if new:
            if self.full():
                return None
            k = len(self.global_index)
            self.global_index[obj_id] = k
            if owner_rank(k, self.world) == self.rank:
                self.local_index[obj_id] = len(self.local_index)
        return k, self.local_index.get(obj_id), new


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
    """OR of the zero-mask bits across ranks: the caller hands an int32 tensor [iters, 2] holding ONE int per bit
    (oo_label_counts' flag_bits), so a single MAX all-reduce is the OR (NCCL has no bitwise reduction)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None

    def allreduce(bits):
        dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=group)
    return allreduce


def max_over_ranks(value_ms, device):
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
