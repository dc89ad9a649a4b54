"""Object sharding across the GPUs of one box: one process per GPU, objects assigned round-robin by
ensemble index, NO gradient collectives.  The only data-path exchange is the OR of the per-step zero-mask flags
(render_rays.py:89-94 couples objects through `(mask_num == 0).any()`), one tiny all-reduce per frame."""
import os

import torch
import torch.distributed as dist


def owner_rank(ensemble_index, world):
    """gpu(k) = k mod G: deterministic, keeps the load balanced as objects appear over time."""
    return ensemble_index % world


def init_from_env(backend=None):
    """(rank, world, local_rank); initialises torch.distributed when launched under torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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
