"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: object -> rank assignment, the OR-reduction of the
zero-mask flags, max-over-ranks timing; and the schedule / sharding invariants that need no GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openobj_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # per-step flags: rank 0 sees an empty object mask in step 1, rank 1 an empty semantic mask in step 2 and both in 3
    flags = torch.tensor([[0, 2, 0, 0], [0, 0, 4, 6]][rank], dtype=torch.int32)
    bits = torch.stack([(flags >> 1) & 1, (flags >> 2) & 1], 1).contiguous()     # what oo_label_counts writes as flag_bits
    D.make_flag_allreduce()(bits)
    flags = (bits[:, 0] << 1) | (bits[:, 1] << 2)                                 # what oo_adam_schedule rebuilds
    # a rank that owns no object joins with all-clear bits (Scene.train) and must not change the result
    empty = torch.zeros(4, 2, dtype=torch.int32) if rank == 1 else bits.clone()
    D.make_flag_allreduce()(empty)
    assert empty.tolist() == bits.tolist()
    t = D.max_over_ranks(10.0 + rank, torch.device("cpu"))
    # objects: ensemble index k -> rank k % world, each object owned exactly once
    owned = [k for k in range(11) if D.owner_rank(k, world) == rank]
    gathered = [None] * world
    dist.all_gather_object(gathered, owned)
    D.barrier()
    q.put((rank, flags.tolist(), t, gathered))
    dist.destroy_process_group()


def test_flag_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, flags, t, gathered in res:
        assert flags == [0, 2, 4, 6]              # bitwise OR across ranks
        assert t == 11.0                           # max over ranks
        assert sorted(gathered[0] + gathered[1]) == list(range(11)) and not set(gathered[0]) & set(gathered[1])
        assert gathered[0] == [0, 2, 4, 6, 8, 10] and gathered[1] == [1, 3, 5, 7, 9]


def test_single_process_helpers_are_noops():
    assert D.make_flag_allreduce() is None
    assert D.max_over_ranks(3.5, torch.device("cpu")) == 3.5
    D.barrier()
    assert [D.owner_rank(k, 8) for k in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]


def test_eval_gather_order_is_global_insertion_order():
    from openobj_b200.eval import global_order
    ks, order = global_order([3, 2, 2], 3)            # 7 objects over 3 ranks: rank r holds k = r, r+3, ...
    assert ks == [0, 3, 6, 1, 4, 2, 5]
    assert [ks[q] for q in order] == list(range(7))
    ks, order = global_order([4], 1)
    assert ks == [0, 1, 2, 3] and order == [0, 1, 2, 3]


def test_shard_book_world_larger_than_object_count():
    """fewer objects than ranks (first frames, background-only ranks).  Every rank runs the same bookkeeping:
    same ensemble indices, owners k mod G, a global models-full cap, and a 'new object somewhere' signal on ALL ranks
    (the reference restacks every model then: Adam state restarts everywhere)."""
    world = 4
    books = [D.ShardBook(r, world, cap=5) for r in range(world)]
    frames = [[7, 3], [3, 7, 9], [1, 2, 3, 4, 5, 6]]
    for ids in frames:
        news = []
        for b in books:
            res = [b.see(i) for i in sorted(ids)]
            news.append([r is not None and r[2] for r in res])
        assert all(n == news[0] for n in news)                      # the reset signal is the same on every rank
    assert all(b.global_index == books[0].global_index for b in books)
    assert books[0].global_index == {3: 0, 7: 1, 9: 2, 1: 3, 2: 4}  # order of first appearance, ties by id; capped at 5
    assert [sorted(b.local_index) for b in books] == [[2, 3], [7], [9], [1]]
    assert books[2].see(6) is None and books[1].see(7) == (1, 0, False)
