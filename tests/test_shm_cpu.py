"""World-size-2 gloo tests (CPU) of the node-wide shared-memory plumbing of the sharded output mode
(``anemoi_graphs_b200/shm.py``): the control block's all-gather / publish log / barrier, the shared host arena
(rank 0 decides which segment backs a request, segments are reused once rank 0's tensors are gone, every rank sees the
same bytes), and the assembly of one complete edge list from per-rank blocks without any collective."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from anemoi_graphs_b200 import device as D
    from anemoi_graphs_b200 import shm

    try:
        group = shm.local_group()
        assert group is not None and (group.rank, group.world) == (rank, world)
        # all-gather of small int64 vectors, many rounds (the ring wraps)
        for i in range(3 * shm.RING // 2):
            got = group.all_gather([rank * 1000 + i, -i])
            assert got == [[r * 1000 + i, -i] for r in range(world)]
        # the log: rank 0 publishes, the others follow
        for i in range(10):
            vals = group.publish([i, i * i, 7] if rank == 0 else [])
            assert vals[:3] == [i, i * i, 7]
        # long waits sleep in read(2) on a FIFO: rank 0 stamps, then wakes every follower with one byte
        import time

        for i in range(3):
            if rank == 0:
                time.sleep(0.02)
                group.order[i, 0] = 100 + i
                group.wake_followers()
            else:
                t0 = time.perf_counter()
                group.sleep_until_woken()
                assert int(group.order[i, 0]) == 100 + i and time.perf_counter() - t0 > 0.005
            group.barrier()
        # uneven per-rank counts (what CutOffEdges exchanges)
        assert D.exchange_counts(5 if rank == 0 else 2, torch.device("cpu")) == [5, 2]

        # ---- arena: same bytes on both ranks -------------------------------------------------------------
        arena = shm.arena()
        counts = [5, 2]
        shard = D.Shard(rank, world, counts)
        full = arena.tensor((2, shard.total), torch.int32)
        assert full.shape == (2, 7) and full.is_shared()
        # names live only until every rank has mapped them: a job that is killed leaves nothing in /dev/shm
        assert not [f for f in os.listdir("/dev/shm") if f.startswith(f"agx_{group.token}")]
        block = torch.full((2, counts[rank]), rank + 1, dtype=torch.int32)
        full[:, shard.offset : shard.offset + counts[rank]] = block  # every rank writes only its own columns
        group.barrier()
        assert full[0].tolist() == [1] * 5 + [2] * 2 and full[1].tolist() == [1] * 5 + [2] * 2
        group.barrier()
        # ---- segments are reused when rank 0 no longer references them, and not before --------------------
        first_ptr_offset = full.data_ptr() - next(iter(arena.segments.values())).data_ptr()
        n_seg = len(arena.segments)
        keep = arena.tensor((1000,), torch.float32)  # same size class while `full` is alive: a second segment
        assert len(arena.segments) == n_seg + 1
        keep[:] = float(rank)
        del full
        again = arena.tensor((2, 7), torch.int32)  # `full` is gone on rank 0: its segment comes back
        assert len(arena.segments) == n_seg + 1
        assert again.data_ptr() - next(iter(arena.segments.values())).data_ptr() == first_ptr_offset
        again.zero_()
        group.barrier()
        again[rank, :] = 10 + rank
        group.barrier()
        assert again[0].tolist() == [10] * 7 and again[1].tolist() == [11] * 7
        # a request that differs between the ranks is a programming error and is reported, not silently mis-mapped
        group.barrier()
        try:
            arena.allocate(64 if rank == 0 else 128)
            bad = rank == 0  # rank 0 cannot see it; the follower must
        except RuntimeError:
            bad = rank != 0
        assert bad
        # ---- replicated tensors: each rank is responsible for its 1/W slice -----------------------------
        lo, hi = D.shard_range(10, rank, world)
        assert (lo, hi) == ((0, 5) if rank == 0 else (5, 10))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_shared_memory_group_and_arena_world2():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
    left = [f for f in os.listdir("/dev/shm") if f.startswith("agx_")]
    assert not left, f"shared-memory segments were not removed: {left}"


def _worker_many(rank: int, world: int, port: int) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from anemoi_graphs_b200 import device as D
    from anemoi_graphs_b200 import shm

    try:
        group, arena = shm.local_group(), shm.arena()
        assert group is not None and group.world == world
        counts = [3 + r for r in range(world)]
        for build in range(3):  # three "builds": the segments of the first are reused by the third
            shard = D.Shard(rank, world, counts)
            full = arena.tensor((2, shard.total), torch.int32)
            keep = arena.tensor((shard.total, 3), torch.float32)
            full[:, shard.offset : shard.offset + counts[rank]] = 100 * build + rank
            keep[shard.offset : shard.offset + counts[rank]] = float(build)
            group.barrier()
            want = np.concatenate([np.full(c, 100 * build + r, dtype=np.int32) for r, c in enumerate(counts)])
            assert np.array_equal(full[0].numpy(), want) and np.array_equal(full[1].numpy(), want)
            assert float(keep.sum()) == build * 3 * shard.total
            group.barrier()
            del full, keep
        assert len(arena.segments) <= 4  # two size classes, at most two generations each
        assert not [f for f in os.listdir("/dev/shm") if f.startswith(f"agx_{group.token}")]
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_arena_rendezvous_world4():
    """Four ranks: every new segment is a rendezvous (all ranks map it, then its name goes away), reuse across builds."""
    mp.spawn(_worker_many, args=(4, _free_port()), nprocs=4, join=True)


def _worker_small_shm(rank: int, world: int, port: int) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from anemoi_graphs_b200 import device as D
    from anemoi_graphs_b200 import shm

    try:
        if rank == 0:  # only rank 0 looks at /dev/shm; its answer reaches the others with the token broadcast
            shm.MIN_SHM_FREE_BYTES = 1 << 62
        assert shm.local_group() is None and shm.arena() is None
        assert shm.local_group() is None  # remembered: no second collective
        # a tensor that must be node-wide cannot be had, and says so
        D.set_sharded_output(True)
        try:
            D.host_tensor((4,), torch.int32, require_shared=True)
            raised = False
        except NotImplementedError:
            raised = True
        assert raised
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_too_small_dev_shm_is_detected_not_crashed_into():
    """A container with the 64 MB default /dev/shm: writing result buffers there would end in SIGBUS. Rank 0 checks the
    free space; every rank learns that there is no node-wide group (bench.py then stays in gathered mode)."""
    mp.spawn(_worker_small_shm, args=(2, _free_port()), nprocs=2, join=True)


def test_shard_bookkeeping():
    from anemoi_graphs_b200.device import Shard

    s = Shard(1, 3, [4, 0, 6])
    assert (s.total, s.offset) == (10, 4) and s.describe()["counts"] == [4, 0, 6]
    r = Shard(2, 4, [9], replicated=True)
    assert (r.total, r.offset) == (9, 0)
    assert np.array_equal(np.cumsum([0] + s.counts)[:-1], [0, 4, 4])
