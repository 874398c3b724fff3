"""Host-side shared memory between the ranks of ONE node (one process per GPU).

A multi-GPU build that leaves its result on the HOST does not need any device-to-device exchange: every rank copies
its shard of every edge tensor over its OWN PCIe link straight into the final position of ONE buffer that all ranks
map - POSIX shared memory, page-locked in each process with ``cudaHostRegister`` so the copy engines can write it.
This module is that plumbing, and nothing in it touches a GPU except the registration call:

* ``LocalGroup``  - a control block in ``/dev/shm`` shared by the ranks: sequence-numbered all-gathers of a few int64
  (edge counts of uneven shards), a one-directional publish log (rank 0 decides, the others follow), a barrier;
* ``HostArena``   - shared, page-locked host buffers.  Rank 0 decides which segment backs a request (it alone tracks
  which segments are still referenced by tensors of an earlier build) and publishes the decision; the other ranks
  map the same segment.  Segments are reused from build to build like a caching allocator's blocks.  Every name in
  ``/dev/shm`` is removed as soon as all ranks hold the object open, so a job that is killed leaves nothing behind;
* the same arena carries the node order of a provisionally numbered node set from the rank that sorts to the ranks
  that wait (``device.Provisional``): one host sort per node instead of one per GPU.

Works without CUDA as well (the gloo tests of the protocol): registration is skipped when there is no device.
"""

from __future__ import annotations

import atexit
import os
import time

import numpy as np
import torch

RING = 256  # entries of each sequence-numbered ring in the control block
PAYLOAD = 16  # int64 words per entry and rank
SEGMENT_ALIGN = 2 << 20
# free space /dev/shm must offer before the ranks of a node share buffers through it (two generations of a large
# graph's host tensors: 2 x 0.7 GB for O1280 -> res 7)
MIN_SHM_FREE_BYTES = int(float(os.environ.get("AGX_MIN_SHM_FREE_BYTES", "4e9")))


def _spin(ready, what: str, timeout_s: float = 120.0) -> None:
    """Wait until ``ready()``: a few hundred polls that only yield the GIL (the ranks of a build run in lockstep, most
    waits end within microseconds), then sleeping polls - a rank must never burn a core for milliseconds: on these
    hosts a spinning hyperthread slows the sibling that runs the node-order sort by a fifth."""
    n = 0
    t0 = None
    while not ready():
        n += 1
        if n < 300:
            time.sleep(0)
        else:
            if t0 is None:
                t0 = time.perf_counter()
            elif time.perf_counter() - t0 > timeout_s:
                raise TimeoutError(f"shared-memory rendezvous timed out waiting for {what}")
            time.sleep(20e-6)


class LocalGroup:
    """Control block shared by the ``world`` ranks of this node.

    Layout (int64 words): ``gather`` ring [RING][world][1 + PAYLOAD] (word 0 = stamp), ``log`` ring
    [RING][1 + PAYLOAD] written by rank 0 only, ``order`` stamps [RING][2]."""

    def __init__(self, rank: int, world: int, token: str) -> None:
        self.rank, self.world, self.token = int(rank), int(world), token
        self.n_gather = RING * self.world * (1 + PAYLOAD)
        self.n_log = RING * (1 + PAYLOAD)
        self.n_order = RING * 2
        words = self.n_gather + self.n_log + self.n_order
        self.path = f"/dev/shm/agx_{token}_ctrl"
        if self.rank == 0:
            fd = os.open(self.path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, words * 8)
            os.close(fd)
        else:
            _spin(lambda: os.path.exists(self.path) and os.path.getsize(self.path) == words * 8, "the control block")
        self._mem = np.memmap(self.path, dtype=np.int64, mode="r+", shape=(words,))
        self.gather = self._mem[: self.n_gather].reshape(RING, self.world, 1 + PAYLOAD)
        self.log = self._mem[self.n_gather : self.n_gather + self.n_log].reshape(RING, 1 + PAYLOAD)
        self.order = self._mem[self.n_gather + self.n_log :].reshape(RING, 2)
        self._gather_seq = 0
        self._log_seq = 0
        # one FIFO per follower rank: a follower that has to wait for MILLISECONDS (the node order rank 0 is sorting)
        # blocks in read(2) - asleep in the kernel, no core burnt - and rank 0 wakes it with one byte
        self._fifo_paths = [f"/dev/shm/agx_{token}_wake{r}" for r in range(self.world)]
        if self.rank == 0:
            for r in range(1, self.world):
                os.mkfifo(self._fifo_paths[r], 0o600)
            self._wake_fds = [None] + [os.open(self._fifo_paths[r], os.O_RDWR) for r in range(1, self.world)]
            self._sleep_fd = None
            atexit.register(self._unlink)
        else:
            _spin(lambda: os.path.exists(self._fifo_paths[self.rank]), "the wake-up FIFO")
            self._wake_fds = None
            self._sleep_fd = os.open(self._fifo_paths[self.rank], os.O_RDWR)  # O_RDWR: never blocks in open(2)

    def _unlink(self) -> None:
        for path in [self.path] + self._fifo_paths[1:]:
            try:
                os.unlink(path)
            except OSError:
                pass

    def wake_followers(self) -> None:
        """Rank 0: one wake-up byte to every follower (after publishing what they wait for)."""
        for fd in self._wake_fds[1:]:
            os.write(fd, b"x")

    def sleep_until_woken(self) -> None:
        """Follower: block (asleep in the kernel, GIL released) until rank 0's next wake-up byte."""
        os.read(self._sleep_fd, 1)

    def drain_wakeups(self) -> None:
        """Follower that found what it waited for by polling: swallow the wake-up byte that belongs to it (it may still
        be on its way: the stamp is written before the byte)."""
        os.read(self._sleep_fd, 1)

    # ---- all-gather of up to PAYLOAD int64 per rank ------------------------------------------------------------
    def all_gather(self, values) -> list[list[int]]:
        """Every rank's ``values`` (same length on all ranks), in rank order.  A rendezvous: returns when all ranks
        have contributed."""
        values = [int(v) for v in values]
        assert len(values) <= PAYLOAD
        seq = self._gather_seq
        self._gather_seq += 1
        entry = self.gather[seq % RING]
        mine = entry[self.rank]
        mine[1 : 1 + len(values)] = values
        mine[0] = seq + 1  # stamp last (x86 stores are ordered; numpy writes straight to the mapping)
        _spin(lambda: bool((entry[:, 0] == seq + 1).all()), f"all-gather {seq}")
        return [[int(v) for v in entry[r, 1 : 1 + len(values)]] for r in range(self.world)]

    def barrier(self) -> None:
        self.all_gather([0])

    # ---- one-directional log: rank 0 decides, the others follow -----------------------------------------------
    def publish(self, values) -> list[int]:
        """Rank 0 appends ``values`` to the log (never waits); every other rank reads the next entry (waits until it
        is there).  All ranks call it at the same point of the program; all return rank 0's values."""
        seq = self._log_seq
        self._log_seq += 1
        entry = self.log[seq % RING]
        if self.rank == 0:
            values = [int(v) for v in values]
            assert len(values) <= PAYLOAD
            entry[1 : 1 + len(values)] = values
            entry[1 + len(values) :] = 0
            entry[0] = seq + 1
            return values
        _spin(lambda: int(entry[0]) == seq + 1, f"log entry {seq}")
        return [int(v) for v in entry[1:]]


class HostArena:
    """Shared page-locked host buffers (see the module docstring).  ``allocate`` is called by all ranks in the same
    order; rank 0's bookkeeping decides, the log carries the decision."""

    def __init__(self, group: LocalGroup) -> None:
        if not hasattr(torch._C, "_storage_Use_Count"):  # what tells rank 0 that a segment's tensors are gone
            raise RuntimeError("shared host arena: this torch build has no torch._C._storage_Use_Count")
        self.group = group
        self.segments: dict[int, torch.Tensor] = {}  # id -> uint8 tensor over the whole segment
        self._capacity: dict[int, int] = {}
        self._baseline: dict[int, int] = {}
        self._next_id = 0
        self._registered: list[tuple[int, int]] = []
        self.n_requests = 0
        if group.rank == 0:
            atexit.register(self._unlink_all)

    def _path(self, seg_id: int) -> str:
        return f"/dev/shm/agx_{self.group.token}_seg{seg_id}"

    def _unlink_all(self) -> None:  # only what a failed rendezvous left behind: names go away as soon as all ranks mapped
        for seg_id in list(self._capacity):
            try:
                os.unlink(self._path(seg_id))
            except OSError:
                pass

    @staticmethod
    def _use_count(t: torch.Tensor) -> int:
        return int(torch._C._storage_Use_Count(t.untyped_storage()._cdata))

    def _create(self, seg_id: int, capacity: int) -> None:
        """Rank 0: the file behind a new segment (sized, not yet mapped)."""
        fd = os.open(self._path(seg_id), os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
        os.ftruncate(fd, capacity)
        os.close(fd)
        self._capacity[seg_id] = capacity

    def _map(self, seg_id: int, capacity: int) -> torch.Tensor:
        """Map (and page-lock) a segment in this process; a NEW segment is a rendezvous: once every rank has mapped it
        rank 0 removes its name, so that nothing is left in /dev/shm however the job ends."""
        if seg_id in self.segments:
            return self.segments[seg_id]
        path = self._path(seg_id)
        seg = torch.from_file(path, shared=True, size=capacity, dtype=torch.uint8)
        if torch.cuda.is_available():
            # page-lock the mapping in THIS process so the copy engines can reach it (what NCCL's shm transport does)
            rc = torch.cuda.cudart().cudaHostRegister(seg.data_ptr(), capacity, 1)  # 1 = cudaHostRegisterPortable
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister({capacity} bytes of shared memory) failed: {rc}")
            self._registered.append((seg.data_ptr(), capacity))
        self.segments[seg_id] = seg
        self._capacity[seg_id] = capacity
        self._baseline[seg_id] = self._use_count(seg)
        self.group.barrier()
        if self.group.rank == 0:
            try:
                os.unlink(path)
            except OSError:
                pass
        return seg

    def _is_free(self, seg_id: int) -> bool:
        return self._use_count(self.segments[seg_id]) <= self._baseline[seg_id]

    def allocate(self, nbytes: int) -> torch.Tensor:
        """uint8 tensor of ``nbytes`` in a segment every rank maps (same bytes on all ranks)."""
        nbytes = max(int(nbytes), 1)
        self.n_requests += 1
        if self.group.rank == 0:
            pick = None
            for seg_id, cap in self._capacity.items():  # smallest free segment that fits without wasting > 2x
                if nbytes <= cap <= max(2 * nbytes, SEGMENT_ALIGN) and self._is_free(seg_id):
                    if pick is None or cap < self._capacity[pick]:
                        pick = seg_id
            if pick is None:
                pick = self._next_id
                self._next_id += 1
                cap = (nbytes + SEGMENT_ALIGN - 1) // SEGMENT_ALIGN * SEGMENT_ALIGN
                self._create(pick, cap)  # the file exists before it is announced; all ranks map and page-lock it together
            seg_id, cap = self.group.publish([pick, self._capacity[pick], nbytes])[:2]
        else:
            seg_id, cap, want = self.group.publish([])[:3]
            if want != nbytes:
                self._map(seg_id, cap)  # keep the rendezvous of a new segment: rank 0 must not hang on this rank's error
                raise RuntimeError(f"shared arena: rank {self.group.rank} asks for {nbytes} bytes where rank 0 asked for {want}")
        return self._map(seg_id, cap)[:nbytes]

    def stats(self) -> dict:
        """Segments mapped so far, their bytes, how many are referenced right now (rank 0's view) and the requests
        served: a steady-state build loop must not grow the first two."""
        busy = sum(0 if self._is_free(i) else 1 for i in self.segments)
        return {"segments": len(self.segments), "bytes": int(sum(self._capacity.values())), "busy": busy, "requests": self.n_requests}

    def tensor(self, shape, dtype: torch.dtype) -> torch.Tensor:
        shape = tuple(int(s) for s in shape)
        n = 1
        for s in shape:
            n *= s
        item = torch.empty((), dtype=dtype).element_size()
        return self.allocate(n * item).view(dtype).view(shape) if n else torch.empty(shape, dtype=dtype)


_group: LocalGroup | None = None
_arena: HostArena | None = None
_unavailable = False  # rank 0 found /dev/shm too small: remembered, so that the question is one collective, not many


def local_group() -> LocalGroup | None:
    """The shared-memory group of this process group, or None (single rank, or ranks spread over several nodes).
    Created on first use - a collective: every rank must call it at the same point."""
    global _group, _arena, _unavailable
    if _group is not None:
        return _group
    if _unavailable:
        return None
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    rank, world = dist.get_rank(), dist.get_world_size()
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if local_world != world:  # several nodes: no common /dev/shm
        return None
    import secrets

    # rank 0 decides for everybody: a token, or None when /dev/shm cannot hold the result buffers (a container with the
    # 64 MB default: touching pages beyond the limit would raise SIGBUS inside cudaHostRegister)
    token = [None]
    if rank == 0:
        try:
            st = os.statvfs("/dev/shm")
            free = st.f_bavail * st.f_frsize
        except OSError:
            free = 0
        token = [secrets.token_hex(8) if free >= MIN_SHM_FREE_BYTES else None]
    dist.broadcast_object_list(token, src=0)
    if token[0] is None:
        _unavailable = True
        return None
    _group = LocalGroup(rank, world, token[0])
    _arena = HostArena(_group)
    _group.barrier()
    if rank == 0:  # every rank holds the control block and its FIFO open: the names can go
        _group._unlink()
    return _group


def arena() -> HostArena | None:
    return _arena if local_group() is not None else None


def reset() -> None:
    """Forget the group (tests that re-initialise torch.distributed)."""
    global _group, _arena, _unavailable
    _group, _arena, _unavailable = None, None, False
