"""Host-side plumbing of the multi-GPU forms of the hot path (one process per GPU, torch.distributed).

    ShardedMedian     frame-sharded temporal median: every rank holds a chunk of the frames; the merge is the
                      count exchange of csrc/median_shard.cu -- ONE pass of window counting around per-launch pilot
                      medians, and the two-round nibble-count exchange when that leaves an element undecided (counts
                      are written straight into the owner rank's memory over NVLink by the counting kernels;
                      torch.distributed only carries the 64-byte buffer handles once and a one-element all-reduce as
                      the barrier between phases).
    frame_chunk       which frames of a job a rank takes (contiguous ranges, like the reference's per-generator
                      frame ranges, /root/reference/Sources/cv_vid_bg_helpers.cpp:84-120)
    element_slices    which elements a rank owns in the exchange
    highlight_frames  which frames of a highlight job a rank takes (round-robin batches; frames are independent,
                      highlight_objects_algo.h:60-69) and the order in which the host re-assembles them
                      (the MatSetIntermediary role, mat_set_intermediary.h:50-68)

Nothing here computes on the CPU: without the CUDA library the classes raise.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

MAX_RANKS = 16  # kMaxShardRanks in csrc/context.hpp


def frame_chunk(nframes: int, rank: int, world: int) -> Tuple[int, int]:
    """(first, count) of the contiguous frame range rank takes; the first `nframes % world` ranks take one more."""
    if world < 1 or not 0 <= rank < world or nframes < 0:
        raise ValueError("bad frame_chunk arguments")
    base, extra = divmod(nframes, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def element_slices(nelem: int, world: int) -> List[Tuple[int, int]]:
    """(first, count) of the elements every rank owns in the exchange: equal slices rounded up to 128 elements so
    that no tile of the counting kernel straddles two owners (csrc/median_shard.cu)."""
    if world < 1 or world > MAX_RANKS or nelem < 1:
        raise ValueError("bad element_slices arguments")
    slice_ = (-(-nelem // world) + 127) // 128 * 128
    out = []
    for r in range(world):
        first = min(r * slice_, nelem)
        out.append((first, min(slice_, nelem - first)))
    return out


def highlight_batches(nframes: int, world: int, batch: int) -> List[Tuple[int, int, int]]:
    """Round-robin assignment of consecutive batches to ranks: [(rank, first_frame, count), ...] in frame order.
    The host consumer walks this list to hand masks to the tracker strictly in frame order."""
    if batch < 1 or world < 1 or nframes < 0:
        raise ValueError("bad highlight_batches arguments")
    out = []
    for i, first in enumerate(range(0, nframes, batch)):
        out.append((i % world, first, min(batch, nframes - first)))
    return out


class ShardedMedian:
    """One rank of a frame-sharded median job.

    barrier: a callable that orders "every rank finished the previous phase" before anything enqueued after it on
    the context's stream (default: a one-element all-reduce on `group`, issued on the context's stream).
    """

    def __init__(self, ctx, nelem: int, rank: int, world: int, barrier: Callable[[], None] | None = None, group=None,
                 max_rank_frames: int | None = None):
        if world > MAX_RANKS:
            raise ValueError(f"at most {MAX_RANKS} ranks")
        self.ctx, self.nelem, self.rank, self.world = ctx, nelem, rank, world
        self._group = group
        self._barrier = barrier
        self._flag = None
        self._stream = None
        self.max_rank_frames = max_rank_frames
        self.last_unresolved = None  # elements the one-pass form left undecided in the last run()
        self.barrier_kind = "caller-supplied"
        ctx.median_shard_begin(nelem, rank, world, max_rank_frames)
        self._open = True

    # -- wiring ----------------------------------------------------------------------------------------------------
    def connect_processes(self):
        """Exchange the buffer handles with the other processes of the group (torch.distributed) and map them."""
        import torch
        import torch.distributed as dist

        mine = self.ctx.median_shard_export()
        handles: List[bytes | None] = [None] * self.world
        dist.all_gather_object(handles, mine, group=self._group)
        for peer, h in enumerate(handles):
            if peer != self.rank:
                self.ctx.median_shard_import(peer, h)
        if self._barrier is None:
            # Every rank on a GPU of its own (what torchrun launches): the library's one-warp flag barrier over the mapped
            # exchange buffers.  Ranks that share a device must not spin on each other on it: a one-element NCCL
            # all-reduce on the context's stream instead.  CVVP_SHARD_BARRIER=nccl|device forces either.
            import os

            dev = torch.device("cuda", self.ctx.device)
            mine_id = (os.uname().nodename, str(torch.cuda.get_device_properties(self.ctx.device).uuid))
            ids = [None] * self.world
            dist.all_gather_object(ids, mine_id, group=self._group)
            own_gpu = len(set(ids)) == self.world
            choice = os.environ.get("CVVP_SHARD_BARRIER", "device" if own_gpu else "nccl")
            if choice == "device" and own_gpu:
                self.barrier_kind = "device flags over peer memory"
                self._barrier = self.ctx.median_shard_barrier
            else:
                self.barrier_kind = "one-element NCCL all-reduce"
                self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
                self._stream = torch.cuda.ExternalStream(self.ctx.stream, device=dev)

                def barrier():
                    with torch.cuda.stream(self._stream):
                        dist.all_reduce(self._flag, group=self._group)

                self._barrier = barrier
        dist.barrier(group=self._group)  # every rank has mapped every buffer before the first store

    @staticmethod
    def connect_local(members: Sequence["ShardedMedian"]):
        """Several ranks inside ONE process (tests; single-process multi-GPU): attach the contexts to each other.
        The barrier then is: synchronize every member's context."""
        for a in members:
            for b in members:
                if a is not b:
                    a.ctx.median_shard_attach(b.rank, b.ctx)

        def barrier():
            for m in members:
                m.ctx.synchronize()

        for m in members:
            m._barrier = barrier

    # -- one job ---------------------------------------------------------------------------------------------------
    def phase(self, p: int, d_frames: int = 0, nframes: int = 0, frame_stride: int = 0):
        self.ctx.median_shard_phase(p, d_frames, nframes, frame_stride)

    def barrier(self):
        self._barrier()

    def window_capacity(self) -> int:
        """frames per rank the one-pass form has record slots for"""
        return max(1024, -(-(self.max_rank_frames or 1024) // 1024) * 1024)

    def run_window(self, d_frames: int, nframes: int, frame_stride: int) -> int:
        """The one-pass form: phase 4, barrier, phase 5, barrier; returns the number of undecided elements (waits for
        the stream; the same number on every rank)."""
        for p in (4, 5):
            self.phase(p, d_frames, nframes, frame_stride)
            self._barrier()
        self.last_unresolved = self.ctx.median_shard_unresolved()
        return self.last_unresolved

    def run_two_round(self, d_frames: int, nframes: int, frame_stride: int, restricted: bool = False):
        """The four phases of the two-round nibble exchange with the barriers in between (exact for any input).
        restricted: only over the 128-element tiles the one-pass form flagged (phases 10..13; after run_window)."""
        for p in range(4):
            self.phase(p + (10 if restricted else 0), d_frames, nframes, frame_stride)
            self._barrier()

    def run(self, d_frames: int, nframes: int, frame_stride: int, window: bool = True) -> int:
        """One job (every rank of the group must call it with the same `window`): the one-pass form when every rank's
        frames fit its record slots, then -- only if it left elements undecided, which every rank learns as the same
        number -- the two-round exchange over the tiles that hold such an element.  Returns the device pointer of the full result image (nelem bytes), complete
        once the stream reaches the last barrier."""
        if not window:
            self.run_two_round(d_frames, nframes, frame_stride)
        elif self.run_window(d_frames, nframes, frame_stride) != 0:
            self.run_two_round(d_frames, nframes, frame_stride, restricted=True)
        return self.ctx.median_shard_result()

    def result_ptr(self) -> int:
        return self.ctx.median_shard_result()

    def close(self):
        if self._open:
            self._open = False
            self.ctx.median_shard_end()
