"""Per-phase device times of the frame-sharded median on ONE GPU (developer aid).

    python tools/probe_shard.py [world]      # `world` ranks emulated as contexts of this process on cuda:0

Every emulated rank holds a full 1080p x 1000-frame chunk (weak scaling: the job is 1000*world frames), so the
phase-0/2 times are what one GPU of a real job spends; phases 1/3 handle 1/world of the elements each.
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cvvidproc_b200 import _cabi, sharded  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    W, H, N = 1920, 1080, 1000
    nelem = W * H
    ctxs = [_cabi.Context(0) for _ in range(world)]
    jobs = [sharded.ShardedMedian(ctxs[r], nelem, r, world, max_rank_frames=N) for r in range(world)]
    sharded.ShardedMedian.connect_local(jobs)
    stack = torch.empty((N, nelem), dtype=torch.uint8, device="cuda:0")
    ctxs[0].synth_frames_device(stack.data_ptr(), nelem, W, H, 0, N, 2, 30)
    ctxs[0].synchronize()
    single = torch.empty(nelem, dtype=torch.uint8, device="cuda:0")
    streams = [torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", 0)) for c in ctxs]
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    times = {p: [] for p in range(6)}
    t_single = []
    for it in range(6):
        a, b = ev(), ev()
        a.record(streams[0])
        ctxs[0].median_device(stack.data_ptr(), N, nelem, nelem, single.data_ptr())
        b.record(streams[0])
        torch.cuda.synchronize()
        t_single.append(a.elapsed_time(b))
        for p in (4, 5, 0, 1, 2, 3):
            a, b = ev(), ev()
            a.record(streams[0])
            jobs[0].phase(p, stack.data_ptr(), N, nelem)
            b.record(streams[0])
            for r in range(1, world):
                jobs[r].phase(p, stack.data_ptr(), N, nelem)
            torch.cuda.synchronize()
            times[p].append(a.elapsed_time(b))
            if p == 5:
                left = jobs[0].ctx.median_shard_unresolved()
                win = np.array_equal(ctxs[0].copy_to_host(jobs[0].result_ptr(), nelem), single.cpu().numpy())
    got = ctxs[0].copy_to_host(jobs[0].result_ptr(), nelem)
    same = np.array_equal(got, single.cpu().numpy())
    print(f"world {world}: one-pass form p4 {np.median(times[4][1:]):.3f} ms, p5 {np.median(times[5][1:]):.3f} ms; "
          f"undecided elements {left}; result == single-GPU: {win}")
    print(f"world {world}: single-GPU select {np.median(t_single[1:]):.3f} ms; rank-0 phases "
          + ", ".join(f"p{p} {np.median(times[p][1:]):.3f} ms" for p in range(4))
          + f"; sum {sum(np.median(times[p][1:]) for p in range(4)):.3f} ms; result == single-GPU: {same}")
    for j in jobs:
        j.close()
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    main()
