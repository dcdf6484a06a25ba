"""world_size-2 (and 3) CPU tests of the N > 1 host logic over the gloo backend: the partition plans of
cvvidproc_b200/sharded.py, the exchange protocol of csrc/median_shard.cu walked with a numpy model
(tests/shard_model.py) where the GPU job uses peer stores, the frame-order re-assembly of the frame-sharded highlight
stage, and bench.py's rule that rank 0 alone runs the reference arm."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

from cvvidproc_b200 import sharded  # noqa: E402


def test_frame_chunk_partitions_every_frame_once():
    for n in (0, 1, 2, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, cnt = sharded.frame_chunk(n, r, world)
                seen.extend(range(first, first + cnt))
            assert seen == list(range(n))
            sizes = [sharded.frame_chunk(n, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_element_slices_cover_and_align():
    for nelem in (1, 127, 128, 129, 640 * 480, 1920 * 1080, 3840 * 2160):
        for world in (1, 2, 3, 4, 8, 16):
            sl = sharded.element_slices(nelem, world)
            assert len(sl) == world
            assert sum(c for _, c in sl) == nelem
            pos = 0
            for first, cnt in sl:
                assert first == pos or cnt == 0
                assert first % 128 == 0 or cnt == 0
                pos += cnt
    with pytest.raises(ValueError):
        sharded.element_slices(100, 17)


def test_highlight_batches_round_robin_in_frame_order():
    plan = sharded.highlight_batches(1000, 4, 64)
    assert [p[1] for p in plan] == list(range(0, 1000, 64))
    assert [p[0] for p in plan] == [i % 4 for i in range(len(plan))]
    assert sum(p[2] for p in plan) == 1000 and plan[-1][2] == 1000 - 64 * 15


def _worker(rank, world, port, nframes, nelem, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import shard_model as sm

        rng = np.random.default_rng(seed)
        stack = rng.integers(90, 150, (nframes, nelem), dtype=np.uint8)  # every rank builds the same job ...
        first, cnt = sharded.frame_chunk(nframes, rank, world)
        mine = stack[first : first + cnt]                                 # ... and keeps only its frame chunk
        slice_len = (-(-nelem // world) + 127) // 128 * 128  # elements per owner, as sharded.element_slices rounds it
        assert sharded.element_slices(nelem, world)[0][1] == min(slice_len, nelem)

        def push(counts):
            """what the counting kernel's peer stores do: owner r receives every rank's counts of ITS elements"""
            padded = np.zeros((world * slice_len, 16), np.uint16)
            padded[:nelem] = counts
            send = [torch.from_numpy(padded[r * slice_len : (r + 1) * slice_len].astype(np.int32)) for r in range(world)]
            recv = [torch.empty_like(send[0]) for _ in range(world)]
            # gloo has no all_to_all on CPU tensors in every build: one gather per owner
            for owner in range(world):
                dist.gather(send[owner], recv if rank == owner else None, dst=owner)
            return np.stack([t.numpy() for t in recv]).astype(np.uint16)  # (world, slice_len, 16)

        def broadcast(own_values, dtype):
            """owner -> every rank (sel array, result image)"""
            t = torch.from_numpy(own_values.astype(np.int64))
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            return np.concatenate([p.numpy() for p in parts])[:nelem].astype(dtype)

        # ---- the one-pass form first (phases 4, 5): window records to the owners, the owners decide what they can
        def push_records(rec):
            padded = np.zeros((world * slice_len, 11), np.int64)
            padded[:nelem] = rec
            send = [torch.from_numpy(padded[r * slice_len : (r + 1) * slice_len].copy()) for r in range(world)]
            recv = [torch.empty_like(send[0]) for _ in range(world)]
            for owner in range(world):
                dist.gather(send[owner], recv if rank == owner else None, dst=owner)
            return np.stack([t.numpy() for t in recv])

        recs = push_records(sm.window_records(mine))                  # phase 4
        dist.barrier()
        res_own, ok_own = sm.owner_window_final(recs)                 # phase 5
        res_w = broadcast(res_own, np.uint8)
        undecided = torch.tensor([int((~ok_own[: max(0, min(slice_len, nelem - rank * slice_len))]).sum())])
        dist.all_reduce(undecided)                                    # every rank learns the same number
        dist.barrier()
        want = np.sort(stack, axis=0)[nframes // 2]
        ok_all = broadcast(ok_own.astype(np.uint8), np.uint8).astype(bool)
        window_ok = bool(np.array_equal(res_w[ok_all], want[ok_all])) and int(undecided[0]) == int((~ok_all).sum())
        np.save(os.path.join(out_dir, f"win_{rank}.npy"), np.array([window_ok, int(undecided[0])]))
        # ---- the two-round exchange (what runs when `undecided` is not zero; exact for any input)
        c1 = push(sm.nibble_counts(mine, "hi"))                       # phase 0
        dist.barrier()
        sel_own = sm.owner_pick_hi(c1)                                # phase 1 (pad elements pick garbage, cut below)
        sel = broadcast(sel_own, np.uint32)
        dist.barrier()
        c2 = push(sm.nibble_counts(mine, "lo", sel))                  # phase 2
        dist.barrier()
        res = broadcast(sm.owner_pick_lo(c2, sel_own), np.uint8)      # phase 3
        dist.barrier()
        np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([np.array_equal(res, want)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nframes,nelem", [(2, 101, 300), (2, 64, 1000), (3, 50, 129), (2, 1, 40), (2, 3, 257)])
def test_exchange_protocol_over_gloo(tmp_path, world, nframes, nelem):
    port = 29500 + (os.getpid() + world * 7 + nframes) % 2000
    mp.spawn(_worker, args=(world, port, nframes, nelem, 11, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / f"ok_{r}.npy")[0], f"rank {r}"
        win = np.load(tmp_path / f"win_{r}.npy")
        assert win[0], f"rank {r}: the one-pass form decided an element wrongly or the ranks disagree on what is left"
    assert len({int(np.load(tmp_path / f"win_{r}.npy")[1]) for r in range(world)}) == 1


def test_window_model_decides_a_noisy_background_and_never_lies(oracle_median):
    """the numpy model of phases 4 + 5: a background with sensor-like noise is decided everywhere; unrelated chunks are
    reported as undecided instead of being answered wrongly"""
    import shard_model as sm

    rng = np.random.default_rng(8)
    n, nelem = 600, 400
    stack = (120 + rng.integers(-3, 5, (n, nelem))).astype(np.uint8)
    want = oracle_median(stack.reshape(n, 1, nelem)).reshape(-1)
    for world in (1, 2, 5):
        recs = np.stack([sm.window_records(stack[sharded.frame_chunk(n, r, world)[0]:][: sharded.frame_chunk(n, r, world)[1]])
                         for r in range(world)])
        res, ok = sm.owner_window_final(recs)
        assert ok.all() and np.array_equal(res, want)
    stack[:300, :100] = rng.integers(0, 256, (300, 100), dtype=np.uint8)
    want = oracle_median(stack.reshape(n, 1, nelem)).reshape(-1)
    recs = np.stack([sm.window_records(stack[:300]), sm.window_records(stack[300:])])
    res, ok = sm.owner_window_final(recs)
    assert ok[100:].all() and not ok[:100].all()
    assert np.array_equal(res[ok], want[ok])


def test_shard_model_matches_oracle(oracle_median):
    """the numpy model itself (single rank) against the CPU oracle"""
    import shard_model as sm

    rng = np.random.default_rng(3)
    for n in (1, 2, 100, 255):
        stack = rng.integers(0, 256, (n, 500), dtype=np.uint8)
        sel = sm.owner_pick_hi(sm.nibble_counts(stack, "hi")[None])
        res = sm.owner_pick_lo(sm.nibble_counts(stack, "lo", sel)[None], sel)
        assert np.array_equal(res, oracle_median(stack.reshape(n, 1, 500)).reshape(-1))


def test_reference_arm_runs_on_rank0_only():
    """bench.py --impl reference under torchrun: ranks other than 0 exit 0 without work or output"""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
