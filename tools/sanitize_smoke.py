"""Small end-to-end pass over every kernel for compute-sanitizer (no torch: ctypes + numpy only, so start-up stays short).

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Sizes are tiny (the tools slow kernels down 10-100x); results are still compared with numpy so that a silent
corruption cannot pass as "no error reported"."""
import os
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
from cvvidproc_b200 import _cabi, sharded  # noqa: E402


def main():
    rng = np.random.default_rng(1)
    ctx = _cabi.Context(0)
    # median: on-chip select, both buffering modes, a narrow-tile variant, ragged geometry
    for n, shape in ((5, (3, 70)), (100, (7, 333)), (600, (2, 130)), (1030, (1, 200)), (2100, (1, 140))):
        fr = rng.integers(0, 256, (n,) + shape, dtype=np.uint8)
        os.environ["CVVP_MEDIAN_TWO_PASS"] = "0"
        got = ctx.median(fr, chunk=512)
        assert np.array_equal(got, np.sort(fr, axis=0)[n // 2]), ("single pass", n)
        os.environ["CVVP_MEDIAN_TWO_PASS"] = "1"
        got = ctx.median(fr, chunk=512)
        assert np.array_equal(got, np.sort(fr, axis=0)[n // 2]), ("two pass", n)
    os.environ.pop("CVVP_MEDIAN_TWO_PASS")
    print("median ok", flush=True)
    # frame-sharded median: three ranks emulated on this device, uneven chunks
    nelem, world = 700, 3
    fr = rng.integers(0, 256, (90, nelem), dtype=np.uint8)
    ctxs = [_cabi.Context(0) for _ in range(world)]
    jobs = [sharded.ShardedMedian(ctxs[r], nelem, r, world) for r in range(world)]
    sharded.ShardedMedian.connect_local(jobs)
    parts = [fr[:50], fr[50:50], fr[50:]]
    import ctypes as C

    cudart = C.CDLL("/usr/local/cuda/lib64/libcudart.so")  # device buffers for the device-resident entry points
    cudart.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    cudart.cudaFree.argtypes = [C.c_void_p]
    stride = (nelem + 127) // 128 * 128
    dptrs = []
    for r in range(world):
        p = C.c_void_p()
        assert cudart.cudaMalloc(C.byref(p), max(parts[r].shape[0], 1) * stride) == 0
        if parts[r].shape[0]:
            padded = np.zeros((parts[r].shape[0], stride), np.uint8)
            padded[:, :nelem] = parts[r]
            assert cudart.cudaMemcpy(p, padded.ctypes.data, padded.nbytes, 1) == 0
        dptrs.append(p)
    for ph in range(4):
        for r in range(world):
            jobs[r].phase(ph, dptrs[r].value, parts[r].shape[0], stride)
        for c in ctxs:
            c.synchronize()
    want = np.sort(fr, axis=0)[fr.shape[0] // 2]
    for r in range(world):
        assert np.array_equal(ctxs[r].copy_to_host(jobs[r].result_ptr(), nelem), want), ("sharded", r)
    for j in jobs:
        j.close()
    for r in range(world):
        cudart.cudaFree(dptrs[r])
        ctxs[r].close()
    print("sharded median ok", flush=True)
    # highlight: fused and per-pixel paths on adversarial and random frames, components with labels
    import hl_cases
    from oracle import highlight_oracle as ho

    cases = [c[1:] for c in hl_cases.adversarial_cases()[:10]] + [hl_cases.random_case(t) for t in (0, 5, 11)]
    for frame, p in cases:
        want = ho.highlight_objects(frame.copy(), p)
        for path in (0, 1):
            ctx.highlight_begin(p.background, np.ascontiguousarray(p.struct_element), p.threshold, p.threshold_lo,
                                p.threshold_hi, p.min_size_hyst, p.min_size_threshold, p.width_border)
            ctx.highlight_set_path(path)
            got = ctx.highlight_frames(frame[None])[0]
            if path == 0:
                m, comps, ncomps, lab = ctx.highlight_frames_cc(frame[None], max_comps=64, labels=True)
                assert np.array_equal(m[0], want) and np.array_equal(lab[0] != 0, want != 0)
            ctx.highlight_end()
            assert np.array_equal(got, want), ("highlight", path)
    print("highlight ok", flush=True)
    # frame source and the asynchronous queue: every mode, odd crops, a misaligned segment end; both highlight builds
    import frame_cases
    from oracle import frames_oracle as fo

    for name, frames, crop, mode in frame_cases.cases()[:20]:
        fmt = _cabi.FrameFormat.of(frames.shape[1:], mode, crop)
        assert np.array_equal(ctx.frames_prepare(frames, fmt), fo.prepare_frames(frames, crop, mode)), name
    frame, p = hl_cases.random_case(3)
    colour = np.stack([frame, frame, frame], axis=-1)[None].repeat(5, axis=0)
    want = ho.highlight_objects(frame.copy(), p)
    for variant in ("small", "large"):
        os.environ["CVVP_HL_VARIANT"] = variant
        ctx.highlight_begin(p.background, np.ascontiguousarray(p.struct_element), p.threshold, p.threshold_lo,
                            p.threshold_hi, p.min_size_hyst, p.min_size_threshold, p.width_border)
        ctx.highlight_queue_begin(2, 2, _cabi.FrameFormat.of(colour.shape[1:], _cabi.FRAMES_CHANNEL0), 16)
        got = []
        for i in range(0, 5, 2):
            if ctx.highlight_queue_pending() == 2:
                got.append(ctx.highlight_next()[0].copy())
            ctx.highlight_submit(colour[i:i + 2])
        while ctx.highlight_queue_pending():
            got.append(ctx.highlight_next()[0].copy())
        ctx.highlight_end()
        got = np.concatenate(got)
        assert got.shape[0] == 5 and all(np.array_equal(g, want) for g in got), ("queue", variant)
    os.environ.pop("CVVP_HL_VARIANT")
    print("frame source and queue ok", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
