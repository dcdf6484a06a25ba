"""Parity at the FULL sizes of BASELINE.json's configs, as SURVEY.md 8(d) "Parity check" specifies it:

    C2 (1920x1080x1000 median)    every pixel against the reference's own class (oracle/_ref) -- or its C restatement where
                                  the reference was never built -- run over all host threads
    C3 (1080p highlight, 10k)     the first 256 frames + every 97th frame of the 10 000 against the cv2 restatement
    C4 (512x256 highlight, 200k)  the first 256 frames + every 97th frame of the 200 000 likewise
    C5 (3840x2160x5000 median)    every pixel of a 2160-row stack of 5000 frames is 41 GB of host memory; the full-size
                                  check is by columns: 5000 frames of 64 full rows against the oracle, and the long-stack
                                  path (window counting + gated two-pass) against the two passes alone on the whole image
                                  (device vs device)

Inputs are generated on the device (csrc/synth.cu, bit-identical to synth.py) and fetched for the CPU side."""
import concurrent.futures as cf
import os

import numpy as np
import pytest
import torch

from cvvidproc_b200 import synth
from oracle import highlight_oracle as ho

pytestmark = pytest.mark.gpu


def _device_frames(ctx, cfg, first, n, width=None, height=None, row0=0, nrows=None):
    p = synth.CONFIG_PARAMS[cfg]
    w, h = width or p["width"], height or p["height"]
    nrows = nrows or h
    t = torch.empty((n, nrows * w), dtype=torch.uint8, device="cuda:0")
    ctx.synth_frames_device(t.data_ptr(), nrows * w, w, h, first, n, p["seed"], p["ndisks"], row0=row0, nrows=nrows)
    ctx.synchronize()
    return t, w, nrows


def test_c2_every_pixel_against_the_reference_class(gpu_ctx, ref_median, oracle_median):
    stack, w, h = _device_frames(gpu_ctx, "C2", 0, 1000)
    out = torch.empty(w * h, dtype=torch.uint8, device="cuda:0")
    gpu_ctx.median_device(stack.data_ptr(), 1000, w * h, w * h, out.data_ptr())
    gpu_ctx.synchronize()
    host = stack.cpu().numpy().reshape(1000, h, w)
    cpu = ref_median or oracle_median
    want = cpu(host, nthreads=os.cpu_count() or 4)
    assert np.array_equal(out.cpu().numpy().reshape(h, w), want)


def _highlight_sample(ctx, cfg, total_frames, batch):
    """first 256 frames + every 97th frame of the config's stream, masks from the device vs the cv2 restatement"""
    p_ = synth.CONFIG_PARAMS[cfg]
    w, h = p_["width"], p_["height"]
    bgstack, _, _ = _device_frames(ctx, cfg, 0, 255)
    bg_d = torch.empty(w * h, dtype=torch.uint8, device="cuda:0")
    ctx.median_device(bgstack.data_ptr(), 255, w * h, w * h, bg_d.data_ptr())
    ctx.synchronize()
    bg = bg_d.cpu().numpy().reshape(h, w)
    del bgstack
    p = ho.canonical_params(bg)
    picks = list(range(256)) + list(range(291, total_frames, 97))
    ctx.highlight_begin(p.background, np.ascontiguousarray(p.struct_element), p.threshold, p.threshold_lo, p.threshold_hi,
                        p.min_size_hyst, p.min_size_threshold, p.width_border)
    bad = []
    try:
        import cv2

        cv2.setNumThreads(1)
        with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            for i0 in range(0, len(picks), batch):
                idx = picks[i0:i0 + batch]
                frames = torch.empty((len(idx), w * h), dtype=torch.uint8, device="cuda:0")
                for j, f in enumerate(idx):  # runs of consecutive frames are generated together
                    if j == 0 or idx[j - 1] != f - 1:
                        run = 1
                        while j + run < len(idx) and idx[j + run] == f + run:
                            run += 1
                        ctx.synth_frames_device(frames[j:].data_ptr(), w * h, w, h, f, run, p_["seed"], p_["ndisks"])
                masks = torch.empty_like(frames)
                ctx.highlight_device(frames.data_ptr(), len(idx), w * h, masks.data_ptr(), w * h)
                ctx.synchronize()
                hf = frames.cpu().numpy().reshape(len(idx), h, w)
                hm = masks.cpu().numpy().reshape(len(idx), h, w)
                want = list(ex.map(lambda k: ho.highlight_objects(hf[k].copy(), p), range(len(idx))))
                bad += [idx[k] for k in range(len(idx)) if not np.array_equal(hm[k], want[k])]
    finally:
        ctx.highlight_end()
    return picks, bad


def test_c3_first_256_and_every_97th_frame_against_cv2(gpu_ctx):
    picks, bad = _highlight_sample(gpu_ctx, "C3", 10_000, 64)
    assert len(picks) == 256 + 101
    assert not bad, f"frames that differ from the oracle: {bad[:20]}"


def test_c4_first_256_and_every_97th_frame_against_cv2(gpu_ctx):
    picks, bad = _highlight_sample(gpu_ctx, "C4", 200_000, 512)
    assert len(picks) == 256 + 2059
    assert not bad, f"frames that differ from the oracle: {bad[:20]}"


def test_c5_long_stack_columns_and_both_device_forms(gpu_ctx, oracle_median, monkeypatch):
    p_ = synth.CONFIG_PARAMS["C5"]
    w, h, n = p_["width"], p_["height"], p_["nframes"]
    # (a) 64 full rows x all 5000 frames against the oracle
    rows = 64
    band, _, _ = _device_frames(gpu_ctx, "C5", 0, n, row0=1000, nrows=rows)
    out = torch.empty(rows * w, dtype=torch.uint8, device="cuda:0")
    gpu_ctx.median_device(band.data_ptr(), n, rows * w, rows * w, out.data_ptr())
    gpu_ctx.synchronize()
    want = oracle_median(band.cpu().numpy().reshape(n, rows, w), nthreads=os.cpu_count() or 4)
    assert np.array_equal(out.cpu().numpy().reshape(rows, w), want)
    del band
    # (b) the whole 3840x2160 image over 1250 frames (10 GB): window counting + gated fallback vs the two passes alone
    m = 1250
    stack, _, _ = _device_frames(gpu_ctx, "C5", 0, m)
    a = torch.empty(w * h, dtype=torch.uint8, device="cuda:0")
    b = torch.empty_like(a)
    c = torch.empty_like(a)
    monkeypatch.setenv("CVVP_MEDIAN_TWO_PASS", "0")
    gpu_ctx.median_device(stack.data_ptr(), m, w * h, w * h, a.data_ptr())  # on-chip select at 64-byte tiles
    monkeypatch.setenv("CVVP_MEDIAN_TWO_PASS", "1")
    gpu_ctx.median_device(stack.data_ptr(), m, w * h, w * h, b.data_ptr())  # long-stack path, window form
    monkeypatch.setenv("CVVP_MEDIAN_WINDOW", "0")
    gpu_ctx.median_device(stack.data_ptr(), m, w * h, w * h, c.data_ptr())  # long-stack path, two passes only
    gpu_ctx.synchronize()
    assert torch.equal(a, b) and torch.equal(a, c)
