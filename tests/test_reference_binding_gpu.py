"""The drop-in boundary, literally: INTEGRATION.md's reference-side bindings (integration/gpu_median_algo.h,
integration/gpu_highlight_algo.h -- TokenProcessorAlgo subclasses over the C ABI, written against cv::Mat) compiled
with the REFERENCE's own token_processor_algo.h (oracle/_ref/cvvp_binding_ref) and driven through that interface
beside the reference's own classes on the same tokens."""
import numpy as np
import pytest

import hl_cases
from cvvidproc_b200 import synth
from oracle import binding_ref as bref
from oracle import highlight_oracle as ho
from oracle import highlight_ref as href

needs_binding = pytest.mark.skipif(not bref.available(), reason="oracle/_ref/cvvp_binding_ref was not built (no /root/reference)")


@needs_binding
def test_binding_fails_loudly_without_a_device():
    """no CPU fallback behind the reference's interface either: the constructor throws the runtime_error a failed
    EXCEPTION_ASSERT would (this test runs wherever there is no GPU)"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bref.load().GpuMedianAlgo(0, 10)


@needs_binding
@pytest.mark.gpu
def test_median_binding_beside_the_reference_class(ref_median):
    """GpuMedianAlgo behind TokenProcessorAlgo: Insert per frame, NotifyNoMoreTokens, TryGetResult -- the token protocol
    of histogram_median_algo.h:66-113 -- against HistogramMedianAlgo on the same tokens"""
    if ref_median is None:
        pytest.skip("oracle/_ref/libcvvp_median_ref.so was not built")
    B = bref.load()
    for shape, n in (((48, 64), 101), ((30, 41, 3), 64), ((1, 130), 1000)):
        frames = np.random.default_rng(n).integers(0, 256, (n,) + shape, dtype=np.uint8)
        algo = B.GpuMedianAlgo(0, n)
        assert algo.insert(None) is None                      # null token: ignored (:69-70)
        for f in frames:
            assert algo.insert(f) is None                     # no result before the stream ends (:110-113)
        got = algo.finish()
        assert got.dtype == np.uint8 and got.shape == shape  # geometry of the first token (:73-77)
        assert np.array_equal(got, ref_median(frames))
        assert algo.finish() is None                          # the result was moved out; nothing pending
    # an operator that never saw a token has no result (:101-108)
    assert B.GpuMedianAlgo(0, -1).finish() is None


@needs_binding
@pytest.mark.gpu
def test_highlight_binding_beside_the_reference_class():
    """GpuHighlightAlgo and the reference's HighlightObjectsAlgo behind the same interface, same tokens"""
    if not href.available():
        pytest.skip("oracle/_ref/cvvp_highlight_ref was not built")
    B = bref.load()
    cases = [c[1:] for c in hl_cases.adversarial_cases()] + [hl_cases.random_case(t) for t in range(0, 120, 4)]
    for frame, p in cases:
        args = (p.background, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi, p.min_size_hyst,
                p.min_size_threshold, p.width_border)
        gpu = B.GpuHighlightAlgo(*args, 0)
        assert gpu.insert(None) is None
        got = gpu.insert(frame)
        assert np.array_equal(got, href.operator(p).insert(frame))
    # one operator, a stream of tokens (the worker-thread pattern), a float structuring element as getStructuringElement
    # users may pass it
    p_ = synth.CONFIG_PARAMS["C4"]
    w, h = p_["width"], p_["height"]
    stack = synth.synth_frames(0, 15, w, h, p_["seed"], p_["ndisks"])
    p = ho.canonical_params(np.sort(stack, axis=0)[7])
    gpu = B.GpuHighlightAlgo(p.background, p.struct_element.astype(np.float32), p.threshold, p.threshold_lo, p.threshold_hi,
                             p.min_size_hyst, p.min_size_threshold, p.width_border, 0)
    ref = href.operator(p)
    for f in range(100, 140):
        fr = synth.synth_frame(f, w, h, p_["seed"], p_["ndisks"])
        assert np.array_equal(gpu.insert(fr), ref.insert(fr)), f"frame {f}"
