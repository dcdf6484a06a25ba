"""Development probe: the constant-memory form of the median job (csrc/median_hist.cu).  A 1080p job whose resident stack
is capped (CVVP_MEDIAN_RESIDENT_MAX) folds its frames into value histograms; reports the end-to-end rate from pinned
host memory against the same job with every frame resident, and checks the two results against each other.
    python tools/probe_median_spill.py [frames] [cap]"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

from cvvidproc_b200 import _cabi, synth


def run(ctx, host, n, nelem, chunk):
    t0 = time.perf_counter()
    ctx.median_begin(nelem, n)
    for i in range(0, n, chunk):
        ctx.median_push(host[i:i + chunk])
    out = ctx.median_finish(nelem=nelem)
    return out, time.perf_counter() - t0


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    cap = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    w, h = 1920, 1080
    nelem = w * h
    pinned = _cabi.PinnedBuffer(n * nelem)
    host = pinned.array[: n * nelem].reshape(n, h, w)
    p = synth.CONFIG_PARAMS["C2"]
    for i in range(0, n, 50):
        host[i:i + 50] = synth.synth_frames(i, min(50, n - i), w, h, p["seed"], p["ndisks"])
    ctx = _cabi.Context(0)
    os.environ.pop("CVVP_MEDIAN_RESIDENT_MAX", None)
    run(ctx, host, n, nelem, 100)
    ref, t_res = run(ctx, host, n, nelem, 100)
    os.environ["CVVP_MEDIAN_RESIDENT_MAX"] = str(cap)
    run(ctx, host, n, nelem, 100)
    got, t_sp = run(ctx, host, n, nelem, 100)
    mpx = n * nelem / 1e6
    print(f"1920x1080 x {n} frames from pinned host memory: resident {t_res * 1e3:.1f} ms ({mpx / t_res:.0f} Mpx-frames/s); "
          f"stack capped at {cap} frames {t_sp * 1e3:.1f} ms ({mpx / t_sp:.0f} Mpx-frames/s); "
          f"results equal: {np.array_equal(ref, got)}")
    ctx.close()


if __name__ == "__main__":
    main()
