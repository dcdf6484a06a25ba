"""TEST INFRASTRUCTURE ONLY -- loader of oracle/_ref/cvvp_background_ref*.so: the reference's own GetVideoBackground
entry point with everything behind it (Sources/cv_vid_bg_helpers.cpp, Sources/Utility/cv_util.cpp, the AsyncTokens
generator / worker threads, CvVidFramesGeneratorAlgo, CvVidFragmentConsumer, HistogramMedianAlgo8/16/32), compiled
UNMODIFIED from /root/reference against oracle/shim_cv2 (oracle/background_ref_driver.cpp, Makefile target ref_background).

Only tests/ import this module.  It pins what no single class does: the crop rule and its quirk, frame_limit, the
bin-width dispatch, the per-generator frame ranges, the strip split and re-assembly, batch sizes from max_threads.

The reference's worker threads reach OpenCV through Python here.  With a Python thread state created and deleted around
every call (what pybind11's gil_scoped_acquire does on threads Python did not start) the pipeline stalled now and then
on a loaded machine; the shim now keeps one thread state per reference thread (shim::Gil) and 600 calls under heavy
load completed.  `run_isolated` stays as a guard: it runs the calls in a child process with a progress watchdog and
starts unfinished ones again, so that a stall of the reference's pipeline (token_storage_limit below its generator
count does stall it, deterministically) costs a retry instead of hanging the test run."""
from __future__ import annotations

import importlib.util
import json
import subprocess
import sys
import sysconfig
import tempfile
from pathlib import Path

import numpy as np

_REF_DIR = Path(__file__).resolve().parent / "_ref"
_mod = None


def path() -> Path:
    return _REF_DIR / ("cvvp_background_ref" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def available() -> bool:
    return path().exists()


def load():
    global _mod
    if _mod is None:
        if not available():
            raise FileNotFoundError(f"{path()} is missing: run `make -C oracle ref_background` where /root/reference is mounted")
        spec = importlib.util.spec_from_file_location("cvvp_background_ref", path())
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def get_video_background(vid_path, **pack):
    """In this process: GetVideoBackground(VidBgPack{vid_path, bg_algo="hist", max_threads=-1, frame_limit=-1, grayscale,
    vid_is_grayscale, crop_x, crop_y, crop_width, crop_height, token_storage_limit=-1, print_timing_report}) -> ndarray
    or None"""
    return load().GetVideoBackground(str(vid_path), **pack)


_MARK = "@@cvvp-background-job@@"


def _worker(job_file: str, out_dir: str) -> None:
    jobs = json.loads(Path(job_file).read_text())
    mod = load()
    for idx, vid_path, pack in jobs:
        for stream in (sys.stdout, sys.stderr):
            print(f"{_MARK} {idx}", file=stream, flush=True)
        bg = mod.GetVideoBackground(vid_path, **pack)
        sys.stdout.flush()
        np.save(Path(out_dir) / f"{idx}.npy", np.zeros(0, np.uint8) if bg is None else bg)


def run_isolated(jobs, stall_timeout: float = 15.0, attempts: int = 10):
    """jobs: [(vid_path, pack dict), ...] -> [(ndarray or None, stdout text, stderr text), ...], each call made in a child
    process; a child that finishes no call for `stall_timeout` seconds is killed and the calls it had not finished are
    started again (at most `attempts` children)."""
    import time

    todo = [(i, str(v), dict(p)) for i, (v, p) in enumerate(jobs)]
    results: dict[int, tuple] = {}
    with tempfile.TemporaryDirectory() as tmp:
        for attempt in range(attempts):
            if not todo:
                break
            job_file = Path(tmp) / "jobs.json"
            job_file.write_text(json.dumps(todo))
            out_f, err_f = Path(tmp) / f"out{attempt}.txt", Path(tmp) / f"err{attempt}.txt"
            cmd = [sys.executable, "-c", "import sys; sys.path.insert(0, sys.argv[1]); from oracle import background_ref as b; "
                   "b._worker(sys.argv[2], sys.argv[3])", str(Path(__file__).resolve().parent.parent), str(job_file), tmp]
            with open(out_f, "w") as fo_, open(err_f, "w") as fe_:
                proc = subprocess.Popen(cmd, stdout=fo_, stderr=fe_)
                done, last = 0, time.monotonic()
                while proc.poll() is None:
                    time.sleep(0.05)
                    n = sum((Path(tmp) / f"{i}.npy").exists() for i, _, _ in todo)
                    if n != done:
                        done, last = n, time.monotonic()
                    elif time.monotonic() - last > stall_timeout + (30.0 if done == 0 else 0.0):  # first call: imports
                        proc.kill()
                        proc.wait()
                        break
            texts = {}
            for name, text in (("out", out_f.read_text()), ("err", err_f.read_text())):
                for part in text.split(_MARK)[1:]:
                    head, _, body = part.partition("\n")
                    texts.setdefault(int(head), {})[name] = body
            for i, _, _ in list(todo):
                f = Path(tmp) / f"{i}.npy"
                if f.exists():
                    try:
                        a = np.load(f)
                    except Exception:  # killed while writing
                        f.unlink()
                        continue
                    results[i] = (None if a.size == 0 else a, texts.get(i, {}).get("out", ""), texts.get(i, {}).get("err", ""))
            todo = [j for j in todo if j[0] not in results]
    if todo:
        raise TimeoutError(f"the reference's GetVideoBackground stalled {attempts} times on {len(todo)} of {len(jobs)} calls")
    return [results[i] for i in range(len(jobs))]
