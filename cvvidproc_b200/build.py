"""In-tree build of the native pieces (sm_100a only).

    libcvvp_cuda.so   CUDA kernels + C ABI (include/cvvp.h), from cvvidproc_b200/csrc/*.cu
    _core*.so         pybind11 module mirroring the reference's py_bindings.cpp (C++ host layer over the C ABI)

The shared objects are written next to this file so that they travel with the repository
snapshot to the GPU box (they are git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
import sysconfig
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libcvvp_cuda.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unknown-pragmas",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def _newer_than(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd, log_name: str, verbose: bool) -> None:
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    (PKG_DIR / "_buildlogs").mkdir(exist_ok=True)
    (PKG_DIR / "_buildlogs" / log_name).write_text(" ".join(map(str, cmd)) + "\n" + proc.stdout)
    if verbose or proc.returncode != 0:
        print(proc.stdout, file=sys.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f"build step failed: {' '.join(map(str, cmd))}")


def build_cuda_lib(force: bool = False, verbose: bool = False) -> Path:
    """Each .cu is compiled to its own object (in parallel, only when it or a header changed), then linked."""
    from concurrent.futures import ThreadPoolExecutor

    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.hpp")) + sorted(CSRC.glob("*.cuh")) + [REPO / "include" / "cvvp.h"]
    objdir = PKG_DIR / "_buildlogs" / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    todo = []
    for src in sources:
        obj = objdir / (src.stem + ".o")
        # a second build of a kernel file includes that .cu (highlight_fused_small.cu): it depends on it like on a header
        included = [CSRC / m for m in re.findall(r'#include "([^"]+\.cu)"', src.read_text())]
        if force or not _newer_than(obj, [src] + headers + included):
            todo.append((src, obj))
    if not todo and _newer_than(LIB_PATH, sources + headers):
        return LIB_PATH

    def compile_one(job):
        src, obj = job
        _run([nvcc, "-c", *NVCC_FLAGS, "-I", str(REPO / "include"), "-o", str(obj), str(src)], f"{src.stem}.log", verbose)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        list(ex.map(compile_one, todo))
    objs = [str(objdir / (src.stem + ".o")) for src in sources]
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", str(LIB_PATH), *objs],
         "libcvvp_cuda.log", verbose)
    return LIB_PATH


def core_module_path() -> Path:
    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return PKG_DIR / f"_core{suffix}"


def build_core_module(force: bool = False, verbose: bool = False) -> Path | None:
    src = CSRC / "host" / "py_core.cpp"
    if not src.exists():
        return None
    import pybind11

    out = core_module_path()
    deps = [src] + sorted((CSRC / "host").glob("*.hpp")) + [REPO / "include" / "cvvp.h"]
    if not force and _newer_than(out, deps) and _newer_than(out, [LIB_PATH]):
        return out
    # the system compiler links libstdc++ dynamically; the image's /opt/gcc wrapper (often in $CXX) links a STATIC
    # libstdc++ whose exported symbols then clash with the copy every other extension (cv2, torch) uses
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else (os.environ.get("CXX") or shutil.which("g++") or "g++")
    cmd = [
        cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall",
        "-I", str(REPO / "include"), "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
        str(src), "-o", str(out),
        "-L", str(PKG_DIR), "-lcvvp_cuda", "-Wl,-rpath,$ORIGIN", "-Wl,--exclude-libs,ALL",
    ]
    _run(cmd, "core.log", verbose)
    return out


def build_oracle(verbose: bool = False) -> None:
    """Checker only (tests / smoke / bench cpu_baseline); building it is not using it."""
    _run(["make", "-C", str(REPO / "oracle")], "oracle.log", verbose)


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda_lib(force=force, verbose=verbose)
    build_core_module(force=force, verbose=verbose)
    build_oracle(verbose=verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
