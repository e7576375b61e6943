"""In-tree build of csrc/liblcr.so (hand-written CUDA for sm_100a behind the C-ABI of include/lcr.h).

    python livecell-instance-segmentation_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box
with the repo snapshot.  No torch headers are involved: the library is plain CUDA runtime + C-ABI.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
OBJ_DIR = os.path.join(CSRC, "_obj")
LIB_PATH = os.path.join(CSRC, "liblcr.so")

SOURCES = ["api.cu", "boxes.cu", "layout.cu", "paste.cu", "nms.cu", "roi_align.cu", "rpn_select.cu", "match.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: liblcr.so cannot be built (there is no CPU fallback)")
    return exe


EXT_SRC = os.path.join(CSRC, "torch_ext.cpp")
EXT_PATH = os.path.join(CSRC, "lcr_torch.so")


def ext_is_stale() -> bool:
    if not os.path.exists(EXT_PATH):
        return True
    newest = max(os.path.getmtime(EXT_SRC), os.path.getmtime(os.path.join(INCLUDE, "lcr.h")))
    return os.path.getmtime(EXT_PATH) < newest


def build_torch_ext(force: bool = False, verbose: bool = False) -> str:
    """g++ -> csrc/lcr_torch.so: the thin torch extension (C++ autograd node of RoIAlign, nms) over liblcr.so.
    Plain g++ against the installed torch headers; links liblcr.so through an $ORIGIN rpath, in-tree."""
    if not force and not ext_is_stale():
        return EXT_PATH
    import fcntl
    import sysconfig
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not ext_is_stale():
            return EXT_PATH
        return _build_torch_ext_locked(verbose, sysconfig)


def _build_torch_ext_locked(verbose, sysconfig) -> str:

    import torch
    tdir = os.path.dirname(torch.__file__)
    tlib = os.path.join(tdir, "lib")
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    defs = ["-DTORCH_EXTENSION_NAME=lcr_torch", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    for name in ("COMPILER_TYPE", "STDLIB", "BUILD_ABI"):
        val = getattr(torch._C, f"_PYBIND11_{name}", None)
        if val is not None:
            defs.append(f'-DPYBIND11_{name}="{val}"')
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-deprecated-declarations", *defs,
           "-I", INCLUDE, "-isystem", os.path.join(tdir, "include"),
           "-isystem", os.path.join(tdir, "include", "torch", "csrc", "api", "include"),
           "-isystem", os.path.join(cuda_home, "include"), "-isystem", sysconfig.get_paths()["include"],
           EXT_SRC, "-o", EXT_PATH + ".tmp",
           "-L", tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
           "-L", CSRC, "-l:liblcr.so", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"g++ failed for torch_ext.cpp:\n{res.stdout}\n{res.stderr}")
    os.replace(EXT_PATH + ".tmp", EXT_PATH)
    return EXT_PATH


def _newest_input() -> float:
    paths = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, "common.cuh"), os.path.join(INCLUDE, "lcr.h")]
    return max(os.path.getmtime(p) for p in paths)


def is_stale() -> bool:
    return not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < _newest_input()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu to an object (in parallel) and link csrc/liblcr.so.  Returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    import fcntl
    lock = open(os.path.join(CSRC, ".build.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)       # ranks started together (torchrun) build once, the others wait and find it fresh
    try:
        if not force and not is_stale():
            return LIB_PATH
        return _build_locked(nvcc, force, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(nvcc: str, force: bool, verbose: bool) -> str:
    header_time = max(os.path.getmtime(os.path.join(CSRC, "common.cuh")), os.path.getmtime(os.path.join(INCLUDE, "lcr.h")))

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        src_path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src_path), header_time):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", src_path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose and res.stderr:
            print(res.stderr, file=sys.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
    print(build_torch_ext(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
