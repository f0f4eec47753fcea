"""In-tree build of libgflow_b200.so (hand-written CUDA for sm_100a, plain nvcc, no torch headers).

The library is the C-ABI drop-in boundary declared in include/gflow_b200.h.  It is built
into gflow_b200/_lib/ so it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
INCLUDE = os.path.join(_ROOT, "include")
LIB_DIR = os.path.join(_PKG, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libgflow_b200.so")
_STAMP = os.path.join(LIB_DIR, "build.stamp")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", INCLUDE, "-I", CSRC]
# geometry.cu must not fuse multiply-adds: its float32 results are bit-compared with the oracle
PER_FILE_FLAGS = {"geometry.cu": ["-fmad=false"], "binning.cu": ["-fmad=false"], "pipeline.cu": ["-fmad=false"],
                  "fit.cu": ["-fmad=false"], "densify.cu": ["-fmad=false"]}
SOURCES = ["capi.cu", "geometry.cu", "binning.cu", "blend.cu", "pipeline.cu", "fit.cu", "densify.cu", "hostpipe.cu"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libgflow_b200.so")


def _fingerprint() -> str:
    """Hash of the sources by NAME and content: the repository is copied to other roots (the GPU box), and a
    fingerprint over absolute paths made every process there rebuild the library on import."""
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(INCLUDE, "gflow_b200.h"), __file__]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode())
            h.update(fh.read())
    return h.hexdigest()


class _BuildLock:
    """Inter-process lock around stale-check + build + stamp: under torchrun every rank imports the package at once,
    and two ranks compiling into the same files handed a half-written libgflow_b200.so to dlopen."""

    def __enter__(self):
        import fcntl

        os.makedirs(LIB_DIR, exist_ok=True)
        self.fh = open(os.path.join(LIB_DIR, ".build.lock"), "w")
        fcntl.flock(self.fh, fcntl.LOCK_EX)
        return self

    def __exit__(self, *a):
        import fcntl

        fcntl.flock(self.fh, fcntl.LOCK_UN)
        self.fh.close()


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(_STAMP):
        return True
    with open(_STAMP) as fh:
        return fh.read().strip() != _fingerprint()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link libgflow_b200.so."""
    if not force and not needs_build():
        return LIB_PATH
    with _BuildLock():
        if not force and not needs_build():  # another process built it while we waited
            return LIB_PATH
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    env = dict(os.environ)
    if os.path.exists("/usr/bin/g++"):
        host = ["-ccbin", "/usr/bin/g++"]
    else:
        host = []
    objs = []
    log = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH_FLAGS, *COMMON_FLAGS, *host, *PER_FILE_FLAGS.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        log.append("$ " + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + log[-1])
        objs.append(obj)
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    cmd = [nvcc, *ARCH_FLAGS, *host, "-shared", "-o", tmp, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    log.append("$ " + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + log[-1])
    os.replace(tmp, LIB_PATH)  # atomic: a concurrent dlopen sees the old or the new file, never a partial one
    with open(os.path.join(LIB_DIR, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(_STAMP, "w") as fh:
        fh.write(_fingerprint())
    if verbose:
        print("\n".join(log))
    return LIB_PATH


# ---------------------------------------------------------------------------------------------
# Thin C++/pybind11 binding (csrc/torch_ext.cpp): host-only C++ against the torch headers, linked to
# libgflow_b200.so.  Optional accelerator for the Python operator surface: when it cannot be built
# the ctypes path in ops.py is used (same kernels either way).
EXT_NAME = "_gfb_torch"
EXT_PATH = os.path.join(LIB_DIR, EXT_NAME + ".so")
_EXT_STAMP = os.path.join(LIB_DIR, "torch_ext.stamp")


def _ext_fingerprint() -> str:
    import torch

    h = hashlib.sha256()
    for f in (os.path.join(CSRC, "torch_ext.cpp"), os.path.join(INCLUDE, "gflow_b200.h")):
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(torch.__version__.encode())
    return h.hexdigest()


def ext_needs_build() -> bool:
    if not os.path.exists(EXT_PATH) or not os.path.exists(_EXT_STAMP):
        return True
    with open(_EXT_STAMP) as fh:
        return fh.read().strip() != _ext_fingerprint()


def build_torch_ext(force: bool = False) -> str:
    """g++ csrc/torch_ext.cpp -> _lib/_gfb_torch.so (needs libgflow_b200.so to exist)."""
    if not force and not ext_needs_build():
        return EXT_PATH
    build()
    with _BuildLock():
        if not force and not ext_needs_build():
            return EXT_PATH
        return _build_torch_ext_locked()


def _build_torch_ext_locked() -> str:
    import sysconfig
    import warnings

    import torch
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from torch.utils import cpp_extension as ce
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    inc = [os.path.join(os.path.dirname(torch.__file__), "include"),
           os.path.join(os.path.dirname(torch.__file__), "include", "torch", "csrc", "api", "include"),
           "/usr/local/cuda/include", sysconfig.get_paths()["include"], INCLUDE]
    abi = getattr(torch._C, "_GLIBCXX_USE_CXX11_ABI", True)
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    extra = list(getattr(ce, "_get_pybind11_abi_build_flags", lambda: [])())
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-w", f"-D_GLIBCXX_USE_CXX11_ABI={int(bool(abi))}",
           f"-DTORCH_EXTENSION_NAME={EXT_NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H", *extra,
           *[f"-I{i}" for i in inc], os.path.join(CSRC, "torch_ext.cpp"), "-o", EXT_PATH + f".tmp{os.getpid()}",
           f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
           f"-L{LIB_DIR}", "-lgflow_b200", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(LIB_DIR, "torch_ext.log"), "w") as fh:
        fh.write("$ " + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("torch extension build failed:\n" + (res.stdout + res.stderr)[-4000:])
    os.replace(EXT_PATH + f".tmp{os.getpid()}", EXT_PATH)
    with open(_EXT_STAMP, "w") as fh:
        fh.write(_ext_fingerprint())
    return EXT_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose=True))
    print(build_torch_ext(force="--force" in sys.argv))
