"""TEST-ONLY: compile gflow_b200/csrc/*.cu for the host against the SIMT shim (simt_emu.h).

The product sources are used unmodified; this script rewrites a private copy under
tests/simt/_build/src/:

  * `kernel<<<grid, block, smem, stream>>>(args)`  ->  gfb_emu::launch(dim3(grid), dim3(block), [&]{ kernel(args); })
  * the bodies of the inline-PTX wrappers listed in PTX_WRAPPERS are replaced by their emulated
    meaning (mbarrier / cp.async.bulk bookkeeping, exp2 / reciprocal, no-ops for PDL);

and then builds tests/simt/_build/libgflow_b200_emu.so with g++ (-ffp-contract=off, matching nvcc
-fmad=false for the translation units whose results are bit-compared).  The result exports the same C
ABI as libgflow_b200.so with host pointers in place of device pointers.  Nothing under gflow_b200/
imports this module: it is a checker for the kernel sources, not a fallback.
"""
from __future__ import annotations

import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gflow_b200", "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD = os.path.join(HERE, "_build")
SRC_OUT = os.path.join(BUILD, "src")
LIB_PATH = os.path.join(BUILD, "libgflow_b200_emu.so")
_STAMP = os.path.join(BUILD, "build.stamp")

# wrapper name -> emulated body
PTX_WRAPPERS = {
    "smem_u32": "return 0u;",
    "mbar_init": "gfb_emu::mbar_init(bar, count);",
    "fence_mbar_init": "",
    "fence_proxy_async": "",
    "mbar_expect_tx": "gfb_emu::mbar_expect_tx(bar, bytes);",
    "mbar_arrive": "gfb_emu::mbar_arrive(bar);",
    "bulk_g2s": "gfb_emu::bulk_g2s(dst, src, bytes, bar);",
    "mbar_wait": "gfb_emu::mbar_wait(bar, parity);",
    "splat_ex2": "return exp2f(x);",
    "trace_now": "return 0ull;",
    "trace_smid": "return 0u;",
    "fma2": "return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));",
    "mul2": "return make_float2(a.x * b.x, a.y * b.y);",
    "add2": "return make_float2(a.x + b.x, a.y + b.y);",
    "sub2": "return make_float2(a.x - b.x, a.y - b.y);",
    "splat_rcp": "return 1.0f / x;",
    "gfb_pdl_wait": "",
    "gfb_pdl_launch_dependents": "",
    "red_add_s32": "++gfb_emu_counts[1]; *p += v;",
    "red_add_f32": "++gfb_emu_counts[1]; *p += v;",
}


def _skip_literal(s: str, i: int) -> int:
    """s[i] opens a string / char literal or a comment: return the index just past it, else i."""
    if s.startswith("//", i):
        j = s.find("\n", i)
        return len(s) if j < 0 else j
    if s.startswith("/*", i):
        j = s.find("*/", i + 2)
        return len(s) if j < 0 else j + 2
    if s.startswith('R"(', i):
        j = s.find(')"', i + 3)
        return len(s) if j < 0 else j + 2
    if s[i] in "\"'":
        q = s[i]
        j = i + 1
        while j < len(s) and s[j] != q:
            j += 2 if s[j] == "\\" else 1
        return j + 1
    return i


def _match(s: str, i: int, open_c: str, close_c: str) -> int:
    """s[i] == open_c: index of the matching close_c (literals and comments skipped)."""
    assert s[i] == open_c, (s[i - 20:i + 20], open_c)
    depth = 0
    while i < len(s):
        j = _skip_literal(s, i)
        if j != i:
            i = j
            continue
        if s[i] == open_c:
            depth += 1
        elif s[i] == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced " + open_c)


def replace_wrapper_bodies(text: str, found: set) -> str:
    for name, body in PTX_WRAPPERS.items():
        pat = re.compile(r"__device__\s+__forceinline__\s+[\w:]+\s*\*?\s+" + name + r"\s*\(")
        pos = 0
        while True:
            m = pat.search(text, pos)
            if not m:
                break
            close = _match(text, m.end() - 1, "(", ")")
            brace = text.find("{", close)
            semi = text.find(";", close)
            if brace < 0 or (0 <= semi < brace):  # a declaration, not a definition
                pos = close
                continue
            end = _match(text, brace, "{", "}")
            new = "{ " + body + " }"
            text = text[:brace] + new + text[end + 1:]
            pos = brace + len(new)
            found.add(name)
    return text


def _split_top(s: str) -> list:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(text: str) -> str:
    pat = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;()]*>)?)\s*<<<")
    while True:
        m = pat.search(text)
        if not m:
            return text
        cfg_start = m.end()
        cfg_end = text.index(">>>", cfg_start)
        cfg = _split_top(text[cfg_start:cfg_end])
        assert 2 <= len(cfg) <= 4, cfg
        paren = cfg_end + 3
        while text[paren].isspace():
            paren += 1
        close = _match(text, paren, "(", ")")
        args = text[paren + 1:close]
        repl = f"gfb_emu::launch(dim3({cfg[0]}), dim3({cfg[1]}), [&]() {{ {m.group(1)}({args}); }})"
        text = text[:m.start()] + repl + text[close + 1:]


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files += [os.path.join(INCLUDE, "gflow_b200.h"), os.path.join(HERE, "simt_emu.h"), os.path.join(HERE, "simt_emu.cpp"),
              __file__]
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(os.environ.get("GFB_EMU_DEFINES", "").encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(_STAMP):
        return True
    with open(_STAMP) as fh:
        return fh.read().strip() != _fingerprint()


def build(force: bool = False, opt: str = "-O2") -> str:
    """Builds (or reuses) the emulated library.  Serialised across processes with a file lock: the world-size-2
    tests load it from two processes at once."""
    import fcntl

    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, opt)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, opt: str) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(SRC_OUT, exist_ok=True)
    for f in os.listdir(SRC_OUT):
        os.remove(os.path.join(SRC_OUT, f))
    found: set = set()
    cpps = []
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith((".cu", ".cuh")):
            continue  # torch_ext.cpp is the torch binding, not kernel code
        with open(os.path.join(CSRC, f)) as fh:
            text = fh.read()
        text = replace_wrapper_bodies(text, found)
        text = rewrite_launches(text)
        code_only = re.sub(r"//[^\n]*", "", text)
        if re.search(r"\basm\b", code_only) or "<<<" in code_only:
            raise RuntimeError(f"{f}: inline PTX or a <<<>>> launch survived the rewrite; add the wrapper to PTX_WRAPPERS")
        out = f[:-3] + ".cpp" if f.endswith(".cu") else f
        with open(os.path.join(SRC_OUT, out), "w") as fh:
            fh.write(f"// GENERATED from gflow_b200/csrc/{f} by tests/simt/build_emu.py -- do not edit\n" + text)
        if f.endswith(".cu"):
            cpps.append(os.path.join(SRC_OUT, out))
    with open(os.path.join(SRC_OUT, "cuda_runtime.h"), "w") as fh:
        fh.write('#pragma once\n#include "simt_emu.h"\n')
    cmd = ["g++", opt, "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing",
           "-Wno-unknown-pragmas", "-Wno-attributes", *os.environ.get("GFB_EMU_DEFINES", "").split(),  # kernel build knobs
           "-I", SRC_OUT, "-I", HERE, "-I", INCLUDE, *cpps, os.path.join(HERE, "simt_emu.cpp"), "-o", LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(BUILD, "build.log"), "w") as fh:
        fh.write("$ " + " ".join(cmd) + "\n" + res.stdout + res.stderr + f"\nwrappers emulated: {sorted(found)}\n")
    if res.returncode != 0:
        raise RuntimeError("simt emulation build failed:\n" + (res.stdout + res.stderr)[-6000:])
    with open(_STAMP, "w") as fh:
        fh.write(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv))


SELFCHECK_LIB = os.path.join(BUILD, "libsimt_selfcheck.so")


def build_selfcheck() -> str:
    """tests/simt/selfcheck.cu (kernels that pin the shim's own semantics) -> _build/libsimt_selfcheck.so."""
    os.makedirs(SRC_OUT, exist_ok=True)
    src = os.path.join(HERE, "selfcheck.cu")
    deps = [src, os.path.join(HERE, "simt_emu.h"), os.path.join(HERE, "simt_emu.cpp"), __file__]
    if os.path.exists(SELFCHECK_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(SELFCHECK_LIB) for d in deps):
        return SELFCHECK_LIB
    with open(src) as fh:
        text = rewrite_launches(fh.read())
    out = os.path.join(SRC_OUT, "selfcheck.cpp")
    with open(out, "w") as fh:
        fh.write(text)
    with open(os.path.join(SRC_OUT, "cuda_runtime.h"), "w") as fh:
        fh.write('#pragma once\n#include "simt_emu.h"\n')
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-attributes",
           "-I", SRC_OUT, "-I", HERE, out, os.path.join(HERE, "simt_emu.cpp"), "-o", SELFCHECK_LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("selfcheck build failed:\n" + (res.stdout + res.stderr)[-4000:])
    return SELFCHECK_LIB


SANITIZED_LIB = os.path.join(BUILD, "libgflow_b200_emu_san.so")


def build_sanitized() -> str:
    """The same generated sources with -fsanitize=address,alignment,bounds (aborting):
      * alignment: a float4 / float2 / ushort4 access through a pointer that is not 16 / 8 byte aligned faults on the
        GPU but is silently tolerated by x86;
      * address: an emulated kernel that writes or reads past a caller's buffer (host tensors get red zones once
        libasan is preloaded) -- the CPU counterpart of compute-sanitizer memcheck;
      * bounds: indexing past a static (shared-memory) array.
    Loading it needs libasan + libubsan preloaded (sanitizer_preload(); see tests/test_simt_kernels.py)."""
    build()  # generated sources are current after this
    if os.path.exists(SANITIZED_LIB) and os.path.getmtime(SANITIZED_LIB) >= os.path.getmtime(LIB_PATH):
        return SANITIZED_LIB
    cpps = [os.path.join(SRC_OUT, f) for f in sorted(os.listdir(SRC_OUT)) if f.endswith(".cpp") and f != "selfcheck.cpp"]
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing",
           "-fsanitize=address,alignment,bounds", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer", "-Wno-unknown-pragmas",
           "-Wno-attributes", "-I", SRC_OUT, "-I", HERE, "-I", INCLUDE, *cpps, os.path.join(HERE, "simt_emu.cpp"), "-o",
           SANITIZED_LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("sanitised emulation build failed:\n" + (res.stdout + res.stderr)[-4000:])
    return SANITIZED_LIB


def sanitizer_preload():
    """LD_PRELOAD value (libasan first, then libubsan) or None when the runtimes are not installed."""
    libs = []
    for name in ("libasan.so", "libubsan.so"):
        path = subprocess.run(["gcc", "-print-file-name=" + name], capture_output=True, text=True).stdout.strip()
        if not os.path.isabs(path) or not os.path.exists(path):
            return None
        libs.append(path)
    return ":".join(libs)
