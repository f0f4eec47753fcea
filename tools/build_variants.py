#!/usr/bin/env python
"""Build A/B variants of the blend kernels: csrc/blend.cu compiled with different -D knobs, each linked with the
other (unchanged) objects of the in-tree build into gflow_b200/_lib/variants/libgfb_<name>.so.  tools/ab_blend.py
times them side by side on the GPU box in ONE gpurun call.

    python tools/build_variants.py            # all variants of VARIANTS
    python tools/build_variants.py name ...   # only these
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gflow_b200 import _build  # noqa: E402

VARIANTS = {
    "base": [],
    "cta12": ["-DGFB_BLEND_MIN_CTAS=12"],
    "list": ["-DGFB_BLEND_LIST=1"],
    "aeff": ["-DGFB_FWD_AEFF=1"],
    "arith1": ["-DGFB_BFLY_ARITH=1"],
    "arith2": ["-DGFB_BFLY_ARITH=2"],
    "trace": ["-DGFB_BLEND_TRACE=1"],
    "pair": ["-DGFB_BLEND_PAIR=1"],
    "pair8": ["-DGFB_BLEND_PAIR=1", "-DGFB_BLEND_MIN_CTAS=8"],
    "pair6": ["-DGFB_BLEND_PAIR=1", "-DGFB_BLEND_MIN_CTAS=6"],
    "signt": ["-DGFB_FWD_SIGNT=1"],
    "pair8_signt": ["-DGFB_BLEND_PAIR=1", "-DGFB_BLEND_MIN_CTAS=8", "-DGFB_FWD_SIGNT=1"],
    "pair_trace": ["-DGFB_BLEND_PAIR=1", "-DGFB_BLEND_MIN_CTAS=8", "-DGFB_BLEND_TRACE=1"],
}
OUT = os.path.join(_build.LIB_DIR, "variants")


def main():
    names = sys.argv[1:] or list(VARIANTS)
    _build.build()
    os.makedirs(OUT, exist_ok=True)
    nvcc = _build._nvcc()
    host = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    others = [os.path.join(_build.LIB_DIR, s.replace(".cu", ".o")) for s in _build.SOURCES if s != "blend.cu"]
    for name in names:
        obj = os.path.join(OUT, f"blend_{name}.o")
        cmd = [nvcc, *_build.ARCH_FLAGS, *_build.COMMON_FLAGS, *host, *VARIANTS[name], "-c",
               os.path.join(_build.CSRC, "blend.cu"), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise SystemExit(f"variant {name}: nvcc failed\n{res.stdout}{res.stderr}")
        regs = [ln for ln in (res.stdout + res.stderr).splitlines() if "registers" in ln]
        lib = os.path.join(OUT, f"libgfb_{name}.so")
        res2 = subprocess.run([nvcc, *_build.ARCH_FLAGS, *host, "-shared", "-o", lib, obj, *others], capture_output=True,
                              text=True)
        if res2.returncode != 0:
            raise SystemExit(f"variant {name}: link failed\n{res2.stdout}{res2.stderr}")
        os.remove(obj)
        spills = sum("spill" in ln and " 0 bytes spill stores" not in ln for ln in (res.stdout + res.stderr).splitlines())
        print(f"{name:12s} {' '.join(VARIANTS[name]) or '(defaults)':50s} kernels with spills: {spills}")


if __name__ == "__main__":
    main()
