#!/usr/bin/env python
"""SHA-1 of the SASS instruction stream of every kernel in libgflow_b200.so's objects.

    python tools/sass_hashes.py                 # print
    python tools/sass_hashes.py --write         # refresh profiles/sass_hashes_measured.json (after a kernel has been
                                                #  changed on purpose AND re-measured on hardware)

profiles/sass_hashes_measured.json lists the kernels whose timings in profiles/ and DESIGN.md were measured; a test
(tests/test_capi_library.py) fails when one of them no longer compiles to the same instructions, so code added without
GPU access cannot silently change a measured kernel."""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(ROOT, "gflow_b200", "_lib")
OUT = os.path.join(ROOT, "profiles", "sass_hashes_measured.json")
CUOBJDUMP = "/usr/local/cuda/bin/cuobjdump"


def kernel_hashes():
    out = {}
    for obj in sorted(f for f in os.listdir(LIB_DIR) if f.endswith(".o")):
        txt = subprocess.run([CUOBJDUMP, "-sass", os.path.join(LIB_DIR, obj)], capture_output=True, text=True).stdout
        cur, lines = None, {}
        for ln in txt.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                cur = m.group(1)
                lines[cur] = []
                continue
            if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
                lines[cur].append(re.sub(r"\s*/\*.*?\*/\s*", "", ln.split("*/", 1)[1]).strip())
        for mangled, ins in lines.items():
            name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "")
            out[f"{obj[:-2]}:{name}"] = {"instructions": len(ins), "sha1": hashlib.sha1("\n".join(ins).encode()).hexdigest()}
    return out


if __name__ == "__main__":
    h = kernel_hashes()
    if "--write" in sys.argv:
        keep = {k: v for k, v in h.items() if k.split(":")[0] in ("geometry", "binning", "blend", "pipeline")
                and "true" not in k.split("<")[-1]}  # experimental instantiations have never been measured
        with open(OUT, "w") as fh:
            json.dump(keep, fh, indent=1, sort_keys=True)
        print(f"wrote {len(keep)} kernels to {OUT}")
    else:
        print(json.dumps(h, indent=1, sort_keys=True))
