"""The C-ABI library builds for sm_100a, loads without a GPU and exports every declared symbol."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gflow_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gfb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from gflow_b200 import capi

    lib = capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gflow_b200.h but not exported"
    # and the ctypes signature table covers exactly the declared set
    assert sorted(capi.SIGNATURES) == declared
    assert lib.gfb_version() >= 100
    assert lib.gfb_build_arch() == b"sm_100a"
    assert b"bad argument" in lib.gfb_error_string(-1)


def test_library_contains_sm100a_sass_with_tma_bulk_copy():
    from gflow_b200 import _build

    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", _build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run([cuobjdump, "-sass", _build.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass, "blend kernels must stage records with cp.async.bulk (TMA)"
    assert "REDG.E.ADD.F32" in sass, "backward must use fire-and-forget float reductions"


def test_size_helpers_and_argument_validation_without_gpu():
    from gflow_b200 import capi

    lib = capi.load()
    assert lib.gfb_sort_workspace_bytes(10) == 80
    assert lib.gfb_blend_geometry_stream_bytes(10) == 320
    assert lib.gfb_blend_feature_stream_bytes(10) == 160
    assert lib.gfb_blend_grad_pack_bytes(10) == 480
    # argument validation happens before any CUDA call
    assert lib.gfb_project_point_fwd(0, 0, 0, -1, 16, 16, 0.2, 1.3, 0, 0, 0) == -1
    assert lib.gfb_compute_sh_fwd(0, 0, 0, 4, 3, 5, 0, 0) == -1
    assert lib.gfb_blend_pack_feature(0, 3, 0, 5, 0, 1, 0, 0) == -1


def test_ops_refuse_cpu_tensors():
    import gflow_b200 as g

    with pytest.raises(RuntimeError, match="CUDA tensor"):
        g.project_point(torch.zeros(4, 3), torch.zeros(4), torch.zeros(3, 4), 32, 32)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        g.compute_cov3d(torch.zeros(4, 3), torch.zeros(4, 4), None)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        g.alpha_blending(torch.zeros(4, 2), torch.zeros(4, 3), torch.zeros(4, 1), torch.zeros(4, 3),
                         torch.zeros(0, dtype=torch.int32), torch.zeros(4, 2, dtype=torch.int32), 0.0, 32, 32)


def test_dropin_module_exposes_msplat_surface():
    import gflow_b200

    gflow_b200.install_dropin()
    import msplat

    for name in ("project_point", "compute_cov3d", "ewa_project", "sort_gaussian", "alpha_blending", "compute_sh",
                 "rasterization"):
        assert callable(getattr(msplat, name))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "splat_oracle" not in txt and "c_oracle" not in txt, f"{f} references the oracle"


def test_measured_kernels_still_compile_to_the_measured_instructions():
    """profiles/sass_hashes_measured.json: SASS hash of every kernel whose timings in profiles/ and DESIGN.md were taken
    on hardware.  Work done without GPU access (new kernels, template variants, refactors) must leave them untouched;
    after an intentional, re-measured change the registry is refreshed with `python tools/sass_hashes.py --write`."""
    import importlib.util
    import json

    if not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    from gflow_b200 import _build

    if _build.needs_build():
        _build.build()
    spec = importlib.util.spec_from_file_location("sass_hashes", os.path.join(ROOT, "tools", "sass_hashes.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.kernel_hashes()
    with open(os.path.join(ROOT, "profiles", "sass_hashes_measured.json")) as fh:
        measured = json.load(fh)
    assert len(measured) >= 25
    def still_there(key, sha1):  # a template that gained a defaulted parameter is renamed, not changed: match by hash
        base = key.split("<")[0]
        return any(v["sha1"] == sha1 for k, v in now.items() if k.split("<")[0] == base)

    changed = [k for k, v in measured.items() if not still_there(k, v["sha1"])]
    assert not changed, f"measured kernels whose SASS changed (re-measure, then refresh the registry): {changed}"
