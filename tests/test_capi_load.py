"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/kpl.h declares; without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "kpl.h")).read()
    return sorted(set(re.findall(r"KPL_API\s+[\w\s\*]+?\b(kpl_\w+)\s*\(", txt)))


def test_header_symbols_are_exported(kpl):
    lib = kpl.load_library()
    names = declared_symbols()
    assert len(names) >= 20, names
    for n in names:
        assert hasattr(lib, n), "libkpl_b200.so does not export %s" % n
    from keypoint_learning_b200.capi import EXPORTS
    assert sorted(EXPORTS) == names          # the ctypes binding covers exactly the header


def test_struct_layouts_match_header(kpl):
    """ctypes mirrors of kpl_params / kpl_timings / kpl_stats must have the C layout (compiled probe)."""
    import subprocess
    import tempfile
    from keypoint_learning_b200 import capi
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "kpl.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(kpl_params), sizeof(kpl_timings), sizeof(kpl_stats),' \
          ' offsetof(kpl_params, grid_origin), offsetof(kpl_stats, grid_origin), offsetof(kpl_stats, n_unscored));' \
          'printf("%zu %zu %zu\\n", offsetof(kpl_params, slab_guard_cells), offsetof(kpl_stats, n_fragile_points), offsetof(kpl_stats, n_views));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "p")
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).split()
    got = [int(x) for x in out]
    exp = [C.sizeof(capi.KplParams), C.sizeof(capi.KplTimings), C.sizeof(capi.KplStats),
           capi.KplParams.grid_origin.offset, capi.KplStats.grid_origin.offset, capi.KplStats.n_unscored.offset,
           capi.KplParams.slab_guard_cells.offset, capi.KplStats.n_fragile_points.offset, capi.KplStats.n_views.offset]
    assert got == exp


def test_version_and_defaults_without_gpu(kpl):
    lib = kpl.load_library()
    assert b"sm_100a" in lib.kpl_version()
    from keypoint_learning_b200.capi import KplParams
    p = KplParams()
    assert lib.kpl_params_default(C.byref(p)) == 0
    # TestDetector defaults (src/main_test_detector.cpp:65-67,105-106,167)
    assert p.radius_features == 20.0 and p.radius_nms == 4.0 and p.threshold == float(np.float32(0.85))
    assert (p.n_annulus, p.n_bins, p.k_normals, p.non_maxima, p.draws_remove) == (5, 10, 10, 1, 0)


def test_no_cpu_fallback(kpl):
    """Without an sm_100 device kpl_create must fail with KPL_E_CUDA: the product never computes on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    lib = kpl.load_library()
    h = C.c_void_p()
    assert lib.kpl_create(0, C.byref(h)) == 6 and not h.value
    with pytest.raises(kpl.KplError):
        kpl.KeypointLearningDetector()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under keypoint_learning_b200/ may reference it."""
    pkg = os.path.join(ROOT, "keypoint_learning_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "kplo_" not in txt and "libkpl_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
