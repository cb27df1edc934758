"""The C-ABI libraries load without a GPU and export every symbol the headers declare."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, text)))


def test_cuda_library_exports_header(pkg):
    lib = pkg.cuda_api.load_library()
    names = declared("dymu_cuda.h", "dymu_")
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), n
    assert set(pkg.cuda_api.CUDA_SYMBOLS) == set(names)


def test_planner_library_exports_header(pkg):
    plib = pkg.planner_lib()
    assert plib.impl == "b200"
    names = declared("dymu_planner_c.h", "dymu_planner_")
    for n in names:
        assert hasattr(plib.lib, n), n
    assert set(pkg.planner_api.PLANNER_SYMBOLS) == set(names)


def test_reference_library_exports_same_api(ref_lib, pkg):
    assert ref_lib.impl == "reference"
    for n in pkg.planner_api.PLANNER_SYMBOLS:
        assert hasattr(ref_lib.lib, n), n


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    import subprocess, sys
    code = ("import os,sys; os.environ['CUDA_VISIBLE_DEVICES']=''; sys.path.insert(0, %r); "
            "import dymu_b200; p=dymu_b200.load(); pl=p.DyMuPathPlanner(1.0,1.5,2.0,1); "
            "print('INIT', pl.initGlobalLayer(1.0,0.1,32,32))" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "INIT False" in out


def test_product_does_not_touch_oracle():
    """Nothing under the package may import, link or execute oracle/."""
    pkg_dir = os.path.join(ROOT, "planning-path_planning_b200")
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "import oracle" not in text and "dymu_oracle" not in text, f
                assert "libdymu_ref" not in text and "orc_" not in text, f
