"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol include/gplvm_b200.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import gplvmhaskell_b200 as g
from gplvmhaskell_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "gplvm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gplvm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol():
    lib = g.load_library()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libgplvm_b200.so lacks %s" % n
    # and the ctypes table covers the header exactly
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_strings():
    lib = g.load_library()
    assert b"sm_100a" in lib.gplvm_version()
    assert lib.gplvm_last_error() is not None


def test_argument_validation_happens_before_cuda():
    lib = g.load_library()
    rc = lib.gplvm_ppca_dense(None, 4, 10, 2, None, 1.0, 0, 1, 0.0, None, None, None, None)
    assert rc == _lib.ERR_ARG
    assert b"null" in lib.gplvm_last_error()
    rc = lib.gplvm_set_devices(0)
    assert rc == _lib.ERR_ARG


def test_typesafe_checks_raise_reference_messages_before_ffi():
    Y = np.zeros((4, 10))
    with pytest.raises(ValueError, match="dimensions x2 and d1 should be equal, but they are not"):
        g.make_ppca_type_safe(Y, 3, g.Left(1), (np.zeros((4, 2)), 1.0))
    with pytest.raises(ValueError, match="dimensions y1 and y2 should be equal, but they are not"):
        g.make_ppca_type_safe(Y, 3, g.Left(1), (np.zeros((5, 3)), 1.0))
    with pytest.raises(ValueError, match="Input matrix should have at least one column"):
        g.make_ppca_type_safe(np.zeros((4, 0)), 3, g.Left(1), (np.zeros((4, 3)), 1.0))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device refusal cannot be observed")
    with pytest.raises(g.GplvmError) as ei:
        g.make_ppca(np.random.default_rng(0).standard_normal((4, 10)), 2, g.Left(1), 0)
    assert ei.value.code == _lib.ERR_CUDA
    assert "no CPU fallback" in ei.value.message


def test_product_never_imports_oracle():
    """The product path must not route through the oracle: no file of the package mentions it."""
    pkg = os.path.join(ROOT, "gplvmhaskell_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_with_list_of_indexes_mirrors_reference():
    assert g.with_list_of_indexes(4, [], lambda ix: ix) == []
    assert g.with_list_of_indexes(4, [0, 3, 4], lambda ix: sum(ix)) == 7      # y < maximum is the failure test (PPCA.hs:86)
    with pytest.raises(IndexError, match="index is out of range"):
        g.with_list_of_indexes(4, [1, 5], lambda ix: ix)


def test_stop_parameter_forms():
    from gplvmhaskell_b200.ppca import _normalise_stop
    assert _normalise_stop(g.Left(7)) == ("left", 7) and _normalise_stop(3) == ("left", 3)
    assert _normalise_stop(g.Right(1e-3)) == ("right", 1e-3) and _normalise_stop(0.5) == ("right", 0.5)
    with pytest.raises(ValueError):
        _normalise_stop(("middle", 1))


def test_header_is_valid_c99_and_links(tmp_path):
    """include/gplvm_b200.h is a plain C header (what cgo / JNI / Haskell's FFI would consume) and a C program links
    against the shared library without any C++ or CUDA on the include path."""
    import subprocess
    src = tmp_path / "abi_check.c"
    src.write_text(
        '#include <stdio.h>\n#include "gplvm_b200.h"\n'
        "int main(void) {\n"
        "  double w0[2] = {1.0, 2.0}, y[4] = {0.0, 1.0, 2.0, 3.0};\n"
        "  int rc = gplvm_ppca_dense(0, 2, 2, 1, w0, 1.0, GPLVM_STOP_ITERATIONS, 1, 0.0, 0, 0, 0, 0);\n"
        '  printf("%d %s %s\\n", rc, gplvm_version(), gplvm_last_error());\n'
        "  (void)y; return rc == GPLVM_ERR_ARG ? 0 : 1;\n}\n")
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(_lib.library_path())
    g.load_library()
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L" + libdir, "-lgplvm_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sm_100a" in r.stdout and "null" in r.stdout


@pytest.mark.parametrize("src", ["imma_peak.cu", "tcgen05_i8_peak.cu", "tcgen05_i8_check.cu", "dense_i8_proto.cu", "dense_i8_proto2.cu", "dense_i8_stream.cu"])
def test_measurement_tools_compile_for_sm_100a(src, tmp_path):
    """The INT8 probes / known-answer checks that DESIGN.md cites (profiles/r1_fp64_peaks.log) must keep compiling."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / (src + ".o")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-c", os.path.join(root, "tools", src), "-o", str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
