"""The C-ABI library loads and exports every symbol include/ukf_batch.h declares; without a
GPU every compute entry point refuses to run (there is no CPU path)."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ukf_batch.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ukfb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from slam_pose_estimation_b200 import _build, _capi

    _build.build()
    return _capi.load()


def test_header_and_binding_agree(lib):
    from slam_pose_estimation_b200 import _capi

    names = declared_functions()
    assert len(names) >= 40
    assert sorted(_capi.SIGNATURES) == names


def test_every_declared_symbol_is_exported(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/ukf_batch.h but not exported by libukfb.so"


def test_no_torch_or_python_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "ukf_batch.h")).read()
    assert "torch" not in text.lower() and "at::" not in text and "PyObject" not in text


def test_meas_dims(lib):
    want = {0: 3, 1: 2, 2: 1, 3: 3, 4: 3, 5: 2, 6: 1, 7: 2, 8: 3, 9: 3}
    for k, m in want.items():
        assert lib.ukfb_meas_dim(k) == m
    assert lib.ukfb_meas_dim(-1) == 0 and lib.ukfb_meas_dim(10) == 0


def test_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.ukfb_create(0, 4, 0, C.byref(h))
    assert rc == -3 and not h.value  # UKFB_ERR_CUDA
    assert b"no CPU path" in lib.ukfb_last_error() or b"CUDA" in lib.ukfb_last_error()
    from slam_pose_estimation_b200 import UkfBatch, UkfbError

    with pytest.raises(UkfbError):
        UkfBatch(0, 4)


def test_bad_arguments(lib):
    h = C.c_void_p()
    assert lib.ukfb_create(7, 4, 0, C.byref(h)) == -1
    assert lib.ukfb_create(0, 0, 0, C.byref(h)) == -1
    assert lib.ukfb_create(0, 4, 0, None) == -1
    assert lib.ukfb_batch(None) == 0 and lib.ukfb_dof(None) == 0
    assert lib.ukfb_predict_dt(None, None, 0) == -1


def test_product_never_touches_the_oracle():
    """nothing under the package or include/ imports, includes or links oracle/"""
    pkg = os.path.join(ROOT, "slam_pose_estimation_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(base, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"


def test_header_is_plain_c_and_links_from_c(lib, tmp_path):
    """the boundary is a C ABI: a C99 translation unit includes the header, links libukfb.so and gets the documented
    refusal (UKFB_ERR_CUDA) or a working handle, with no C++ or Python in between"""
    import subprocess

    from slam_pose_estimation_b200 import _build

    src = tmp_path / "abi.c"
    src.write_text('''#include <stdio.h>
#include <ukf_batch.h>
int main(void) {
    ukfb_handle* h = 0;
    int rc = ukfb_create(UKFB_POSE, 4, 0, &h);
    printf("%d %d %d %d\\n", rc, ukfb_meas_dim(UKFB_MEAS_POSE_XY), UKFB_RBS_DOUBLES, UKFB_EVENT_KIND_COUNT);
    if (rc == UKFB_OK) { printf("%d %d\\n", ukfb_dof(h), ukfb_mu_size(h)); ukfb_destroy(h); }
    else printf("%s\\n", ukfb_last_error());
    return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_build.LIB)
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-L", libdir, "-lukfb", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    rc, m, rbs, nk = (int(v) for v in out[0].split())
    assert (m, rbs, nk) == (2, 49, 13)
    assert rc in (0, -3)
    if rc == 0:
        assert out[1].split() == ["12", "13"]
    else:
        assert "CUDA" in out[1] or "device" in out[1]


def test_library_staleness_is_by_content_not_by_time():
    """The built library travels with a snapshot of the tree (the GPU box, the round-end run): a copy keeps contents, not
    necessarily time stamps, and must not trigger a rebuild -- nor may a changed source go unnoticed."""
    import os

    from slam_pose_estimation_b200 import _build

    if os.environ.get("UKFB_LIB"):
        pytest.skip("a library chosen through UKFB_LIB is taken as it is")
    _build.build()
    assert not _build.stale()
    dep = os.path.join(_build.CSRC, "so3.cuh")
    st = os.stat(dep)
    try:
        os.utime(dep, (st.st_atime, os.path.getmtime(_build.LIB) + 3600.0))  # "newer" than the library, same content
        assert not _build.stale()
        saved = open(_build.LIB + ".srchash").read()
        open(_build.LIB + ".srchash", "w").write("0" * 64 + "\n")  # as if a source had changed since the build
        assert _build.stale()
        open(_build.LIB + ".srchash", "w").write(saved)
        assert not _build.stale()
    finally:
        os.utime(dep, (st.st_atime, st.st_mtime))
