"""CPU-side checks of the drop-in boundary: the library builds/loads without a GPU, exports every
symbol include/sfgpu.h declares, and refuses to work without a device (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as H
from sigfish_b200 import build as B
from sigfish_b200 import capi


@pytest.fixture(scope="module")
def lib():
    B.build_gpu()
    return capi.lib()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(H.ROOT, "include", "sfgpu.h")).read()
    declared = set(re.findall(r"\b(sfgpu_[a-z_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_struct_layouts_match_header():
    assert C.sizeof(capi.Opt) == 48
    assert C.sizeof(capi.Result) == 64
    assert capi.RESULT_DTYPE.itemsize == 64
    assert C.sizeof(capi.Timing) == 64


def test_no_device_is_an_error_not_a_fallback(lib):
    if lib.sfgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SfgpuError):
        capi.Context(np.zeros(4 ** 6, np.float32), 6)


def test_argument_validation_happens_before_device_use(lib):
    h = C.c_void_p()
    opt = capi.Opt(device=0, flags=0, query_size=0, prefix_size=50, kmer_size=6, n_slots=2)
    lm = np.zeros(4 ** 6, np.float32)
    assert lib.sfgpu_create(C.byref(h), C.byref(opt), lm.ctypes.data_as(C.c_void_p)) == -2
    opt.query_size = 2000
    assert lib.sfgpu_create(C.byref(h), C.byref(opt), lm.ctypes.data_as(C.c_void_p)) == -5
    opt.query_size = 250
    opt.prefix_size = -1
    assert lib.sfgpu_create(C.byref(h), C.byref(opt), lm.ctypes.data_as(C.c_void_p)) == -2
    assert b"prefix_size" in lib.sfgpu_strerror(None)


def test_product_never_touches_the_oracle_or_the_test_double():
    """oracle/ and tests/mockdev/ are test infrastructure: no product source may include, link, load or name them, and
    the built product libraries must not depend on them"""
    import glob
    import re
    import subprocess
    ROOT = H.ROOT
    pat = re.compile(r"oracle[/_.]|liboracle|mockdev|sfgpu_oracle|sfgpu_null", re.I)
    srcs = [f for ext in ("c", "h", "cu", "cuh", "py") for f in glob.glob(os.path.join(ROOT, "sigfish_b200", "**", "*." + ext), recursive=True)]
    srcs += glob.glob(os.path.join(ROOT, "include", "*.h")) + [os.path.join(ROOT, "Makefile")]
    assert len(srcs) > 20
    for f in srcs:
        for n, line in enumerate(open(f, errors="replace"), 1):
            if f.endswith("Makefile") and line.lstrip().startswith(("#", "oracle:", ".PHONY", "test:", "\t$(MAKE) -C oracle")):
                continue  # the convenience targets that build / run the test infrastructure
            assert not pat.search(line), f"{f}:{n}: {line.strip()}"
    for so in ("libsfgpu.so", "libsfhost.so", "sigfish-b200"):
        p = os.path.join(ROOT, "sigfish_b200", so)
        if os.path.exists(p):
            needed = subprocess.run(["readelf", "-d", p], capture_output=True, text=True).stdout
            assert "oracle" not in needed and "mockdev" not in needed, so
