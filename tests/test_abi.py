"""CPU suite: the C-ABI library loads and exports every symbol include/nipgpu.h
declares (no compute calls — there is no GPU here), and refuses to compute
without a device instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import nip_b200.api as api
from cases import Case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nipgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nipgpu_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(api.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    from nip_b200 import build
    build.build_device_library()
    lib = C.CDLL(api.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), "libnipgpu.so does not export " + s


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    c = Case("hmm5")
    with pytest.raises(api.NipGpuError) as e:
        api.Model(c.fm)
    assert e.value.code == 2  # NIPGPU_ENODEVICE


def test_missing_library_is_loud(tmp_path):
    api_lib, api._lib = api._lib, None
    try:
        with pytest.raises(FileNotFoundError):
            api.load_library(str(tmp_path / "libnipgpu.so"))
    finally:
        api._lib = api_lib
