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


from cases import ALL_CASES


@pytest.mark.parametrize("name", ALL_CASES)
def test_factor_extraction_on_the_host(name):
    """engine 3's factor extraction and verification are host code (no device): every fixture the
    reference generated — structural zeros, no interface, shared observations included — factors
    into its families' CPTs"""
    ok, why = api.factorable(Case(name).fm)
    assert ok, why


def test_factor_extraction_refuses_tables_that_do_not_factor():
    """one entry of a 6-variable clique of the factorial model changed: the table is no longer a
    product of 3-variable CPTs, the check says which clique, and a malformed description is an
    error, not a 'no'"""
    fm = Case("factorial4x3").fm
    big = int(np.argmax(np.diff(fm.clique_tab_off)))
    lo, hi = int(fm.clique_tab_off[big]), int(fm.clique_tab_off[big + 1])
    fm.clique_tables[lo + int(np.argmax(fm.clique_tables[lo:hi]))] *= 1.25
    ok, why = api.factorable(fm)
    assert not ok and ("clique %d" % big) in why
    bad = Case("hmm5").fm
    bad.var_family[0] = 99
    with pytest.raises(api.NipGpuError):
        api.factorable(bad)
