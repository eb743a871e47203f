"""Host-side plumbing that needs no GPU: the packed data files of nip_b200/host/nip_data_bin.c
against the reference's own text reader (src/nip.c:512-667), through the reference's CPU code."""
import ctypes as C
import os

import numpy as np
import pytest

from nip_b200.synth import HmmSpec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BACKEND = os.path.join(ROOT, "nip_b200", "libnip_gpu_backend.so")
vp, i32 = C.c_void_p, C.c_int


def test_binary_data_files_round_trip_through_the_reference(ref_lib, tmp_path):
    if not os.path.exists(BACKEND):
        pytest.skip("needs nip_b200/libnip_gpu_backend.so (built where the reference headers are)")
    from oracle.bindings import REF_SO
    C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
    be = C.CDLL(BACKEND)
    be.nip_gpu_write_timeseries_bin.argtypes = [vp, i32, C.c_char_p]
    be.nip_gpu_read_timeseries_bin.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    be.nip_gpu_forget_set.argtypes = [vp]
    h = HmmSpec(9, 4, seed=6)
    net = tmp_path / "h.net"
    net.write_text(h.net_text())
    model = ref_lib.parse(net)
    data = h.sample(7, 12, seed=8, missing=0.2)
    series = [data[i, :3 + i] for i in range(7)]
    ts = [model.timeseries(h.obs_vars, s) for s in series]
    arr = (vp * len(ts))(*ts)
    path = str(tmp_path / "set.nipb").encode()
    assert be.nip_gpu_write_timeseries_bin(arr, len(ts), path) == 0
    assert os.path.getsize(path) == 4 + 12 + (4 + 2) + 4 * 7 + 4 * sum(len(s) for s in series)
    out = vp()
    n = be.nip_gpu_read_timeseries_bin(model.h, path, C.byref(out))
    assert n == len(ts)
    loaded = C.cast(out, C.POINTER(vp))
    for i in range(n):   # the loaded series are ordinary time_series: the reference smooths them
        model._T[loaded[i]] = len(series[i])
        want, ll_want = model.infer(ts[i], [1, 0])
        got, ll_got = model.infer(loaded[i], [1, 0])
        assert np.array_equal(got, want) and ll_got == ll_want
    be.nip_gpu_forget_set(out)
    # a truncated file is rejected, not half-loaded
    blob = open(path, "rb").read()
    open(path, "wb").write(blob[:-5])
    assert be.nip_gpu_read_timeseries_bin(model.h, path, C.byref(out)) == 0
