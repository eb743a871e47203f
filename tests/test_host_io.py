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


def _posteriors(model, obs_names, series, query_names, names):
    idx = {s: i for i, s in enumerate(names)}
    obs = [idx[s] for s in obs_names]
    q = [idx[s] for s in query_names]
    out = []
    for s in series:
        ts = model.timeseries(obs, s)
        post, ll = model.infer(ts, q)
        out.append((post, ll))
    return out


@pytest.mark.parametrize("kind", ["hmm", "coupled"])
def test_exact_model_writer_round_trips(ref_lib, kind, tmp_path):
    """write_model() prints six decimals and reloads differently;
    nip_gpu_write_model_exact() -> parse_model() gives the same posteriors and likelihoods to 1e-12,
    also after EM has moved the parameters away from anything the text had"""
    import re
    if not os.path.exists(BACKEND):
        pytest.skip("needs nip_b200/libnip_gpu_backend.so")
    from oracle.bindings import REF_SO
    from nip_b200.synth import net_text_generic
    C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
    be = C.CDLL(BACKEND)
    be.nip_gpu_write_model_exact.argtypes = [vp, C.c_char_p]
    rng = np.random.default_rng(3)
    if kind == "hmm":
        text = HmmSpec(7, 4, seed=9).net_text()
        obs_names, query_names = ["M1"], ["P1", "P0"]
        series = [rng.integers(-1, 4, size=(int(T), 1)).astype(np.int32) for T in (5, 9, 1, 12)]
    else:
        text = net_text_generic(
            [("YA", 2, None), ("YB", 4, None), ("A1", 3, None), ("B1", 3, None), ("A0", 3, "A1"), ("B0", 3, "B1")],
            [("YA", ["A1"], rng.random((3, 2)) + 0.05), ("YB", ["B1"], rng.random((3, 4)) + 0.05),
             ("A1", ["B0", "A0"], rng.random((3, 3, 3)) + 0.05), ("B1", ["A0", "B0"], rng.random((3, 3, 3)) + 0.05),
             ("A0", [], (rng.random(3) + 0.1)[None, :]), ("B0", [], (rng.random(3) + 0.1)[None, :])])
        obs_names, query_names = ["YA", "YB"], ["A1", "B1", "A0"]
        series = [np.stack([rng.integers(-1, 2, size=int(T)), rng.integers(-1, 4, size=int(T))], axis=1).astype(np.int32)
                  for T in (4, 7, 2)]
    for s in series:
        s[0] = np.abs(s[0])
    src = tmp_path / "src.net"
    src.write_text(text)
    names = re.findall(r"^\s*node\s+(\w+)", text, flags=re.M)
    model = ref_lib.parse(src)
    idx = {s: i for i, s in enumerate(names)}
    ts = [model.timeseries([idx[s] for s in obs_names], s) for s in series]
    model.em_learn(ts, 1.0, 5)                       # three EM iterations from random parameters
    want = _posteriors(model, obs_names, series, query_names, names)
    out = tmp_path / "exact.net"
    assert be.nip_gpu_write_model_exact(model.h, str(out).encode()) == 0
    written = out.read_text()
    names2 = re.findall(r"^\s*node\s+(\w+)", written, flags=re.M)
    assert sorted(names2) == sorted(names)
    got = _posteriors(ref_lib.parse(out), obs_names, series, query_names, names2)
    for (pw, lw), (pg, lg) in zip(want, got):
        assert np.allclose(pg, pw, rtol=1e-12, atol=1e-15)
        assert abs(lg - lw) <= 1e-12 * abs(lw)
