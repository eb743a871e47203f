"""GPU suite: the nip.h drop-in (nip_b200/host/nip_gpu_backend.c) called exactly as
util/nipinference.c and util/niptrain.c call libnip, on a model parsed by the
reference's own host code, against the reference's own functions in the same
process (same parsed nip_model, same in-memory time_series)."""
import ctypes as C
import os

import numpy as np
import pytest

from cases import assert_close
from nip_b200.synth import HmmSpec, net_text_generic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BACKEND = os.path.join(ROOT, "nip_b200", "libnip_gpu_backend.so")
vp, i32, f64 = C.c_void_p, C.c_int, C.c_double


@pytest.fixture(scope="module")
def libs(gpu_lib):
    from oracle.bindings import REF_SO, RefLib, have_ref
    if not have_ref() or not os.path.exists(BACKEND):
        pytest.skip("needs the prebuilt oracle/_ref/libnip_ref.so and nip_b200/libnip_gpu_backend.so")
    C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)      # the unchanged host code: parser, lists, error handler
    ref = RefLib()
    gpu = C.CDLL(BACKEND)
    for f in (gpu.forward_backward_inference, gpu.forward_inference):
        f.restype = vp
        f.argtypes = [vp, vp, i32, C.POINTER(f64)]
    gpu.em_learn.argtypes = [vp, i32, f64, vp]
    gpu.make_consistent.argtypes = [vp]
    gpu.nip_gpu_release.argtypes = [vp]
    L = ref.L
    L.refh_variable.restype = vp
    L.refh_variable.argtypes = [vp, i32]
    L.refh_ucs_length.argtypes = [vp]
    L.refh_flatten_ucs.argtypes = [vp, vp]
    L.refh_free_ucs.argtypes = [vp]
    L.refh_new_double_list.restype = vp
    L.refh_double_list_to_array.argtypes = [vp, vp, i32]
    L.refh_free_double_list.argtypes = [vp]
    L.refh_seed.argtypes = [C.c_long]
    return ref, gpu


def _vars(ref, model, idx):
    return (vp * len(idx))(*[ref.L.refh_variable(model.h, v) for v in idx])


def _flat(ref, ucs, row):
    out = np.zeros((ref.L.refh_ucs_length(ucs), row))
    ref.L.refh_flatten_ucs(ucs, out.ctypes.data_as(vp))
    ref.L.refh_free_ucs(ucs)
    return out


def test_forward_backward_inference_dropin(libs, tmp_path):
    ref, gpu = libs
    h = HmmSpec(24, 7, seed=21)
    p = tmp_path / "h.net"
    p.write_text(h.net_text())
    model = ref.parse(p)
    query = [1, 2, 0]
    row = 24 + 24 + 7
    for s in h.sample(5, 30, seed=3, missing=0.1):
        ts = model.timeseries(h.obs_vars, s)
        want, ll_want = model.infer(ts, query)
        ll = f64()
        got = _flat(ref, gpu.forward_backward_inference(ts, _vars(ref, model, query), 3, C.byref(ll)), row)
        assert_close(got, want, "forward_backward_inference posteriors")
        assert_close(ll.value, ll_want, "forward_backward_inference loglikelihood")
        want, ll_want = model.infer(ts, query, forward_only=True)
        got = _flat(ref, gpu.forward_inference(ts, _vars(ref, model, query), 3, C.byref(ll)), row)
        assert_close(got, want, "forward_inference posteriors")
        assert_close(ll.value, ll_want, "forward_inference loglikelihood")
        # loglikelihood == NULL is allowed (util/nipmap.c:145)
        got = _flat(ref, gpu.forward_backward_inference(ts, _vars(ref, model, [1]), 1, None), 24)
        assert_close(got, model.infer(ts, [1])[0], "posteriors without likelihood")
    # unmarked variables are ignored (nip_unmark_variable, src/nip.c:993)
    model.mark(0, False)
    ts = model.timeseries(h.obs_vars, h.sample(1, 12, seed=9)[0])
    want, _ = model.infer(ts, [1])
    got = _flat(ref, gpu.forward_backward_inference(ts, _vars(ref, model, [1]), 1, None), 24)
    assert_close(got, want, "unmarked evidence column")
    gpu.nip_gpu_release(model.h)


def test_em_learn_dropin(libs, tmp_path):
    ref, gpu = libs
    text = net_text_generic(
        [("Y1", 3, None), ("X1", 4, None), ("X0", 4, "X1")],
        [("Y1", ["X1"], np.ones((4, 3))), ("X1", ["X0"], np.ones((4, 4))), ("X0", [], np.ones((1, 4)))])
    p = tmp_path / "e.net"
    p.write_text(text)
    rng = np.random.default_rng(4)
    series = [rng.integers(0, 3, size=(int(rng.integers(5, 20)), 1)).astype(np.int32) for _ in range(6)]
    m_ref, m_gpu = ref.parse(p), ref.parse(p)
    ts_ref = [m_ref.timeseries([0], s) for s in series]
    ts_gpu = [m_gpu.timeseries([0], s) for s in series]
    st_ref, curve_ref = m_ref.em_learn(ts_ref, 1.0, 99)          # threshold 1.0 => exactly 3 iterations
    ref.L.refh_seed(99)
    lc = ref.L.refh_new_double_list()
    arr = (vp * len(ts_gpu))(*ts_gpu)
    st_gpu = gpu.em_learn(arr, len(ts_gpu), 1.0, lc)
    curve = np.zeros(64)
    n = ref.L.refh_double_list_to_array(lc, curve.ctypes.data_as(vp), 64)
    ref.L.refh_free_double_list(lc)
    assert st_gpu == st_ref == 0
    assert n == len(curve_ref) == 3
    assert_close(curve[:n], curve_ref, "learning curve")
    t_ref, p_ref = m_ref.parameters()
    t_gpu, p_gpu = m_gpu.parameters()                             # trained CPTs are back on the host model
    assert_close(t_gpu, t_ref, "trained original_p")
    assert_close(p_gpu, p_ref, "trained priors")
    gpu.nip_gpu_release(m_gpu.h)


def _demo1_like_net():
    """general tree: three cliques, in_clique != out_clique, a parentless non-interface variable"""
    rng = np.random.default_rng(5)
    t = lambda *shape: rng.random(shape) + 0.05
    return net_text_generic(
        [("D1", 2, None), ("C1", 3, None), ("B1", 3, None), ("A1", 4, None), ("A0", 4, "A1")],
        [("D1", [], t(1, 2)), ("C1", ["A1", "D1"], t(8, 3)), ("B1", ["C1"], t(3, 3)),
         ("A1", ["A0"], t(4, 4)), ("A0", [], t(1, 4))])


@pytest.mark.parametrize("memo", ["1", "0"])
def test_make_consistent_dropin(libs, tmp_path, memo, monkeypatch):
    """make_consistent() must work on WHATEVER tree state the host holds (src/nip.c:1600-1617),
    not on a reconstruction: evidence entered on an already consistent tree, and the
    inter-slice message multiplied into in_clique->p by finish_timeslice_message_pass
    (generate_data, src/nip.c:2433-2461).  Compared: every clique->p, both potentials of every
    sepset, the mass and every marginal, after every step, with and without the memo."""
    ref, gpu = libs
    monkeypatch.setenv("NIP_GPU_SLICE_MEMO", memo)
    nets = {"hmm": HmmSpec(6, 4, seed=2).net_text(), "tree": _demo1_like_net()}
    for name, text in nets.items():
        p = tmp_path / (name + ".net")
        p.write_text(text)
        m_ref, m_gpu = ref.parse(p), ref.parse(p)
        cards = m_ref._cards()
        rng = np.random.default_rng(11)

        def both(f):
            f(m_ref)
            f(m_gpu)

        def check(what):
            m_ref.make_consistent()
            gpu.make_consistent(m_gpu.h)
            assert_close(m_gpu.tree_state(), m_ref.tree_state(), "%s: %s: cliques and sepsets" % (name, what))
            assert_close(m_gpu.mass(), m_ref.mass(), "%s: %s: model_prob_mass" % (name, what))
            for v in range(len(cards)):
                assert_close(m_gpu.marginal(v), m_ref.marginal(v), "%s: %s: get_probability(%d)" % (name, what, v))

        for rep in range(2):                      # the second round is served from the memo when it is on
            both(lambda m: (m.reset(), m.use_priors(0)))
            check("priors only")
            lik0 = np.array([0.1, 0.7, 0.0, 0.2, 0.5, 0.3])[:cards[0]]
            both(lambda m: m.enter_evidence(0, lik0))
            check("soft evidence")
            hard = np.zeros(cards[1]); hard[1] = 1.0
            both(lambda m: m.enter_evidence(1, hard))   # on the consistent tree: needs new/old division
            check("second evidence on a consistent tree")
            for t in range(3):                       # three slices of generate_data's loop
                both(lambda m: m.next_slice())
                check("slice %d after the inter-slice message" % (t + 1))
                e = rng.random(cards[0]) + 0.01
                both(lambda m: m.enter_evidence(0, e))
                check("slice %d with evidence" % (t + 1))
        gpu.nip_gpu_release(m_gpu.h)


@pytest.mark.parametrize("n_dev", [2, 4, 8])
def test_em_learn_dropin_on_several_devices(libs, tmp_path, n_dev, monkeypatch):
    """NIP_GPU_DEVICES=0,1,...: em_learn (src/nip.c:2076-2250, called by util/niptrain.c:151)
    shards the series over the devices inside the C boundary, one ncclAllReduce of the expected
    counts per iteration.  Learning curve and trained CPTs: 1e-12 against the one-device run,
    1e-9 against the reference"""
    import torch
    if torch.cuda.device_count() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    ref, gpu = libs
    h = HmmSpec(9, 4, seed=17)
    p = tmp_path / "e.net"
    p.write_text(h.net_text())
    data = h.sample(37, 21, seed=5, missing=0.05)
    data[:, 0, 0] = np.abs(data[:, 0, 0])
    rng = np.random.default_rng(3)
    series = [data[i, :int(rng.integers(3, 22))] for i in range(37)]

    def train(lib_em, model, seed):
        ts = [model.timeseries(h.obs_vars, s) for s in series]
        arr = (vp * len(ts))(*ts)
        ref.L.refh_seed(seed)
        lc = ref.L.refh_new_double_list()
        st = lib_em(arr, len(ts), 1e-4, lc)
        curve = np.zeros(4096)
        n = ref.L.refh_double_list_to_array(lc, curve.ctypes.data_as(vp), 4096)
        ref.L.refh_free_double_list(lc)
        return st, curve[:n].copy(), model.parameters()

    m_ref = ref.parse(p)
    st_ref, curve_ref = m_ref.em_learn([m_ref.timeseries(h.obs_vars, s) for s in series], 1e-4, 31)
    t_ref, p_ref = m_ref.parameters()
    monkeypatch.delenv("NIP_GPU_DEVICES", raising=False)
    m_one = ref.parse(p)
    st_one, curve_one, (t_one, p_one) = train(gpu.em_learn, m_one, 31)
    gpu.nip_gpu_release(m_one.h)
    monkeypatch.setenv("NIP_GPU_DEVICES", ",".join(str(d) for d in range(n_dev)))
    m_many = ref.parse(p)
    st_many, curve_many, (t_many, p_many) = train(gpu.em_learn, m_many, 31)
    gpu.nip_gpu_release(m_many.h)
    assert st_ref == st_one == st_many == 0
    assert len(curve_ref) == len(curve_one) == len(curve_many) and len(curve_ref) >= 3
    assert_close(curve_many, curve_one, "learning curve, %d devices vs one" % n_dev, rtol=1e-12)
    assert_close(t_many, t_one, "trained tables, %d devices vs one" % n_dev, rtol=1e-10)
    assert_close(curve_many, curve_ref, "learning curve vs the reference")
    assert_close(t_many, t_ref, "trained original_p vs the reference", rtol=1e-8)
    assert_close(p_many, p_ref, "trained priors vs the reference", rtol=1e-8)


def test_inference_after_training_is_not_served_stale(libs, tmp_path):
    """infer -> em_learn -> infer on a registered set: the second answer must come from the
    trained parameters (the parked pass of the first call is stale)"""
    ref, gpu = libs
    gpu.nip_gpu_register_set.argtypes = [vp, i32]
    gpu.nip_gpu_forget_set.argtypes = [vp]
    h = HmmSpec(5, 3, seed=6)
    p = tmp_path / "s.net"
    p.write_text(h.net_text())
    model = ref.parse(p)
    data = h.sample(6, 14, seed=2)
    ts = [model.timeseries(h.obs_vars, s) for s in data]
    arr = (vp * len(ts))(*ts)
    q = _vars(ref, model, [1])
    gpu.nip_gpu_register_set(arr, len(ts))
    before = _flat(ref, gpu.forward_backward_inference(ts[0], q, 1, None), 5)
    ref.L.refh_seed(5)
    lc = ref.L.refh_new_double_list()
    assert gpu.em_learn(arr, len(ts), 1.0, lc) == 0
    ref.L.refh_free_double_list(lc)
    after = _flat(ref, gpu.forward_backward_inference(ts[1], q, 1, None), 5)
    want, _ = model.infer(ts[1], [1])               # the reference on the trained host model
    assert_close(after, want, "series 1 after em_learn")
    again = _flat(ref, gpu.forward_backward_inference(ts[0], q, 1, None), 5)
    assert_close(again, model.infer(ts[0], [1])[0], "series 0 after em_learn")
    assert np.abs(again - before).max() > 1e-3      # training did change the answer
    gpu.nip_gpu_forget_set(arr)
    gpu.nip_gpu_release(model.h)


def test_set_with_reordered_columns_is_rejected(libs, tmp_path):
    """series of one set must list their observed variables in the same order; anything else
    would enter evidence on the wrong variable"""
    ref, gpu = libs
    gpu.nip_gpu_smooth_set.argtypes = [vp, i32, vp, i32, i32, vp, vp]
    h = HmmSpec(4, 4, seed=1)
    p = tmp_path / "r.net"
    p.write_text(h.net_text())
    model = ref.parse(p)
    a = model.timeseries([0, 1], np.zeros((3, 2), dtype=np.int32))
    b = model.timeseries([1, 0], np.zeros((3, 2), dtype=np.int32))
    arr = (vp * 2)(a, b)
    res = (vp * 2)()
    assert gpu.nip_gpu_smooth_set(arr, 2, _vars(ref, model, [1]), 1, 0, res, None) == 3   # NIP_ERROR_INVALID_ARGUMENT
    gpu.nip_gpu_release(model.h)


def test_transparent_batching_of_per_series_calls(libs, tmp_path):
    """util/nipinference.c:125-129 calls forward_backward_inference once per series of a set;
    once the set is registered the first call smooths the whole set in one device pass and the
    others are served from it — same results as the reference, a fraction of the launches"""
    import nip_b200.api as api
    ref, gpu = libs
    gpu.nip_gpu_register_set.argtypes = [vp, i32]
    gpu.nip_gpu_forget_set.argtypes = [vp]
    h = HmmSpec(16, 5, seed=8)
    p = tmp_path / "b.net"
    p.write_text(h.net_text())
    model = ref.parse(p)
    data = h.sample(12, 25, seed=4, missing=0.1)
    series = [data[i, :10 + i] for i in range(12)]            # ragged
    ts = [model.timeseries(h.obs_vars, s) for s in series]
    arr = (vp * len(ts))(*ts)
    q = _vars(ref, model, [1])
    gpu.nip_gpu_register_set(arr, len(ts))
    api.launch_count(reset=True)
    launches = []
    for k, t in enumerate(ts):
        ll = f64()
        got = _flat(ref, gpu.forward_backward_inference(t, q, 1, C.byref(ll)), 16)
        want, ll_want = model.infer(t, [1])
        assert_close(got, want, "batched series %d posterior" % k)
        assert_close(ll.value, ll_want, "batched series %d loglikelihood" % k)
        launches.append(api.launch_count())
    assert launches[-1] == launches[0], "series 2..n must be served from the first pass"
    # a different query invalidates the parked pass; a repeated request is computed again
    got = _flat(ref, gpu.forward_backward_inference(ts[3], _vars(ref, model, [1, 0]), 2, None), 16 + 5)
    assert_close(got, model.infer(ts[3], [1, 0])[0], "new query after a parked pass")
    got = _flat(ref, gpu.forward_inference(ts[3], q, 1, None), 16)
    assert_close(got, model.infer(ts[3], [1], forward_only=True)[0], "filtering after smoothing")
    # changed marks change the evidence: no stale result may be served
    model.mark(0, False)
    got = _flat(ref, gpu.forward_backward_inference(ts[5], q, 1, None), 16)
    assert_close(got, model.infer(ts[5], [1])[0], "marks changed")
    model.mark(0, True)
    gpu.nip_gpu_forget_set(arr)
    gpu.nip_gpu_release(model.h)


# ---- the reference's unchanged command-line tools, relinked -------------------------------
TOOLS = os.path.join(ROOT, "oracle", "_ref")


def _run(tool, *args):
    import subprocess
    exe = os.path.join(TOOLS, tool)
    if not os.path.exists(exe):
        pytest.skip("needs the prebuilt oracle/_ref/%s (make -C oracle, where /root/reference is present)" % tool)
    r = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def _observed_only(src, dst, name):
    lines = open(src).read().split("\n")
    k = lines[0].split(",").index(name)
    open(dst, "w").write("\n".join([name] + [l.split(",")[k] if l.strip() else "" for l in lines[1:]]))


def test_unchanged_cli_tools_on_the_gpu_backend(gpu_lib, tmp_path):
    """util/nipinference.c and util/niptrain.c, compiled from the reference's sources without a
    single edit and linked as INTEGRATION.md says, against the same tools linked with the
    reference's own code: same files in, same files out"""
    h = HmmSpec(12, 5, seed=3)
    net = tmp_path / "h.net"
    net.write_text(h.net_text())
    _run("sample_driver_cpu", net, 40, 15, 11, tmp_path / "all.txt")
    _observed_only(tmp_path / "all.txt", tmp_path / "m1.txt", "M1")
    out_cpu = _run("nipinference_cpu", net, tmp_path / "m1.txt", "P1", tmp_path / "post_cpu.txt")
    out_gpu = _run("nipinference_gpu", net, tmp_path / "m1.txt", "P1", tmp_path / "post_gpu.txt")
    a = np.genfromtxt(tmp_path / "post_cpu.txt", delimiter=",", skip_header=1)
    b = np.genfromtxt(tmp_path / "post_gpu.txt", delimiter=",", skip_header=1)
    assert a.shape == b.shape and a.shape[1] == 12 and np.isfinite(a).sum() == 40 * 15 * 12
    assert np.nanmax(np.abs(a - b)) <= 1.01e-6           # the tool prints six decimals
    ll = lambda s: float(s.split("Average log. likelihood =")[1].split()[0])
    assert abs(ll(out_cpu) - ll(out_gpu)) <= 1e-5 * abs(ll(out_cpu))
    # niptrain seeds rand() from the clock, so two runs never agree digit for digit: train on the
    # GPU backend, then let the REFERENCE tool score the model it wrote
    log = _run("niptrain_gpu", net, tmp_path / "m1.txt", 0.0001, -5, tmp_path / "trained.net")
    final = float(log.strip().split("average loglikelihood =")[-1].split()[0])
    scored = ll(_run("nipinference_cpu", tmp_path / "trained.net", tmp_path / "m1.txt", "P1", tmp_path / "p2.txt"))
    assert np.isfinite(final) and final < 0
    # write_model keeps six decimals (a probability below 5e-7 reloads as 0), so only a loose check
    if np.isfinite(scored):
        assert abs(final - scored) <= 0.2


def _table(text):
    """all numbers of a tool's stdout, line by line"""
    rows = []
    for line in text.split("\n"):
        vals = []
        for tok in line.replace("=", " ").replace(",", " ").replace("(", " ").replace(")", " ").split():
            try:
                vals.append(float(tok))
            except ValueError:
                pass
        if vals:
            rows.append(vals)
    return rows


def _same_numbers(a, b, what, rtol=2e-5):
    ta, tb = _table(a), _table(b)
    assert len(ta) == len(tb) and len(ta) > 0, what
    for ra, rb in zip(ta, tb):
        assert len(ra) == len(rb), what
        assert np.allclose(ra, rb, rtol=rtol, atol=1.5e-6), (what, ra, rb)    # %g / %f: six digits


def _sample_files(tmp_path, h, n=30, T=12, seed=7):
    net = tmp_path / "h.net"
    net.write_text(h.net_text())
    _run("sample_driver_cpu", net, n, T, seed, tmp_path / "all.txt")
    _observed_only(tmp_path / "all.txt", tmp_path / "m1.txt", "M1")
    return net


def test_generate_data_on_the_gpu_backend(gpu_lib, tmp_path):
    """generate_data (src/nip.c:2325-2478) unchanged, its make_consistent calls on the device:
    with the same rand() seed the GPU-linked build draws the same series as the reference
    (the slices of a series depend on each other through in_clique->p, which the replacement
    must honour), and the unchanged nipsample tool produces series whose previous-slice column
    repeats the interface state of the slice before"""
    h = HmmSpec(7, 4, seed=5)
    net = tmp_path / "g.net"
    net.write_text(h.net_text())
    _run("sample_driver_cpu", net, 12, 9, 123, tmp_path / "cpu.txt")
    _run("sample_driver_gpu", net, 12, 9, 123, tmp_path / "gpu.txt")
    assert open(tmp_path / "cpu.txt").read() == open(tmp_path / "gpu.txt").read()
    _run("nipsample_gpu", net, 20, 10, tmp_path / "s.txt")
    lines = open(tmp_path / "s.txt").read().split("\n")
    cols = lines[0].split(",")
    p0, p1 = cols.index("P0"), cols.index("P1")
    n_pairs, prev = 0, None
    for l in lines[1:]:
        if not l.strip():
            prev = None
            continue
        row = l.split(",")
        if prev is not None:
            assert row[p0] == prev[p1], "P0 of a slice must be the P1 drawn in the slice before"
            n_pairs += 1
        prev = row
    assert n_pairs == 20 * 9


def test_niplikelihood_nipjoint_nipmap_on_the_gpu_backend(gpu_lib, tmp_path):
    """the remaining callers of the fine-grained API (util/niplikelihood.c:111-135,
    util/nipjoint.c:77-96) and of the smoother (util/nipmap.c:139-145), unchanged and relinked:
    same files in, same numbers out as the reference build"""
    h = HmmSpec(9, 5, seed=13)
    net = _sample_files(tmp_path, h)
    for memo in ("1", "0"):
        os.environ["NIP_GPU_SLICE_MEMO"] = memo
        try:
            _same_numbers(_run("niplikelihood_gpu", net, tmp_path / "m1.txt", "M1"),
                          _run("niplikelihood_cpu", net, tmp_path / "m1.txt", "M1"), "niplikelihood memo=" + memo)
        finally:
            del os.environ["NIP_GPU_SLICE_MEMO"]
    # (without variable arguments util/nipjoint.c:114-119 takes ts->hidden and frees it twice,
    # :142 and free_timeseries: undefined behaviour in the reference itself, not exercised)
    _same_numbers(_run("nipjoint_gpu", net, tmp_path / "m1.txt", "P1"),
                  _run("nipjoint_cpu", net, tmp_path / "m1.txt", "P1"), "nipjoint P1")
    _run("nipmap_gpu", net, tmp_path / "m1.txt", tmp_path / "map_gpu.txt")
    _run("nipmap_cpu", net, tmp_path / "m1.txt", tmp_path / "map_cpu.txt")
    a, b = open(tmp_path / "map_cpu.txt").read().split("\n"), open(tmp_path / "map_gpu.txt").read().split("\n")
    assert len(a) == len(b) and len(a) > 30 * 12
    differ = sum(x.split() != y.split() for x, y in zip(a, b))
    assert differ <= 2, "MAP states differ in %d rows (only exact posterior ties may)" % differ


def test_packed_file_straight_into_a_batch(libs, tmp_path):
    """a set written by nip_gpu_write_timeseries_bin() is smoothed straight
    from the file (one bulk read, one upload, no time_series structs) — same posteriors and
    log-likelihoods as the reference computes on the series it was written from"""
    ref, gpu = libs
    gpu.nip_gpu_write_timeseries_bin.argtypes = [vp, i32, C.c_char_p]
    gpu.nip_gpu_smooth_bin.argtypes = [vp, C.c_char_p, vp, i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    h = HmmSpec(11, 6, seed=19)
    p = tmp_path / "f.net"
    p.write_text(h.net_text())
    model = ref.parse(p)
    data = h.sample(9, 14, seed=6, missing=0.1)
    series = [data[i, :4 + i] for i in range(9)]
    ts = [model.timeseries(h.obs_vars, s) for s in series]
    arr = (vp * len(ts))(*ts)
    path = str(tmp_path / "set.bin").encode()
    assert gpu.nip_gpu_write_timeseries_bin(arr, len(ts), path) == 0
    post, ll, lens = vp(), vp(), vp()
    n = gpu.nip_gpu_smooth_bin(model.h, path, _vars(ref, model, [1, 0]), 2, 0, C.byref(post), C.byref(ll), C.byref(lens))
    assert n == 9
    L = np.ctypeslib.as_array(C.cast(lens, C.POINTER(C.c_int32)), (n,))
    assert list(L) == [len(s) for s in series]
    P = np.ctypeslib.as_array(C.cast(post, C.POINTER(f64)), (int(L.sum()), 11 + 6))
    LL = np.ctypeslib.as_array(C.cast(ll, C.POINTER(f64)), (n,))
    r0 = 0
    for i, t in enumerate(ts):
        want, ll_want = model.infer(t, [1, 0])
        assert_close(P[r0:r0 + L[i]], want, "series %d from the packed file" % i)
        assert_close(LL[i], ll_want, "series %d loglikelihood" % i)
        r0 += L[i]
    libc = C.CDLL(None)
    libc.free.argtypes = [vp]
    for x in (post, ll, lens):
        libc.free(x)
    gpu.nip_gpu_release(model.h)


def test_generate_set_dropin(libs, tmp_path):
    """nip_gpu_generate_set(): a whole set of fully observed series drawn on the device, usable by
    the reference's own code (here: its smoother reproduces the sampled hidden states when every
    variable is observed)"""
    ref, gpu = libs
    gpu.nip_gpu_generate_set.argtypes = [vp, i32, i32, C.c_ulong, C.POINTER(vp)]
    gpu.nip_gpu_forget_set.argtypes = [vp]
    h = HmmSpec(6, 4, seed=12)
    p = tmp_path / "g.net"
    p.write_text(h.net_text())
    model = ref.parse(p)
    out = vp()
    n = gpu.nip_gpu_generate_set(model.h, 5, 7, 42, C.byref(out))
    assert n == 5
    series = C.cast(out, C.POINTER(vp))
    for i in range(n):
        model._T[series[i]] = 7
        post, ll = model.infer(series[i], [1])          # P1 is observed in the generated data
        assert np.all(np.isclose(post.max(axis=1), 1.0)) and np.isfinite(ll)
    gpu.nip_gpu_forget_set(out)
    gpu.nip_gpu_release(model.h)


def test_dropin_on_the_factor_engine(libs, tmp_path, monkeypatch):
    """NIP_GPU_ENGINE=3: the reference's entry points on the engine that evaluates the join tree
    factor by factor (what a model with huge cliques gets by itself), on a general tree with
    in_clique != out_clique: smoothing, filtering, and em_learn's learning curve + trained CPTs"""
    ref, gpu = libs
    monkeypatch.setenv("NIP_GPU_ENGINE", "3")
    p = tmp_path / "d.net"
    p.write_text(_demo1_like_net())
    rng = np.random.default_rng(8)
    series = [np.stack([rng.integers(0, 2, size=n), rng.integers(0, 3, size=n)], axis=1).astype(np.int32)
              for n in (7, 12, 3, 9)]
    for s in series:
        s[rng.random(s.shape) < 0.15] = -1
        s[0] = np.abs(s[0])
    model = ref.parse(p)
    obs, query = [0, 2], [1, 3, 4]          # D1, B1 observed; C1, A1, A0 queried
    row = 3 + 4 + 4
    for s in series:
        ts = model.timeseries(obs, s)
        want, ll_want = model.infer(ts, query)
        ll = f64()
        got = _flat(ref, gpu.forward_backward_inference(ts, _vars(ref, model, query), 3, C.byref(ll)), row)
        assert_close(got, want, "factor engine: forward_backward_inference posteriors")
        assert_close(ll.value, ll_want, "factor engine: loglikelihood")
        want, ll_want = model.infer(ts, query, forward_only=True)
        got = _flat(ref, gpu.forward_inference(ts, _vars(ref, model, query), 3, C.byref(ll)), row)
        assert_close(got, want, "factor engine: forward_inference posteriors")
    gpu.nip_gpu_release(model.h)
    m_ref, m_gpu = ref.parse(p), ref.parse(p)
    ts_ref = [m_ref.timeseries(obs, s) for s in series]
    ts_gpu = [m_gpu.timeseries(obs, s) for s in series]
    st_ref, curve_ref = m_ref.em_learn(ts_ref, 1e-3, 7)
    ref.L.refh_seed(7)
    lc = ref.L.refh_new_double_list()
    arr = (vp * len(ts_gpu))(*ts_gpu)
    st_gpu = gpu.em_learn(arr, len(ts_gpu), 1e-3, lc)
    curve = np.zeros(256)
    n = ref.L.refh_double_list_to_array(lc, curve.ctypes.data_as(vp), 256)
    ref.L.refh_free_double_list(lc)
    assert st_gpu == st_ref
    assert n == len(curve_ref)
    assert_close(curve[:n], curve_ref, "factor engine: learning curve")
    t_ref, p_ref = m_ref.parameters()
    t_gpu, p_gpu = m_gpu.parameters()
    assert_close(t_gpu, t_ref, "factor engine: trained original_p", rtol=1e-8)
    assert_close(p_gpu, p_ref, "factor engine: trained priors", rtol=1e-8)
    gpu.nip_gpu_release(m_gpu.h)
