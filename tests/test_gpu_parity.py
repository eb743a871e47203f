"""GPU parity suite (run with -m gpu on a B200): the CUDA path, called through the
C ABI (include/nipgpu.h), against
  * the golden fixtures generated from the reference itself, and
  * the C oracle on seeded synthetic inputs.
Tolerance: 1e-9 RELATIVE on posterior marginals, log-likelihoods, expected counts
and re-estimated CPTs (BASELINE.json north star); exact zeros must stay exact
zeros (atol = 0) and -DBL_MAX must be reproduced exactly.
"""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from cases import (ALL_CASES, EM_CASES, LIKELIHOOD_CASES, SLICE_CASES, Case, assert_close, unhex)

pytestmark = pytest.mark.gpu

ENGINES = [1, 0]  # NIPGPU_ENGINE_JTREE, NIPGPU_ENGINE_AUTO (chain engine when the model allows)
ENGINES3 = [1, 0, 3]  # ... and NIPGPU_ENGINE_FACTOR (the join tree factor by factor, any model)


@pytest.mark.parametrize("engine", ENGINES3)
@pytest.mark.parametrize("name", ALL_CASES)
def test_inference_golden(gpu_lib, name, engine):
    c = Case(name)
    m = gpu_lib.Model(c.fm, engine=engine)
    b = m.batch(c.obs_vars, c.series)
    for kind, fwd in (("smooth", False), ("filter", True)):
        posts, lls = c.expected(kind)
        post, ll = b.infer(c.query, forward_only=fwd)
        for i, got in enumerate(b.split(post)):
            assert_close(got, posts[i], "%s %s series %d posterior (engine %d)" % (name, kind, i, m.engine))
        assert_close(ll, lls, "%s %s loglik" % (name, kind))
        # interface-variable-only query: the path the chain engine serves directly
        if c.fm.n_interface == 1:
            v = int(c.fm.outgoing[0])
            k = c.query.index(v) if v in c.query else None
            if k is not None:
                off = int(sum(c.fm.var_card[q] for q in c.query[:k]))
                card = int(c.fm.var_card[v])
                post1, ll1 = b.infer([v], forward_only=fwd)
                for i, got in enumerate(b.split(post1)):
                    assert_close(got, posts[i][:, off:off + card], "%s %s interface posterior" % (name, kind))
                assert_close(ll1, lls, "%s %s loglik (interface query)" % (name, kind))
    b.close()
    m.close()


@pytest.mark.parametrize("engine", ENGINES3)
def test_unmarked_columns_are_ignored(gpu_lib, oracle_lib, engine):
    """NIP_MARK semantics (src/nip.c:993): only marked variables enter evidence"""
    c = Case("hmm12_two_leaves")
    om = oracle_lib.model(c.fm)
    m = gpu_lib.Model(c.fm, engine=engine)
    b = m.batch(c.obs_vars, c.series)
    for keep in ([1, 0], [0, 1], [0, 0]):
        mask = np.zeros(c.fm.n_vars, dtype=np.uint8)
        for k, on in zip(c.obs_vars, keep):
            mask[k] = on
        post, ll = b.infer([2], use_evidence=mask)
        for i, got in enumerate(b.split(post)):
            want, llw = om.infer(c.obs_vars, c.series[i], [2], use_evidence=mask)
            assert_close(got, want, "masked posterior %r" % (keep,))
            assert_close(ll[i], llw, "masked loglik %r" % (keep,), atol=1e-12)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("S,M,B,T", [(64, 32, 48, 40), (7, 3, 100, 17), (33, 5, 19, 9),
                                     (30, 4, 77, 23), (32, 6, 9, 12), (57, 3, 130, 11),   # 4 / 8 state tiles: warp pairs
                                     (72, 5, 150, 9), (130, 4, 21, 6), (200, 3, 300, 5),
                                     (72, 4, 700, 6)])  # > 64: GEMM E-step; 700: two concurrent halves
def test_hmm_em_vs_oracle(gpu_lib, oracle_lib, engine, S, M, B, T):
    """E-step of seeded synthetic HMMs (ragged, missing data) against the oracle"""
    from nip_b200.synth import HmmSpec
    h = HmmSpec(S, M, seed=S + M)
    fm = h.flat()
    data = h.sample(B, T, seed=3, missing=0.1)
    data[:, 0, 0] = np.abs(data[:, 0, 0])      # observed first slices
    rng = np.random.default_rng(5)
    series = [data[i, :int(rng.integers(1, T + 1))] for i in range(B)]
    want, ll_want, st_want = oracle_lib.model(fm).estep(h.obs_vars, series)
    m = gpu_lib.Model(fm, engine=engine)
    counts, ll, st = m.batch(h.obs_vars, series).estep()
    assert st == st_want == 0
    assert_close(counts, want, "HMM-%d expected counts (engine %d)" % (S, m.engine))
    assert_close(ll, ll_want, "HMM-%d EM loglik" % S)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("S,M,B,T", [(64, 32, 48, 40), (7, 3, 100, 17), (33, 5, 19, 9),
                                     (30, 4, 77, 23), (32, 6, 9, 12), (57, 3, 130, 11),   # 4 / 8 state tiles: warp pairs
                                     (72, 5, 150, 9), (130, 4, 21, 6), (72, 4, 700, 6)])   # > 64 states: tiled-GEMM engine
def test_hmm_vs_oracle(gpu_lib, oracle_lib, engine, S, M, B, T):
    """seeded synthetic HMMs of the benchmark family, ragged lengths, missing data"""
    from nip_b200.synth import HmmSpec
    h = HmmSpec(S, M, seed=S + M)
    fm = h.flat()
    data = h.sample(B, T, seed=3, missing=0.1)
    rng = np.random.default_rng(5)
    series = [data[i, :int(rng.integers(1, T + 1))] for i in range(B)]
    om = oracle_lib.model(fm)
    m = gpu_lib.Model(fm, engine=engine)
    b = m.batch(h.obs_vars, series)
    post, ll = b.infer(h.hidden_query)
    fpost, fll = b.infer(h.hidden_query, forward_only=True)
    for i, (got, fgot) in enumerate(zip(b.split(post), b.split(fpost))):
        want, llw = om.infer(h.obs_vars, series[i], h.hidden_query)
        assert_close(got, want, "HMM-%d series %d smoothed" % (S, i))
        # atol: a series made of missing observations only has ll == 0 here and +-1e-16 of
        # rounding noise in the reference
        assert_close(ll[i], llw, "HMM-%d series %d loglik" % (S, i), atol=1e-12)
        want, llw = om.infer(h.obs_vars, series[i], h.hidden_query, forward_only=True)
        assert_close(fgot, want, "HMM-%d series %d filtered" % (S, i))
        assert_close(fll[i], llw, "HMM-%d series %d loglik (filter)" % (S, i), atol=1e-12)


@pytest.mark.parametrize("name", ["demo1_net", "coupled2x3"])
def test_hbm_workspace_path(gpu_lib, name, monkeypatch):
    """cliques too large for shared memory are staged in a per-CTA HBM workspace (config C3's
    16^6-entry cliques); force that path on small models and check it against the goldens"""
    monkeypatch.setenv("NIPGPU_FORCE_HBM_WORKSPACE", "1")
    c = Case(name)
    m = gpu_lib.Model(c.fm, engine=1)
    b = m.batch(c.obs_vars, c.series)
    posts, lls = c.expected("smooth")
    post, ll = b.infer(c.query)
    for i, got in enumerate(b.split(post)):
        assert_close(got, posts[i], "%s series %d posterior (HBM workspace)" % (name, i))
    assert_close(ll, lls, "%s loglik (HBM workspace)" % name)
    it = c.j["em"]["iters"][0]
    m.mstep(unhex(c.j["em"]["init"]))
    counts, L, st = b.estep()
    assert st == 0
    assert_close(counts, unhex(it["counts"]), "%s expected counts (HBM workspace)" % name)


@pytest.mark.parametrize("mode", ["warp", "cta", "hbm", "grid"])
@pytest.mark.parametrize("name", ALL_CASES)
def test_generic_engine_team_modes(gpu_lib, name, mode, monkeypatch):
    """the generic engine's kernels are instantiated for three teams (warp / CTA / whole grid) and
    two table homes (shared memory / HBM); every combination must reproduce every golden: smoothing,
    filtering, E-step counts, the likelihood loop and the single-slice API"""
    monkeypatch.setenv("NIPGPU_JT_MODE", mode)
    c = Case(name)
    try:
        m = gpu_lib.Model(c.fm, engine=1)
    except Exception as e:          # tables larger than shared memory in a shared-memory mode
        assert mode in ("warp", "cta") and "do not fit" in str(e)
        pytest.skip("model does not fit shared memory in mode " + mode)
    b = m.batch(c.obs_vars, c.series)
    for kind, fwd in (("smooth", False), ("filter", True)):
        posts, lls = c.expected(kind)
        post, ll = b.infer(c.query, forward_only=fwd)
        for i, got in enumerate(b.split(post)):
            assert_close(got, posts[i], "%s series %d %s (%s)" % (name, i, kind, mode))
        assert_close(ll, lls, "%s loglik %s (%s)" % (name, kind, mode))
    if "likelihood" in c.j and name in LIKELIHOOD_CASES:
        on = np.zeros(c.fm.n_vars, dtype=np.uint8)
        on[c.j["likelihood"]["marked"]] = 1
        out = b.likelihood(1 - on, on)
        for i, got in enumerate(b.split(out)):
            assert_close(got.reshape(-1), unhex(c.j["likelihood"]["out"][i]), "%s likelihood (%s)" % (name, mode))
    if name in EM_CASES:
        it = c.j["em"]["iters"][0]
        m.mstep(unhex(c.j["em"]["init"]))
        counts, L, st = b.estep()
        assert st == 0
        assert_close(counts, unhex(it["counts"]), "%s expected counts (%s)" % (name, mode))
        assert_close(L, float.fromhex(it["ll"]), "%s EM loglik (%s)" % (name, mode))
    if name in SLICE_CASES:
        m2 = gpu_lib.Model(c.fm, engine=1)
        step = c.j["slice"][-1]
        m2.slice_reset()
        m2.slice_use_priors(step["has_history"])
        for var, lik in step["evidence"]:
            m2.slice_enter_evidence(var, unhex(lik))
        m2.slice_make_consistent()
        assert_close(m2.slice_mass(), float.fromhex(step["mass"]), "mass (%s)" % mode)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("mode", [None, "grid", "hbm"])
def test_empty_and_single_slice(gpu_lib, engine, mode, monkeypatch):
    """T = 0 and T = 1 series, an empty batch, and their E-steps (every engine and team)"""
    if mode:
        if engine != 1:
            pytest.skip("team modes belong to the generic engine")
        monkeypatch.setenv("NIPGPU_JT_MODE", mode)
    c = Case("hmm5")
    m = gpu_lib.Model(c.fm, engine=engine)
    b = m.batch(c.obs_vars, [np.zeros((0, 1), dtype=np.int32), c.series[6]])   # T = 0 and T = 1
    post, ll = b.infer(c.query)
    assert ll[0] == 0.0
    posts, lls = c.expected("smooth")
    assert_close(post, posts[6], "single-slice series")
    assert_close(ll[1], lls[6], "single-slice loglik")
    counts, L, st = b.estep(add_pseudocount=False)
    assert st == 0
    assert_close(L, lls[6], "single-slice EM loglik")
    b0 = m.batch(c.obs_vars, [])
    post, ll = b0.infer(c.query)
    assert post.shape[0] == 0 and ll.shape[0] == 0
    counts, L, st = b0.estep(add_pseudocount=True)
    assert st == 0 and L == 0.0 and np.all(counts == 1.0)      # only the pseudo-count


@pytest.mark.parametrize("engine", ENGINES3)
@pytest.mark.parametrize("name", EM_CASES)
def test_em_golden(gpu_lib, name, engine):
    """each EM iteration from the reference's own inputs: M-step tables/priors, then
    E-step expected counts and log-likelihood (engine 1: per-slice family marginals;
    engine 2: transition counts as one DMMA GEMM)"""
    c = Case(name)
    m = gpu_lib.Model(c.fm, engine=engine)
    b = m.batch(c.obs_vars, c.series)
    counts_in = unhex(c.j["em"]["init"])
    for k, it in enumerate(c.j["em"]["iters"]):
        m.mstep(counts_in)
        tables, prior = m.parameters()
        assert_close(tables, unhex(it["tables"]), "%s iter %d re-estimated clique tables" % (name, k))
        assert_close(prior, unhex(it["prior"]), "%s iter %d priors" % (name, k))
        counts, ll, st = b.estep()
        assert (st != 0) == (it["status"] != 0)
        if it["status"] == 0:
            assert_close(counts, unhex(it["counts"]), "%s iter %d expected counts" % (name, k))
            assert_close(ll, float.fromhex(it["ll"]), "%s iter %d loglik" % (name, k))
        counts_in = unhex(it["counts"])


@pytest.mark.parametrize("engine", ENGINES3)
def test_bad_luck_is_reported(gpu_lib, oracle_lib, engine):
    """an impossible observation (m2 == 0) must surface as NIP_ERROR_BAD_LUCK
    (src/nip.c:1827-1854), a clean set must not"""
    c = Case("model_net")
    m = gpu_lib.Model(c.fm, engine=engine)
    om = oracle_lib.model(c.fm)
    impossible = [np.array([1, 0, 4]).reshape(-1, 1)]   # P(1,0,4) == 0 in examples/model.net
    fine = [np.array([2, 3, 2, 3, 2, 4]).reshape(-1, 1)]
    for series, want in ((impossible, 8), (fine, 0), (fine + impossible, 8)):
        _, _, st_o = om.estep(c.obs_vars, series)
        _, _, st = m.batch(c.obs_vars, series).estep()
        assert st == want and st_o == want


def test_evidence_free_first_slice_is_not_bad_luck(gpu_lib):
    """DESIGN.md, deviations: for a slice without evidence the reference computes
    log(m2) - log(m1) from two masses that differ only by rounding and may flag
    BAD_LUCK by chance; the device path defines m2 := m1 there."""
    c = Case("hmm5")
    m = gpu_lib.Model(c.fm)
    series = [np.array([-1, 2, 1, -1, 0]).reshape(-1, 1), np.array([-1]).reshape(-1, 1)]
    b = m.batch(c.obs_vars, series)
    _, ll, st = b.estep()
    assert st == 0 and ll < 0
    _, lls = b.infer(c.query)
    assert lls[1] == 0.0


def test_em_pseudocount_once_and_device_mstep(gpu_lib):
    """counts without the 1.0 pseudo-count + 1 == counts with it; an M-step straight
    from the device accumulator equals an M-step from the same counts uploaded"""
    c = Case("hmm5")
    m = gpu_lib.Model(c.fm)
    b = m.batch(c.obs_vars, c.series)
    with_p, ll1, _ = b.estep(add_pseudocount=True)
    without, ll0, _ = b.estep(add_pseudocount=False)
    assert_close(without + 1.0, with_p, "pseudo-count")
    assert ll0 == ll1
    b.estep(add_pseudocount=True, want_counts=False)
    m.mstep(None)
    t_dev, p_dev = m.parameters()
    m.mstep(with_p)
    t_up, p_up = m.parameters()
    assert np.array_equal(t_dev, t_up) and np.array_equal(p_dev, p_up)


@pytest.mark.parametrize("name", LIKELIHOOD_CASES)
def test_likelihood_golden(gpu_lib, name):
    c = Case(name)
    m = gpu_lib.Model(c.fm)
    b = m.batch(c.obs_vars, c.series)
    on = np.zeros(c.fm.n_vars, dtype=np.uint8)
    on[c.j["likelihood"]["marked"]] = 1
    out = b.likelihood(1 - on, on)
    for i, got in enumerate(b.split(out)):
        assert_close(got.reshape(-1), unhex(c.j["likelihood"]["out"][i]), "%s likelihood series %d" % (name, i))


@pytest.mark.parametrize("name", SLICE_CASES)
def test_slice_api_golden(gpu_lib, name):
    c = Case(name)
    m = gpu_lib.Model(c.fm)
    for step in c.j["slice"]:
        m.slice_reset()
        m.slice_use_priors(step["has_history"])
        for var, lik in step["evidence"]:
            m.slice_enter_evidence(var, unhex(lik))
        m.slice_make_consistent()
        assert_close(m.slice_mass(), float.fromhex(step["mass"]), "mass")
        for v in range(c.fm.n_vars):
            assert_close(m.slice_marginal(v), unhex(step["marginals"][v]), "marginal of %d" % v)
        for k in range(c.fm.n_cliques):
            assert_close(m.slice_clique(k), unhex(step["cliques"][k]), "clique %d" % k)


def test_launches_are_counted(gpu_lib):
    c = Case("hmm5")
    m = gpu_lib.Model(c.fm)
    b = m.batch(c.obs_vars, c.series)
    gpu_lib.launch_count(reset=True)
    b.infer(c.query)
    assert gpu_lib.launch_count() >= 2


# ---- factorial DBN (config C3): cliques too large for shared memory ------------------------
@pytest.mark.parametrize("coupled", [True, False])
@pytest.mark.parametrize("ns,mode", [(6, "hbm"), (6, "grid"), (8, None), (8, "grid"), (6, "factor"), (8, "factor"),
                                     (2, "factor"), (10, "factor")])   # 2 x 2 tiles over pairs of even-cardinality variables
def test_factorial_vs_oracle(gpu_lib, oracle_lib, ns, mode, coupled, monkeypatch):
    """4 ring-coupled chains (C3's topology) with cliques too large for shared memory, in the
    per-CTA HBM workspace and with the whole grid streaming one sequence: smoothing, filtering
    and the E-step against the oracle"""
    from nip_b200.synth import FactorialSpec
    if mode and mode != "factor":
        monkeypatch.setenv("NIPGPU_JT_MODE", mode)
    sp = FactorialSpec(ns, 3, seed=4, coupled=coupled)
    fm = sp.flat()
    data = sp.sample(3, 4, seed=5, missing=0.2)
    data[:, 0, :] = np.abs(data[:, 0, :])          # observed first slices
    series = [data[0], data[1][:2], data[2][:1]]
    query = [4, 7, 9, 2]                            # X0, X3, W1, Y2
    om = oracle_lib.model(fm)
    m = gpu_lib.Model(fm, engine=3 if mode == "factor" else 1)
    assert m.engine == (3 if mode == "factor" else 1)
    b = m.batch(sp.obs_vars, series)
    for fwd in (False, True):
        post, ll = b.infer(query, forward_only=fwd)
        for i, got in enumerate(b.split(post)):
            want, llw = om.infer(sp.obs_vars, series[i], query, forward_only=fwd)
            assert_close(got, want, "factorial ns=%d series %d posterior (fwd=%s)" % (ns, i, fwd))
            assert_close(ll[i], llw, "factorial ns=%d series %d loglik" % (ns, i))
    want, ll_want, st_want = om.estep(sp.obs_vars, series)
    counts, L, st = b.estep()
    assert st == st_want == 0
    assert_close(counts, want, "factorial ns=%d expected counts" % ns)
    assert_close(L, ll_want, "factorial ns=%d EM loglik" % ns)


@pytest.mark.parametrize("engine", [1, 0])
def test_c3_full_size_vs_oracle(gpu_lib, oracle_lib, engine):
    """config C3 at its real size — 16 states per chain, three 16^6-entry cliques (403 MB of
    tables), interface of 65 536 states — one short series against the oracle (which needs
    about ten seconds per slice), plus the size-independent checks"""
    from nip_b200.synth import FactorialSpec
    sp = FactorialSpec(16, 4, seed=1)
    fm = sp.flat()
    series = [sp.sample(1, 2, seed=2)[0]]
    query = [4, 6, 9]                               # X0, X2, W1
    m = gpu_lib.Model(fm, engine=engine)
    assert m.engine == (1 if engine == 1 else 3)    # left to itself the library evaluates C3 factor by factor
    b = m.batch(sp.obs_vars, series)
    post, ll = b.infer(query)
    want, llw = oracle_lib.model(fm).infer(sp.obs_vars, series[0], query)
    assert_close(post, want, "C3 posterior")
    assert_close(ll[0], llw, "C3 loglik")
    assert np.allclose(post.reshape(2, 3, 16).sum(-1), 1.0, rtol=0, atol=1e-12)
    counts, L, st = b.estep(add_pseudocount=False)
    assert st == 0
    assert_close(L, llw, "C3 EM loglik")
    # every variable's expected family counts sum to the number of slices it is counted in
    off = m.counts_offsets()
    for v in range(fm.n_vars):
        slices = 1 if v >= 8 else 2                 # W^i (previous slice) only at t = 0
        assert abs(counts[off[v]:off[v + 1]].sum() - slices) < 1e-9


@pytest.mark.parametrize("name", LIKELIHOOD_CASES)
def test_likelihood_memo_equals_direct(gpu_lib, name, monkeypatch):
    """with few evidence configurations the likelihood loop evaluates each configuration once
    and the records gather; that path must give the same bits as evaluating every record"""
    c = Case(name)
    rng = np.random.default_rng(7)
    cards = [int(c.fm.var_card[v]) for v in c.obs_vars]
    series = [np.stack([rng.integers(-1, k, size=int(T)) for k in cards], axis=1).astype(np.int32)
              for T in rng.integers(1, 40, size=60)]
    on = np.zeros(c.fm.n_vars, dtype=np.uint8)
    on[c.j["likelihood"]["marked"]] = 1
    m = gpu_lib.Model(c.fm)
    memo = m.batch(c.obs_vars, series).likelihood(1 - on, on)
    monkeypatch.setenv("NIPGPU_NO_LIKELIHOOD_MEMO", "1")
    direct = m.batch(c.obs_vars, series).likelihood(1 - on, on)
    assert np.array_equal(memo, direct)


@pytest.mark.parametrize("order", [[0], [1], [0, 1], [1, 0]])
def test_composite_interface_queries_on_the_chain_engine(gpu_lib, order):
    """two coupled chains: the interface is the pair (A1, B1); the chain engine smooths the joint
    state and the queried interface variables are its marginals"""
    c = Case("coupled2x3")
    iface = [int(v) for v in c.fm.outgoing]                 # A1, B1
    q = [iface[k] for k in order]
    cols = {v: (int(sum(c.fm.var_card[u] for u in c.query[:c.query.index(v)])), int(c.fm.var_card[v])) for v in iface}
    m = gpu_lib.Model(c.fm, engine=0)
    b = m.batch(c.obs_vars, c.series)
    for kind, fwd in (("smooth", False), ("filter", True)):
        posts, lls = c.expected(kind)
        post, ll = b.infer(q, forward_only=fwd)
        for i, got in enumerate(b.split(post)):
            want = np.concatenate([posts[i][:, cols[v][0]:cols[v][0] + cols[v][1]] for v in q], axis=1)
            assert_close(got, want, "coupled2x3 %s series %d, interface query %r" % (kind, i, order))
        assert_close(ll, lls, "coupled2x3 %s loglik" % kind)


@pytest.mark.parametrize("S,M,B,T", [(7, 3, 60, 11), (64, 32, 40, 25), (130, 4, 21, 6)])
def test_hmm_all_variables_vs_oracle(gpu_lib, oracle_lib, S, M, B, T):
    """observation, state and previous-state posteriors of seeded HMMs (missing data, ragged
    lengths) through the chain engines' query projection, against the oracle"""
    from nip_b200.synth import HmmSpec
    h = HmmSpec(S, M, seed=S + M)
    fm = h.flat()
    data = h.sample(B, T, seed=3, missing=0.25)
    rng = np.random.default_rng(5)
    series = [data[i, :int(rng.integers(1, T + 1))] for i in range(B)]
    om = oracle_lib.model(fm)
    m = gpu_lib.Model(fm, engine=0)
    b = m.batch(h.obs_vars, series)
    for q in ([0, 1, 2], [2, 0], [0]):          # M1, P1, P0
        post, ll = b.infer(q)
        for i, got in enumerate(b.split(post)):
            want, llw = om.infer(h.obs_vars, series[i], q)
            assert_close(got, want, "HMM-%d series %d query %r" % (S, i, q), atol=1e-300)
            assert_close(ll[i], llw, "HMM-%d series %d loglik" % (S, i), atol=1e-12)
    post, ll = b.infer([0, 1], forward_only=True)
    for i, got in enumerate(b.split(post)):
        want, llw = om.infer(h.obs_vars, series[i], [0, 1], forward_only=True)
        assert_close(got, want, "HMM-%d series %d filtered (M1, P1)" % (S, i), atol=1e-300)


def test_batch_update_checks_the_observation_range(gpu_lib):
    """a re-uploaded observation >= the cardinality of its variable is refused (it would index
    past the evidence tables); valid data is accepted and used"""
    from nip_b200.api import NipGpuError
    c = Case("hmm5")
    m = gpu_lib.Model(c.fm)
    s = [np.array(x, dtype=np.int32) for x in c.series]
    b = m.batch(c.obs_vars, s)
    flat = np.concatenate([x.reshape(-1) for x in s]).astype(np.int32)
    good = np.where(flat >= 0, (flat + 1) % int(c.fm.var_card[c.obs_vars[0]]), flat).astype(np.int32)
    b.update(good)
    post, ll = b.infer(c.query)
    b2 = m.batch(c.obs_vars, [g.reshape(-1, 1) for g in np.split(good, np.cumsum([len(x) for x in s])[:-1])])
    post2, ll2 = b2.infer(c.query)
    assert np.array_equal(post, post2) and np.array_equal(ll, ll2)
    bad = good.copy()
    bad[len(bad) // 2] = int(c.fm.var_card[c.obs_vars[0]])
    with pytest.raises(NipGpuError):
        b.update(bad)


@pytest.mark.parametrize("name", ["hmm5", "coupled2x3", "model_net"])
def test_device_sampler_matches_the_model(gpu_lib, name):
    """nipgpu_sample draws whole series on the device; with 200 000 series the empirical marginal of
    every variable at every slice must sit within 5 sigma of the exact evidence-free marginal, and a
    previous-slice variable must repeat its successor's value of the slice before.  examples/model.net
    is declared parent first, so its tables are not proper CPTs (the parser's dimension-0
    normalisation): there the forward sampler agrees with exact inference on the first slice only,
    like the reference's generate_data, which never looks at future slices either."""
    c = Case(name)
    m = gpu_lib.Model(c.fm, engine=0)
    N, T = 200000, 5
    x = m.sample(N, T, seed=11)
    assert x.shape == (N, T, c.fm.n_vars)
    for k in range(c.fm.n_interface):
        assert np.array_equal(x[:, 1:, int(c.fm.prev_outgoing[k])], x[:, :-1, int(c.fm.outgoing[k])])
    allvars = list(range(c.fm.n_vars))
    proper = name != "model_net"
    b = m.batch(c.obs_vars, [np.full((T if proper else 1, len(c.obs_vars)), -1, dtype=np.int32)])
    post, _ = b.infer(allvars)                      # no evidence: the prior predictive marginals
    off = 0
    for v in allvars:
        card = int(c.fm.var_card[v])
        for t in range(T if proper else 1):
            want = post[t, off:off + card]
            got = np.bincount(x[:, t, v], minlength=card) / N
            sigma = np.sqrt(np.maximum(want * (1 - want), 1e-12) / N)
            assert np.all(np.abs(got - want) <= 5 * sigma + 1e-9), (name, v, t, got, want)
        off += card
    assert not np.array_equal(x, m.sample(N, T, seed=12))
    assert np.array_equal(x, m.sample(N, T, seed=11))


def test_c4_shape_vs_oracle(gpu_lib, oracle_lib):
    """BASELINE configs[3] at its real shape: 1024-state interface clique (8 x 8 tiles of the
    per-slice DGEMM, the split-K count GEMM, two concurrent halves), 64 symbols; ragged series
    with missing data; smoothing, filtering and the E-step against the oracle"""
    from nip_b200.synth import HmmSpec
    h = HmmSpec(1024, 64, seed=41)
    fm = h.flat()
    data = h.sample(9, 6, seed=3, missing=0.15)
    data[:, 0, 0] = np.abs(data[:, 0, 0])
    lens = [6, 5, 1, 3, 6, 2, 4, 6, 5]
    series = [data[i, :lens[i]] for i in range(9)]
    om = oracle_lib.model(fm)
    m = gpu_lib.Model(fm)
    assert m.engine == gpu_lib.ENGINE_CHAIN
    b = m.batch(h.obs_vars, series)
    post, ll = b.infer(h.hidden_query)
    fpost, fll = b.infer(h.hidden_query, forward_only=True)
    for i, (got, fgot) in enumerate(zip(b.split(post), b.split(fpost))):
        want, llw = om.infer(h.obs_vars, series[i], h.hidden_query)
        assert_close(got, want, "C4 series %d smoothed" % i)
        assert_close(ll[i], llw, "C4 series %d loglik" % i)
        want, llw = om.infer(h.obs_vars, series[i], h.hidden_query, forward_only=True)
        assert_close(fgot, want, "C4 series %d filtered" % i)
        assert_close(fll[i], llw, "C4 series %d loglik (filter)" % i)
    want, ll_want, st_want = om.estep(h.obs_vars, series)
    counts, L, st = b.estep()
    assert st == st_want == 0
    assert_close(counts, want, "C4 expected counts")
    assert_close(L, ll_want, "C4 EM loglik")
    b.close()
    m.close()


@pytest.mark.parametrize("mode", ["", "cta", "hbm", "grid"])
@pytest.mark.parametrize("name", ["model_net", "demo1_net", "factorial4x3", "no_interface", "two_layer"])
def test_slice_propagate_vs_oracle(gpu_lib, oracle_lib, name, mode, monkeypatch):
    """nipgpu_slice_propagate = make_consistent on the caller's own state: start from the oracle's
    tree after priors + evidence, then perturb the consistent tree (new evidence multiplied into
    a clique, as nip_enter_evidence and finish_timeslice_message_pass do) and propagate again —
    the second pass divides by the sepsets the first one left"""
    if mode:
        monkeypatch.setenv("NIPGPU_JT_MODE", mode)
    c = Case(name)
    fm = c.fm
    try:
        m = gpu_lib.Model(fm, engine=1)
    except gpu_lib.NipGpuError as e:
        pytest.skip(str(e))
    om = oracle_lib.model(fm)
    om.reset()
    om.use_priors(0)
    rng = np.random.default_rng(3)
    lik = rng.random(int(fm.var_card[c.obs_vars[0]])) + 0.05
    lik[0] = 0.0
    om.enter_evidence(c.obs_vars[0], lik)
    start = np.concatenate([om.clique(k) for k in range(fm.n_cliques)])
    n_msg = int(sum(np.prod([fm.var_card[v] for v in fm.sepset_vars[fm.sepset_var_off[s]:fm.sepset_var_off[s + 1]]])
                    for s in range(fm.n_sepsets)))
    ones = np.ones(n_msg)
    tab, new, old = m.slice_propagate(start, ones)
    om.make_consistent()
    want = np.concatenate([om.clique(k) for k in range(fm.n_cliques)])
    assert_close(tab, want, "%s consistent cliques" % name)
    # mass = sum cliques - sum sepsets (src/nipjointree.c:1156-1188)
    assert_close(tab.sum() - new.sum(), om.mass(), "%s mass" % name)
    # second round on the consistent tree: scale one clique along one of its variables
    v = [int(x) for x in fm.clique_vars if int(x) != int(c.obs_vars[0])][-1]
    lik2 = rng.random(int(fm.var_card[v])) + 0.05
    om.enter_evidence(v, lik2)          # multiplied into the family clique of v, nothing else touched
    start2 = np.concatenate([om.clique(k) for k in range(fm.n_cliques)])
    tab2, new2, old2 = m.slice_propagate(start2, new)
    om.make_consistent()
    want2 = np.concatenate([om.clique(k) for k in range(fm.n_cliques)])
    assert_close(tab2, want2, "%s second propagation on a consistent tree" % name)
    assert_close(tab2.sum() - new2.sum(), om.mass(), "%s mass after the second propagation" % name)
    m.close()


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
@pytest.mark.parametrize("name", ["hmm64", "tree"])
def test_group_em_matches_one_device(gpu_lib, n_dev, name):
    """EM over several devices of one box (nipgpu_group_*): series sharded over the members,
    one ncclAllReduce of the expected counts per iteration.  Reduced counts, log-likelihood and
    the parameters after three iterations must equal the one-device run to 1e-12"""
    if _device_count() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    from nip_b200.dist import shard_series
    from nip_b200.synth import HmmSpec
    if name == "hmm64":
        h = HmmSpec(64, 8, seed=9)
        fm, obs_vars = h.flat(), h.obs_vars
        data = h.sample(203, 30, seed=4, missing=0.05)
        data[:, 0, 0] = np.abs(data[:, 0, 0])
        rng = np.random.default_rng(2)
        series = [data[i, :int(rng.integers(2, 31))] for i in range(203)]
    else:
        c = Case("demo1_net")
        fm, obs_vars = c.fm, c.obs_vars
        series = [s for s in c.series for _ in range(7)]
    init = np.random.default_rng(1).random(0)
    single = gpu_lib.Model(fm, device=0)
    init = np.random.default_rng(1).random(single.counts_size()) + 0.1
    sb = single.batch(obs_vars, series)
    models = [gpu_lib.Model(fm, device=d) for d in range(n_dev)]
    parts = shard_series([len(s) for s in series], n_dev)
    batches = [models[d].batch(obs_vars, [series[i] for i in parts[d]]) for d in range(n_dev)]
    g = gpu_lib.Group(models)
    single.mstep(init)
    g.mstep(init)
    for it in range(3):
        want, ll_want, st_want = sb.estep()
        got, ll, st = g.estep(batches)
        assert st == st_want == 0
        assert_close(got, want, "%s iteration %d reduced counts on %d devices" % (name, it, n_dev), rtol=1e-12)
        assert_close(ll, ll_want, "%s iteration %d loglik" % (name, it), rtol=1e-12)
        single.mstep()
        g.mstep()
    t0, p0 = single.parameters()
    for mdl in models:
        t, p = mdl.parameters()
        assert_close(t, t0, "trained tables on device %d" % mdl.device, rtol=1e-12)
        assert_close(p, p0, "trained priors on device %d" % mdl.device, rtol=1e-12)
    g.close()
    for b in batches:
        b.close()
    for mdl in models:
        mdl.close()
    sb.close()
    single.close()


@pytest.mark.parametrize("env", [{"NIPGPU_CHAIN_TEAM": "2"}, {"NIPGPU_CHAIN_TEAM": "4"}, {"NIPGPU_CHAIN_PAIR": "0"}])
def test_chain_kernel_flavours(gpu_lib, env):
    """the chain engine's three kernel flavours — teams of two warps on one scheduler (what a
    batch above 3072 sequences runs on), teams of four warps on four schedulers (small batches:
    everything else in this suite), one warp per group (interfaces that are not 4 or 8 state
    tiles wide) — forced one by one on the same ragged sets: 1e-9 against the oracle, and
    bit-identical results from two runs (the warps of a team exchange halves of every vector
    through shared memory behind a named barrier)"""
    import os
    import subprocess
    import sys
    helper = os.path.join(os.path.dirname(os.path.abspath(__file__)), "team_kernels_check.py")
    outs = []
    for _ in range(2):
        r = subprocess.run([sys.executable, helper], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert r.stdout.strip().startswith("OK")
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1], "results differ from run to run"


def test_thread_per_sequence_kernels(gpu_lib):
    """interfaces of 1..8 joint states on the thread-per-sequence kernels (what a batch of 2048+
    sequences of a one-tile model such as examples/model.net runs on), forced on small ragged sets
    with missing data: 1e-9 against the oracle for smoothing, filtering, likelihood, queries
    through the joint projection, and the E-step afterwards on the same batch"""
    import os
    import subprocess
    import sys
    helper = os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_kernels_check.py")
    r = subprocess.run([sys.executable, helper], env=dict(os.environ, NIPGPU_CHAIN_SMALL="1"), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.strip().endswith("OK")


def test_thread_per_sequence_kernels_large_batch_matches_dmma_kernels(gpu_lib):
    """4096 sequences of the 4-state model.net shape: the automatic choice (thread per sequence)
    against the DMMA kernels forced with NIPGPU_CHAIN_SMALL=0 in a subprocess, at 1e-12"""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "import nip_b200.api as api\n"
        "from nip_b200.synth import HmmSpec\n"
        "h = HmmSpec(4, 5, seed=1); data = h.sample(4096, 30, seed=2, missing=0.05)\n"
        "m = api.Model(h.flat()); b = m.batch(h.obs_vars, [d for d in data])\n"
        "post, ll = b.infer(h.hidden_query)\n"
        "np.save(sys.argv[1], np.concatenate([post.ravel(), ll.ravel()]))\n" % ROOT)
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for flag in (None, "0"):
            env = dict(os.environ)
            env.pop("NIPGPU_CHAIN_SMALL", None)
            if flag is not None:
                env["NIPGPU_CHAIN_SMALL"] = flag
            f = os.path.join(d, "o%s.npy" % flag)
            r = subprocess.run([sys.executable, "-c", code, f], env=env, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            outs.append(np.load(f))
    assert_close(outs[0], outs[1], "thread-per-sequence vs DMMA kernels", rtol=1e-12)


def test_chunked_host_copy_matches_single_pass(gpu_lib, oracle_lib):
    """host-buffered smoothing of a large equal-length set runs in four chunks of sequences whose
    rows are copied to the host while the next chunk is computed: same posteriors and
    log-likelihoods as the single pass (NIPGPU_CHUNKED_COPY=0), and the oracle's on the first and
    last series (the threshold is lowered to 1 MB so that a 2148-series set qualifies)"""
    import subprocess
    import sys
    import tempfile
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "import nip_b200.api as api\n"
        "from nip_b200.synth import HmmSpec\n"
        "h = HmmSpec(16, 4, seed=5); data = h.sample(2148, 24, seed=6, missing=0.1)\n"
        "m = api.Model(h.flat()); b = m.batch(h.obs_vars, [d for d in data])\n"
        "post, ll = b.infer(h.hidden_query)\n"
        "b.update(np.ascontiguousarray(data.reshape(-1, 1))); post2, ll2 = b.infer(h.hidden_query)\n"
        "assert np.array_equal(post, post2) and np.array_equal(ll, ll2)\n"
        "print('launches', api.launch_count())\n"
        "np.save(sys.argv[1], np.concatenate([post.ravel(), ll.ravel()]))\n" % ROOT)
    outs, launches = [], []
    with tempfile.TemporaryDirectory() as d:
        for flag in ("1", "0"):
            env = dict(os.environ, NIPGPU_CHUNKED_COPY=flag, NIPGPU_CHUNKED_COPY_MIN_MB="1")
            f = os.path.join(d, "o%s.npy" % flag)
            r = subprocess.run([sys.executable, "-c", code, f], env=env, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            launches.append(int(r.stdout.split("launches")[1].split()[0]))
            outs.append(np.load(f))
    assert launches[0] > launches[1]            # the chunked run really launched per chunk
    assert_close(outs[0], outs[1], "chunked vs single pass", rtol=1e-13)
    from nip_b200.synth import HmmSpec
    h = HmmSpec(16, 4, seed=5)
    data = h.sample(2148, 24, seed=6, missing=0.1)
    om = oracle_lib.model(h.flat())
    post = outs[0][:2148 * 24 * 16].reshape(2148, 24, 16)
    ll = outs[0][2148 * 24 * 16:]
    for i in (0, 2147):
        want, llw = om.infer(h.obs_vars, data[i], h.hidden_query)
        assert_close(post[i], want, "chunked copy: series %d" % i)
        assert_close(ll[i], llw, "chunked copy: loglik %d" % i, atol=1e-12)


def test_factor_engine_falls_back_when_it_cannot_plan(gpu_lib, oracle_lib, monkeypatch):
    """engine 3 chosen by the library itself (tables too large for shared memory) hands a request
    it cannot plan — here: the operand limit lowered to 1 — to engine 1 instead of failing; asked
    for explicitly it reports NIPGPU_EUNSUPPORTED"""
    from nip_b200.synth import FactorialSpec
    sp = FactorialSpec(6, 3, seed=4)
    fm = sp.flat()
    data = sp.sample(2, 3, seed=5)
    series = [data[0], data[1][:2]]
    monkeypatch.setenv("NIPGPU_FACTOR_MAX_OPS", "1")
    m = gpu_lib.Model(fm, engine=0)
    assert m.engine == 3
    b = m.batch(sp.obs_vars, series)
    post, ll = b.infer([4, 9])
    assert m.engine == 1
    om = oracle_lib.model(fm)
    for i, got in enumerate(b.split(post)):
        want, llw = om.infer(sp.obs_vars, series[i], [4, 9])
        assert_close(got, want, "fallback posterior, series %d" % i)
        assert_close(ll[i], llw, "fallback loglik, series %d" % i)
    b.close()
    m.close()
    m = gpu_lib.Model(fm, engine=3)
    b = m.batch(sp.obs_vars, series)
    with pytest.raises(gpu_lib.NipGpuError):
        b.infer([4, 9])
    b.close()
    m.close()


def test_factor_engine_follows_parameter_changes(gpu_lib, oracle_lib):
    """engine 3 works on factors extracted from the clique tables: nipgpu_model_set_parameters
    (new tables from the host: extraction + verification again) and nipgpu_em_mstep (factors =
    the normalised counts, on the device) must both be followed; tables that are NOT products of
    their families' CPTs are still served (the model drops to engine 1)"""
    from nip_b200.synth import FactorialSpec
    sp1, sp2 = FactorialSpec(6, 3, seed=4), FactorialSpec(6, 3, seed=9)
    fm1, fm2 = sp1.flat(), sp2.flat()
    data = sp1.sample(3, 3, seed=5, missing=0.2)
    data[:, 0, :] = np.abs(data[:, 0, :])
    series = [data[0], data[1][:2], data[2][:1]]
    query = [5, 10, 1]
    m = gpu_lib.Model(fm1, engine=0)
    assert m.engine == 3
    b = m.batch(sp1.obs_vars, series)

    def check(fm, what):
        om = oracle_lib.model(fm)
        post, ll = b.infer(query)
        for i, got in enumerate(b.split(post)):
            want, llw = om.infer(sp1.obs_vars, series[i], query)
            assert_close(got, want, "%s: posterior %d" % (what, i))
            assert_close(ll[i], llw, "%s: loglik %d" % (what, i))
        counts, L, st = b.estep()
        want, Lw, stw = om.estep(sp1.obs_vars, series)
        assert st == stw == 0
        assert_close(counts, want, what + ": expected counts")
        return counts

    check(fm1, "as created")
    m.set_parameters(fm2.clique_tables, fm2.var_prior)
    assert m.engine == 3
    counts = check(fm2, "after set_parameters")
    m.mstep(counts)                                 # factors <- normalised counts, on the device
    t, p = m.parameters()
    fm3 = sp2.flat()
    fm3.clique_tables[:] = t
    fm3.var_prior[:] = p
    check(fm3, "after an M-step")
    # a table that is no product of CPTs: multiply one entry of a big clique
    t2 = t.copy()
    big = int(np.argmax(np.diff(fm3.clique_tab_off)))
    t2[int(fm3.clique_tab_off[big]) + 7] *= 1.5
    m.set_parameters(t2, p)
    assert m.engine == 1
    fm4 = sp2.flat()
    fm4.clique_tables[:] = t2
    fm4.var_prior[:] = p
    check(fm4, "tables that do not factor (engine 1)")
    b.close()
    m.close()
