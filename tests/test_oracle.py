"""CPU suite: the C restatement (oracle/nip_oracle.c) against
  * the reference's own known answers (test/potentialtest.c, test/cliquetest.c,
    SURVEY.md Appendix C), restated here as data,
  * the golden fixtures generated from the reference itself (tests/golden/), and
  * the reference library side by side when oracle/_ref is available.
Bit-exact where the same additions happen in the same order (everything below).
"""
import ctypes as C

import numpy as np
import pytest

from cases import ALL_CASES, EM_CASES, LIKELIHOOD_CASES, SLICE_CASES, Case, unhex


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_potentialtest_known_answers(oracle_lib):
    """test/potentialtest.c:28-113 — 2x3x4x5x6 table holding 0..719, marginalised with
    mapping {1,3,2,0} into a 3x5x4x2 table (SURVEY Appendix C values)."""
    src_card = np.array([2, 3, 4, 5, 6], dtype=np.int32)
    # the test numbers the entries with the LAST index fastest (its nested loops), while
    # storage is first-index fastest (src/nippotential.c:58-68)
    src = np.zeros(720)
    for i in range(2):
        for j in range(3):
            for k in range(4):
                for l in range(5):
                    for m in range(6):
                        src[i + 2 * (j + 3 * (k + 4 * (l + 5 * m)))] = (((i * 3 + j) * 4 + k) * 5 + l) * 6 + m
    dst_card = np.array([3, 5, 4, 2], dtype=np.int32)
    mapping = np.array([1, 3, 2, 0], dtype=np.int32)
    dst = np.zeros(120)
    oracle_lib.L.orc_general_marginalise(_p(src), 5, _p(src_card), _p(dst), 4, _p(dst_card), _p(mapping))

    def at(i0, i1, i2, i3):
        return dst[i0 + 3 * (i1 + 5 * (i2 + 4 * i3))]
    # the printed order of the test is (a,b,c,d) with d fastest; its tuples index the
    # destination dimensions in order
    assert at(0, 0, 0, 0) == 15 and at(0, 0, 0, 1) == 2175
    assert at(0, 0, 1, 0) == 195 and at(0, 0, 1, 1) == 2355 and at(0, 0, 2, 0) == 375
    assert at(2, 4, 2, 1) == 4119 and at(2, 4, 3, 0) == 2139 and at(2, 4, 3, 1) == 4299
    assert dst.sum() == src.sum()
    # total_marginalise over dimension 2 (cardinality 4)
    out = np.zeros(4)
    oracle_lib.L.orc_total_marginalise(_p(src), 5, _p(src_card), _p(out), 2)
    ref = src.reshape(6, 5, 4, 3, 2).sum(axis=(0, 1, 3, 4))  # numpy view: last axis = dim 0
    assert np.array_equal(out, ref)


def test_update_rules(oracle_lib):
    """x/0 -> 0 in update_potential (src/nippotential.c:486-491) but 'skip' in
    update_evidence (:512-514); zero-sum normalise is a no-op (:354-355)."""
    card = np.array([2, 3], dtype=np.int32)
    sub = np.array([3], dtype=np.int32)
    mapping = np.array([1], dtype=np.int32)
    t = np.arange(1, 7, dtype=np.float64)
    num = np.array([2.0, 3.0, 5.0]); den = np.array([4.0, 0.0, 1.0])
    oracle_lib.L.orc_update_potential(_p(num), _p(den), 1, _p(sub), _p(t), 2, _p(card), _p(mapping))
    assert np.array_equal(t, [1 * 2 / 4, 2 * 2 / 4, 0, 0, 5 * 5, 6 * 5])
    t = np.arange(1, 7, dtype=np.float64)
    oracle_lib.L.orc_update_evidence(_p(num), _p(den), _p(t), 2, _p(card), 1)
    assert np.array_equal(t, [1 * 2 / 4, 2 * 2 / 4, 3 * 3, 4 * 3, 5 * 5, 6 * 5])
    z = np.zeros(4)
    oracle_lib.L.orc_normalise_array(_p(z), 4)
    assert np.array_equal(z, np.zeros(4))
    c = np.array([1.0, 3.0, 0.0, 0.0, 2.0, 2.0])
    oracle_lib.L.orc_normalise_cpd(_p(c), 6, 2)
    assert np.array_equal(c, [0.25, 0.75, 0, 0, 0.5, 0.5])


def test_appendix_c_model_net(oracle_lib):
    """examples/model.net, M1 = 2,3,2,3,2,4."""
    c = Case("model_net")
    m = oracle_lib.model(c.fm)
    post, ll = m.infer(c.obs_vars, c.series[0], [1])
    assert ll == -9.8679199821617463
    want = [[0.0082602323049822506, 0.77832532441191715, 0.21341444328310069, 0],
            [0, 0.50655316486499258, 0.49344417435194848, 2.6607830588873714e-06],
            [2.6607830588873714e-06, 0.47096885163067492, 0.52902848758626619, 0],
            [0, 0.18629700653919257, 0.81370299346080754, 0],
            [0, 0.1507121850654142, 0.84928781493458583, 0],
            [0, 0, 0.60033514591313608, 0.39966485408686386]]
    assert np.array_equal(post, np.array(want))
    post, ll = m.infer(c.obs_vars, c.series[0], [1], forward_only=True)
    assert ll == -9.8679199821617463
    assert np.array_equal(post[0], [0.12091503267973856, 0.81045751633986929, 0.068627450980392177, 0])
    assert np.array_equal(post[5], [0, 0, 0.60033514591313619, 0.39966485408686392])


def test_appendix_c_demo1(oracle_lib):
    c = Case("demo1_net")
    m = oracle_lib.model(c.fm)
    _, ll = m.infer(c.obs_vars, c.series[0], c.query)
    assert ll == -9.8096262260286338


def test_appendix_c_em(oracle_lib):
    """examples/empty.net, seed 1234, three iterations: curve, prior and first CPT entries."""
    c = Case("empty_net_em")
    m = oracle_lib.model(c.fm)
    counts = unhex(c.j["em"]["init"])
    assert np.array_equal(counts[:3], [0.2231180734108752, 0.21679622084684494, 0.44755924281084875])
    curve = []
    for _ in range(3):
        m.mstep(counts)
        counts, ll, st = m.estep(c.obs_vars, c.series)
        assert st == 0
        curve.append(ll / 72)
    # the model keeps the parameters of the 3rd M-step
    assert curve == [-1.5953246988199119, -1.5372731344499233, -1.5364557913908865]
    tables, prior = m.parameters()
    assert np.array_equal(prior, [0.23952055776429151, 0.24365651307737121, 0.25542881178161464,
                                  0.26139411737672263])
    assert np.array_equal(tables[:6], [0.31345421856208083, 0.35330201058319272, 0.14941931745525419,
                                       0.15988027582544151, 0.25654485213231976, 0.355966314017365])


@pytest.mark.parametrize("name", ALL_CASES)
def test_inference_golden(oracle_lib, name):
    c = Case(name)
    m = oracle_lib.model(c.fm)
    for kind, fwd in (("smooth", False), ("filter", True)):
        posts, lls = c.expected(kind)
        for s, want, ll_want in zip(c.series, posts, lls):
            post, ll = m.infer(c.obs_vars, s, c.query, forward_only=fwd)
            assert np.array_equal(post, want), (name, kind)
            assert ll == ll_want, (name, kind, ll, ll_want)


@pytest.mark.parametrize("name", EM_CASES)
def test_em_golden(oracle_lib, name):
    c = Case(name)
    m = oracle_lib.model(c.fm)
    counts = unhex(c.j["em"]["init"])
    for it in c.j["em"]["iters"]:
        cpt = m.mstep(counts)
        assert np.array_equal(cpt, unhex(it["cpt"]))
        tables, prior = m.parameters()
        assert np.array_equal(tables, unhex(it["tables"]))
        assert np.array_equal(prior, unhex(it["prior"]))
        counts, ll, st = m.estep(c.obs_vars, c.series)
        assert st == it["status"]
        if st == 0:
            assert np.array_equal(counts, unhex(it["counts"]))
            assert ll == float.fromhex(it["ll"])


@pytest.mark.parametrize("name", LIKELIHOOD_CASES)
def test_likelihood_golden(oracle_lib, name):
    c = Case(name)
    m = oracle_lib.model(c.fm)
    on = np.zeros(c.fm.n_vars, dtype=np.uint8)
    on[c.j["likelihood"]["marked"]] = 1
    for s, want in zip(c.series, c.j["likelihood"]["out"]):
        out = m.likelihood(c.obs_vars, s, 1 - on, on)
        assert np.array_equal(out.reshape(-1), unhex(want))


@pytest.mark.parametrize("name", SLICE_CASES)
def test_slice_api_golden(oracle_lib, name):
    c = Case(name)
    m = oracle_lib.model(c.fm)
    for step in c.j["slice"]:
        m.reset()
        m.use_priors(step["has_history"])
        for var, lik in step["evidence"]:
            m.enter_evidence(var, unhex(lik))
        m.make_consistent()
        assert m.mass() == float.fromhex(step["mass"])
        for v in range(c.fm.n_vars):
            assert np.array_equal(m.marginal(v), unhex(step["marginals"][v]))
        for k in range(c.fm.n_cliques):
            assert np.array_equal(m.clique(k), unhex(step["cliques"][k]))


def test_against_reference_side_by_side(oracle_lib, ref_lib, tmp_path):
    """fresh random models parsed by the reference: oracle == reference, bit for bit"""
    from nip_b200.synth import HmmSpec
    for S, M, seed in [(3, 2, 1), (9, 4, 2), (16, 5, 3)]:
        h = HmmSpec(S, M, seed=seed)
        p = tmp_path / ("h%d.net" % S)
        p.write_text(h.net_text())
        rm = ref_lib.parse(p)
        fm = rm.export()
        assert fm.structure_equal(h.flat())
        assert np.array_equal(fm.clique_tables, h.flat().clique_tables)
        assert np.array_equal(fm.var_prior, h.flat().var_prior)
        om = oracle_lib.model(fm)
        data = h.sample(4, 11, seed=seed + 10, missing=0.1)
        for s in data:
            ts = rm.timeseries(h.obs_vars, s)
            pr, lr = rm.infer(ts, [1, 2, 0])
            po, lo = om.infer(h.obs_vars, s, [1, 2, 0])
            assert np.array_equal(pr, po) and lr == lo


def test_factorial_generator_matches_reference_parse():
    """nip_b200.synth.FactorialSpec states the join tree the reference builds for the factorial
    family; for ns = 3 it must reproduce the golden fixture (generated from the reference's own
    parser + triangulation) bit for bit, tables included"""
    from nip_b200.synth import FactorialSpec
    g = Case("factorial4x3").fm
    f = FactorialSpec(3, 2, seed=15).flat()
    for k in ("var_card", "var_flags", "var_parent_off", "var_parents", "var_family", "var_prior_off",
              "var_prior", "clique_var_off", "clique_vars", "clique_tab_off", "clique_tables",
              "sepset_cliques", "sepset_var_off", "sepset_vars", "clique_adj_off", "clique_adj",
              "outgoing", "prev_outgoing"):
        assert np.array_equal(np.asarray(getattr(f, k)), np.asarray(getattr(g, k))), k
    assert (f.in_clique, f.out_clique) == (g.in_clique, g.out_clique)


@pytest.mark.parametrize("coupled", [True, False])
def test_factorial_generators_vs_reference_parser(ref_lib, coupled, tmp_path):
    """both factorial families (ring-coupled and uncoupled chains): the FlatModel stated by
    nip_b200.synth equals what the reference's parser + triangulation build for the same text"""
    from nip_b200.synth import FactorialSpec
    sp = FactorialSpec(3, 2, seed=15, coupled=coupled)
    p = tmp_path / "f.net"
    p.write_text(sp.net_text())
    g, f = ref_lib.parse(p).export(), sp.flat()
    for k in ("var_card", "var_flags", "var_parent_off", "var_parents", "var_family", "var_prior_off",
              "var_prior", "clique_var_off", "clique_vars", "clique_tab_off", "clique_tables",
              "sepset_cliques", "sepset_var_off", "sepset_vars", "clique_adj_off", "clique_adj",
              "outgoing", "prev_outgoing"):
        assert np.array_equal(np.asarray(getattr(f, k)), np.asarray(getattr(g, k))), k
    assert (f.in_clique, f.out_clique) == (g.in_clique, g.out_clique)
