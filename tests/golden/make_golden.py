#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ from the REFERENCE ITSELF
(oracle/_ref/libnip_ref.so, i.e. /root/reference compiled by oracle/Makefile).

Run here (the container that has /root/reference); the JSON it writes is
committed so that the GPU box, which has no /root/reference, can check both the
C oracle and the CUDA path against genuine reference outputs.

    python tests/golden/make_golden.py

Every double is stored as a C99 hex float, so fixtures are bit exact.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from nip_b200.desc import save_json  # noqa: E402
from nip_b200.synth import HmmSpec, net_text_generic  # noqa: E402
from oracle.bindings import RefLib  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_EXAMPLES = "/root/reference/examples"


def hx(a):
    return [float.hex(float(x)) for x in np.asarray(a, dtype=np.float64).reshape(-1)]


def rand_series(rng, cards, n, tmin, tmax, missing):
    out = []
    for _ in range(n):
        T = int(rng.integers(tmin, tmax + 1))
        d = np.stack([rng.integers(0, c, size=T) for c in cards], axis=1).astype(np.int32)
        d[rng.random(d.shape) < missing] = -1
        out.append(d)
    return out


def observed_start(series):
    """EM fixtures: make the first slice of every series carry an observation.  When a
    series starts with an evidence-free slice the reference's running log-likelihood is
    log(m2) - log(m1) of two masses that differ only by rounding, and its BAD_LUCK test
    `ll > 0` (src/nip.c:1827-1831) then fires or not by chance; see DESIGN.md."""
    out = []
    for s in series:
        s = np.array(s, dtype=np.int32)
        if s.shape[0] and (s[0] < 0).all():
            s[0, 0] = 0
        out.append(s)
    return out


ONLY = None   # set from the command line: only=name1,name2 regenerates just those fixtures


def make_case(R, name, net_path, obs_names, series, query_names, em_seed=None, em_iters=3,
              likelihood_marked=None, slice_script=None):
    if ONLY is not None and name not in ONLY:
        return
    if em_seed is not None:
        series = observed_start([np.asarray(s).reshape(-1, len(obs_names)) for s in series])
    m = R.parse(net_path)
    fm = m.export()
    names = []
    # variable names in model order: recover through a parse of the text
    import re
    names = re.findall(r"^\s*(?:discrete\s+)?node\s+(\w+)", open(net_path).read(), flags=re.M)
    fm.var_names = names
    idx = {s: i for i, s in enumerate(names)}
    obs_vars = [idx[s] for s in obs_names]
    query = [idx[s] for s in query_names]
    case = {"name": name, "model": fm.to_json(), "obs_vars": obs_vars, "query": query,
            "series": [np.asarray(s, dtype=np.int32).reshape(-1, len(obs_vars)).tolist() for s in series]}
    ts = [m.timeseries(obs_vars, s) for s in series]
    sm, fl = [], []
    for t in ts:
        post, ll = m.infer(t, query)
        sm.append({"post": hx(post), "ll": float.hex(ll)})
        post, ll = m.infer(t, query, forward_only=True)
        fl.append({"post": hx(post), "ll": float.hex(ll)})
    case["smooth"], case["filter"] = sm, fl
    if likelihood_marked is not None:
        marked = [idx[s] for s in likelihood_marked]
        m.mark_all(False)
        for v in marked:
            m.mark(v, True)
        case["likelihood"] = {"marked": marked, "out": [hx(m.likelihood(t)) for t in ts]}
        m.mark_all(True)
    if slice_script is not None:
        # fine-grained API: reset / priors / soft evidence / make_consistent / mass / marginals
        res = []
        for step in slice_script:
            m.reset()
            m.use_priors(step["has_history"])
            for sym, lik in step["evidence"]:
                m.enter_evidence(idx[sym], lik)
            m.make_consistent()
            res.append({"has_history": step["has_history"],
                        "evidence": [[idx[s], hx(l)] for s, l in step["evidence"]],
                        "mass": float.hex(m.mass()),
                        "marginals": [hx(m.marginal(v)) for v in range(fm.n_vars)],
                        "cliques": [hx(m.clique(c, original=False)) for c in range(fm.n_cliques)]})
        m.reset()
        case["slice"] = res
    if em_seed is not None:
        # a second, fresh parse: EM mutates the model
        m2 = R.parse(net_path)
        ts2 = [m2.timeseries(obs_vars, s) for s in series]
        counts = m2.random_parameters(em_seed)
        em = {"seed": em_seed, "init": hx(counts), "iters": []}
        for _ in range(em_iters):
            cpt = m2.mstep(counts)
            tables, prior = m2.parameters()
            counts, ll, st = m2.estep(ts2)
            em["iters"].append({"cpt": hx(cpt), "tables": hx(tables), "prior": hx(prior),
                                "counts": hx(counts), "ll": float.hex(ll), "status": int(st)})
        case["em"] = em
    save_json(case, os.path.join(HERE, name + ".json"))
    print("wrote", name, "vars", names, "cliques", [fm.clique(c) for c in range(fm.n_cliques)])


def main():
    R = RefLib()
    rng = np.random.default_rng(20261018)
    tmp = tempfile.mkdtemp()

    def write(name, text):
        p = os.path.join(tmp, name)
        open(p, "w").write(text)
        return p

    # 1. examples/model.net: SURVEY Appendix C sequence + plausible and impossible series
    walk = []
    for _ in range(4):
        T = int(rng.integers(1, 12))
        x = int(rng.integers(0, 5))
        s = []
        for _t in range(T):
            x = int(np.clip(x + rng.integers(-1, 2), 0, 4))
            s.append(-1 if rng.random() < 0.15 else x)
        walk.append(np.array(s).reshape(-1, 1))
    series = [np.array([2, 3, 2, 3, 2, 4]).reshape(-1, 1)] + walk + [np.array([0, 4, 0, 4]).reshape(-1, 1),
                                                                     np.array([3]).reshape(-1, 1)]
    make_case(R, "model_net", os.path.join(REF_EXAMPLES, "model.net"), ["M1"], series, ["P1", "P0", "M1"],
              likelihood_marked=["M1"],
              slice_script=[{"has_history": 0, "evidence": []},
                            {"has_history": 0, "evidence": [["M1", [0, 0, 1, 0, 0]]]},
                            {"has_history": 1, "evidence": [["M1", [0.2, 0.5, 0.1, 0.1, 0.1]], ["P0", [0.5, 0.5, 0, 0]]]}])

    # 2. examples/demo1.net: general tree (in_clique != out_clique), missing data
    series = [np.array([[0, 0], [1, -1], [2, 1], [-1, 1], [1, 0]])] + rand_series(rng, [3, 2], 5, 1, 9, 0.25)
    make_case(R, "demo1_net", os.path.join(REF_EXAMPLES, "demo1.net"), ["A1", "B1"], series,
              ["A1", "B1", "C0", "C1", "D1"], em_seed=77, likelihood_marked=["A1"],
              slice_script=[{"has_history": 0, "evidence": [["A1", [1, 0, 0]]]},
                            {"has_history": 1, "evidence": [["B1", [0, 1]], ["D1", [0.3, 0.7]]]}])

    # 3. examples/empty.net: the EM known answer of SURVEY Appendix C (seed 1234, 3 iterations)
    S = ["233440440034430010011232", "312121222222344011010001", "044044100111100111012001"]
    series = [np.array([int(c) for c in s]).reshape(-1, 1) for s in S]
    make_case(R, "empty_net_em", os.path.join(REF_EXAMPLES, "empty.net"), ["M1"], series, ["P1"], em_seed=1234)

    # 4. child-first HMM, 5 states x 3 symbols (the layout of the benchmark models)
    h = HmmSpec(5, 3, seed=5)
    series = [d for d in h.sample(6, 9, seed=6, missing=0.2)] + [h.sample(1, 1, seed=8)[0], h.sample(1, 23, seed=9)[0]]
    make_case(R, "hmm5", write("hmm5.net", h.net_text()), ["M1"], series, ["P1"], em_seed=11,
              likelihood_marked=["M1"])

    # 5. HMM with 12 states (padding 12 -> 16 in the DMMA path) and two observation leaves
    r2 = np.random.default_rng(12)
    A = r2.random((12, 12)) + 0.02; A[r2.random((12, 12)) < 0.3] = 0; A += np.eye(12) * 0.1
    E1 = r2.random((12, 4)) + 0.05
    E2 = r2.random((12, 3)) + 0.05; E2[3] = [0, 1, 0]
    pi = r2.random(12) + 0.1
    text = net_text_generic(
        [("Y1", 4, None), ("Z1", 3, None), ("X1", 12, None), ("X0", 12, "X1")],
        [("Y1", ["X1"], E1), ("Z1", ["X1"], E2), ("X1", ["X0"], A), ("X0", [], pi[None, :])])
    series = rand_series(r2, [4, 3], 7, 1, 12, 0.2)
    make_case(R, "hmm12_two_leaves", write("hmm12.net", text), ["Y1", "Z1"], series, ["X1"], em_seed=5,
              likelihood_marked=["Y1"])

    # 6. two coupled chains (factorial-style): interface of two variables, per-chain observations
    r3 = np.random.default_rng(13)
    TA = r3.random((3, 3, 3)) + 0.05   # [B0][A0][A1]   A1 | A0 B0
    TB = r3.random((3, 3, 3)) + 0.05   # [A0][B0][B1]   B1 | B0 A0
    EA = r3.random((3, 2)) + 0.05
    EB = r3.random((3, 4)) + 0.05
    text = net_text_generic(
        [("YA", 2, None), ("YB", 4, None), ("A1", 3, None), ("B1", 3, None), ("A0", 3, "A1"), ("B0", 3, "B1")],
        [("YA", ["A1"], EA), ("YB", ["B1"], EB), ("A1", ["B0", "A0"], TA), ("B1", ["A0", "B0"], TB),
         ("A0", [], (r3.random(3) + 0.1)[None, :]), ("B0", [], (r3.random(3) + 0.1)[None, :])])
    series = rand_series(r3, [2, 4], 6, 1, 8, 0.15)
    make_case(R, "coupled2x3", write("coupled.net", text), ["YA", "YB"], series, ["A1", "B1", "A0", "YA"],
              em_seed=3, likelihood_marked=["YA", "YB"])

    # 6b. config C3's topology in small: 4 chains, ring-coupled X^i_t | X^i_{t-1}, X^{i-1}_{t-1},
    #     per-chain observations; the reference's triangulation yields 3^6-entry cliques
    r5 = np.random.default_rng(15)
    K, ns, ny = 4, 3, 2
    nodes = [("Y%d" % i, ny, None) for i in range(K)] + [("X%d" % i, ns, None) for i in range(K)] + \
            [("W%d" % i, ns, "X%d" % i) for i in range(K)]
    pots = [("Y%d" % i, ["X%d" % i], r5.random((ns, ny)) + 0.05) for i in range(K)]
    pots += [("X%d" % i, ["W%d" % ((i - 1) % K), "W%d" % i], r5.random((ns, ns, ns)) + 0.05) for i in range(K)]
    pots += [("W%d" % i, [], (r5.random(ns) + 0.1)[None, :]) for i in range(K)]
    text = net_text_generic(nodes, pots)
    series = rand_series(r5, [ny] * K, 5, 1, 6, 0.15)
    make_case(R, "factorial4x3", write("fact.net", text), ["Y%d" % i for i in range(K)], series,
              ["X0", "X3", "W1", "Y2"], em_seed=21, likelihood_marked=["Y0", "Y1"])

    # 7. no time-slice interface at all: independent slices (scalar alpha)
    r4 = np.random.default_rng(14)
    text = net_text_generic(
        [("U", 3, None), ("V", 2, None), ("W", 4, None)],
        [("V", ["U"], r4.random((3, 2)) + 0.1), ("W", ["V"], r4.random((2, 4)) + 0.1),
         ("U", [], (r4.random(3) + 0.1)[None, :])])
    series = rand_series(r4, [4], 4, 1, 5, 0.2)
    make_case(R, "no_interface", write("noif.net", text), ["W"], series, ["U", "V", "W"], em_seed=9)

    # 8. a hidden middle layer inside the slice: X0 -> X1 -> H1, Y1 | H1 X1, Z1 | H1
    r6 = np.random.default_rng(16)
    text = net_text_generic(
        [("Y1", 4, None), ("Z1", 2, None), ("H1", 2, None), ("X1", 3, None), ("X0", 3, "X1")],
        [("Y1", ["X1", "H1"], r6.random((3, 2, 4)) + 0.05), ("Z1", ["H1"], r6.random((2, 2)) + 0.05),
         ("H1", ["X1"], r6.random((3, 2)) + 0.05), ("X1", ["X0"], r6.random((3, 3)) + 0.05),
         ("X0", [], (r6.random(3) + 0.1)[None, :])])
    series = rand_series(r6, [4, 2], 6, 1, 9, 0.2)
    make_case(R, "two_layer", write("two_layer.net", text), ["Y1", "Z1"], series, ["X1", "H1", "Y1", "X0"],
              em_seed=31, likelihood_marked=["Y1"])

    # 9. two chains with independent dynamics but a SHARED observation Y1 | A1 B1 (explaining away)
    r7 = np.random.default_rng(17)
    text = net_text_generic(
        [("Y1", 3, None), ("U1", 2, None), ("A1", 2, None), ("B1", 3, None), ("A0", 2, "A1"), ("B0", 3, "B1")],
        [("Y1", ["B1", "A1"], r7.random((3, 2, 3)) + 0.05), ("U1", ["A1"], r7.random((2, 2)) + 0.05),
         ("A1", ["A0"], r7.random((2, 2)) + 0.05), ("B1", ["B0"], r7.random((3, 3)) + 0.05),
         ("A0", [], (r7.random(2) + 0.1)[None, :]), ("B0", [], (r7.random(3) + 0.1)[None, :])])
    series = rand_series(r7, [3, 2], 6, 1, 8, 0.2)
    make_case(R, "shared_obs", write("shared_obs.net", text), ["Y1", "U1"], series, ["A1", "B1", "B0", "Y1"],
              em_seed=32, likelihood_marked=["Y1", "U1"])

    # 10. structural zeros: a cyclic, almost deterministic chain and observations that rule states out
    A = np.array([[0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0.5, 0.5], [1, 0, 0, 0]], dtype=float)
    E = np.array([[1, 0, 0], [0.5, 0.5, 0], [0, 1, 0], [0, 0, 1]], dtype=float)
    text = net_text_generic(
        [("M1", 3, None), ("P1", 4, None), ("P0", 4, "P1")],
        [("M1", ["P1"], E), ("P1", ["P0"], A), ("P0", [], np.array([[0.5, 0.5, 0, 0]]))])
    series = [np.array(x).reshape(-1, 1) for x in
              ([0, 1, 1, 2, 0], [1, 1, 2, 0, 0, 1], [0, 0], [2, 2], [-1, 1, -1, 2], [0, 1, 2, 1], [1], [-1, -1, 0])]
    make_case(R, "structural_zeros", write("zeros.net", text), ["M1"], series, ["P1", "P0", "M1"],
              likelihood_marked=["M1"])


if __name__ == "__main__":
    for a in sys.argv[1:]:
        if a.startswith("only="):
            ONLY = set(a[5:].split(","))
    main()
