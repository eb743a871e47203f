"""Golden fixtures (tests/golden/*.json) decoded into numpy."""
import os

import numpy as np

from nip_b200.desc import FlatModel, load_json

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL_CASES = ["model_net", "demo1_net", "empty_net_em", "hmm5", "hmm12_two_leaves", "coupled2x3",
             "factorial4x3", "no_interface", "two_layer", "shared_obs", "structural_zeros"]
EM_CASES = ["demo1_net", "empty_net_em", "hmm5", "hmm12_two_leaves", "coupled2x3", "factorial4x3",
            "no_interface", "two_layer", "shared_obs"]
LIKELIHOOD_CASES = ["model_net", "demo1_net", "hmm5", "hmm12_two_leaves", "coupled2x3", "factorial4x3",
                    "two_layer", "shared_obs", "structural_zeros"]
SLICE_CASES = ["model_net", "demo1_net"]
RTOL = 1e-9   # north star: 1e-9 relative on posteriors, log-likelihood, re-estimated CPTs


def unhex(a):
    return np.array([float.fromhex(x) for x in a], dtype=np.float64)


class Case:
    def __init__(self, name):
        j = load_json(os.path.join(GOLDEN, name + ".json"))
        self.name = name
        self.j = j
        self.fm = FlatModel.from_json(j["model"])
        self.obs_vars = j["obs_vars"]
        self.query = j["query"]
        n_obs = len(self.obs_vars)
        self.series = [np.array(s, dtype=np.int32).reshape(-1, n_obs) for s in j["series"]]
        self.row = int(sum(self.fm.var_card[v] for v in self.query))

    def expected(self, kind):
        """kind: 'smooth' | 'filter' -> (list of post arrays, ll array)"""
        posts = [unhex(e["post"]).reshape(-1, self.row) for e in self.j[kind]]
        ll = np.array([float.fromhex(e["ll"]) for e in self.j[kind]])
        return posts, ll


def close(a, b, rtol=RTOL, atol=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.allclose(a, b, rtol=rtol, atol=atol)


def assert_close(a, b, what="", rtol=RTOL, atol=0.0):
    """relative 1e-9; `atol` only where the reference value is an exact 0 or a
    difference of nearly equal numbers (stated at the call site)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = rtol * np.abs(b) + atol
    bad = err > tol
    if bad.any():
        i = int(np.argmax(err - tol))
        raise AssertionError("%s: %d/%d entries differ, worst at %d: got %r want %r"
                             % (what, int(bad.sum()), a.size, i, a.reshape(-1)[i], b.reshape(-1)[i]))
