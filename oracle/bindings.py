"""TEST INFRASTRUCTURE — ctypes bindings for the two CPU oracles.

  RefLib     oracle/_ref/libnip_ref.so : the reference itself (built by
             oracle/Makefile from /root/reference, prebuilt on the GPU box)
  OracleLib  oracle/liboracle.so       : the C restatement (oracle/nip_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl
reference` legs may import this module.  The product (nip_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from nip_b200.desc import FlatModel, ModelDesc

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libnip_ref.so")
REF_O0_SO = os.path.join(HERE, "_ref", "libnip_ref_O0.so")
ORACLE_SO = os.path.join(HERE, "liboracle.so")

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double


def build(verbose=False):
    """compile liboracle.so and, when /root/reference exists, oracle/_ref"""
    r = subprocess.run(["make", "-s", "-C", HERE, "all"], capture_output=True, text=True)
    if r.returncode != 0 or verbose:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_ref():
    return os.path.exists(REF_SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _mask(a, n):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.shape == (n,)
    return a


# --------------------------------------------------------------------------
class OracleLib:
    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build()
        L = self.L = C.CDLL(path)
        L.orc_model_new.restype = _vp
        L.orc_model_new.argtypes = [C.POINTER(ModelDesc)]
        L.orc_model_free.argtypes = [_vp]
        L.orc_counts_size.restype = C.c_long
        L.orc_counts_size.argtypes = [_vp]
        L.orc_prob_mass.restype = _d
        L.orc_prob_mass.argtypes = [_vp]
        for name, args in {
            "orc_infer": [_vp, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _vp, _vp],
            "orc_estep": [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp],
            "orc_mstep": [_vp, _vp],
            "orc_likelihood": [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp],
            "orc_reset_model": [_vp], "orc_total_reset": [_vp], "orc_use_priors": [_vp, _i],
            "orc_enter_evidence": [_vp, _i, _vp], "orc_enter_index_observation": [_vp, _i, _i],
            "orc_make_consistent": [_vp], "orc_marginal": [_vp, _i, _vp],
            "orc_get_clique": [_vp, _i, _i, _vp], "orc_get_parameters": [_vp, _vp, _vp],
            "orc_set_parameters": [_vp, _vp, _vp],
            "orc_general_marginalise": [_vp, _i, _vp, _vp, _i, _vp, _vp],
            "orc_total_marginalise": [_vp, _i, _vp, _vp, _i],
            "orc_update_potential": [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp],
            "orc_update_evidence": [_vp, _vp, _vp, _i, _vp, _i],
            "orc_normalise_array": [_vp, _i], "orc_normalise_cpd": [_vp, _i, _i],
        }.items():
            getattr(L, name).argtypes = args

    def model(self, fm: FlatModel) -> "OracleModel":
        return OracleModel(self, fm)


class OracleModel:
    """orc_model handle; one series per call like the reference API."""

    def __init__(self, lib: OracleLib, fm: FlatModel):
        self.lib, self.fm = lib, fm
        self._desc = fm.to_ctypes()
        self.h = lib.L.orc_model_new(C.byref(self._desc))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.L.orc_model_free(self.h)
            self.h = None

    def infer(self, obs_vars, data, query, forward_only=False, want_ll=True, use_evidence=None):
        fm = self.fm
        data = _i32(data).reshape(-1, len(obs_vars))
        T = data.shape[0]
        ov, q = _i32(obs_vars), _i32(query)
        row = int(sum(fm.var_card[v] for v in query))
        post = np.zeros((T, row))
        ll = C.c_double(0)
        m = _mask(use_evidence, fm.n_vars)
        self.lib.L.orc_infer(self.h, T, len(ov), _p(ov), _p(data), _p(m), len(q), _p(q),
                             int(forward_only), int(want_ll), _p(post), C.byref(ll))
        return post, ll.value

    def estep(self, obs_vars, series, use_evidence=None, counts=None):
        fm = self.fm
        n = int(self.lib.L.orc_counts_size(self.h))
        counts = np.ones(n) if counts is None else np.ascontiguousarray(counts, dtype=np.float64)
        ov = _i32(obs_vars)
        m = _mask(use_evidence, fm.n_vars)
        total, status = 0.0, 0
        for data in series:
            data = _i32(data).reshape(-1, len(ov))
            ll = C.c_double(0)
            status = self.lib.L.orc_estep(self.h, data.shape[0], len(ov), _p(ov), _p(data), _p(m),
                                          _p(counts), C.byref(ll))
            if status:
                break
            total += ll.value
        return counts, total, status

    def mstep(self, counts):
        counts = np.ascontiguousarray(counts, dtype=np.float64).copy()
        self.lib.L.orc_mstep(self.h, _p(counts))
        return counts

    def parameters(self):
        t = np.zeros(len(self.fm.clique_tables))
        p = np.zeros(len(self.fm.var_prior))
        self.lib.L.orc_get_parameters(self.h, _p(t), _p(p))
        return t, p

    def set_parameters(self, tables, prior):
        t = np.ascontiguousarray(tables, dtype=np.float64)
        p = np.ascontiguousarray(prior, dtype=np.float64)
        self.lib.L.orc_set_parameters(self.h, _p(t), _p(p))

    def likelihood(self, obs_vars, data, evidence_off, evidence_on):
        ov = _i32(obs_vars)
        data = _i32(data).reshape(-1, len(ov))
        out = np.zeros((data.shape[0], 2))
        self.lib.L.orc_likelihood(self.h, data.shape[0], len(ov), _p(ov), _p(data),
                                  _p(_mask(evidence_off, self.fm.n_vars)),
                                  _p(_mask(evidence_on, self.fm.n_vars)), _p(out))
        return out

    # fine-grained API
    def reset(self): self.lib.L.orc_reset_model(self.h)
    def use_priors(self, has_history): self.lib.L.orc_use_priors(self.h, int(has_history))
    def make_consistent(self): self.lib.L.orc_make_consistent(self.h)
    def mass(self): return self.lib.L.orc_prob_mass(self.h)

    def enter_evidence(self, var, lik):
        lik = np.ascontiguousarray(lik, dtype=np.float64)
        self.lib.L.orc_enter_evidence(self.h, int(var), _p(lik))

    def marginal(self, var):
        out = np.zeros(int(self.fm.var_card[var]))
        self.lib.L.orc_marginal(self.h, int(var), _p(out))
        return out

    def clique(self, c, original=False):
        out = np.zeros(int(self.fm.clique_tab_off[c + 1] - self.fm.clique_tab_off[c]))
        self.lib.L.orc_get_clique(self.h, int(c), int(original), _p(out))
        return out


# --------------------------------------------------------------------------
class RefLib:
    """The reference library.  One model per parse; models mutate in place."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with `make -C oracle`; needs /root/reference)")
        L = self.L = C.CDLL(path)
        L.refh_parse_model.restype = _vp
        L.refh_parse_model.argtypes = [C.c_char_p]
        L.refh_free_model.argtypes = [_vp]
        L.refh_export.restype = C.POINTER(ModelDesc)
        L.refh_export.argtypes = [_vp]
        L.refh_free_desc.argtypes = [C.POINTER(ModelDesc)]
        L.refh_mark_all.argtypes = [_vp, _i]
        L.refh_mark_var.argtypes = [_vp, _i, _i]
        L.refh_new_timeseries.restype = _vp
        L.refh_new_timeseries.argtypes = [_vp, _i, _vp, _i, _vp]
        L.refh_free_timeseries.argtypes = [_vp]
        L.refh_infer.argtypes = [_vp, _i, _vp, _i, _i, _vp, _vp]
        L.refh_counts_size.restype = C.c_long
        L.refh_counts_size.argtypes = [_vp]
        L.refh_mstep.argtypes = [_vp, _vp]
        L.refh_estep.argtypes = [_vp, _i, _vp, _vp]
        L.refh_em_learn.argtypes = [_vp, _i, _d, C.c_long, _vp, _i, _vp]
        L.refh_random_parameters.argtypes = [_vp, C.c_long, _vp]
        L.refh_likelihood.argtypes = [_vp, _vp]
        L.refh_clique_size.argtypes = [_vp, _i]
        L.refh_get_clique.argtypes = [_vp, _i, _i, _vp]
        L.refh_get_prior.argtypes = [_vp, _i, _vp]
        L.refh_enter_evidence.argtypes = [_vp, _i, _vp]
        L.refh_marginal.argtypes = [_vp, _i, _vp]
        L.refh_next_slice.argtypes = [_vp]
        L.refh_sepset_size.argtypes = [_vp, _i]
        L.refh_get_sepset.argtypes = [_vp, _i, _i, _vp]
        L.refh_time_infer.restype = _d
        L.refh_time_infer.argtypes = [_vp, _i, _i, _vp, _i, _i]
        L.refh_time_em_iteration.restype = _d
        L.refh_time_em_iteration.argtypes = [_vp, _i, _vp, _vp]
        for f in ("reset_model", "make_consistent", "total_reset"):
            getattr(L, f).argtypes = [_vp]
        L.use_priors.argtypes = [_vp, _i]
        L.model_prob_mass.restype = _d
        L.model_prob_mass.argtypes = [_vp]

    def parse(self, path) -> "RefModel":
        h = self.L.refh_parse_model(str(path).encode())
        if not h:
            raise RuntimeError("reference parse_model failed for %s" % path)
        return RefModel(self, h)


class RefModel:
    def __init__(self, lib: RefLib, h):
        self.lib, self.h = lib, h
        lib.L.refh_mark_all(h, 1)
        self._ts = []
        self._T = {}

    def export(self) -> FlatModel:
        d = self.lib.L.refh_export(self.h)
        fm = FlatModel.from_ctypes(d.contents)
        self.lib.L.refh_free_desc(d)
        return fm

    def mark_all(self, on=True): self.lib.L.refh_mark_all(self.h, int(on))
    def mark(self, var, on=True): self.lib.L.refh_mark_var(self.h, int(var), int(on))

    def timeseries(self, obs_vars, data):
        ov = _i32(obs_vars)
        data = _i32(data).reshape(-1, len(ov))
        ts = self.lib.L.refh_new_timeseries(self.h, len(ov), _p(ov), data.shape[0], _p(data))
        self._ts.append(ts)
        self._T[ts] = data.shape[0]
        return ts

    def infer(self, ts, query, forward_only=False, want_ll=True):
        fm_card = self._cards()
        q = _i32(query)
        row = int(sum(fm_card[v] for v in query))
        post = np.zeros((self._T[ts], row))
        ll = C.c_double(0)
        r = self.lib.L.refh_infer(ts, len(q), _p(q), int(forward_only), int(want_ll), _p(post),
                                  C.byref(ll))
        if r:
            raise RuntimeError("reference inference failed (%d)" % r)
        return post, ll.value

    def _cards(self):
        if not hasattr(self, "_card_cache"):
            self._card_cache = self.export().var_card
        return self._card_cache

    def counts_size(self): return int(self.lib.L.refh_counts_size(self.h))

    def mstep(self, counts):
        c = np.ascontiguousarray(counts, dtype=np.float64).copy()
        self.lib.L.refh_mstep(self.h, _p(c))
        return c

    def estep(self, ts_list):
        arr = (C.c_void_p * len(ts_list))(*ts_list)
        counts = np.zeros(self.counts_size())
        ll = C.c_double(0)
        st = self.lib.L.refh_estep(arr, len(ts_list), _p(counts), C.byref(ll))
        return counts, ll.value, st

    def em_learn(self, ts_list, threshold, seed, cap=4096):
        arr = (C.c_void_p * len(ts_list))(*ts_list)
        curve = np.zeros(cap)
        n = C.c_int(0)
        st = self.lib.L.refh_em_learn(arr, len(ts_list), float(threshold), int(seed), _p(curve), cap,
                                      C.byref(n))
        return st, curve[:min(n.value, cap)].copy()

    def random_parameters(self, seed):
        c = np.zeros(self.counts_size())
        self.lib.L.refh_random_parameters(self.h, int(seed), _p(c))
        return c

    def likelihood(self, ts):
        out = np.zeros((self._T[ts], 2))
        self.lib.L.refh_likelihood(ts, _p(out))
        return out

    def clique(self, c, original=True):
        out = np.zeros(self.lib.L.refh_clique_size(self.h, int(c)))
        self.lib.L.refh_get_clique(self.h, int(c), int(original), _p(out))
        return out

    def prior(self, var):
        out = np.zeros(int(self._cards()[var]))
        self.lib.L.refh_get_prior(self.h, int(var), _p(out))
        return out

    def parameters(self):
        fm = self.export()
        return fm.clique_tables.copy(), fm.var_prior.copy()

    # fine-grained API straight through nip.h
    def reset(self): self.lib.L.reset_model(self.h)
    def use_priors(self, has_history): self.lib.L.use_priors(self.h, int(has_history))
    def make_consistent(self): self.lib.L.make_consistent(self.h)
    def mass(self): return self.lib.L.model_prob_mass(self.h)

    def enter_evidence(self, var, lik):
        lik = np.ascontiguousarray(lik, dtype=np.float64)
        self.lib.L.refh_enter_evidence(self.h, int(var), _p(lik))

    def marginal(self, var):
        out = np.zeros(int(self._cards()[var]))
        self.lib.L.refh_marginal(self.h, int(var), _p(out))
        return out

    def next_slice(self):
        """generate_data's slice hand-over: alpha out, reset, priors with history, alpha in"""
        if self.lib.L.refh_next_slice(self.h):
            raise RuntimeError("reference timeslice message pass failed")

    def sepset(self, s, new=True):
        out = np.zeros(self.lib.L.refh_sepset_size(self.h, int(s)))
        self.lib.L.refh_get_sepset(self.h, int(s), int(new), _p(out))
        return out

    def tree_state(self):
        """every clique->p, then every sepset's new and old potential, concatenated"""
        fm = self.export()
        parts = [self.clique(c, original=False) for c in range(fm.n_cliques)]
        parts += [self.sepset(s, True) for s in range(fm.n_sepsets)]
        parts += [self.sepset(s, False) for s in range(fm.n_sepsets)]
        return np.concatenate(parts) if parts else np.zeros(0)

    def time_infer(self, ts_list, query, want_ll=True, nproc=1):
        arr = (C.c_void_p * len(ts_list))(*ts_list)
        q = _i32(query)
        return self.lib.L.refh_time_infer(arr, len(ts_list), len(q), _p(q), int(want_ll), int(nproc))

    def time_em_iteration(self, ts_list, counts):
        arr = (C.c_void_p * len(ts_list))(*ts_list)
        c = np.ascontiguousarray(counts, dtype=np.float64).copy()
        ll = C.c_double(0)
        s = self.lib.L.refh_time_em_iteration(arr, len(ts_list), _p(c), C.byref(ll))
        return s, c, ll.value
