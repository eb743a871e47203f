"""ctypes binding of the C ABI (include/nipgpu.h) — plumbing for the Python test
and benchmark harness.  The product is the shared library; this module adds
nothing to the computation.

There is no CPU fallback: if `libnipgpu.so` is missing, or no B200-class device
is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .desc import FlatModel, ModelDesc

PKG = os.path.dirname(os.path.abspath(__file__))
# NIPGPU_LIB_PATH: development only (A/B timing of kernel variants built side by side)
LIB_PATH = os.environ.get("NIPGPU_LIB_PATH") or os.path.join(PKG, "libnipgpu.so")

ENGINE_AUTO, ENGINE_JTREE, ENGINE_CHAIN, ENGINE_FACTOR = 0, 1, 2, 3
EBADLUCK = 8

# every symbol include/nipgpu.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "nipgpu_last_error", "nipgpu_device_check", "nipgpu_model_create", "nipgpu_model_destroy",
    "nipgpu_model_engine", "nipgpu_model_factorable", "nipgpu_model_set_parameters", "nipgpu_model_get_parameters",
    "nipgpu_batch_create", "nipgpu_batch_destroy", "nipgpu_batch_update", "nipgpu_infer", "nipgpu_infer_device",
    "nipgpu_em_estep", "nipgpu_model_counts_size", "nipgpu_model_counts_offsets",
    "nipgpu_em_counts_device", "nipgpu_em_mstep", "nipgpu_likelihood", "nipgpu_slice_reset",
    "nipgpu_slice_use_priors", "nipgpu_slice_enter_prior", "nipgpu_slice_enter_evidence", "nipgpu_slice_get_sepset", "nipgpu_slice_make_consistent",
    "nipgpu_slice_mass", "nipgpu_slice_marginal", "nipgpu_slice_get_clique",
    "nipgpu_launch_count", "nipgpu_last_kernel_ms", "nipgpu_last_forward_ms", "nipgpu_jt_trace", "nipgpu_sample", "nipgpu_model_stream", "nipgpu_probe_peaks",
    "nipgpu_probe_dmma_chain", "nipgpu_probe_sweep", "nipgpu_probe_dmma_dfma", "nipgpu_slice_propagate", "nipgpu_group_create", "nipgpu_group_destroy",
    "nipgpu_group_size", "nipgpu_group_em_estep", "nipgpu_group_em_mstep",
]

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_lib = None


class NipGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nipgpu error %d: %s" % (code, msg))
        self.code = code


def load_library(path=LIB_PATH):
    """dlopen the device library and declare prototypes; raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s is missing: build it with `python -m nip_b200.build` "
            "(there is no CPU fallback)" % path)
    L = C.CDLL(path)
    L.nipgpu_last_error.restype = C.c_char_p
    L.nipgpu_device_check.argtypes = [_i]
    L.nipgpu_model_create.argtypes = [C.POINTER(ModelDesc), _i, _i, C.POINTER(_vp)]
    L.nipgpu_model_destroy.argtypes = [_vp]
    L.nipgpu_model_destroy.restype = None
    L.nipgpu_model_engine.argtypes = [_vp]
    L.nipgpu_model_factorable.argtypes = [_vp]
    L.nipgpu_model_set_parameters.argtypes = [_vp, _vp, _vp]
    L.nipgpu_model_get_parameters.argtypes = [_vp, _vp, _vp]
    L.nipgpu_batch_create.argtypes = [_vp, _i, _vp, _i, _vp, _vp, C.POINTER(_vp)]
    L.nipgpu_batch_destroy.argtypes = [_vp]
    L.nipgpu_batch_destroy.restype = None
    L.nipgpu_batch_update.argtypes = [_vp, _vp]
    L.nipgpu_infer.argtypes = [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp]
    L.nipgpu_infer_device.argtypes = [_vp, _vp, _vp, _i, _vp, _i, _i, C.POINTER(_vp), C.POINTER(_vp)]
    L.nipgpu_em_estep.argtypes = [_vp, _vp, _vp, _i, _vp, C.POINTER(_d), C.POINTER(_i)]
    L.nipgpu_model_counts_size.restype = C.c_int64
    L.nipgpu_model_counts_size.argtypes = [_vp]
    L.nipgpu_model_counts_offsets.argtypes = [_vp, _vp]
    L.nipgpu_em_counts_device.argtypes = [_vp, C.POINTER(_vp), C.POINTER(C.c_int64)]
    L.nipgpu_em_mstep.argtypes = [_vp, _vp]
    L.nipgpu_likelihood.argtypes = [_vp, _vp, _vp, _vp, _vp]
    L.nipgpu_slice_reset.argtypes = [_vp]
    L.nipgpu_slice_use_priors.argtypes = [_vp, _i]
    L.nipgpu_slice_enter_evidence.argtypes = [_vp, _i, _vp]
    L.nipgpu_slice_enter_prior.argtypes = [_vp, _i]
    L.nipgpu_slice_get_sepset.argtypes = [_vp, _i, _vp]
    L.nipgpu_slice_make_consistent.argtypes = [_vp]
    L.nipgpu_slice_mass.argtypes = [_vp, C.POINTER(_d)]
    L.nipgpu_slice_marginal.argtypes = [_vp, _i, _vp]
    L.nipgpu_slice_get_clique.argtypes = [_vp, _i, _vp]
    L.nipgpu_launch_count.restype = C.c_int64
    L.nipgpu_launch_count.argtypes = [_i]
    L.nipgpu_last_kernel_ms.argtypes = [_vp, C.POINTER(_d), C.POINTER(C.c_int32)]
    L.nipgpu_last_forward_ms.argtypes = [_vp, C.POINTER(_d)]
    L.nipgpu_jt_trace.argtypes = [_vp, _vp, _i, _i]
    L.nipgpu_sample.argtypes = [_vp, _i, _i, C.c_uint64, _vp]
    L.nipgpu_probe_peaks.argtypes = [_i, C.POINTER(_d), C.POINTER(_d), C.POINTER(_d)]
    L.nipgpu_probe_dmma_chain.argtypes = [_i, _i, _i, C.POINTER(_d)]
    L.nipgpu_probe_sweep.argtypes = [_i, _i, C.POINTER(_d)]
    L.nipgpu_probe_dmma_dfma.argtypes = [_i, _i, _i, C.POINTER(_d)]
    L.nipgpu_slice_propagate.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]
    L.nipgpu_group_create.argtypes = [_vp, _i, C.POINTER(_vp)]
    L.nipgpu_group_destroy.argtypes = [_vp]
    L.nipgpu_group_destroy.restype = None
    L.nipgpu_group_size.argtypes = [_vp]
    L.nipgpu_group_em_estep.argtypes = [_vp, _vp, _vp, _i, _vp, C.POINTER(_d), C.POINTER(_i)]
    L.nipgpu_group_em_mstep.argtypes = [_vp, _vp]
    L.nipgpu_model_stream.restype = _vp
    L.nipgpu_model_stream.argtypes = [_vp]
    _lib = L
    return L


def _check(code):
    if code != 0:
        raise NipGpuError(code, (_lib.nipgpu_last_error() or b"").decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _mask(a, n):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.shape == (n,)
    return a


def probe_peaks(device=0):
    """(DMMA TFLOP/s, DFMA TFLOP/s, copy GB/s) measured live on `device`"""
    a, b, c = _d(), _d(), _d()
    _check(load_library().nipgpu_probe_peaks(int(device), C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def factorable(fm: FlatModel):
    """host-only: can engine 3 serve this model (every clique table = product of its families' CPTs)?
    Returns (bool, reason)."""
    L = load_library()
    desc = fm.to_ctypes()
    rc = int(L.nipgpu_model_factorable(C.byref(desc)))
    if rc < 0:
        raise NipGpuError(-rc, (L.nipgpu_last_error() or b"").decode())
    return rc == 1, ("" if rc == 1 else (L.nipgpu_last_error() or b"").decode())


def launch_count(reset=False):
    return int(load_library().nipgpu_launch_count(int(reset)))


class Model:
    """nipgpu_model handle."""

    def __init__(self, fm: FlatModel, device=0, engine=ENGINE_AUTO):
        self.L = load_library()
        self.fm = fm
        self._desc = fm.to_ctypes()
        h = _vp()
        _check(self.L.nipgpu_model_create(C.byref(self._desc), int(device), int(engine), C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.nipgpu_model_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def engine(self):
        return int(self.L.nipgpu_model_engine(self.h))

    def batch(self, obs_vars, series) -> "Batch":
        return Batch(self, obs_vars, series)

    def set_parameters(self, tables, prior):
        t = np.ascontiguousarray(tables, dtype=np.float64)
        p = np.ascontiguousarray(prior, dtype=np.float64)
        _check(self.L.nipgpu_model_set_parameters(self.h, _p(t), _p(p)))

    def parameters(self):
        t = np.zeros(len(self.fm.clique_tables))
        p = np.zeros(max(len(self.fm.var_prior), 1))
        _check(self.L.nipgpu_model_get_parameters(self.h, _p(t), _p(p)))
        return t, p[:len(self.fm.var_prior)]

    def counts_size(self):
        return int(self.L.nipgpu_model_counts_size(self.h))

    def counts_offsets(self):
        """[n_vars + 1] offsets of every variable's family counts inside the E-step vector"""
        off = np.zeros(self.fm.n_vars + 1, dtype=np.int64)
        _check(self.L.nipgpu_model_counts_offsets(self.h, _p(off)))
        return off

    def mstep(self, counts=None):
        c = None if counts is None else np.ascontiguousarray(counts, dtype=np.float64)
        _check(self.L.nipgpu_em_mstep(self.h, _p(c)))

    def counts_device(self):
        ptr, n = _vp(), C.c_int64()
        _check(self.L.nipgpu_em_counts_device(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, int(n.value)

    def last_kernel_ms(self):
        ms, n = _d(), C.c_int32()
        _check(self.L.nipgpu_last_kernel_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_forward_ms(self):
        ms = _d()
        _check(self.L.nipgpu_last_forward_ms(self.h, C.byref(ms)))
        return ms.value

    def sample(self, n_series, length, seed=1):
        """[n_series, length, n_vars] int32 states drawn from the model (chain-structured models)"""
        out = np.zeros((n_series, length, self.fm.n_vars), dtype=np.int32)
        _check(self.L.nipgpu_sample(self.h, int(n_series), int(length), int(seed), _p(out)))
        return out

    def jt_trace(self, reset=True, cap=8192):
        """(tag, ns) per grid barrier of the generic engine's grid team (NIPGPU_JT_TRACE=1)"""
        buf = np.zeros(2 * cap, dtype=np.uint64)
        n = self.L.nipgpu_jt_trace(self.h, _p(buf), cap, int(reset))
        if n < 0:
            _check(1)
        return buf[:2 * n].reshape(-1, 2)

    # ---- single-slice API ----
    def slice_reset(self): _check(self.L.nipgpu_slice_reset(self.h))
    def slice_use_priors(self, has_history): _check(self.L.nipgpu_slice_use_priors(self.h, int(has_history)))
    def slice_make_consistent(self): _check(self.L.nipgpu_slice_make_consistent(self.h))

    def slice_enter_evidence(self, var, lik):
        lik = np.ascontiguousarray(lik, dtype=np.float64)
        _check(self.L.nipgpu_slice_enter_evidence(self.h, int(var), _p(lik)))

    def slice_mass(self):
        m = _d()
        _check(self.L.nipgpu_slice_mass(self.h, C.byref(m)))
        return m.value

    def slice_marginal(self, var):
        out = np.zeros(int(self.fm.var_card[var]))
        _check(self.L.nipgpu_slice_marginal(self.h, int(var), _p(out)))
        return out

    def slice_propagate(self, tables, sepsets):
        """make_consistent on the caller's tree state: (tables, sepset new) -> (tables, new, old)"""
        t = np.ascontiguousarray(tables, dtype=np.float64)
        sp = np.ascontiguousarray(sepsets, dtype=np.float64)
        out, new, old = np.zeros_like(t), np.zeros_like(sp), np.zeros_like(sp)
        _check(self.L.nipgpu_slice_propagate(self.h, _p(t), _p(sp) if sp.size else None, _p(out),
                                             _p(new) if sp.size else None, _p(old) if sp.size else None))
        return out, new, old

    def slice_clique(self, c):
        out = np.zeros(int(self.fm.clique_tab_off[c + 1] - self.fm.clique_tab_off[c]))
        _check(self.L.nipgpu_slice_get_clique(self.h, int(c), _p(out)))
        return out


class Group:
    """nipgpu_group handle: the same model on several devices of one box, one batch (shard) per
    member; EM sums the expected counts with one ncclAllReduce per iteration."""

    def __init__(self, models):
        self.L = load_library()
        self.models = list(models)
        arr = (_vp * len(self.models))(*[m.h for m in self.models])
        h = _vp()
        _check(self.L.nipgpu_group_create(arr, len(self.models), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.nipgpu_group_destroy(self.h)
            self.h = None

    __del__ = close

    def estep(self, batches, use_evidence=None, add_pseudocount=True, want_counts=True):
        n = self.models[0].counts_size()
        counts = np.zeros(n) if want_counts else None
        ll, st = _d(), _i()
        arr = (_vp * len(batches))(*[b.h for b in batches])
        m = _mask(use_evidence, self.models[0].fm.n_vars)
        _check(self.L.nipgpu_group_em_estep(self.h, arr, _p(m), int(add_pseudocount), _p(counts),
                                            C.byref(ll), C.byref(st)))
        return counts, ll.value, st.value

    def mstep(self, counts=None):
        c = None if counts is None else np.ascontiguousarray(counts, dtype=np.float64)
        _check(self.L.nipgpu_group_em_mstep(self.h, _p(c)))


class Batch:
    """nipgpu_batch handle: a set of (possibly ragged) series resident in HBM."""

    def __init__(self, model: Model, obs_vars, series):
        self.model, self.L = model, model.L
        self.obs_vars = np.ascontiguousarray(obs_vars, dtype=np.int32)
        n_obs = len(self.obs_vars)
        if isinstance(series, np.ndarray) and series.ndim == 3:
            self.lengths = np.full(series.shape[0], series.shape[1], dtype=np.int32)
            data = np.ascontiguousarray(series, dtype=np.int32).reshape(-1, max(n_obs, 1))
        else:
            mats = [np.asarray(s, dtype=np.int32).reshape(-1, max(n_obs, 1)) for s in series]
            self.lengths = np.array([m.shape[0] for m in mats], dtype=np.int32)
            data = np.concatenate(mats, axis=0) if mats else np.zeros((0, max(n_obs, 1)), dtype=np.int32)
            data = np.ascontiguousarray(data, dtype=np.int32)
        self.rows = int(self.lengths.sum())
        self.row_off = np.concatenate([[0], np.cumsum(self.lengths)]).astype(np.int64)
        self.h2d_bytes = data.nbytes + self.lengths.nbytes
        h = _vp()
        _check(self.L.nipgpu_batch_create(model.h, len(self.lengths), _p(self.lengths), n_obs,
                                          _p(self.obs_vars), _p(data), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.model, "h", None):
            self.L.nipgpu_batch_destroy(self.h)
        self.h = None

    __del__ = close

    def update(self, data):
        """re-upload observations (int32, same shape); `data` may live in pinned memory"""
        assert data.dtype == np.int32 and data.size == self.rows * max(len(self.obs_vars), 1)
        _check(self.L.nipgpu_batch_update(self.h, _p(data)))

    def _row(self, query):
        return int(sum(self.model.fm.var_card[v] for v in query))

    def infer(self, query, forward_only=False, want_ll=True, use_evidence=None, out=None, ll_out=None):
        """host results: (post [rows, sum card], loglik [n_series])"""
        q = np.ascontiguousarray(query, dtype=np.int32)
        post = out if out is not None else np.zeros((self.rows, self._row(query)))
        ll = (ll_out if ll_out is not None else np.zeros(len(self.lengths))) if want_ll else None
        m = _mask(use_evidence, self.model.fm.n_vars)
        _check(self.L.nipgpu_infer(self.model.h, self.h, _p(m), len(q), _p(q), int(forward_only),
                                   _p(post) if len(q) else None, _p(ll)))
        return post, ll

    def infer_device(self, query, forward_only=False, want_ll=True, use_evidence=None):
        """results stay in HBM; returns (post_ptr, ll_ptr) device addresses"""
        q = np.ascontiguousarray(query, dtype=np.int32)
        m = _mask(use_evidence, self.model.fm.n_vars)
        pp, lp = _vp(), _vp()
        _check(self.L.nipgpu_infer_device(self.model.h, self.h, _p(m), len(q), _p(q), int(forward_only),
                                          int(want_ll), C.byref(pp), C.byref(lp)))
        return pp.value, lp.value

    def estep(self, use_evidence=None, add_pseudocount=True, want_counts=True):
        n = self.model.counts_size()
        counts = np.zeros(n) if want_counts else None
        ll, st = _d(), _i()
        m = _mask(use_evidence, self.model.fm.n_vars)
        _check(self.L.nipgpu_em_estep(self.model.h, self.h, _p(m), int(add_pseudocount), _p(counts),
                                      C.byref(ll), C.byref(st)))
        return counts, ll.value, st.value

    def likelihood(self, evidence_off, evidence_on):
        out = np.zeros((self.rows, 2))
        n = self.model.fm.n_vars
        _check(self.L.nipgpu_likelihood(self.model.h, self.h, _p(_mask(evidence_off, n)),
                                        _p(_mask(evidence_on, n)), _p(out)))
        return out

    def split(self, flat):
        """[rows, k] -> list of per-series arrays"""
        return [flat[self.row_off[i]:self.row_off[i + 1]] for i in range(len(self.lengths))]
