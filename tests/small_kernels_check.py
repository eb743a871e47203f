"""Helper of test_gpu_parity.py::test_thread_per_sequence_kernels (run as a subprocess with
NIPGPU_CHAIN_SMALL=1: the library reads it once per process).  Interfaces of 1..8 joint states on
the thread-per-sequence kernels (chain_small.cuh): smoothing, filtering and log-likelihood of
ragged sets with missing data against the oracle at 1e-9, queries that go through the joint
projection, the golden model.net fixture, and the E-step (DMMA kernels) on the same batch object
afterwards (both kernel families share the forward-row store)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import nip_b200.api as api
from cases import assert_close
from nip_b200.synth import HmmSpec
from oracle.bindings import OracleLib

assert os.environ.get("NIPGPU_CHAIN_SMALL") == "1"
O = OracleLib()
for S, M, B, T in ((1, 3, 5, 4), (2, 3, 40, 9), (3, 2, 33, 12), (4, 5, 150, 24), (5, 4, 21, 7), (6, 3, 64, 5),
                   (7, 2, 9, 11), (8, 6, 70, 9)):
    h = HmmSpec(S, M, seed=3 + S)
    fm = h.flat()
    data = h.sample(B, T, seed=2, missing=0.15)
    data[:, 0, 0] = np.abs(data[:, 0, 0])
    rng = np.random.default_rng(S)
    series = [data[i, :int(rng.integers(1, T + 1))] for i in range(B)]
    series[B // 2] = series[B // 2][:0] if S == 4 else series[B // 2]      # an empty series in the middle
    om = O.model(fm)
    m = api.Model(fm)
    b = m.batch(h.obs_vars, series)
    k0 = api.launch_count()
    post, ll = b.infer(h.hidden_query)
    fpost, fll = b.infer(h.hidden_query, forward_only=True)
    apost, all_ = b.infer(list(range(fm.n_vars)))           # every variable: joint rows + projection
    _, ll_only = b.infer([])
    counts, L, st = b.estep()
    post2, ll2 = b.infer(h.hidden_query)                    # after the E-step reused the row store
    assert np.array_equal(post, post2) and np.array_equal(ll, ll2)
    for i, (got, fgot, agot) in enumerate(zip(b.split(post), b.split(fpost), b.split(apost))):
        if len(series[i]) == 0:
            assert ll[i] == 0.0
            continue
        want, llw = om.infer(h.obs_vars, series[i], h.hidden_query)
        assert_close(got, want, "S=%d series %d smoothed" % (S, i))
        assert_close(ll[i], llw, "S=%d series %d loglik" % (S, i), atol=1e-12)
        assert_close(ll_only[i], llw, "S=%d series %d loglik (no query)" % (S, i), atol=1e-12)
        want, llw = om.infer(h.obs_vars, series[i], h.hidden_query, forward_only=True)
        assert_close(fgot, want, "S=%d series %d filtered" % (S, i))
        want, _ = om.infer(h.obs_vars, series[i], list(range(fm.n_vars)))
        assert_close(agot, want, "S=%d series %d all variables" % (S, i))
    nonempty = [s for s in series if len(s)]
    want, Lw, stw = om.estep(h.obs_vars, nonempty)
    assert st == stw == 0
    assert_close(counts, want, "S=%d expected counts" % S)
    b.close()
    m.close()
print("OK")
