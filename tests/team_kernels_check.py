"""Helper of test_gpu_parity.py::test_chain_kernel_flavours (run as a subprocess because the
library reads NIPGPU_CHAIN_TEAM / NIPGPU_CHAIN_PAIR once per process): smoothing, filtering and the
E-step of ragged HMM sets with missing data against the oracle at 1e-9, and a hash of the raw
results (a data race between the warps of a team would show as run-to-run differences)."""
import hashlib
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

O = OracleLib()
digest = hashlib.sha256()
for S, M, B, T in ((64, 8, 70, 9), (57, 3, 33, 12), (30, 4, 21, 7)):
    h = HmmSpec(S, M, seed=3)
    fm = h.flat()
    data = h.sample(B, T, seed=2, missing=0.1)
    data[:, 0, 0] = np.abs(data[:, 0, 0])
    rng = np.random.default_rng(1)
    series = [data[i, :int(rng.integers(1, T + 1))] for i in range(B)]
    om = O.model(fm)
    m = api.Model(fm)
    b = m.batch(h.obs_vars, series)
    post, ll = b.infer(h.hidden_query)
    fpost, fll = b.infer(h.hidden_query, forward_only=True)
    counts, L, st = b.estep()
    for i, (got, fgot) in enumerate(zip(b.split(post), b.split(fpost))):
        want, llw = om.infer(h.obs_vars, series[i], h.hidden_query)
        assert_close(got, want, "S=%d series %d smoothed" % (S, i))
        assert_close(ll[i], llw, "S=%d series %d loglik" % (S, i), atol=1e-12)
        want, llw = om.infer(h.obs_vars, series[i], h.hidden_query, forward_only=True)
        assert_close(fgot, want, "S=%d series %d filtered" % (S, i))
    want, Lw, stw = om.estep(h.obs_vars, series)
    assert st == stw == 0
    assert_close(counts, want, "S=%d expected counts" % S)
    assert_close(L, Lw, "S=%d EM loglik" % S)
    for a in (post, ll, fpost, counts):
        digest.update(np.ascontiguousarray(a).tobytes())
    b.close()
    m.close()
print("OK", digest.hexdigest())
