import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nip_b200.api as api
from nip_b200.synth import HmmSpec
h = HmmSpec(64, 32, seed=1); fm = h.flat()
m = api.Model(fm)
for B, T in [(4096, 200), (4096, 1000), (4096, 3000), (2048, 1000), (8192, 1000), (4736, 1000), (16384, 500)]:
    data = h.sample(B, T, seed=2)
    b = m.batch(h.obs_vars, data)
    res = []
    for kw in (dict(forward_only=True, want_ll=True), dict(forward_only=True, want_ll=False), dict(forward_only=False, want_ll=True)):
        for _ in range(3):
            b.infer_device(h.hidden_query, **kw)
        res.append(m.last_kernel_ms()[0])
    fwd_ll, fwd, both = res
    print("B=%5d T=%4d  filter+ll %.3f ms (%.3f us/step)  filter %.3f ms (%.3f us/step)  smooth %.3f ms -> bwd %.3f us/step   %.3e slice-steps/s" % (
        B, T, fwd_ll, fwd_ll * 1e3 / T, fwd, fwd * 1e3 / T, both, (both - fwd_ll) * 1e3 / T, B * T / (both * 1e-3)), flush=True)
    b.close()
