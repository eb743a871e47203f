"""mid-size cliques: per-CTA HBM workspace vs the grid team's concurrent groups"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nip_b200.api as api
from nip_b200.synth import FactorialSpec
for ns, B, T in ((6, 600, 6), (7, 600, 6), (8, 600, 6)):
    sp = FactorialSpec(ns, 3, seed=1)
    fm = sp.flat()
    data = sp.sample(B, T, seed=2)
    for mode in ("hbm", "grid"):
        os.environ["NIPGPU_JT_MODE"] = mode
        m = api.Model(fm, engine=1)
        b = m.batch(sp.obs_vars, data)
        for rep in range(2):
            b.infer_device([4, 5])
        ms = m.last_kernel_ms()[0]
        print("ns=%d (%d-entry cliques) %s: %.1f ms = %.3e slice-steps/s" % (ns, ns ** 6, mode, ms, B * T / ms * 1e3), flush=True)
        b.close(); m.close()
