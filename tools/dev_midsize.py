"""mid-size cliques (factorial DBN, 4 ring-coupled chains of ns states: three ns^6-entry cliques):
engine 1 with per-CTA HBM workspaces / the grid team's concurrent groups, against engine 3
(factor by factor).  Smoothing of B series x T slices, two queried variables."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nip_b200.api as api
from nip_b200.synth import FactorialSpec
for ns, B, T in ((4, 2000, 6), (6, 600, 6), (8, 600, 6), (12, 128, 6)):
    sp = FactorialSpec(ns, 3, seed=1)
    fm = sp.flat()
    data = sp.sample(B, T, seed=2)
    ref = None
    for mode in ("hbm", "grid", "factor"):
        if mode != "factor":
            os.environ["NIPGPU_JT_MODE"] = mode
        else:
            os.environ.pop("NIPGPU_JT_MODE", None)
        try:
            m = api.Model(fm, engine=3 if mode == "factor" else 1)
        except Exception as e:
            print("ns=%d %s: %s" % (ns, mode, e))
            continue
        b = m.batch(sp.obs_vars, data)
        for rep in range(2):
            post, ll = b.infer([4, 5])
        ms = m.last_kernel_ms()[0]
        for rep in range(2):
            b.estep(want_counts=False)
        ems = m.last_kernel_ms()[0]
        if ref is None:
            ref = post
        err = float(np.max(np.abs(post - ref)))
        print("ns=%2d (%8d-entry cliques) %-6s: smoothing %8.1f ms = %.3e slice-steps/s, E-step %8.1f ms = %.3e /s, max abs diff %.1e"
              % (ns, ns ** 6, mode, ms, B * T / ms * 1e3, ems, B * T / ems * 1e3, err), flush=True)
        b.close(); m.close()
