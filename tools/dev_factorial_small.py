"""generic engine (CTA / warp teams) on small factorial DBNs: throughput of the universal path"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nip_b200.api as api
from nip_b200.synth import FactorialSpec
for ns, ny, B, T in ((2, 2, 8192, 32), (3, 2, 8192, 32), (4, 3, 4096, 16), (5, 3, 1024, 8)):
    sp = FactorialSpec(ns, ny, seed=1)
    fm = sp.flat()
    data = sp.sample(B, T, seed=2)
    m = api.Model(fm, engine=0)
    b = m.batch(sp.obs_vars, data)
    for rep in range(2):
        b.infer_device([4, 5])
    ms = m.last_kernel_ms()[0]
    for rep in range(2):
        b.estep(want_counts=False)
    ms2 = m.last_kernel_ms()[0]
    print("factorial ns=%d (cliques %d entries), %d x %d: smoothing %.2f ms = %.3e slice-steps/s; E-step %.2f ms = %.3e"
          % (ns, ns ** 6, B, T, ms, B * T / ms * 1e3, ms2, B * T / ms2 * 1e3), flush=True)
    b.close(); m.close()
