"""coupled2x3 golden model on a large random batch: which engine serves interface queries?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import nip_b200.api as api
from cases import Case
c = Case("coupled2x3")
rng = np.random.default_rng(0)
B, T = 4096, 200
cards = [int(c.fm.var_card[v]) for v in c.obs_vars]
data = np.stack([rng.integers(0, k, size=(B, T)) for k in cards], axis=2).astype(np.int32)
m = api.Model(c.fm, engine=0)
b = m.batch(c.obs_vars, data)
iface = [int(v) for v in c.fm.outgoing]
for q in (iface, [iface[0]], c.query):
    for rep in range(2):
        b.infer_device(q)
    ms = m.last_kernel_ms()[0]
    print("query %r: kernels %.3f ms -> %.3e slice-steps/s" % (q, ms, B * T / ms * 1e3))
