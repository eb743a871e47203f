import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nip_b200.api as api
from nip_b200.synth import HmmSpec
h = HmmSpec(64, 32, seed=1); fm = h.flat()
B, T = int(os.environ.get("B", 4096)), int(os.environ.get("T", 1000))
data = h.sample(B, T, seed=2)
m = api.Model(fm); b = m.batch(h.obs_vars, data)
rng = np.random.default_rng(0)
m.mstep(rng.random(m.counts_size()) + 0.1)
for it in range(5):
    t0 = time.perf_counter()
    c, ll, st = b.estep(want_counts=False)
    m.mstep(None)
    dt = time.perf_counter() - t0
    print("EM iter %d: wall %.2f ms, kernels %.3f ms, ll/step %.6f st %d" % (it, dt * 1e3, m.last_kernel_ms()[0], ll / (B * T), st), flush=True)
