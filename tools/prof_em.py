"""minimal C2 E-step for ncu captures"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nip_b200.api as api
from nip_b200.synth import HmmSpec
h = HmmSpec(64, 32, seed=1)
data = h.sample(4096, int(os.environ.get("T", 1000)), seed=2)
m = api.Model(h.flat())
b = m.batch(h.obs_vars, data)
m.mstep(np.random.default_rng(0).random(m.counts_size()) + 0.1)
for _ in range(int(os.environ.get("REPS", 2))):
    b.estep(want_counts=False)
    print(m.last_kernel_ms())
