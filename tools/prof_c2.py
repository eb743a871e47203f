"""minimal C2 pass for ncu captures"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nip_b200.api as api
from nip_b200.synth import HmmSpec
h = HmmSpec(64, 32, seed=1)
data = h.sample(4096, int(os.environ.get("T", 1000)), seed=2)
m = api.Model(h.flat())
b = m.batch(h.obs_vars, data)
for _ in range(int(os.environ.get("REPS", 3))):
    b.infer_device(h.hidden_query)
    print(m.last_kernel_ms())
