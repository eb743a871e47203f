"""one smoothing pass of the ns=4 factorial model through the generic engine (for ncu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nip_b200.api as api
from nip_b200.synth import FactorialSpec
sp = FactorialSpec(int(os.environ.get("NS", 4)), 3, seed=1)
m = api.Model(sp.flat(), engine=0)
b = m.batch(sp.obs_vars, sp.sample(1184, 4, seed=2))
b.infer_device([4, 5]); b.infer_device([4, 5])
print(m.last_kernel_ms())
