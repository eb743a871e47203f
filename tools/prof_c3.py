"""one E-step of config C3 (factorial DBN 4 x 16 ring-coupled) on engine 3 over a few series, for
ncu captures of k_fac_contract:  N=8 T=2 python tools/prof_c3.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nip_b200.api as api
from nip_b200.synth import FactorialSpec
N, T = int(os.environ.get("N", 8)), int(os.environ.get("T", 2))
sp = FactorialSpec(16, 4, seed=1)
m = api.Model(sp.flat())
assert m.engine == api.ENGINE_FACTOR
b = m.batch(sp.obs_vars, sp.sample(N, T, seed=2))
counts, L, st = b.estep()
print("C3 E-step %d x %d: %.2f ms, ll %.6f, status %d" % (N, T, m.last_kernel_ms()[0], L, st))
