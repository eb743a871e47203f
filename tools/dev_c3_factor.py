"""C3 (factorial DBN, 4 ring-coupled chains x 16 states) on engine 3 (factor by factor) against
engine 1 (grid team streaming the materialised 16^6 cliques): timing and agreement.
N=32 T=8 python tools/dev_c3_factor.py        (GRID=1 also runs engine 1 on min(N, 8) series)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nip_b200.api as api
from nip_b200.synth import FactorialSpec

NS, NY = int(os.environ.get("NS", 16)), int(os.environ.get("NY", 4))
N, T = int(os.environ.get("N", 32)), int(os.environ.get("T", 8))
sp = FactorialSpec(NS, NY, seed=1)
fm = sp.flat()
data = sp.sample(N, T, seed=2)
q = [4, 5, 6, 7]
t0 = time.perf_counter()
m = api.Model(fm, engine=api.ENGINE_FACTOR)
print("engine 3 model on device in %.2fs (factor extraction + verification on the host included)" % (time.perf_counter() - t0), flush=True)
b = m.batch(sp.obs_vars, data)
res = {}
for rep in range(3):
    k0 = api.launch_count()
    post, ll = b.infer(q)
    ms = m.last_kernel_ms()[0]
    print("[factor] smoothing %d x %d: %.2f ms -> %.3e slice-steps/s (%d launches)" % (N, T, ms, N * T / (ms * 1e-3), api.launch_count() - k0), flush=True)
for rep in range(3):
    k0 = api.launch_count()
    counts, L, st = b.estep()
    ms = m.last_kernel_ms()[0]
    print("[factor] E-step   %d x %d: %.2f ms -> %.3e slice-steps/s (%d launches), ll %.6f status %d" % (N, T, ms, N * T / (ms * 1e-3), api.launch_count() - k0, L, st), flush=True)
b.close(); m.close()
if os.environ.get("GRID"):
    n1 = min(N, 8)
    m1 = api.Model(fm, engine=api.ENGINE_JTREE)
    b1 = m1.batch(sp.obs_vars, data[:n1])
    post1, ll1 = b1.infer(q)
    print("[grid]   smoothing %d x %d: %.2f ms" % (n1, T, m1.last_kernel_ms()[0]))
    c1, L1, st1 = b1.estep()
    print("[grid]   E-step   %d x %d: %.2f ms" % (n1, T, m1.last_kernel_ms()[0]))
    rel = lambda x, y: float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1e-300)))
    print("factor vs grid: posteriors %.2e  ll %.2e" % (rel(post[:n1 * T], post1), rel(ll[:n1], ll1)))
    m2 = api.Model(fm, engine=api.ENGINE_FACTOR)
    b2 = m2.batch(sp.obs_vars, data[:n1])
    c2, L2, st2 = b2.estep()
    print("factor vs grid: counts %.2e  L %.2e" % (rel(c2, c1), rel(np.array(L2), np.array(L1))))
