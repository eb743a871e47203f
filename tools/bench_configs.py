"""All five BASELINE.json configurations on one GPU, device-timed (the contract bench is bench.py;
this is the side table of DESIGN.md section 5).   python tools/bench_configs.py > profiles/...txt"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import nip_b200.api as api
from nip_b200.synth import FactorialSpec, HmmSpec

HBM = 6542.1e9   # MEASURED_PEAKS.json
dmma, dfma, copy = api.probe_peaks(0)
print("probe: DMMA %.1f TFLOP/s, DFMA %.1f TFLOP/s, copy %.0f GB/s" % (dmma, dfma, copy))


def best(fn, reps=3):
    out = None
    for _ in range(reps):
        r = fn()
        out = r if out is None or r < out else out
    return out


def smooth_ms(m, b, q):
    def one():
        b.infer_device(q)
        return m.last_kernel_ms()[0]
    return best(one)


def estep_ms(m, b):
    def one():
        b.estep(want_counts=False)
        return m.last_kernel_ms()[0]
    return best(one)


# ---- C1 / C5: examples/model.net-sized model (4 states, 5 symbols), 1 M x 50 ----
from cases import Case
c = Case("model_net")
N, T = 1000000, 50
rng = np.random.default_rng(1)
data = rng.integers(0, 5, size=(N, T, 1), dtype=np.int32)
m = api.Model(c.fm)
b = m.batch(c.obs_vars, data)
ms = smooth_ms(m, b, [1])
print("C1/C5 model, smoothing 1M x 50 (chain engine): %.2f ms, %.3e slice-steps/s, %.0f GB/s = %.2f of HBM peak"
      % (ms, N * T / ms * 1e3, 104.0 * N * T / ms * 1e3 / 1e9, 104.0 * N * T / ms * 1e3 / HBM))
on = np.zeros(c.fm.n_vars, dtype=np.uint8)
on[2] = 1
def lik():
    b.likelihood(1 - on, on)
    return m.last_kernel_ms()[0]
ms = best(lik)
print("C5 likelihood loop 1M x 50 (per-configuration table + gather): %.3f ms, %.3e records/s, %.0f GB/s = %.2f of HBM peak"
      % (ms, N * T / ms * 1e3, 21.0 * N * T / ms * 1e3 / 1e9, 21.0 * N * T / ms * 1e3 / HBM))
b.close(); m.close()

# ---- C2: HMM 64 x 32, 4096 x 1000 ----
h = HmmSpec(64, 32, seed=1)
data = h.sample(4096, 1000, seed=2)
m = api.Model(h.flat())
b = m.batch(h.obs_vars, data)
ms = smooth_ms(m, b, h.hidden_query)
n = 4096 * 1000
print("C2 smoothing 4096 x 1000: %.3f ms, %.3e slice-steps/s, %.1f TFLOP/s = %.2f of DMMA peak"
      % (ms, n / ms * 1e3, 16384.0 * n / ms * 1e3 / 1e12, 16384.0 * n / ms * 1e3 / 1e12 / dmma))
m.mstep(np.random.default_rng(0).random(m.counts_size()) + 0.1)
ms = estep_ms(m, b)
print("C2 E-step 4096 x 1000: %.3f ms, %.3e slice-steps/s" % (ms, n / ms * 1e3))
b.close(); m.close()

# ---- C4: HMM 1024 x 64, 4096 x 100 ----
h = HmmSpec(1024, 64, seed=1)
data = h.sample(4096, 100, seed=2)
m = api.Model(h.flat())
b = m.batch(h.obs_vars, data)
ms = smooth_ms(m, b, h.hidden_query)
n = 4096 * 100
print("C4 smoothing 4096 x 100: %.2f ms, %.3e slice-steps/s, %.1f TFLOP/s = %.2f of DMMA peak"
      % (ms, n / ms * 1e3, 4.0 * 1024 * 1024 * n / ms * 1e3 / 1e12, 4.0 * 1024 * 1024 * n / ms * 1e3 / 1e12 / dmma))
m.mstep(np.random.default_rng(0).random(m.counts_size()) + 0.1)
ms = estep_ms(m, b)
print("C4 E-step 4096 x 100: %.2f ms, %.3e slice-steps/s, %.1f TFLOP/s"
      % (ms, n / ms * 1e3, 6.0 * 1024 * 1024 * n / ms * 1e3 / 1e12))
b.close(); m.close()

# ---- C3: factorial 4 x 16 states, three 16^6-entry cliques ----
sp = FactorialSpec(16, 4, seed=1)
fm = sp.flat()
data = sp.sample(8, 8, seed=2)
m = api.Model(fm, engine=1)
b = m.batch(sp.obs_vars, data)
ms = smooth_ms(m, b, [4, 5], ) if False else best(lambda: (b.infer([4, 5]), m.last_kernel_ms()[0])[1], 2)
n = 64
print("C3 smoothing 8 x 8 (grid team, 403 MB of tables): %.1f ms, %.1f slice-steps/s, %.2f TB/s = %.2f of HBM peak (6.2 GB per slice-step)"
      % (ms, n / ms * 1e3, 6.2e9 * n / ms * 1e3 / 1e12, 6.2e9 * n / ms * 1e3 / HBM))
ms = best(lambda: (b.estep(), m.last_kernel_ms()[0])[1], 2)
print("C3 E-step 8 x 8: %.1f ms, %.1f slice-steps/s" % (ms, n / ms * 1e3))

# ---- C3, uncoupled variant: four 16^5-entry cliques ----
b.close(); m.close()
sp = FactorialSpec(16, 4, seed=1, coupled=False)
data = sp.sample(64, 8, seed=2)
m = api.Model(sp.flat(), engine=1)
b = m.batch(sp.obs_vars, data)
ms = best(lambda: (b.infer([4, 5]), m.last_kernel_ms()[0])[1], 2)
n = 64 * 8
print("C3 uncoupled variant (four 16^5-entry cliques), smoothing 64 x 8: %.1f ms, %.1f slice-steps/s" % (ms, n / ms * 1e3))
ms = best(lambda: (b.estep(), m.last_kernel_ms()[0])[1], 2)
print("C3 uncoupled variant, E-step 64 x 8: %.1f ms, %.1f slice-steps/s" % (ms, n / ms * 1e3))
