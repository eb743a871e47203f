import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nip_b200.api as api
from nip_b200.synth import HmmSpec
S, M = int(os.environ.get("S", 1024)), 64
B, T = int(os.environ.get("B", 4096)), int(os.environ.get("T", 100))
t0 = time.perf_counter(); h = HmmSpec(S, M, seed=1); fm = h.flat(); data = h.sample(B, T, seed=2)
print("model+data %.1fs" % (time.perf_counter() - t0), flush=True)
t0 = time.perf_counter(); m = api.Model(fm); print("model create %.2fs engine %d" % (time.perf_counter() - t0, m.engine), flush=True)
b = m.batch(h.obs_vars, data)
for rep in range(3):
    b.infer_device(h.hidden_query)
    ms, n = m.last_kernel_ms()
    print("C4 S=%d B=%d T=%d: kernels %.2f ms -> %.3e slice-steps/s, %.2f TFLOP/s" % (S, B, T, ms, B * T / (ms * 1e-3), 4.0 * S * S * B * T / (ms * 1e-3) / 1e12), flush=True)
rng = np.random.default_rng(0)
m.mstep(rng.random(m.counts_size()) + 0.1)
for it in range(3):
    t0 = time.perf_counter()
    c, ll, st = b.estep(want_counts=False)
    m.mstep(None)
    dt = time.perf_counter() - t0
    ms, n = m.last_kernel_ms()
    print("C4 EM iter %d: wall %.1f ms, E-step kernels %.1f ms (%d launches timed) -> %.3e slice-steps/s, ll/step %.6f st %d" % (it, dt * 1e3, ms, n, B * T / (ms * 1e-3), ll / (B * T), st), flush=True)
