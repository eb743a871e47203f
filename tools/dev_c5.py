import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import nip_b200.api as api
from cases import Case
c = Case("model_net")
N, T = int(os.environ.get("N", 1000000)), 50
rng = np.random.default_rng(1)
data = rng.integers(0, 5, size=(N, T, 1), dtype=np.int32)
m = api.Model(c.fm)
t0 = time.perf_counter(); b = m.batch(c.obs_vars, data); print("upload %.2fs" % (time.perf_counter() - t0), flush=True)
on = np.zeros(c.fm.n_vars, dtype=np.uint8); on[2] = 1
for rep in range(2):
    t0 = time.perf_counter()
    out = b.likelihood(1 - on, on)
    print("likelihood: wall %.1f ms, kernel %.2f ms -> %.3e records/s (kernel)" % ((time.perf_counter() - t0) * 1e3, m.last_kernel_ms()[0], N * T / (m.last_kernel_ms()[0] * 1e-3)), flush=True)
print(out[:3])
for rep in range(2):
    b.infer_device([1]); print("smoothing 1M x 50: kernels %.2f ms -> %.3e slice-steps/s" % (m.last_kernel_ms()[0], N * T / (m.last_kernel_ms()[0] * 1e-3)))
