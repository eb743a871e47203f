"""development probe: peaks + C2 timing (not the contract bench)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nip_b200.api as api
from nip_b200.synth import HmmSpec

print("peaks (DMMA TF, DFMA TF, copy GB/s):", api.probe_peaks(0), flush=True)
S, M = int(os.environ.get("S", 64)), 32
B, T = int(os.environ.get("B", 4096)), int(os.environ.get("T", 1000))
h = HmmSpec(S, M, seed=1)
fm = h.flat()
t0 = time.time(); data = h.sample(B, T, seed=2); print("sample %.1fs" % (time.time() - t0), flush=True)
for engine in (api.ENGINE_AUTO,):
    m = api.Model(fm, engine=engine)
    b = m.batch(h.obs_vars, data)
    for rep in range(4):
        t0 = time.time()
        b.infer_device(h.hidden_query)
        dt = time.time() - t0
        ms, n = m.last_kernel_ms()
        print("engine %d rep %d: wall %.1f ms, kernels %.3f ms (%d launches) -> %.3e slice-steps/s" % (m.engine, rep, dt * 1e3, ms, n, B * T / (ms * 1e-3)), flush=True)
    b.infer_device(h.hidden_query, forward_only=True)
    print("forward only kernels %.3f ms" % m.last_kernel_ms()[0])
if os.environ.get("JT"):
    Bj = 256
    m = api.Model(fm, engine=api.ENGINE_JTREE)
    b = m.batch(h.obs_vars, data[:Bj, :100])
    for rep in range(2):
        b.infer_device(h.hidden_query)
        ms, n = m.last_kernel_ms()
        print("jtree rep %d: kernels %.3f ms -> %.3e slice-steps/s" % (rep, ms, Bj * 100 / (ms * 1e-3)), flush=True)
