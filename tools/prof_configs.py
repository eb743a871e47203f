"""minimal runs of the side configurations for ncu captures:  python tools/prof_configs.py C1|C4|C5"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nip_b200.api as api
import bench
name = sys.argv[1]
cfg = bench.make_config(name)
if name == "C1":
    cfg.n_series, cfg.T = 1000000, 50        # the size that fills the machine
if name == "C4":
    cfg.n_series, cfg.T = 4096, 12           # a few slices of the real shape are enough for a capture
fm = cfg.build(2)
m = api.Model(fm)
b = m.batch(cfg.obs_vars, cfg.data)
for _ in range(2):
    print(name, cfg.device_step(m, b), "ms")
