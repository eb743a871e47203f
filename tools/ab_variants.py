"""A/B timing of kernel variants built side by side (nip_b200/variants/libnipgpu_<tag>.so): one
C2 smoothing pass and one E-step per variant in a fresh process each, CUDA-event kernel times."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np
import nip_b200.api as api
from nip_b200.synth import HmmSpec
S, M, B, T = [int(x) for x in os.environ.get("SHAPE", "64,32,4096,1000").split(",")]
h = HmmSpec(S, M, seed=1)
data = h.sample(B, T, seed=2)
m = api.Model(h.flat())
b = m.batch(h.obs_vars, data)
ms, fw = [], []
for _ in range(6):
    b.infer_device(h.hidden_query)
    ms.append(m.last_kernel_ms()[0])
    fw.append(m.last_forward_ms())
fs = []
for _ in range(4):
    b.infer_device(h.hidden_query, forward_only=True)
    fs.append(m.last_kernel_ms()[0])
ns = []
for _ in range(4):
    b.infer_device(h.hidden_query, want_ll=False)
    ns.append(m.last_kernel_ms()[0])
post, ll = b.infer(h.hidden_query)
m.mstep(np.random.default_rng(0).random(m.counts_size()) + 0.1)
es = []
for _ in range(4):
    c, L, st = b.estep()
    es.append(m.last_kernel_ms()[0])
print("%%-8s smooth %%.3f ms = fwd %%.3f + bwd %%.3f (min of %%s)  filter %%.3f  smooth-without-ll %%.3f  estep %%.3f ms   checksum %%.15g %%.15g %%.15g" %% (
    os.environ["TAG"], min(ms[1:]), min(fw[1:]), min(ms[1:]) - min(fw[1:]), " ".join("%%.3f" %% x for x in ms[1:]), min(fs[1:]), min(ns[1:]), min(es[1:]), post.sum(), ll.sum(), c.sum()))
''' % ROOT
vdir = os.path.join(ROOT, "nip_b200", "variants")
tags = sys.argv[1:] or sorted(f[len("libnipgpu_"):-3] for f in os.listdir(vdir) if f.endswith(".so"))
for t in tags:
    if "=" in t:     # an environment switch of the shipped library, e.g. NIPGPU_CHAIN_PAIR=0
        k, v = t.split("=", 1)
        env = dict(os.environ, TAG=t[-8:], **{k: v})
    else:
        env = dict(os.environ, NIPGPU_LIB_PATH=os.path.join(vdir, "libnipgpu_%s.so" % t), TAG=t)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print((r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
