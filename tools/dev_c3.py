"""C3 (factorial DBN, 4 ring-coupled chains): generic engine, grid team streaming the 16^6 cliques.
NS=16 NY=4 N=4 T=4 python tools/dev_c3.py      (CHECK=1 also runs the per-CTA HBM mode and compares)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nip_b200.api as api
from nip_b200.synth import FactorialSpec

NS, NY = int(os.environ.get("NS", 16)), int(os.environ.get("NY", 4))
N, T = int(os.environ.get("N", 4)), int(os.environ.get("T", 4))
t0 = time.perf_counter()
sp = FactorialSpec(NS, NY, seed=1)
fm = sp.flat()
data = sp.sample(N, T, seed=2)
print("model built in %.1fs: tables %.1f MB, |I| = %d" % (time.perf_counter() - t0, fm.clique_tables.nbytes / 1e6, NS ** 4), flush=True)
q = [4, 5]   # X0, X1


def run(mode):
    if mode:
        os.environ["NIPGPU_JT_MODE"] = mode
    else:
        os.environ.pop("NIPGPU_JT_MODE", None)
    t0 = time.perf_counter()
    m = api.Model(fm, engine=1)
    b = m.batch(sp.obs_vars, data)
    print("[%s] model+batch on device in %.2fs" % (mode or "auto", time.perf_counter() - t0), flush=True)
    res = {}
    for rep in range(2):
        post, ll = b.infer(q)
        ms = m.last_kernel_ms()[0]
        print("[%s] smoothing %d x %d: kernels %.2f ms -> %.3e slice-steps/s" % (mode or "auto", N, T, ms, N * T / (ms * 1e-3)), flush=True)
    res["post"], res["ll"] = post, ll
    for rep in range(2):
        counts, L, st = b.estep()
        ms = m.last_kernel_ms()[0]
        print("[%s] E-step: kernels %.2f ms -> %.3e slice-steps/s, ll %.6f status %d" % (mode or "auto", ms, N * T / (ms * 1e-3), L, st), flush=True)
    res["counts"], res["L"] = counts, L
    if os.environ.get("NIPGPU_JT_TRACE") and mode is None:
        m.jt_trace()
        b.infer(q)
        tr = m.jt_trace()
        names = {1: "marg", 2: "marg2", 3: "update", 4: "quot", 5: "sum", 6: "norm", 0: "other"}
        agg = {}
        for k in range(1, len(tr)):
            tag, dt = int(tr[k, 0]), (int(tr[k, 1]) - int(tr[k - 1, 1])) / 1e3
            if dt > 1e5:       # gap between the two kernels
                continue
            key = (names.get(tag >> 16, "?"), tag & 0xffff)
            a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += dt
        tot = sum(v[1] for v in agg.values())
        print("trace: %d barriers, %.1f us total" % (len(tr), tot))
        for key, (cnt, us) in sorted(agg.items(), key=lambda z: -z[1][1])[:40]:
            print("   %-7s proj %3d: %4d x %8.1f us = %9.1f us (%.1f%%)" % (key[0], key[1], cnt, us / cnt, us, 100 * us / tot))
    b.close(); m.close()
    return res


a = run(None)
print("posterior rows sum to", a["post"][:, :NS].sum(1)[:3], "ll", a["ll"][:3])
if os.environ.get("CHECK"):
    h = run("hbm")
    rel = lambda x, y: float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1e-300)))
    print("grid vs hbm: post %.2e ll %.2e counts %.2e L %.2e" % (rel(a["post"], h["post"]), rel(a["ll"], h["ll"]), rel(a["counts"], h["counts"]), rel(np.array(a["L"]), np.array(h["L"]))))
