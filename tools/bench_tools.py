"""The reference's UNCHANGED command-line tools relinked on the GPU backend (oracle/_ref/*_gpu),
timed beside the same tools built from the reference's own code (*_cpu) on the same files:
  niplikelihood (util/niplikelihood.c:111-135: two make_consistent per data record, the
                 launch-bound fine-grained API) with the memo of make_consistent on and off,
  nipinference  (one forward_backward_inference per series; the backend smooths the set in one pass).
Wall time of the whole process, text parsing included (it is the same code in both builds).
  python tools/bench_tools.py [n_series] [length]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from nip_b200.synth import HmmSpec

TOOLS = os.path.join(ROOT, "oracle", "_ref")
n, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 20000), (int(sys.argv[2]) if len(sys.argv) > 2 else 50)
tmp = "/tmp/nip_tools"
os.makedirs(tmp, exist_ok=True)
h = HmmSpec(4, 5, seed=3)                       # the shape of examples/model.net: 4 states, 5 symbols
open(tmp + "/m.net", "w").write(h.net_text())
data = h.sample(n, T, seed=5)
with open(tmp + "/d.txt", "w") as f:
    f.write("M1\n")
    for s in data:
        f.write("\n".join(str(int(x)) for x in s[:, 0]) + "\n\n")


def run(tool, *args, env=None):
    e = dict(os.environ, **(env or {}))
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(TOOLS, tool)] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr.decode()[-500:]
    return dt, r.stdout, r.stderr.decode()


rec = n * T
t_cpu, out_cpu, _ = run("niplikelihood_cpu", tmp + "/m.net", tmp + "/d.txt", "M1")
print("niplikelihood, %d records: reference build %.2f s = %.3e records/s" % (rec, t_cpu, rec / t_cpu))
for memo in ("1", "0"):
    t, out, err = run("niplikelihood_gpu", tmp + "/m.net", tmp + "/d.txt", "M1", env={"NIP_GPU_SLICE_MEMO": memo, "NIP_GPU_STATS": "1"})
    a = np.array([[float(x) for x in l.split()] for l in out_cpu.decode().split("\n")[1:] if len(l.split()) == 3])
    b = np.array([[float(x) for x in l.split()] for l in out.decode().split("\n")[1:] if len(l.split()) == 3])
    same = a.shape == b.shape and np.allclose(a, b, rtol=2e-5, atol=1e-6)
    print("niplikelihood, GPU backend, memo %s: %.2f s = %.3e records/s; output equal to the reference build's to the printed digits: %s"
          % (memo, t, rec / t, same))
t_cpu, _, _ = run("nipinference_cpu", tmp + "/m.net", tmp + "/d.txt", "P1", tmp + "/p_cpu.txt")
t_gpu, _, _ = run("nipinference_gpu", tmp + "/m.net", tmp + "/d.txt", "P1", tmp + "/p_gpu.txt")
print("nipinference, %d slice-steps: reference build %.2f s, GPU backend %.2f s (both include parsing %d text rows and writing %d)"
      % (rec, t_cpu, t_gpu, rec, rec))
