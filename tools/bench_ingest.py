"""data ingest.  read_timeseries() (src/nip.c:512-667) tokenises the text file twice
with a calloc per token; nip_gpu_read_timeseries_bin() reads packed int32 rows.  Rows per second of
both on the same set (CPU only), then the loaded set goes to the device in one batch.
  python tools/bench_ingest.py [n_series] [length]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from nip_b200.synth import HmmSpec
from oracle.bindings import REF_SO, RefLib

n, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 40000), (int(sys.argv[2]) if len(sys.argv) > 2 else 50)
tmp = "/tmp/nip_ingest"
os.makedirs(tmp, exist_ok=True)
h = HmmSpec(4, 5, seed=3)
open(tmp + "/m.net", "w").write(h.net_text())
data = h.sample(n, T, seed=5)
with open(tmp + "/d.txt", "w") as f:
    f.write("M1\n")
    for s in data:
        f.write("\n".join(str(int(x)) for x in s[:, 0]) + "\n\n")
C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
ref = RefLib()
be = C.CDLL(os.path.join(ROOT, "nip_b200", "libnip_gpu_backend.so"))
vp = C.c_void_p
ref.L.ref_read_timeseries = getattr(ref.L, "read_timeseries")
ref.L.read_timeseries.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
be.nip_gpu_write_timeseries_bin.argtypes = [vp, C.c_int, C.c_char_p]
be.nip_gpu_read_timeseries_bin.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
model = ref.parse(tmp + "/m.net")
out = vp()
t0 = time.perf_counter()
k = ref.L.read_timeseries(model.h, (tmp + "/d.txt").encode(), C.byref(out))
t_text = time.perf_counter() - t0
assert k == n
assert be.nip_gpu_write_timeseries_bin(out, n, (tmp + "/d.bin").encode()) == 0
out2 = vp()
t0 = time.perf_counter()
k2 = be.nip_gpu_read_timeseries_bin(model.h, (tmp + "/d.bin").encode(), C.byref(out2))
t_bin = time.perf_counter() - t0
assert k2 == n
rows = n * T
print("read_timeseries (text, %d MB): %.2f s = %.3e rows/s" % (os.path.getsize(tmp + "/d.txt") >> 20, t_text, rows / t_text))
print("nip_gpu_read_timeseries_bin (%d MB): %.3f s = %.3e rows/s (%.0fx)" % (os.path.getsize(tmp + "/d.bin") >> 20, t_bin, rows / t_bin, t_text / t_bin))
print("at C5's 50 M rows: text %.0f s, packed %.1f s (linear extrapolation)" % (5e7 / (rows / t_text), 5e7 / (rows / t_bin)))

# straight from the packed file into a device batch (needs a GPU): one bulk read, one upload
import torch
if torch.cuda.is_available():
    be.nip_gpu_smooth_bin.argtypes = [vp, C.c_char_p, vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    ref.L.refh_variable.restype = vp
    ref.L.refh_variable.argtypes = [vp, C.c_int]
    q = (vp * 1)(ref.L.refh_variable(model.h, 1))
    post, ll, lens = vp(), vp(), vp()
    for rep in range(2):
        t0 = time.perf_counter()
        k3 = be.nip_gpu_smooth_bin(model.h, (tmp + "/d.bin").encode(), q, 1, 0, C.byref(post), C.byref(ll), C.byref(lens))
        t_dir = time.perf_counter() - t0
    assert k3 == n
    print("nip_gpu_smooth_bin (file -> batch -> posteriors in host memory): %.3f s = %.3e rows/s ingested AND smoothed" % (t_dir, rows / t_dir))
