"""clocks per sweep (128 DMMAs, floor 2048) of the B-fragment delivery variants of sweep.cuh"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nip_b200.api as api
L = api.load_library()
for v in (0, 1, 2, 3, 10, 11, 13):
    out = C.c_double()
    if L.nipgpu_probe_sweep(v, 148, C.byref(out)) != 0:
        continue
    w = 2 if v >= 10 else 1
    print("variant %d, %d warp(s) per scheduler: %.0f clocks per sweep per warp = %.2f per DMMA per scheduler (%.0f%% of the pipe)"
          % (v % 10, w, out.value, out.value / 128 / w, 204800 * w / out.value))
