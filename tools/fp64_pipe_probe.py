"""Do DMMA.8x8x4 and scalar FP64 instructions share a pipe on B200?  clocks per DMMA of a warp that
issues J independent DFMAs after every DMMA, one and two warps per scheduler"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nip_b200.api as api
L = api.load_library()
for warps in (4, 8):
    row = []
    for j in (0, 1, 2, 4):
        v = C.c_double()
        L.nipgpu_probe_dmma_dfma(j, warps, 148, C.byref(v))
        row.append("%d DFMA: %6.2f" % (j, v.value))
    print("%d warp(s) per scheduler, clocks per DMMA (one warp's view) | " % (warps // 4) + " | ".join(row))
