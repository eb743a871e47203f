import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nip_b200.api as api
L = api.load_library()
L.nipgpu_probe_dmma_chain.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
print("clocks per DMMA (one warp's view):  chains x warps/block (148 blocks)")
for warps in (1, 4, 8, 16):
    row = []
    for ch in (1, 2, 4, 8, 16):
        v = C.c_double()
        L.nipgpu_probe_dmma_chain(ch, warps, 148, C.byref(v))
        row.append("%d ch: %6.1f" % (ch, v.value))
    print("warps/block %2d | " % warps + " | ".join(row))
