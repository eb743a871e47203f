"""clocks one warp spends per DMMA.8x8x4: `chains` independent accumulators round-robin, one warp
per scheduler (4 warps per block) or more; operand modes: 0 = every DMMA reads the same A and B
registers, 1 = A changes every `chains` DMMAs and B with every DMMA (a k-step-major sweep),
2 = A and B change with every DMMA"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nip_b200.api as api
L = api.load_library()
L.nipgpu_probe_dmma_chain.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
for warps in (4, 8):
    for mode in (0, 1, 2):
        row = []
        for ch in ((1, 2, 4, 8, 16) if mode == 0 else (2, 4, 8)):
            v = C.c_double()
            L.nipgpu_probe_dmma_chain(ch + 100 * mode, warps, 148, C.byref(v))
            row.append("%d ch: %6.2f" % (ch, v.value))
        print("warps/block %2d operand mode %d | " % (warps, mode) + " | ".join(row))
