"""MMA chain of one engine layer in isolation, with / without concurrent TMA and store traffic (debug)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from big_dreamer_b200 import _lib
lib = _lib.load_test()      # libbd_b200_test.so
lib.bd_tc_mmabench2.restype = C.c_int
lib.bd_tc_mmabench2.argtypes = [C.c_int] * 5 + [C.c_void_p] * 3
out = torch.zeros(4, dtype=torch.int64, device="cuda")
src = torch.zeros(64 * 13312, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for N in (208, 64):
    for tma, stw in ((0, 0), (0, 256 + 2), (0, 256 + 2 + 512), (1, 256 + 2 + 512)):
        res = []
        for nrep in (4, 36):
            out.zero_()
            rc = lib.bd_tc_mmabench2(N, 13, nrep, tma, stw, src.data_ptr(), out.data_ptr(), s)
            torch.cuda.synchronize()
            res.append(out.tolist())
        per = (res[1][0] - res[0][0]) / (32 * 13)
        print(f"N={N:3d} tma={tma} stores={stw}: {per:6.1f} cyc/MMA (tensor floor {N/2:.0f})  tma copies {res[1][2]} stores/thread {res[1][3]}", flush=True)
