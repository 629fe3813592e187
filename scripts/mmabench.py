import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from big_dreamer_b200 import _lib
lib = _lib.load_test()      # libbd_b200_test.so
lib.bd_tc_mmabench.restype = C.c_int
lib.bd_tc_mmabench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
out = torch.zeros(1, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for layout in (0, 1):
    for dep in (1, 0):
        for N in (16, 64, 112, 128, 192, 208, 256):
            res = []
            for nmma in (64, 576):
                rc = lib.bd_tc_mmabench(N, nmma, layout, dep, out.data_ptr(), s)
                torch.cuda.synchronize()
                res.append(int(out.item()))
            per = (res[1] - res[0]) / 512
            print(f"layout={'KM8' if layout == 0 else 'SW128'} dep={dep} N={N:3d}: {per:6.1f} cyc/MMA (ideal {N/2:.0f})  bytes/MMA={(128+N)*32}  -> {(128+N)*32/per:5.1f} B/clk", flush=True)
