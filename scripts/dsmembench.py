"""Cluster exchange micro-benchmark (debug): cycles to broadcast an operand-tile slice to the peers."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from big_dreamer_b200 import _lib
lib = _lib.load_test()      # libbd_b200_test.so
lib.bd_tc_dsmembench.restype = C.c_int
lib.bd_tc_dsmembench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for R in (2, 4):
    for nbytes in (2048, 8192, 16384, 26624):
        for mode in (0, 1):
            rc = lib.bd_tc_dsmembench(R, nbytes, mode, 20, out.data_ptr(), s)
            torch.cuda.synchronize()
            cyc = int(out[0].item())
            print(f"R={R} bytes/peer={nbytes:6d} mode={'st.cluster' if mode == 0 else 'bulk copy '}: {cyc:6d} cycles  "
                  f"-> {nbytes * (R - 1) / max(cyc, 1):6.1f} B/clk out per CTA", flush=True)
