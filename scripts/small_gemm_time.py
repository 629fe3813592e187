"""GPU-side time of the fp32 small-batch GEMMs without host launch overhead (CUDA graph replay)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
bd.set_precision("fp32")
rows = int(os.environ.get("ROWS", 50))
dm = bd.DenseModel(230, 200, 1, "ELU").cuda()
x = torch.randn(rows, 230, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = dm(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                y = dm(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        g.replay()
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"rows={rows}: {e0.elapsed_time(e1) * 1e3 / (10 * 20 * 5):.2f} us per GEMM launch (graph replay, 5 GEMMs per forward)")
