"""Does programmatic dependent launch help a chain of short kernels?  200 GEMMs 4096x1280x1280 (21 us each) back to back,
eager stream and CUDA-graph replay; run once with TAIR_PDL=1 and once with TAIR_PDL=0 (the flag is read at first launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
res = {}
for (M, N, K) in ((4096, 1280, 1280), (65536, 320, 320), (1024, 1280, 1280)):
    a = torch.randn((M, K), device=dev, generator=g).bfloat16(); w = (torch.randn((N, K), device=dev, generator=g) / K ** 0.5).bfloat16()
    bufs = [torch.empty((M, N), device=dev, dtype=torch.bfloat16) for _ in range(2)]
    def chain(n=100):
        for i in range(n):
            ops.gemm(a, w, out=bufs[i & 1])
    chain(5); torch.cuda.synchronize()
    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t_eager = timeit(chain)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        chain()
    t_graph = timeit(gr.replay)
    res[f"{M}x{N}x{K}"] = dict(eager_us_per_gemm=round(10 * t_eager, 2), graph_us_per_gemm=round(10 * t_graph, 2))
print("TAIR_PDL=" + os.environ.get("TAIR_PDL", "1"), res)
