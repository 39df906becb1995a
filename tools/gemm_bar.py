"""Library bar for the rollout-sized GEMMs: torch.matmul (cuBLASLt) vs genrl_gemm, each captured in a CUDA graph of 50 back-to-back launches
on rotating operands (so host launch overhead is excluded and the weights come from L2 as in the step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genrl_b200 import ops

dev = "cuda:0"
shapes = [(2400, 1024, 1024), (2400, 1024, 2048), (2400, 3072, 2048), (2400, 1024, 1040), (2400, 2048, 3072), (38400, 1024, 1024), (38400, 1024, 2048)]
R = 50


def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(R):
            fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * R) * 1e3


print(f"{'M':>6s} {'N':>5s} {'K':>5s} | {'cuBLAS us':>9s} {'TFLOP/s':>8s} | {'genrl us':>9s} {'TFLOP/s':>8s} | {'pair us':>9s} {'TFLOP/s':>8s}")
with ops.use_dtype(torch.bfloat16):
    for M, N, K in shapes:
        xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(4)]
        w = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        t_lib = graph_time(lambda i: torch.matmul(xs[i % 4], w.t(), out=out))
        t_own = graph_time(lambda i: ops.gemm(xs[i % 4], w, out=out))
        t_pair = graph_time(lambda i: ops.gemm_pair(xs[i % 4], w, out=out)) if os.environ.get("NO_PAIR") != "1" else float("nan")
        print(f"{M:6d} {N:5d} {K:5d} | {t_lib:9.2f} {fl / t_lib / 1e6:8.1f} | {t_own:9.2f} {fl / t_own / 1e6:8.1f} | {t_pair:9.2f} {fl / t_pair / 1e6:8.1f}", flush=True)
