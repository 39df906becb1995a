"""Chain kernel (csrc/mlp_chain.cu) vs separate GEMM + LayerNorm launches, per shape, CUDA graphs of back-to-back launches between CUDA events."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genrl_b200 import ops
from genrl_b200._lib import lib

dev = "cuda:0"
shapes = [(2400, 2048, 1024, 4), (2400, 2048, 1024, 1), (2400, 1024, 1024, 1), (2400, 1040, 1024, 1), (2400, 1024, 1024, 3),
          (38400, 2048, 1024, 1), (38400, 1024, 1024, 1), (38400, 2048, 1024, 4), (36000, 1024, 1024, 1)]
R = int(os.environ.get("REPS", "20"))


def timeit(fn):
    """CUDA graph of R back-to-back launches (no host time between them), replayed 3 times"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(R):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * R) * 1e3


with ops.use_dtype(torch.bfloat16):
    print(f"{'M':>6s} {'K0':>5s} {'N':>5s} {'L':>2s} | {'chain us':>9s} {'TFLOP/s':>8s} | {'gemm+ln us':>10s} {'TFLOP/s':>8s}")
    for M, K0, N, L in shapes:
        x = torch.randn(M, K0, device=dev).bfloat16()
        Ws = [(torch.randn(N, K0 if l == 0 else N, device=dev) / math.sqrt(K0 if l == 0 else N)).bfloat16() for l in range(L)]
        bs = [torch.zeros(N, device=dev) for _ in range(L)]
        gs = [torch.ones(N, device=dev) for _ in range(L)]
        pre = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(L)]
        y = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(L)]
        st = [torch.empty(M, 2, device=dev) for _ in range(L)]
        flops = 2.0 * M * N * (K0 + (L - 1) * N)

        def chain():
            ops.mlp_chain_fwd(x, Ws, bs, gs, bs, 1e-5, 1, pre, y, st)

        def sep():
            inp = x
            for l in range(L):
                ops.gemm(inp, Ws[l], out=pre[l], bias=bs[l])
                lib.genrl_ln_act_fwd(ops._st(), pre[l].data_ptr(), N, gs[l].data_ptr(), bs[l].data_ptr(), 1e-5, 1, y[l].data_ptr(), N, st[l].data_ptr(), M, N, 1)
                inp = y[l]
        tc, ts = timeit(chain), timeit(sep)
        print(f"{M:6d} {K0:5d} {N:5d} {L:2d} | {tc:9.1f} {flops / tc / 1e6:8.1f} | {ts:10.1f} {flops / ts / 1e6:8.1f}")
