"""Warm per-call timings of the hot kernels at the benchmark shapes (back-to-back launches, CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genrl_b200 import ops
from genrl_b200._lib import lib

dev = "cuda:0"
bf = torch.bfloat16

def timeit(name, fn, n=50, flops=None, bytes_=None):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    extra = ""
    if flops: extra += f"  {flops / us / 1e6:8.1f} TFLOP/s"
    if bytes_: extra += f"  {bytes_ / us / 1e3:8.1f} GB/s"
    print(f"{us:9.2f} us  {name}{extra}")

def R(*s, dt=bf): return torch.randn(*s, device=dev).to(dt)

with ops.use_dtype(bf):
    for (M, N, K) in [(2400, 1024, 1024), (2400, 3072, 2048), (2400, 1024, 2048), (38400, 1024, 1024), (2400, 255, 1024), (2400, 10, 1024)]:
        a, w = R(M, K), R(N, K)
        out = ops.empty_act(M, N, dev, bf)
        timeit(f"gemm fwd  NT  M={M} N={N} K={K} bf16 out", lambda: ops.gemm(a, w, out=out), flops=2.0 * M * N * K)
        dy = R(M, N)
        dx = ops.empty_act(M, K, dev, bf)
        timeit(f"gemm dgrad (tW) M={M} N={K} K={N}", lambda: ops.gemm(dy, w, out=dx, trans_w=True), flops=2.0 * M * N * K)
        dw = torch.zeros(N, K, device=dev)
        timeit(f"gemm wgrad (tA,tW,acc) M={N} N={K} K={M}", lambda: ops.gemm(dy, a, out=dw, trans_a=True, trans_w=True, accumulate=True), flops=2.0 * M * N * K)
    for (M, D) in [(2400, 1024), (38400, 1024), (2306400, 48)]:
        x = R(M, D); g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
        y = torch.empty_like(x); st = torch.empty(M, 2, device=dev); dy = R(M, D); dx = torch.empty_like(x)
        dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev); dl = torch.zeros(D, device=dev)
        P = lambda t: t.data_ptr()
        timeit(f"ln_act_fwd {M}x{D}", lambda: lib.genrl_ln_act_fwd(ops._st(), P(x), D, P(g), P(b), 1e-5, 1, P(y), D, P(st), M, D, 1), bytes_=M * D * 4)
        timeit(f"ln_act_bwd (frozen) {M}x{D}", lambda: lib.genrl_ln_act_bwd(ops._st(), P(dy), D, P(x), D, P(g), P(b), P(st), 1, P(dx), D, None, None, None, M, D, 1), bytes_=M * D * 6)
        timeit(f"ln_colgrad alone {M}x{D}", lambda: lib.genrl_ln_colgrad(ops._st(), P(dy), D, P(x), D, P(dx), D, P(g), P(b), P(st), 1, P(dg), P(db), P(dl), M, D, 1), bytes_=M * D * 6)
        timeit(f"ln_act_bwd (+colgrad) {M}x{D}", lambda: lib.genrl_ln_act_bwd(ops._st(), P(dy), D, P(x), D, P(g), P(b), P(st), 1, P(dx), D, P(dg), P(db), P(dl), M, D, 1), bytes_=M * D * 12)
    M = 2400
    a, b2 = R(M, 1024), R(M, 1024)
    outc = ops.empty_act(M, 2048, dev, bf)
    timeit("copy2d 2400x1024 bf16->bf16 (strided dst)", lambda: ops.copy2d(a, outc[:, :1024]), bytes_=M * 1024 * 4)
    f32 = torch.randn(M, 1024, device=dev)
    timeit("copy2d 2400x1024 f32->bf16", lambda: ops.copy2d(f32, outc[:, 1024:]), bytes_=M * 1024 * 6)
    z = R(M, 3072); h = torch.randn(M, 1024, device=dev); g3 = torch.ones(3072, device=dev); b3 = torch.zeros(3072, device=dev)
    ho = torch.empty(M, 1024, device=dev); st = torch.empty(M, 2, device=dev); dz = torch.empty_like(z); dhp = torch.empty_like(h)
    P = lambda t: t.data_ptr()
    timeit("gru_fwd 2400x1024", lambda: lib.genrl_gru_fwd(ops._st(), P(z), 3072, P(g3), P(b3), P(h), 1024, P(ho), 1024, None, 0, P(st), M, 1024, 1))
    timeit("gru_bwd 2400x1024 (frozen)", lambda: lib.genrl_gru_bwd(ops._st(), P(h), 1024, P(z), 3072, P(g3), P(b3), P(h), 1024, P(st), P(dz), 3072, P(dhp), 1024, None, None, M, 1024, 1))
    lg = torch.randn(M, 1024, device=dev)
    timeit("cat_sample fwd 2400x32x32 (philox)", lambda: ops.cat_sample(lg, 32, 32, None, 1, 1, True))
    tiny = torch.zeros(8, device=dev)
    timeit("torch tiny elementwise (launch floor)", lambda: tiny.add_(1.0))
    # dependent chain of small kernels: launch latency when each kernel waits for the previous one
    x = R(M, 1024); g = torch.ones(1024, device=dev); b = torch.zeros(1024, device=dev); w = R(1024, 1024)
    y = torch.empty_like(x); st2 = torch.empty(M, 2, device=dev); pre = ops.empty_act(M, 1024, dev, bf)
    def chain():
        ops.gemm(x, w, out=pre)
        lib.genrl_ln_act_fwd(ops._st(), P(pre), 1024, P(g), P(b), 1e-5, 1, P(y), 1024, P(st2), M, 1024, 1)
    timeit("gemm(2400x1024x1024) + ln_act chain", chain)
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): chain()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(gr):
        for _ in range(20): chain()
    timeit("  same chain x20 inside one CUDA graph (per chain)", lambda: gr.replay(), n=20)
