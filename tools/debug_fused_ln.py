"""Cluster-fused Linear+LayerNorm+SiLU vs torch (run on the GPU box): errors + timing against the unfused pair."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from genrl_b200 import ops
from genrl_b200._lib import lib
DEV = "cuda:0"; bf = torch.bfloat16
P = lambda t: None if t is None else t.data_ptr()
torch.manual_seed(0)
for (M, K, N) in [(200, 264, 128), (333, 1034, 512), (2400, 1024, 1024), (2400, 2048, 1024), (38400, 1024, 1024)]:
    x = torch.randn(M, K, device=DEV).to(bf); w = (torch.randn(N, K, device=DEV) / K ** 0.5).to(bf)
    b = torch.randn(N, device=DEV) * 0.1; g = 1 + 0.1 * torch.randn(N, device=DEV); be = 0.1 * torch.randn(N, device=DEV)
    pre_ref = x.float() @ w.float().t() + b
    y_ref = F.silu(F.layer_norm(pre_ref, (N,), g, be, 1e-5))
    with ops.use_dtype(bf):
        xa = ops.cast_to(x, bf); wa = ops.cast_to(w, bf)
        pre = ops.empty_act(M, N, DEV); y = ops.empty_act(M, N, DEV); st = torch.empty(M, 2, device=DEV)
        fused = lambda: lib.genrl_gemm_ln_fwd(ops._st(), M, N, K, P(xa), xa.stride(0), P(wa), wa.stride(0), P(b), P(g), P(be), 1e-5, 1, P(pre), pre.stride(0),
                                              P(y), y.stride(0), P(st))
        fused(); torch.cuda.synchronize()
        mean_ref = pre_ref.mean(-1); rstd_ref = 1.0 / torch.sqrt(pre_ref.var(-1, unbiased=False) + 1e-5)
        e = lambda a, r: float((a.float() - r).abs().max() / (r.abs().max() + 1e-12))
        print(f"M={M} K={K} N={N}: pre {e(pre, pre_ref):.2e}  y {e(y, y_ref):.2e}  mean {e(st[:, 0], mean_ref):.2e}  rstd {e(st[:, 1], rstd_ref):.2e}", flush=True)
        pre2 = ops.empty_act(M, N, DEV); y2 = ops.empty_act(M, N, DEV)
        def unfused():
            ops.gemm(xa, wa, out=pre2, bias=b)
            lib.genrl_ln_act_fwd(ops._st(), P(pre2), pre2.stride(0), P(g), P(be), 1e-5, 1, P(y2), y2.stride(0), P(st), M, N, 1)
        for name, fn in (("fused", fused), ("unfused", unfused)):
            gr = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fn()
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(gr):
                for _ in range(20): fn()
            gr.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): gr.replay()
            e1.record(); torch.cuda.synchronize()
            print(f"    {name:8s} {e0.elapsed_time(e1) / 100 * 1e3:8.2f} us per layer (20 in a graph)", flush=True)
