"""Sweep BN / split-K of the tcgen05 GEMM for the shapes of the benchmark step; times 20 back-to-back launches in a CUDA graph."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GENRL_GEMM_FORCE"] = "1"
import torch
from genrl_b200 import ops
bf = torch.bfloat16; dev = "cuda:0"
def R(*s): return torch.randn(*s, device=dev).to(bf)
shapes = [  # M N K tA tW acc outdtype
    (2400, 1024, 1024, 0, 0, 0, bf), (2400, 1024, 1024, 0, 1, 0, bf), (1024, 1024, 2400, 1, 1, 1, torch.float32),
    (2400, 3072, 2048, 0, 0, 0, bf), (2400, 2048, 3072, 0, 1, 0, bf), (1024, 2048, 2400, 1, 1, 1, torch.float32),
    (2400, 1024, 2048, 0, 0, 0, bf), (38400, 1024, 1024, 0, 0, 0, bf), (38400, 1024, 1024, 0, 1, 0, bf),
    (1024, 1024, 36000, 1, 1, 1, torch.float32), (2306400, 48, 48, 0, 0, 0, bf), (405600, 1728, 96, 0, 1, 0, bf),
    (48, 48, 2306400, 1, 1, 0, torch.float32), (96, 1728, 405600, 1, 1, 0, torch.float32),
]
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * n) * 1e3
with ops.use_dtype(bf):
    for (M, N, K, ta, tw, acc, odt) in shapes:
        a = R(K, M) if ta else R(M, K)
        w = R(K, N) if tw else R(N, K)
        out = torch.zeros(M, N, device=dev, dtype=odt) if odt == torch.float32 else ops.empty_act(M, N, dev, bf)
        res = []
        for bn in (64, 128, 160, 192, 256):
            if tw and bn % 64: continue
            if bn > 64 and bn - 64 >= N: continue
            for sp in ((1, 2, 4, 8, 16, 32, 64, 128) if odt == torch.float32 else (1,)):
                if sp > max(1, K // 64 // 4): continue
                os.environ["GENRL_GEMM_BN"] = str(bn); os.environ["GENRL_GEMM_SPLITS"] = str(sp)
                try:
                    us = bench(lambda: ops.gemm(a, w, out=out, trans_a=bool(ta), trans_w=bool(tw), accumulate=bool(acc)))
                except Exception as e:
                    us = float("nan")
                res.append((us, bn, sp))
        res.sort()
        best = res[0]
        print(f"M={M} N={N} K={K} tA={ta} tW={tw} acc={acc}: best {best[0]:.1f}us bn={best[1]} s={best[2]} ({2.0*M*N*K/best[0]/1e6:.0f} TFLOP/s) | " +
              " ".join(f"bn{b}/s{s}:{u:.1f}" for u, b, s in res[:8]))
