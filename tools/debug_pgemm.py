"""Persistent 256x128 GEMM vs the 128x128 kernel: error and time at the large-M shapes (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genrl_b200 import ops
DEV = "cuda:0"; bf = torch.bfloat16
torch.manual_seed(0)
for (M, N, K, tw) in [(700, 255, 136, False), (38400, 1024, 1024, False), (38400, 1024, 1024, True), (38400, 1024, 2048, False), (38400, 2048, 1024, True),
                      (36000, 1024, 1024, False), (38400, 255, 1024, False)]:
    a = torch.randn(M, K, device=DEV).to(bf); w = torch.randn(N, K, device=DEV).to(bf)
    wd = ops.cast_to(w.t().contiguous() if tw else w, bf); ad = ops.cast_to(a, bf)
    ref = a.float() @ w.float().t()
    res = {}
    for mode in ("0", "2"):
        os.environ["GENRL_GEMM_PERSIST"] = mode
        out = ops.empty_act(M, N, DEV, bf)
        fn = lambda: ops.gemm(ad, wd, out=out, trans_w=tw)
        fn(); torch.cuda.synchronize()
        err = float((out.float() - ref).abs().max() / ref.abs().max())
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        res[mode] = (err, us)
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K} tW={int(tw)}: 128x128 err {res['0'][0]:.1e} {res['0'][1]:7.1f} us {fl/res['0'][1]/1e6:6.0f} TF | persistent 256x128 err {res['2'][0]:.1e} "
          f"{res['2'][1]:7.1f} us {fl/res['2'][1]/1e6:6.0f} TF", flush=True)
