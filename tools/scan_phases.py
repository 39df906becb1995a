"""Per-phase timing of the persistent RSSM scan (block 0 globaltimer stamps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, math
from genrl_b200 import ops
from genrl_b200._lib import lib
B, T, D, Hd = 50, 48, 1024, 1024
dev = "cuda:0"
torch.manual_seed(0)
w = torch.nn.Parameter((torch.randn(3 * D, Hd + D) / math.sqrt(Hd + D)).to(dev))
g = torch.nn.Parameter(torch.ones(3 * D, device=dev)); b = torch.nn.Parameter(torch.zeros(3 * D, device=dev))
gx = torch.randn(B * T, 3 * D, device=dev)
mask = torch.ones(B * T, device=dev)
dbg = torch.zeros(T, 6, dtype=torch.int64, device=dev)
for dt in (torch.bfloat16, torch.float32):
    with ops.use_dtype(dt), torch.no_grad():
        for _ in range(3):
            ops.gru_scan(gx, w, g, b, mask, None, B, T, Hd)
        torch.cuda.synchronize()
        lib.genrl_rssm_scan_set_debug(dbg.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gru_scan(gx, w, g, b, mask, None, B, T, Hd); e1.record(); torch.cuda.synchronize()
        lib.genrl_rssm_scan_set_debug(None)
        d = dbg.cpu().numpy().astype("float64")
        names = ["matvec", "stats-atomics", "barrier1", "read-stats", "gates", "barrier2"]
        per = []
        for t in range(2, T - 1):
            row = [d[t, 1] - d[t, 0], d[t, 2] - d[t, 1], d[t, 3] - d[t, 2], d[t, 4] - d[t, 3], d[t, 5] - d[t, 4], d[t + 1, 0] - d[t, 5]]
            per.append(row)
        import numpy as np
        per = np.array(per)
        print(dt, f"total {e0.elapsed_time(e1)*1e3:.0f} us; per-step ns (median):", {n: int(np.median(per[:, i])) for i, n in enumerate(names)},
              "sum", int(np.median(per.sum(1))))
