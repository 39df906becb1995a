"""One forward + backward of the weight-streaming scan at the configs[4] shard shape (D = 4096, 16 sequences; T from argv, default 16) for ncu."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genrl_b200 import ops

DEV = "cuda:0"
B, T, D = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 16, 4096
with ops.use_dtype(torch.bfloat16):
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(3 * D, 2 * D, device=DEV) / (2 * D) ** 0.5)
    g = torch.nn.Parameter(torch.ones(3 * D, device=DEV)); b = torch.nn.Parameter(torch.zeros(3 * D, device=DEV))
    gx = torch.randn(B * T, 3 * D, device=DEV).requires_grad_(True)
    mask = torch.ones(B * T, device=DEV); h0 = torch.zeros(B, D, device=DEV)
    dd = torch.randn(B * T, D, device=DEV)
    for _ in range(2):
        w.grad = g.grad = b.grad = gx.grad = None
        ops.gru_scan(gx, w, g, b, mask, h0, B, T, D).backward(dd)
    torch.cuda.synchronize()
