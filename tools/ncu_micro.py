"""A handful of hot kernels at benchmark shapes, each launched a few times (for `ncu --set full -k regex:...`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genrl_b200 import ops
from genrl_b200._lib import lib

dev = "cuda:0"; bf = torch.bfloat16
P = lambda t: None if t is None else t.data_ptr()
R = lambda *s, dt=bf: torch.randn(*s, device=dev).to(dt)
which = set(sys.argv[1].split(",")) if len(sys.argv) > 1 else {"ln", "conv", "gemm"}
reps = 2
with ops.use_dtype(bf):
    if "ln" in which:
        for (M, D) in [(38400, 1024), (2400, 1024), (2306400, 48)]:
            x = R(M, D); g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
            y = torch.empty_like(x); st = torch.empty(M, 2, device=dev); dy = R(M, D); dx = torch.empty_like(x)
            dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev); dl = torch.zeros(D, device=dev)
            for _ in range(reps):
                lib.genrl_ln_act_fwd(ops._st(), P(x), D, P(g), P(b), 1e-5, 1, P(y), D, P(st), M, D, 1)
                lib.genrl_ln_act_bwd(ops._st(), P(dy), D, P(x), D, P(g), P(b), P(st), 1, P(dx), D, None, None, None, M, D, 1)
                lib.genrl_ln_act_bwd(ops._st(), P(dy), D, P(x), D, P(g), P(b), P(st), 1, P(dx), D, P(dg), P(db), P(dl), M, D, 1)
    if "gemm" in which:
        for (M, N, K) in [(2400, 1024, 1024), (38400, 1024, 1024)]:
            a, w = R(M, K), R(N, K); out = ops.empty_act(M, N, dev, bf)
            for _ in range(reps):
                ops.gemm(a, w, out=out)
    if "conv" in which:
        N = 2400
        for (Hi, Cin, Cout, k, last) in [(30, 48, 3, 6, True), (13, 96, 48, 6, False)]:
            conv = torch.nn.ConvTranspose2d(Cin, Cout, k, stride=2).to(dev)
            x = R(N * Hi * Hi, Cin).requires_grad_(True)
            for _ in range(reps):
                y, Ho, Wo = ops.conv_transpose2d_s2(x, conv.weight, conv.bias, N, Hi, Hi, k, nchw_out=last)
                y.backward(torch.randn_like(y))
        conv = torch.nn.Conv2d(48, 96, 4, stride=2).to(dev)
        x = R(N * 31 * 31, 48).requires_grad_(True)
        for _ in range(reps):
            y, Ho, Wo = ops.conv2d_s2(x, conv.weight, conv.bias, N, 31, 31, 4)
            y.backward(torch.randn_like(y))
torch.cuda.synchronize()
