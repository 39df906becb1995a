"""Error report of the implicit-GEMM conv passes vs torch fp32 (bf16-rounded operands), every pass separately (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from genrl_b200 import ops

DEV = "cuda:0"
bf = torch.bfloat16


def err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-12)


def conv(N, H, Cin, Cout, k):
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(Cin, Cout, k, stride=2)
    x = torch.randn(N, Cin, H, H).to(bf).float()
    wr = conv.weight.detach().to(bf).float().requires_grad_(True); br = conv.bias.detach().clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y = F.conv2d(xr, wr, br, stride=2)
    dy = torch.randn(*y.shape).to(bf).float()
    y.backward(dy)
    cd = conv.to(DEV); cd.weight.grad = None
    with ops.use_dtype(bf):
        xd = x.permute(0, 2, 3, 1).reshape(N * H * H, Cin).to(DEV).to(bf).requires_grad_(True)
        yd, Ho, Wo = ops.conv2d_s2(xd, cd.weight, cd.bias, N, H, H, k)
        torch.cuda.synchronize()
        e_y = err(yd.reshape(N, Ho, Wo, Cout).permute(0, 3, 1, 2), y)
        yd.backward(dy.permute(0, 2, 3, 1).reshape(N * Ho * Wo, Cout).to(DEV).to(bf))
        torch.cuda.synchronize()
    print(f"conv  N={N} H={H} {Cin}->{Cout} k{k}: y {e_y:.2e}  dW {err(cd.weight.grad, wr.grad):.2e}  db {err(cd.bias.grad, br.grad):.2e}  "
          f"dx {err(xd.grad.reshape(N, H, H, Cin).permute(0, 3, 1, 2), xr.grad):.2e}", flush=True)


def convT(N, Hi, Cin, Cout, k, last):
    torch.manual_seed(0)
    conv = torch.nn.ConvTranspose2d(Cin, Cout, k, stride=2)
    x = torch.randn(N, Cin, Hi, Hi).to(bf).float()
    wr = conv.weight.detach().to(bf).float().requires_grad_(True); br = conv.bias.detach().clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y = F.conv_transpose2d(xr, wr, br, stride=2)
    dy = torch.randn(*y.shape).to(bf).float()
    y.backward(dy)
    cd = conv.to(DEV)
    with ops.use_dtype(bf):
        xd = x.permute(0, 2, 3, 1).reshape(N * Hi * Hi, Cin).to(DEV).to(bf).requires_grad_(True)
        yd, Ho, Wo = ops.conv_transpose2d_s2(xd, cd.weight, cd.bias, N, Hi, Hi, k, nchw_out=last)
        torch.cuda.synchronize()
        if last:
            e_y = err(yd, y); yd.backward(dy.to(DEV))
        else:
            e_y = err(yd.reshape(N, Ho, Wo, Cout).permute(0, 3, 1, 2), y)
            yd.backward(dy.permute(0, 2, 3, 1).reshape(N * Ho * Wo, Cout).to(DEV).to(bf))
        torch.cuda.synchronize()
    print(f"convT N={N} Hi={Hi} {Cin}->{Cout} k{k} last={last}: y {e_y:.2e}  dW {err(cd.weight.grad, wr.grad):.2e}  db {err(cd.bias.grad, br.grad):.2e}  "
          f"dx {err(xd.grad.reshape(N, Hi, Hi, Cin).permute(0, 3, 1, 2), xr.grad):.2e}", flush=True)


if __name__ == "__main__":
    for a in [(2, 31, 8, 16, 4), (2, 14, 16, 24, 4), (3, 31, 48, 96, 4), (2, 14, 96, 192, 4), (7, 6, 192, 384, 4)]:
        conv(*a)
    for a in [(2, 5, 16, 8, 5, False), (2, 13, 8, 8, 6, False), (2, 30, 8, 3, 6, True), (5, 5, 192, 96, 5, False), (3, 13, 96, 48, 6, False),
              (2, 30, 48, 3, 6, True)]:
        convT(*a)
