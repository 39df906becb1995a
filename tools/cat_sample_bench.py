import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genrl_b200 import ops
from genrl_b200._lib import lib
from genrl_b200.ops import _st, _p, _CODE
dev = "cuda:0"
N, S, K = 2400, 32, 32
logit = torch.randn(N, S * K, device=dev)
stoch = torch.empty(N, S * K, device=dev, dtype=torch.bfloat16); st32 = torch.empty(N, S * K, device=dev)
ds = torch.randn(N, S * K, device=dev).bfloat16(); dl = torch.empty(N, S * K, device=dev, dtype=torch.bfloat16)
step = torch.zeros(1, device=dev, dtype=torch.int64)
def fwd():
    lib.genrl_cat_sample_fwd(_st(), _p(logit), None, 3, 1, step.data_ptr(), 1, _p(stoch), S * K, _p(st32), S * K, None, N, S, K, _CODE[torch.bfloat16])
def bwd():
    lib.genrl_cat_sample_bwd2(_st(), _p(logit), _p(ds), S * K, _p(ds), S * K, _p(dl), S * K, N, S, K, _CODE[torch.bfloat16])
for name, fn in (("fwd", fwd), ("bwd2", bwd)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) * 1e3 / 100, "us")
