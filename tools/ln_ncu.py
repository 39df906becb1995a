"""One launch each of the large-row LayerNorm kernels (forward, frozen backward, backward with parameter gradients) for an ncu capture."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genrl_b200._lib import lib
from genrl_b200.ops import _st, _p, _CODE

DEV = "cuda:0"
M, D = int(sys.argv[1]) if len(sys.argv) > 1 else 38400, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dt = torch.bfloat16
x = torch.randn(M, D, device=DEV).to(dt); dy = torch.randn(M, D, device=DEV).to(dt); y = torch.empty_like(x)
st = torch.empty(M, 2, device=DEV); g = torch.ones(D, device=DEV); b = torch.zeros(D, device=DEV)
dg = torch.zeros(D, device=DEV); db = torch.zeros(D, device=DEV); dlb = torch.zeros(D, device=DEV)
for _ in range(2):
    lib.genrl_ln_act_fwd(_st(), _p(x), D, _p(g), _p(b), 1e-5, 1, _p(y), D, _p(st), M, D, _CODE[dt])
    lib.genrl_ln_act_bwd(_st(), _p(dy), D, _p(x), D, _p(g), _p(b), _p(st), 1, _p(y), D, None, None, None, M, D, _CODE[dt])
    lib.genrl_ln_act_bwd(_st(), _p(dy), D, _p(x), D, _p(g), _p(b), _p(st), 1, _p(y), D, _p(dg), _p(db), _p(dlb), M, D, _CODE[dt])
torch.cuda.synchronize()
