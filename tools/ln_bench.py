"""Times the LayerNorm(+SiLU) kernels on the step's large shapes: CUDA graph of 20 launches on rotating buffers, us per launch and
GB/s of algorithmic bytes (forward: read x, write y; backward: read dy, x, write dx).  GENRL_LN_PIPE=0 selects the register kernels."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genrl_b200 import ops
from genrl_b200._lib import lib
from genrl_b200.ops import _st, _p, _ld, _CODE

DEV = "cuda:0"
SHAPES = [(38400, 1024), (36000, 1024), (19200, 1024), (2457600, 48), (614400, 96), (153600, 192), (38400, 384), (9600, 1024)]


def timeit(fn, nbuf, reps=20):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i % nbuf)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * reps)


def main():
    dt = torch.bfloat16
    print(f"GENRL_LN_PIPE={os.environ.get('GENRL_LN_PIPE', '1')} GENRL_LN_PIPE_FWD={os.environ.get('GENRL_LN_PIPE_FWD', '1')}")
    print(f"{'M':>9} {'D':>5} | {'fwd us':>8} {'GB/s':>7} | {'bwd us':>8} {'GB/s':>7} | {'bwd+colg us':>11} {'GB/s':>7}")
    for M, D in SHAPES:
        nbuf = max(2, int(400e6 // (M * D * 2 * 4)) + 1)
        nbuf = min(nbuf, 4)
        x = [torch.randn(M, D, device=DEV).to(dt) for _ in range(nbuf)]
        dy = [torch.randn(M, D, device=DEV).to(dt) for _ in range(nbuf)]
        y = [torch.empty(M, D, device=DEV, dtype=dt) for _ in range(nbuf)]
        st = torch.empty(M, 2, device=DEV)
        g = torch.ones(D, device=DEV); b = torch.zeros(D, device=DEV)
        dg = torch.zeros(D, device=DEV); db = torch.zeros(D, device=DEV); dlb = torch.zeros(D, device=DEV)
        eps = 1e-5 if D >= 512 else 1e-3

        def fwd(i):
            lib.genrl_ln_act_fwd(_st(), _p(x[i]), D, _p(g), _p(b), eps, 1, _p(y[i]), D, _p(st), M, D, _CODE[dt])

        def bwd(i):
            lib.genrl_ln_act_bwd(_st(), _p(dy[i]), D, _p(x[i]), D, _p(g), _p(b), _p(st), 1, _p(y[i]), D, None, None, None, M, D, _CODE[dt])

        def bwdc(i):
            lib.genrl_ln_act_bwd(_st(), _p(dy[i]), D, _p(x[i]), D, _p(g), _p(b), _p(st), 1, _p(y[i]), D, _p(dg), _p(db), _p(dlb), M, D, _CODE[dt])

        tf, tb, tc = timeit(fwd, nbuf), timeit(bwd, nbuf), timeit(bwdc, nbuf)
        bf, bb = 2 * M * D * 2, 3 * M * D * 2
        print(f"{M:9d} {D:5d} | {tf:8.1f} {bf / tf / 1e3:7.0f} | {tb:8.1f} {bb / tb / 1e3:7.0f} | {tc:11.1f} {bb / tc / 1e3:7.0f}")


if __name__ == "__main__":
    main()
