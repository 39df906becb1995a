"""Times the weight-streaming RSSM scan (csrc/rssm_scan_stream.cu) at the BASELINE configs[4] per-GPU shard (D = 4096, 16 sequences x 64 steps)
against the op-level recurrence it replaces: us per time step and GB/s of streamed weight bytes (3 D^2 bf16 per step)."""
import os
import sys
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genrl_b200 import ops

DEV = "cuda:0"


def run(B, T, D, Hd, reps=5):
    with ops.use_dtype(torch.bfloat16):
        torch.manual_seed(0)
        w = torch.nn.Parameter(torch.randn(3 * D, Hd + D, device=DEV) / (Hd + D) ** 0.5)
        g = torch.nn.Parameter(torch.ones(3 * D, device=DEV)); b = torch.nn.Parameter(torch.zeros(3 * D, device=DEV))
        gx = torch.randn(B * T, 3 * D, device=DEV)
        mask = torch.ones(B * T, device=DEV)
        h0 = torch.zeros(B, D, device=DEV)
        out = {}
        for name, env in (("stream", "1"), ("op-level", "0")):
            os.environ["GENRL_SCAN_STREAM"] = env
            with torch.no_grad():
                for _ in range(2):
                    ops.gru_scan(gx, w, g, b, mask, h0, B, T, Hd)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    ops.gru_scan(gx, w, g, b, mask, h0, B, T, Hd)
                e1.record()
                torch.cuda.synchronize()
            out[name] = e0.elapsed_time(e1) * 1e3 / reps
            # forward + backward (dgx, dW_h, dgamma / dbeta, dh0)
            gxg = gx.clone().requires_grad_(True)
            dd = torch.randn(B * T, D, device=DEV)
            for k in range(reps + 1):
                if k == 1:
                    torch.cuda.synchronize()
                    e0.record()
                w.grad = g.grad = b.grad = gxg.grad = None
                ops.gru_scan(gxg, w, g, b, mask, h0, B, T, Hd).backward(dd)
            e1.record()
            torch.cuda.synchronize()
            out[name + " fwd+bwd"] = e0.elapsed_time(e1) * 1e3 / reps
        wbytes = 3 * D * D * 2
        for name, us in out.items():
            print(f"B={B} T={T} D={D}  {name:9s} forward {us:9.1f} us  = {us / T:7.2f} us/step  weights {wbytes / 1e6:.0f} MB/step -> {wbytes / (us / T) / 1e3:7.0f} GB/s")
        return out


if __name__ == "__main__":
    run(16, 64, 4096, 4096)
    run(16, 64, 2048, 2048)
