"""Per-kernel / per-shape CUDA-event profile of one full update at the benchmark configuration (run on the GPU box)."""
import sys, os, io, contextlib, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from genrl_b200 import ops
from genrl_b200._lib import lib
from genrl_b200.agent.genrl import GenRLAgent

class A: pass
args = A(); args.batch, args.length, args.horizon, args.action_dim, args.sequence, args.preset = 50, 48, 15, 10, "S2", "c2"
precision = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg, obs_space, act_spec = bench.make_cfg(args, "cuda:0", precision)
torch.manual_seed(1)
with contextlib.redirect_stdout(io.StringIO()):
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
data = {k: v.cuda() for k, v in bench.synthetic(50, 48, 10, 0).items()}
def step(i):
    state, outputs, metrics = agent.update_wm(data, i)
    agent.update_acting_behavior(state, outputs, metrics, data)
for i in range(3): step(i)
torch.cuda.synchronize()
import time
t0 = time.perf_counter(); step(3); t_cpu = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print(f"host-side enqueue time {t_cpu*1e3:.1f} ms, until GPU idle {t_all*1e3:.1f} ms")
lib.profile = []
step(4)
torch.cuda.synchronize()
prof, lib.profile = lib.profile, None
agg = collections.OrderedDict()
for name, e0, e1, a3, allargs in prof:
    if name == "genrl_gemm":
        M, N, K = allargs[1:4]; key = (name, M, N, K, f"tA{allargs[6]} tW{allargs[9]} in{allargs[10]} c{allargs[13]} acc{allargs[15]}")
    elif name in ("genrl_ln_act_fwd", "genrl_ln_act_bwd"):
        key = (name, allargs[-3], allargs[-2])
    elif name in ("genrl_im2col", "genrl_col2im"):
        key = (name,) + tuple(allargs[-7:])
    elif name == "genrl_copy2d":
        key = (name, allargs[-3], allargs[-2])
    else:
        key = (name,)
    d = agg.setdefault(key, [0.0, 0])
    d[0] += e0.elapsed_time(e1); d[1] += 1
tot = sum(v[0] for v in agg.values())
print(f"sum of per-call event times: {tot:.2f} ms over {len(prof)} calls")
for key, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    extra = ""
    if key[0] == "genrl_gemm":
        fl = 2.0 * key[1] * key[2] * key[3] * n
        extra = f"  {fl / (ms * 1e-3) / 1e12:7.1f} TFLOP/s"
    print(f"{ms:8.3f} ms  {n:4d} calls  {ms / n * 1e3:8.1f} us/call  {key}{extra}")
