"""One full update at the benchmark configuration between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import sys, os, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
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
for i in range(2): step(i)
torch.cuda.synchronize()
from genrl_b200._lib import lib
import json
_orig = lib.genrl_gemm
shapes = []
def _logged(*a):
    shapes.append([a[1], a[2], a[3], a[6], a[9], a[13], a[15]])     # M N K transA transW c_dtype accumulate
    return _orig(*a)
lib.__dict__["genrl_gemm"] = _logged
torch.cuda.profiler.start()
step(2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

os.makedirs("gpurun_out", exist_ok=True)
json.dump(shapes, open("gpurun_out/gemm_shapes.json", "w"))
