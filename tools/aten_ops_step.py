"""Which torch (aten) ops still launch kernels inside one step, by op and input shape (run on the GPU box)."""
import sys, os, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from genrl_b200.agent.genrl import GenRLAgent
from torch.profiler import profile, ProfilerActivity

class A: pass
args = A(); args.batch, args.length, args.horizon, args.action_dim, args.sequence, args.preset = 50, 48, 15, 10, "S2", "c2"
cfg, obs_space, act_spec = bench.make_cfg(args, "cuda:0", 16)
torch.manual_seed(1)
with contextlib.redirect_stdout(io.StringIO()):
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
data = {k: v.cuda() for k, v in bench.synthetic(50, 48, 10, 0).items()}
def step(i):
    state, outputs, metrics = agent.update_wm(data, i)
    agent.update_acting_behavior(state, outputs, metrics, data)
for i in range(3): step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step(3)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    dt = getattr(e, "self_device_time_total", None)
    if dt is None: dt = e.self_cuda_time_total
    if dt > 0 and e.key.startswith("aten::"):
        rows.append((dt, e.count, e.key, str(e.input_shapes)[:110]))
tot = sum(r[0] for r in rows)
print(f"aten ops with GPU time: {tot/1e3:.3f} ms, {sum(r[1] for r in rows)} calls")
for dt, n, k, sh in sorted(rows, reverse=True)[:45]:
    print(f"{dt/1e3:8.3f} ms {n:5d}  {k:28s} {sh}")
