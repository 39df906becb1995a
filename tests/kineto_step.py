"""Real (warm, concurrent-free) per-kernel GPU durations of one eager step via torch.profiler (CUPTI)."""
import sys, os, io, contextlib, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from genrl_b200.agent.genrl import GenRLAgent
from torch.profiler import profile, ProfilerActivity

class A: pass
args = A(); args.batch, args.length, args.horizon, args.action_dim = 50, 48, 15, 10
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
from genrl_b200._lib import lib
_orig = lib.genrl_gemm
shapes = []
def _logged(*a):
    shapes.append((a[1], a[2], a[3], a[6], a[9], a[13], a[15]))     # M N K transA transW c_dtype accumulate
    return _orig(*a)
lib.__dict__["genrl_gemm"] = _logged
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(3)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type.name != "CUDA":
        continue
    name = re.sub(r"<.*", "", e.name); name = re.sub(r"\(.*", "", name)
    agg[name][0] += 1; agg[name][1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
tot = sum(v[1] for v in agg.values())
print(f"total GPU kernel time {tot/1e3:.3f} ms over {sum(v[0] for v in agg.values())} kernels/memops")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:36]:
    print(f"{us/1e3:9.3f} ms  {100*us/tot:5.1f}%  {n:5d}  {us/n:9.1f} us avg  {name}")

gem = sorted([e for e in prof.events() if e.device_type.name == "CUDA" and "gemm_tc_kernel" in e.name], key=lambda e: e.time_range.start)
print(len(gem), len(shapes))
if len(gem) == len(shapes):
    ag = collections.OrderedDict()
    for e, sh in zip(gem, shapes):
        d = ag.setdefault(sh, [0.0, 0]); d[0] += (e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total); d[1] += 1
    print("GEMM by shape (warm, kineto):   ms  calls  us/call  TFLOP/s  (M N K tA tW cdt acc)")
    for sh, (us, n) in sorted(ag.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{us/1e3:7.3f} {n:5d} {us/n:9.1f} {2.0*sh[0]*sh[1]*sh[2]*n/us/1e6:9.1f}   {sh}")
