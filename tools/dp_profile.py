"""Launched under torchrun with 2+ ranks: CUPTI per-kernel durations of graph-replayed steps with data parallelism attached (rank 0 prints):
which NCCL kernels run, how long, and how much longer the step's own kernels take next to them."""
import os, sys, io, contextlib, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from genrl_b200 import ops
from genrl_b200.agent.genrl import GenRLAgent
from genrl_b200.dist import DataParallel
from torch.profiler import profile, ProfilerActivity

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))


class A: pass
args = A(); args.batch, args.length, args.horizon, args.action_dim, args.sequence, args.preset = 50, 48, 15, 10, "S2", "c2"
cfg, obs_space, act_spec = bench.make_cfg(args, dev, 16)
cfg.cuda_graph, cfg.overlap_streams = True, True
torch.manual_seed(1)
with contextlib.redirect_stdout(io.StringIO()):
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
if os.environ.get("NO_DP") != "1":
    DataParallel().attach(agent)
ops.set_noise(ops.PhiloxNoise(1000 + rank))
data = {k: v.to(dev) for k, v in bench.synthetic(50, 48, 10, rank).items()}


def step(i):
    state, outputs, metrics = agent.update_wm(data, i)
    agent.update_acting_behavior(state, outputs, metrics, data)


for i in range(5):
    step(i)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    step(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
if rank == 0:
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type.name != "CUDA":
            continue
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", e.name)).replace("void ", "").strip()
        agg[name][0] += 1; agg[name][1] += e.device_time_total
    tot = sum(v[1] for v in agg.values())
    print(f"world {world} dp {'off' if os.environ.get('NO_DP') == '1' else 'on'}: {ms:.2f} ms/step; kernel time {tot / 3e3:.2f} ms/step")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"{us / 3e3:9.3f} ms {n / 3:7.1f} launches {us / n:9.1f} us avg  {name}")
    for name, (n, us) in agg.items():
        if "nccl" in name.lower():
            print(f"NCCL: {us / 3e3:9.3f} ms/step {n / 3:5.1f} launches {us / n:9.1f} us avg  {name}")
agent._graphs.clear()
torch.cuda.synchronize(); dist.barrier()
sys.stdout.flush()
os._exit(0)
