"""Real (warm, concurrent-free) per-kernel GPU durations of one eager step via torch.profiler (CUPTI)."""
import sys, os, io, contextlib, collections, re
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
from genrl_b200._lib import lib
_orig = lib.genrl_gemm
shapes = []
def _logged(*a):
    shapes.append((a[1], a[2], a[3], a[6], a[9], a[13], a[15]))     # M N K transA transW c_dtype accumulate
    return _orig(*a)
lib.__dict__["genrl_gemm"] = _logged
ln_calls = {"genrl_ln_act_fwd": [], "genrl_ln_act_bwd": []}
def _wrap(name, mi):
    orig = getattr(lib, name)
    def f(*a):
        ln_calls[name].append((a[mi], a[mi + 1], bool(a[11]) if name.endswith("bwd") else None))
        return orig(*a)
    lib.__dict__[name] = f
_wrap("genrl_ln_act_fwd", 10); _wrap("genrl_ln_act_bwd", 14)
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

# implicit-GEMM conv launches in issue order (fwd: enc L1..L3 gather, dec D1..D3 sub-pixel; bwd: D3..D1 gather dgrad, L3..L1 sub-pixel dgrad)
for pat in ("conv_fwd_kernel", "conv_wgrad_kernel", "im2col", "col2im"):
    evs = sorted([e for e in prof.events() if e.device_type.name == "CUDA" and pat in e.name], key=lambda e: e.time_range.start)
    print(pat, " ".join(f"{(e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total):.0f}" for e in evs), "us")

# LayerNorm launches by shape
def _t(e): return e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
for cname, pats in (("genrl_ln_act_fwd", ("ln_fwd_fast_kernel", "ln_act_fwd_kernel")), ("genrl_ln_act_bwd", ("ln_bwd_colg_kernel", "ln_bwd_fast_kernel", "ln_act_bwd_kernel"))):
    evs = sorted([e for e in prof.events() if e.device_type.name == "CUDA" and any(p in e.name for p in pats)], key=lambda e: e.time_range.start)
    calls = ln_calls[cname]
    print(cname, len(evs), len(calls))
    if len(evs) == len(calls):
        ag = collections.OrderedDict()
        for e, c in zip(evs, calls):
            d = ag.setdefault(c, [0.0, 0]); d[0] += _t(e); d[1] += 1
        for c, (us, n) in sorted(ag.items(), key=lambda kv: -kv[1][0])[:14]:
            M, D = c[0], c[1]
            print(f"  {us/1e3:7.3f} ms {n:4d} calls {us/n:8.1f} us/call  {M*D*2*(2 if c[2] is None else 3)/(us/n)/1e3:7.0f} GB/s  {c}")
