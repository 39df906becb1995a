"""Where the step's GEMM time goes: every genrl_gemm / genrl_gemm_pair / chain call of one eager step is recorded (shape, operand majors, output
dtype, accumulate), each distinct call is then timed alone as a CUDA graph of R back-to-back launches on rotating operands (weights L2-hot,
like in the step), and the table is sorted by time x count."""
import os, sys, io, contextlib, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from genrl_b200 import ops
from genrl_b200._lib import lib
from genrl_b200.agent.genrl import GenRLAgent


class A: pass
args = A(); args.batch, args.length, args.horizon, args.action_dim, args.sequence, args.preset = 50, 48, 15, 10, "S2", "c2"
cfg, obs_space, act_spec = bench.make_cfg(args, "cuda:0", 16)
cfg.overlap_streams = True
torch.manual_seed(1)
with contextlib.redirect_stdout(io.StringIO()):
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
data = {k: v.cuda() for k, v in bench.synthetic(50, 48, 10, 0).items()}


def step(i):
    state, outputs, metrics = agent.update_wm(data, i)
    agent.update_acting_behavior(state, outputs, metrics, data)


for i in range(2):
    step(i)
torch.cuda.synchronize()
calls = collections.Counter()
orig = lib.genrl_gemm


def logged(*a):
    calls[(a[1], a[2], a[3], a[6], a[9], a[10], a[13], a[15], a[14] is not None)] += 1      # M N K tA tW in_dt c_dt acc bias
    return orig(*a)


lib.__dict__["genrl_gemm"] = logged
step(2)
torch.cuda.synchronize()
lib.__dict__["genrl_gemm"] = orig
R = 20
dev = "cuda:0"
rows = []
for (M, N, K, tA, tW, idt, cdt, acc, hb), cnt in calls.items():
    if idt != 1:
        continue
    pad = lambda v: (v + 7) // 8 * 8
    As = [torch.randn((K, pad(M)) if tA else (M, pad(K)), device=dev).bfloat16() for _ in range(3)]
    W = torch.randn((K, pad(N)) if tW else (N, pad(K)), device=dev).bfloat16()
    C = torch.zeros(M, pad(N), device=dev, dtype=torch.bfloat16 if cdt == 1 else torch.float32)
    bias = torch.zeros(N, device=dev) if hb else None

    def call(i):
        a = As[i % 3]
        orig(torch.cuda.current_stream().cuda_stream, M, N, K, a.data_ptr(), a.stride(0), tA, W.data_ptr(), W.stride(0), tW, 1, C.data_ptr(), C.stride(0), cdt,
             None if bias is None else bias.data_ptr(), acc)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        call(0); call(1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(R):
            call(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (2 * R) * 1e3
    rows.append((us * cnt, cnt, us, 2.0 * M * N * K / us / 1e6, (M, N, K, tA, tW, cdt, acc)))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
fl = sum(r[1] * 2.0 * r[4][0] * r[4][1] * r[4][2] for r in rows)
print(f"{len(rows)} distinct bf16 GEMM calls, {sum(r[1] for r in rows)} launches per step, {tot / 1e3:.2f} ms stand-alone, {fl / 1e12:.2f} TFLOP -> {fl / tot / 1e6:.0f} TFLOP/s")
print("   ms/step  calls   us/call  TFLOP/s   (M, N, K, transA, transW, c_dtype, accumulate)")
for t, cnt, us, tf, sh in rows[:45]:
    print(f"{t / 1e3:9.3f} {cnt:6d} {us:9.1f} {tf:8.0f}   {sh}")
