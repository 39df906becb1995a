#!/usr/bin/env python
"""bench.py — seq-frames/s through one full GenRL update (world model + connector + acting behaviour) on N B200s.

  python bench.py --gpus N --steps K --warmup W              # this framework (CUDA kernels through the agent API)
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1], with T=48 because GenRLAgent asserts batch_length % 8 == 0):
stickman (A=10), per-GPU batch B=50 x T=48, 64x64x3 uint8 observations, deter=hidden=1024, stoch 32x32, imag_horizon=15,
step sequence S2 = agent.update_wm() [WM + in-update connector step] + agent.update_acting_behavior().
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", type=int, default=16, choices=[16, 32], help="16 = bf16 tensor-core mode, 32 = fp32 parity mode")
    ap.add_argument("--batch", type=int, default=50)
    ap.add_argument("--length", type=int, default=48)
    ap.add_argument("--horizon", type=int, default=15)
    ap.add_argument("--action-dim", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="single-stream execution (no connector/critic overlap)")
    ap.add_argument("--eager", action="store_true", help="do not capture the update calls into CUDA graphs")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def make_cfg(args, device, precision):
    from genrl_b200.config import genrl_config, make_specs
    cfg = genrl_config(device=device, precision=precision, batch_size=args.batch, batch_length=args.length, imag_horizon=args.horizon)
    obs_space, act_spec = make_specs(action_dim=args.action_dim, clip_dim=512)
    return cfg, obs_space, act_spec


def synthetic(B, T, A, seed):
    from genrl_b200.synthetic import synthetic_batch
    return synthetic_batch(B, T, A, clip_dim=512, seed=seed)


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """the reference algorithm (oracle port: same torch-CPU ops as the reference's own code path) on the host cores"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import genrl_oracle as O
    from genrl_b200.config import genrl_config, make_specs
    from genrl_b200.agent.genrl import GenRLAgent
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg, obs_space, act_spec = make_cfg(args, "cpu", 32)
    torch.manual_seed(1)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        init = {k: v.detach().clone() for k, v in GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec).state_dict().items()}
    # bounded sample of the workload: pick the sample batch so the whole run ends within a few minutes
    probe_B = 2
    agent = O.OracleAgent(init, cfg, kind="genrl")
    data = synthetic(probe_B, args.length, args.action_dim, 0)
    t0 = time.perf_counter(); agent.update(data); per_row = (time.perf_counter() - t0) / probe_B
    budget = 150.0
    Bs = int(max(1, min(args.batch, budget / max(per_row, 1e-3) / max(1, args.steps + args.warmup))))
    data = synthetic(Bs, args.length, args.action_dim, 1)
    for _ in range(args.warmup):
        agent.update(data)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        agent.update(data)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = Bs * args.length / dt
    sample = f"B={Bs} of {args.batch} sequences x T={args.length}, H={args.horizon}, fp32, full update (WM+connector+behaviour)"
    line = {"impl": "reference", "metric": "seq_frames_per_s_full_genrl_update", "value": value, "unit": "seq-frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": "seq-frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "seq-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": f"GenRL full update S2 (update_wm incl. connector + update_acting_behavior), stickman A={args.action_dim}, "
                        f"per-GPU batch {args.batch}x{args.length}, 64x64x3 u8 obs, deter=1024, stoch=32x32, imag_horizon={args.horizon}",
            "batch": args.batch, "length": args.length, "horizon": args.horizon, "parallelism": f"dp{args.gpus}",
            "l2": "5 rotating input batches (150 MB) and a multi-GB per-step working set exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist
    from genrl_b200 import ops
    from genrl_b200._lib import lib
    from genrl_b200.agent.genrl import GenRLAgent
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: the genrl_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator comes up: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(dev))
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    cfg, obs_space, act_spec = make_cfg(args, dev, args.precision)
    cfg.cuda_graph = not args.eager
    cfg.overlap_streams = not args.no_overlap
    torch.manual_seed(1)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    if world > 1:
        from genrl_b200.dist import DataParallel
        DataParallel().attach(agent)
    ops.set_noise(ops.PhiloxNoise(1000 + rank))
    nb = 5
    host = [{k: v.pin_memory() for k, v in synthetic(args.batch, args.length, args.action_dim, 100 * rank + i).items()} for i in range(nb)]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    def step(data, i):
        state, outputs, metrics = agent.update_wm(data, i)
        start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
        return metrics

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n

    l_before = lib.launches
    step(devb[0], 0)                                        # first call is always eager: count the kernels of one step
    launches = lib.launches - l_before
    for i in range(1, max(args.warmup, 3)):
        step(devb[i % nb], i)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_step = timed(lambda i: step(devb[i % nb], i), args.steps)

    # end-to-end: host (pinned) batch -> device copy, public API call, device -> host read of the step's loss, every step.
    # The loader side is double-buffered the way a replay iterator would be: the H2D copy of batch i+1 is enqueued on a copy stream
    # before step i is launched, so it overlaps the step; every copy and every loss read-back is inside the timed region.
    d2h_bytes = 4
    main_s = torch.cuda.current_stream()
    copy_s = torch.cuda.Stream(device=dev)
    slots = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        slot = i % 2
        copy_s.wait_event(free[slot])                       # the step that last read this slot has finished
        with torch.cuda.stream(copy_s):
            for k, v in host[i % nb].items():
                slots[slot][k].copy_(v, non_blocking=True)
            ready[slot].record(copy_s)

    def e2e_step(i):
        slot = i % 2
        if i == 0:
            prefetch(0)
        main_s.wait_event(ready[slot])
        if i + 1 < args.steps:
            prefetch(i + 1)
        m = step(slots[slot], i)
        free[slot].record(main_s)
        float(m["model_loss"])                              # D2H read of this step's result (synchronises)
    ms_e2e = timed(e2e_step, args.steps)
    clk = clocks.stop() if rank == 0 else None

    frames = world * args.batch * args.length
    line = {"metric": "seq_frames_per_s_full_genrl_update", "value": frames / (ms_step * 1e-3), "unit": "seq-frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == 16 else "f32", "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "seq-frames/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h_bytes, "pipeline": "pinned host batch i+1 copied on a second stream while step i runs; loss of every step read back"},
            "gpu_launches": int(launches), "clocks": clk}

    if rank == 0 and world == 1 and not args.no_roofline:
        line["roofline"], line["kernel_breakdown"] = roofline_pass(step, devb, nb, lib, args, agent)
    elif rank == 0 and world > 1:
        # the roofline pass re-runs eager steps on ONE rank; with data parallelism attached those steps contain NCCL collectives
        # the other ranks would never join (it hung an 8-GPU run), and the kernel metric is a single-GPU quantity anyway
        line["roofline_note"] = "single-GPU kernel metric: reported by the N=1 run only"
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, agent)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # captured graphs hold NCCL work: drop them, drain the device, and leave without the (blocking) communicator teardown
        agent._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def roofline_pass(step, devb, nb, lib, args, agent):
    """Roofline of the dominant kernel (the tcgen05 GEMM: 40 % of the step's GPU time, profiles/).

    1. two instrumented EAGER steps with CUDA events on the launching stream around every C-ABI call give each entry point's share
       of the step (the graph-captured steps of the timed region cannot be instrumented per kernel);
    2. every distinct GEMM call of the step (shape, operand majors, dtypes, accumulate) is then replayed on fresh operands, R launches
       back to back between two CUDA events, so the per-launch duration excludes the ~8 us event/launch floor that dominates
       a 10 us kernel when it is timed alone.  achieved = sum(2MNK x calls) / sum(duration x calls)."""
    from genrl_b200 import ops
    pk = peaks()
    graphs_on, agent.cfg.cuda_graph = agent.cfg.cuda_graph, False
    step(devb[0], 0)
    torch.cuda.synchronize()
    lib.profile = []
    n = 2
    for i in range(n):
        step(devb[i % nb], i)
    torch.cuda.synchronize()
    prof, lib.profile = lib.profile, None
    agent.cfg.cuda_graph = graphs_on
    agg, shapes = {}, {}
    for name, e0, e1, a, allargs in prof:
        d = agg.setdefault(name, {"ms": 0.0, "calls": 0, "flops": 0.0})
        d["ms"] += e0.elapsed_time(e1); d["calls"] += 1
        if name == "genrl_gemm":
            d["flops"] += 2.0 * a[0] * a[1] * a[2]
            key = (allargs[1], allargs[2], allargs[3], allargs[6], allargs[9], allargs[10], allargs[13], allargs[15], allargs[14] is not None)
            shapes[key] = shapes.get(key, 0) + 1
    total = sum(d["ms"] for d in agg.values())
    breakdown = {k: {"ms_per_step": v["ms"] / n, "calls_per_step": v["calls"] // n, "share": v["ms"] / total} for k, v in
                 sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:8]}
    g = agg.get("genrl_gemm")
    roof = None
    if g and g["ms"] > 0:
        dev = devb[0]["action"].device
        st = torch.cuda.current_stream()
        flops = dur_ms = 0.0
        R = 20
        for (M, N, K, tA, tW, in_dt, c_dt, acc, has_bias), cnt in shapes.items():
            idt = torch.bfloat16 if in_dt == 1 else torch.float32
            cdt = torch.bfloat16 if c_dt == 1 else torch.float32
            pad = lambda v: (v + 7) // 8 * 8
            A = torch.randn((K, pad(M)) if tA else (M, pad(K)), device=dev).to(idt)
            W = torch.randn((K, pad(N)) if tW else (N, pad(K)), device=dev).to(idt)
            C = torch.zeros(M, pad(N), device=dev, dtype=cdt)
            bias = torch.zeros(N, device=dev) if has_bias else None
            call = lambda: lib.genrl_gemm(st.cuda_stream, M, N, K, A.data_ptr(), A.stride(0), tA, W.data_ptr(), W.stride(0), tW, in_dt,
                                          C.data_ptr(), C.stride(0), c_dt, None if bias is None else bias.data_ptr(), acc)
            call(); call()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(st)
            for _ in range(R):
                call()
            e1.record(st)
            torch.cuda.synchronize()
            per = e0.elapsed_time(e1) / R
            flops += 2.0 * M * N * K * cnt
            dur_ms += per * cnt
        achieved = flops / (dur_ms * 1e-3) / 1e12
        bf16 = args.precision == 16
        peak = pk["tf_sustained"] if bf16 else 75.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
        if bf16 and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")      # ncu dram__bytes_read+write, mean over the step's GEMM launches
        calls = sum(shapes.values())
        roof = {"kernel": "gemm_tc_kernel (tcgen05)" if bf16 else "gemm_simt_kernel (fp32 parity mode)", "bound": "tensor", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": (pk["source"] + " bf16 sustained") if bf16 else "nominal fp32 FFMA",
                "share_of_step": g["ms"] / total, "launch_avg_us": dur_ms / calls * 1e3, "launches_per_step": calls // n,
                "algorithmic_flops_per_launch": flops / calls, "algorithmic_flops_per_step": flops / n,
                "method": f"{len(shapes)} distinct GEMM calls of the step replayed {R}x back to back between CUDA events"}
    return roof, breakdown


def cpu_baseline(args, agent):
    """the oracle port (reference algorithm, torch CPU ops) on this box's host cores, on a bounded sample of the workload"""
    from oracle import genrl_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg, _, _ = make_cfg(args, "cpu", 32)
    init = {k: v.detach().cpu().clone() for k, v in agent.state_dict().items()}
    Bs = max(1, min(args.batch, 6))
    data = synthetic(Bs, args.length, args.action_dim, 7)
    oracle = O.OracleAgent(init, cfg, kind="genrl")
    oracle.update(data)
    t0 = time.perf_counter()
    n = 2
    for _ in range(n):
        oracle.update(data)
    dt = (time.perf_counter() - t0) / n
    return {"value": Bs * args.length / dt, "unit": "seq-frames/s", "cores": threads, "kind": "port",
            "sample": f"B={Bs} of {args.batch} sequences x T={args.length}, H={args.horizon}, fp32 torch-CPU, {n} timed updates after 1 warm-up"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
