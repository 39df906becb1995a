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
    ap.add_argument("--preset", default="c2", choices=["c2", "c4", "c5"],
                    help="BASELINE.json configs: c2 stickman (default) | c4 kitchen (A = 9, extra low-dimensional `state` observation key) | "
                         "c5 large RSSM per-GPU shard (deter = hidden = 4096, stoch 48x48, batch 16x64, imag_horizon 32)")
    ap.add_argument("--global-batch", type=int, default=0, help="strong scaling: total sequences, split evenly over the ranks (default: --batch per rank, weak scaling)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"], help="--impl reference: host cores (default) or torch-eager on the GPU")
    ap.add_argument("--ref-autocast", action="store_true", help="--impl reference --ref-device cuda: fp16 autocast (the reference's precision: 16)")
    ap.add_argument("--ref-batch", type=int, default=0, help="--impl reference: sample size in sequences (default REF_SAMPLE_B)")
    ap.add_argument("--sequence", default="S2", choices=["S1", "S2", "S3"],
                    help="S1 update_wm; S2 update_wm + update_acting_behavior (default, BASELINE configs[1]); S3 train.py:275,280,340 "
                         "(update_wm + second connector step + update_imag_behavior with the VLM reward)")
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


def apply_preset(args):
    if args.preset == "c4":
        args.action_dim = 9
    elif args.preset == "c5":
        args.batch, args.length, args.horizon = 16, 64, 32


def make_cfg(args, device, precision):
    from genrl_b200.config import genrl_config, make_specs, Spec
    over = {"imag_reward_fn": "video_text_reward"} if getattr(args, "sequence", "S2") == "S3" else {}
    if args.preset == "c5":
        big = dict(ensemble=1, hidden=4096, deter=4096, stoch=48, discrete=48, norm="layer", std_act="softplus", min_std=0.1)
        over.update(rssm=dict(big, single_obs_posterior=True), connector_rssm=dict(big, single_obs_posterior=False, learn_initial=True))
    cfg = genrl_config(device=device, precision=precision, batch_size=args.batch, batch_length=args.length, imag_horizon=args.horizon, **over)
    obs_space, act_spec = make_specs(action_dim=args.action_dim, clip_dim=512)
    if args.preset == "c4":
        obs_space["state"] = Spec((30,), "float32")
    return cfg, obs_space, act_spec


EXTRA_KEYS = {}


def synthetic(B, T, A, seed):
    from genrl_b200.synthetic import synthetic_batch
    d = synthetic_batch(B, T, A, clip_dim=512, seed=seed)
    for k, n in EXTRA_KEYS.items():
        d[k] = torch.zeros(B, T, n)
    return d


# ------------------------------------------------------------------------------------------------------------------
REF_SAMPLE_B = 25        # fixed sample of the reference arm: half of the workload's 50 sequences, every run, every N


def run_reference(args):
    """The reference algorithm (oracle port: the same torch ops in the same order as the reference's own code path) on the host cores.
    Fixed sample (REF_SAMPLE_B of the workload's sequences, full T / H / architecture) so the number is comparable across runs.
    --ref-device cuda runs the same port as torch-eager kernels on the GPU (cuDNN / cuBLAS via torch): the 'existing Blackwell path' bar
    of BASELINE.md §3 (reported under profiles/, not the driver's reference arm)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import genrl_oracle as O
    from genrl_b200.agent.genrl import GenRLAgent
    on_gpu = args.ref_device == "cuda"
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg, obs_space, act_spec = make_cfg(args, "cpu", 32)
    torch.manual_seed(1)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        init = {k: v.detach().clone() for k, v in GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec).state_dict().items()}
    Bs = args.batch if on_gpu else min(args.batch, args.ref_batch or REF_SAMPLE_B)
    data = synthetic(Bs, args.length, args.action_dim, 1)
    if on_gpu:
        torch.set_default_device("cuda")
        init = {k: v.cuda() for k, v in init.items()}
        data = {k: v.cuda() for k, v in data.items()}
    agent = O.OracleAgent(init, cfg, kind="genrl")
    amp = torch.autocast("cuda", dtype=torch.float16) if (on_gpu and args.ref_autocast) else contextlib.nullcontext()

    def sync():
        if on_gpu:
            torch.cuda.synchronize()
    with amp:
        for _ in range(max(1, args.warmup if on_gpu else min(args.warmup, 1))):
            agent.update(data)
        sync()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            agent.update(data)
        sync()
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = Bs * args.length / dt
    where = ("torch-eager CUDA kernels on the GPU, " + ("fp16 autocast" if args.ref_autocast else "fp32")) if on_gpu else "fp32 on the host CPU"
    sample = f"B={Bs} of {args.batch} sequences x T={args.length}, H={args.horizon}, {where}, full update (WM+connector+behaviour)"
    line = {"impl": "reference", "metric": "seq_frames_per_s_full_genrl_update", "value": value, "unit": "seq-frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": (args.warmup if on_gpu else min(args.warmup, 1)), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16" if (on_gpu and args.ref_autocast) else "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": "seq-frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "seq-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if on_gpu:
        line["device"] = "cuda (torch eager)"
    print(json.dumps(line))


SEQ_TEXT = {"S1": "S1 (update_wm incl. connector)", "S2": "S2 (update_wm incl. connector + update_acting_behavior)",
            "S3": "S3 (train.py:275,280,340: update_wm + second connector step + update_imag_behavior with the VLM reward)"}


def workload_config(args):
    env = {"c2": "stickman", "c4": "kitchen (extra `state` key)", "c5": "stickman, large RSSM"}[args.preset]
    dims = "deter=4096, stoch=48x48" if args.preset == "c5" else "deter=1024, stoch=32x32"
    return {"workload": f"GenRL full update {SEQ_TEXT[args.sequence]}, {env} A={args.action_dim}, "
                        f"per-GPU batch {args.batch}x{args.length}, 64x64x3 u8 obs, {dims}, imag_horizon={args.horizon}", "preset": args.preset,
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
    if args.global_batch:
        assert args.global_batch % world == 0, f"--global-batch {args.global_batch} does not divide over {world} ranks"
        args.batch = args.global_batch // world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: the genrl_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator comes up: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        # The four gradient all-reduces (219 MB) sit on the critical path; with NCCL's default CTA count they run at ~330 GB/s on 2 GPUs.
        # Measured (profiles/r02_notes.md §24): NCCL_MIN_CTAS=32 takes the 2-GPU step from 21.0 to 20.4 ms and is neutral on 8 GPUs (NVLS already
        # uses 24-32 channels).  An explicit setting in the environment wins.
        os.environ.setdefault("NCCL_MIN_CTAS", "32")
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

    if args.sequence == "S3":
        # train.py:289-319 caches the connector's rollout from the task prompt once; synthetic stand-in: random one-hot latents
        Hn, Nn = args.horizon + 1, args.batch * args.length
        idx = torch.randint(0, 32, (Hn, Nn, 32), device=dev, generator=torch.Generator(device=dev).manual_seed(5))
        tgt = torch.nn.functional.one_hot(idx, 32).float()
        agent.unconditional_target = {"stoch": tgt, "logit": torch.zeros_like(tgt), "deter": torch.zeros(Hn, Nn, 1024, device=dev)}

    def step(data, i):
        state, outputs, metrics = agent.update_wm(data, i)
        if args.sequence == "S2":
            start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
        elif args.sequence == "S3":
            _, metrics = agent.wm.update_additional_detached_modules(data, outputs, metrics)
            start, metrics = agent.update_imag_behavior(state, outputs, metrics, seq_data=data)
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

    # the same public call fed by the on-GPU replay store: per step the host samples B (episode, offset) pairs with the
    # reference's numpy calls, uploads B int64 start rows from pinned memory, one gather kernel per key builds the batch in HBM
    import numpy as np
    from genrl_b200.replay import ReplayBuffer, make_replay_loader
    ep_len, n_eps = 4 * args.length, 24
    rbuf = ReplayBuffer(length=args.length, capacity=n_eps * ep_len, device=dev, max_episode_steps=ep_len)
    for e in range(n_eps):
        epi = {k: v[0].numpy() for k, v in synthetic(1, ep_len, args.action_dim, 1000 + 97 * rank + e).items()}
        rbuf.store_episode(episode=epi, run_checks=False)
    np.random.seed(7 + rank)
    riter = iter(make_replay_loader(rbuf, args.batch))

    def replay_step(i):
        m = step(next(riter), i)
        float(m["model_loss"])
    for i in range(2):
        replay_step(i)
    ms_replay = timed(replay_step, args.steps)

    frames = world * args.batch * args.length
    line = {"metric": "seq_frames_per_s_full_genrl_update", "value": frames / (ms_step * 1e-3), "unit": "seq-frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if args.global_batch else "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == 16 else "f32", "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "seq-frames/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h_bytes, "pipeline": "pinned host batch i+1 copied on a second stream while step i runs; loss of every step read back"},
            "e2e_replay": {"value": frames / (ms_replay * 1e-3), "unit": "seq-frames/s", "ms_per_step": ms_replay, "h2d_bytes_per_step": 8 * args.batch,
                           "d2h_bytes_per_step": d2h_bytes, "pipeline": "batch gathered in HBM from the on-GPU episode store (genrl_b200.replay) behind "
                           "next(replay_iter); host samples the windows with the reference's numpy calls; loss of every step read back"},
            "gpu_launches": int(launches), "clocks": clk}

    if rank == 0 and world == 1 and not args.no_roofline:
        line.update(roofline_pass(step, devb, nb, lib, args, agent, ms_step))
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


def conv_algorithmic_flops(frames, depth=48, img=64):
    """Algorithmic FLOPs of the conv stacks per update: encoder k4 s2 x4, decoder convT k5,k5,k6,k6 s2.  The decoder's first
    layer acts on a 1x1 map and runs as a plain GEMM (counted with the GEMMs), the rest on the implicit-GEMM kernels.
    Returns (fwd-type kernel flops: forward + dgrad, wgrad kernel flops)."""
    enc, h, cin = [], img, 3
    for i in range(4):
        cout, ho = depth * 2 ** i, (h - 4) // 2 + 1
        enc.append(ho * ho * cout * cin * 16); h, cin = ho, cout                  # MACs per frame
    dec, h, cin = [], 1, 32 * depth
    for i, k in enumerate((5, 5, 6, 6)):
        cout = 3 if i == 3 else depth * 2 ** (2 - i)
        dec.append(h * h * cin * cout * k * k); h, cin = (h - 1) * 2 + k, cout
    fwd = sum(enc) + sum(dec[1:])
    dgrad = sum(enc[1:]) + sum(dec[1:])                                           # no gradient to the pixels
    wgrad = sum(enc) + sum(dec[1:])
    return 2.0 * frames * (fwd + dgrad), 2.0 * frames * wgrad


def rssm_algorithmic_bytes(B, T, A, E=1536, S=32, K=32, D=1024, Hd=1024, s=4, s_w=4):
    """SURVEY §8d 'Algorithmic bytes - RSSM observe fwd': (compulsory, streamed)."""
    N = B * T
    io = N * (E + A + 1) * s + N * (4 * S * K + D) * s
    p_rec = (S * K + A) * Hd + (Hd + D) * 3 * D + D * Hd + Hd * S * K
    p_post = E * Hd + Hd * S * K
    return io + (p_rec + p_post) * s_w, io + p_post * s_w + T * p_rec * s_w


def cupti_kernels(fn, n):
    """per-kernel device durations (CUPTI via torch.profiler) of n calls of fn, graph replays included; {name: [launches, us]}"""
    import collections, re
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(n):
            fn(i)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type.name != "CUDA":
            continue
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", e.name)).replace("void ", "").strip()
        t = e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        agg[name][0] += 1; agg[name][1] += t
    return agg


def roofline_pass(step, devb, nb, lib, args, agent, ms_step):
    """Rooflines of the three kernel families BASELINE's metric names, measured LIVE on the graph-replayed steps of the timed
    configuration: per-kernel device durations come from CUPTI (torch.profiler) over n replayed steps, algorithmic work from the
    arguments of the C-ABI calls one eager step makes (GEMM: 2MNK) or from the layer table (conv stacks).
      roofline       tcgen05 GEMM family (dominant: largest share of the step)  -> tensor roofline, measured sustained bf16 peak
      roofline_conv  implicit-GEMM conv kernels                                 -> tensor roofline
      roofline_rssm  persistent scan kernel + the whole observe() forward       -> HBM roofline, §8d compulsory and streamed bytes"""
    pk = peaks()
    bf16 = args.precision == 16
    # 1. work of one step, from one instrumented eager step (no timing taken here)
    graphs_on, agent.cfg.cuda_graph = agent.cfg.cuda_graph, False
    counts = {"gemm_flops": 0.0, "gemm_calls": 0, "scan_calls": 0}
    hooks = {}

    def hook(name, fn):
        orig = getattr(lib, name)
        hooks[name] = orig
        def wrapped(*a):
            fn(a)
            return orig(*a)
        lib.__dict__[name] = wrapped
    def on_gemm(a):
        counts["gemm_flops"] += 2.0 * a[1] * a[2] * a[3]; counts["gemm_calls"] += 1
    hook("genrl_gemm", on_gemm)
    hook("genrl_gemm_ln_fwd", on_gemm)

    def on_chain(a):          # (stream, M, N, L, x, ldx, K0, ...): layer 0 has K0 inputs, the others N
        counts["gemm_flops"] += 2.0 * a[1] * a[2] * (a[6] + (a[3] - 1) * a[2]); counts["gemm_calls"] += 1
    hook("genrl_mlp_chain_fwd", on_chain)
    hook("genrl_rssm_scan_fwd", lambda a: counts.__setitem__("scan_calls", counts["scan_calls"] + 1))
    step(devb[0], 0)
    torch.cuda.synchronize()
    for k, v in hooks.items():
        lib.__dict__[k] = v
    agent.cfg.cuda_graph = graphs_on
    for i in range(3):                      # back on the captured graphs (the eager step dropped nothing, but re-warm anyway)
        step(devb[i % nb], i)
    # 2. CUPTI durations over replayed steps
    n = 3
    agg = cupti_kernels(lambda i: step(devb[i % nb], i), n)
    total_us = sum(v[1] for v in agg.values())
    # 2b. the same with the two-stream overlap off: every kernel runs alone, so its duration is the kernel's own (in the overlapped step a
    #     GEMM's duration also contains the time it shares the SMs with the other stream's kernels)
    agg_alone = None
    if getattr(agent.cfg, "overlap_streams", False):
        ov = agent.cfg.overlap_streams
        agent.cfg.overlap_streams = False
        for beh in (getattr(agent, "_acting_behavior", None), getattr(agent, "_imag_behavior", None)):
            if beh is not None and getattr(beh, "cfg", None) is not agent.cfg:
                beh.cfg.overlap_streams = False
        saved_graphs, agent._graphs = agent._graphs, {}
        for i in range(4):
            step(devb[i % nb], i)
        agg_alone = cupti_kernels(lambda i: step(devb[i % nb], i), n)
        agent.cfg.overlap_streams = ov
        agent._graphs = saved_graphs
        for i in range(2):
            step(devb[i % nb], i)
    breakdown = {k: {"ms_per_step": v[1] / n / 1e3, "launches_per_step": v[0] / n, "avg_us": v[1] / v[0], "share_of_kernel_time": v[1] / total_us}
                 for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]}
    breakdown["_total"] = {"kernel_ms_per_step": total_us / n / 1e3, "launches_per_step": sum(v[0] for v in agg.values()) / n,
                           "step_ms": ms_step, "source": f"CUPTI (torch.profiler) over {n} graph-replayed steps"}
    fam = lambda *pats: [sum(v[0] for k, v in agg.items() if any(p in k for p in pats)), sum(v[1] for k, v in agg.items() if any(p in k for p in pats))]
    GEMM_FAMILY = ("gemm_tc_kernel", "gemm_tc_persist_kernel", "gemm_ln_kernel", "gemm_simt_kernel", "gemm_pair_kernel", "mlp_chain_fwd_kernel")
    g_l, g_us = fam(*GEMM_FAMILY)
    alone_us = sum(v[1] for k, v in agg_alone.items() if any(p in k for p in GEMM_FAMILY)) if agg_alone else None
    roof = None
    if g_us > 0:
        in_step = counts["gemm_flops"] * n / (g_us * 1e-6) / 1e12
        achieved = counts["gemm_flops"] * n / (alone_us * 1e-6) / 1e12 if alone_us else in_step
        peak = pk["tf_sustained"] if bf16 else 75.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02b_gemm_traffic.json")          # ncu pass of the second half of round 2
        if not os.path.exists(tpath):
            tpath = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
        if not os.path.exists(tpath):
            tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
        if bf16 and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")      # ncu dram__bytes_read+write, mean over the step's GEMM launches
        roof = {"kernel": "tcgen05 GEMM family: gemm_tc_kernel (128xBN), gemm_pair_kernel (2-CTA 256x256), mlp_chain_fwd_kernel (Linear+LN chains), "
                          "gemm_tc_persist_kernel, gemm_ln_kernel" if bf16 else "gemm_simt_kernel (fp32 parity mode)",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": (pk["source"] + " bf16 sustained") if bf16 else "nominal fp32 FFMA",
                "achieved_in_overlapped_step": in_step, "frac_in_overlapped_step": in_step / peak,
                "share_of_step": g_us / total_us, "launch_avg_us": (alone_us or g_us) / g_l, "launches_per_step": g_l / n,
                "algorithmic_flops_per_launch": counts["gemm_flops"] * n / g_l, "algorithmic_flops_per_step": counts["gemm_flops"],
                "method": f"sum of 2MNK over the step's {counts['gemm_calls']} GEMM / chain calls / sum of their CUPTI kernel durations over {n} graph-replayed "
                          "steps with the two-stream overlap off (each kernel alone on the GPU); *_in_overlapped_step: the same inside the timed two-stream steps"}
    out = {"roofline": roof, "kernel_breakdown": breakdown}
    # conv stacks
    frames = args.batch * args.length
    f_fwd, f_wg = conv_algorithmic_flops(frames)
    src = agg_alone or agg               # kernel durations with each kernel alone on the GPU when that pass ran
    fam = lambda *pats: [sum(v[0] for k, v in src.items() if any(p in k for p in pats)), sum(v[1] for k, v in src.items() if any(p in k for p in pats))]
    c_l, c_us = fam("conv_fwd_kernel")
    w_l, w_us = fam("conv_wgrad_kernel")
    if bf16 and c_us > 0 and w_us > 0:
        peak = pk["tf_sustained"]
        out["roofline_conv"] = {
            "bound": "tensor", "unit": "TFLOP/s", "peak": peak, "peak_source": pk["source"] + " bf16 sustained",
            "fwd_dgrad": {"kernel": "conv_fwd_kernel (implicit GEMM, tcgen05)", "launches_per_step": c_l / n, "ms_per_step": c_us / n / 1e3,
                          "algorithmic_flops_per_step": f_fwd, "achieved": f_fwd * n / (c_us * 1e-6) / 1e12, "frac": f_fwd * n / (c_us * 1e-6) / 1e12 / peak},
            "wgrad": {"kernel": "conv_wgrad_kernel (implicit GEMM, tcgen05)", "launches_per_step": w_l / n, "ms_per_step": w_us / n / 1e3,
                      "algorithmic_flops_per_step": f_wg, "achieved": f_wg * n / (w_us * 1e-6) / 1e12, "frac": f_wg * n / (w_us * 1e-6) / 1e12 / peak},
            "achieved": (f_fwd + f_wg) * n / ((c_us + w_us) * 1e-6) / 1e12, "frac": (f_fwd + f_wg) * n / ((c_us + w_us) * 1e-6) / 1e12 / peak,
            "tensor_pipe_util": "ncu sm__pipe_tensor* per launch: profiles/ (conv digest)"}
    # RSSM
    s_l, s_us = fam("rssm_scan_fwd_kernel")
    if s_us > 0:
        B, T, D = args.batch, args.length, 1024
        N = B * T
        sa = 2 if bf16 else 4
        # what the scan kernel itself has to move once: gx in (fp32), deter out (fp32), m*h_prev out (compute dtype), pre-activations z
        # and LayerNorm statistics out (saved for the backward), the W_h slice in (fp32 master, converted on load)
        kbytes = N * 3 * D * 4 + N * D * 4 + N * D * sa + N * 3 * D * 4 + N * 2 * 4 + 3 * D * D * 4
        us = s_us / s_l
        comp, streamed = rssm_algorithmic_bytes(B, T, args.action_dim)
        obs_us = time_observe(agent, devb[0])
        out["roofline_rssm"] = {
            "kernel": "rssm_scan_fwd_kernel (persistent, one launch per sequence batch)", "bound": "hbm", "unit": "GB/s", "peak": pk["hbm"],
            "peak_source": pk["source"] + " HBM copy", "launches_per_step": s_l / n, "us_per_launch": us, "us_per_time_step": us / T,
            "kernel_bytes_per_launch": kbytes, "achieved": kbytes / (us * 1e-6) / 1e9, "frac": kbytes / (us * 1e-6) / 1e9 / pk["hbm"],
            "observe_fwd": {"what": "whole EnsembleRSSM.observe forward (batched posterior / input / prior GEMMs + LayerNorms + samples + the scan), "
                                    "eager launches between CUDA events, best of 5", "us": obs_us,
                            "compulsory_bytes_fp32": comp, "streamed_bytes_fp32": streamed,
                            "gbs_vs_compulsory": comp / (obs_us * 1e-6) / 1e9, "frac_vs_compulsory": comp / (obs_us * 1e-6) / 1e9 / pk["hbm"],
                            "gbs_vs_streamed": streamed / (obs_us * 1e-6) / 1e9, "frac_vs_streamed": streamed / (obs_us * 1e-6) / 1e9 / pk["hbm"],
                            "note": "SURVEY §8d: 'streamed' is the figure when the recurrent weights exceed the 33 MB of aggregate shared memory "
                                    "(fp32: 37.8 MB) and are re-read every time step; this design keeps the W_h slice resident instead"}}
    # the weight-streaming regime of the same time loop (BASELINE configs[4] per-GPU shard: deter 4096, 16 sequences x 64 steps), measured live as a
    # stand-alone launch: this is the shape the HBM roofline of the north star is about (3 D^2 bf16 weight bytes per time step)
    if "roofline_rssm" in out and bf16:
        try:
            out["roofline_rssm"]["weight_streaming_d4096"] = time_stream_scan(pk)
        except Exception as e:          # noqa: BLE001 - a bench line without this sub-object is still valid
            out["roofline_rssm"]["weight_streaming_d4096"] = {"error": str(e)[:200]}
    ss_l, ss_us = fam("rssm_scan_stream_fwd_kernel")
    if ss_us > 0:
        # weight-streaming regime (preset c5, deter = 4096): per time step the kernel streams 3 D^2 bf16 weight bytes; per launch also gx in
        # (fp32), deter out (fp32), m*h_prev out and the pre-activations z out (bf16)
        B, T = args.batch, args.length
        D = int(agent.wm.rssm._deter)
        N = B * T
        wbytes = 3 * D * D * 2
        kbytes = T * wbytes + N * (3 * D * 4 + D * 4 + D * 2 + 3 * D * 2 + 8)
        us = ss_us / ss_l
        out["roofline_rssm"] = {
            "kernel": "rssm_scan_stream_fwd_kernel (persistent, W_h streamed once per time step into mma.sync A fragments)", "bound": "hbm", "unit": "GB/s",
            "peak": pk["hbm"], "peak_source": pk["source"] + " HBM copy", "launches_per_step": ss_l / n, "us_per_launch": us, "us_per_time_step": us / T,
            "weight_bytes_per_time_step": wbytes, "kernel_bytes_per_launch": kbytes, "achieved": kbytes / (us * 1e-6) / 1e9,
            "frac": kbytes / (us * 1e-6) / 1e9 / pk["hbm"],
            "note": "up to 75 MB of the weights are kept in L2 between time steps (evict_last), so part of the streamed bytes is served by L2, not HBM"}
    return out


def time_stream_scan(pk, B=16, T=64, D=4096):
    """rssm_scan_stream_fwd_kernel at the configs[4] shard shape: CUDA events around 5 launches after 2 warm-ups"""
    from genrl_b200 import ops
    dev = "cuda:0"
    with torch.no_grad(), ops.use_dtype(torch.bfloat16):
        g = torch.Generator(device=dev).manual_seed(0)
        w = torch.randn(3 * D, 2 * D, device=dev, generator=g) / (2 * D) ** 0.5
        gam, bet = torch.ones(3 * D, device=dev), torch.zeros(3 * D, device=dev)
        gx = torch.randn(B * T, 3 * D, device=dev, generator=g)
        mask, h0 = torch.ones(B * T, device=dev), torch.zeros(B, D, device=dev)
        wc = w.to(torch.bfloat16)
        ops.compute_weight, keep = (lambda _w: wc), ops.compute_weight          # one bf16 copy instead of a cast per call
        try:
            for _ in range(2):
                ops._GruScanStream.apply(gx, w, gam, bet, mask, h0, B, T, D)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops._GruScanStream.apply(gx, w, gam, bet, mask, h0, B, T, D)
            e1.record()
            torch.cuda.synchronize()
        finally:
            ops.compute_weight = keep
    us = e0.elapsed_time(e1) * 1e3 / 5
    wbytes = 3 * D * D * 2
    kbytes = T * wbytes + B * T * (3 * D * 4 + D * 4 + D * 2 + 3 * D * 2 + 8)
    return {"kernel": "rssm_scan_stream_fwd_kernel (W_h streamed once per time step into mma.sync A fragments, 128 CTAs)", "B": B, "T": T, "D": D,
            "us_per_launch": us, "us_per_time_step": us / T, "weight_bytes_per_time_step": wbytes, "kernel_bytes_per_launch": kbytes,
            "achieved": kbytes / (us * 1e-6) / 1e9, "unit": "GB/s", "peak": pk["hbm"], "frac": kbytes / (us * 1e-6) / 1e9 / pk["hbm"],
            "note": "up to 75 MB of the 101 MB of weights stay in L2 between time steps (evict_last), the rest is re-read from HBM"}


def time_observe(agent, data):
    """device time of one EnsembleRSSM.observe forward at the benchmark shape (no grad), best of 5"""
    from genrl_b200 import ops
    best = 1e30
    with torch.no_grad(), agent._ctx():
        embed = agent.wm.encoder(data)
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            agent.wm.rssm.observe(embed, data["action"], data["is_first"])
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


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
    apply_preset(a)
    if a.preset == "c4":
        EXTRA_KEYS["state"] = 30
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
