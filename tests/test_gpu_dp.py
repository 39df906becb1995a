"""Data-parallel equivalence on the CUDA path (SURVEY §4 iii / §8e): two ranks on halves of a batch == one rank on the whole batch.

Two processes are spawned; with >= 2 GPUs each takes its own device and talks NCCL, on a single-GPU box both share cuda:0 and talk
gloo (the kernels, the flat-gradient all-reduce, the early side-stream reduce and the RewardEMA all-gather are the same code).  Every
rank first runs the WHOLE batch alone while recording every random draw, then replays its rows of those draws on its half of the batch
with the DataParallel hooks attached, and compares: gradient norms (norm of the averaged gradients = norm of the global-batch
gradient), parameters after two full updates, the RewardEMA buffer, and the losses averaged over the ranks.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class RecordedNoise:
    """records device draws (world = 1) or replays this rank's rows of a recorded tape (rows are batch-major everywhere)"""

    def __init__(self, items=None, rank=0, world=1):
        self.items, self.replay = (items, True) if items is not None else ([], False)
        self.rank, self.world, self.pos = rank, world, 0

    def _take(self, shape, device, kind):
        if not self.replay:
            t = torch.empty(shape, device=device, dtype=torch.float32)
            t = t.exponential_() if kind == "exp" else t.normal_()
            self.items.append((kind, t))
            return t
        k, full = self.items[self.pos]
        self.pos += 1
        assert k == kind and full.shape[0] == shape[0] * self.world and tuple(full.shape[1:]) == tuple(shape[1:]), (k, kind, full.shape, shape)
        return full[self.rank * shape[0]:(self.rank + 1) * shape[0]].contiguous()

    def exp(self, rows, K, device):
        return self._take((rows, K), device, "exp"), 0, 0

    def randn(self, shape, device):
        return self._take(tuple(shape), device, "randn")

    def randn_into(self, out):
        out.copy_(self._take(tuple(out.shape), out.device, "randn"))
        return out

    def randint(self):
        return None

    def tick(self, device):
        pass

    def observe(self, B, T, S, K, device):
        return self.exp(B * T * S, K, device), self.exp(B * T * S, K, device)


def _worker(rank, world, port, overlap, out):
    try:
        _worker_body(rank, world, port, overlap, out)
    except Exception:
        import traceback
        out.put((rank, {"bad": [("exception", traceback.format_exc())], "n_allreduce": -1}))
        raise


def _worker_body(rank, world, port, overlap, out):
    import torch.distributed as dist
    from genrl_b200 import ops
    from genrl_b200.agent.genrl import GenRLAgent
    from genrl_b200.config import genrl_config, tiny_overrides, make_specs
    from genrl_b200.dist import DataParallel
    from genrl_b200.synthetic import synthetic_batch
    ngpu = torch.cuda.device_count()
    dev = f"cuda:{rank}" if ngpu >= world else "cuda:0"
    torch.cuda.set_device(dev)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    backend = "nccl" if ngpu >= world else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world, **({"device_id": torch.device(dev)} if backend == "nccl" else {}))
    B, T, A, H = 4, 16, 6, 5

    def make():
        cfg = genrl_config(device=dev, precision=32, batch_size=B, batch_length=T, imag_horizon=H, **tiny_overrides())
        cfg.overlap_streams = overlap
        obs_space, act_spec = make_specs(action_dim=A, clip_dim=32)
        torch.manual_seed(1)
        return GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)

    batches = [{k: v.to(dev) for k, v in synthetic_batch(B, T, A, clip_dim=32, seed=40 + i, resets=0.1).items()} for i in range(2)]
    scal = lambda m: {k: float(torch.as_tensor(v).float().cpu().reshape(-1)[0]) for k, v in m.items()}
    # ---- whole batch on one rank, recording the draws
    torch.manual_seed(123)
    torch.cuda.manual_seed(123)
    rec = ops.set_noise(RecordedNoise())
    whole = make()
    hist_w = []
    for i, data in enumerate(batches):
        state, outputs, metrics = whole.update_wm(data, i)
        _, metrics = whole.update_acting_behavior(state, outputs, metrics, data)
        hist_w.append(scal(metrics))
    # ---- this rank's half, DataParallel attached, replaying its rows of the same draws
    ops.set_noise(RecordedNoise(rec.items, rank, world))
    part = make()
    dp = DataParallel().attach(part)
    hist_p = []
    for i, data in enumerate(batches):
        shard = part.dp.shard(data)
        state, outputs, metrics = part.update_wm(shard, i)
        _, metrics = part.update_acting_behavior(state, outputs, metrics, shard)
        hist_p.append(scal(metrics))
    assert ops.noise.pos == len(rec.items)
    torch.cuda.synchronize()
    res = {"n_allreduce": part.dp.n_allreduce, "backend": backend, "bad": []}
    keys_mean = ("model_loss", "kl_loss", "observation_loss", "reward_loss", "connector_model_loss", "actor_loss", "critic_loss")
    keys_same = ("model_grad_norm", "connector_model_grad_norm", "actor_grad_norm", "critic_grad_norm", "reward_ema_005", "reward_ema_095")
    for i in range(2):
        for k in keys_same:
            a, b = hist_w[i][k], hist_p[i][k]
            if abs(a - b) > 2e-4 * max(1.0, abs(a)):
                res["bad"].append((i, k, a, b))
        t = torch.tensor([hist_p[i][k] for k in keys_mean], dtype=torch.float64, device=dev if backend == "nccl" else "cpu")
        dist.all_reduce(t)
        for k, v in zip(keys_mean, (t / world).tolist()):
            if abs(v - hist_w[i][k]) > 1e-4 * max(1.0, abs(hist_w[i][k])):
                res["bad"].append((i, k, hist_w[i][k], v))
    sw, sp = whole.state_dict(), part.state_dict()
    worst = 0.0
    for k in sw:
        if sw[k].is_floating_point():
            d = float((sw[k] - sp[k]).abs().max())
            worst = max(worst, d / (2e-6 + 1e-4 * float(sw[k].abs().max())))
    res["worst_param_err_over_tol"] = worst
    if worst > 1.0:
        res["bad"].append(("params", worst))
    if not torch.allclose(whole._acting_behavior.ema_vals, part._acting_behavior.ema_vals, rtol=1e-5, atol=1e-7):
        res["bad"].append(("ema_vals", whole._acting_behavior.ema_vals.tolist(), part._acting_behavior.ema_vals.tolist()))
    out.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True])
def test_two_ranks_on_half_batches_equal_one_rank_on_the_whole_batch(overlap):
    import torch.multiprocessing as mp
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, overlap, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    try:
        import queue, time
        deadline = time.time() + 300
        while len(results) < world and time.time() < deadline:
            try:
                rank, res = q.get(timeout=2)
                results[rank] = res
            except queue.Empty:
                if any(p.exitcode not in (None, 0) for p in procs):      # a worker died: do not wait for its result
                    break
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
    assert len(results) == world and all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for rank, res in results.items():
        assert not res["bad"], (rank, res)
        # WM, connector, actor, critic: four flat-buffer collectives per update round (two of them hidden beside the other stream's backward)
        assert res["n_allreduce"] == 2 * 4, res
