"""BASELINE configs[4] shape (deter = hidden = 4096, stoch 48x48, H = 32) at a small batch: does the op-level path run in bf16?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import numpy as np
import torch
from genrl_b200 import ops
from genrl_b200.config import genrl_config, make_specs
from genrl_b200.agent.genrl import GenRLAgent
from genrl_b200.synthetic import synthetic_batch
DEV = "cuda:0"
B, T, A, H = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 16, 10, 8
cfg = genrl_config(device=DEV, precision=16, batch_size=B, batch_length=T, imag_horizon=H,
                   rssm=dict(ensemble=1, hidden=4096, deter=4096, stoch=48, discrete=48, norm="layer", std_act="softplus", min_std=0.1, single_obs_posterior=True),
                   connector_rssm=dict(ensemble=1, hidden=4096, deter=4096, stoch=48, discrete=48, norm="layer", std_act="softplus", min_std=0.1,
                                       single_obs_posterior=False, learn_initial=True))
obs_space, act_spec = make_specs(action_dim=A, clip_dim=512)
torch.manual_seed(1)
agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
ops.set_noise(ops.PhiloxNoise(3))
for i in range(3):
    data = {k: v.to(DEV) for k, v in synthetic_batch(B, T, A, clip_dim=512, seed=i).items()}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    state, outputs, metrics = agent.update_wm(data, i)
    start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
    torch.cuda.synchronize()
    vals = {k: float(torch.as_tensor(metrics[k]).float().cpu().reshape(-1)[0]) for k in ("model_loss", "kl_loss", "connector_model_loss", "actor_loss", "critic_loss")}
    print(i, f"{(time.perf_counter() - t0) * 1e3:.1f} ms", vals, flush=True)
    assert all(np.isfinite(v) for v in vals.values())
print("c5-shaped update ok; peak memory", torch.cuda.max_memory_allocated() // 2**20, "MiB")
