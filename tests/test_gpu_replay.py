"""the on-GPU replay store yields the same batches as the reference's sampler (oracle/replay_oracle.py restates
tools/replay.py:175-236) for the same numpy seed — ragged episode lengths, eviction at capacity, min_t_sampling, every dtype of the data dict."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.replay_oracle import ReplayOracle


from oracle.make_golden_replay import script, batch_digest, L, B


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_device_replay_batches_equal_the_reference_sampler(ci):
    """same episodes, same numpy seeds as oracle/make_golden_replay.py ran through the unmodified reference: every batch byte-identical to
    the oracle's and its sha256 equal to the reference's committed digest"""
    import json, os
    from genrl_b200.replay import ReplayBuffer, make_replay_loader
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replay_small.json")))
    case = gold["cases"][ci]
    buf = ReplayBuffer(length=L, capacity=case["capacity"], device="cuda:0", min_t_sampling=case["min_t"], max_episode_steps=128)
    ora = ReplayOracle(L, case["capacity"], case["min_t"])
    it = iter(make_replay_loader(buf, B))
    n = 0
    for op, arg in script(case):
        if op == "store":
            buf.store_episode(episode=arg)
            ora.store_episode(arg)
            assert buf.stats["loaded_episodes"] == len(ora.lens) and buf.stats["loaded_steps"] == ora.loaded_steps
            continue
        np.random.seed(arg)
        mine = next(it)
        np.random.seed(arg)
        ref = ora.sample(B)
        assert set(mine) == set(ref)
        host = {k: v.cpu().numpy() for k, v in mine.items()}
        for k, v in ref.items():
            assert host[k].shape == v.shape and host[k].dtype == v.dtype and mine[k].device.type == "cuda", k
            assert np.array_equal(host[k], v), f"batch {n}, key {k}"
        assert batch_digest(host) == case["digests"][n], f"batch {n} differs from the reference's"
        n += 1
    assert n == len(case["digests"])
    assert mine["observation"].dtype == torch.uint8 and mine["is_first"].dtype == torch.bool and mine["action"].dtype == torch.float32


def test_device_replay_feeds_the_agent_update():
    """the yielded dict is what agent.update_wm consumes (train.py:272-275)"""
    from genrl_b200.replay import ReplayBuffer, make_replay_loader
    from genrl_b200 import ops
    from genrl_b200.agent.genrl import GenRLAgent
    from genrl_b200.config import genrl_config, tiny_overrides, make_specs
    from genrl_b200.synthetic import synthetic_batch
    B, T, A = 3, 16, 6
    buf = ReplayBuffer(length=T, capacity=5000, device="cuda:0", max_episode_steps=64)
    for s in range(4):
        ep = {k: v[0].numpy() for k, v in synthetic_batch(1, 40, A, clip_dim=32, seed=s).items()}
        buf.add_episode(ep)
    it = iter(make_replay_loader(buf, B))
    cfg = genrl_config(device="cuda:0", precision=32, batch_size=B, batch_length=T, imag_horizon=4, **tiny_overrides())
    obs_space, act_spec = make_specs(action_dim=A, clip_dim=32)
    torch.manual_seed(1)
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    ops.set_noise(ops.PhiloxNoise(2))
    np.random.seed(0)
    for i in range(2):
        data = next(it)
        assert data["observation"].shape == (B, T, 3, 64, 64)
        state, metrics = agent.update(data, i)
        assert np.isfinite(float(metrics["model_loss"]))
    ops.set_noise(ops.PhiloxNoise(0))
