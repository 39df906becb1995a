"""CPU: the oracle restatement reproduces the reference's golden vectors (tests/golden/, made by
oracle/make_golden.py from the unmodified reference: outputs, metrics, gradients, parameters after two updates)."""
import numpy as np
import pytest
import torch

from oracle import genrl_oracle as O
from tests.golden_utils import load_case


@pytest.mark.parametrize("case", ["tiny_genrl", "tiny_dreamer", "tiny_dreamer_v3"])
def test_oracle_matches_reference_golden(case):
    torch.set_num_threads(1)
    g, meta, cfg = load_case(case)
    agent = O.OracleAgent(g["init"], cfg, kind=meta["kind"])
    agent.tape = O.NoiseTape(replay=g["tape"])          # the recorded draws of the reference run
    grads = []
    for step in range(2):
        state, outs, mets = agent.update(g["data"])
        ref = g[f"ref{step}"]
        for grp in ("post", "prior"):
            # bit-exact categorical sample indices
            assert torch.equal(outs[grp]["stoch"], ref[f"outputs/{grp}/stoch"]), (step, grp)
            for k in ("deter", "logit"):
                torch.testing.assert_close(outs[grp][k], ref[f"outputs/{grp}/{k}"], rtol=2e-5, atol=1e-6)
        for k in ("embed", "feat", "kl"):
            torch.testing.assert_close(outs[k], ref[f"outputs/{k}"], rtol=2e-5, atol=2e-6)
        for k in ("observation", "reward"):
            torch.testing.assert_close(outs["likes"][k], ref[f"outputs/likes/{k}"], rtol=2e-5, atol=1e-5)
        for k, v in meta["metrics"][step].items():
            assert k in mets, k
            np.testing.assert_allclose(float(mets[k]), v, rtol=1e-4, atol=1e-6, err_msg=k)
        gl = [agent.last_wm_grads]
        if meta["kind"] == "genrl":
            gl.append(agent.last_connector_grads)
        grads += gl + [agent.last_actor_grads, agent.last_critic_grads]
    assert agent.tape.pos == len(agent.tape.items)      # every recorded draw consumed, in order
    for i, og in enumerate(grads):
        rg = g["grads"][i]
        assert set(rg) == set(og)
        for k in rg:
            s = float(rg[k].abs().max()) + 1e-12
            torch.testing.assert_close(og[k] / s, rg[k] / s, rtol=0, atol=1e-4, msg=lambda m: f"grad {i} {k}: {m}")
    for k, v in g["ref1"].items():
        if k.startswith("params/"):
            torch.testing.assert_close(agent.P[k[len("params/"):]], v, rtol=1e-5, atol=2e-6, msg=lambda m: f"{k}: {m}")


def test_multinomial_identity():
    """torch.multinomial(p,1,True) == argmax(p / Exp(1)) under the same seed."""
    p = torch.softmax(torch.randn(64, 32), -1)
    torch.manual_seed(5)
    a = torch.multinomial(p, 1, True).squeeze(-1)
    torch.manual_seed(5)
    q = torch.empty_like(p).exponential_(1)
    assert torch.equal(a, torch.argmax(p / q, -1))


def _load_vlm_case():
    import json, os
    from tests.golden_utils import GOLDEN
    from genrl_b200.config import genrl_config
    z = np.load(os.path.join(GOLDEN, "tiny_genrl_vlm.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "tiny_genrl_vlm.json")))
    g = {"init": {}, "data": {}, "target": {}, "reward": {}, "params": {}, "tape": [], "target_tape": []}
    for k in z.files:
        arr = torch.from_numpy(z[k])
        if k == "prompt_embedding":
            g["prompt"] = arr
            continue
        head, rest = k.split("/", 1)
        if head in ("tape", "target_tape"):
            g[head].append((rest, arr))
        elif head == "ref1":
            g["params"][rest[len("params/"):]] = arr
        elif head == "reward":
            g["reward"][int(rest)] = arr
        else:
            g[head][rest] = arr
    for t in ("tape", "target_tape"):
        g[t] = [(name.split("_", 1)[1], v) for name, v in sorted(g[t], key=lambda kv: kv[0])]
    cfg = genrl_config(device="cpu", precision=32, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"], agent="genrl",
                       **meta["overrides"])
    return g, meta, cfg


def test_oracle_vlm_reward_imag_behavior_matches_reference_golden(monkeypatch):
    """video_text_reward (max_cosine on the decoder projection, sliding-window sequence alignment against the cached
    connector target) + update_imag_behavior of the UNMODIFIED reference, reproduced by the oracle: target, rewards, metrics, parameters."""
    torch.set_num_threads(1)
    g, meta, cfg = _load_vlm_case()
    agent = O.OracleAgent(g["init"], cfg, kind="genrl")
    args = cfg["imag_reward_args"]
    N, steps = meta["B"] * meta["T"], meta["H"] + 1
    target = O.unconditional_target(agent.P, g["prompt"], N, steps, cfg, O.NoiseTape(replay=g["target_tape"]),
                                    skip_first=args["skip_first_target"], sample=args["sample_for_target"])
    assert torch.equal(target["stoch"], g["target"]["stoch"])                 # one-hot target latents bit-exact
    for k in ("deter", "logit"):
        torch.testing.assert_close(target[k], g["target"][k], rtol=2e-5, atol=2e-6)
    rewards = []
    real = O.vlm_reward
    monkeypatch.setattr(O, "vlm_reward", lambda *a, **k: rewards.append(real(*a, **k)) or rewards[-1])
    agent.tape = O.NoiseTape(replay=g["tape"])
    for step in range(2):
        state, outs, mets = agent.update_wm(g["data"])
        mets.update(agent.update_imag_behavior(outs, target))
        torch.testing.assert_close(rewards[step].detach(), g["reward"][step], rtol=2e-5, atol=2e-6)
        assert rewards[step].shape == (steps, N, 1)
        for k, v in meta["metrics"][step].items():
            assert k in mets, k
            np.testing.assert_allclose(float(mets[k]), v, rtol=1e-4, atol=1e-6, err_msg=k)
    assert agent.tape.pos == len(agent.tape.items)
    assert any(k.startswith("imag_") for k in meta["metrics"][0])
    for k, v in g["params"].items():
        torch.testing.assert_close(agent.P[k], v, rtol=1e-5, atol=2e-6, msg=lambda m: f"{k}: {m}")


def _load_report_case():
    import json, os
    from tests.golden_utils import GOLDEN
    from genrl_b200.config import genrl_config
    z = np.load(os.path.join(GOLDEN, "tiny_genrl_report.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "tiny_genrl_report.json")))
    shared = np.load(os.path.join(GOLDEN, meta["init_from"]))
    g = {"init": {k[len("init/"):]: torch.from_numpy(shared[k]) for k in shared.files if k.startswith("init/")}, "data": {}, "report": {}, "tape": []}
    for k in z.files:
        head, rest = k.split("/", 1)
        arr = torch.from_numpy(z[k])
        if head == "tape":
            g["tape"].append((rest, arr))
        else:
            g[head][rest] = arr
    g["tape"] = [(name.split("_", 1)[1], v) for name, v in sorted(g["tape"], key=lambda kv: kv[0])]
    cfg = genrl_config(device="cpu", precision=32, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"], agent="genrl",
                       **meta["overrides"])
    return g, meta, cfg


def test_oracle_report_videos_match_reference_golden():
    """agent.report() of the UNMODIFIED reference — `openl_observation` (WorldModel.video_pred, dreamer.py:307-321) and
    `video_clip_pred` (GenRLAgent.report, genrl.py:64-106) — reproduced by the oracle on the reference's recorded draws."""
    torch.set_num_threads(1)
    g, meta, cfg = _load_report_case()
    agent = O.OracleAgent(g["init"], cfg, kind="genrl")
    agent.tape = O.NoiseTape(replay=g["tape"])
    data = agent.cast(g["data"])
    with torch.no_grad():
        vids = {"openl_observation": O.video_pred(agent.P, data, cfg, agent.tape), "video_clip_pred": O.video_clip_pred(agent.P, data, cfg, agent.tape)}
    assert agent.tape.pos == len(agent.tape.items)
    for k, v in vids.items():
        assert v.shape == (meta["B"], meta["T"], 3, 192, 64)
        model = v[:, :, :, 64:128]
        torch.testing.assert_close(model, g["report"][k].float(), rtol=0, atol=1e-3)          # the fixture stores the model third in fp16
        truth = g["data"]["observation"].float() / 255.0
        torch.testing.assert_close(v[:, :, :, :64], truth, rtol=0, atol=1e-6)
        torch.testing.assert_close(v[:, :, :, 128:], (model - truth + 1) / 2, rtol=0, atol=1e-6)


def _load_act_case():
    import json, os
    from tests.golden_utils import GOLDEN
    from genrl_b200.config import genrl_config
    z = np.load(os.path.join(GOLDEN, "tiny_genrl_act.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "tiny_genrl_act.json")))
    shared = np.load(os.path.join(GOLDEN, meta["init_from"]))
    g = {"init": {k[len("init/"):]: torch.from_numpy(shared[k]) for k in shared.files if k.startswith("init/")},
         "observation": z["observation"], "actions": z["actions"],
         "tape": [(k.split("/", 1)[1].split("_", 1)[1], torch.from_numpy(z[k])) for k in sorted(f for f in z.files if f.startswith("tape/"))]}
    cfg = genrl_config(device="cpu", precision=32, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"], agent="genrl",
                       **meta["overrides"])
    return g, meta, cfg


def act_observation(obs_u8, is_first):
    return dict(observation=obs_u8, reward=np.zeros((), np.float32), is_first=np.array(bool(is_first)), is_last=np.array(False),
                is_terminal=np.array(False), discount=np.ones((), np.float32))


def test_oracle_act_matches_reference_golden():
    """DreamerAgent.act (dreamer.py:41-64) of the UNMODIFIED reference over a 6-step episode (sampled and mean actions, one
    reset, carried state) reproduced by the oracle on the recorded draws."""
    torch.set_num_threads(1)
    g, meta, cfg = _load_act_case()
    agent = O.OracleAgent(g["init"], cfg, kind="genrl")
    agent.tape = O.NoiseTape(replay=g["tape"])
    state = None
    with torch.no_grad():
        for i, (first, ev) in enumerate(meta["steps"]):
            a, state = O.act_step(agent.P, act_observation(g["observation"][i], first), state, cfg, agent.tape, ev, meta["A"])
            np.testing.assert_allclose(a.numpy(), g["actions"][i], rtol=0, atol=2e-6, err_msg=f"step {i}")
    assert agent.tape.pos == len(agent.tape.items)


def seeded_init_for(meta):
    """initial parameters of a slim fixture: the mirror built on CPU under the recorded seed (module creation order = the reference's),
    verified against the sha256 of the reference's initial state_dict stored in the fixture's json"""
    import hashlib
    from genrl_b200.agent.genrl import GenRLAgent
    from genrl_b200.config import genrl_config, make_specs, Spec
    cfg = genrl_config(device="cpu", precision=32, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"], agent=meta["kind"],
                       **meta["overrides"])
    obs_space, act_spec = make_specs(action_dim=meta["A"], clip_dim=meta["clip"])
    for k, d in meta.get("extra_obs", {}).items():
        obs_space[k] = Spec((d,), "float32")
    torch.manual_seed(meta["seed_init"])
    init = {k: v.detach().clone() for k, v in GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec).state_dict().items()}
    sha = hashlib.sha256()
    for k in sorted(init):
        sha.update(k.encode()); sha.update(init[k].numpy().tobytes())
    assert sha.hexdigest() == meta["init_sha256"], "seeded initial parameters differ from the reference's"
    return init


def test_oracle_matches_reference_golden_kitchen_shaped_single_sequence():
    """BASELINE configs[3] flavour at the ragged edge: A = 9, an extra `state` observation key the model ignores, B = 1 sequence with resets.
    Slim fixture (no parameter dumps): outputs of both updates, bit-exact sample indices and every metric (losses, grad norms) are compared."""
    torch.set_num_threads(1)
    g, meta, cfg = load_case("tiny_kitchen")
    assert not g["init"] and meta["B"] == 1 and "state" in g["data"]
    agent = O.OracleAgent(seeded_init_for(meta), cfg, kind=meta["kind"])
    agent.tape = O.NoiseTape(replay=g["tape"])
    for step in range(2):
        state, outs, mets = agent.update(g["data"])
        ref = g[f"ref{step}"]
        for grp in ("post", "prior"):
            assert torch.equal(outs[grp]["stoch"], ref[f"outputs/{grp}/stoch"]), (step, grp)
            for k in ("deter", "logit"):
                torch.testing.assert_close(outs[grp][k], ref[f"outputs/{grp}/{k}"], rtol=2e-5, atol=1e-6)
        for k in ("embed", "feat", "kl"):
            torch.testing.assert_close(outs[k], ref[f"outputs/{k}"], rtol=2e-5, atol=2e-6)
        for k, v in meta["metrics"][step].items():
            np.testing.assert_allclose(float(mets[k]), v, rtol=1e-4, atol=1e-6, err_msg=k)
    assert agent.tape.pos == len(agent.tape.items)


def test_sharded_gradients_average_to_the_global_batch_gradient():
    """SURVEY §8e (why data parallelism needs only a gradient all-reduce): every loss on the path is a mean of per-row terms, so the
    average of the shard gradients — each shard seeing its rows of the batch AND its rows of every random draw — equals the gradient of
    the global batch.  Checked with the oracle for the world-model and the connector gradients of one update, 2 shards x 2 sequences."""
    torch.set_num_threads(1)
    g, meta, cfg = load_case("tiny_genrl")
    B, world = 4, 2
    data = O.synthetic_batch(B, 16, meta["A"], clip_dim=meta["clip"], seed=21, resets=0.1)
    glob = O.OracleAgent(g["init"], cfg, kind="genrl")
    torch.manual_seed(3)
    glob.update_wm(data)
    ref = {"wm": glob.last_wm_grads, "connector": glob.last_connector_grads}
    acc = {"wm": {}, "connector": {}}
    per = B // world
    for r in range(world):
        items = []
        for kind, v in glob.tape.items:
            if kind == "randint":
                items.append((kind, v))
            else:                                         # draws are batch-major in their leading dimension
                rows = v.shape[0] // B
                items.append((kind, v[r * per * rows:(r + 1) * per * rows]))
        shard = O.OracleAgent(g["init"], cfg, kind="genrl")
        shard.tape = O.NoiseTape(replay=items)
        shard.update_wm({k: v[r * per:(r + 1) * per] for k, v in data.items()})
        assert shard.tape.pos == len(items)
        for name, grads in (("wm", shard.last_wm_grads), ("connector", shard.last_connector_grads)):
            for k, v in grads.items():
                acc[name][k] = acc[name].get(k, 0) + v / world
    for name in ref:
        assert set(ref[name]) == set(acc[name])
        for k, v in ref[name].items():
            s = float(v.abs().max()) + 1e-12
            torch.testing.assert_close(acc[name][k] / s, v / s, rtol=0, atol=2e-5, msg=lambda m: f"{name} {k}: {m}")


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_replay_oracle_reproduces_the_reference_batches(ci):
    """oracle/replay_oracle.py (tools/replay.py:175-236 restated) against the sha256 of the batches the unmodified
    reference ReplayBuffer yielded for the same episodes and numpy seeds (oracle/make_golden_replay.py)"""
    import json, os
    from tests.golden_utils import GOLDEN
    from oracle.replay_oracle import ReplayOracle
    from oracle.make_golden_replay import script, batch_digest, L, B
    case = json.load(open(os.path.join(GOLDEN, "replay_small.json")))["cases"][ci]
    ora = ReplayOracle(L, case["capacity"], case["min_t"])
    n = 0
    for op, arg in script(case):
        if op == "store":
            ora.store_episode(arg)
        else:
            np.random.seed(arg)
            assert batch_digest(ora.sample(B)) == case["digests"][n], n
            n += 1
    assert n == len(case["digests"]) and n > 20
