"""GPU parity of the drop-in agent API (GenRLAgent / DreamerAgent mirrors) against the reference's golden vectors.

The fixtures (tests/golden/*.npz) hold the UNMODIFIED reference's outputs, metrics, gradients and parameters for two
consecutive `update_wm` + `update_acting_behavior` calls; the CUDA path is fed the reference's own random draws
(TapeNoise) so categorical sample indices must come out bit-exact and losses within 1e-4 (fp32 mode).
"""
import copy
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from genrl_b200.config import genrl_config, make_specs
from tests.golden_utils import load_case

DEV = "cuda:0"


def build_agent(case, precision=32):
    from genrl_b200.agent.genrl import GenRLAgent
    from genrl_b200.agent.dreamer import DreamerAgent
    g, meta, _ = load_case(case)
    cfg = genrl_config(device=DEV, precision=precision, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"],
                       agent=meta["kind"], **meta["overrides"])
    obs_space, act_spec = make_specs(action_dim=meta["A"], clip_dim=meta["clip"])
    torch.manual_seed(1)
    cls = GenRLAgent if meta["kind"] == "genrl" else DreamerAgent
    agent = cls(name=meta["kind"], cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    missing = agent.load_state_dict({k: v for k, v in g["init"].items()}, strict=True)
    data = {k: v.to(DEV) for k, v in g["data"].items()}
    return agent, data, g, meta


def relclose(a, b, rtol, atol, msg):
    a = torch.as_tensor(a).detach().float().cpu(); b = torch.as_tensor(b).detach().float().cpu()
    scale = float(b.abs().max()) + 1e-12
    err = float((a - b).abs().max())
    assert err <= atol + rtol * scale, f"{msg}: max abs err {err:.3e} (scale {scale:.3e})"


@pytest.mark.parametrize("case", ["tiny_genrl", "tiny_dreamer", "tiny_dreamer_v3"])
def test_agent_update_matches_reference_golden_fp32(case):
    from genrl_b200 import ops
    agent, data, g, meta = build_agent(case, 32)
    # same initial weights as the reference from the seed alone (module creation order mirrors the reference)
    ops.set_noise(ops.TapeNoise(g["tape"], DEV))
    flips = 0
    for step in range(2):
        state, outputs, metrics = agent.update_wm(data, step)
        start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
        ref = g[f"ref{step}"]
        for grp in ("post", "prior"):
            mine = outputs[grp]["stoch"].float().cpu()
            flips += int((mine != ref[f"outputs/{grp}/stoch"]).any(-1).sum())
            for k in ("deter", "logit"):
                relclose(outputs[grp][k], ref[f"outputs/{grp}/{k}"], 1e-4, 2e-5, f"step {step} {grp}.{k}")
        assert flips == 0, f"{flips} categorical sample indices differ from the reference"
        relclose(outputs["embed"], ref["outputs/embed"], 1e-4, 2e-5, "embed")
        relclose(outputs["feat"], ref["outputs/feat"], 1e-4, 2e-5, "feat")
        relclose(outputs["kl"], ref["outputs/kl"], 1e-4, 2e-5, "kl value")
        for k in ("observation", "reward"):
            relclose(outputs["likes"][k], ref[f"outputs/likes/{k}"], 1e-4, 1e-4, f"likes.{k}")
        for k, v in meta["metrics"][step].items():
            assert k in metrics, f"metric {k} missing"
            mv = float(torch.as_tensor(metrics[k]).float().cpu().reshape(-1)[0])
            tol = 2e-3 if k.endswith("grad_norm") else 2e-4
            assert abs(mv - v) <= tol * max(1.0, abs(v)) + 1e-5, f"step {step} metric {k}: {mv} vs reference {v}"
    assert ops.noise.pos == len(ops.noise.items), "every reference draw must be consumed"
    sd = agent.state_dict()
    worst = 0.0
    for k, v in g["ref1"].items():
        if not k.startswith("params/"):
            continue
        name = k[len("params/"):]
        err = float((sd[name].float().cpu() - v).abs().max())
        worst = max(worst, err)
        assert err <= 5e-6 + 1e-4 * float(v.abs().max()), f"parameter {name} after 2 updates: err {err:.3e}"
    ops.set_noise(ops.PhiloxNoise(0))


def test_agent_bf16_mode_runs_and_tracks_fp32():
    from genrl_b200 import ops
    agent, data, g, meta = build_agent("tiny_genrl", 16)
    ops.set_noise(ops.TapeNoise(g["tape"], DEV))
    state, outputs, metrics = agent.update_wm(data, 0)
    ref = meta["metrics"][0]
    for k in ("kl_loss", "observation_loss", "reward_loss", "model_loss"):
        mv = float(metrics[k])
        assert abs(mv - ref[k]) <= 2e-2 * max(1.0, abs(ref[k])), f"bf16 {k}: {mv} vs {ref[k]}"
    ops.set_noise(ops.PhiloxNoise(0))
    start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
    for k, v in metrics.items():
        assert np.isfinite(float(torch.as_tensor(v).float().cpu().reshape(-1)[0])), k


def test_agent_fast_noise_pickle_and_act():
    from genrl_b200 import ops
    agent, data, g, meta = build_agent("tiny_genrl", 32)
    ops.set_noise(ops.PhiloxNoise(123))
    state, metrics = agent.update(data, 0)
    assert all(np.isfinite(float(torch.as_tensor(v).float().cpu().reshape(-1)[0])) for v in metrics.values())
    buf = io.BytesIO()
    torch.save({"agent": agent}, buf)                       # train.py:387-398 pickles the whole agent
    # the snapshot carries parameters + Adam moments once (no flat gradient buffers, weight shadows or duplicate stores)
    pbytes = sum(p.numel() * 4 for p in agent.parameters())
    assert buf.tell() <= 3.4 * pbytes + (1 << 20), f"snapshot {buf.tell()} bytes for {pbytes} parameter bytes"
    buf.seek(0)
    agent2 = torch.load(buf, weights_only=False)["agent"]
    assert agent2.wm.model_opt.store is None and not hasattr(agent2._acting_behavior, "dp_gather")
    for (k, a), (_, b) in zip(agent.state_dict().items(), agent2.state_dict().items()):
        assert torch.equal(a, b), k
    state2, metrics2 = agent2.update(data, 1)
    assert np.isfinite(float(metrics2["model_loss"]))
    obs = {"observation": np.zeros((3, 64, 64), np.uint8), "reward": np.zeros((), np.float32), "is_first": np.array(True),
           "is_last": np.array(False), "is_terminal": np.array(False), "discount": np.ones((), np.float32)}
    action, st = agent2.act(obs, None, 0, eval_mode=False, state=None)
    assert action.shape == (meta["A"],) and np.all(np.abs(action) <= 2.0)
    action, st = agent2.act(obs, None, 1, eval_mode=True, state=st)
    assert action.shape == (meta["A"],)


def test_cuda_graph_replay_matches_eager():
    """the graph-captured update (cfg.cuda_graph) reproduces the eager update step for step (same Philox keys)"""
    from genrl_b200 import ops
    results = []
    for graphed in (False, True):
        agent, data, g, meta = build_agent("tiny_genrl", 32)
        agent.cfg.cuda_graph = graphed
        ops.set_noise(ops.PhiloxNoise(77))
        hist = []
        for step in range(5):
            state, outputs, metrics = agent.update_wm(data, step)
            start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
            hist.append({k: float(torch.as_tensor(v).float().cpu().reshape(-1)[0]) for k, v in metrics.items()})
            stoch = outputs["post"]["stoch"].clone()
        results.append((hist, stoch, {k: v.clone() for k, v in agent.state_dict().items()}))
        if graphed:
            assert agent._graphs and all(gc.graph is not None for gc in agent._graphs.values()), "graphs were not captured"
    (h0, s0, p0), (h1, s1, p1) = results
    assert torch.equal(s0, s1), "sampled latents differ between eager and graph replay"
    for a, b in zip(h0, h1):
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-3 * max(1.0, abs(a[k])), (k, a[k], b[k])
    assert h0[1]["kl_loss"] != h0[2]["kl_loss"]          # the step counter gives fresh noise each step
    for k in p0:
        assert torch.allclose(p0[k].float(), p1[k].float(), rtol=1e-3, atol=1e-5), k
    ops.set_noise(ops.PhiloxNoise(0))


def test_overlapped_streams_match_single_stream():
    """cfg.overlap_streams (connector beside WM backward, critic beside actor backward) gives the same update"""
    from genrl_b200 import ops
    import gc
    results = []
    for overlap in (False, True):
        agent = None
        gc.collect(); torch.cuda.synchronize()
        agent, data, g, meta = build_agent("tiny_genrl", 32)
        agent.cfg.overlap_streams = overlap
        agent.cfg.cuda_graph = overlap          # exercise capture of the forked streams too
        ops.set_noise(ops.PhiloxNoise(5))
        hist = []
        for step in range(4):
            state, outputs, metrics = agent.update_wm(data, step)
            start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
            hist.append({k: float(torch.as_tensor(v).float().cpu().reshape(-1)[0]) for k, v in metrics.items()})
        results.append(hist)
    # noise call order differs between the two schedules, so compare the noise-free first-step quantities and finiteness
    a, b = results
    for k in ("observation_loss", "reward_loss", "kl_loss", "model_loss", "model_grad_norm"):
        assert abs(a[0][k] - b[0][k]) <= 1e-4 * max(1.0, abs(a[0][k])), (k, a[0][k], b[0][k])
    for h in b:
        assert all(np.isfinite(v) for v in h.values())
    assert b[3]["model_loss"] < b[0]["model_loss"]
    ops.set_noise(ops.PhiloxNoise(0))


def test_overlapped_connector_step_has_its_own_norm_and_matches_single_stream():
    """with overlap_streams the connector's gradients are already in their flat buffer when the world-model step
    runs; they must not enter the world-model norm / clip / Adam pass.  One update_wm: connector_model_grad_norm > 0 and equal to the
    single-stream schedule's, model_grad_norm equal, connector weights equal (the only difference left is the 1e-6 decay order)."""
    from genrl_b200 import ops
    res = []
    for overlap in (False, True):
        agent, data, g, meta = build_agent("tiny_genrl", 32)
        agent.cfg.overlap_streams = overlap
        ops.set_noise(ops.PhiloxNoise(9))
        state, outputs, metrics = agent.update_wm(data, 0)
        res.append(({k: float(torch.as_tensor(v).float().cpu().reshape(-1)[0]) for k, v in metrics.items()},
                    {k: v.clone() for k, v in agent.wm.connector.state_dict().items()}))
    (m0, c0), (m1, c1) = res
    assert m1["connector_model_grad_norm"] > 0
    for k in ("connector_model_grad_norm", "model_grad_norm", "connector_model_loss", "model_loss"):
        assert abs(m0[k] - m1[k]) <= 1e-4 * max(1.0, abs(m0[k])), (k, m0[k], m1[k])
    for k in c0:
        assert torch.allclose(c0[k], c1[k], rtol=1e-4, atol=2e-6), k
    ops.set_noise(ops.PhiloxNoise(0))


def test_finetune_mode_keeps_the_imag_behaviours_adam_state():
    """after finetune_mode() the acting behaviour IS the imagination behaviour; its flat group must inherit that
    behaviour's Adam moments (the reference's per-module torch Adam follows the module), not the old acting behaviour's."""
    from genrl_b200 import ops
    from tests.test_oracle_golden import _load_vlm_case
    from genrl_b200.agent.genrl import GenRLAgent
    g, meta, _ = _load_vlm_case()
    cfg = genrl_config(device=DEV, precision=32, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"], agent="genrl",
                       **meta["overrides"])
    obs_space, act_spec = make_specs(action_dim=meta["A"], clip_dim=meta["clip"])
    torch.manual_seed(1)
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    agent.load_state_dict(g["init"], strict=True)
    agent.unconditional_target = {k: v.to(DEV) for k, v in g["target"].items()}
    data = {k: v.to(DEV) for k, v in g["data"].items()}
    ops.set_noise(ops.PhiloxNoise(3))
    state, outputs, metrics = agent.update_wm(data, 0)
    agent.update_imag_behavior(state, outputs, metrics, seq_data=data)
    agent.update_acting_behavior(state, outputs, metrics, data)
    st = agent.ensure_store()
    m_imag = st.groups["_imag_behavior.actor"].m.clone()
    m_act = st.groups["_acting_behavior.actor"].m.clone()
    assert float(m_imag.abs().sum()) > 0 and not torch.equal(m_imag, m_act)
    # through a snapshot as well: the state saved under the group name must follow the module, not the name
    buf = io.BytesIO(); torch.save({"agent": agent}, buf); buf.seek(0)
    for ag in (agent, torch.load(buf, weights_only=False)["agent"]):
        ag.finetune_mode()
        st2 = ag.ensure_store()
        assert "_imag_behavior.actor" not in st2.groups
        assert torch.equal(st2.groups["_acting_behavior.actor"].m, m_imag)
        assert float(st2.groups["_acting_behavior.actor"].state[0]) == 1.0
    ops.set_noise(ops.PhiloxNoise(0))


def test_direct_wm_calls_of_train_py_work_without_a_prior_update():
    """train.py:277-280 (train_world_model: False, train_connector: True): wm.observe_data + wm.update_additional_detached_modules are
    the first calls on a fresh agent; they must build the flat store themselves and step the connector only."""
    from genrl_b200 import ops
    agent, data, g, meta = build_agent("tiny_genrl", 32)
    ops.set_noise(ops.PhiloxNoise(4))
    before = {k: v.clone() for k, v in agent.state_dict().items()}
    with torch.no_grad():
        outputs, metrics = agent.wm.observe_data(data)
    assert set(outputs) >= {"embed", "post", "prior", "is_terminal"} and np.isfinite(float(metrics["model_kl"]))
    _, metrics = agent.wm.update_additional_detached_modules(data, outputs, metrics)
    assert float(metrics["connector_model_grad_norm"]) > 0
    after = agent.state_dict()
    changed = {k for k in before if not torch.equal(before[k], after[k])}
    assert any(k.startswith("wm.connector.") for k in changed)
    assert all(k.startswith("wm.connector.") for k in changed), sorted(changed)[:5]
    ops.set_noise(ops.PhiloxNoise(0))


def test_kitchen_shaped_agent_full_conv_stack_bf16_tracks_fp32():
    """BASELINE configs[3]: kitchen has A = 9 and an extra low-dimensional `state` observation key that the GenRL encoder ignores
    (conf/env/dmc_pixels.yaml: mlp_keys '$^').  Full-size conv stacks (depth 48, 64x64), small batch: the bf16 tensor-core path
    (implicit-GEMM convolutions) must track the fp32 parity path of the same weights on the same draws."""
    from genrl_b200 import ops
    from genrl_b200.agent.genrl import GenRLAgent
    from genrl_b200.config import Spec
    from genrl_b200.synthetic import synthetic_batch
    B, T, A, H = 3, 16, 9, 4
    results = {}
    for precision in (32, 16):
        cfg = genrl_config(device=DEV, precision=precision, batch_size=B, batch_length=T, imag_horizon=H)
        obs_space, act_spec = make_specs(action_dim=A, clip_dim=512)
        obs_space["state"] = Spec((30,), "float32")
        torch.manual_seed(1)
        agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
        data = {k: v.to(DEV) for k, v in synthetic_batch(B, T, A, clip_dim=512, seed=3, resets=0.1).items()}
        data["state"] = torch.zeros(B, T, 30, device=DEV)
        ops.set_noise(ops.PhiloxNoise(7))
        state, outputs, metrics = agent.update_wm(data, 0)
        start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
        assert outputs["embed"].shape == (B, T, 1536) and outputs["post"]["stoch"].shape == (B, T, 32, 32)
        results[precision] = {k: float(torch.as_tensor(v).float().cpu().reshape(-1)[0]) for k, v in metrics.items()}
        assert all(np.isfinite(v) for v in results[precision].values())
    for k in ("observation_loss", "reward_loss", "kl_loss", "model_loss", "connector_model_loss"):
        a, b = results[16][k], results[32][k]
        assert abs(a - b) <= 2e-2 * max(1.0, abs(b)), f"{k}: bf16 {a} vs fp32 {b}"


def test_ragged_shapes_resets_and_graph_replay_bf16():
    """odd batch / action sizes, frequent is_first resets, short horizon; eager and CUDA-graph replays give the same losses"""
    from genrl_b200 import ops
    from genrl_b200.agent.genrl import GenRLAgent
    from genrl_b200.config import tiny_overrides
    from genrl_b200.synthetic import synthetic_batch
    B, T, A, H = 5, 24, 7, 3
    losses = {}
    for graph in (False, True):
        cfg = genrl_config(device=DEV, precision=16, batch_size=B, batch_length=T, imag_horizon=H, **tiny_overrides())
        cfg.cuda_graph = graph
        obs_space, act_spec = make_specs(action_dim=A, clip_dim=32)
        torch.manual_seed(1)
        agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
        ops.set_noise(ops.PhiloxNoise(11))
        out = []
        for i in range(4):
            data = {k: v.to(DEV) for k, v in synthetic_batch(B, T, A, clip_dim=32, seed=20 + i, resets=0.3).items()}
            state, outputs, metrics = agent.update_wm(data, i)
            start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
            out.append([float(metrics[k]) for k in ("model_loss", "connector_model_loss", "actor_loss", "critic_loss")])
            assert all(np.isfinite(v) for v in out[-1]), (graph, i, out[-1])
            assert outputs["post"]["stoch"].shape == (B, T) + tuple(outputs["post"]["stoch"].shape[2:])
        losses[graph] = out
    # calls 0-1 are eager in both runs; the captured replays (calls 2-3) must reproduce the eager numbers of the same step
    for a, b in zip(losses[False], losses[True]):
        for x, y in zip(a, b):
            assert abs(x - y) <= 2e-2 * max(1.0, abs(x)), (losses[False], losses[True])


def test_large_rssm_config_runs_on_the_op_level_path():
    """BASELINE configs[4] shape: deter = hidden = 4096, stoch 48x48 (K = 48 takes the generic categorical kernels, D = 4096 the
    op-level recurrence because the resident-weight scan holds at most 8 x #SMs hidden units); small batch, bf16."""
    from genrl_b200 import ops
    from genrl_b200.agent.genrl import GenRLAgent
    from genrl_b200.synthetic import synthetic_batch
    B, T, A, H = 2, 16, 10, 3
    big = dict(ensemble=1, hidden=4096, deter=4096, stoch=48, discrete=48, norm="layer", std_act="softplus", min_std=0.1)
    cfg = genrl_config(device=DEV, precision=16, batch_size=B, batch_length=T, imag_horizon=H,
                       rssm=dict(big, single_obs_posterior=True), connector_rssm=dict(big, single_obs_posterior=False, learn_initial=True))
    obs_space, act_spec = make_specs(action_dim=A, clip_dim=512)
    torch.manual_seed(1)
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    assert not ops.scan_supported(4096) and ops.scan_supported(1024)
    ops.set_noise(ops.PhiloxNoise(3))
    first = None
    for i in range(3):
        data = {k: v.to(DEV) for k, v in synthetic_batch(B, T, A, clip_dim=512, seed=5).items()}
        state, outputs, metrics = agent.update_wm(data, i)
        start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
        vals = {k: float(torch.as_tensor(metrics[k]).float().cpu().reshape(-1)[0]) for k in ("model_loss", "connector_model_loss", "actor_loss", "critic_loss")}
        assert all(np.isfinite(v) for v in vals.values()), vals
        assert outputs["post"]["stoch"].shape == (B, T, 48, 48) and state["deter"].shape == (B, 4096)
        first = first or vals
    assert vals["model_loss"] < first["model_loss"]          # three updates on the same batch reduce the world-model loss


def test_agent_vlm_reward_imag_behavior_matches_reference_golden_fp32():
    """SURVEY §8 a19/a20 on the CUDA path: cached connector target, video_text_reward and update_imag_behavior vs the reference's golden
    vectors (tests/golden/tiny_genrl_vlm.*, pinned on CPU by tests/test_oracle_golden.py)."""
    from genrl_b200 import ops
    from genrl_b200.agent.genrl import GenRLAgent
    from tests.test_oracle_golden import _load_vlm_case
    g, meta, _ = _load_vlm_case()
    cfg = genrl_config(device=DEV, precision=32, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"], agent="genrl",
                       **meta["overrides"])
    obs_space, act_spec = make_specs(action_dim=meta["A"], clip_dim=meta["clip"])
    torch.manual_seed(1)
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    agent.load_state_dict(g["init"], strict=True)
    data = {k: v.to(DEV) for k, v in g["data"].items()}
    args = cfg.imag_reward_args
    N, steps = meta["B"] * meta["T"], meta["H"] + 1
    # the cached target, created the way genrl_utils.video_text_reward would (skip_first_target), on the reference's draws
    ops.set_noise(ops.TapeNoise(g["target_tape"], DEV))
    with torch.no_grad(), agent._ctx():
        agent.ensure_store()
        ve = g["prompt"].to(DEV).reshape(1, 1, -1).repeat(N, steps + 1, 1)
        stats = agent.wm.connector.video_imagine(ve, dreamer_init=None, sample=args["sample_for_target"], reset_every_n_frames=False, denoise=True)
        agent.unconditional_target = {k: v[:, 1:].permute([1, 0] + list(range(2, v.dim()))).contiguous() for k, v in stats.items()}
    assert torch.equal(agent.unconditional_target["stoch"].float().cpu(), g["target"]["stoch"])
    relclose(agent.unconditional_target["logit"], g["target"]["logit"], 1e-4, 2e-5, "target logit")
    ops.set_noise(ops.TapeNoise(g["tape"], DEV))
    for step in range(2):
        state, outputs, metrics = agent.update_wm(data, step)
        start, metrics = agent.update_imag_behavior(state, outputs, metrics, seq_data=data)
        for k, v in meta["metrics"][step].items():
            assert k in metrics, f"metric {k} missing"
            mv = float(torch.as_tensor(metrics[k]).float().cpu().reshape(-1)[0])
            tol = 2e-3 if k.endswith("grad_norm") else 2e-4
            assert abs(mv - v) <= tol * max(1.0, abs(v)) + 1e-5, f"step {step} metric {k}: {mv} vs reference {v}"
    sd = agent.state_dict()
    for k, v in g["params"].items():
        err = float((sd[k].float().cpu() - v).abs().max())
        assert err <= 5e-6 + 1e-4 * float(v.abs().max()), f"parameter {k} after 2 updates: err {err:.3e}"
    ops.set_noise(ops.PhiloxNoise(0))


def test_agent_report_videos_match_reference_golden_fp32():
    """SURVEY §8 f3 on the CUDA path: DreamerAgent.report / GenRLAgent.report vs the reference's golden videos
    (tests/golden/tiny_genrl_report.*, pinned on CPU by tests/test_oracle_golden.py)."""
    from genrl_b200 import ops
    from genrl_b200.agent.genrl import GenRLAgent
    from tests.test_oracle_golden import _load_report_case
    g, meta, _ = _load_report_case()
    cfg = genrl_config(device=DEV, precision=32, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"], agent="genrl",
                       **meta["overrides"])
    obs_space, act_spec = make_specs(action_dim=meta["A"], clip_dim=meta["clip"])
    torch.manual_seed(1)
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    agent.load_state_dict(g["init"], strict=True)
    data = {k: v.to(DEV) for k, v in g["data"].items()}
    ops.set_noise(ops.TapeNoise(g["tape"], DEV))
    rep = agent.report(data)
    assert ops.noise.pos == len(ops.noise.items), "every reference draw must be consumed"
    assert set(rep) == {"openl_observation", "video_clip_pred"}
    for k, ref in g["report"].items():
        v = rep[k].float().cpu()
        assert v.shape == (meta["B"], meta["T"], 3, 192, 64)
        relclose(v[:, :, :, 64:128], ref.float(), 0.0, 2e-3, f"{k} model frames")
    ops.set_noise(ops.PhiloxNoise(0))


def test_agent_act_matches_reference_golden_fp32():
    """SURVEY §8 f4 on the CUDA path: agent.act() over the reference's 6-step episode (tests/golden/tiny_genrl_act.*)."""
    from genrl_b200 import ops
    from genrl_b200.agent.genrl import GenRLAgent
    from tests.test_oracle_golden import _load_act_case, act_observation
    g, meta, _ = _load_act_case()
    cfg = genrl_config(device=DEV, precision=32, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"], agent="genrl",
                       **meta["overrides"])
    obs_space, act_spec = make_specs(action_dim=meta["A"], clip_dim=meta["clip"])
    torch.manual_seed(1)
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    agent.load_state_dict(g["init"], strict=True)
    ops.set_noise(ops.TapeNoise(g["tape"], DEV))
    state = None
    for i, (first, ev) in enumerate(meta["steps"]):
        a, state = agent.act(act_observation(g["observation"][i], first), None, i, eval_mode=ev, state=state)
        assert np.abs(np.asarray(a) - g["actions"][i]).max() <= 1e-4, f"step {i}: {a} vs {g['actions'][i]}"
    assert ops.noise.pos == len(ops.noise.items), "every reference draw must be consumed"
    ops.set_noise(ops.PhiloxNoise(0))


def test_agent_update_matches_reference_golden_kitchen_shaped_single_sequence_fp32():
    """tests/golden/tiny_kitchen.* on the CUDA path: A = 9, extra `state` key, B = 1 sequence with resets (slim fixture: outputs + metrics)."""
    from genrl_b200 import ops
    from genrl_b200.agent.genrl import GenRLAgent
    from genrl_b200.config import Spec
    from tests.test_oracle_golden import seeded_init_for
    g, meta, _ = load_case("tiny_kitchen")
    cfg = genrl_config(device=DEV, precision=32, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"], agent="genrl",
                       **meta["overrides"])
    obs_space, act_spec = make_specs(action_dim=meta["A"], clip_dim=meta["clip"])
    for k, d in meta["extra_obs"].items():
        obs_space[k] = Spec((d,), "float32")
    torch.manual_seed(1)
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    agent.load_state_dict(seeded_init_for(meta), strict=True)
    data = {k: v.to(DEV) for k, v in g["data"].items()}
    ops.set_noise(ops.TapeNoise(g["tape"], DEV))
    for step in range(2):
        state, outputs, metrics = agent.update_wm(data, step)
        start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
        ref = g[f"ref{step}"]
        for grp in ("post", "prior"):
            assert torch.equal(outputs[grp]["stoch"].float().cpu(), ref[f"outputs/{grp}/stoch"]), (step, grp)
            for k in ("deter", "logit"):
                relclose(outputs[grp][k], ref[f"outputs/{grp}/{k}"], 1e-4, 2e-5, f"step {step} {grp}.{k}")
        for k, v in meta["metrics"][step].items():
            mv = float(torch.as_tensor(metrics[k]).float().cpu().reshape(-1)[0])
            tol = 2e-3 if k.endswith("grad_norm") else 2e-4
            assert abs(mv - v) <= tol * max(1.0, abs(v)) + 1e-5, f"step {step} metric {k}: {mv} vs reference {v}"
    assert ops.noise.pos == len(ops.noise.items)
    ops.set_noise(ops.PhiloxNoise(0))
