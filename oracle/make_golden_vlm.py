"""TEST INFRASTRUCTURE ONLY — golden vectors for the VLM-reward imagination behaviour from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_vlm
The reference's `video_text_reward` (tools/genrl_utils.py:279-370) needs the InternVideo2 text encoder only to obtain ONE prompt embedding;
here a fixed unit-norm vector stands in for it and `agent.unconditional_target` is pre-set exactly as genrl_utils.py:305-316 would cache it
(SURVEY appendix A6).  Then two rounds of `update_wm` + `update_imag_behavior` run on the reference GenRLAgent (CPU, fp32, tiny config),
the rewards returned by the reference's reward function are recorded by wrapping it, the oracle is asserted to reproduce everything, and
the vectors are written to tests/golden/tiny_genrl_vlm.{npz,json}.
"""
import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genrl_b200.config import genrl_config, tiny_overrides, make_specs   # noqa: E402
from oracle import genrl_oracle as O                                        # noqa: E402
from oracle.ref_harness import make_reference_agent, import_reference      # noqa: E402

warnings.filterwarnings("ignore")
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASE = dict(B=2, T=16, H=10, A=5, clip=32, resets=0.1, kind="genrl", over={**tiny_overrides(), "imag_reward_fn": "video_text_reward"})
NAME = "tiny_genrl_vlm"
to_np = lambda x: x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def prompt_embedding(dim, seed=5):
    g = torch.Generator().manual_seed(seed)
    return F.normalize(torch.randn(dim, generator=g), dim=-1)


def main():
    torch.set_num_threads(1)
    c = CASE
    cfg = genrl_config(device="cpu", precision=32, batch_size=c["B"], batch_length=c["T"], imag_horizon=c["H"], agent="genrl", **c["over"])
    obs_space, act_spec = make_specs(action_dim=c["A"], clip_dim=c["clip"], with_clip=True)
    ref_genrl, _, _, _ = import_reference()
    agent = make_reference_agent(cfg, obs_space, act_spec, kind="genrl", seed=1)
    init = {k: v.detach().clone() for k, v in agent.state_dict().items()}
    data = O.synthetic_batch(c["B"], c["T"], c["A"], clip_dim=c["clip"], seed=3, resets=c["resets"])
    emb = prompt_embedding(c["clip"])
    N, steps = c["B"] * c["T"], c["H"] + 1
    args = cfg.imag_reward_args
    # --- the reference's cached target (genrl_utils.py:305-316) from the stand-in prompt embedding
    # (VideoSSM.initial samples the initial stoch even when sample_for_target is False, video_utils.py:109: the target depends on the
    #  generator state at caching time, so it is seeded explicitly here and in the oracle)
    with torch.no_grad():
        torch.manual_seed(11)
        ve = emb.reshape(1, 1, -1).repeat(N, steps + 1, 1)
        stats = agent.wm.connector.video_imagine(ve, dreamer_init=None, sample=args["sample_for_target"], reset_every_n_frames=False, denoise=True)
        agent.unconditional_target = {k: v[:, 1:].permute([1, 0] + list(range(2, len(v.shape)))) for k, v in stats.items()}
    ref_target = {k: v.detach().clone() for k, v in agent.unconditional_target.items()}
    # --- record what the reference's reward function returns
    rewards = []
    orig = ref_genrl.video_text_reward

    def recording(agent_, seq, **kw):
        r = orig(agent_, seq, **kw)
        rewards.append(r.detach().clone())
        return r
    ref_genrl.video_text_reward = recording
    ref = []
    try:
        torch.manual_seed(7)
        for step in range(2):
            with contextlib.redirect_stdout(io.StringIO()):
                state, outputs, metrics = agent.update_wm(data, step)
                start, metrics = agent.update_imag_behavior(state, outputs, metrics, seq_data=data)
            ref.append(dict(metrics={k: float(np.asarray(to_np(v)).reshape(-1)[0]) for k, v in metrics.items()},
                            params={k: v.detach().clone() for k, v in agent.state_dict().items()}))
    finally:
        ref_genrl.video_text_reward = orig

    # --- oracle
    oagent = O.OracleAgent(init, cfg, kind="genrl")
    scratch = O.NoiseTape()                                   # draws of the target rollout: not part of the update tape
    torch.manual_seed(11)
    o_target = O.unconditional_target(oagent.P, emb, N, steps, cfg, scratch, skip_first=args["skip_first_target"], sample=args["sample_for_target"])
    for k in ("stoch", "deter", "logit"):
        assert torch.allclose(o_target[k], ref_target[k], rtol=2e-5, atol=2e-6), f"target {k}"
    assert torch.equal(o_target["stoch"], ref_target["stoch"])
    torch.manual_seed(7)
    orc = []
    for step in range(2):
        state, outs, mets = oagent.update_wm(data)
        mets.update(oagent.update_imag_behavior(outs, o_target))
        orc.append(dict(metrics={k: float(v) for k, v in mets.items()}, params={k: v.clone() for k, v in oagent.P.items()}))
    bad = []
    for step in range(2):
        for k, v in ref[step]["metrics"].items():
            ov = orc[step]["metrics"].get(k)
            assert ov is not None, f"oracle lacks metric {k}"
            if not np.allclose(v, ov, rtol=1e-4, atol=1e-6):
                bad.append(("metric", step, k, v, ov))
        for k, v in ref[step]["params"].items():
            if not torch.allclose(v.float(), orc[step]["params"][k].float(), rtol=1e-5, atol=2e-6):
                bad.append(("param", step, k, float((v - orc[step]["params"][k]).abs().max())))
    print(f"[{NAME}] oracle vs reference: {len(bad)} mismatches; reward mean per step {[float(r.mean()) for r in rewards]}")
    for b in bad[:20]:
        print("   MISMATCH", b)
    assert not bad, "oracle does not reproduce the reference"

    out = {}
    for k, v in init.items():
        out[f"init/{k}"] = to_np(v)
    for k, v in data.items():
        out[f"data/{k}"] = to_np(v)
    out["prompt_embedding"] = to_np(emb)
    for k, v in ref_target.items():
        out[f"target/{k}"] = to_np(v)
    for i, (kind, v) in enumerate(oagent.tape.items):
        out[f"tape/{i:04d}_{kind}"] = to_np(v)
    for i, (kind, v) in enumerate(scratch.items):
        out[f"target_tape/{i:04d}_{kind}"] = to_np(v)
    for step in range(2):
        out[f"reward/{step}"] = to_np(rewards[step])
    for k, v in ref[1]["params"].items():
        if k.startswith("_imag_behavior."):
            out[f"ref1/params/{k}"] = to_np(v)
    meta = dict(case=NAME, **{k: v for k, v in c.items() if k != "over"}, overrides=c["over"], torch=torch.__version__, threads=torch.get_num_threads(),
                seed_init=1, seed_data=3, seed_run=7, seed_prompt=5, seed_target=11, metrics=[r["metrics"] for r in ref], reference_commit="4cd8394")
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{NAME}.npz"), **out)
    with open(os.path.join(GOLDEN_DIR, f"{NAME}.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(f"[{NAME}] wrote {len(out)} arrays, {os.path.getsize(os.path.join(GOLDEN_DIR, NAME + '.npz')) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
