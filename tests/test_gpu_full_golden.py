"""GPU parity of the BENCHMARKED code path against the unmodified reference at the benchmarked architecture.

tests/golden/full_genrl.* (oracle/make_golden_full.py) holds the reference GenRLAgent's results at deter = hidden = 1024, stoch 32x32,
conv depth 48, A = 10 on a 4 x 16 batch with resets, for two rounds of
    update_wm -> wm.update_additional_detached_modules -> update_acting_behavior -> update_imag_behavior (video_text_reward)
i.e. the S2 sequence bench.py times plus the S3 calls of train.py:275,280,340.

* fp32 mode  : north_star tolerance 1e-4 relative, categorical sample indices bit-exact (0 flips).
* bf16 mode  : the tensor-core path the bench runs (tcgen05 GEMM, implicit-GEMM convolutions, bf16 scan, fast row kernels);
               north_star tolerance 1e-2 relative on losses, likes and reconstructed pixels; the number of posterior / prior sample
               indices that differ from the reference is counted, reported and bounded.
Every comparison is collected into gpurun_out/full_golden_<mode>.json (when that directory exists) before asserting.
"""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from genrl_b200.config import genrl_config, make_specs

DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_full():
    z = np.load(os.path.join(GOLDEN, "full_genrl.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "full_genrl.json")))
    g = {"data": {}, "tape": [], "target_tape": [], "ref0": {}, "ref1": {}, "params": {}, "reward": {}}
    for k in z.files:
        arr = torch.from_numpy(z[k])
        if "/" not in k:
            g[k] = arr
            continue
        head, rest = k.split("/", 1)
        if head in ("tape", "target_tape"):
            g[head].append((rest, arr))
        elif head == "ref1" and rest.startswith("params/"):
            g["params"][rest[len("params/"):]] = arr
        elif head == "reward":
            g["reward"][int(rest)] = arr
        elif head == "target":
            g["target_idx"] = arr
        else:
            g[head][rest] = arr
    for t in ("tape", "target_tape"):
        g[t] = [(name.split("_", 1)[1], v) for name, v in sorted(g[t], key=lambda kv: kv[0])]
    return g, meta


def one_hot(idx, K):
    return torch.nn.functional.one_hot(idx.long(), K).float()


def build(meta, precision):
    from genrl_b200.agent.genrl import GenRLAgent
    cfg = genrl_config(device=DEV, precision=precision, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"], agent="genrl",
                       **meta["overrides"])
    obs_space, act_spec = make_specs(action_dim=meta["A"], clip_dim=meta["clip"])
    torch.manual_seed(meta["seed_init"])
    agent = GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    sd = agent.state_dict()
    sha = hashlib.sha256()
    for k in sorted(sd):
        sha.update(k.encode()); sha.update(sd[k].detach().cpu().numpy().tobytes())
    assert sha.hexdigest() == meta["init_sha256"], "seeded initial parameters differ from the reference's"
    return agent, cfg


class Report:
    def __init__(self, mode):
        self.mode, self.rows, self.fail = mode, [], []

    def close(self, name, mine, ref, rtol, atol=0.0):
        a = torch.as_tensor(mine).detach().float().cpu().reshape(-1)
        b = torch.as_tensor(ref).detach().float().cpu().reshape(-1)
        assert a.shape == b.shape, (name, a.shape, b.shape)
        scale = max(float(b.abs().max()), 1e-12)
        err = float((a - b).abs().max())
        ok = err <= atol + rtol * scale
        self.rows.append({"what": name, "max_abs_err": err, "scale": scale, "rel": err / scale, "rtol": rtol, "atol": atol, "ok": ok})
        if not ok:
            self.fail.append(f"{name}: max abs err {err:.3e} vs scale {scale:.3e} (rtol {rtol}, atol {atol})")

    def metric(self, name, mine, ref, rtol):
        mv = float(torch.as_tensor(mine).float().cpu().reshape(-1)[0])
        ok = abs(mv - ref) <= rtol * max(1.0, abs(ref)) + 1e-6
        self.rows.append({"what": name, "mine": mv, "ref": ref, "rel": abs(mv - ref) / max(1.0, abs(ref)), "rtol": rtol, "ok": ok})
        if not ok:
            self.fail.append(f"{name}: {mv} vs reference {ref} (rtol {rtol})")

    def note(self, **kw):
        self.rows.append(kw)

    def finish(self):
        out = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out):
            json.dump({"mode": self.mode, "failures": self.fail, "rows": self.rows}, open(os.path.join(out, f"full_golden_{self.mode}.json"), "w"), indent=1)
        assert not self.fail, f"{len(self.fail)} mismatches vs the reference golden ({self.mode}):\n  " + "\n  ".join(self.fail[:25])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_full_size_agent_matches_reference_golden(mode):
    from genrl_b200 import ops
    g, meta = load_full()
    precision = 32 if mode == "fp32" else 16
    # tolerances: north_star — 1e-4 relative in fp32, 1e-2 in bf16 (losses, likes, reconstructed pixels); gradient norms get 5e-2 in
    # bf16 (measured: <= 4e-3).  Parameters: Adam's first steps move every element by ~lr * sign(g), so wherever a gradient component
    # sits at the bf16 rounding-noise level two correct implementations may disagree by 2 lr per step; the fp32 mode is held to
    # 1e-4 relative + 5e-6 (measured 0.17 of that), the bf16 mode to 1e-2 relative + 2.2 lr x (number of Adam steps the tensor took).
    tol = 1e-4 if precision == 32 else 1e-2
    tol_gn = 2e-3 if precision == 32 else 5e-2
    agent, cfg = build(meta, precision)
    rep = Report(mode)
    B, T, H, S, K = meta["B"], meta["T"], meta["H"], 32, 32
    N = B * T
    data = {k: v.to(DEV) for k, v in g["data"].items()}
    tgt = one_hot(g["target_idx"], K).to(DEV)                                  # [H+1, N, S, K]
    agent.unconditional_target = {"stoch": tgt, "deter": torch.zeros(H + 1, N, 1024, device=DEV), "logit": torch.zeros_like(tgt)}
    # reconstructed pixels: the decoder (initial weights) on the reference's posterior sample of sequence 0
    with torch.no_grad(), agent._ctx():
        agent.ensure_store()
        st0 = one_hot(g["ref0"]["outputs/post/stoch_idx"][:1], K).to(DEV)      # [1, T, S, K]
        recon = agent.wm.heads["decoder"](agent.wm.decoder_input_fn({"stoch": st0}))["observation"].mean[0]
    rep.close("reconstructed pixels (initial decoder on the reference's post.stoch[0])", recon, g["recon0"], tol, atol=tol * 1e-2)
    ops.set_noise(ops.TapeNoise(g["tape"], DEV))
    flips_total = {"post": 0, "prior": 0}
    for step in range(2):
        ref = g[f"ref{step}"]
        state, outputs, metrics = agent.update_wm(data, step)
        for k, v in meta["metrics_wm"][step].items():
            rep.metric(f"round {step} update_wm {k}", metrics[k], v, tol_gn if k.endswith("grad_norm") else tol)
        for grp in ("post", "prior"):
            mine = outputs[grp]["stoch"].float().argmax(-1).cpu()
            assert float(outputs[grp]["stoch"].float().sum(-1).min()) == 1.0 == float(outputs[grp]["stoch"].float().max())
            flips = int((mine != ref[f"outputs/{grp}/stoch_idx"].long()).sum())
            flips_total[grp] += flips
            rep.note(what=f"round {step} {grp} sample-index flips", flips=flips, of=B * T * S)
        # element-wise activations: round 0 runs on identical weights; in round 1 the bf16 mode's weights already carry the Adam
        # sign noise described above (measured 2-4e-2 of the tensor's range on single elements), so only round 0 is held to 1e-2 there
        tol_e = tol if (precision == 32 or step == 0) else 5e-2
        rep.close(f"round {step} embed", outputs["embed"], ref["outputs/embed"], tol_e, atol=tol * 1e-1)
        rep.close(f"round {step} post.logit", outputs["post"]["logit"], ref["outputs/post/logit"], tol_e, atol=tol * 1e-1)
        if precision == 32:       # deter / prior depend on the sampled stoch history: comparable only while no index flipped
            for grp, k in (("post", "deter"), ("prior", "logit")):
                rep.close(f"round {step} {grp}.{k}", outputs[grp][k], ref[f"outputs/{grp}/{k}"], tol, atol=2e-5)
            rep.close(f"round {step} kl value", outputs["kl"], ref["outputs/kl"], tol, atol=2e-5)
        # per-frame likelihoods: a frame whose posterior sample differs from the reference's decodes a different latent, so the
        # element-wise comparison covers the frames whose 32 sample indices all agree (every frame in fp32); the batch means
        # (observation_loss / reward_loss above) cover all frames
        same = (outputs["post"]["stoch"].float().argmax(-1).cpu() == ref["outputs/post/stoch_idx"].long()).all(-1)      # [B, T]
        rep.note(what=f"round {step} frames with an identical posterior sample", frames=int(same.sum()), of=B * T)
        for k in ("observation", "reward"):
            rep.close(f"round {step} likes.{k} (frames with identical samples)", outputs["likes"][k].float().cpu()[same],
                      ref[f"outputs/likes/{k}"][same], tol, atol=tol)
        _, metrics = agent.wm.update_additional_detached_modules(data, outputs, metrics)
        for k, v in meta["metrics_conn2"][step].items():
            rep.metric(f"round {step} second connector step {k}", metrics[k], v, tol_gn if k.endswith("grad_norm") else tol)
        _, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
        _, metrics = agent.update_imag_behavior(state, outputs, metrics, seq_data=data)
        for k, v in meta["metrics"][step].items():
            assert k in metrics, f"metric {k} missing"
            rep.metric(f"round {step} {k}", metrics[k], v, tol_gn if k.endswith("grad_norm") else tol)
    assert ops.noise.pos == len(ops.noise.items), "every reference draw must be consumed"
    sd = agent.state_dict()
    worst = 0.0
    for k, v in g["params"].items():
        err = float((sd[k].float().cpu() - v).abs().max())
        if precision == 32:
            allowed = 5e-6 + 1e-4 * float(v.abs().max())
        else:
            lr = cfg.model_opt["lr"] if k.startswith("wm.") else cfg.actor_opt["lr"]
            adam_steps = 2 * (2 if k.startswith("wm.connector.") else 1)
            allowed = 2.2 * lr * adam_steps + 1e-2 * float(v.abs().max())
        worst = max(worst, err / allowed)
    rep.rows.append({"what": "1-D parameters after two rounds: worst err / tolerance", "ratio": worst, "ok": worst <= 1.0})
    if worst > 1.0:
        rep.fail.append(f"parameters after two rounds: worst error is {worst:.2f}x the tolerance")
    rep.note(what="sample-index flips over both rounds", **flips_total, draws_per_group=2 * B * T * S)
    if precision == 32:
        if flips_total["post"] or flips_total["prior"]:
            rep.fail.append(f"fp32 mode: {flips_total} categorical sample indices differ from the reference (must be bit-exact)")
    else:
        # bf16 logits carry ~3 significant digits: a draw flips when its two best p/q ratios are closer than that.  Posterior draws are
        # independent (single_obs_posterior), prior draws additionally inherit the history of their sequence.
        if flips_total["post"] > 0.02 * 2 * B * T * S:
            rep.fail.append(f"bf16 mode: {flips_total['post']} posterior index flips (> 2 % of {2 * B * T * S})")
        if flips_total["prior"] > 0.10 * 2 * B * T * S:
            rep.fail.append(f"bf16 mode: {flips_total['prior']} prior index flips (> 10 % of {2 * B * T * S})")
    ops.set_noise(ops.PhiloxNoise(0))
    rep.finish()
