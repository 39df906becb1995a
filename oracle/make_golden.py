"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
For each case it
  1. instantiates the reference GenRLAgent (agent/genrl.py:27) on CPU, fp32, with a small config,
  2. runs `agent.update(data, step)` twice under a fixed seed, recording outputs, metrics,
     gradients (captured at the clip_grad_norm_ call inside Optimizer.__call__,
     dreamer_utils.py:913) and the post-update parameters,
  3. runs oracle/genrl_oracle.OracleAgent under the same seed and asserts it agrees,
  4. writes inputs, initial parameters, the oracle's noise tape and the REFERENCE's results.
The recorded environment (torch version, threads) is stored alongside.
"""
import os
import sys
import json
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genrl_b200.config import genrl_config, tiny_overrides, make_specs   # noqa: E402
from oracle import genrl_oracle as O                                        # noqa: E402
from oracle.ref_harness import make_reference_agent                        # noqa: E402

warnings.filterwarnings("ignore")
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (B, T, H, A, clip_dim, resets, overrides)
    "tiny_genrl": dict(B=3, T=16, H=5, A=6, clip=32, resets=0.15, kind="genrl", over=tiny_overrides()),
    "tiny_dreamer": dict(B=2, T=6, H=4, A=4, clip=32, resets=0.0, kind="dreamer",
                                over={**tiny_overrides(units=32, stoch=4, discrete=8),
                                      "actor_ent": 3e-4}),
    # BASELINE configs[3] flavour: 9 actions and an extra low-dimensional `state` observation key that the GenRL encoder/decoder ignore
    # (conf/env/dmc_pixels.yaml mlp_keys '$^'); single sequence (B = 1) as the ragged edge
    # conf/defaults/dreamer_v3.yaml:5 — single_obs_posterior: false: the posterior reads [deter | embed], so nothing can be hoisted out of
    # the time loop and observe() runs step by step (EnsembleRSSM._observe_stepwise); resets included
    "tiny_dreamer_v3": dict(B=3, T=8, H=4, A=4, clip=32, resets=0.2, kind="dreamer",
                            over=(lambda o: {**o, "rssm": {**o["rssm"], "single_obs_posterior": False}})(tiny_overrides(units=32, stoch=4, discrete=8))),
    "tiny_kitchen": dict(B=1, T=16, H=4, A=9, clip=32, resets=0.2, kind="genrl", over=tiny_overrides(), extra_obs={"state": 30}, slim=True),
}


def to_np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def run_reference(case, seed=7):
    c = CASES[case]
    cfg = genrl_config(device="cpu", precision=32, batch_size=c["B"], batch_length=c["T"],
                       imag_horizon=c["H"], agent=c["kind"], **c["over"])
    obs_space, act_spec = make_specs(action_dim=c["A"], clip_dim=c["clip"], with_clip=True)
    from genrl_b200.config import Spec
    for k, d in c.get("extra_obs", {}).items():
        obs_space[k] = Spec((d,), "float32")
    agent = make_reference_agent(cfg, obs_space, act_spec, kind=c["kind"], seed=1)
    init = {k: v.detach().clone() for k, v in agent.state_dict().items()}
    data = O.synthetic_batch(c["B"], c["T"], c["A"], clip_dim=c["clip"], seed=3, resets=c["resets"])
    for k, d in c.get("extra_obs", {}).items():
        data[k] = torch.rand(c["B"], c["T"], d, generator=torch.Generator().manual_seed(4))
    name_of = {id(p): n for n, p in agent.named_parameters()}
    grads_log = []
    orig_clip = torch.nn.utils.clip_grad_norm_

    def recording_clip(params, max_norm, *a, **k):
        params = list(params)
        grads_log.append({name_of[id(p)]: p.grad.detach().clone() for p in params if p.grad is not None})
        return orig_clip(params, max_norm, *a, **k)

    torch.nn.utils.clip_grad_norm_ = recording_clip
    results = []
    try:
        torch.manual_seed(seed)
        for step in range(2):
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                state, outputs, metrics = agent.update_wm(data, step)
                start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
            results.append(dict(
                state={k: v.detach().clone() for k, v in state.items()},
                outputs={k: ({kk: vv.detach().clone() for kk, vv in v.items()} if isinstance(v, dict) else v.detach().clone())
                         for k, v in outputs.items()},
                metrics={k: float(np.asarray(to_np(v)).reshape(-1)[0]) for k, v in metrics.items()},
                params={k: v.detach().clone() for k, v in agent.state_dict().items()},
            ))
    finally:
        torch.nn.utils.clip_grad_norm_ = orig_clip
    return cfg, init, data, results, grads_log


def run_oracle(cfg, init, data, kind, seed=7):
    agent = O.OracleAgent(init, cfg, kind=kind)
    torch.manual_seed(seed)
    res, grads = [], []
    for step in range(2):
        state, outs, mets = agent.update(data)
        res.append(dict(state=state, outputs=outs, metrics={k: float(v) for k, v in mets.items()},
                        params={k: v.clone() for k, v in agent.P.items()}))
        g = [agent.last_wm_grads]
        if kind == "genrl":
            g.append(agent.last_connector_grads)
        g += [agent.last_actor_grads, agent.last_critic_grads]
        grads += g
    return agent, res, grads


def compare(name, a, b, rtol=2e-5, atol=1e-6):
    a, b = to_np(a).astype(np.float64), to_np(b).astype(np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    err = np.abs(a - b).max() if a.size else 0.0
    ok = np.allclose(a, b, rtol=rtol, atol=atol)
    return ok, err


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(1)          # deterministic reduction order for the fixtures
    selected = sys.argv[1:] or list(CASES)            # python -m oracle.make_golden [case ...]: regenerate only the named fixtures
    for case, c in CASES.items():
        if case not in selected:
            continue
        cfg, init, data, ref, ref_grads = run_reference(case)
        oagent, orc, orc_grads = run_oracle(cfg, init, data, c["kind"])
        worst = 0.0
        bad = []
        for step in range(2):
            for k in ("stoch", "deter", "logit"):
                for grp in ("post", "prior"):
                    ok, e = compare(f"{grp}.{k}", ref[step]["outputs"][grp][k], orc[step]["outputs"][grp][k])
                    worst = max(worst, e); bad += [] if ok else [(step, grp, k, e)]
            for k in ("embed", "feat", "kl"):
                ok, e = compare(k, ref[step]["outputs"][k], orc[step]["outputs"][k])
                worst = max(worst, e); bad += [] if ok else [(step, k, e)]
            for k, v in ref[step]["metrics"].items():
                ov = orc[step]["metrics"].get(k)
                assert ov is not None, f"oracle lacks metric {k}"
                ok, e = compare(k, v, ov, rtol=1e-4, atol=1e-6)
                worst = max(worst, e); bad += [] if ok else [("metric", step, k, v, ov)]
            for k, v in ref[step]["params"].items():
                ok, e = compare(k, v, orc[step]["params"][k], rtol=1e-5, atol=2e-6)
                worst = max(worst, e); bad += [] if ok else [("param", step, k, e)]
        assert len(ref_grads) == len(orc_grads), (len(ref_grads), len(orc_grads))
        for i, (rg, og) in enumerate(zip(ref_grads, orc_grads)):
            assert set(rg) == set(og), (i, set(rg) ^ set(og))
            for k in rg:
                scale = float(rg[k].abs().max()) + 1e-12      # error relative to the tensor's own scale
                ok, e = compare(k, rg[k] / scale, og[k] / scale, rtol=0, atol=1e-4)
                bad += [] if ok else [("grad", i, k, e)]
        idx_equal = all(torch.equal(ref[s]["outputs"][g]["stoch"], orc[s]["outputs"][g]["stoch"])
                        for s in range(2) for g in ("post", "prior"))
        print(f"[{case}] oracle vs reference: worst abs err {worst:.3e}; sample indices bit-equal: {idx_equal}; "
              f"mismatches: {len(bad)}")
        for b in bad[:20]:
            print("   MISMATCH", b)
        assert not bad and idx_equal, "oracle does not reproduce the reference"

        out = {}
        slim = bool(c.get("slim"))       # slim fixtures: no initial / final parameters and gradients (the seeded init is re-created by the
        if not slim:                     # mirror and checked by hash; losses and grad norms of BOTH updates still pin the optimizer steps)
            for k, v in init.items():
                out[f"init/{k}"] = to_np(v)
        for k, v in data.items():
            out[f"data/{k}"] = to_np(v)
        for i, (kind, v) in enumerate(oagent.tape.items):
            out[f"tape/{i:04d}_{kind}"] = to_np(v)
        for step in range(2):
            r = ref[step]
            for grp in ("post", "prior"):
                for k, v in r["outputs"][grp].items():
                    out[f"ref{step}/outputs/{grp}/{k}"] = to_np(v)
            for k in ("embed", "feat", "kl"):
                out[f"ref{step}/outputs/{k}"] = to_np(r["outputs"][k])
            for k, v in r["outputs"]["likes"].items():
                out[f"ref{step}/outputs/likes/{k}"] = to_np(v)
            if step == 1 and not slim:          # parameters after both updates (keeps the fixture small)
                for k, v in r["params"].items():
                    out[f"ref{step}/params/{k}"] = to_np(v)
        if not slim:
            for i, rg in enumerate(ref_grads):
                for k, v in rg.items():
                    out[f"grads/{i}/{k}"] = to_np(v)
        import hashlib
        sha = hashlib.sha256()
        for k in sorted(init):
            sha.update(k.encode()); sha.update(to_np(init[k]).tobytes())
        meta = dict(case=case, **{k: v for k, v in c.items() if k != "over"}, overrides=c["over"], init_sha256=sha.hexdigest(),
                    torch=torch.__version__, threads=torch.get_num_threads(), seed_init=1, seed_data=3, seed_run=7,
                    metrics=[r["metrics"] for r in ref], reference_commit="4cd8394",
                    grad_order=(["wm", "connector", "actor", "critic"] if c["kind"] == "genrl"
                                else ["wm", "actor", "critic"]))
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{case}.npz"), **out)
        with open(os.path.join(GOLDEN_DIR, f"{case}.json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        print(f"[{case}] wrote {len(out)} arrays, {os.path.getsize(os.path.join(GOLDEN_DIR, case + '.npz'))/1e6:.2f} MB")


if __name__ == "__main__":
    main()
