"""TEST INFRASTRUCTURE ONLY — golden vectors for agent.act() from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_act
The reference DreamerAgent.act (agent/dreamer.py:41-64) is stepped over a short synthetic episode — three steps in training mode (sampled
actions), an episode reset, three steps in eval mode (mean actions), carrying the (latent, action) state — the oracle is asserted to
reproduce every action, and inputs, draws and actions go to tests/golden/tiny_genrl_act.{npz,json} (initial parameters shared with
tests/golden/tiny_genrl.npz: same architecture and seed).
"""
import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genrl_b200.config import genrl_config, tiny_overrides, make_specs   # noqa: E402
from oracle import genrl_oracle as O                                        # noqa: E402
from oracle.ref_harness import make_reference_agent                        # noqa: E402

warnings.filterwarnings("ignore")
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASE = dict(B=3, T=16, H=5, A=6, clip=32, kind="genrl", over=tiny_overrides())
NAME = "tiny_genrl_act"
STEPS = [(True, False), (False, False), (False, False), (True, True), (False, True), (False, True)]      # (is_first, eval_mode)


def episode(seed=9):
    g = torch.Generator().manual_seed(seed)
    return [dict(observation=torch.randint(0, 256, (3, 64, 64), generator=g, dtype=torch.uint8).numpy(), reward=np.zeros((), np.float32),
                 is_first=np.array(first), is_last=np.array(False), is_terminal=np.array(False), discount=np.ones((), np.float32))
            for first, _ in STEPS]


def main():
    torch.set_num_threads(1)
    c = CASE
    cfg = genrl_config(device="cpu", precision=32, batch_size=c["B"], batch_length=c["T"], imag_horizon=c["H"], agent="genrl", **c["over"])
    obs_space, act_spec = make_specs(action_dim=c["A"], clip_dim=c["clip"], with_clip=True)
    agent = make_reference_agent(cfg, obs_space, act_spec, kind="genrl", seed=1)
    init = {k: v.detach().clone() for k, v in agent.state_dict().items()}
    shared = np.load(os.path.join(GOLDEN_DIR, "tiny_genrl.npz"))
    for k, v in init.items():
        assert np.array_equal(shared[f"init/{k}"], v.numpy()), f"initial parameter {k} differs from tiny_genrl.npz"
    obs_list = episode()
    torch.manual_seed(7)
    ref_actions, state = [], None
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        for obs, (_, ev) in zip(obs_list, STEPS):
            a, state = agent.act(obs, None, 0, eval_mode=ev, state=state)
            ref_actions.append(np.asarray(a))
    oagent = O.OracleAgent(init, cfg, kind="genrl")
    torch.manual_seed(7)
    state = None
    with torch.no_grad():
        for i, (obs, (_, ev)) in enumerate(zip(obs_list, STEPS)):
            a, state = O.act_step(oagent.P, obs, state, cfg, oagent.tape, ev, c["A"])
            err = float(np.abs(a.numpy() - ref_actions[i]).max())
            print(f"[{NAME}] step {i} eval={ev}: max abs err {err:.2e}")
            assert err < 2e-6, i
    out = {"observation": np.stack([o["observation"] for o in obs_list]), "actions": np.stack(ref_actions)}
    for i, (kind, v) in enumerate(oagent.tape.items):
        out[f"tape/{i:04d}_{kind}"] = v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)
    meta = dict(case=NAME, **{k: v for k, v in c.items() if k != "over"}, overrides=c["over"], steps=STEPS, torch=torch.__version__,
                threads=torch.get_num_threads(), seed_init=1, seed_episode=9, seed_run=7, init_from="tiny_genrl.npz", reference_commit="4cd8394")
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{NAME}.npz"), **out)
    with open(os.path.join(GOLDEN_DIR, f"{NAME}.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(f"[{NAME}] wrote {len(out)} arrays, {os.path.getsize(os.path.join(GOLDEN_DIR, NAME + '.npz')) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
