"""CPU: the oracle restatement reproduces the reference's golden vectors (tests/golden/, made by
oracle/make_golden.py from the unmodified reference: outputs, metrics, gradients, parameters after two updates)."""
import numpy as np
import pytest
import torch

from oracle import genrl_oracle as O
from tests.golden_utils import load_case


@pytest.mark.parametrize("case", ["tiny_genrl", "tiny_dreamer"])
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
