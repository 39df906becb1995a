"""TEST INFRASTRUCTURE ONLY — golden vectors for agent.report() from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_report
GenRLAgent.report (agent/genrl.py:64-106) = DreamerAgent.report (dreamer.py:98-108: `openl_observation` from WorldModel.video_pred
:307-321) + `video_clip_pred`.  `additional_report_fns` is emptied (report_text2video needs the InternVideo2 text encoder).  The oracle
is asserted to reproduce both videos; inputs, the draw tape and the reference's videos go to tests/golden/tiny_genrl_report.{npz,json}.
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
# same architecture and seed as the tiny_genrl case: the initial parameters are shared with tests/golden/tiny_genrl.npz (asserted below)
CASE = dict(B=2, T=16, H=5, A=6, clip=32, resets=0.1, kind="genrl", over={**tiny_overrides(), "additional_report_fns": []})
NAME = "tiny_genrl_report"
to_np = lambda x: x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def main():
    torch.set_num_threads(1)
    c = CASE
    cfg = genrl_config(device="cpu", precision=32, batch_size=c["B"], batch_length=c["T"], imag_horizon=c["H"], agent="genrl", **c["over"])
    obs_space, act_spec = make_specs(action_dim=c["A"], clip_dim=c["clip"], with_clip=True)
    agent = make_reference_agent(cfg, obs_space, act_spec, kind="genrl", seed=1)
    init = {k: v.detach().clone() for k, v in agent.state_dict().items()}
    data = O.synthetic_batch(c["B"], c["T"], c["A"], clip_dim=c["clip"], seed=3, resets=c["resets"])
    torch.manual_seed(7)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        ref = agent.report(dict(data))
    assert set(ref) == {"openl_observation", "video_clip_pred"}, set(ref)

    oagent = O.OracleAgent(init, cfg, kind="genrl")
    torch.manual_seed(7)
    with torch.no_grad():
        d = oagent.cast(data)
        mine = {"openl_observation": O.video_pred(oagent.P, d, cfg, oagent.tape), "video_clip_pred": O.video_clip_pred(oagent.P, d, cfg, oagent.tape)}
    for k in ref:
        err = float((ref[k] - mine[k]).abs().max())
        print(f"[{NAME}] {k}: shape {tuple(ref[k].shape)} max abs err {err:.2e}")
        assert ref[k].shape == mine[k].shape and err < 2e-5, k

    shared = np.load(os.path.join(GOLDEN_DIR, "tiny_genrl.npz"))
    for k, v in init.items():
        assert np.array_equal(shared[f"init/{k}"], to_np(v)), f"initial parameter {k} differs from tiny_genrl.npz"
    out = {}
    for k, v in data.items():
        out[f"data/{k}"] = to_np(v)
    for i, (kind, v) in enumerate(oagent.tape.items):
        out[f"tape/{i:04d}_{kind}"] = to_np(v)
    for k, v in ref.items():
        # only the model third of [truth | model | error] (truth is the input, error follows from both), videos in [0, 1] as fp16 (err 5e-4)
        out[f"report/{k}"] = to_np(v)[:, :, :, 64:128].astype(np.float16)
    meta = dict(case=NAME, **{k: v for k, v in c.items() if k != "over"}, overrides=c["over"], torch=torch.__version__, threads=torch.get_num_threads(),
                seed_init=1, seed_data=3, seed_run=7, init_from="tiny_genrl.npz", reference_commit="4cd8394")
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{NAME}.npz"), **out)
    with open(os.path.join(GOLDEN_DIR, f"{NAME}.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(f"[{NAME}] wrote {len(out)} arrays, {os.path.getsize(os.path.join(GOLDEN_DIR, NAME + '.npz')) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
