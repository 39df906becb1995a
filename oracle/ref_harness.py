"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference). Used by
oracle/make_golden.py to generate tests/golden/* fixtures and to pin oracle/genrl_oracle.py
against the real reference. Nothing in the product (genrl_b200/) imports this.
"""
import sys
import contextlib
import io

REFERENCE_ROOT = "/root/reference"


def import_reference():
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        import agent.genrl as ref_genrl          # noqa: F401  (reference agent/genrl.py)
        import agent.dreamer as ref_dreamer      # noqa: F401
        import agent.dreamer_utils as ref_utils  # noqa: F401
        import agent.video_utils as ref_video    # noqa: F401
    return ref_genrl, ref_dreamer, ref_utils, ref_video


def make_reference_agent(cfg, obs_space, act_spec, kind="genrl", seed=1):
    """Instantiate the reference agent the way train.py:32-37 does (minus Hydra)."""
    import torch
    ref_genrl, ref_dreamer, _, _ = import_reference()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "genrl":
            agent = ref_genrl.GenRLAgent(name="genrl", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
        else:
            agent = ref_dreamer.DreamerAgent(name="dreamer", cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    return agent
