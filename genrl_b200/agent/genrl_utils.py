"""Reward of the imagination behaviour: the video-text reward on the kernel path.

What the reference computes (tools/genrl_utils.py:279-370 with agent/genrl.yaml:21 — score_fn max_cosine, align_sequence, skip_first_target):
a target latent sequence is rolled out ONCE by the connector from the task prompt's embedding and cached on the agent; at every behaviour
update the imagined latents and the target latents go through the decoder's input projection, the target is delayed by the window offset
at which its first n_frames match the imagined sequence best, and the per-step reward is the max-norm cosine between the two projections.

Here: two GEMMs (the projections, on the tcgen05 kernel, with the straight-through gradient flowing back into the rollout) and ONE fused
kernel for alignment scores + arg-max + delayed gather + reward (csrc/vlm_reward.cu), with its own backward kernel.
The InternVideo2 / ViCLIP text encoder is out of scope: provide `agent.prompt_embedding` (any [clip_dim] tensor) or pre-set
`agent.unconditional_target` the way train.py would find it cached.
"""
import torch

from .. import ops

SUPPORTED_SCORE = "max_cosine"


def _project(agent, stoch):
    """decoder input projection of flattened latents: [..., S, K] -> [..., E] in the compute dtype"""
    lin = agent.wm.heads["decoder"]._conv_in[0]
    flat = stoch.reshape(-1, stoch.shape[-2] * stoch.shape[-1])
    y = ops.linear(ops.to_compute(flat), lin.weight, lin.bias)
    return y.reshape(tuple(stoch.shape[:-2]) + (y.shape[-1],))


def build_unconditional_target(agent, steps, rows, sample=False, skip_first=True):
    """the cached target of genrl_utils.py:289-319: connector rollout from the prompt embedding, time-major [steps, rows, ...]"""
    emb = getattr(agent, "prompt_embedding", None)
    if emb is None:
        raise RuntimeError("video_text_reward needs agent.prompt_embedding (embedding of the task prompt) or a pre-set "
                           "agent.unconditional_target; the InternVideo2 encoder itself is out of scope")
    n = steps + 1 if skip_first else steps
    with torch.no_grad():
        ve = emb.to(agent.device).reshape(1, 1, -1).repeat(rows, n, 1)
        roll = agent.wm.connector.video_imagine(ve, dreamer_init=None, sample=sample, reset_every_n_frames=False, denoise=True)
        first = 1 if skip_first else 0
        return {k: v[:, first:].transpose(0, 1).contiguous() for k, v in roll.items()}


def compute_reward(agent, agent_seq, target_seq, score_fn=SUPPORTED_SCORE, n_frames=8, align=False):
    """[T, N] reward between an imagined sequence and a target sequence (both dicts with 'stoch' [T, N, S, K])"""
    if score_fn != SUPPORTED_SCORE:
        raise NotImplementedError(f"score_fn={score_fn!r}: the fused reward kernel implements '{SUPPORTED_SCORE}' (agent/genrl.yaml:21)")
    pa = _project(agent, agent_seq["stoch"])
    with torch.no_grad():
        pt = _project(agent, target_seq["stoch"].detach())
    reward, _ = ops.vlm_reward(pa, pt, n_frames, align)
    return reward


def video_text_reward(agent, seq, score_fn=SUPPORTED_SCORE, sample_for_target=False, weighted_align=False, align_initial=False,
                      align_sequence=False, task_prompt="", skip_first_target=False, **kwargs):
    if weighted_align or align_initial:
        raise NotImplementedError("weighted_align / align_initial are not GenRL defaults (agent/genrl.yaml:21) and have no fused kernel")
    steps, rows = seq["deter"].shape[:2]
    if not hasattr(agent, "unconditional_target"):
        agent.unconditional_target = build_unconditional_target(agent, steps, rows, sample=sample_for_target, skip_first=skip_first_target)
    reward = compute_reward(agent, seq, agent.unconditional_target, score_fn, agent.wm.connector.n_frames, bool(align_sequence))
    return reward.unsqueeze(-1)
