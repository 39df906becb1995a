"""Reward functions used inside the imagination-behaviour step (reference tools/genrl_utils.py:240-370).

Only the pieces on the training hot path are mirrored: `compute_reward` and `video_text_reward` given a target sequence.
The InternVideo2 / ViCLIP text encoder is out of scope: supply the prompt embedding with
`agent.set_prompt_embedding(e)` or pre-set `agent.unconditional_target` exactly as genrl_utils.py:289-319 caches it.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .. import ops


def max_cosine_similarity(u, v, dim=-1):
    max_norm = torch.max(torch.norm(u, dim=dim), torch.norm(v, dim=dim)).unsqueeze(-1)
    return torch.sum((u / max_norm) * (v / max_norm), dim=dim)


def neg_mse_fn(a, b, dim=-1, scale=True):
    dist = -torch.norm(a - b, dim=dim)
    if scale:
        dist = dist / np.sqrt(a.shape[-1]).item()
    return dist


def _project(agent, stoch):
    """decoder._conv_in[0] applied to the flattened stoch (genrl_utils.py:255-256) on the GEMM kernel"""
    lin = agent.wm.heads["decoder"]._conv_in[0]
    lead = tuple(stoch.shape[:-1])
    y = ops.linear(ops.to_compute(stoch.reshape(-1, stoch.shape[-1])), lin.weight, lin.bias, out_fp32=True)
    return y.reshape(lead + (y.shape[-1],))


def compute_reward(agent, agent_seq, target_seq, score_fn="cosine"):
    """reference genrl_utils.py:250-277 (distance scores on the decoder's input projection)"""
    if score_fn in ["cosine", "max_cosine", "neg_mse", "exp_neg_mse"]:
        distance_fn = dict(cosine=F.cosine_similarity, max_cosine=max_cosine_similarity, neg_mse=neg_mse_fn, exp_neg_mse=neg_mse_fn)[score_fn]
        target_stoch = agent.wm.connector.get_stoch(target_seq)
        agent_stoch = agent.wm.rssm.get_stoch(agent_seq)
        conv_target = _project(agent, target_stoch)
        conv_agent = _project(agent, agent_stoch)
        reward = distance_fn(conv_target, conv_agent, dim=-1)
        if score_fn == "exp_neg_mse":
            reward = torch.exp(reward)
        return reward
    raise NotImplementedError(f"{score_fn} reward not implemented on the fused path")


def video_text_reward(agent, seq, score_fn="cosine", sample_for_target=False, weighted_align=False, align_initial=False,
                      align_sequence=False, task_prompt="", skip_first_target=False, **kwargs):
    """reference genrl_utils.py:279-370"""
    wm = agent.wm
    connector = wm.connector
    n_frames = connector.n_frames
    T, B = seq["deter"].shape[:2]
    if not hasattr(agent, "unconditional_target"):
        video_embed = getattr(agent, "prompt_embedding", None)
        if video_embed is None:
            raise RuntimeError("video_text_reward needs agent.prompt_embedding (text/video embedding of the task prompt) or a "
                               "pre-set agent.unconditional_target; the InternVideo2 encoder itself is out of scope")
        video_embed = video_embed.to(agent.device)
        with torch.no_grad():
            steps = T + 1 if skip_first_target else T
            ve = video_embed.reshape(1, 1, -1).repeat(B, steps, 1)
            stats = connector.video_imagine(ve, dreamer_init=None, sample=sample_for_target, reset_every_n_frames=False, denoise=True)
            if skip_first_target:
                stats = {k: v[:, 1:] for k, v in stats.items()}
            stats = {k: v.permute([1, 0] + list(range(2, len(v.shape)))).contiguous() for k, v in stats.items()}
        agent.unconditional_target = stats
    target_seq = agent.unconditional_target
    if align_initial:
        raise NotImplementedError("align_initial is not the GenRL default (agent/genrl.yaml:21)")
    if align_sequence:
        align_score = []
        window = lambda d, a, b: {k: v[a:b] for k, v in d.items()}
        shorter_target_seq = window(target_seq, 0, n_frames)
        for t in range(T - n_frames):
            score = compute_reward(agent, window(seq, t, t + n_frames), shorter_target_seq, score_fn=score_fn).mean(dim=0)
            align_score.append(score)
        align_score = torch.stack(align_score, dim=0)
        if weighted_align:
            w = torch.cumprod(0.99 * torch.ones_like(align_score), dim=1)
            align_score = w * align_score
        best = F.one_hot(torch.argmax(align_score, dim=0), num_classes=target_seq["stoch"].shape[0])
        ts_idx = torch.clip(torch.cumsum(torch.cumsum(best, dim=1), dim=1) - 1, min=0).T
        new_target_seq = {}
        for k, v in target_seq.items():
            if len(v.shape) == 4:
                new_ts = ts_idx.unsqueeze(-1).unsqueeze(-1).repeat(1, 1, v.shape[-2], v.shape[-1])
            else:
                new_ts = ts_idx.unsqueeze(-1).repeat(1, 1, v.shape[-1])
            new_target_seq[k] = torch.gather(v, 0, new_ts)
        return compute_reward(agent, seq, new_target_seq, score_fn=score_fn).unsqueeze(-1)
    return compute_reward(agent, seq, target_seq, score_fn=score_fn).unsqueeze(-1)
