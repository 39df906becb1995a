"""TEST INFRASTRUCTURE ONLY — CPU oracle for the GenRL/Dreamer world-model training hot path.

This is a functional restatement (torch CPU ops, fp32 or fp64, autograd for the gradients —
the reference's own arithmetic is PyTorch, requirements.txt:27) of the reference algorithm,
written against a flat parameter dict that uses the reference's state_dict key names. It is
NOT the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it. The product (genrl_b200/) never does.

Pinned against the unmodified reference by oracle/make_golden.py (run in the build
container, fixtures committed under tests/golden/) — see tests/test_oracle_golden.py.

Every function cites the reference file:line it restates (paths relative to the reference
root, commit 4cd8394).

Random draws follow the reference's order on the *global* torch CPU generator, so that
`torch.manual_seed(s)` + oracle == `torch.manual_seed(s)` + reference, draw for draw:
  * categorical sample  -> torch.multinomial(p,1) == argmax(p / Exp(1))   (dreamer_utils.py:189-197)
  * ensemble index      -> torch.randint(0, ensemble, ())                  (dreamer_utils.py:477)
  * actor sample        -> Normal.rsample == mean + std * randn            (dreamer_utils.py:814-819)
All draws can be recorded on a `NoiseTape` so the CUDA path can be fed the identical noise.
"""
import math
import os
import time

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# noise tape
# --------------------------------------------------------------------------------------
class NoiseTape:
    """Records (or replays) every random draw of one oracle call, in order."""

    def __init__(self, replay=None):
        self.items = [] if replay is None else list(replay)
        self.replay = replay is not None
        self.pos = 0

    def _next(self, kind, make):
        if self.replay:
            k, v = self.items[self.pos]
            assert k == kind, (k, kind)
            self.pos += 1
            return v
        v = make()
        self.items.append((kind, v))
        return v

    def exponential(self, like):
        # reference: torch.multinomial(probs_2d, 1, True) -> q = empty_like(p).exponential_(1)
        return self._next("exp", lambda: torch.empty_like(like).exponential_(1))

    def randn(self, like):
        return self._next("randn", lambda: torch.randn(like.shape, dtype=like.dtype))

    def randint(self, n):
        return self._next("randint", lambda: torch.randint(0, n, ()))

    def of_kind(self, kind):
        return [v for k, v in self.items if k == kind]


# --------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------
def symlog(x):   # dreamer_utils.py:13-14
    return torch.sign(x) * torch.log(torch.abs(x) + 1.0)


def symexp(x):   # dreamer_utils.py:16-17
    return torch.sign(x) * (torch.exp(torch.abs(x)) - 1.0)


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def linear_ln_silu(P, lin, norm, x, eps=1e-5):
    """Linear(bias) -> LayerNorm(eps 1e-5) -> SiLU   (NormLayer dreamer_utils.py:844-859, MLP :738-744)."""
    y = F.linear(x, P[lin + ".weight"], P.get(lin + ".bias"))
    y = layer_norm(y, P[norm + ".weight"], P[norm + ".bias"], eps)
    return F.silu(y)


def preprocess_obs(obs_u8):
    """uint8 -> float / 255 - 0.5   (dreamer.py:289-295)."""
    return obs_u8.to(torch.get_default_dtype()) / 255.0 - 0.5


# --------------------------------------------------------------------------------------
# encoder / decoder (dreamer_utils.py:558-706, ImgChLayerNorm :1031-1040 eps=1e-3)
# --------------------------------------------------------------------------------------
def encoder(P, pre, obs):
    """obs [..., 3, 64, 64] float -> embed [..., 8*depth*4]; 4x [Conv2d(k4,s2) -> LN over C -> SiLU]."""
    lead = obs.shape[:-3]
    x = obs.reshape((-1,) + tuple(obs.shape[-3:]))
    for i in range(4):
        x = F.conv2d(x, P[f"{pre}._conv_model.{3*i}.weight"], P[f"{pre}._conv_model.{3*i}.bias"], stride=2)
        x = x.permute(0, 2, 3, 1)
        x = layer_norm(x, P[f"{pre}._conv_model.{3*i+1}.norm.weight"], P[f"{pre}._conv_model.{3*i+1}.norm.bias"], 1e-3)
        x = F.silu(x.permute(0, 3, 1, 2))
    x = x.reshape(x.shape[0], -1)           # flatten order (C, H, W)  dreamer_utils.py:621
    return x.reshape(tuple(lead) + (x.shape[-1],))


def decoder(P, pre, feat):
    """feat [..., in] -> image mean [..., 3, 64, 64]   (dreamer_utils.py:695-706)."""
    lead = feat.shape[:-1]
    x = F.linear(feat, P[f"{pre}._conv_in.0.weight"], P[f"{pre}._conv_in.0.bias"])
    x = x.reshape(-1, x.shape[-1], 1, 1)
    for i in range(4):
        x = F.conv_transpose2d(x, P[f"{pre}._conv_model.{3*i}.weight"], P[f"{pre}._conv_model.{3*i}.bias"], stride=2)
        if i < 3:
            x = x.permute(0, 2, 3, 1)
            x = layer_norm(x, P[f"{pre}._conv_model.{3*i+1}.norm.weight"], P[f"{pre}._conv_model.{3*i+1}.norm.bias"], 1e-3)
            x = F.silu(x.permute(0, 3, 1, 2))
    return x.reshape(tuple(lead) + tuple(x.shape[1:]))


def mse_logprob(mean, value):
    """MSEDist.log_prob, agg=sum over all dims past the first two (dreamer_utils.py:74-83)."""
    d = (mean - value) ** 2
    return -d.sum(dim=list(range(2, d.dim())))


# --------------------------------------------------------------------------------------
# MLP heads and distributions (dreamer_utils.py:718-747, :787-841, :120-171)
# --------------------------------------------------------------------------------------
def mlp_trunk(P, pre, x, layers=4):
    for i in range(layers):
        x = linear_ln_silu(P, f"{pre}.dense{i}", f"{pre}.norm{i}._layer", x)
    return x


def mlp_twohot_logits(P, pre, x, layers=4):
    z = mlp_trunk(P, pre, x, layers)
    out = F.linear(z, P[f"{pre}._out._out.weight"], P[f"{pre}._out._out.bias"])
    return out.float() if out.dtype in (torch.float16, torch.bfloat16) else out   # DistLayer: out.float()


def twohot_buckets(dtype=torch.float32):
    return torch.linspace(-20.0, 20.0, steps=255).to(dtype)


def twohot_mean(logits):
    """TwoHotDist.mean: symexp(sum softmax * buckets), keepdim (dreamer_utils.py:139-142)."""
    probs = torch.softmax(logits, -1)
    return symexp(torch.sum(probs * twohot_buckets(logits.dtype), dim=-1, keepdim=True))


def twohot_logprob(logits, x):
    """TwoHotDist.log_prob (dreamer_utils.py:149-171). x has a trailing singleton dim."""
    buckets = twohot_buckets(logits.dtype)
    x = symlog(x)
    below = torch.sum((buckets <= x[..., None]).to(torch.int32), dim=-1) - 1
    above = len(buckets) - torch.sum((buckets > x[..., None]).to(torch.int32), dim=-1)
    below = torch.clip(below, 0, len(buckets) - 1)
    above = torch.clip(above, 0, len(buckets) - 1)
    equal = below == above
    dist_to_below = torch.where(equal, 1, torch.abs(buckets[below] - x))
    dist_to_above = torch.where(equal, 1, torch.abs(buckets[above] - x))
    total = dist_to_below + dist_to_above
    weight_below = dist_to_above / total
    weight_above = dist_to_below / total
    target = (F.one_hot(below, num_classes=255) * weight_below[..., None]
              + F.one_hot(above, num_classes=255) * weight_above[..., None])
    log_pred = logits - torch.logsumexp(logits, -1, keepdim=True)
    target = target.squeeze(-2)
    return (target * log_pred).sum(-1)


def actor_dist(P, pre, feat, layers=4):
    """DistLayer 'normal': mean=tanh(out), std=0.9*sigmoid(std+2)+0.1 (dreamer_utils.py:814-819)."""
    z = mlp_trunk(P, pre, feat, layers)
    out = F.linear(z, P[f"{pre}._out._out.weight"], P[f"{pre}._out._out.bias"])
    std = F.linear(z, P[f"{pre}._out._std.weight"], P[f"{pre}._out._std.bias"])
    mean = torch.tanh(out)
    std = (1.0 - 0.1) * torch.sigmoid(std + 2.0) + 0.1
    return mean, std


def normal_entropy(std):
    """Independent(Normal).entropy(): sum_A (0.5 + 0.5 log 2pi + log std)."""
    return (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std)).sum(-1)


# --------------------------------------------------------------------------------------
# categorical latent (OneHotDist dreamer_utils.py:177-197, get_dist :410-416)
# --------------------------------------------------------------------------------------
def unimix_probs(logit, unif_mix=0.99):
    """probs the reference's OneHotDist ends up holding: normalise(0.99 softmax + 0.01/K)."""
    logit = logit.float() if logit.dtype not in (torch.float32, torch.float64) else logit
    p = torch.softmax(logit, -1)                      # Categorical(logits=...).probs
    p = unif_mix * p + (1 - unif_mix) * torch.ones_like(p) / p.shape[-1]
    return p / p.sum(-1, keepdim=True)                # Categorical(probs=...) normalises


def sample_onehot(logit, tape, sample=True):
    """OneHotDist.sample with straight-through gradients; returns (stoch, index)."""
    probs = unimix_probs(logit)
    K = probs.shape[-1]
    if sample:
        p2 = probs.detach().reshape(-1, K)
        q = tape.exponential(p2)
        idx = torch.argmax(p2 / q, dim=-1).reshape(probs.shape[:-1])
        onehot = F.one_hot(idx, K).to(probs.dtype)
        return onehot + (probs - probs.detach()), idx
    logits_n = torch.log(probs)                       # super().logits of the probs-constructed dist
    idx = torch.argmax(logits_n, dim=-1)
    onehot = F.one_hot(idx, K)
    return onehot.detach() + logits_n - logits_n.detach(), idx


def cat_entropy(logit):
    """Independent(OneHotDist(logit),1).entropy(): sum over the S latents."""
    p = unimix_probs(logit)
    logp = torch.log(p)
    ent = -(torch.clamp(logp, min=torch.finfo(logp.dtype).min) * p).sum(-1)
    return ent.sum(-1)


def cat_kl(logit_p, logit_q):
    """kl_divergence(Independent(OneHotDist(p)), Independent(OneHotDist(q))) summed over S.

    torch's _kl_categorical_categorical: t = p * (log p - log q); t[q==0]=inf; t[p==0]=0; sum(-1).
    With the 1 % uniform mix neither special case can trigger.
    """
    p = unimix_probs(logit_p)
    q = unimix_probs(logit_q)
    t = p * (torch.log(p) - torch.log(q))
    return t.sum(-1).sum(-1)


def kl_loss(post_logit, prior_logit, forward, balance, free, free_avg=False):
    """EnsembleRSSM.kl_loss (dreamer_utils.py:534-555). Returns (loss, value)."""
    lhs, rhs = (prior_logit, post_logit) if forward else (post_logit, prior_logit)
    mix = balance if forward else (1 - balance)
    free_t = torch.tensor([free], dtype=post_logit.dtype)
    if balance == 0.5:
        value = cat_kl(lhs, rhs)
        loss = torch.maximum(value, free_t).mean()
    else:
        value_lhs = value = cat_kl(lhs, rhs.detach())
        value_rhs = cat_kl(lhs.detach(), rhs)
        if free_avg:
            loss_lhs = torch.maximum(value_lhs.mean(), free_t)
            loss_rhs = torch.maximum(value_rhs.mean(), free_t)
        else:
            loss_lhs = torch.maximum(value_lhs, free_t).mean()
            loss_rhs = torch.maximum(value_rhs, free_t).mean()
        loss = mix * loss_lhs + (1 - mix) * loss_rhs
    return loss, value


# --------------------------------------------------------------------------------------
# RSSM (EnsembleRSSM dreamer_utils.py:302-521, GRUCell :750-785)
# --------------------------------------------------------------------------------------
def gru_cell(P, pre, x, h):
    """GRUCell.forward with norm=True, update_bias=-1, act=tanh, no bias (dreamer_utils.py:771-785)."""
    parts = F.linear(torch.cat([x, h], -1), P[f"{pre}._cell._layer.weight"])
    parts = layer_norm(parts, P[f"{pre}._cell._norm.weight"], P[f"{pre}._cell._norm.bias"], 1e-5)
    reset, cand, update = torch.chunk(parts, 3, -1)
    reset = torch.sigmoid(reset)
    cand = torch.tanh(reset * cand)
    update = torch.sigmoid(update - 1.0)
    return update * cand + (1 - update) * h


def prior_logit_from_deter(P, pre, deter, S, K):
    """_suff_stats_ensemble with ensemble=1 (dreamer_utils.py:489-521)."""
    y = linear_ln_silu(P, f"{pre}._ensemble_img_out.0.0", f"{pre}._ensemble_img_out.0.1._layer", deter)
    logit = F.linear(y, P[f"{pre}._ensemble_img_dist.0.weight"], P[f"{pre}._ensemble_img_dist.0.bias"])
    return logit.reshape(tuple(logit.shape[:-1]) + (S, K))


def img_step(P, pre, prev, action, tape, S, K, sample=True):
    """EnsembleRSSM.img_step (dreamer_utils.py:459-487), cell_input='stoch', ensemble=1."""
    stoch_flat = prev["stoch"].reshape(tuple(prev["stoch"].shape[:-2]) + (S * K,))
    x = torch.cat([stoch_flat, action], -1)
    x = linear_ln_silu(P, f"{pre}._img_in.0", f"{pre}._img_in.1._layer", x)
    deter = gru_cell(P, pre, x, prev["deter"])
    logit = prior_logit_from_deter(P, pre, deter, S, K)
    tape.randint(1)                                        # torch.randint(0, ensemble, ()) :477
    stoch, _ = sample_onehot(logit, tape, sample)
    return {"stoch": stoch, "deter": deter, "logit": logit}


def post_logit_from_embed(P, pre, embed, deter, S, K, single_obs_posterior=True):
    """get_post_stoch up to the logits (dreamer_utils.py:442-453)."""
    x = embed if single_obs_posterior else torch.cat([deter, embed], -1)
    z = linear_ln_silu(P, f"{pre}._obs_out.0", f"{pre}._obs_out.1._layer", x)
    logit = F.linear(z, P[f"{pre}._obs_dist.weight"], P[f"{pre}._obs_dist.bias"])
    return logit.reshape(tuple(logit.shape[:-1]) + (S, K))


def obs_step(P, pre, prev, prev_action, embed, is_first, tape, S, K, single_obs_posterior=True, sample=True):
    """EnsembleRSSM.obs_step (dreamer_utils.py:432-440)."""
    m = 1.0 - is_first.to(embed.dtype)
    prev = {k: torch.einsum("b,b...->b...", m, v) for k, v in prev.items()}
    prev_action = torch.einsum("b,b...->b...", m, prev_action)
    prior = img_step(P, pre, prev, prev_action, tape, S, K, sample)
    logit = post_logit_from_embed(P, pre, embed, prior["deter"], S, K, single_obs_posterior)
    stoch, _ = sample_onehot(logit, tape, sample)
    post = {"stoch": stoch, "deter": prior["deter"], "logit": logit}
    return post, prior


def rssm_initial(B, S, K, D, dtype):
    """EnsembleRSSM.initial, discrete (dreamer_utils.py:348-360)."""
    return {"logit": torch.zeros(B, S, K, dtype=dtype), "stoch": torch.zeros(B, S, K, dtype=dtype),
            "deter": torch.zeros(B, D, dtype=dtype)}


def observe(P, pre, embed, action, is_first, tape, S, K, D, state=None, single_obs_posterior=True):
    """EnsembleRSSM.observe (dreamer_utils.py:362-371): loop obs_step over T; returns [B,T,...] dicts."""
    B, T = action.shape[:2]
    if state is None:
        state = rssm_initial(B, S, K, D, embed.dtype)
    posts, priors = [], []
    prev = state
    for t in range(T):
        post, prior = obs_step(P, pre, prev, action[:, t], embed[:, t], is_first[:, t], tape, S, K,
                               single_obs_posterior)
        posts.append(post)
        priors.append(prior)
        prev = post
    stack = lambda lst: {k: torch.stack([d[k] for d in lst], 1) for k in lst[0]}
    return stack(posts), stack(priors)


def imagine_prior(P, pre, action, state, tape, S, K, sample=True):
    """EnsembleRSSM.imagine (dreamer_utils.py:373-382): open-loop prior rollout over given actions."""
    priors = []
    prev = state
    for t in range(action.shape[1]):
        prev = img_step(P, pre, prev, action[:, t], tape, S, K, sample)
        priors.append(prev)
    return {k: torch.stack([d[k] for d in priors], 1) for k in priors[0]}


def get_feat(state):
    """[stoch_flat | deter] (dreamer_utils.py:405-408)."""
    st = state["stoch"]
    return torch.cat([st.reshape(tuple(st.shape[:-2]) + (-1,)), state["deter"]], -1)


# --------------------------------------------------------------------------------------
# world-model loss (WorldModel.loss dreamer.py:219-252)
# --------------------------------------------------------------------------------------
def wm_loss(P, data, cfg, tape, grad_heads=("decoder",)):
    """Returns (model_loss, last_state, outs, metrics) exactly as WorldModel.loss does."""
    r = cfg["rssm"]
    S, K, D = r["stoch"], r["discrete"], r["deter"]
    obs = preprocess_obs(data["observation"])
    embed = encoder(P, "wm.encoder", obs)
    post, prior = observe(P, "wm.rssm", embed, data["action"], data["is_first"], tape, S, K, D,
                          single_obs_posterior=r["single_obs_posterior"])
    kl, kl_value = kl_loss(post["logit"], prior["logit"], **cfg["kl"])
    feat = get_feat(post)
    likes, losses = {}, {"kl": kl}
    dec_in = feat if cfg["decoder_inputs"] == "feat" else post["stoch"].reshape(post["stoch"].shape[:2] + (S * K,))
    if "decoder" not in grad_heads:
        dec_in = dec_in.detach()
    mean = decoder(P, "wm.heads.decoder", dec_in)
    likes["observation"] = mse_logprob(mean, obs)
    losses["observation"] = -likes["observation"].mean()
    rew_in = feat if "reward" in grad_heads else feat.detach()
    rew_logits = mlp_twohot_logits(P, "wm.heads.reward", rew_in, cfg["reward_head"]["layers"])
    likes["reward"] = twohot_logprob(rew_logits, data["reward"])
    losses["reward"] = -likes["reward"].mean()
    model_loss = sum(cfg["loss_scales"].get(k, 1.0) * v for k, v in losses.items())
    outs = dict(embed=embed, feat=feat, post=post, prior=prior, likes=likes, kl=kl_value, recon=mean)
    metrics = {f"{k}_loss": v for k, v in losses.items()}
    metrics["model_kl"] = kl_value.mean()
    metrics["prior_ent"] = cat_entropy(prior["logit"]).mean()
    metrics["post_ent"] = cat_entropy(post["logit"]).mean()
    last = {k: v[:, -1] for k, v in post.items()}
    return model_loss, last, outs, metrics


# --------------------------------------------------------------------------------------
# connector / aligner (agent/video_utils.py)
# --------------------------------------------------------------------------------------
def residual_linear(P, pre, x, norm=True, act=True):
    """ResidualLinear, prenorm=False (video_utils.py:8-25)."""
    h = F.linear(x, P[f"{pre}.layer.weight"], P[f"{pre}.layer.bias"])
    if norm:
        h = layer_norm(h, P[f"{pre}.norm_layer._layer.weight"], P[f"{pre}.norm_layer._layer.bias"], 1e-5)
    if act:
        h = F.silu(h)
    if f"{pre}.res_proj.weight" in P:
        res = F.linear(x, P[f"{pre}.res_proj.weight"], P[f"{pre}.res_proj.bias"])
    else:
        res = x
    return h + res


def aligner(P, pre, x):
    """UNetDenoiser with n_layers=2 (video_utils.py:27-61)."""
    d0 = residual_linear(P, f"{pre}.down.0", x)
    d1 = residual_linear(P, f"{pre}.down.1", d0)
    m = residual_linear(P, f"{pre}.mid.0", d1)
    m = residual_linear(P, f"{pre}.mid.1", m)
    u = residual_linear(P, f"{pre}.up.0", torch.cat([m, d1], -1), norm=False, act=False)
    u = residual_linear(P, f"{pre}.up.1", torch.cat([u, d0], -1))
    return u


def connector_get_action(video_embed, n_frames=8, rescale=True):
    """VideoSSM.get_action with temporal_embeds=False (video_utils.py:114-125)."""
    E = video_embed.shape[-1]
    if rescale:
        video_embed = video_embed * np.sqrt(E).item()
    zeros = torch.zeros(tuple(video_embed.shape[:-1]) + (n_frames,), dtype=video_embed.dtype)
    return torch.cat([video_embed, zeros], -1)


def connector_initial(P, pre, init_embed, tape, S, K):
    """VideoSSM.initial with learn_initial (video_utils.py:100-112)."""
    x = linear_ln_silu(P, f"{pre}.initial_state_pred.0", f"{pre}.initial_state_pred.1._layer", init_embed)
    x = linear_ln_silu(P, f"{pre}.initial_state_pred.3", f"{pre}.initial_state_pred.4._layer", x)
    deter = F.linear(x, P[f"{pre}.initial_state_pred.6.weight"], P[f"{pre}.initial_state_pred.6.bias"])
    logit = prior_logit_from_deter(P, pre, deter, S, K)
    tape.randint(1)
    stoch, _ = sample_onehot(logit, tape, True)
    return {"logit": logit, "stoch": stoch, "deter": deter}


def connector_update(P, clip_video, wm_post, cfg, tape, pre="wm.connector", n_frames=8):
    """connector_update_fn + VideoSSM.update (genrl.py:7-25, video_utils.py:127-207). Returns (loss, metrics)."""
    r = cfg["connector_rssm"]
    S, K = r["stoch"], r["discrete"]
    B, T = clip_video.shape[:2]
    ve = clip_video[:, n_frames - 1::n_frames]
    ve = ve.reshape(B, T // n_frames, 1, -1).repeat(1, 1, n_frames, 1).reshape(B, T, -1)
    orig = ve
    loss = 0
    metrics = {}
    if cfg["clip_add_noise"] > 0:
        ve = ve + tape.randn(ve) * cfg["clip_add_noise"]
        ve = F.normalize(ve, dim=-1)
    if cfg["clip_lafite_noise"] > 0:
        nn_ = F.normalize(tape.randn(ve), dim=-1)
        ve = (1 - cfg["clip_lafite_noise"]) * ve + cfg["clip_lafite_noise"] * nn_
        ve = F.normalize(ve, dim=-1)
    den = F.normalize(aligner(P, f"{pre}.aligner", ve), dim=-1)
    den_loss = 1 - F.cosine_similarity(den, orig, dim=-1).mean()
    loss = loss + den_loss
    metrics["aligner_cosine_distance"] = den_loss
    ve = orig
    acts = connector_get_action(ve, n_frames, cfg["connector"]["rescale_embeds"])
    wm_post = {k: v.detach() for k, v in wm_post.items()}
    priors = []
    prior = None
    for t in range(T):
        a = acts[:, t]
        if t == 0:
            prev = connector_initial(P, pre, a, tape, S, K)
        else:
            prev = dict(prior)
            prev["stoch"] = wm_post["stoch"][:, t - 1]
        prior = img_step(P, pre, prev, a, tape, S, K)
        priors.append(prior)
    prior_seq = {k: torch.stack([d[k] for d in priors], 1) for k in priors[0]}
    kl, kl_value = kl_loss(wm_post["logit"], prior_seq["logit"], **cfg["connector_kl"])
    metrics["connector_kl"] = kl_value.mean()
    loss = loss + cfg["connector"]["loss_scale"] * kl
    # initial-KL metric pass (video_utils.py:197-205)
    ve2 = ve.reshape(B, T // n_frames, n_frames, -1)[:, 1:, 0].reshape(B * (T // n_frames - 1), 1, -1)
    acts2 = connector_get_action(ve2, n_frames, cfg["connector"]["rescale_embeds"])
    post2 = {k: v.reshape((B, T // n_frames, n_frames) + tuple(v.shape[2:]))[:, 1:, 0]
             .reshape((B * (T // n_frames - 1),) + tuple(v.shape[2:])) for k, v in wm_post.items()}
    a0 = acts2[:, 0]
    prev = connector_initial(P, pre, a0, tape, S, K)
    prior0 = img_step(P, pre, prev, a0, tape, S, K)
    _, kv = kl_loss(post2["logit"], prior0["logit"], **cfg["connector_kl"])
    metrics["connector_initial_kl"] = kv.mean()
    return loss, metrics


# --------------------------------------------------------------------------------------
# imagination + actor-critic (dreamer.py:254-287, :366-462; dreamer_utils.py:228-253, :1014-1029)
# --------------------------------------------------------------------------------------
def imagine(P, actor_pre, start, horizon, cfg, tape):
    """WorldModel.imagine without discount head / task_cond. start: dict of [B,T,...] (detached)."""
    r = cfg["rssm"]
    S, K = r["stoch"], r["discrete"]
    flat = lambda x: x.reshape((-1,) + tuple(x.shape[2:]))
    start = {k: flat(v) for k, v in start.items()}
    start["feat"] = get_feat(start)
    mean, std = actor_dist(P, actor_pre, start["feat"], cfg["actor"]["layers"])
    tape.randn(mean)                                   # policy_dist.sample() consumed, zeros_like'd (dreamer.py:259-260)
    start["action"] = torch.zeros_like(mean)
    seq = {k: [v] for k, v in start.items()}
    for _ in range(horizon):
        inp = seq["feat"][-1].detach()
        mean, std = actor_dist(P, actor_pre, inp, cfg["actor"]["layers"])
        action = mean + std * tape.randn(mean)
        state = img_step(P, "wm.rssm", {k: v[-1] for k, v in seq.items() if k in ("stoch", "deter", "logit")},
                         action, tape, S, K)
        feat = get_feat(state)
        for key, value in {**state, "action": action, "feat": feat}.items():
            seq[key].append(value)
    seq = {k: torch.stack(v, 0) for k, v in seq.items()}
    disc = torch.ones(tuple(seq["feat"].shape[:-1]) + (1,), dtype=seq["feat"].dtype)
    seq["discount"] = disc * cfg["discount"]
    seq["weight"] = torch.cumprod(torch.cat([torch.ones_like(disc[:1]), disc[:-1]], 0), 0)
    return seq


def lambda_return(reward, value, pcont, bootstrap, lambda_):
    """lambda_return along axis 0 (dreamer_utils.py:228-253)."""
    next_values = torch.cat([value[1:], bootstrap[None]], 0)
    inputs = reward + pcont * next_values * (1 - lambda_)
    last = bootstrap
    outs = []
    for t in reversed(range(reward.shape[0])):
        last = inputs[t] + pcont[t] * lambda_ * last
        outs.append(last)
    return torch.stack(list(reversed(outs)), 0)


def reward_ema_update(target, ema_vals, alpha=1e-2):
    """RewardEMA.__call__ (dreamer_utils.py:1022-1029). Updates ema_vals in place."""
    flat = torch.flatten(target.detach())
    q = torch.quantile(flat, torch.tensor([0.05, 0.95], dtype=flat.dtype))
    ema_vals[:] = alpha * q + (1 - alpha) * ema_vals
    scale = torch.clip(ema_vals[1] - ema_vals[0], min=1.0)
    return ema_vals[0].detach().clone(), scale.detach().clone()


def env_reward(P, seq, cfg):
    """env_reward: reward-head two-hot mean (dreamer.py:16-17)."""
    return twohot_mean(mlp_twohot_logits(P, "wm.heads.reward", seq["feat"], cfg["reward_head"]["layers"]))


def behavior_losses(P, beh_pre, start, cfg, tape, ema_vals, reward_fn=None, actor_ent=None):
    """ActorCritic.update forward part (dreamer.py:366-453). Returns (actor_loss, critic_loss_fn, seq, metrics)."""
    H = cfg["imag_horizon"]
    seq = imagine(P, f"{beh_pre}.actor", start, H, cfg, tape)
    reward = env_reward(P, seq, cfg) if reward_fn is None else reward_fn(seq)
    seq["reward"] = reward                                               # StreamNorm momentum=1 -> identity
    mets = {"reward_mean": reward.mean(), "reward_std": reward.std(),
            "reward_normed_mean": reward.mean(), "reward_normed_std": reward.std()}
    value = twohot_mean(mlp_twohot_logits(P, f"{beh_pre}._target_critic", seq["feat"], cfg["critic"]["layers"]))
    target = lambda_return(reward[:-1], value[:-1], seq["discount"][:-1], value[-1], cfg["discount_lambda"])
    mets["critic_slow"] = value.mean()
    mets["critic_target"] = target.mean()
    # actor loss (dreamer.py:392-429), actor_grad == 'dynamics'
    mean, std = actor_dist(P, f"{beh_pre}.actor", seq["feat"][:-2].detach(), cfg["actor"]["layers"])
    offset, scale = reward_ema_update(target, ema_vals)
    normed_target = (target - offset) / scale
    mets["normed_target_mean"] = normed_target.mean()
    mets["normed_target_std"] = normed_target.std()
    mets["reward_ema_005"] = ema_vals[0].clone()
    mets["reward_ema_095"] = ema_vals[1].clone()
    objective = normed_target[1:]
    ent = normal_entropy(std)[:, :, None]
    ent_scale = cfg["actor_ent"] if actor_ent is None else actor_ent
    objective = objective + ent_scale * ent
    mets["actor_ent"] = ent.mean()
    mets["actor_ent_scale"] = ent_scale
    weight = seq["weight"].detach()
    actor_loss = -(weight[:-2] * objective).mean()

    def critic_loss_fn():
        feat = seq["feat"][:-1].detach()
        logits = mlp_twohot_logits(P, f"{beh_pre}.critic", feat, cfg["critic"]["layers"])
        loss = -(twohot_logprob(logits, target.detach())[:, :, None] * weight[:-1]).mean()
        return loss, {"critic": twohot_mean(logits).mean()}

    return actor_loss, critic_loss_fn, seq, mets, target


# --------------------------------------------------------------------------------------
# VLM reward of the imagination behaviour (tools/genrl_utils.py:240-370, agent/genrl.yaml:21; agent/video_utils.py:209-239)
# --------------------------------------------------------------------------------------
def connector_video_imagine(P, video_embed, cfg, tape, sample=True, denoise=False, dreamer_init=None, pre="wm.connector", n_frames=8):
    """VideoSSM.video_imagine with reset_every_n_frames=False (video_utils.py:209-239): optional denoising of the embeddings by the
    aligner, learned initial state from the first action (its stoch replaced by dreamer_init's when given: cell_input == 'stoch'),
    open-loop prior rollout over the embedding actions."""
    r = cfg["connector_rssm"]
    S, K = r["stoch"], r["discrete"]
    if denoise and cfg["connector"]["denoising_ae"]:
        video_embed = F.normalize(aligner(P, f"{pre}.aligner", video_embed), dim=-1)
    action = connector_get_action(video_embed, n_frames, cfg["connector"]["rescale_embeds"])
    init = connector_initial(P, pre, action[:, 0], tape, S, K)
    if dreamer_init is not None:
        init["stoch"] = dreamer_init["stoch"]
    return imagine_prior(P, pre, action, init, tape, S, K, sample)


def unconditional_target(P, prompt_embed, n_rows, steps, cfg, tape, skip_first=True, sample=False):
    """The cached target of video_text_reward (genrl_utils.py:305-316): the connector's rollout from the prompt embedding repeated over
    time, denoised, first step dropped when skip_first_target, returned time-major [steps, n_rows, ...]."""
    ve = prompt_embed.reshape(1, 1, -1).repeat(n_rows, steps + (1 if skip_first else 0), 1)
    stats = connector_video_imagine(P, ve, cfg, tape, sample=sample, denoise=True)
    if skip_first:
        stats = {k: v[:, 1:] for k, v in stats.items()}
    return {k: v.transpose(0, 1).contiguous() for k, v in stats.items()}


def max_cosine(u, v):
    """genrl_utils.py:240-242: dot product of the two vectors after dividing both by the larger of their norms."""
    m = torch.maximum(u.norm(dim=-1), v.norm(dim=-1)).unsqueeze(-1)
    return ((u / m) * (v / m)).sum(-1)


def vlm_reward(P, seq, target, n_frames=8, align_sequence=True):
    """video_text_reward with score_fn='max_cosine' on the decoder's input projection of the flattened stoch (compute_reward,
    genrl_utils.py:250-258).  align_sequence (:344-366): score every n_frames window of the agent sequence against the first n_frames of
    the target, take the best window start t* per row, and compare the agent sequence with the target shifted to start at t*
    (target index clip(cumsum(cumsum(one_hot(t*))) - 1, 0), i.e. 0 before t* and counting up from it)."""
    W, b = P["wm.heads.decoder._conv_in.0.weight"], P["wm.heads.decoder._conv_in.0.bias"]
    proj = lambda st: F.linear(st.reshape(tuple(st.shape[:-2]) + (-1,)), W, b)
    score = lambda agent_stoch, target_stoch: max_cosine(proj(target_stoch), proj(agent_stoch))
    if not align_sequence:
        return score(seq["stoch"], target["stoch"]).unsqueeze(-1)
    T = seq["deter"].shape[0]
    head = target["stoch"][:n_frames]
    window_scores = torch.stack([score(seq["stoch"][t:t + n_frames], head).mean(0) for t in range(T - n_frames)], 0)     # [T-nf, N]
    best = F.one_hot(window_scores.argmax(0), num_classes=target["stoch"].shape[0])                                     # [N, T_target]
    idx = torch.clip(torch.cumsum(torch.cumsum(best, 1), 1) - 1, min=0).T                                                # [T_target, N]
    ts = target["stoch"]
    shifted = torch.gather(ts, 0, idx.reshape(idx.shape + (1, 1)).expand(idx.shape + tuple(ts.shape[2:])))
    return score(seq["stoch"], shifted).unsqueeze(-1)


# --------------------------------------------------------------------------------------
# report(): open-loop video prediction (dreamer.py:98-108, :307-321) and text/video-conditioned prediction (genrl.py:64-106)
# --------------------------------------------------------------------------------------
def video_pred(P, data, cfg, tape, nvid=8, ctx=5):
    """WorldModel.video_pred: reconstruct the first `ctx` frames from the posterior, imagine the rest open loop from the recorded actions,
    return [truth | model | error] stacked along the image height, [nvid, T, 3, 3*64, 64]."""
    r = cfg["rssm"]
    S, K, D = r["stoch"], r["discrete"], r["deter"]
    obs = preprocess_obs(data["observation"])
    truth = obs[:nvid] + 0.5
    embed = encoder(P, "wm.encoder", obs)
    states, _ = observe(P, "wm.rssm", embed[:nvid, :ctx], data["action"][:nvid, :ctx], data["is_first"][:nvid, :ctx], tape, S, K, D,
                        single_obs_posterior=r["single_obs_posterior"])
    dec_in = lambda st: st["stoch"].reshape(tuple(st["stoch"].shape[:-2]) + (-1,))          # decoder_inputs: stoch (dreamer.py:138-140)
    recon = decoder(P, "wm.heads.decoder", dec_in(states))
    init = {k: v[:, -1] for k, v in states.items()}
    prior = imagine_prior(P, "wm.rssm", data["action"][:nvid, ctx:], init, tape, S, K)
    prior_recon = decoder(P, "wm.heads.decoder", dec_in(prior))
    model = torch.clip(torch.cat([recon[:, :ctx] + 0.5, prior_recon + 0.5], 1), 0, 1)
    return torch.cat([truth, model, (model - truth + 1) / 2], 3)


def video_clip_pred(P, data, cfg, tape, nvid=8, n_frames=8):
    """GenRLAgent.report (viclip_encode: true): posterior reconstruction of the first n_frames, then the connector's rollout from the
    batch's video embeddings (one per n_frames chunk, held over the chunk) starting at the last posterior stoch, decoded."""
    r = cfg["rssm"]
    S, K, D = r["stoch"], r["discrete"], r["deter"]
    obs = preprocess_obs(data["observation"])
    truth = obs[:nvid] + 0.5
    B, T = truth.shape[0], truth.shape[1] - n_frames
    embed = encoder(P, "wm.encoder", obs)
    states, _ = observe(P, "wm.rssm", embed[:nvid, :n_frames], data["action"][:nvid, :n_frames], data["is_first"][:nvid, :n_frames], tape,
                        S, K, D, single_obs_posterior=r["single_obs_posterior"])
    dec_in = lambda st: st["stoch"].reshape(tuple(st["stoch"].shape[:-2]) + (-1,))
    recon = decoder(P, "wm.heads.decoder", dec_in(states))
    last = {k: v[:, -1] for k, v in states.items()}
    ve = data["clip_video"][:nvid, 2 * n_frames - 1::n_frames]
    ve = ve.reshape(B, T // n_frames, 1, -1).repeat(1, 1, n_frames, 1).reshape(B, T, -1)
    prior = connector_video_imagine(P, ve, cfg, tape, sample=True, denoise=False, dreamer_init=last)
    prior_recon = decoder(P, "wm.heads.decoder", dec_in(prior))
    model = torch.clip(torch.cat([recon[:, :n_frames] + 0.5, prior_recon + 0.5], 1), 0, 1)
    return torch.cat([truth, model, (model - truth + 1) / 2], 3)


# --------------------------------------------------------------------------------------
# acting (DreamerAgent.act, dreamer.py:41-64): one encoder + obs_step + actor evaluation per environment step, batch 1
# --------------------------------------------------------------------------------------
def act_step(P, obs, state, cfg, tape, eval_mode, act_dim):
    """obs: dict of un-batched numpy/torch values (observation uint8 [3,64,64], is_first bool, ...); state: None or (latent, action).
    Returns (action [act_dim], new_state).  The latent is always sampled (eval_state_mean: False); the action is the policy mean in eval
    mode and a reparameterised sample otherwise."""
    r = cfg["rssm"]
    S, K, D = r["stoch"], r["discrete"], r["deter"]
    ob = preprocess_obs(torch.as_tensor(obs["observation"]).unsqueeze(0))
    is_first = torch.as_tensor(obs["is_first"]).reshape(1)
    if state is None:
        latent, action = rssm_initial(1, S, K, D, ob.dtype), torch.zeros(1, act_dim, dtype=ob.dtype)
    else:
        latent, action = state
    embed = encoder(P, "wm.encoder", ob)
    should_sample = (not eval_mode) or (not cfg["eval_state_mean"])
    latent, _ = obs_step(P, "wm.rssm", latent, action, embed, is_first, tape, S, K, r["single_obs_posterior"], should_sample)
    mean, std = actor_dist(P, "_acting_behavior.actor", get_feat(latent), cfg["actor"]["layers"])
    action = mean if eval_mode else mean + std * tape.randn(mean)
    return action[0], (latent, action)


# --------------------------------------------------------------------------------------
# optimizer (Optimizer.__call__ dreamer_utils.py:892-932 around torch.optim.Adam)
# --------------------------------------------------------------------------------------
class AdamState:
    def __init__(self, lr, eps, clip, wd):
        self.lr, self.eps, self.clip, self.wd = lr, eps, clip, wd
        self.m, self.v, self.step = {}, {}, {}


def optimizer_step(P, names, grads, st: AdamState):
    """clip_grad_norm_ -> multiplicative weight decay on every passed param -> Adam on params with grad.

    P: dict name->leaf tensor (updated in place), names: params handed to the optimizer call,
    grads: dict name->grad (only names that have a gradient). Returns grad norm (before clipping).
    """
    gl = [grads[n] for n in names if n in grads]
    norm = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in gl])) if gl else torch.tensor(0.0)
    coef = torch.clamp(st.clip / (norm + 1e-6), max=1.0)
    with torch.no_grad():
        for n in names:
            P[n].mul_(1 - st.wd)
        for n in names:
            if n not in grads:
                continue
            g = grads[n] * coef
            if n not in st.m:
                st.m[n] = torch.zeros_like(P[n]); st.v[n] = torch.zeros_like(P[n]); st.step[n] = 0
            st.step[n] += 1
            t = st.step[n]
            st.m[n].lerp_(g, 1 - 0.9)
            st.v[n].mul_(0.999).addcmul_(g, g, value=1 - 0.999)
            bc1 = 1 - 0.9 ** t
            bc2 = 1 - 0.999 ** t
            step_size = st.lr / bc1
            denom = (st.v[n].sqrt() / math.sqrt(bc2)).add_(st.eps)
            P[n].addcdiv_(st.m[n], denom, value=-step_size)
    return norm


# --------------------------------------------------------------------------------------
# whole-agent oracle: GenRLAgent.update = update_wm (+ in-update connector step) + acting behaviour
# --------------------------------------------------------------------------------------
class OracleAgent:
    """Functional twin of reference GenRLAgent/DreamerAgent holding a flat {state_dict key: tensor}."""

    def __init__(self, params, cfg, kind="genrl", dtype=torch.float32):
        self.cfg = cfg
        self.kind = kind
        self.dtype = dtype
        self.P = {k: v.detach().clone().to(dtype) if v.is_floating_point() else v.clone() for k, v in params.items()}
        self.P.setdefault("_acting_behavior.ema_vals", torch.zeros(2, dtype=dtype))
        self.ema = {"_acting_behavior": self.P["_acting_behavior.ema_vals"]}      # aliases: updated in place
        if "_imag_behavior.ema_vals" in self.P:
            self.ema["_imag_behavior"] = self.P["_imag_behavior.ema_vals"]
        mo, ao, co = cfg["model_opt"], cfg["actor_opt"], cfg["critic_opt"]
        self.model_opt = AdamState(mo["lr"], mo["eps"], mo["clip"], mo["wd"])
        self.beh_opts = {}
        for b in ("_acting_behavior", "_imag_behavior"):
            self.beh_opts[b] = (AdamState(ao["lr"], ao["eps"], ao["clip"], ao["wd"]),
                                AdamState(co["lr"], co["eps"], co["clip"], co["wd"]))
        self.updates = {"_acting_behavior": 0, "_imag_behavior": 0}
        self.tape = NoiseTape()

    # parameter groups exactly as nn.Module.parameters() would enumerate them
    def names(self, prefix, exclude=()):
        return [k for k in self.P if k.startswith(prefix) and not k.endswith("ema_vals")
                and not any(k.startswith(e) for e in exclude)]

    def _grad(self, loss, names):
        leaves = [self.P[n] for n in names]
        for p in leaves:
            p.requires_grad_(True)
        gs = torch.autograd.grad(loss, leaves, allow_unused=True)
        for p in leaves:
            p.requires_grad_(False)
        return {n: g for n, g in zip(names, gs) if g is not None}

    def _with_grad(self, names):
        for n in names:
            self.P[n].requires_grad_(True)

    def _no_grad(self):
        for p in self.P.values():
            if p.is_floating_point():
                p.requires_grad_(False)

    def cast(self, data):
        return {k: (v.to(self.dtype) if v.is_floating_point() else v) for k, v in data.items()}

    def update_wm(self, data):
        """DreamerAgent.update_wm -> WorldModel.update (dreamer.py:66-71, :166-187)."""
        data = self.cast(data)
        wm_names = self.names("wm.")
        self._with_grad(wm_names)
        loss, state, outs, mets = wm_loss(self.P, data, self.cfg, self.tape, tuple(self.cfg["grad_heads"]))
        gs = torch.autograd.grad(loss, [self.P[n] for n in wm_names], allow_unused=True)
        self._no_grad()
        grads = {n: g for n, g in zip(wm_names, gs) if g is not None}
        mets = {k: v.detach() for k, v in mets.items()}
        mets["model_loss"] = loss.detach()
        mets["model_grad_norm"] = optimizer_step(self.P, wm_names, grads, self.model_opt)
        outs = {k: ({kk: vv.detach() for kk, vv in v.items()} if isinstance(v, dict) else v.detach())
                for k, v in outs.items()}
        outs["is_terminal"] = data["is_terminal"]
        self.last_wm_grads = grads
        if self.kind == "genrl":
            _, mets = self.update_connector(data, outs, mets)
        state = {k: v.detach() for k, v in state.items()}
        return state, outs, mets

    def update_connector(self, data, outs, mets):
        """WorldModel.update_additional_detached_modules (dreamer.py:189-200)."""
        data = self.cast(data)
        names = self.names("wm.connector.")
        self._with_grad(names)
        loss, m = connector_update(self.P, data["clip_video"], outs["post"], self.cfg, self.tape)
        gs = torch.autograd.grad(loss, [self.P[n] for n in names], allow_unused=True)
        self._no_grad()
        grads = {n: g for n, g in zip(names, gs) if g is not None}
        mets.update({k: v.detach() for k, v in m.items()})
        mets["connector_model_loss"] = loss.detach()
        mets["connector_model_grad_norm"] = optimizer_step(self.P, names, grads, self.model_opt)
        self.last_connector_grads = grads
        return loss.detach(), mets

    def update_behavior(self, outs, beh="_acting_behavior", reward_fn=None, prefix=""):
        """ActorCritic.update (dreamer.py:366-390)."""
        start = {k: v.detach() for k, v in outs["post"].items()}
        a_names = self.names(f"{beh}.actor.")
        c_names = self.names(f"{beh}.critic.")
        aopt, copt = self.beh_opts[beh]
        self._with_grad(a_names)
        actor_loss, critic_fn, seq, mets, target = behavior_losses(
            self.P, beh, start, self.cfg, self.tape, self.ema[beh], reward_fn)
        gs = torch.autograd.grad(actor_loss, [self.P[n] for n in a_names], allow_unused=True)
        self._no_grad()
        agr = {n: g for n, g in zip(a_names, gs) if g is not None}
        out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in mets.items()}
        out["actor_loss"] = actor_loss.detach()
        out["actor_grad_norm"] = optimizer_step(self.P, a_names, agr, aopt)
        seq = {k: v.detach() for k, v in seq.items()}
        self._with_grad(c_names)
        critic_loss, cm = critic_fn()
        gs = torch.autograd.grad(critic_loss, [self.P[n] for n in c_names], allow_unused=True)
        self._no_grad()
        cgr = {n: g for n, g in zip(c_names, gs) if g is not None}
        out["critic_loss"] = critic_loss.detach()
        out["critic_grad_norm"] = optimizer_step(self.P, c_names, cgr, copt)
        out.update({k: v.detach() for k, v in cm.items()})
        # update_slow_target (dreamer.py:455-462)
        if self.cfg["slow_target"]:
            if self.updates[beh] % self.cfg["slow_target_update"] == 0:
                mix = 1.0 if self.updates[beh] == 0 else float(self.cfg["slow_target_fraction"])
                for n in c_names:
                    tn = n.replace(f"{beh}.critic.", f"{beh}._target_critic.")
                    self.P[tn] = mix * self.P[n] + (1 - mix) * self.P[tn]
            self.updates[beh] += 1
        self.last_actor_grads, self.last_critic_grads = agr, cgr
        self.last_seq, self.last_target = seq, target.detach()
        return {f"{prefix}{k}": v for k, v in out.items()}

    def update_imag_behavior(self, outs, target, n_frames=8):
        """GenRLAgent.update_imag_behavior (genrl.py:108-124) with imag_reward_fn = video_text_reward and the cached `target`."""
        args = self.cfg["imag_reward_args"]
        assert args["score_fn"] == "max_cosine" and not args["align_initial"] and not args["weighted_align"]
        fn = lambda seq: vlm_reward(self.P, seq, target, n_frames, bool(args["align_sequence"]))
        return self.update_behavior(outs, beh="_imag_behavior", reward_fn=fn, prefix="imag_")

    def update(self, data):
        """DreamerAgent.update (dreamer.py:94-97)."""
        state, outs, mets = self.update_wm(data)
        mets.update(self.update_behavior(outs))
        return state, outs, mets


# --------------------------------------------------------------------------------------
# synthetic batches and default-initialised parameters
# --------------------------------------------------------------------------------------
def synthetic_batch(B, T, A, clip_dim=512, seed=0, resets=0.0, img=64):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randint(0, 256, (B, T, 3, img, img), generator=g, dtype=torch.uint8)
    action = torch.rand(B, T, A, generator=g) * 2 - 1
    reward = torch.rand(B, T, 1, generator=g) * 2
    is_first = torch.zeros(B, T, dtype=torch.bool)
    is_first[:, 0] = True
    if resets > 0:
        is_first |= torch.rand(B, T, generator=g) < resets
    clip = F.normalize(torch.randn(B, T, clip_dim, generator=g), dim=-1)
    return dict(observation=obs, action=action, reward=reward, discount=torch.ones(B, T, 1),
                is_first=is_first, is_last=torch.zeros(B, T, dtype=torch.bool),
                is_terminal=torch.zeros(B, T, dtype=torch.bool), clip_video=clip)


def time_cpu_update(agent, data, warmup=1, steps=1, fn="update"):
    """Seconds per call of the oracle's full update on this host's CPU threads."""
    for _ in range(warmup):
        getattr(agent, fn)(data)
    t0 = time.perf_counter()
    for _ in range(steps):
        getattr(agent, fn)(data)
    return (time.perf_counter() - t0) / steps, os.cpu_count()
