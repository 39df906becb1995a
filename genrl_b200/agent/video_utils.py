"""B200-native mirror of the reference's agent/video_utils.py: the connector (VideoSSM) and aligner (UNetDenoiser).

The connector is an RSSM driven by video/text embeddings as "actions" and teacher-forced on the world model's posterior
stoch, so — exactly like EnsembleRSSM.observe with single_obs_posterior — everything but the GRU recurrence is batched
over all B*T rows and the recurrence runs in the persistent scan kernel.

Reference: agent/video_utils.py (ResidualLinear :8, UNetDenoiser :27, VideoSSM :64, .update :127, .video_imagine :209).
"""
from collections import defaultdict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from . import dreamer_utils as common
from .dreamer_utils import _rows, dense_norm_act


class ResidualLinear(nn.Module):
    def __init__(self, in_channels, out_channels, norm="layer", act="SiLU", prenorm=False):
        super().__init__()
        assert not prenorm
        self.norm_layer = common.NormLayer(norm, in_channels if prenorm else out_channels)
        self.act = common.get_act(act)
        self.layer = nn.Linear(in_channels, out_channels)
        self.prenorm = prenorm
        self.res_proj = nn.Identity() if in_channels == out_channels else nn.Linear(in_channels, out_channels)

    def forward(self, x):
        """x: [M, in] compute dtype"""
        if self.norm_layer._layer is not None:
            h = dense_norm_act(x, self.layer, self.norm_layer, common._act_code(self.act))
        else:
            assert isinstance(self.act, nn.Identity)
            h = ops.linear(x, self.layer.weight, self.layer.bias)
        res = x if isinstance(self.res_proj, nn.Identity) else ops.linear(x, self.res_proj.weight, self.res_proj.bias)
        return h + res


class UNetDenoiser(nn.Module):
    def __init__(self, in_channels: int, mid_channels: int, n_layers: int, norm="layer", act="SiLU"):
        super().__init__()
        out_channels = in_channels
        self.down = nn.ModuleList()
        for i in range(n_layers):
            if i == (n_layers - 1):
                self.down.append(ResidualLinear(in_channels, mid_channels, norm=norm, act=act))
            else:
                self.down.append(ResidualLinear(in_channels, in_channels, norm=norm, act=act))
        self.mid = nn.ModuleList()
        for i in range(n_layers):
            self.mid.append(ResidualLinear(mid_channels, mid_channels, norm=norm, act=act))
        self.up = nn.ModuleList()
        for i in range(n_layers):
            if i == 0:
                self.up.append(ResidualLinear(mid_channels * 2, out_channels, norm="none", act="Identity"))
            else:
                self.up.append(ResidualLinear(out_channels * 2, out_channels, norm=norm, act=act))

    def forward(self, x):
        lead = tuple(x.shape[:-1])
        x = ops.to_compute(_rows(x))
        down_res = []
        for down_layer in self.down:
            x = down_layer(x)
            down_res.append(x)
        for mid_layer in self.mid:
            x = mid_layer(x)
        down_res.reverse()
        for up_layer, res in zip(self.up, down_res):
            x = up_layer(ops.concat(x, res))
        x = ops.to_f32(x)
        return x.reshape(lead + (x.shape[-1],))


class VideoSSM(common.EnsembleRSSM):
    def __init__(self, *args, connector_kl={}, temporal_embeds=False, detached_post=True, n_frames=8, token_dropout=0.,
                 loss_scale=1, clip_add_noise=0, clip_lafite_noise=0, rescale_embeds=False, denoising_ae=False, learn_initial=True,
                 **kwargs):
        super().__init__(*args, **kwargs)
        self.n_frames = n_frames
        self.viclip_emb_dim = kwargs["action_dim"] - self.n_frames
        self.temporal_embeds = temporal_embeds
        self.detached_post = detached_post
        self.connector_kl = connector_kl
        self.token_dropout = token_dropout
        self.loss_scale = loss_scale
        self.rescale_embeds = rescale_embeds
        self.clip_add_noise = clip_add_noise
        self.clip_lafite_noise = clip_lafite_noise
        self.clip_const = np.sqrt(self.viclip_emb_dim).item()
        self.denoising_ae = denoising_ae
        assert token_dropout == 0 and detached_post, "connector fast path: token_dropout 0, detached posterior (GenRL configs)"
        if self.denoising_ae:
            self.aligner = UNetDenoiser(self.viclip_emb_dim, self.viclip_emb_dim // 2, n_layers=2, norm="layer", act="SiLU")
        self.learn_initial = learn_initial
        if self.learn_initial:
            self.initial_state_pred = nn.Sequential(
                nn.Linear(kwargs["action_dim"], kwargs["hidden"]),
                common.NormLayer(kwargs["norm"], kwargs["hidden"]), common.get_act("SiLU"),
                nn.Linear(kwargs["hidden"], kwargs["hidden"]),
                common.NormLayer(kwargs["norm"], kwargs["hidden"]), common.get_act("SiLU"),
                nn.Linear(kwargs["hidden"], kwargs["deter"]))
        del self._obs_out
        del self._obs_dist

    def _initial_deter(self, init_embed):
        """initial_state_pred MLP: [M, E+8] -> deter fp32 [M, D]"""
        sp = self.initial_state_pred
        x = dense_norm_act(ops.to_compute(_rows(init_embed)), sp[0], sp[1])
        x = dense_norm_act(x, sp[3], sp[4])
        return ops.linear(x, sp[6].weight, sp[6].bias, out_fp32=True)

    def initial(self, batch_size, init_embed=None, ignore_learned=False):
        """reference :100-112"""
        init = super().initial(batch_size)
        if self.learn_initial and not ignore_learned and hasattr(self, "initial_state_pred"):
            assert init_embed is not None
            if init_embed.shape[-1] == self.viclip_emb_dim:
                patcher = torch.zeros((*init_embed.shape[:-1], 8), device=init_embed.device)
                init_embed = torch.cat([init_embed, patcher], dim=-1)
            S, K = self._stoch, self._discrete
            deter = self._initial_deter(init_embed.float())
            logit = self._prior_logit(deter)
            ops.noise.randint()
            stoch = self._sample(logit, True)
            init["deter"] = deter
            init["stoch"] = stoch.reshape(-1, S, K)
            init["logit"] = logit.reshape(-1, S, K)
        return init

    def get_action(self, video_embed):
        """reference :114-125"""
        n_frames = self.n_frames
        B, T = video_embed.shape[:2]
        if self.rescale_embeds:
            video_embed = video_embed * self.clip_const
        temporal_embeds = F.one_hot(torch.arange(T, device=video_embed.device) % n_frames, n_frames).reshape(1, T, n_frames).repeat(B, 1, 1)
        if not self.temporal_embeds:
            temporal_embeds = temporal_embeds * 0
        return torch.cat([video_embed, temporal_embeds.to(video_embed.dtype)], dim=-1)

    def update(self, video_embed, wm_post):
        """reference :127-207 — returns (loss, metrics)"""
        n_frames = self.n_frames
        B, T = video_embed.shape[:2]
        S, K, D, Hd = self._stoch, self._discrete, self._deter, self._hidden
        dev = video_embed.device
        N = B * T
        metrics = {}
        video_embed = video_embed[:, n_frames - 1::n_frames].float()
        video_embed = video_embed.reshape(B, T // n_frames, 1, -1).repeat(1, 1, n_frames, 1).reshape(B, T, -1)
        orig_video_embed = video_embed
        if self.clip_add_noise > 0:
            video_embed = video_embed + ops.noise.randn(tuple(video_embed.shape), dev) * self.clip_add_noise
            video_embed = F.normalize(video_embed, dim=-1)
        if self.clip_lafite_noise > 0:
            normed_noise = F.normalize(ops.noise.randn(tuple(video_embed.shape), dev), dim=-1)
            video_embed = (1 - self.clip_lafite_noise) * video_embed + self.clip_lafite_noise * normed_noise
            video_embed = F.normalize(video_embed, dim=-1)
        loss = 0
        if self.denoising_ae:
            assert (self.clip_lafite_noise + self.clip_add_noise) > 0, "Nothing to denoise"
            denoised_embed = F.normalize(self.aligner(video_embed), dim=-1)
            denoising_loss = 1 - F.cosine_similarity(denoised_embed, orig_video_embed, dim=-1).mean()
            loss = loss + denoising_loss
            metrics["aligner_cosine_distance"] = denoising_loss.detach()
            video_embed = orig_video_embed
        embed_actions = self.get_action(video_embed)                                   # [B, T, E+8]
        wm_stoch = wm_post["stoch"].detach().reshape(N, S * K).float().contiguous()
        wm_logit = wm_post["logit"].detach().reshape(N, S * K).float().contiguous()
        prev_idx, next_idx, first_idx = self._indices(B, T, dev)
        # t = 0: learned initial state from the first action (with gradient); t >= 1: teacher-forced posterior stoch
        init = self.initial(B, init_embed=embed_actions[:, 0])
        row0 = (torch.arange(B, device=dev, dtype=torch.int32) * T).contiguous()
        in_stoch = ops.shift_mask(init["stoch"].reshape(B, S * K), first_idx, None, row0, None)
        ops.gather_rows(wm_stoch, prev_idx, None, in_stoch.detach(), accumulate=True)
        acts = ops.as_compute(embed_actions.reshape(N, -1).contiguous())
        x = dense_norm_act(ops.concat(in_stoch, acts), self._img_in[0], self._img_in[1])
        w = self._cell._layer.weight
        gx = ops.linear(x, w, None, out_fp32=True, wcols=(0, Hd))
        deter = ops.gru_scan(gx, w, self._cell._norm.weight, self._cell._norm.bias, None, init["deter"], B, T, Hd)
        prior_logit = self._prior_logit(deter)
        if isinstance(ops.noise, ops.TapeNoise):       # the reference samples (and discards) a prior stoch per step
            for _ in range(T):
                ops.noise.randint(); ops.noise.exp(B * S, K, dev)
        kl_loss, kl_value = self.kl_loss({"logit": wm_logit.reshape(B, T, S, K)}, {"logit": prior_logit.reshape(B, T, S, K)}, **self.connector_kl)
        metrics["connector_kl"] = self._last_kl_stats[0][1]
        loss = loss + self.loss_scale * kl_loss
        # initial-KL metric pass (reference :197-205)
        with torch.no_grad():
            ve2 = video_embed.reshape(B, T // n_frames, n_frames, -1)[:, 1:, 0].reshape(B * (T // n_frames - 1), 1, -1)
            act2 = self.get_action(ve2)[:, 0]
            sel = lambda v: v.reshape(B, T // n_frames, n_frames, S * K)[:, 1:, 0].reshape(B * (T // n_frames - 1), S * K)
            prev2 = self.initial(act2.shape[0], init_embed=act2)
            prior2 = self.img_step(prev2, act2)
            self.kl_loss({"logit": sel(wm_logit).reshape(-1, S, K).contiguous()}, {"logit": prior2["logit"]}, **self.connector_kl)
            metrics["connector_initial_kl"] = self._last_kl_stats[0][1]
        return loss, metrics

    def video_imagine(self, video_embed, dreamer_init=None, sample=True, reset_every_n_frames=True, denoise=False):
        """reference :209-239"""
        n_frames = self.n_frames
        B, T = video_embed.shape[:2]
        if self.denoising_ae and denoise:
            denoised_embed = self.aligner(video_embed.float())
            video_embed = F.normalize(denoised_embed, dim=-1)
        action = self.get_action(video_embed)
        init = self.initial(batch_size=B, init_embed=action[:, 0])
        if dreamer_init is not None:
            init[self.cell_input] = dreamer_init[self.cell_input]
        if reset_every_n_frames:
            prior_states = defaultdict(list)
            for action_chunk in torch.chunk(action, T // n_frames, dim=1):
                prior = self.imagine(action_chunk, init, sample=sample)
                for k in prior:
                    prior_states[k].append(prior[k])
                init = self.initial(batch_size=B, ignore_learned=True)
                init[self.cell_input] = prior[self.cell_input][:, -1]
            prior = {k: torch.cat(v, dim=1) for k, v in prior_states.items()}
        else:
            prior = self.imagine(action, init, sample=sample)
        return prior
