"""B200-native mirror of the reference's agent/dreamer_utils.py building blocks.

Same class names, constructor arguments, parameter names/shapes (state_dict compatible) and module creation order
(so `torch.manual_seed(s)` yields the same initial weights as the reference) — but every forward/backward runs in the
hand-written sm_100a kernels of libgenrl_b200.so through genrl_b200.ops.  The nn.Linear / nn.LayerNorm / nn.Conv2d
children are parameter containers only; their own forward() is never called.

Reference: agent/dreamer_utils.py (EnsembleRSSM :302, GRUCell :750, Encoder :558, Decoder :631, MLP :718, DistLayer :787,
NormLayer :844, Optimizer :871, StreamNorm :934, RewardEMA :1014, ImgChLayerNorm :1031).
"""
import re

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_NONE, ACT_SILU

Module = nn.Module


def get_act(name):
    if name in ("SiLU", "silu"):
        return nn.SiLU()
    if name in ("none", "Identity"):
        return nn.Identity()
    if hasattr(nn, name):
        return getattr(nn, name)()
    raise NotImplementedError(name)


def _act_code(act):
    if isinstance(act, nn.SiLU):
        return ACT_SILU
    if isinstance(act, nn.Identity):
        return ACT_NONE
    raise NotImplementedError(f"activation {act} has no fused kernel (the GenRL configs use SiLU)")


def _rows(x):
    return x.reshape(-1, x.shape[-1])


# ------------------------------------------------------------------------------------------------------------
# distributions (value objects over kernel outputs)
# ------------------------------------------------------------------------------------------------------------
class MSEDist:
    """reference dreamer_utils.py:62-83"""

    def __init__(self, mode, agg="sum"):
        self._mode = mode
        self._agg = agg

    @property
    def mean(self):
        return self._mode

    def mode(self):
        return self._mode

    def log_prob(self, value):
        lead = self._mode.shape[:-3]
        m = self._mode.reshape((-1,) + tuple(self._mode.shape[-3:]))
        if value.dtype == torch.uint8:
            like = ops.mse_logprob(m, value.reshape(m.shape))
        else:   # already-preprocessed float target (off the hot path: reports only)
            like = -((m - value.reshape(m.shape).float()) ** 2).sum(dim=(1, 2, 3))
        return like.reshape(lead)


class TwoHotDist:
    """reference dreamer_utils.py:120-171; logits fp32 [..., 255]"""

    def __init__(self, logits):
        assert logits.shape[-1] == 255
        self.logits = logits

    @property
    def mean(self):
        m = ops.twohot_mean(_rows(self.logits))
        return m.reshape(tuple(self.logits.shape[:-1]) + (1,))

    @property
    def mode(self):
        return self.mean

    def log_prob(self, x):
        lp = ops.twohot_logprob(_rows(self.logits), x.reshape(-1))
        return lp.reshape(tuple(self.logits.shape[:-1]))


class NormalDist:
    """Independent(Normal(tanh(out), 0.9*sigmoid(std+2)+0.1), 1) — DistLayer 'normal' (dreamer_utils.py:814-819)."""

    def __init__(self, out, shape):
        self._out = out                      # fp32 [..., 2A]: mean pre-activations | std pre-activations
        self._A = int(np.prod(shape))
        self._cache = None

    def _run(self, use_mean):
        rows = _rows(self._out)
        eps = None if use_mean else ops.noise.randn((rows.shape[0], self._A), rows.device)
        action, ent = ops.actor_head(rows, eps, self._A, use_mean)
        lead = tuple(self._out.shape[:-1])
        return action.reshape(lead + (self._A,)), ent.reshape(lead)

    def sample(self):
        a, ent = self._run(False)
        self._cache = ent
        return a

    rsample = sample

    @property
    def mean(self):
        return self._run(True)[0]

    def mode(self):
        return self.mean

    def entropy(self):
        if self._cache is not None:
            return self._cache
        return self._run(True)[1]


class CatDist:
    """Independent(OneHotDist(logit), 1) of the reference (dreamer_utils.py:177-197, :410-416)."""

    def __init__(self, logit):
        self.logit = logit.float()
        self.S, self.K = logit.shape[-2], logit.shape[-1]

    def _flat(self):
        return self.logit.reshape(-1, self.S * self.K)

    def sample(self):
        rows = self._flat()
        nz, seed, off = ops.noise.exp(rows.shape[0] * self.S, self.K, rows.device)
        st, _ = ops.cat_sample(rows, self.S, self.K, nz, seed, off, True)
        return st.reshape(self.logit.shape)

    def mode(self):
        st, _ = ops.cat_sample(self._flat(), self.S, self.K, None, 0, 0, False)
        return st.reshape(self.logit.shape)

    @property
    def mean(self):
        """probs of the unimix categorical (OneHotCategorical.mean)"""
        p = torch.softmax(self.logit, -1)
        p = 0.99 * p + 0.01 / self.K
        return p / p.sum(-1, keepdim=True)

    def entropy(self):
        """per-row entropy summed over the S latents (off the hot path; the WM loss takes the means from the KL kernel)"""
        p = self.mean
        return -(p * torch.log(p)).sum(-1).sum(-1)


# ------------------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------------------
class NormLayer(Module):
    def __init__(self, name, dim=None):
        super().__init__()
        if name == "none":
            self._layer = None
        elif name == "layer":
            assert dim is not None
            self._layer = nn.LayerNorm(dim)
        else:
            raise NotImplementedError(name)

    def forward(self, features):
        if self._layer is None:
            return features
        y = ops.ln_act(_rows(features), self._layer.weight, self._layer.bias, self._layer.eps, ACT_NONE)
        return y.reshape(features.shape)


class ImgChLayerNorm(nn.Module):
    def __init__(self, ch, eps=1e-03):
        super().__init__()
        self.norm = torch.nn.LayerNorm(ch, eps=eps)


def dense_norm_act(x, lin, norm, act_code=ACT_SILU):
    """Linear -> NormLayer('layer') -> act as one fused node; falls back to Linear only when norm is 'none'."""
    if norm._layer is None:
        raise NotImplementedError("norm='none' blocks are not on the GenRL path")
    return ops.linear_ln_act(x, lin.weight, lin.bias, norm._layer.weight, norm._layer.bias, norm._layer.eps, act_code)


class DistLayer(Module):
    def __init__(self, in_dim, shape, dist="mse", min_std=0.1, max_std=1.0, init_std=0.0, bias=True):
        super().__init__()
        self._in_dim = in_dim
        self._shape = shape if type(shape) in [list, tuple] else [shape]
        self._dist = dist
        self._min_std = min_std
        self._init_std = init_std
        self._max_std = max_std
        self._out = nn.Linear(in_dim, int(np.prod(shape)), bias=bias)
        if dist in ("normal", "tanh_normal", "trunc_normal"):
            self._std = nn.Linear(in_dim, int(np.prod(shape)))

    def forward(self, inputs):
        x = _rows(inputs)
        lead = tuple(inputs.shape[:-1])
        out = ops.linear(x, self._out.weight, self._out.bias, out_fp32=True)
        if self._dist == "twohot":
            return TwoHotDist(out.reshape(lead + (255,)))
        if self._dist == "normal":
            assert (self._min_std, self._max_std) == (0.1, 1.0), "fused actor head assumes std = 0.9*sigmoid(.+2)+0.1"
            std = ops.linear(x, self._std.weight, self._std.bias, out_fp32=True)
            both = torch.cat([out, std], -1)
            return NormalDist(both.reshape(lead + (both.shape[-1],)), self._shape)
        if self._dist == "mse":
            return MSEDist(out.reshape(lead + tuple(self._shape)))
        raise NotImplementedError(self._dist)


class MLP(Module):
    def __init__(self, in_shape, shape, layers, units, act="SiLU", norm="none", **out):
        super().__init__()
        self._in_shape = in_shape
        if out["dist"] == "twohot":
            shape = 255
        self._shape = (shape,) if isinstance(shape, int) else shape
        self._layers = layers
        self._units = units
        self._norm = norm
        self._act = get_act(act)
        last_units = in_shape
        for index in range(self._layers):
            self.add_module(f"dense{index}", nn.Linear(last_units, units, bias=norm != "none"))
            self.add_module(f"norm{index}", NormLayer(norm, units))
            last_units = units
        self._out = DistLayer(units, shape, **out)

    def trunk(self, features):
        x = ops.to_compute(_rows(features))
        code = _act_code(self._act)
        for index in range(self._layers):
            x = dense_norm_act(x, getattr(self, f"dense{index}"), getattr(self, f"norm{index}"), code)
        return x

    def forward(self, features):
        x = self.trunk(features)
        x = x.reshape(tuple(features.shape[:-1]) + (x.shape[-1],))
        return self._out(x)


class GRUCell(Module):
    def __init__(self, inp_size, size, norm=False, act="Tanh", update_bias=-1, device="cuda", **kwargs):
        super().__init__()
        assert norm and act == "Tanh" and update_bias == -1, "fused GRU kernel implements the GenRL configuration"
        self._inp_size = inp_size
        self._size = size
        self._update_bias = update_bias
        self.device = device
        self._layer = nn.Linear(inp_size + size, 3 * size, bias=False, **kwargs)
        self._norm = nn.LayerNorm(3 * size)

    def get_initial_state(self, inputs=None, batch_size=None, dtype=None):
        return torch.zeros((batch_size), self._size, device=self.device)

    @property
    def state_size(self):
        return self._size

    def forward(self, inputs, deter_state):
        """inputs [B, hidden] (compute dtype), deter_state = [h fp32]"""
        h = deter_state[0]
        xh = ops.concat(inputs, h)
        z = ops.linear(xh, self._layer.weight, None)
        out = ops.gru_gates(z, self._norm.weight, self._norm.bias, h)
        return out, [out]


# ------------------------------------------------------------------------------------------------------------
# RSSM
# ------------------------------------------------------------------------------------------------------------
class EnsembleRSSM(Module):
    def __init__(self, ensemble=5, stoch=30, deter=200, hidden=200, discrete=False, act="SiLU", norm="none", std_act="softplus",
                 min_std=0.1, action_dim=None, embed_dim=1536, device="cuda", single_obs_posterior=False, cell_input="stoch",
                 cell_type="gru"):
        super().__init__()
        assert action_dim is not None
        assert discrete, "the fused path implements the categorical latent (discrete > 0) of the GenRL configs"
        assert ensemble == 1 and norm == "layer" and cell_input == "stoch" and cell_type == "gru"
        self.device = device
        self._embed_dim = embed_dim
        self._action_dim = action_dim
        self._ensemble = ensemble
        self._stoch = stoch
        self._deter = deter
        self._hidden = hidden
        self._discrete = discrete
        self._act = get_act(act)
        self._norm = norm
        self._std_act = std_act
        self._min_std = min_std
        self._cell_type = cell_type
        self.cell_input = cell_input
        self._cell = GRUCell(self._hidden, self._deter, norm=True, device=self.device)
        self.single_obs_posterior = single_obs_posterior
        self._ensemble_img_dist = nn.ModuleList([nn.Linear(hidden, stoch * discrete) for _ in range(ensemble)])
        self._obs_dist = nn.Linear(hidden, stoch * discrete)
        cell_state_input_size = self.get_stoch_size()
        self._img_in = nn.Sequential(nn.Linear(cell_state_input_size + action_dim, hidden, bias=True), NormLayer(norm, hidden))
        self._ensemble_img_out = nn.ModuleList(
            [nn.Sequential(nn.Linear(self.get_deter_size(), hidden, bias=True), NormLayer(norm, hidden)) for _ in range(ensemble)])
        if self.single_obs_posterior:
            self._obs_out = nn.Sequential(nn.Linear(embed_dim, hidden, bias=True), NormLayer(norm, hidden))
        else:
            self._obs_out = nn.Sequential(nn.Linear(deter + embed_dim, hidden, bias=True), NormLayer(norm, hidden))
        self._idx_cache = {}

    # ---- sizes / accessors (reference :384-408) ----
    def initial(self, batch_size):
        return dict(logit=torch.zeros([batch_size, self._stoch, self._discrete], device=self.device),
                    stoch=torch.zeros([batch_size, self._stoch, self._discrete], device=self.device),
                    deter=self._cell.get_initial_state(None, batch_size))

    def get_stoch_size(self):
        return self._stoch * self._discrete

    def get_deter_size(self):
        return self._cell.state_size

    def get_feat_size(self):
        return self.get_deter_size() + self.get_stoch_size()

    def get_stoch(self, state):
        stoch = state["stoch"]
        return stoch.reshape(list(stoch.shape[:-2]) + [self._stoch * self._discrete])

    def get_deter(self, state):
        return state["deter"]

    def get_feat(self, state):
        """[stoch_flat | deter] (reference :405-408), assembled directly in the compute dtype (fp32 in the parity mode, bf16 in
        the tensor-core mode) so that the heads consume it without another cast"""
        stoch, deter = self.get_stoch(state), self.get_deter(state)
        if not stoch.is_cuda:
            return torch.cat([stoch, deter], -1)
        lead = tuple(deter.shape[:-1])
        f = ops.concat(stoch.reshape(-1, stoch.shape[-1]), deter.reshape(-1, deter.shape[-1]))
        return f.reshape(lead + (f.shape[-1],))

    def get_dist(self, state, ensemble=False):
        return CatDist(state["logit"])

    def get_unif_dist(self, state):
        return CatDist(torch.ones_like(state["logit"]))

    # ---- single steps (op-level path: imagination, acting, non-hoistable configs) ----
    def _sample(self, logit_rows, sample):
        S, K = self._stoch, self._discrete
        if sample:
            nz, seed, off = ops.noise.exp(logit_rows.shape[0] * S, K, logit_rows.device)
        else:
            nz, seed, off = None, 0, 0
        st, _ = ops.cat_sample(logit_rows, S, K, nz, seed, off, bool(sample))
        return st

    def _prior_logit(self, deter):
        """deter fp32 [M, D] -> prior logits fp32 [M, S*K] (_suff_stats_ensemble, ensemble = 1)"""
        lin, norm = self._ensemble_img_out[0][0], self._ensemble_img_out[0][1]
        y = dense_norm_act(ops.to_compute(deter), lin, norm)
        d = self._ensemble_img_dist[0]
        return ops.linear(y, d.weight, d.bias, out_fp32=True)

    def img_step(self, prev_state, prev_action, sample=True):
        """reference :459-487"""
        S, K = self._stoch, self._discrete
        lead = tuple(prev_action.shape[:-1])
        stoch = prev_state["stoch"].reshape(-1, S * K)
        x = ops.concat(stoch, _rows(prev_action))
        x = dense_norm_act(x, self._img_in[0], self._img_in[1])
        deter, _ = self._cell(x, [prev_state["deter"].reshape(-1, self._deter)])
        logit = self._prior_logit(deter)
        ops.noise.randint()
        st = self._sample(logit, sample)
        return {"stoch": st.reshape(lead + (S, K)), "deter": deter.reshape(lead + (self._deter,)), "logit": logit.reshape(lead + (S, K))}

    def get_post_stoch(self, embed, prior, should_sample=True):
        """reference :442-457"""
        S, K = self._stoch, self._discrete
        lead = tuple(embed.shape[:-1])
        if self.single_obs_posterior:
            x = ops.to_compute(_rows(embed))
        else:
            x = ops.concat(_rows(prior["deter"]), _rows(embed))
        z = dense_norm_act(x, self._obs_out[0], self._obs_out[1])
        logit = ops.linear(z, self._obs_dist.weight, self._obs_dist.bias, out_fp32=True)
        st = self._sample(logit, should_sample)
        return st.reshape(lead + (S, K)), {"logit": logit.reshape(lead + (S, K))}

    def obs_step(self, prev_state, prev_action, embed, is_first, should_sample=True):
        """reference :432-440"""
        m = 1.0 - is_first.float()
        prev_state = {k: v * m.reshape((-1,) + (1,) * (v.dim() - 1)) for k, v in prev_state.items()}
        prev_action = prev_action * m[:, None]
        prior = self.img_step(prev_state, prev_action, should_sample)
        stoch, stats = self.get_post_stoch(embed, prior, should_sample)
        post = {"stoch": stoch, "deter": prior["deter"], **stats}
        return post, prior

    # ---- sequence paths ----
    def _indices(self, B, T, device):
        key = (B, T, str(device))
        c = self._idx_cache.get(key)
        if c is None:
            n = torch.arange(B * T, device=device, dtype=torch.int32)
            t = n % T
            prev = torch.where(t > 0, n - 1, torch.full_like(n, -1))
            nxt = torch.where(t < T - 1, n + 1, torch.full_like(n, -1))
            first = torch.where(t == 0, n // T, torch.full_like(n, -1))
            c = (prev.contiguous(), nxt.contiguous(), first.contiguous())
            self._idx_cache[key] = c
        return c

    def observe(self, embed, action, is_first, state=None):
        """reference :362-371.  With single_obs_posterior the posterior, the GRU input projection and the prior head are
        batched over all B*T rows and only the GRU recurrence runs in the persistent scan kernel."""
        if not self.single_obs_posterior:
            return self._observe_stepwise(embed, action, is_first, state)
        B, T = action.shape[:2]
        S, K, D, Hd = self._stoch, self._discrete, self._deter, self._hidden
        dev = embed.device
        N = B * T
        prev_idx, next_idx, first_idx = self._indices(B, T, dev)
        mask = (1.0 - is_first.reshape(N).float()).contiguous()
        nz_post, nz_prior = ops.noise.observe(B, T, S, K, dev)
        # posterior from the embedding alone
        z = dense_norm_act(ops.to_compute(embed.reshape(N, -1)), self._obs_out[0], self._obs_out[1])
        post_logit = ops.linear(z, self._obs_dist.weight, self._obs_dist.bias, out_fp32=True)
        post_stoch, _ = ops.cat_sample(post_logit, S, K, nz_post[0], nz_post[1], nz_post[2], True)
        # GRU input: [m * stoch_{t-1} | m * action_t]
        mask_next = ops.gather_rows(mask.reshape(N, 1), next_idx, None, torch.empty(N, 1, device=dev, dtype=torch.float32)).reshape(N)
        prev_stoch = ops.shift_mask(post_stoch, prev_idx, mask, next_idx, mask_next)
        act_f = action.reshape(N, -1).float().contiguous()
        act_m = ops.copy2d(act_f, ops.empty_act(N, act_f.shape[1], dev), mask)
        h0 = None
        if state is not None:
            st0 = state["stoch"].reshape(B, S * K).float().contiguous()
            ops.gather_rows(st0, first_idx, mask, prev_stoch.detach(), accumulate=True)
            h0 = state["deter"].float().contiguous()
        x = dense_norm_act(ops.concat(prev_stoch, act_m), self._img_in[0], self._img_in[1])
        w = self._cell._layer.weight
        gx = ops.linear(x, w, None, out_fp32=True, wcols=(0, Hd))
        deter = ops.gru_scan(gx, w, self._cell._norm.weight, self._cell._norm.bias, mask, h0, B, T, Hd)
        prior_logit = self._prior_logit(deter)
        prior_stoch, _ = ops.cat_sample(prior_logit.detach(), S, K, nz_prior[0], nz_prior[1], nz_prior[2], True)
        shp = lambda x, *d: x.reshape((B, T) + d)
        post = {"stoch": shp(post_stoch, S, K), "deter": shp(deter, D), "logit": shp(post_logit, S, K)}
        prior = {"stoch": shp(prior_stoch, S, K), "deter": shp(deter, D), "logit": shp(prior_logit, S, K)}
        return post, prior

    def _observe_stepwise(self, embed, action, is_first, state=None):
        B, T = action.shape[:2]
        if state is None:
            state = self.initial(B)
        posts, priors = [], []
        prev = state
        for t in range(T):
            post, prior = self.obs_step(prev, action[:, t].float(), embed[:, t], is_first[:, t])
            posts.append(post); priors.append(prior)
            prev = post
        stack = lambda lst: {k: torch.stack([d[k] for d in lst], 1) for k in lst[0]}
        return stack(posts), stack(priors)

    def imagine(self, action, state=None, sample=True):
        """reference :373-382 — open-loop prior rollout"""
        B, T = action.shape[:2]
        if state is None:
            state = self.initial(B)
        priors = []
        prev = state
        for t in range(T):
            prev = self.img_step(prev, action[:, t].float(), sample)
            priors.append(prev)
        return {k: torch.stack([d[k] for d in priors], 1) for k in priors[0]}

    def kl_loss(self, post, prior, forward, balance, free, free_avg):
        """reference :534-555; returns (loss 0-d, value [B,T])"""
        assert not free_avg, "free_avg=True is not on the GenRL path"
        S, K = self._stoch, self._discrete
        lead = tuple(post["logit"].shape[:-2])
        loss, value, stats = ops.kl_balance(post["logit"].reshape(-1, S * K), prior["logit"].reshape(-1, S * K), S, K, forward, balance, free)
        self._last_kl_stats = (stats, forward)
        return loss.reshape(()), value.reshape(lead)


# ------------------------------------------------------------------------------------------------------------
# conv encoder / decoder
# ------------------------------------------------------------------------------------------------------------
class Encoder(Module):
    def __init__(self, shapes, cnn_keys=r".*", mlp_keys=r".*", act="SiLU", norm="none", cnn_depth=48, cnn_kernels=(4, 4, 4, 4),
                 mlp_layers=[400, 400, 400, 400], symlog_inputs=False):
        super().__init__()
        self.shapes = shapes
        self.cnn_keys = [k for k, v in shapes.items() if re.match(cnn_keys, k) and len(v) == 3]
        self.mlp_keys = [k for k, v in shapes.items() if re.match(mlp_keys, k) and len(v) == 1]
        assert len(self.cnn_keys) == 1 and not self.mlp_keys, "fused encoder implements the pixel configuration (one image key)"
        assert norm == "layer"
        self._act = get_act(act)
        self._norm = norm
        self._cnn_depth = cnn_depth
        self._cnn_kernels = cnn_kernels
        layers = []
        for i, kernel in enumerate(self._cnn_kernels):
            prev_depth = 3 if i == 0 else 2 ** (i - 1) * self._cnn_depth
            depth = 2 ** i * self._cnn_depth
            layers.append(nn.Conv2d(prev_depth, depth, kernel, stride=2))
            layers.append(ImgChLayerNorm(depth))
            layers.append(self._act)
        self._conv_model = nn.Sequential(*layers)

    def embed_dim(self):
        h = self.shapes[self.cnn_keys[0]][-1]
        for k in self._cnn_kernels:
            h = (h - k) // 2 + 1
        return 2 ** (len(self._cnn_kernels) - 1) * self._cnn_depth * h * h

    def forward(self, data):
        """data[key]: uint8 [..., 3, H, W] (preprocessing fused into the first im2col) -> embed [..., E] (compute dtype)"""
        key = self.cnn_keys[0]
        obs = data[key]
        if obs.dtype != torch.uint8:
            raise NotImplementedError("fused encoder takes the raw uint8 observation (preprocess is fused)")
        lead = tuple(obs.shape[:-3])
        C, H, W = obs.shape[-3:]
        x = obs.reshape(-1, C, H, W).contiguous()
        N = x.shape[0]
        code = _act_code(self._act)
        h, w = H, W
        for i, k in enumerate(self._cnn_kernels):
            conv, ln = self._conv_model[3 * i], self._conv_model[3 * i + 1].norm
            x, h, w = ops.conv2d_s2(x, conv.weight, conv.bias, N, h, w, k, u8=(i == 0))
            x = ops.ln_act(x, ln.weight, ln.bias, ln.eps, code)
        Cl = x.shape[1]
        x = ops.permute_abp(x.reshape(N, h * w * Cl), N, h * w, Cl)          # NHWC rows -> (C,H,W) flatten order (:621)
        return x.reshape(lead + (x.shape[-1],))


class Decoder(Module):
    def __init__(self, shapes, cnn_keys=r".*", mlp_keys=r".*", act="SiLU", norm="none", cnn_depth=48, cnn_kernels=(4, 4, 4, 4),
                 mlp_layers=[400, 400, 400, 400], embed_dim=1024, mlp_dist="mse", image_dist="mse"):
        super().__init__()
        self._embed_dim = embed_dim
        self._shapes = shapes
        self.cnn_keys = [k for k, v in shapes.items() if re.match(cnn_keys, k) and len(v) == 3]
        self.mlp_keys = [k for k, v in shapes.items() if re.match(mlp_keys, k) and len(v) == 1]
        assert len(self.cnn_keys) == 1 and not self.mlp_keys and norm == "layer" and image_dist == "mse"
        self._act = get_act(act)
        self._norm = norm
        self._cnn_depth = cnn_depth
        self._cnn_kernels = cnn_kernels
        self.channels = {k: self._shapes[k][0] for k in self.cnn_keys}
        self._conv_in = nn.Sequential(nn.Linear(embed_dim, 32 * self._cnn_depth))
        layers = []
        for i, kernel in enumerate(self._cnn_kernels):
            if i == 0:
                prev_depth = 32 * self._cnn_depth
            else:
                prev_depth = 2 ** (len(self._cnn_kernels) - (i - 1) - 2) * self._cnn_depth
            depth = 2 ** (len(self._cnn_kernels) - i - 2) * self._cnn_depth
            act_, norm_ = self._act, self._norm
            if i == len(self._cnn_kernels) - 1:
                depth, act_, norm_ = sum(self.channels.values()), nn.Identity(), "none"
            layers.append(nn.ConvTranspose2d(prev_depth, depth, kernel, stride=2))
            layers.append(ImgChLayerNorm(depth) if norm_ == "layer" else NormLayer(norm_, depth))
            layers.append(act_)
        self._conv_model = nn.Sequential(*layers)

    def forward(self, features):
        lead = tuple(features.shape[:-1])
        x = ops.to_compute(_rows(features))
        N = x.shape[0]
        lin = self._conv_in[0]
        x = ops.linear(x, lin.weight, lin.bias)
        h = w = 1
        code = _act_code(self._act)
        n_l = len(self._cnn_kernels)
        for i, k in enumerate(self._cnn_kernels):
            conv = self._conv_model[3 * i]
            last = i == n_l - 1
            x, h, w = ops.conv_transpose2d_s2(x, conv.weight, conv.bias, N, h, w, k, nchw_out=last)
            if not last:
                ln = self._conv_model[3 * i + 1].norm
                x = ops.ln_act(x, ln.weight, ln.bias, ln.eps, code)
        mean = x.reshape(lead + tuple(x.shape[1:]))
        return {self.cnn_keys[0]: MSEDist(mean)}


# ------------------------------------------------------------------------------------------------------------
# optimizer / normalisers
# ------------------------------------------------------------------------------------------------------------
class Optimizer:
    """Mirror of reference Optimizer (dreamer_utils.py:871-932) over the flat ParamStore buffers.

    __call__(loss, params): backward through the kernel tape, then grad-norm / clip / multiplicative weight decay on every
    parameter handed in / Adam on the groups that received gradients — all on device, no host synchronisation."""

    def __init__(self, name, parameters, lr, eps=1e-4, clip=None, wd=None, opt="adam", wd_pattern=r".*", use_amp=False):
        assert 0 <= wd < 1
        assert not clip or 1 <= clip
        assert opt == "adam", "fused optimizer implements Adam (the GenRL configs)"
        assert wd_pattern == r".*"
        self._name = name
        self._lr, self._eps, self._clip, self._wd = float(lr), float(eps), clip, wd
        self._once = True
        self.store = None              # set by the agent (ParamStore)
        self.allreduce = None          # set by the data-parallel wrapper

    def __call__(self, loss, params):
        self.backward(loss)
        return self.apply(loss, params)

    def backward(self, loss):
        assert len(loss.shape) == 0 or (len(loss.shape) == 1 and loss.shape[0] == 1), (self._name, loss.shape)
        loss.backward()

    def __getstate__(self):
        # snapshots (train.py pickles the whole agent) must not carry the flat store (params, grads, moments, shadows: several
        # times the model size) nor a process-group hook; the agent re-attaches both in ensure_store()
        state = self.__dict__.copy()
        state["store"] = None
        state["allreduce"] = None
        return state

    def reduce(self, params):
        """data parallel: average the gradients of the groups of `params` now, on the current stream (no-op on one GPU)"""
        if self.allreduce is not None and _EARLY_REDUCE:
            self.store.reduce(self.store.groups_of(list(params)), self.allreduce)

    def apply(self, loss, params, active=None):
        """grad-norm / clip / weight decay / Adam on the groups of `params` (the backward has already run).  `active` restricts the
        norm / clip / Adam part to the groups of those parameters; the other groups of `params` only get the weight decay."""
        metrics = {f"{self._name}_loss": loss.detach()}
        groups = self.store.groups_of(list(params))
        only = None if active is None else self.store.groups_of(list(active))
        norm = self.store.step(groups, self._lr, self._eps, self._clip, self._wd, self.allreduce, only=only)
        if self._clip:
            metrics[f"{self._name}_grad_norm"] = norm
        return metrics


_SIDE_STREAMS = {}
# Data parallel: issue a finished group's gradient all-reduce on the side stream right away (beside the other stream's backward).  Measured on
# 8 B200s (profiles/r02_notes.md §11): 24.69 ms with it, 24.39 ms without — the NCCL kernels take SMs from the persistent one-CTA-per-SM
# kernels they run beside — so it is opt-in.
_EARLY_REDUCE = __import__("os").environ.get("GENRL_DP_EARLY", "0") == "1"


def side_stream(device, index=0):
    """extra CUDA streams per device: independent kernel chains (connector vs world-model backward, critic vs actor backward,
    row chunks of the imagination rollout) are enqueued there so their latency-bound small kernels fill the SMs the other chain
    leaves idle"""
    k = (str(device), index)
    s = _SIDE_STREAMS.get(k)
    if s is None:
        s = _SIDE_STREAMS[k] = torch.cuda.Stream(device=device)
    return s


class StreamNorm:
    """reference :934-1001; momentum == 1 (the GenRL setting) is the identity"""

    def __init__(self, shape=(), momentum=0.99, scale=1.0, eps=1e-8, device="cuda"):
        assert momentum == 1, "StreamNorm with momentum < 1 is not on the GenRL path"
        self.device = device
        self._momentum = momentum

    def __call__(self, inputs):
        ms = ops.mean_std(inputs.detach())
        metrics = {"mean": ms[0], "std": ms[1], "normed_mean": ms[0], "normed_std": ms[1]}
        return inputs, metrics


class RequiresGrad:
    def __init__(self, model):
        self._model = model

    def __enter__(self):
        self._model.requires_grad_(requires_grad=True)

    def __exit__(self, *args):
        self._model.requires_grad_(requires_grad=False)
