"""Attribute-dict configuration mirroring the reference's Hydra/OmegaConf config.

The reference agent only uses `.attr`, `.get`, `**mapping`, `.update(**kw)` and
`getattr(cfg, k, default)` on its cfg object (reference agent/dreamer.py:25-26,
:127-141; train.py:32-37), so a plain attribute-dict is a drop-in for DictConfig.
Values restate conf/defaults/genrl.yaml, conf/env/dmc_pixels.yaml, agent/genrl.yaml
and train.yaml:24-31 of the reference.
"""
import copy


class AttrDict(dict):
    """dict with attribute access; nested dicts are converted on construction."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def __getstate__(self):
        return dict(self)

    def __setstate__(self, state):
        for k, v in state.items():
            self[k] = v


def genrl_config(device="cuda:0", precision=32, batch_size=48, batch_length=48,
                 imag_horizon=16, task="stickman_walk", agent="genrl", **overrides):
    """GenRL defaults (conf/defaults/genrl.yaml + conf/env/dmc_pixels.yaml + agent/genrl.yaml)."""
    cfg = AttrDict(
        img_size=64,
        rssm=dict(ensemble=1, hidden=1024, deter=1024, stoch=32, discrete=32, norm="layer",
                  std_act="softplus", min_std=0.1, single_obs_posterior=True),
        discount_head=dict(layers=4, units=512, norm="none", dist="binary"),
        reward_head=dict(layers=4, units=1024, norm="layer", dist="twohot"),
        kl=dict(free=1.0, forward=False, balance=0.85, free_avg=False),
        loss_scales=dict(kl=0.6, reward=1.0, discount=1.0, proprio=1.0),
        model_opt=dict(opt="adam", lr=1e-4, eps=1e-8, clip=1000, wd=1e-6),
        decoder_inputs="stoch",
        image_dist="mse",
        actor=dict(layers=4, units=1024, norm="layer", dist="normal", min_std=0.1),
        critic=dict(layers=4, units=1024, norm="layer", dist="twohot"),
        actor_opt=dict(opt="adam", lr=3e-5, eps=1e-5, clip=100, wd=1e-6),
        critic_opt=dict(opt="adam", lr=3e-5, eps=1e-5, clip=100, wd=1e-6),
        discount=0.99,
        discount_lambda=0.95,
        slow_target=True,
        slow_target_update=100,
        slow_target_fraction=1,
        slow_baseline=True,
        reward_ema=True,
        acting_reward_fn="env_reward",
        clip_rewards="identity",
        batch_size=batch_size,
        batch_length=batch_length,
        imag_horizon=imag_horizon,
        eval_state_mean=False,
        precision=precision,
        train_every_actions=10,
        only_random_actions=False,
        # conf/env/dmc_pixels.yaml
        obs_type="pixels",
        action_repeat=2,
        encoder=dict(mlp_keys="$^", cnn_keys="observation", norm="layer", cnn_depth=48,
                     cnn_kernels=[4, 4, 4, 4], mlp_layers=[400, 400, 400, 400]),
        decoder=dict(mlp_keys="$^", cnn_keys="observation", norm="layer", cnn_depth=48,
                     cnn_kernels=[5, 5, 6, 6], mlp_layers=[400, 400, 400, 400]),
        pred_discount=False,
        imag_actor_grad="dynamics",
        actor_grad="dynamics",
        # train.yaml
        task=task,
        seed=1,
        device=device,
        viclip_encode=True,
        viclip_model="internvideo2",
        clip_hd_rendering=False,
    )
    if agent == "genrl":
        # agent/genrl.yaml extras (merged by cfg.update(**kwargs), reference dreamer.py:26)
        cfg.update(
            grad_heads=["decoder"],
            reward_norm=dict(momentum=1.0, scale=1.0, eps=1e-8),
            actor_ent=0,
            additional_report_fns=["report_text2video"],
            clip_add_noise=0.0,
            clip_lafite_noise=0.5,
            connector=dict(token_dropout=0, loss_scale=1, denoising_ae=True, detached_post=True,
                           temporal_embeds=False, rescale_embeds=True),
            connector_rssm=dict(ensemble=1, hidden=1024, deter=1024, stoch=32, discrete=32,
                                norm="layer", std_act="softplus", min_std=0.1,
                                single_obs_posterior=False, learn_initial=True),
            connector_kl=dict(free=0.0, forward=True, balance=0.8, free_avg=False),
            imag_reward_fn=None,
            imag_reward_norm=dict(momentum=1.00, scale=1.0, eps=1e-8),
            imag_reward_args=dict(score_fn="max_cosine", sample_for_target=False, align_initial=False,
                                  weighted_align=False, align_sequence=True, skip_first_target=True),
        )
    else:
        # agent/dreamer.yaml extras
        cfg.update(
            grad_heads=["decoder", "reward"],
            reward_norm=dict(momentum=1.0, scale=1.0, eps=1e-8),
            actor_ent=3e-4,
        )
    for k, v in overrides.items():
        if isinstance(v, dict) and k in cfg and isinstance(cfg[k], dict):
            cfg[k].update(v)
        else:
            cfg[k] = v
    return cfg


def tiny_overrides(units=64, stoch=8, discrete=8, depth=4):
    """Small-dimension overrides used by golden fixtures and fast parity tests."""
    return dict(
        rssm=dict(hidden=units, deter=units, stoch=stoch, discrete=discrete),
        reward_head=dict(units=units),
        actor=dict(units=units),
        critic=dict(units=units),
        encoder=dict(cnn_depth=depth),
        decoder=dict(cnn_depth=depth),
        connector_rssm=dict(hidden=units, deter=units, stoch=stoch, discrete=discrete),
    )


class Spec:
    """Stand-in for the gym/dm_env specs the reference passes (only .shape/.dtype are read)."""

    def __init__(self, shape, dtype="float32"):
        self.shape = tuple(shape)
        self.dtype = dtype

    def __repr__(self):
        return f"Spec(shape={self.shape}, dtype={self.dtype})"


def make_specs(action_dim=10, clip_dim=512, img_size=64, with_clip=True):
    """obs_space / act_spec with the keys the reference's replay provides."""
    obs_space = {
        "observation": Spec((3, img_size, img_size), "uint8"),
        "reward": Spec((), "float32"),
        "discount": Spec((), "float32"),
        "is_first": Spec((), "bool"),
        "is_last": Spec((), "bool"),
        "is_terminal": Spec((), "bool"),
    }
    if with_clip:
        obs_space["clip_video"] = Spec((clip_dim,), "float32")
    act_spec = Spec((action_dim,), "float32")
    return obs_space, act_spec
