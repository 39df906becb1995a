"""B200-native mirror of the reference's agent/genrl.py: GenRLAgent = Dreamer + connector/aligner + imagination behaviour.

Reference: agent/genrl.py (connector_update_fn :7, GenRLAgent :27, .update_imag_behavior :108).
"""
import torch

from .. import ops
from .dreamer import DreamerAgent, ActorCritic, stop_gradient, env_reward, _REWARD_FNS   # noqa: F401
from . import dreamer_utils as common
from . import video_utils
from .genrl_utils import video_text_reward, compute_reward                               # noqa: F401

_REWARD_FNS["video_text_reward"] = video_text_reward


def connector_update_fn(self, module_name, data, outputs, metrics):
    """reference genrl.py:7-25 (viclip_encode: true — the video embeddings come with the batch)"""
    connector = getattr(self, module_name)
    if not getattr(self.cfg, "viclip_encode", False):
        raise NotImplementedError("online InternVideo2 encoding is out of scope; train with viclip_encode: true (train.yaml:29)")
    video_embed = data["clip_video"]
    wm_post = outputs["post"]
    return connector.update(video_embed, wm_post)


class GenRLAgent(DreamerAgent):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.n_frames = 8
        self.viclip_emb_dim = 512
        assert self.cfg.batch_length % self.n_frames == 0, "Fix batch length param"
        if "clip_video" in self.obs_space:
            self.viclip_emb_dim = self.obs_space["clip_video"].shape[0]
        connector = video_utils.VideoSSM(**self.cfg.connector, **self.cfg.connector_rssm, connector_kl=self.cfg.connector_kl,
                                         n_frames=self.n_frames, action_dim=self.viclip_emb_dim + self.n_frames,
                                         clip_add_noise=self.cfg.clip_add_noise, clip_lafite_noise=self.cfg.clip_lafite_noise,
                                         device=self.device, cell_input="stoch")
        connector.to(self.device)
        connector.requires_grad_(False)
        self.wm.add_module_to_update("connector", connector, connector_update_fn, detached=self.cfg.connector.detached_post)
        if getattr(self.cfg, "imag_reward_fn", None) is not None:
            self.instantiate_imag_behavior()

    def instantiate_imag_behavior(self):
        self._imag_behavior = ActorCritic(self.cfg, self.act_spec, self.wm.inp_size, name="imag").to(self.device)
        self._imag_behavior.requires_grad_(False)
        self._imag_behavior.rewnorm = common.StreamNorm(**self.cfg.imag_reward_norm, device=self.device)

    def finetune_mode(self):
        self._acting_behavior = self._imag_behavior
        self.wm.detached_update_fns = {}
        self.wm.e2e_update_fns = {}
        self.wm.grad_heads.append("reward")

    def update_wm(self, data, step):
        return super().update_wm(data, step)

    def update_imag_behavior(self, state=None, outputs=None, metrics={}, seq_data=None):
        """reference genrl.py:108-124"""
        if getattr(self.cfg, "imag_reward_fn", None) is None:
            return outputs["post"], metrics
        with self._ctx():
            self.ensure_store()
            if outputs is not None:
                post = outputs["post"]
            else:
                with torch.no_grad():
                    embed = self.wm.encoder(seq_data)
                    post, _ = self.wm.rssm.observe(embed, seq_data["action"], seq_data["is_first"])
            start = {k: stop_gradient(v) for k, v in post.items()}
            metrics.update(self._update_behavior("_imag_behavior", self.cfg.imag_reward_fn, start))
        return start, metrics
