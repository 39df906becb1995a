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

    def report(self, data, key="observation", nvid=8):
        """reference genrl.py:64-106 with viclip_encode: true — the open-loop Dreamer video plus `video_clip_pred`: the first n_frames are
        reconstructed from the posterior, the rest is imagined by the connector from the clip's video embeddings (held for n_frames each)
        starting at the last posterior state, decoded, and stacked [truth | model | error] along the image height.
        Inference-only composition of the kernels the update uses (encoder, observe, connector rollout, decoder)."""
        report_data = super().report(data)
        if not getattr(self.cfg, "viclip_encode", False):
            raise NotImplementedError("online InternVideo2 encoding is out of scope; report needs data['clip_video'] (viclip_encode: true)")
        wm = self.wm
        nf = wm.connector.n_frames
        with torch.no_grad(), self._ctx():
            self.ensure_store()
            decoder = wm.heads["decoder"]
            truth = data[key][:nvid].float() / 255.0
            B, T = truth.shape[0], truth.shape[1] - nf                      # frames after the n_frames warm-up
            embed = wm.encoder(data)
            states, _ = wm.rssm.observe(embed[:nvid, :nf], data["action"][:nvid, :nf], data["is_first"][:nvid, :nf])
            recon = decoder(wm.decoder_input_fn(states))[key].mean[:nvid]
            last = {k: v[:, -1] for k, v in states.items()}
            # one embedding per n_frames chunk (the chunk's last frame), repeated over the chunk: these are the connector's "actions"
            clip = data["clip_video"][:nvid, 2 * nf - 1::nf].to(self.device)
            clip = clip.reshape(B, T // nf, 1, -1).expand(B, T // nf, nf, clip.shape[-1]).reshape(B, T, -1)
            prior = wm.connector.video_imagine(clip, last, reset_every_n_frames=False)
            imagined = decoder(wm.decoder_input_fn(prior))[key].mean
            model = torch.clip(torch.cat([recon[:, :nf] + 0.5, imagined + 0.5], 1), 0, 1)
            report_data["video_clip_pred"] = torch.cat([truth, model, (model - truth + 1) / 2], 3)
        return report_data

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
