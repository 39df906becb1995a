"""B200-native mirror of the reference's agent/dreamer.py: DreamerAgent, WorldModel, ActorCritic.

Same constructor signature, methods, returned structures and metric names as the reference, so it is
selected with `agent._target_=genrl_b200.agent.dreamer.DreamerAgent` (or by shadowing `agent/` on sys.path) and
`train.py` runs unchanged.  All arithmetic runs in libgenrl_b200.so kernels via genrl_b200.ops.

Reference: agent/dreamer.py (DreamerAgent :19, WorldModel :120, ActorCritic :323).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..graphs import GraphedCall
from ..store import ParamStore
from . import dreamer_utils as common


def stop_gradient(x):
    return x.detach()


Module = nn.Module


def env_reward(agent, seq):
    """reference dreamer.py:16-17"""
    return agent.wm.heads["reward"](seq["feat"]).mean


_REWARD_FNS = {"env_reward": env_reward}


def register_reward_fn(name, fn):
    _REWARD_FNS[name] = fn


def _compute_dtype(cfg):
    return torch.bfloat16 if cfg.precision == 16 else torch.float32


class DreamerAgent(Module):
    def __init__(self, name, cfg, obs_space, act_spec, **kwargs):
        super().__init__()
        self.name = name
        self.cfg = cfg
        self.cfg.update(**kwargs)
        self.obs_space = obs_space
        self.act_spec = act_spec
        self._use_amp = (cfg.precision == 16)
        self.device = cfg.device
        self.act_dim = act_spec.shape[0]
        self.wm = WorldModel(cfg, obs_space, self.act_dim)
        self.instantiate_acting_behavior()
        self.to(cfg.device)
        self.requires_grad_(requires_grad=False)
        self._store = None
        self._store_sig = None
        self._moment_bank = None
        self._graphs = {}
        self.dp = None                      # genrl_b200.dist.DataParallel when running one process per GPU
        self.wm._owner = [self]             # lets train.py's direct wm.* calls (train.py:277-280) reach ensure_store()

    # ---- flat parameter store (lazy: train.py swaps sub-modules after construction) ----
    def _param_groups(self):
        groups = OrderedDict()
        conn = getattr(self.wm, "connector", None)
        conn_ids = {id(p) for p in conn.parameters()} if conn is not None else set()
        groups["wm"] = ([p for p in self.wm.parameters() if id(p) not in conn_ids], True)
        if conn is not None:
            groups["connector"] = (list(conn.parameters()), True)
        seen = set()
        for attr in ("_acting_behavior", "_imag_behavior"):
            beh = getattr(self, attr, None)
            if beh is None or id(beh) in seen:
                continue
            seen.add(id(beh))
            groups[f"{attr}.actor"] = (list(beh.actor.parameters()), True)
            groups[f"{attr}.critic"] = (list(beh.critic.parameters()), True)
            if beh._target_critic is not beh.critic:
                groups[f"{attr}.target"] = (list(beh._target_critic.parameters()), False)
        return groups

    def ensure_store(self):
        groups = self._param_groups()
        sig = tuple((k, tuple(id(p) for p in v[0])) for k, v in groups.items())
        if self._store is not None and sig == self._store_sig:
            return self._store
        # Adam moments follow the PARAMETERS (in the reference every ActorCritic / WorldModel owns its torch Adam), not the group
        # name: after finetune_mode() the group "_acting_behavior.actor" holds the imagination behaviour's parameters and must
        # inherit that behaviour's moments.  Bank = {ids of the group's parameters: (m, v, state)} from the previous store and
        # from an unpickled snapshot (bound to identities in __setstate__).
        bank = dict(getattr(self, "_moment_bank", None) or {})
        if self._store is not None:
            for og in self._store.groups.values():
                if og.trainable:
                    bank[tuple(id(p) for p in og.params)] = (og.m, og.v, og.state)
        store = ParamStore(self.device, _compute_dtype(self.cfg))
        with ops.use_dtype(store.compute_dtype):
            for k, (params, trainable) in groups.items():
                g = store.add_group(k, params, trainable)
                prev = bank.get(tuple(id(p) for p in params)) if trainable else None
                if prev is not None and prev[0].numel() == g.numel:
                    g.m.copy_(prev[0]); g.v.copy_(prev[1]); g.state.copy_(prev[2])
        self._store, self._store_sig = store, sig
        self._moment_bank = None
        self.wm._owner = [self]
        self._graphs = {}                   # captured graphs hold pointers into the old flat buffers
        opts = [self.wm.model_opt]
        for attr in ("_acting_behavior", "_imag_behavior"):
            beh = getattr(self, attr, None)
            if beh is not None:
                opts += [beh.actor_opt, beh.critic_opt]
        for o in opts:
            o.store = store
            o.allreduce = self.dp.allreduce if self.dp is not None else None
        return store

    def __getstate__(self):
        state = self.__dict__.copy()
        st = state.pop("_store", None)
        state["_store"] = None
        state["_store_sig"] = None
        state["_graphs"] = {}
        state["dp"] = None
        state["_moment_bank"] = None
        if st is not None:
            state["_store_state"] = {k: (g.m.clone(), g.v.clone(), g.state.clone()) for k, g in st.groups.items() if g.trainable}
        elif getattr(self, "_moment_bank", None):
            groups = self._param_groups()
            state["_store_state"] = {k: self._moment_bank[tuple(id(p) for p in ps)] for k, (ps, tr) in groups.items()
                                     if tr and tuple(id(p) for p in ps) in self._moment_bank}
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        # bind the snapshot's optimizer state (saved per group name) to the identities of the just-restored parameters, so that
        # later attribute swaps (finetune_mode, train.py:142-157) cannot hand it to a different module
        saved = self.__dict__.pop("_store_state", None)
        self._moment_bank = None
        if saved:
            groups = self._param_groups()
            self._moment_bank = {tuple(id(p) for p in ps): saved[k] for k, (ps, tr) in groups.items() if tr and k in saved}
        if getattr(self, "wm", None) is not None:
            self.wm._owner = [self]

    def _ctx(self):
        return ops.use_dtype(_compute_dtype(self.cfg))

    def instantiate_acting_behavior(self):
        self._acting_behavior = ActorCritic(self.cfg, self.act_spec, self.wm.inp_size).to(self.device)

    def act(self, obs, meta, step, eval_mode, state):
        """reference dreamer.py:41-64 (batch-1 inference; runs on the same kernels, op by op)"""
        if self.cfg.only_random_actions:
            return np.random.uniform(-1, 1, self.act_dim).astype(self.act_spec.dtype), (None, None)
        with torch.no_grad(), self._ctx():
            self.ensure_store()
            obs = {k: torch.as_tensor(np.copy(v), device=self.device).unsqueeze(0) for k, v in obs.items()}
            if state is None:
                latent = self.wm.rssm.initial(len(obs["reward"]))
                action = torch.zeros((len(obs["reward"]),) + tuple(self.act_spec.shape), device=self.device)
            else:
                latent, action = state
            embed = self.wm.encoder(obs)
            should_sample = (not eval_mode) or (not self.cfg.eval_state_mean)
            latent, _ = self.wm.rssm.obs_step(latent, action, embed, obs["is_first"], should_sample)
            feat = self.wm.rssm.get_feat(latent)
            actor = self._acting_behavior.actor(feat)
            action = actor.mean if eval_mode else actor.sample()
            return action.cpu().numpy()[0], (latent, action)

    # ---- CUDA-graph plumbing: cfg.cuda_graph=True captures each update call after two eager warm-ups ----
    def _use_graphs(self):
        return bool(getattr(self.cfg, "cuda_graph", False)) and isinstance(ops.noise, ops.PhiloxNoise)

    def _run(self, key, fn, *args):
        if not self._use_graphs():
            return fn(*args)
        gc = self._graphs.get(key)
        if gc is None:
            gc = self._graphs[key] = GraphedCall(fn)
        return gc(*args)

    _WM_KEYS = ("observation", "action", "reward", "is_first", "clip_video")

    def _update_wm_impl(self, data):
        ops.noise_tick(self.device)
        return self.wm.update(data, state=None)

    def update_wm(self, data, step):
        metrics = {}
        with self._ctx():
            self.ensure_store()
            used = {k: data[k] for k in self._WM_KEYS if k in data}
            state, outputs, mets = self._run("wm", self._update_wm_impl, used)
        outputs = dict(outputs)
        outputs["is_terminal"] = data["is_terminal"]
        metrics.update(mets)
        return state, outputs, metrics

    def _behavior_impl(self, attr, fn_name, post):
        ops.noise_tick(self.device)
        beh = getattr(self, attr)
        start = {k: stop_gradient(v) for k, v in post.items()}
        reward_fn = self._reward_closure(fn_name)
        return beh.update(self.wm, start, None, reward_fn, slow_target=False)

    def _reward_closure(self, fn_name):
        if callable(fn_name):
            return lambda seq: fn_name(self, seq)
        if fn_name == getattr(self.cfg, "imag_reward_fn", None) and fn_name != self.cfg.acting_reward_fn:
            return lambda seq: _REWARD_FNS[fn_name](self, seq, **self.cfg.imag_reward_args)
        return lambda seq: _REWARD_FNS[fn_name](self, seq)

    def _update_behavior(self, attr, fn_name, post):
        key = (attr, fn_name if isinstance(fn_name, str) else id(fn_name))
        mets = self._run(key, lambda p: self._behavior_impl(attr, fn_name, p), post)
        getattr(self, attr).update_slow_target()          # host-side counter + occasional copy: outside the graph
        return mets

    def update_acting_behavior(self, state=None, outputs=None, metrics={}, data=None, reward_fn=None):
        if self.cfg.only_random_actions:
            return {}, metrics
        with self._ctx():
            self.ensure_store()
            if outputs is not None:
                post = outputs["post"]
            else:
                with torch.no_grad():
                    embed = self.wm.encoder(data)
                    post, _ = self.wm.rssm.observe(embed, data["action"], data["is_first"])
            start = {k: stop_gradient(v) for k, v in post.items()}
            metrics.update(self._update_behavior("_acting_behavior", reward_fn or self.cfg.acting_reward_fn, start))
        return start, metrics

    def update(self, data, step):
        state, outputs, metrics = self.update_wm(data, step)
        start, metrics = self.update_acting_behavior(state, outputs, metrics, data)
        return state, metrics

    def report(self, data):
        report = {}
        with torch.no_grad(), self._ctx():
            self.ensure_store()
            for key in self.wm.heads["decoder"].cnn_keys:
                name = key.replace("/", "_")
                report[f"openl_{name}"] = self.wm.video_pred(data, key)
        return report

    def get_meta_specs(self):
        return tuple()

    def init_meta(self):
        return OrderedDict()

    def update_meta(self, meta, global_step, time_step, finetune=False):
        return meta


class WorldModel(Module):
    def __init__(self, config, obs_space, act_dim):
        super().__init__()
        shapes = {k: tuple(v.shape) for k, v in obs_space.items()}
        self.shapes = shapes
        self.cfg = config
        self.device = config.device
        self.encoder = common.Encoder(shapes, **config.encoder)
        self.embed_dim = self.encoder.embed_dim()
        self.rssm = common.EnsembleRSSM(**config.rssm, action_dim=act_dim, embed_dim=self.embed_dim, device=self.device)
        self.heads = {}
        self._use_amp = (config.precision == 16)
        self.inp_size = self.rssm.get_feat_size()
        self.decoder_input_fn = getattr(self.rssm, f"get_{config.decoder_inputs}")
        self.decoder_input_size = getattr(self.rssm, f"get_{config.decoder_inputs}_size")()
        self.heads["decoder"] = common.Decoder(shapes, **config.decoder, embed_dim=self.decoder_input_size, image_dist=config.image_dist)
        self.heads["reward"] = common.MLP(self.inp_size, (1,), **config.reward_head)
        with torch.no_grad():
            for p in self.heads["reward"]._out.parameters():
                p.data = p.data * 0
        assert not config.pred_discount, "discount head is not on the GenRL path (pred_discount: False)"
        for name in config.grad_heads:
            assert name in self.heads, name
        self.grad_heads = config.grad_heads
        self.heads = nn.ModuleDict(self.heads)
        self.model_opt = common.Optimizer("model", self.parameters(), **config.model_opt, use_amp=self._use_amp)
        self.e2e_update_fns = {}
        self.detached_update_fns = {}
        self.eval()

    def add_module_to_update(self, name, module, update_fn, detached=False):
        self.add_module(name, module)
        if detached:
            self.detached_update_fns[name] = update_fn
        else:
            self.e2e_update_fns[name] = update_fn
        self.model_opt = common.Optimizer("model", self.parameters(), **self.cfg.model_opt, use_amp=self._use_amp)

    def update(self, data, state=None):
        """reference dreamer.py:166-187"""
        self._in_agent_call = True
        try:
            return self._update(data, state)
        finally:
            self._in_agent_call = False

    def _update(self, data, state=None):
        self.train()
        overlap = getattr(self.cfg, "overlap_streams", False) in (True, "wm") and len(self.detached_update_fns) > 0 and not self.e2e_update_fns
        if overlap:
            state, outputs, metrics = self._update_overlapped(data, state)
            self.eval()
            return state, outputs, metrics
        with common.RequiresGrad(self):
            model_loss, state, outputs, metrics = self.loss(data, state)
            model_loss, metrics = self.update_additional_e2e_modules(data, outputs, model_loss, metrics)
            metrics.update(self.model_opt(model_loss, self.parameters()))
        outputs = {k: ({kk: vv.detach() for kk, vv in v.items()} if isinstance(v, dict) else v.detach()) for k, v in outputs.items()}
        state = {k: v.detach() for k, v in state.items()}
        if len(self.detached_update_fns) > 0:
            metrics = self._detached_steps(data, outputs, metrics)
        self.eval()
        return state, outputs, metrics

    def _update_overlapped(self, data, state=None):
        """Same update, two streams: the detached modules' forward+backward (connector: a latency-bound chain of small kernels and
        the persistent scan) run beside the world model's backward (conv/GEMM heavy).  The optimizer steps keep the reference
        order (world model, then connector); the only difference is that the connector's forward reads its weights before the
        world-model step's 1e-6 multiplicative decay instead of after."""
        main = torch.cuda.current_stream()
        side = common.side_stream(self.device)
        with common.RequiresGrad(self):
            model_loss, state, outputs, metrics = self.loss(data, state)
            outputs = {k: ({kk: vv.detach() for kk, vv in v.items()} if isinstance(v, dict) else v.detach()) for k, v in outputs.items()}
            state = {k: v.detach() for k, v in state.items()}
            side.wait_stream(main)
            det = []
            with torch.cuda.stream(side):
                for k in self.detached_update_fns:
                    add_loss, add_metrics = self.detached_update_fns[k](self, k, data, outputs, metrics)
                    self.model_opt.backward(add_loss)
                    det.append((k, add_loss, add_metrics))
                    # data parallel: this module's gradients are final -> average them on this stream, beside the world model's backward
                    self.model_opt.reduce(getattr(self, k).parameters())
            self.model_opt.backward(model_loss)
            main.wait_stream(side)
            # the reference steps the world model on model_loss' gradients alone (dreamer.py:183); the detached modules' gradients are
            # already in their flat buffers here, so they are excluded from this norm / clip / Adam pass (they still get the
            # multiplicative weight decay every handed-in parameter receives, dreamer_utils.py:926-932) and stepped below with their own norm
            det_ids = {id(p) for k in self.detached_update_fns for p in getattr(self, k).parameters()}
            metrics.update(self.model_opt.apply(model_loss, self.parameters(), active=[p for p in self.parameters() if id(p) not in det_ids]))
            for k, add_loss, add_metrics in det:
                metrics.update(add_metrics)
                opt_metrics = self.model_opt.apply(add_loss, getattr(self, k).parameters())
                metrics.update({f"{k}_{m}": opt_metrics[m] for m in opt_metrics})
        return state, outputs, metrics

    def update_additional_detached_modules(self, data, outputs, metrics):
        """reference dreamer.py:189-200.  Called from update() and, by train.py:280, directly (a second connector step per iteration);
        the direct call draws a fresh Philox key and, with cfg.cuda_graph, is captured into a CUDA graph like the agent's update calls."""
        detached_loss = 0
        if getattr(self, "_in_agent_call", False):
            return detached_loss, self._detached_steps(data, outputs, metrics)
        owner = self._ensure_store()
        with ops.use_dtype(_compute_dtype(self.cfg)):
            if owner is not None and owner._use_graphs() and set(self.detached_update_fns) <= {"connector"}:
                used = {"clip_video": data["clip_video"]}
                post = {k: outputs["post"][k] for k in ("stoch", "logit")}
                mets = owner._run("detached", self._detached_direct, used, post)
            else:
                mets = self._detached_direct(data, outputs["post"])
        metrics.update(mets)
        return detached_loss, metrics

    def _detached_direct(self, data, post):
        ops.noise_tick(self.device)
        return self._detached_steps(data, {"post": post}, {})

    def _detached_steps(self, data, outputs, metrics):
        with ops.use_dtype(_compute_dtype(self.cfg)):
            for k in self.detached_update_fns:
                detached_module = getattr(self, k)
                with common.RequiresGrad(detached_module):
                    add_loss, add_metrics = self.detached_update_fns[k](self, k, data, outputs, metrics)
                    metrics.update(add_metrics)
                    opt_metrics = self.model_opt(add_loss, detached_module.parameters())
                    metrics.update({f"{k}_{m}": opt_metrics[m] for m in opt_metrics})
        return metrics

    def _ensure_store(self):
        """train.py calls wm.observe_data / wm.update_additional_detached_modules directly (train.py:277-280): make sure the flat
        parameter store exists; returns the owning agent"""
        owner = getattr(self, "_owner", None)
        if owner and owner[0] is not None:
            owner[0].ensure_store()
            return owner[0]
        if self.model_opt.store is None:
            raise RuntimeError("WorldModel is not attached to an agent: call agent.ensure_store() first")
        return None

    def update_additional_e2e_modules(self, data, outputs, model_loss, metrics):
        for k in self.e2e_update_fns:
            add_loss, add_metrics = self.e2e_update_fns[k](self, k, data, outputs, metrics)
            model_loss = model_loss + add_loss
            metrics.update(add_metrics)
        return model_loss, metrics

    def observe_data(self, data, state=None):
        """reference dreamer.py:210-217"""
        self._ensure_store()
        ops.noise_tick(self.device)
        with torch.no_grad(), ops.use_dtype(_compute_dtype(self.cfg)):
            embed = self.encoder(data)
            post, prior = self.rssm.observe(embed, data["action"], data["is_first"], state)
            kl_loss, kl_value = self.rssm.kl_loss(post, prior, **self.cfg.kl)
            stats, _ = self.rssm._last_kl_stats
        outs = dict(embed=embed, post=post, prior=prior, is_terminal=data["is_terminal"])
        return outs, {"model_kl": stats[1]}

    def loss(self, data, state=None):
        """reference dreamer.py:219-252"""
        embed = self.encoder(data)
        post, prior = self.rssm.observe(embed, data["action"], data["is_first"], state)
        kl_loss, kl_value = self.rssm.kl_loss(post, prior, **self.cfg.kl)
        stats, kl_forward = self.rssm._last_kl_stats
        likes = {}
        losses = {"kl": kl_loss}
        feat = self.rssm.get_feat(post)
        for name, head in self.heads.items():
            grad_head = (name in self.grad_heads)
            inp = self.decoder_input_fn(post) if name == "decoder" else feat
            inp = inp if grad_head else stop_gradient(inp)
            out = head(inp)
            dists = out if isinstance(out, dict) else {name: out}
            for key, dist in dists.items():
                like = dist.log_prob(data[key])
                likes[key] = like
                losses[key] = ops.scaled_mean(like, -1.0).reshape(())
        model_loss = sum(self.cfg.loss_scales.get(k, 1.0) * v for k, v in losses.items())
        outs = dict(embed=embed, feat=feat, post=post, prior=prior, likes=likes, kl=kl_value)
        metrics = {f"{name}_loss": value.detach() for name, value in losses.items()}
        metrics["model_kl"] = stats[1]
        # kernel stats are (ent lhs, ent rhs) with lhs = post unless the KL is the forward one
        metrics["prior_ent"] = stats[2] if kl_forward else stats[3]
        metrics["post_ent"] = stats[3] if kl_forward else stats[2]
        last_state = {k: v[:, -1] for k, v in post.items()}
        return model_loss, last_state, outs, metrics

    def imagine(self, policy, start, is_terminal, horizon, task_cond=None, eval_policy=False):
        """reference dreamer.py:254-287"""
        assert task_cond is None
        flatten = lambda x: x.reshape([-1] + list(x.shape[2:]))
        start = {k: flatten(v) for k, v in start.items()}
        return self._imagine_rows(policy, start, horizon, eval_policy)

    def _imagine_rows(self, policy, start, horizon, eval_policy=False):
        from .. import rollout
        if getattr(self.cfg, "fused_rollout", True) and start["deter"].is_cuda and rollout.supported(policy, self.rssm):
            # explicit-buffer rollout: one autograd node for all H steps (genrl_b200/rollout.py)
            seq = rollout.imagine_rows(policy, self.rssm, start, horizon, eval_policy)
            disc = torch.ones(list(seq["feat"].shape[:-1]) + [1], device=seq["feat"].device)
            seq["discount"] = disc * self.cfg.discount
            seq["weight"] = torch.cumprod(torch.cat([torch.ones_like(disc[:1]), disc[:-1]], 0), 0)
            return seq
        start = dict(start)
        start["feat"] = self.rssm.get_feat(start)
        policy_dist = policy(start["feat"])
        start["action"] = torch.zeros_like(policy_dist.sample())
        seq = {k: [v] for k, v in start.items()}
        ents = []
        for _ in range(horizon):
            inp = seq["feat"][-1]
            policy_dist = policy(stop_gradient(inp))
            action = policy_dist.sample() if not eval_policy else policy_dist.mean
            if not eval_policy and hasattr(policy_dist, "entropy"):
                ents.append(policy_dist.entropy())          # same inputs/weights as actor_loss's re-forward (dreamer.py:397)
            state = self.rssm.img_step({k: v[-1] for k, v in seq.items()}, action)
            feat = self.rssm.get_feat(state)
            for key, value in {**state, "action": action, "feat": feat}.items():
                seq[key].append(value)
        seq = {k: torch.stack(v, 0) for k, v in seq.items()}
        if len(ents) == horizon and horizon > 0:
            seq["policy_ent"] = torch.stack(ents, 0)            # [H, N]: entropy of pi(. | sg feat_t), t = 0..H-1
        disc = torch.ones(list(seq["feat"].shape[:-1]) + [1], device=seq["feat"].device)
        seq["discount"] = disc * self.cfg.discount
        seq["weight"] = torch.cumprod(torch.cat([torch.ones_like(disc[:1]), disc[:-1]], 0), 0)
        return seq

    def preprocess(self, obs):
        """reference dreamer.py:289-305 (API parity; the training path fuses this into the first conv's im2col)"""
        obs = obs.copy()
        for key, value in obs.items():
            if key.startswith("log_"):
                continue
            if value.dtype in [np.uint8, torch.uint8]:
                value = value / 255.0 - 0.5
            obs[key] = value
        obs["reward"] = {"identity": nn.Identity(), "sign": torch.sign, "tanh": torch.tanh}[self.cfg.clip_rewards](obs["reward"])
        obs["discount"] = (1.0 - obs["is_terminal"].float())
        if len(obs["discount"].shape) < len(obs["reward"].shape):
            obs["discount"] = obs["discount"].unsqueeze(-1)
        return obs

    def video_pred(self, data, key, nvid=8):
        """reference dreamer.py:307-321 (takes the raw uint8 batch)"""
        decoder = self.heads["decoder"]
        truth = data[key][:nvid].float() / 255.0
        embed = self.encoder(data)
        states, _ = self.rssm.observe(embed[:nvid, :5], data["action"][:nvid, :5], data["is_first"][:nvid, :5])
        recon = decoder(self.decoder_input_fn(states))[key].mean[:nvid]
        init = {k: v[:, -1] for k, v in states.items()}
        prior = self.rssm.imagine(data["action"][:nvid, 5:], init)
        prior_recon = decoder(self.decoder_input_fn(prior))[key].mean
        model = torch.clip(torch.cat([recon[:, :5] + 0.5, prior_recon + 0.5], 1), 0, 1)
        error = (model - truth + 1) / 2
        return torch.cat([truth, model, error], 3)


class ActorCritic(Module):
    def __init__(self, config, act_spec, feat_size, name=""):
        super().__init__()
        self.name = name
        self.cfg = config
        self.act_spec = act_spec
        self._use_amp = (config.precision == 16)
        self.device = config.device
        assert not getattr(self.cfg, "discrete_actions", False), "continuous actions only (GenRL configs)"
        self.actor_grad = getattr(self.cfg, f"{self.name}_actor_grad".strip("_"))
        assert self.actor_grad == "dynamics", "fused behaviour update implements actor_grad: dynamics"
        inp_size = feat_size
        self.actor = common.MLP(inp_size, act_spec.shape[0], **self.cfg.actor)
        self.critic = common.MLP(inp_size, (1,), **self.cfg.critic)
        if self.cfg.slow_target:
            self._target_critic = common.MLP(inp_size, (1,), **self.cfg.critic)
            self._updates = 0
        else:
            self._target_critic = self.critic
        self.actor_opt = common.Optimizer("actor", self.actor.parameters(), **self.cfg.actor_opt, use_amp=self._use_amp)
        self.critic_opt = common.Optimizer("critic", self.critic.parameters(), **self.cfg.critic_opt, use_amp=self._use_amp)
        assert self.cfg.reward_ema, "reward_ema: True is the GenRL configuration"
        self.register_buffer("ema_vals", torch.zeros((2,)))
        self.rewnorm = common.StreamNorm(momentum=1, scale=1.0, device=self.device)
        with torch.no_grad():
            for p in self.critic._out.parameters():
                p.data = p.data * 0
        for s, d in zip(self.critic.parameters(), self._target_critic.parameters()):
            d.data = s.data.clone()

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("dp_gather", None)        # bound to a process group; re-attached by DataParallel.attach after a resume
        return state

    def update(self, world_model, start, is_terminal, reward_fn, slow_target=True):
        """reference dreamer.py:366-390"""
        metrics = {}
        hor = self.cfg.imag_horizon
        if getattr(self.cfg, "overlap_streams", False) in (True, "beh"):
            # critic forward+backward (large 36 000-row kernels) on a second stream beside the actor's backward through the rollout
            main = torch.cuda.current_stream()
            side = common.side_stream(self.device)
            with common.RequiresGrad(self.actor), common.RequiresGrad(self.critic), ops.defer_param_grads(self.actor.parameters()) as dg:
                seq = world_model.imagine(self.actor, start, is_terminal, hor)
                reward = reward_fn(seq)
                seq["reward"], mets1 = self.rewnorm(reward)
                mets1 = {f"reward_{k}": v for k, v in mets1.items()}
                target, mets2, baseline = self.target(seq)
                actor_loss, mets3 = self.actor_loss(seq, target, baseline)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    seq_d = {k: stop_gradient(v) for k, v in seq.items()}
                    critic_loss, mets4 = self.critic_loss(seq_d, target)
                    self.critic_opt.backward(critic_loss)
                    self.critic_opt.reduce(self.critic.parameters())      # collective beside the actor's backward
                self.actor_opt.backward(actor_loss)
                dg.flush()
                main.wait_stream(side)
                metrics.update(self.actor_opt.apply(actor_loss, self.actor.parameters()))
                metrics.update(self.critic_opt.apply(critic_loss, self.critic.parameters()))
        else:
            with common.RequiresGrad(self.actor), ops.defer_param_grads(self.actor.parameters()) as dg:
                seq = world_model.imagine(self.actor, start, is_terminal, hor)
                reward = reward_fn(seq)
                seq["reward"], mets1 = self.rewnorm(reward)
                mets1 = {f"reward_{k}": v for k, v in mets1.items()}
                target, mets2, baseline = self.target(seq)
                actor_loss, mets3 = self.actor_loss(seq, target, baseline)
                self.actor_opt.backward(actor_loss)
                dg.flush()
                metrics.update(self.actor_opt.apply(actor_loss, self.actor.parameters()))
            with common.RequiresGrad(self.critic):
                seq = {k: stop_gradient(v) for k, v in seq.items()}
                critic_loss, mets4 = self.critic_loss(seq, target)
                metrics.update(self.critic_opt(critic_loss, self.critic.parameters()))
        metrics.update(**mets1, **mets2, **mets3, **mets4)
        if slow_target:
            self.update_slow_target()
        return {f"{self.name}_{k}".strip("_"): v for k, v in metrics.items()}

    def reward_ema(self, target):
        """RewardEMA.__call__ (reference dreamer_utils.py:1022-1029) as one radix-select kernel; returns (offset, scale)"""
        flat = target.detach().reshape(-1).contiguous().float()
        if getattr(self, "dp_gather", None) is not None:
            flat = self.dp_gather(flat)                 # quantiles are a global-batch statistic
        out4 = torch.empty(4, device=flat.device, dtype=torch.float32)
        ops.lib.genrl_quantile_ema(ops._st(), flat.data_ptr(), flat.numel(), 0.05, 0.95, 1e-2, self.ema_vals.data_ptr(), out4.data_ptr())
        return out4[0], out4[1]

    def actor_loss(self, seq, target, baseline):
        """reference dreamer.py:392-429 (actor_grad == 'dynamics')"""
        # The reference re-runs the actor on sg(feat[:-2]) only to read its entropy (dreamer.py:397,418); the rollout already
        # evaluated the same policy on the same inputs, so its entropies (values and gradients) are reused.
        if "policy_ent" in seq and seq["policy_ent"].shape[0] == seq["feat"].shape[0] - 1:
            ent_all = seq["policy_ent"][:-1]
        else:
            ent_all = self.actor(stop_gradient(seq["feat"][:-2])).entropy()
        metrics = {}
        offset, scale = self.reward_ema(target)
        normed_target = (target - offset) / scale
        ms = ops.mean_std(normed_target.detach())
        metrics["normed_target_mean"] = ms[0]
        metrics["normed_target_std"] = ms[1]
        metrics["reward_ema_005"] = self.ema_vals[0]
        metrics["reward_ema_095"] = self.ema_vals[1]
        objective = normed_target[1:]
        ent = ent_all[:, :, None]
        ent_scale = self.cfg.actor_ent
        if ent_scale:
            objective = objective + ent_scale * ent
        metrics["actor_ent"] = ops.mean_of(ent.detach())[0]
        metrics["actor_ent_scale"] = ent_scale
        weight = stop_gradient(seq["weight"])
        actor_loss = ops.scaled_mean(weight[:-2] * objective, -1.0).reshape(())
        return actor_loss, metrics

    def critic_loss(self, seq, target):
        """reference dreamer.py:431-438"""
        feat = seq["feat"][:-1]
        target = stop_gradient(target)
        weight = stop_gradient(seq["weight"])
        dist = self.critic(feat)
        critic_loss = ops.scaled_mean(dist.log_prob(target)[:, :, None] * weight[:-1], -1.0).reshape(())
        metrics = {"critic": ops.mean_of(dist.mean.detach())[0]}
        return critic_loss, metrics

    def target(self, seq):
        """reference dreamer.py:440-453"""
        reward = seq["reward"]
        value = self._target_critic(seq["feat"]).mean
        target = ops.lambda_return(reward, value, self.cfg.discount, self.cfg.discount_lambda)
        metrics = {"critic_slow": ops.mean_of(value.detach())[0], "critic_target": ops.mean_of(target.detach())[0]}
        return target, metrics, value[:-1]

    def update_slow_target(self):
        """reference dreamer.py:455-462"""
        if self.cfg.slow_target:
            if self._updates % self.cfg.slow_target_update == 0:
                mix = 1.0 if self._updates == 0 else float(self.cfg.slow_target_fraction)
                with torch.no_grad():
                    for s, d in zip(self.critic.parameters(), self._target_critic.parameters()):
                        d.data.copy_(mix * s.data + (1 - mix) * d.data)
                info = ops.param_info(next(self._target_critic.parameters()))
                if info is not None:
                    info.group.refresh()
            self._updates += 1
