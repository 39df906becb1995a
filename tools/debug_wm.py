"""Stage-by-stage comparison of the CUDA world-model forward against the golden reference (debug aid, run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_agent import build_agent
from genrl_b200 import ops
from oracle import genrl_oracle as O

case = sys.argv[1] if len(sys.argv) > 1 else "tiny_genrl"
agent, data, g, meta = build_agent(case, 32)
ops.set_noise(ops.TapeNoise(g["tape"], "cuda:0"))
ref = g["ref0"]
def err(name, a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    print(f"{name:28s} shape {tuple(a.shape)} max|err| {float((a-b).abs().max()):.3e}  scale {float(b.abs().max()):.3e}")
with torch.no_grad(), ops.use_dtype(torch.float32):
    agent.ensure_store()
    wm = agent.wm
    embed = wm.encoder(data)
    err("embed", embed, ref["outputs/embed"])
    post, prior = wm.rssm.observe(embed, data["action"], data["is_first"])
    for grp, d in (("post", post), ("prior", prior)):
        for k in ("logit", "stoch", "deter"):
            err(f"{grp}.{k}", d[k], ref[f"outputs/{grp}/{k}"])
    # feed the reference embed to isolate the rssm
    ops.set_noise(ops.TapeNoise(g["tape"], "cuda:0"))
    post, prior = wm.rssm.observe(ref["outputs/embed"].to("cuda:0"), data["action"], data["is_first"])
    for grp, d in (("post", post), ("prior", prior)):
        for k in ("logit", "stoch", "deter"):
            err(f"[ref embed] {grp}.{k}", d[k], ref[f"outputs/{grp}/{k}"])
    kl_loss, kl_value = wm.rssm.kl_loss(post, prior, **agent.cfg.kl)
    err("kl value", kl_value, ref["outputs/kl"])
    dec = wm.heads["decoder"](wm.decoder_input_fn({k: v.to("cuda:0") for k, v in {"stoch": ref["outputs/post/stoch"]}.items()}))
    like = dec["observation"].log_prob(data["observation"])
    err("likes.observation", like, ref["outputs/likes/observation"])
    rew = wm.heads["reward"](ref["outputs/feat"].to("cuda:0")).log_prob(data["reward"])
    err("likes.reward", rew, ref["outputs/likes/reward"])

print("---- training path ----")
agent, data, g, meta = build_agent(case, 32)
ops.set_noise(ops.TapeNoise(g["tape"], "cuda:0"))
state, outputs, metrics = agent.update_wm(data, 0)
torch.cuda.synchronize()
def report(tag):
    for grp in ("post", "prior"):
        for k in ("logit", "stoch", "deter"):
            err(f"{tag} {grp}.{k}", outputs[grp][k], ref[f"outputs/{grp}/{k}"])
    err(f"{tag} embed", outputs["embed"], ref["outputs/embed"])
    err(f"{tag} feat", outputs["feat"], ref["outputs/feat"])
report("after update_wm")
for k, v in meta["metrics"][0].items():
    if k in metrics:
        print(f"   metric {k:28s} mine {float(torch.as_tensor(metrics[k]).float().cpu().reshape(-1)[0]): .6f}  ref {v: .6f}")
start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
torch.cuda.synchronize()
report("after behaviour")
for k, v in meta["metrics"][0].items():
    if k in metrics:
        print(f"   metric {k:28s} mine {float(torch.as_tensor(metrics[k]).float().cpu().reshape(-1)[0]): .6f}  ref {v: .6f}")
    else:
        print("   MISSING metric", k)
