import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_agent import build_agent
from genrl_b200 import ops
for prec in (32, 16):
    for mode in ("wm", "beh", True):
        try:
            agent, data, g, meta = build_agent("tiny_genrl", prec)
            agent.cfg.overlap_streams = mode
            agent.cfg.cuda_graph = True
            ops.set_noise(ops.PhiloxNoise(3))
            for step in range(4):
                state, outputs, metrics = agent.update_wm(data, step)
                start, metrics = agent.update_acting_behavior(state, outputs, metrics, data)
            torch.cuda.synchronize()
            print("OK  ", prec, mode, float(metrics["model_loss"]))
        except Exception as e:
            print("FAIL", prec, mode, repr(e)[:200])
            tb = traceback.extract_tb(e.__traceback__)
            print("     at", [(f.name, f.lineno) for f in tb[-6:]])
            torch.cuda.synchronize()
