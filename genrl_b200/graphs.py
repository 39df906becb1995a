"""CUDA-graph capture of whole update calls.

One GenRL update is ~1 800 kernel launches sequenced by Python + autograd (~45 ms of host time at the benchmark shape);
the kernels themselves need far less.  After a couple of eager warm-up calls the complete call — forward kernels, the
autograd-sequenced backward kernels, grad-norm/clip/weight-decay/Adam, weight-shadow refresh and (data-parallel) the NCCL
all-reduce — is captured once into a CUDA graph and replayed with a single launch.  Inputs are copied into static
buffers; outputs are static tensors that stay valid until the next call (the contract train.py relies on).
Anything host-side that must differ between steps lives on the device (Adam step counters, the Philox step counter,
RewardEMA buffer) or runs outside the graph (slow-target copy every N updates).
"""
import torch


def _flatten(obj, out, path=""):
    """collect tensors of a nested dict/tuple/list structure in deterministic order"""
    if torch.is_tensor(obj):
        out.append((path, obj))
    elif isinstance(obj, dict):
        for k in obj:
            _flatten(obj[k], out, f"{path}/{k}")
    elif isinstance(obj, (tuple, list)):
        for i, v in enumerate(obj):
            _flatten(v, out, f"{path}/{i}")
    return out


def _rebuild(obj, it):
    if torch.is_tensor(obj):
        return next(it)
    if isinstance(obj, dict):
        return {k: _rebuild(v, it) for k, v in obj.items()}
    if isinstance(obj, tuple):
        return tuple(_rebuild(v, it) for v in obj)
    if isinstance(obj, list):
        return [_rebuild(v, it) for v in obj]
    return obj


class GraphedCall:
    def __init__(self, fn, warmup=2):
        self.fn = fn
        self.warmup = warmup
        self.calls = 0
        self.graph = None
        self.static_in = None
        self.static_out = None
        self.sig = None

    def reset(self):
        self.__init__(self.fn, self.warmup)

    def _stage_inputs(self, args):
        flat = _flatten(args, [])
        sig = tuple((p, tuple(t.shape), t.dtype) for p, t in flat)
        if sig != self.sig:
            self.reset()
            self.sig = sig
            self.static_in = [t.detach().clone() for _, t in flat]
        else:
            for s, (_, t) in zip(self.static_in, flat):
                if s.data_ptr() != t.data_ptr():
                    s.copy_(t.detach(), non_blocking=True)
        return _rebuild(args, iter(self.static_in))

    def __call__(self, *args):
        sargs = self._stage_inputs(args)
        if self.graph is not None:
            self.graph.replay()
            return self.static_out
        if self.calls < self.warmup:
            self.calls += 1
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                out = self.fn(*sargs)
            torch.cuda.current_stream().wait_stream(s)
            return out
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            out = self.fn(*sargs)
        self.graph = g
        self.static_out = out
        g.replay()                       # capture records without executing
        return out
