"""Flat parameter storage for the fused optimizer step.

Every optimizer group (world model, connector, actor, critic, target critic) lives in ONE flat fp32 buffer with a
parallel flat gradient buffer and flat Adam moments, so that
  * wgrad kernels accumulate straight into the gradient buffer (no per-tensor .grad traffic),
  * grad-norm, clip, weight decay and Adam are a handful of launches over the flat buffers
    (reference: Optimizer.__call__ dreamer_utils.py:892-932),
  * a data-parallel step needs a single NCCL all-reduce per group set,
  * compute-layout weight shadows (bf16 casts, K-padded rows, (kh,kw,Cin)-packed conv filters) are refreshed by one
    multi-tensor launch after the step.
The nn.Parameters keep their reference names/shapes (state_dict compatible); their .data are views into the flat buffer.
"""
import ctypes

import torch

from ._lib import lib
from . import ops


class _PackEntry(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("n", ctypes.c_longlong), ("kind", ctypes.c_int),
                ("B", ctypes.c_int), ("P", ctypes.c_int), ("cols", ctypes.c_int), ("ld", ctypes.c_int)]


def _align(n, a):
    return (n + a - 1) // a * a


class ParamGroup:
    def __init__(self, name, params, device, compute_dtype, trainable=True):
        self.name = name
        self.params = [p for p in params]
        self.touched = False
        self.reduced = False               # gradients already averaged over the data-parallel group (reduce() ran early)
        self.trainable = trainable
        self.compute_dtype = compute_dtype
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += _align(p.numel(), 4)
        self.offsets, self.numel = offs, total
        self.flat = torch.zeros(total, device=device, dtype=torch.float32)
        self.grad = torch.zeros(total, device=device, dtype=torch.float32) if trainable else None
        self.m = torch.zeros(total, device=device, dtype=torch.float32) if trainable else None
        self.v = torch.zeros(total, device=device, dtype=torch.float32) if trainable else None
        self.state = torch.zeros(2, device=device, dtype=torch.float32)        # {adam step, last grad norm}
        self.attach()

    def attach(self):
        """(re)point every parameter at its slice of the flat buffers and rebuild the shadow table"""
        dev = self.flat.device
        entries, sh_total = [], 0
        plan = []
        for p, off in zip(self.params, self.offsets):
            n = p.numel()
            view = self.flat[off:off + n].view(p.shape)
            if p.data.data_ptr() != view.data_ptr():
                view.copy_(p.data.to(dev, torch.float32))
                p.data = view
            gview = self.grad[off:off + n].view(p.shape) if self.trainable else None
            info = ops.register_param(p, gview, None, self)
            if p.dim() == 4:                                   # conv filter: packed layout in every mode
                A, B, kh, kw = p.shape
                plan.append((p, off, 1, A, kh * kw * B, kh * kw * B, (B, kh * kw)))
                sh_total += _align(A * kh * kw * B, 8)
            elif p.dim() == 2 and self.compute_dtype != torch.float32:
                N, K = p.shape
                ld = _align(K, 8)
                plan.append((p, off, 0 if ld == K else 2, N, K, ld, None))
                sh_total += _align(N * ld, 8)
        self.shadow = torch.zeros(max(sh_total, 8), device=dev, dtype=self.compute_dtype)
        soff = 0
        es = self.shadow.element_size()
        for p, off, kind, rows, cols, ld, bp in plan:
            n_sh = rows * ld
            sh = self.shadow[soff:soff + n_sh].view(rows, ld)[:, :cols]
            ops.param_info(p).shadow = sh
            e = _PackEntry()
            e.src = self.flat.data_ptr() + off * 4
            e.dst = self.shadow.data_ptr() + soff * es
            e.n = p.numel()
            e.kind = kind
            e.B, e.P = bp if bp else (0, 0)
            e.cols, e.ld = cols, ld
            entries.append(e)
            soff += _align(n_sh, 8)
        self.n_entries = len(entries)
        if entries:
            assert ctypes.sizeof(_PackEntry) == lib.genrl_pack_entry_size()
            arr = (_PackEntry * len(entries))(*entries)
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
            self.table = raw.to(dev)
        else:
            self.table = None
        self.refresh()

    def refresh(self):
        """recompute the compute-layout shadows from the fp32 masters (one launch)"""
        if self.table is not None:
            lib.genrl_pack_weights(ops._st(), self.table.data_ptr(), self.n_entries, ops._CODE[self.compute_dtype])

    def zero_grad(self):
        if self.grad is not None:
            self.grad.zero_()
        self.touched = False


class ParamStore:
    """Owns the ParamGroups of one agent; maps parameters -> groups for the Optimizer mirror."""

    def __init__(self, device, compute_dtype):
        self.device = torch.device(device)
        self.compute_dtype = compute_dtype
        self.groups = {}
        self.sumsq = torch.zeros(1, device=self.device, dtype=torch.float32)

    def add_group(self, name, params, trainable=True):
        g = ParamGroup(name, list(params), self.device, self.compute_dtype, trainable)
        self.groups[name] = g
        return g

    def groups_of(self, params):
        seen, out = set(), []
        for p in params:
            info = ops.param_info(p)
            if info is None:
                raise RuntimeError("parameter is not registered in the ParamStore")
            g = info.group
            if id(g) not in seen:
                seen.add(id(g))
                out.append(g)
        return out

    def reduce(self, groups, allreduce):
        """average the gradients of `groups` over the data-parallel ranks NOW, on the current stream (called right after the backward
        that produced them, so the collective overlaps whatever the other stream is still computing); step() then skips them"""
        if allreduce is None:
            return
        todo = [g for g in groups if g.trainable and g.touched and not g.reduced]
        if todo:
            allreduce([g.grad for g in todo])
            for g in todo:
                g.reduced = True

    def refresh_all(self):
        for g in self.groups.values():
            g.refresh()

    def step(self, groups, lr, eps, clip, wd, allreduce=None, only=None):
        """clip_grad_norm_ -> weight decay on every group handed in -> Adam on the groups that received gradients.

        Returns the device scalar holding the (pre-clip) gradient norm."""
        st = ops._st()
        active = [g for g in groups if g.trainable and g.touched and (only is None or any(g is o for o in only))]
        if allreduce is not None:
            pending = [g for g in active if not g.reduced]
            if pending:
                allreduce([g.grad for g in pending])
        for g in active:
            g.reduced = False
        first = True
        for g in active:
            lib.genrl_grad_sumsq(st, g.grad.data_ptr(), g.numel, self.sumsq.data_ptr(), int(first))
            first = False
        if first:
            self.sumsq.zero_()
        decay = 1.0 - wd if wd else 1.0
        if wd:
            for g in groups:
                if not any(g is a for a in active):       # groups without gradients still decay (reference :926-932 runs on every param)
                    lib.genrl_scale(st, g.flat.data_ptr(), g.numel, decay)
        norm = None
        for g in active:                                   # decay folded into the Adam pass: p <- p*(1-wd), then the Adam update
            lib.genrl_adam_step(st, g.flat.data_ptr(), g.grad.data_ptr(), g.m.data_ptr(), g.v.data_ptr(), g.numel, lr, eps,
                                float(clip or 0.0), self.sumsq.data_ptr(), g.state.data_ptr(), 1, decay)
            g.touched = False
            norm = g.state[1]
        for g in groups:
            g.refresh()
        return norm if norm is not None else self.sumsq[0].sqrt()
