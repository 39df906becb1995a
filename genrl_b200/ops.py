"""Host-side operators of the GenRL hot path: thin torch.autograd.Function wrappers around the C-ABI CUDA kernels.

PyTorch is plumbing only (device memory, streams, the autograd tape that sequences hand-written forward/backward
kernels); every arithmetic op on the path is a kernel from libgenrl_b200.so.  There is no CPU fallback: tensors must
live on a CUDA device and the library must load, otherwise these functions raise.

Activations are stored in the *compute dtype* (fp32 for the 1e-4 parity mode, bf16 for the tensor-core mode); logits,
statistics, losses and parameter gradients are always fp32.
"""
import contextlib
import ctypes

import torch
from torch.autograd import Function

from ._lib import lib, GenrlError

F32, BF16 = 0, 1
ACT_NONE, ACT_SILU = 0, 1
_CODE = {torch.float32: F32, torch.bfloat16: BF16}


class _State:
    dtype = torch.float32


def compute_dtype():
    return _State.dtype


@contextlib.contextmanager
def use_dtype(dtype):
    old = _State.dtype
    _State.dtype = dtype
    try:
        yield
    finally:
        _State.dtype = old


def _st():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _check(t, name="tensor"):
    if not t.is_cuda:
        raise GenrlError(f"{name} must be a CUDA tensor: the genrl_b200 path has no CPU fallback")
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise GenrlError(f"{name} must be 2-D with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    return t


def _ld(t):
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _rows2d(t):
    """view [..., D] as [M, D] without copying when possible"""
    if t.dim() == 2:
        return t
    return t.reshape(-1, t.shape[-1])


def _pad8(n):
    return (n + 7) // 8 * 8


def empty_act(rows, cols, device, dtype=None):
    """activation buffer whose leading dimension satisfies the TMA 16-byte stride rule"""
    dtype = dtype or _State.dtype
    ld = _pad8(cols) if dtype == torch.bfloat16 else cols
    buf = torch.empty(rows, ld, device=device, dtype=dtype)
    if ld != cols:
        buf[:, cols:].zero_()
    return buf[:, :cols]


# ---------------------------------------------------------------------------------------------------------------
# weights in compute layout
# ---------------------------------------------------------------------------------------------------------------
def cast_to(t, dtype):
    """t (2-D) -> new tensor of dtype with padded leading dimension, via the library's copy kernel"""
    _check(t)
    out = empty_act(t.shape[0], t.shape[1], t.device, dtype)
    lib.genrl_copy2d(_st(), _p(t), _ld(t), _CODE[t.dtype], _p(out), _ld(out), _CODE[dtype], t.shape[0], t.shape[1], None)
    return out


class ParamInfo:
    __slots__ = ("ref", "grad", "shadow", "group")


_PARAM_INFO = {}


def register_param(p, grad, shadow, group):
    """ParamStore side-table: parameter -> (flat-gradient view, compute-layout shadow, group)."""
    import weakref
    info = ParamInfo()
    info.ref, info.grad, info.shadow, info.group = weakref.ref(p), grad, shadow, group
    _PARAM_INFO[id(p)] = info
    return info


def param_info(p):
    info = _PARAM_INFO.get(id(p))
    if info is not None and info.ref() is p:
        return info
    return None


def compute_weight(w):
    """Linear weight [N,K] (fp32 master) in the compute dtype; uses the ParamStore shadow when one is attached."""
    if _State.dtype == torch.float32:
        return w.detach()
    info = param_info(w)
    if info is not None and info.shadow is not None and info.shadow.dtype == _State.dtype:
        return info.shadow
    return cast_to(w.detach(), _State.dtype)


def packed_conv_weight(w):
    """[A, B, kh, kw] -> [A, kh*kw*B] in the compute dtype (A = Cout for Conv2d, Cin for ConvTranspose2d)."""
    info = param_info(w)
    if info is not None and info.shadow is not None and info.shadow.dtype == _State.dtype:
        return info.shadow
    A, B, kh, kw = w.shape
    out = torch.empty(A, kh * kw * B, device=w.device, dtype=_State.dtype)
    lib.genrl_permute_abp(_st(), _p(w.detach()), F32, _p(out), _CODE[_State.dtype], A, B, kh * kw, 0)
    return out


def _grad_buf(p):
    """(buffer the kernels accumulate into, value handed back to autograd)"""
    info = param_info(p)
    if info is not None and info.grad is not None:
        info.group.touched = True
        return info.grad, None
    z = torch.zeros(p.shape, device=p.device, dtype=torch.float32)
    return z, z


# ---------------------------------------------------------------------------------------------------------------
# raw kernel wrappers
# ---------------------------------------------------------------------------------------------------------------
def gemm(a, w, out=None, bias=None, trans_a=False, trans_w=False, accumulate=False, out_dtype=None):
    """out[M,N] (+)= op(a) @ op(w)^T + bias."""
    _check(a, "gemm A"); _check(w, "gemm W")
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    N, K2 = (w.shape[1], w.shape[0]) if trans_w else (w.shape[0], w.shape[1])
    if K != K2:
        raise GenrlError(f"gemm: inner dimensions differ ({K} vs {K2})")
    if a.dtype != w.dtype:
        raise GenrlError(f"gemm: operand dtypes differ ({a.dtype} vs {w.dtype})")
    if a.dtype == torch.bfloat16:
        if _ld(a) % 8 or a.data_ptr() % 16:
            a = cast_to(a, torch.bfloat16)
        if _ld(w) % 8 or w.data_ptr() % 16:
            w = cast_to(w, torch.bfloat16)
    if out is None:
        out_dtype = out_dtype or a.dtype
        out = empty_act(M, N, a.device, out_dtype)
    _check(out, "gemm C")
    if a.dtype == torch.float32 and out.dtype == torch.float32 and FP32_TC and M * N * K >= FP32_TC_MIN_WORK:
        return _gemm_fp32_x3(a, w, out, bias, trans_a, trans_w, accumulate, M, N, K)
    lib.genrl_gemm(_st(), M, N, K, _p(a), _ld(a), int(trans_a), _p(w), _ld(w), int(trans_w), _CODE[a.dtype], _p(out), _ld(out),
                   _CODE[out.dtype], _p(bias), int(accumulate))
    return out


# fp32 parity mode on the tensor cores: each fp32 operand is split into three bf16 terms (8 + 8 + 8 significant bits) and the six products
# a_i w_j with i + j <= 2 are accumulated in fp32 by the tcgen05 GEMM (smallest terms first).  Relative error ~2^-22 of |a||w|, the same
# order as the fp32 FFMA kernel's accumulation-order noise, at a sixth of the bf16 rate instead of the SIMT kernel's ~7 TFLOP/s.
FP32_TC = __import__("os").environ.get("GENRL_FP32_TC", "1") != "0"
FP32_TC_MIN_WORK = 1 << 24


def _split3(t):
    rows, cols = t.shape
    ld = _pad8(cols)
    buf = torch.zeros(3, rows, ld, device=t.device, dtype=torch.bfloat16) if ld != cols else torch.empty(3, rows, ld, device=t.device, dtype=torch.bfloat16)
    lib.genrl_split3_bf16(_st(), _p(t), _ld(t), buf[0].data_ptr(), buf[1].data_ptr(), buf[2].data_ptr(), ld, rows, cols)
    return [buf[i][:, :cols] for i in range(3)]


def _gemm_fp32_x3(a, w, out, bias, trans_a, trans_w, accumulate, M, N, K):
    a3, w3 = _split3(a), _split3(w)
    first = True
    for i, j in ((0, 2), (2, 0), (1, 1), (0, 1), (1, 0), (0, 0)):
        lib.genrl_gemm(_st(), M, N, K, _p(a3[i]), _ld(a3[i]), int(trans_a), _p(w3[j]), _ld(w3[j]), int(trans_w), BF16, _p(out), _ld(out), F32,
                       _p(bias) if first else None, int(accumulate or not first))
        first = False
    return out


def gemm_pair(a, w, out=None, bias=None, trans_w=False, out_dtype=None):
    """out[M,N] = a @ op(w)^T + bias on the persistent 2-CTA kernel (csrc/gemm_pair.cu); bf16 operands.  Returns None if the shape is not eligible."""
    _check(a, "gemm A"); _check(w, "gemm W")
    M, K = a.shape
    N, K2 = (w.shape[1], w.shape[0]) if trans_w else (w.shape[0], w.shape[1])
    if K != K2 or a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise GenrlError("gemm_pair: bf16 operands with matching inner dimension required")
    if out is None:
        out = empty_act(M, N, a.device, out_dtype or a.dtype)
    rc = lib.load().genrl_gemm_pair(_st(), M, N, K, _p(a), _ld(a), _p(w), _ld(w), int(trans_w), _p(out), _ld(out), _CODE[out.dtype], _p(bias))
    if rc == 2:
        return None
    if rc != 0:
        raise GenrlError(f"genrl_gemm_pair failed: {lib.load().genrl_last_error().decode()}")
    lib.launches += 1
    return out


def as_compute(t):
    t = _check(t)
    return t if t.dtype == _State.dtype else cast_to(t, _State.dtype)


def copy2d(src, dst, row_scale=None):
    lib.genrl_copy2d(_st(), _p(src), _ld(src), _CODE[src.dtype], _p(dst), _ld(dst), _CODE[dst.dtype], src.shape[0], src.shape[1], _p(row_scale))
    return dst


def gather_rows(src, idx, scale, dst, accumulate=False):
    lib.genrl_gather_rows(_st(), _p(src), _ld(src), _CODE[src.dtype], _p(idx), _p(scale), _p(dst), _ld(dst), _CODE[dst.dtype], dst.shape[0],
                          src.shape[1], int(accumulate))
    return dst


def mean_std(x):
    out = torch.empty(2, device=x.device, dtype=torch.float32)
    x = x.contiguous().float()
    lib.genrl_mean_std(_st(), _p(x), x.numel(), _p(out))
    return out


def mean_of(x, scale=1.0, out=None, accumulate=False):
    """scalar tensor = scale * mean(x) (fp32)"""
    x = x.contiguous()
    if out is None:
        out = torch.empty(1, device=x.device, dtype=torch.float32)
    lib.genrl_mean_scale(_st(), _p(x), x.numel(), float(scale), _p(out), int(accumulate))
    return out


# ---------------------------------------------------------------------------------------------------------------
# deferred parameter gradients: a module applied at every step of a rollout (the actor: H+1 calls on N rows) gets ONE
# wgrad GEMM / column pass over all H*N rows at the end of the backward instead of H small ones
# ---------------------------------------------------------------------------------------------------------------
class DeferredGrads:
    def __init__(self, params):
        self.ids = {id(p) for p in params}
        self.jobs = {}

    def wants(self, w):
        return id(w) in self.ids

    def add(self, w, b, dy, x, ln=None):
        j = self.jobs.get(id(w))
        if j is None:
            j = self.jobs[id(w)] = dict(w=w, b=b, dy=[], x=[], ln=None, pre=[], stats=[], dlin=[])
        j["x"].append(x)
        if ln is None:
            j["dy"].append(dy)
        else:
            gamma, beta, act, pre, stats, dpre = ln
            j["ln"] = (gamma, beta, act)
            j["dy"].append(dy); j["pre"].append(pre); j["stats"].append(stats); j["dlin"].append(dpre)

    def flush(self):
        st = _st()
        for j in self.jobs.values():
            w, b = j["w"], j["b"]
            cat = lambda lst: lst[0] if len(lst) == 1 else torch.cat(lst, 0)
            X = cat(j["x"])
            if j["ln"] is None:
                DY = cat(j["dy"])
                gbuf, _ = _grad_buf(w)
                gemm(DY, X, out=gbuf.view(w.shape[0], -1), trans_a=True, trans_w=True, accumulate=True)
                if b is not None:
                    bbuf, _ = _grad_buf(b)
                    lib.genrl_colsum(st, _p(DY), _ld(DY), _p(bbuf), DY.shape[0], DY.shape[1], _CODE[DY.dtype])
            else:
                gamma, beta, act = j["ln"]
                DY, PRE, ST, DP = cat(j["dy"]), cat(j["pre"]), cat(j["stats"]), cat(j["dlin"])
                gbuf, _ = _grad_buf(w)
                gemm(DP, X, out=gbuf.view(w.shape[0], -1), trans_a=True, trans_w=True, accumulate=True)
                gb, _ = _grad_buf(gamma)
                bb, _ = _grad_buf(beta)
                lb = _grad_buf(b)[0] if b is not None else None
                lib.genrl_ln_colgrad(st, _p(DY), _ld(DY), _p(PRE), _ld(PRE), _p(DP), _ld(DP), _p(gamma), _p(beta), _p(ST), act, _p(gb), _p(bb),
                                     _p(lb), DP.shape[0], DP.shape[1], _CODE[DP.dtype])
        self.jobs = {}


_DEFER = None


@contextlib.contextmanager
def defer_param_grads(params):
    """inside the scope, weight/bias/LayerNorm gradients of `params` are accumulated lazily; call .flush() after backward"""
    global _DEFER
    old, _DEFER = _DEFER, DeferredGrads(list(params))
    try:
        yield _DEFER
    finally:
        _DEFER = old


def _deferred_for(w):
    d = _DEFER
    if d is not None and d.wants(w) and param_info(w) is not None:
        return d
    return None


# ---------------------------------------------------------------------------------------------------------------
# Linear (+ LayerNorm + SiLU)
# ---------------------------------------------------------------------------------------------------------------
def _wview(t, wcols):
    return t if wcols is None else t[:, wcols[0]:wcols[1]]


def _linear_backward(ctx_x, w, b, dy, need_dx, need_dw, bias_done=False, wcols=None):
    """shared dgrad / wgrad / bias-grad of y = x w[:, cols]^T + b; dy already in the compute dtype"""
    dx = dwr = dbr = None
    if need_dx:
        dx = gemm(dy, _wview(compute_weight(w), wcols), trans_w=True, out_dtype=ctx_x.dtype)
    if need_dw and wcols is None and not bias_done:
        d = _deferred_for(w)
        if d is not None:
            d.add(w, b, dy, ctx_x)
            return dx, None, None
    if need_dw:
        gbuf, dwr = _grad_buf(w)
        gemm(dy, ctx_x, out=_wview(gbuf.view(w.shape[0], -1), wcols), trans_a=True, trans_w=True, accumulate=True)
        if b is not None and not bias_done:
            bbuf, dbr = _grad_buf(b)
            lib.genrl_colsum(_st(), _p(dy), _ld(dy), _p(bbuf), dy.shape[0], dy.shape[1], _CODE[dy.dtype])
    return dx, dwr, dbr


class _Linear(Function):
    @staticmethod
    def forward(ctx, x, w, b, out_fp32, wcols):
        x = as_compute(x)
        y = gemm(x, _wview(compute_weight(w), wcols), bias=b.detach() if b is not None else None,
                 out_dtype=torch.float32 if out_fp32 else _State.dtype)
        ctx.save_for_backward(x)
        ctx.w, ctx.b, ctx.wcols = w, b, wcols
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = as_compute(dy if dy.stride(-1) == 1 else dy.contiguous())
        dx, dw, db = _linear_backward(x, ctx.w, ctx.b, dy, ctx.needs_input_grad[0], ctx.needs_input_grad[1], wcols=ctx.wcols)
        return dx, dw, db, None, None


def linear(x, w, b=None, out_fp32=False, wcols=None):
    """nn.Linear forward on [M,K] rows (reference: every nn.Linear of dreamer_utils.py); wcols selects weight columns."""
    return _Linear.apply(x, w, b, out_fp32, wcols)


class _LnAct(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, act):
        x = as_compute(x)
        M, D = x.shape
        y = empty_act(M, D, x.device)
        stats = torch.empty(M, 2, device=x.device, dtype=torch.float32)
        lib.genrl_ln_act_fwd(_st(), _p(x), _ld(x), _p(gamma), _p(beta), eps, act, _p(y), _ld(y), _p(stats), M, D, _CODE[x.dtype])
        ctx.save_for_backward(x, stats)
        ctx.gamma, ctx.beta, ctx.act = gamma, beta, act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats = ctx.saved_tensors
        dy = as_compute(dy if dy.stride(-1) == 1 else dy.contiguous())
        M, D = x.shape
        dx = empty_act(M, D, x.device, x.dtype)
        gb = bb = dg = db = None
        if ctx.gamma.requires_grad:
            gb, dg = _grad_buf(ctx.gamma)
            bb, db = _grad_buf(ctx.beta)
        lib.genrl_ln_act_bwd(_st(), _p(dy), _ld(dy), _p(x), _ld(x), _p(ctx.gamma), _p(ctx.beta), _p(stats), ctx.act, _p(dx), _ld(dx),
                             _p(gb), _p(bb), None, M, D, _CODE[x.dtype])
        return dx, dg, db, None, None


def ln_act(x, gamma, beta, eps=1e-5, act=ACT_SILU):
    """LayerNorm over the last dim (+SiLU): NormLayer / ImgChLayerNorm + act of the reference."""
    return _LnAct.apply(x, gamma, beta, eps, act)


# Widest layer that takes the cluster-fused Linear+LayerNorm kernel.  Measured on B200 (tools/debug_fused_ln.py, 20 layers in a CUDA
# graph): N = 128 / 512 fused 10.5 / 13.9 us vs 12.7 / 15.9 us unfused; at N = 1024 (cluster of 8) the un-overlapped epilogue + two
# cluster barriers cost as much as the separate LayerNorm launch (20.2 vs 19.5 us), so the 1024-wide layers stay on the two-kernel path.
FUSED_LN_MAX_N = 512
CHAIN_MIN_ROWS = 64            # below that a 128-row stripe is mostly padding
# Measured on B200 (tools/chain_bench.py, profiles/r02_notes.md): at 36 000-38 400 rows the chain kernel's un-overlapped epilogue loses to the
# double-buffered persistent 256x128 GEMM + row LayerNorm (188 vs 140 us per 1024-wide layer); it pays at the rollout's 2 400 rows.
CHAIN_MAX_ROWS = int(__import__("os").environ.get("GENRL_CHAIN_MAX_ROWS", "8192"))


def fused_linear_ln_enabled():
    import os
    return os.environ.get("GENRL_FUSED_LN", "1") != "0"


def chain_supported(N, dtype=None):
    """the persistent cluster chain kernel (csrc/mlp_chain.cu): bf16, layer width a multiple of 256 up to 1024"""
    import os
    dtype = dtype or _State.dtype
    return dtype == torch.bfloat16 and N % 256 == 0 and 256 <= N <= 1024 and os.environ.get("GENRL_CHAIN", "1") != "0"


def _ptr_array(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(int(q) if q else 0) for q in ptrs])


def mlp_chain_fwd(x, weights, biases, gammas, betas, eps, act, pre, y, stats):
    """y_l = act(LN(y_{l-1} W_l^T + b_l)) for l = 0..L-1 in one launch.  x [M, K0] bf16; weights: compute-layout bf16 [N, K_l]; pre / y: lists of
    [M, N] bf16 output views sharing one leading dimension each; stats: list of [M, 2] fp32."""
    L = len(weights)
    M, K0 = x.shape
    N = weights[0].shape[0]
    ldw = (ctypes.c_longlong * L)(*[_ld(w) for w in weights])
    ldpre, ldy = _ld(pre[0]), _ld(y[0])
    assert all(_ld(t) == ldpre for t in pre) and all(_ld(t) == ldy for t in y)
    args = (_st(), M, N, L, _p(x), _ld(x), K0, _ptr_array([_p(w) for w in weights]), ldw,
            _ptr_array([_p(b) for b in biases]), _ptr_array([_p(g) for g in gammas]), _ptr_array([_p(b) for b in betas]),
            float(eps), int(act), _ptr_array([_p(t) for t in pre]), ldpre, _ptr_array([_p(t) for t in y]), ldy,
            _ptr_array([_p(t) for t in stats]))
    lib.genrl_mlp_chain_fwd(*args)


class _LinearLnAct(Function):
    """y = act(LN(x w^T + b)) as one node: Linear -> NormLayer -> SiLU (dreamer_utils.py:738-744, :339-346)."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, eps, act):
        x = as_compute(x)
        N = w.shape[0]
        if _State.dtype == torch.bfloat16 and fused_linear_ln_enabled() and N % 128 == 0 and 128 <= N <= FUSED_LN_MAX_N and x.shape[0] > 0:
            # one kernel: GEMM + bias + LayerNorm + SiLU, row statistics merged across the cluster of column-tile CTAs through DSMEM
            wc = compute_weight(w)
            if _ld(x) % 8 or x.data_ptr() % 16:
                x = cast_to(x, torch.bfloat16)
            M = x.shape[0]
            pre = empty_act(M, N, x.device)
            y = empty_act(M, N, x.device)
            stats = torch.empty(M, 2, device=x.device, dtype=torch.float32)
            lib.genrl_gemm_ln_fwd(_st(), M, N, x.shape[1], _p(x), _ld(x), _p(wc), _ld(wc), _p(b.detach() if b is not None else None),
                                  _p(gamma.detach()), _p(beta.detach()), float(eps), act, _p(pre), _ld(pre), _p(y), _ld(y), _p(stats))
            ctx.save_for_backward(x, pre, stats)
            ctx.w, ctx.b, ctx.gamma, ctx.beta, ctx.act = w, b, gamma, beta, act
            return y
        if chain_supported(N) and N > FUSED_LN_MAX_N and CHAIN_MIN_ROWS <= x.shape[0] <= CHAIN_MAX_ROWS:
            # one persistent cluster launch: GEMM + bias + LayerNorm + SiLU (clusters of N/256 CTAs, csrc/mlp_chain.cu)
            wc = compute_weight(w)
            if _ld(x) % 8 or x.data_ptr() % 16:
                x = cast_to(x, torch.bfloat16)
            M = x.shape[0]
            pre = empty_act(M, N, x.device)
            y = empty_act(M, N, x.device)
            stats = torch.empty(M, 2, device=x.device, dtype=torch.float32)
            mlp_chain_fwd(x, [wc], [b.detach() if b is not None else None], [gamma.detach()], [beta.detach()], eps, act, [pre], [y], [stats])
            ctx.save_for_backward(x, pre, stats)
            ctx.w, ctx.b, ctx.gamma, ctx.beta, ctx.act = w, b, gamma, beta, act
            return y
        pre = gemm(x, compute_weight(w), bias=b.detach() if b is not None else None, out_dtype=_State.dtype)
        M, D = pre.shape
        y = empty_act(M, D, x.device)
        stats = torch.empty(M, 2, device=x.device, dtype=torch.float32)
        lib.genrl_ln_act_fwd(_st(), _p(pre), _ld(pre), _p(gamma), _p(beta), eps, act, _p(y), _ld(y), _p(stats), M, D, _CODE[pre.dtype])
        ctx.save_for_backward(x, pre, stats)
        ctx.w, ctx.b, ctx.gamma, ctx.beta, ctx.act = w, b, gamma, beta, act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pre, stats = ctx.saved_tensors
        dy = as_compute(dy if dy.stride(-1) == 1 else dy.contiguous())
        M, D = pre.shape
        dpre = empty_act(M, D, x.device, pre.dtype)
        train = ctx.needs_input_grad[1]
        gb = bb = dg = db = lb = dlb = None
        d = _deferred_for(ctx.w) if train else None
        if d is not None:
            # rollout modules (the actor: H calls on N rows): frozen row pass now (10 us at 2400x1024 against 23 us with the fused column
            # partials, which only pay off from ~10^4 rows), parameter gradients once over all H*N rows when the scope is flushed
            lib.genrl_ln_act_bwd(_st(), _p(dy), _ld(dy), _p(pre), _ld(pre), _p(ctx.gamma), _p(ctx.beta), _p(stats), ctx.act, _p(dpre), _ld(dpre),
                                 None, None, None, M, D, _CODE[pre.dtype])
            d.add(ctx.w, ctx.b, dy, x, ln=(ctx.gamma, ctx.beta, ctx.act, pre, stats, dpre))
            dx = gemm(dpre, compute_weight(ctx.w), trans_w=True, out_dtype=x.dtype) if ctx.needs_input_grad[0] else None
            return dx, None, None, None, None, None, None
        if train:
            gb, dg = _grad_buf(ctx.gamma)
            bb, db = _grad_buf(ctx.beta)
            if ctx.b is not None:
                lb, dlb = _grad_buf(ctx.b)
        lib.genrl_ln_act_bwd(_st(), _p(dy), _ld(dy), _p(pre), _ld(pre), _p(ctx.gamma), _p(ctx.beta), _p(stats), ctx.act, _p(dpre), _ld(dpre),
                             _p(gb), _p(bb), _p(lb), M, D, _CODE[pre.dtype])
        dx, dw, _ = _linear_backward(x, ctx.w, ctx.b, dpre, ctx.needs_input_grad[0], train, bias_done=True)
        return dx, dw, dlb, dg, db, None, None


def linear_ln_act(x, w, b, gamma, beta, eps=1e-5, act=ACT_SILU):
    return _LinearLnAct.apply(x, w, b, gamma, beta, eps, act)


# ---------------------------------------------------------------------------------------------------------------
# concat / slicing helpers (explicit copies keep every GEMM operand K-contiguous)
# ---------------------------------------------------------------------------------------------------------------
class _Concat(Function):
    @staticmethod
    def forward(ctx, *parts):
        parts = [_check(p) for p in parts]
        widths = [p.shape[1] for p in parts]
        M = parts[0].shape[0]
        out = empty_act(M, sum(widths), parts[0].device)
        off = 0
        for p, wd in zip(parts, widths):
            copy2d(p, out[:, off:off + wd])
            off += wd
        ctx.widths = widths
        ctx.dtypes = [p.dtype for p in parts]
        return out

    @staticmethod
    def backward(ctx, dout):
        outs, off = [], 0
        for i, (wd, dt) in enumerate(zip(ctx.widths, ctx.dtypes)):
            if not ctx.needs_input_grad[i]:
                outs.append(None)
            else:
                sl = dout[:, off:off + wd]
                if dt == dout.dtype and dout.stride(1) == 1:
                    outs.append(sl)                     # column-slice view: every consumer kernel takes a leading dimension
                else:
                    g = torch.empty(dout.shape[0], wd, device=dout.device, dtype=dt)
                    copy2d(sl, g)
                    outs.append(g)
            off += wd
        return tuple(outs)


def concat(*parts):
    """[M, sum(widths)] in the compute dtype"""
    return _Concat.apply(*parts)


class _ShiftMask(Function):
    """out[n] = scale[n] * src[idx[n]] (idx<0 -> 0); adjoint uses the inverse map (ridx, rscale)."""

    @staticmethod
    def forward(ctx, src, idx, scale, ridx, rscale):
        src = _check(src)
        out = empty_act(idx.shape[0], src.shape[1], src.device)
        gather_rows(src, idx, scale, out)
        ctx.ridx, ctx.rscale, ctx.nsrc, ctx.sdtype = ridx, rscale, src.shape[0], src.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        dsrc = torch.empty(ctx.nsrc, dout.shape[1], device=dout.device, dtype=ctx.sdtype)
        gather_rows(dout if dout.stride(-1) == 1 else dout.contiguous(), ctx.ridx, ctx.rscale, dsrc)
        return dsrc, None, None, None, None


def shift_mask(src, idx, scale, ridx, rscale):
    return _ShiftMask.apply(src, idx, scale, ridx, rscale)


class _ToF32(Function):
    @staticmethod
    def forward(ctx, x):
        ctx.dt = x.dtype
        return cast_to(x, torch.float32) if x.dtype != torch.float32 else x

    @staticmethod
    def backward(ctx, g):
        return cast_to(g if g.stride(-1) == 1 else g.contiguous(), ctx.dt) if ctx.dt != torch.float32 else g


def to_f32(x):
    return _ToF32.apply(x)


class _ToCompute(Function):
    @staticmethod
    def forward(ctx, x):
        ctx.dt = x.dtype
        return as_compute(x)

    @staticmethod
    def backward(ctx, g):
        g = g if g.stride(-1) == 1 else g.contiguous()
        return cast_to(g, ctx.dt) if g.dtype != ctx.dt else g


def to_compute(x):
    return _ToCompute.apply(x)


# ---------------------------------------------------------------------------------------------------------------
# GRU cell tail
# ---------------------------------------------------------------------------------------------------------------
class _Gru(Function):
    """h' = gates(LN_3D(z), h): GRUCell.forward after the Linear (dreamer_utils.py:777-785). h, h' fp32 [M, D]."""

    @staticmethod
    def forward(ctx, z, gamma, beta, hprev):
        z = as_compute(z)
        M, D3 = z.shape
        D = D3 // 3
        hprev = hprev.contiguous().float()
        hout = torch.empty(M, D, device=z.device, dtype=torch.float32)
        stats = torch.empty(M, 2, device=z.device, dtype=torch.float32)
        lib.genrl_gru_fwd(_st(), _p(z), _ld(z), _p(gamma), _p(beta), _p(hprev), D, _p(hout), D, None, 0, _p(stats), M, D, _CODE[z.dtype])
        ctx.save_for_backward(z, hprev, stats)
        ctx.gamma, ctx.beta = gamma, beta
        return hout

    @staticmethod
    def backward(ctx, dh):
        z, hprev, stats = ctx.saved_tensors
        M, D3 = z.shape
        D = D3 // 3
        dh = dh.contiguous()
        dz = empty_act(M, D3, z.device, z.dtype)
        dhp = torch.empty(M, D, device=z.device, dtype=torch.float32)
        gb = bb = dg = db = None
        if ctx.gamma.requires_grad:
            gb, dg = _grad_buf(ctx.gamma)
            bb, db = _grad_buf(ctx.beta)
        lib.genrl_gru_bwd(_st(), _p(dh), D, _p(z), _ld(z), _p(ctx.gamma), _p(ctx.beta), _p(hprev), D, _p(stats), _p(dz), _ld(dz), _p(dhp), D,
                          _p(gb), _p(bb), M, D, _CODE[z.dtype])
        return dz, dg, db, dhp


def gru_gates(z, gamma, beta, hprev):
    return _Gru.apply(z, gamma, beta, hprev)


# ---------------------------------------------------------------------------------------------------------------
# categorical latent
# ---------------------------------------------------------------------------------------------------------------
class _CatSample(Function):
    """stoch = one_hot(argmax(pi/q)) with straight-through gradient (OneHotDist.sample / .mode)."""

    @staticmethod
    def forward(ctx, logit, noise, seed, offset, sample, S, K):
        logit = logit.contiguous()
        rows = logit.numel() // (S * K)
        stoch = torch.empty(rows, S * K, device=logit.device, dtype=torch.float32)
        idx = torch.empty(rows, S, device=logit.device, dtype=torch.int32)
        step_dev = noise_step_ptr(logit.device) if (noise is None and sample) else None
        lib.genrl_cat_sample_fwd(_st(), _p(logit), _p(noise), seed, offset, step_dev, int(sample), None, 0, _p(stoch), S * K, _p(idx), rows, S, K, F32)
        ctx.save_for_backward(logit)
        ctx.sample, ctx.S, ctx.K = sample, S, K
        ctx.mark_non_differentiable(idx)
        return stoch, idx

    @staticmethod
    def backward(ctx, dstoch, _didx):
        (logit,) = ctx.saved_tensors
        S, K = ctx.S, ctx.K
        rows = logit.numel() // (S * K)
        dstoch = dstoch if dstoch.stride(-1) == 1 else dstoch.contiguous()
        dlogit = torch.empty_like(logit)
        lib.genrl_cat_sample_bwd(_st(), _p(logit), _p(dstoch), _ld(dstoch), int(ctx.sample), _p(dlogit), 0, rows, S, K, _CODE[dstoch.dtype])
        return dlogit, None, None, None, None, None, None


def cat_sample(logit, S, K, noise=None, seed=0, offset=0, sample=True):
    """logit fp32 [rows, S*K] -> (stoch fp32 one-hot [rows, S*K], idx int32 [rows, S])."""
    return _CatSample.apply(logit, noise, seed, offset, sample, S, K)


class _KlBalance(Function):
    @staticmethod
    def forward(ctx, lhs, rhs, free, w_lhs, w_rhs, S, K):
        lhs, rhs = lhs.contiguous(), rhs.contiguous()
        rows = lhs.numel() // (S * K)
        value = torch.empty(rows, device=lhs.device, dtype=torch.float32)
        out4 = torch.empty(4, device=lhs.device, dtype=torch.float32)
        lib.genrl_kl_fwd(_st(), _p(lhs), _p(rhs), free, _p(value), _p(out4), rows, S, K)
        ctx.save_for_backward(lhs, rhs, value)
        ctx.cfg = (free, w_lhs, w_rhs, S, K)
        ctx.mark_non_differentiable(value)
        stats = out4.detach().clone()
        ctx.mark_non_differentiable(stats)
        return out4[0:1].clone(), value, stats

    @staticmethod
    def backward(ctx, gloss, _gv, _gs):
        lhs, rhs, value = ctx.saved_tensors
        free, w_lhs, w_rhs, S, K = ctx.cfg
        rows = value.numel()
        dl, dr = torch.empty_like(lhs), torch.empty_like(rhs)
        gloss = gloss.contiguous().float()
        lib.genrl_kl_bwd(_st(), _p(lhs), _p(rhs), _p(value), free, w_lhs, w_rhs, _p(gloss), _p(dl), _p(dr), rows, S, K)
        return dl, dr, None, None, None, None, None


def kl_balance(post_logit, prior_logit, S, K, forward, balance, free):
    """EnsembleRSSM.kl_loss. Returns (loss [1], value [rows], stats [4] = loss, mean value, ent(lhs), ent(rhs))."""
    lhs, rhs = (prior_logit, post_logit) if forward else (post_logit, prior_logit)
    mix = balance if forward else (1 - balance)
    if balance == 0.5:
        w_l, w_r = 1.0, 1.0           # single KL term, gradients to both sides
    else:
        w_l, w_r = mix, 1 - mix
    return _KlBalance.apply(lhs, rhs, float(free), float(w_l), float(w_r), S, K)


# ---------------------------------------------------------------------------------------------------------------
# two-hot head
# ---------------------------------------------------------------------------------------------------------------
class _TwoHot(Function):
    @staticmethod
    def forward(ctx, logits, x, mode):
        logits = _check(logits)
        rows = logits.shape[0]
        out = torch.empty(rows, device=logits.device, dtype=torch.float32)
        xx = None if x is None else x.contiguous().float().reshape(-1)
        lib.genrl_twohot_fwd(_st(), _p(logits), _ld(logits), _p(xx), mode, _p(out), rows)
        ctx.save_for_backward(logits, xx)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, dout):
        logits, xx = ctx.saved_tensors
        rows = logits.shape[0]
        dl = torch.empty(rows, 255, device=logits.device, dtype=torch.float32)
        dout = dout.contiguous()
        lib.genrl_twohot_bwd(_st(), _p(logits), _ld(logits), _p(xx), ctx.mode, _p(dout), 1.0, _p(dl), 255, rows)
        return dl, None, None


def twohot_logprob(logits, x):
    """TwoHotDist(logits).log_prob(x); logits fp32 [rows,255], x [rows] -> [rows]"""
    return _TwoHot.apply(logits, x, 0)


def twohot_mean(logits):
    return _TwoHot.apply(logits, None, 1)


# ---------------------------------------------------------------------------------------------------------------
# actor head
# ---------------------------------------------------------------------------------------------------------------
class _ActorHead(Function):
    @staticmethod
    def forward(ctx, out, eps, use_mean, A):
        out = _check(out)
        rows = out.shape[0]
        action = torch.empty(rows, A, device=out.device, dtype=torch.float32)
        ent = torch.empty(rows, device=out.device, dtype=torch.float32)
        lib.genrl_actor_head_fwd(_st(), _p(out), _ld(out), _p(eps), int(use_mean), _p(action), A, None, 0, _p(ent), rows, A, F32)
        ctx.save_for_backward(out, eps)
        ctx.A, ctx.use_mean = A, use_mean
        return action, ent

    @staticmethod
    def backward(ctx, daction, dent):
        out, eps = ctx.saved_tensors
        rows, A = out.shape[0], ctx.A
        dout = torch.empty(rows, 2 * A, device=out.device, dtype=torch.float32)
        daction = None if daction is None else daction.contiguous().float()
        dent = None if dent is None else dent.contiguous().float()
        if eps is None:
            eps = torch.zeros(rows, A, device=out.device, dtype=torch.float32)
        lib.genrl_actor_head_bwd(_st(), _p(out), _ld(out), _p(eps), _p(daction), A, _p(dent), 1.0 if dent is not None else 0.0, _p(dout), 2 * A,
                                 F32, rows, A, F32)
        return dout, None, None, None


def actor_head(out, eps, A, use_mean=False):
    """out fp32 [rows, 2A] -> (action fp32 [rows, A], entropy [rows])"""
    return _ActorHead.apply(out, eps, use_mean, A)


# ---------------------------------------------------------------------------------------------------------------
# lambda return
# ---------------------------------------------------------------------------------------------------------------
class _LambdaReturn(Function):
    @staticmethod
    def forward(ctx, reward, value, gamma, lam):
        reward, value = reward.contiguous().float(), value.contiguous().float()
        H = reward.shape[0] - 1
        N = reward.numel() // (H + 1)
        target = torch.empty((H,) + tuple(reward.shape[1:]), device=reward.device, dtype=torch.float32)
        lib.genrl_lambda_return_fwd(_st(), _p(reward), _p(value), gamma, lam, _p(target), H, N)
        ctx.cfg = (gamma, lam, H, N, reward.shape)
        return target

    @staticmethod
    def backward(ctx, dt):
        gamma, lam, H, N, shape = ctx.cfg
        dt = dt.contiguous().float()
        dr = torch.empty(shape, device=dt.device, dtype=torch.float32)
        dv = torch.empty(shape, device=dt.device, dtype=torch.float32)
        lib.genrl_lambda_return_bwd(_st(), _p(dt), None, gamma, lam, _p(dr), _p(dv), H, N)
        return dr, dv, None, None


def lambda_return(reward, value, gamma, lam):
    """reward, value [H+1, N, 1] -> target [H, N, 1] (bootstrap = value[-1])"""
    return _LambdaReturn.apply(reward, value, float(gamma), float(lam))


# ---------------------------------------------------------------------------------------------------------------
# video-text reward (max-cosine score with sliding-window alignment)
# ---------------------------------------------------------------------------------------------------------------
class _VlmReward(Function):
    """reward[t,n] = maxcos(pa[t,n], pt[max(t - b_n, 0), n]), b_n = best alignment offset (tools/genrl_utils.py:250-277, :344-366)."""

    @staticmethod
    def forward(ctx, pa, pt, n_frames, align):
        T, N, E = pa.shape
        pa, pt = pa.contiguous(), pt.contiguous().to(pa.dtype)
        reward = torch.empty(T, N, device=pa.device, dtype=torch.float32)
        shift = torch.empty(N, device=pa.device, dtype=torch.int32)
        lib.genrl_vlm_reward_fwd(_st(), _p(pa), _p(pt), _p(reward), _p(shift), T, N, E, int(n_frames), int(bool(align)), _CODE[pa.dtype])
        ctx.save_for_backward(pa, pt, shift)
        ctx.mark_non_differentiable(shift)
        return reward, shift

    @staticmethod
    def backward(ctx, dreward, _dshift):
        pa, pt, shift = ctx.saved_tensors
        T, N, E = pa.shape
        dpa = torch.empty_like(pa)
        lib.genrl_vlm_reward_bwd(_st(), _p(pa), _p(pt), _p(shift), _p(dreward.contiguous().float()), _p(dpa), T, N, E, _CODE[pa.dtype])
        return dpa, None, None, None


def vlm_reward(pa, pt, n_frames=8, align=True):
    """pa, pt: [T, N, E] projections (compute dtype) -> (reward fp32 [T, N], shift int32 [N])"""
    return _VlmReward.apply(pa, pt, n_frames, align)


# ---------------------------------------------------------------------------------------------------------------
# conv stacks
# ---------------------------------------------------------------------------------------------------------------

# ---------------------------------------------------------------------------------------------------------------
# implicit-GEMM convolutions (conv_igemm.cu): bf16 tensor-core mode, channel counts that are multiples of 8
# ---------------------------------------------------------------------------------------------------------------
def _pad16(n):
    return (n + 15) // 16 * 16


def conv_igemm_enabled(C):
    import os
    return _State.dtype == torch.bfloat16 and C % 8 == 0 and os.environ.get("GENRL_CONV_IGEMM", "1") != "0"


def _rows_contig(t, C):
    """NHWC rows [M, C] with ld == C (the gather addresses pixels as in[pix*C + c])"""
    t = _check(t)
    if t.dtype != torch.bfloat16 or _ld(t) != C or t.data_ptr() % 16:
        out = torch.empty(t.shape[0], C, device=t.device, dtype=torch.bfloat16)
        copy2d(t, out)
        return out
    return t


def _igemm_gather(x, N, IH, IW, C, RH, RW, k, wm, ncols, bias, out=None):
    """out[(n,yy,xx), col] = sum_{ky,kx,c} x[n, 2yy+ky, 2xx+kx, c] wm[col, (ky*k+kx)*C + c] (+bias): Conv2d fwd / ConvT dgrad"""
    if out is None:
        out = torch.empty(N * RH * RW, ncols, device=x.device, dtype=torch.bfloat16)
    lib.genrl_conv_igemm_fwd(_st(), _p(x), N, IH, IW, C, RH, RW, 2, 1, k, k, _p(wm), _ld(wm), _pad16(ncols), ncols, 1, RH, RW, _p(out), 0, _p(bias))
    return out


def _igemm_subpixel(x, N, IH, IW, C, OH, OW, k, w_master, A, B, bias, nchw=False):
    """out[n, oy, ox, b] = sum x[n, iy, ix, a] W4[a][b][oy-2iy][ox-2ix] (+bias): ConvTranspose2d fwd (W4 = its weight) / Conv2d dgrad"""
    T = (k + 1) // 2
    Cpad = _pad8(B)
    wq = torch.empty(4 * Cpad, T * T * A, device=x.device, dtype=torch.bfloat16)
    lib.genrl_conv_pack_parity(_st(), _p(w_master), _p(wq), A, B, k, Cpad)
    if nchw:
        out = torch.empty(N, B, OH, OW, device=x.device, dtype=torch.float32)
    else:
        out = torch.empty(N * OH * OW, B, device=x.device, dtype=torch.bfloat16)
    lib.genrl_conv_igemm_fwd(_st(), _p(x), N, IH, IW, C, (OH + 1) // 2, (OW + 1) // 2, 1, -1, T, T, _p(wq), T * T * A, Cpad, B, 4, OH, OW, _p(out),
                             int(nchw), _p(bias))
    return out


def _igemm_wgrad(src, N, IH, IW, C, RH, RW, k, S, Cs, w, A, B):
    """grad(w)[a][b][ky][kx] += sum_rows S[row, a] src[n, 2yy+ky, 2xx+kx, b]   (src channel count C >= B, zero padded)"""
    gbuf, ret = _grad_buf(w)
    dwp = torch.zeros(Cs, k * k * C, device=src.device, dtype=torch.float32)
    lib.genrl_conv_igemm_wgrad(_st(), _p(src), N, IH, IW, C, RH, RW, 2, k, k, _p(S), _ld(S), Cs, _p(dwp), k * k * C)
    lib.genrl_conv_unpack_wgrad(_st(), _p(dwp), k * k * C, _p(gbuf), A, B, k, C)
    return ret


class _Conv2dS2(Function):
    """Conv2d(k, stride 2, no padding) on NHWC rows; x is uint8 NCHW for the first layer (preprocessing fused)."""

    @staticmethod
    def forward(ctx, x, w, b, geom):
        N, H, W, C, k, Ho, Wo, u8 = geom
        Cout = w.shape[0]
        Kc = k * k * C
        dt = _State.dtype
        ctx.igemm = (not u8) and conv_igemm_enabled(C) and Cout % 8 == 0
        ctx.igemm_u8 = u8 and C <= 8 and conv_igemm_enabled(8) and Cout % 8 == 0
        if ctx.igemm_u8:
            # first encoder layer: uint8 NCHW -> preprocessed NHWC8 bf16 (zero channels 3..7), then the same implicit GEMM (K = k*k*8)
            x8 = torch.empty(N * H * W, 8, device=x.device, dtype=torch.bfloat16)
            lib.genrl_u8_nchw_to_nhwc8(_st(), _p(x), _p(x8), N, C, H, W)
            wg = torch.empty(Cout, k * k * 8, device=x.device, dtype=torch.bfloat16)
            lib.genrl_conv_pack_gather(_st(), _p(w.detach()), _p(wg), Cout, C, k, 8)
            y = _igemm_gather(x8, N, H, W, 8, Ho, Wo, k, wg, Cout, b.detach())
            ctx.save_for_backward(x)
            ctx.w, ctx.b, ctx.geom = w, b, geom
            return y
        if ctx.igemm:
            x = _rows_contig(x, C)
            y = _igemm_gather(x, N, H, W, C, Ho, Wo, k, packed_conv_weight(w), Cout, b.detach())
            ctx.save_for_backward(x)
            ctx.w, ctx.b, ctx.geom = w, b, geom
            return y
        col = empty_act(N * Ho * Wo, Kc, w.device, dt)
        lib.genrl_im2col(_st(), _p(x), -1 if u8 else _CODE[x.dtype], _p(col), _ld(col), _CODE[dt], N, H, W, C, k, Ho, Wo)
        wp = packed_conv_weight(w)
        y = gemm(col, wp, bias=b.detach(), out_dtype=dt)
        ctx.save_for_backward(col)
        ctx.w, ctx.b, ctx.geom = w, b, geom
        return y

    @staticmethod
    def backward(ctx, dy):
        (col,) = ctx.saved_tensors
        N, H, W, C, k, Ho, Wo, u8 = ctx.geom
        w, b = ctx.w, ctx.b
        dy = as_compute(dy if dy.stride(-1) == 1 else dy.contiguous())
        dx = dw = db = None
        if ctx.igemm_u8:
            Cout = w.shape[0]
            dy = _rows_contig(dy, Cout)
            if w.requires_grad:
                x8 = torch.empty(N * H * W, 8, device=dy.device, dtype=torch.bfloat16)           # recomputed: cheaper than keeping 157 MB alive
                lib.genrl_u8_nchw_to_nhwc8(_st(), _p(col), _p(x8), N, C, H, W)
                dw = _igemm_wgrad(x8, N, H, W, 8, Ho, Wo, k, dy, Cout, w, Cout, C)
                bbuf, db = _grad_buf(b)
                lib.genrl_colsum(_st(), _p(dy), _ld(dy), _p(bbuf), dy.shape[0], dy.shape[1], _CODE[dy.dtype])
            return None, dw, db, None
        if ctx.igemm:
            x, Cout = col, w.shape[0]
            dy = _rows_contig(dy, Cout)
            if w.requires_grad:
                dw = _igemm_wgrad(x, N, H, W, C, Ho, Wo, k, dy, Cout, w, Cout, C)
                bbuf, db = _grad_buf(b)
                lib.genrl_colsum(_st(), _p(dy), _ld(dy), _p(bbuf), dy.shape[0], dy.shape[1], _CODE[dy.dtype])
            if ctx.needs_input_grad[0]:
                dx = _igemm_subpixel(dy, N, Ho, Wo, Cout, H, W, k, w.detach(), Cout, C, None)
            return dx, dw, db, None
        if w.requires_grad:
            gbuf, dw = _grad_buf(w)
            dwp = gemm(dy, col, trans_a=True, trans_w=True, out_dtype=torch.float32)            # [Cout, k*k*C]
            lib.genrl_permute_abp(_st(), _p(dwp), F32, _p(gbuf), F32, w.shape[0], k * k, C, 1)     # -> [Cout, C, k*k] +=
            bbuf, db = _grad_buf(b)
            lib.genrl_colsum(_st(), _p(dy), _ld(dy), _p(bbuf), dy.shape[0], dy.shape[1], _CODE[dy.dtype])
        if ctx.needs_input_grad[0] and not u8:
            dcol = gemm(dy, packed_conv_weight(w), trans_w=True, out_dtype=dy.dtype)             # [M, k*k*C]
            dx = torch.empty(N * H * W, C, device=dy.device, dtype=dy.dtype)
            lib.genrl_col2im(_st(), _p(dcol), _ld(dcol), _CODE[dcol.dtype], None, _p(dx), _CODE[dx.dtype], 0, N, Ho, Wo, C, k, H, W)
        return dx, dw, db, None


def conv2d_s2(x, w, b, N, H, W, k, u8=False):
    """x: [N*H*W, C] NHWC rows (or uint8 [N,C,H,W]); returns [N*Ho*Wo, Cout] rows, Ho = (H-k)//2+1."""
    C = w.shape[1]
    Ho, Wo = (H - k) // 2 + 1, (W - k) // 2 + 1
    return _Conv2dS2.apply(x, w, b, (N, H, W, C, k, Ho, Wo, u8)), Ho, Wo


class _ConvT2dS2(Function):
    """ConvTranspose2d(k, stride 2): y = col2im(x Wt) + b. Output NHWC rows, or fp32 NCHW for the last layer."""

    @staticmethod
    def forward(ctx, x, w, b, geom):
        N, Hi, Wi, k, Ho, Wo, nchw = geom
        Cin, Cout = w.shape[0], w.shape[1]
        x = as_compute(x)
        dt = _State.dtype
        ctx.igemm = Hi > 1 and conv_igemm_enabled(Cin) and (nchw or Cout % 8 == 0)
        if ctx.igemm:
            x = _rows_contig(x, Cin)
            y = _igemm_subpixel(x, N, Hi, Wi, Cin, Ho, Wo, k, w.detach(), Cin, Cout, b.detach(), nchw=nchw)
            ctx.save_for_backward(x)
            ctx.w, ctx.b, ctx.geom = w, b, geom
            return y
        wt = packed_conv_weight(w)                                         # [Cin, k*k*Cout]
        col = gemm(x, wt, trans_w=True, out_dtype=dt)                      # [M, k*k*Cout]
        if nchw:
            y = torch.empty(N, Cout, Ho, Wo, device=x.device, dtype=torch.float32)
        else:
            y = torch.empty(N * Ho * Wo, Cout, device=x.device, dtype=dt)
        lib.genrl_col2im(_st(), _p(col), _ld(col), _CODE[dt], _p(b.detach()), _p(y), _CODE[y.dtype], int(nchw), N, Hi, Wi, Cout, k, Ho, Wo)
        ctx.save_for_backward(x)
        ctx.w, ctx.b, ctx.geom = w, b, geom
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        N, Hi, Wi, k, Ho, Wo, nchw = ctx.geom
        w, b = ctx.w, ctx.b
        Cin, Cout = w.shape[0], w.shape[1]
        dt = x.dtype
        if ctx.igemm:
            Cp = _pad8(Cout)
            if nchw:
                dyr = torch.empty(N * Ho * Wo, Cp, device=x.device, dtype=torch.bfloat16)
                lib.genrl_nchw_to_nhwc_pad(_st(), _p(dy.contiguous().float()), _p(dyr), N, Cout, Ho, Wo, Cp)
            else:
                dyr = _rows_contig(dy if dy.stride(-1) == 1 else dy.contiguous(), Cout)
            dx = dw = db = None
            if w.requires_grad:
                dw = _igemm_wgrad(dyr, N, Ho, Wo, Cp, Hi, Wi, k, x, Cin, w, Cin, Cout)
                bbuf, db = _grad_buf(b)
                lib.genrl_colsum(_st(), _p(dyr), Cp, _p(bbuf), dyr.shape[0], Cout, _CODE[dyr.dtype])
            if ctx.needs_input_grad[0]:
                if Cp == Cout:
                    wg = packed_conv_weight(w)                                                     # [Cin, k*k*Cout]
                else:
                    wg = torch.empty(Cin, k * k * Cp, device=x.device, dtype=torch.bfloat16)
                    lib.genrl_conv_pack_gather(_st(), _p(w.detach()), _p(wg), Cin, Cout, k, Cp)
                dx = _igemm_gather(dyr, N, Ho, Wo, Cp, Hi, Wi, k, wg, Cin, None)
            return dx, dw, db, None
        if nchw:
            dyr = torch.empty(N * Ho * Wo, Cout, device=x.device, dtype=dt)
            lib.genrl_nchw_to_nhwc(_st(), _p(dy.contiguous().float()), _p(dyr), _CODE[dt], N, Cout, Ho, Wo)
        else:
            dyr = as_compute(dy if dy.stride(-1) == 1 else dy.contiguous())
        dcol = empty_act(N * Hi * Wi, k * k * Cout, x.device, dt)
        lib.genrl_im2col(_st(), _p(dyr), _CODE[dt], _p(dcol), _ld(dcol), _CODE[dt], N, Ho, Wo, Cout, k, Hi, Wi)
        dx = dw = db = None
        if w.requires_grad:
            gbuf, dw = _grad_buf(w)
            dwt = gemm(x, dcol, trans_a=True, trans_w=True, out_dtype=torch.float32)            # [Cin, k*k*Cout]
            lib.genrl_permute_abp(_st(), _p(dwt), F32, _p(gbuf), F32, Cin, k * k, Cout, 1)        # -> [Cin, Cout, k*k] +=
            bbuf, db = _grad_buf(b)
            lib.genrl_colsum(_st(), _p(dyr), _ld(dyr), _p(bbuf), dyr.shape[0], Cout, _CODE[dt])
        if ctx.needs_input_grad[0]:
            dx = gemm(dcol, packed_conv_weight(w), out_dtype=dt)                                  # [M, Cin]
        return dx, dw, db, None


def conv_transpose2d_s2(x, w, b, N, Hi, Wi, k, nchw_out=False):
    Ho, Wo = (Hi - 1) * 2 + k, (Wi - 1) * 2 + k
    return _ConvT2dS2.apply(x, w, b, (N, Hi, Wi, k, Ho, Wo, nchw_out)), Ho, Wo


class _PermuteABP(Function):
    """[A, B, P] -> [A, P, B] rows (used for NHWC <-> (C,H,W) flatten order of the encoder output)."""

    @staticmethod
    def forward(ctx, x, A, B, P):
        x = x.contiguous()
        out = torch.empty(A, P * B, device=x.device, dtype=x.dtype)
        lib.genrl_permute_abp(_st(), _p(x), _CODE[x.dtype], _p(out), _CODE[x.dtype], A, B, P, 0)
        ctx.cfg = (A, B, P)
        return out

    @staticmethod
    def backward(ctx, g):
        A, B, P = ctx.cfg
        g = g.contiguous()
        out = torch.empty(A, B * P, device=g.device, dtype=g.dtype)
        lib.genrl_permute_abp(_st(), _p(g), _CODE[g.dtype], _p(out), _CODE[g.dtype], A, P, B, 0)
        return out, None, None, None


def permute_abp(x, A, B, P):
    return _PermuteABP.apply(x, A, B, P)


class _MseLogProb(Function):
    @staticmethod
    def forward(ctx, mean, obs_u8):
        mean = mean.contiguous()
        obs_u8 = obs_u8.contiguous()
        rows = mean.shape[0]
        E = mean.numel() // rows
        like = torch.empty(rows, device=mean.device, dtype=torch.float32)
        lib.genrl_mse_fwd(_st(), _p(mean), _p(obs_u8), _p(like), rows, E)
        ctx.save_for_backward(mean, obs_u8)
        return like

    @staticmethod
    def backward(ctx, dlike):
        mean, obs = ctx.saved_tensors
        rows = mean.shape[0]
        E = mean.numel() // rows
        dmean = torch.empty_like(mean)
        dlike = dlike.contiguous().float()
        lib.genrl_mse_bwd(_st(), _p(mean), _p(obs), _p(dlike), None, 1.0, _p(dmean), rows, E)
        return dmean, None


def mse_logprob(mean, obs_u8):
    """MSEDist(mean).log_prob(obs/255-0.5) summed over (C,H,W): mean fp32 [rows,C,H,W], obs uint8 -> [rows]"""
    return _MseLogProb.apply(mean, obs_u8)


class _MeanScale(Function):
    """scale * mean(x) as a [1] fp32 tensor (losses are plain means of per-row terms)."""

    @staticmethod
    def forward(ctx, x, scale):
        ctx.n, ctx.scale, ctx.shape = x.numel(), scale, x.shape
        return mean_of(x.float(), scale)

    @staticmethod
    def backward(ctx, g):
        return (g * (ctx.scale / ctx.n)).expand(ctx.shape).contiguous(), None


def scaled_mean(x, scale=1.0):
    return _MeanScale.apply(x, float(scale))


# ---------------------------------------------------------------------------------------------------------------
# persistent RSSM scan
# ---------------------------------------------------------------------------------------------------------------
_WS = {}


def _scan_workspace(B, T, D, device):
    # one workspace per stream: the world-model and connector scans may run concurrently on two streams
    key = (B, T, D, str(device), _st())
    ws = _WS.get(key)
    if ws is None:
        nbytes = lib.load().genrl_rssm_scan_workspace_bytes(B, T, D)
        ws = torch.zeros((nbytes + 3) // 4, device=device, dtype=torch.int32)
        _WS[key] = ws
    return ws


class _GruScan(Function):
    """deter[b,t] = GRU(gx[b,t] + W_h (m[b,t] deter[b,t-1])) for t = 0..T-1 in ONE persistent cooperative kernel."""

    @staticmethod
    def forward(ctx, gx, w_gru, gamma, beta, mask, h0, B, T, Hd):
        gx = gx.contiguous().float()
        D = gx.shape[1] // 3
        dev = gx.device
        wh = w_gru.detach()[:, Hd:]
        deter = torch.empty(B * T, D, device=dev, dtype=torch.float32)
        hprev = empty_act(B * T, D, dev)                 # m * h_{t-1} in the compute dtype: operand of the W_h wgrad GEMM
        z = torch.empty(B * T, 3 * D, device=dev, dtype=torch.float32)
        stats = torch.empty(B * T, 2, device=dev, dtype=torch.float32)
        h0c = None if h0 is None else h0.contiguous().float()
        ws = _scan_workspace(B, T, D, dev)
        lib.genrl_rssm_scan_fwd(_st(), B, T, D, _p(gx), _p(wh), w_gru.stride(0), _p(gamma), _p(beta), _p(mask), _p(h0c), _p(deter), None, 0,
                                _p(hprev), _ld(hprev), _CODE[hprev.dtype], _p(z), _p(stats), _p(ws))
        ctx.save_for_backward(z, stats, deter, hprev, mask, h0c)
        ctx.w, ctx.gamma, ctx.beta, ctx.cfg = w_gru, gamma, beta, (B, T, D, Hd)
        return deter

    @staticmethod
    def backward(ctx, ddeter):
        z, stats, deter, hprev, mask, h0c = ctx.saved_tensors
        B, T, D, Hd = ctx.cfg
        w, dev = ctx.w, z.device
        wh = w.detach()[:, Hd:]
        ddeter = ddeter.contiguous().float()
        dz = empty_act(B * T, 3 * D, dev)
        dh0 = torch.empty(B, D, device=dev, dtype=torch.float32) if (h0c is not None and ctx.needs_input_grad[5]) else None
        train = w.requires_grad
        gb = bb = dg = db = None
        if train:
            gb, dg = _grad_buf(ctx.gamma)
            bb, db = _grad_buf(ctx.beta)
        ws = _scan_workspace(B, T, D, dev)
        lib.genrl_rssm_scan_bwd(_st(), B, T, D, _p(ddeter), _p(z), _p(stats), _p(wh), w.stride(0), _p(ctx.gamma), _p(ctx.beta), _p(mask),
                                _p(deter), _p(h0c), _p(dz), _ld(dz), _CODE[dz.dtype], _p(dh0), _p(gb), _p(bb), _p(ws))
        dw = None
        if train:
            gbuf, dw_ret = _grad_buf(w)
            # dW_gru[:, Hd:] += dz^T (m * h_prev)
            gemm(dz, hprev, out=gbuf[:, Hd:], trans_a=True, trans_w=True, accumulate=True)
            dw = dw_ret
        dgx = dz if dz.dtype == torch.float32 else cast_to(dz, torch.float32)
        ctx.dz_compute = dz
        return dgx, dw, dg, db, None, dh0, None, None, None


_SM_COUNT = None


def scan_supported(D):
    """the persistent scan keeps 8 hidden units (24 rows of W_h) per CTA resident in shared memory: D/8 CTAs must be co-resident"""
    global _SM_COUNT
    if _SM_COUNT is None:
        sm, ma, mi, opt = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong()
        lib.load().genrl_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(opt))
        _SM_COUNT = max(int(sm.value), 1)
    return D % 8 == 0 and D // 8 <= _SM_COUNT


def _gru_scan_stepwise(gx, w_gru, gamma, beta, mask, h0, B, T, Hd):
    """Op-level recurrence for hidden sizes the resident-weight scan cannot hold (D > 8 x #SMs, e.g. BASELINE configs[4] with D = 4096):
    per step one W_h GEMM (weights streamed from L2/HBM) + the GRU gate kernel, sequenced by autograd."""
    D = gx.shape[1] // 3
    gx3 = gx.reshape(B, T, 3 * D)
    m = None if mask is None else mask.reshape(B, T, 1).float()
    h = torch.zeros(B, D, device=gx.device, dtype=torch.float32) if h0 is None else h0.float()
    outs = []
    for t in range(T):
        hp = h if m is None else h * m[:, t]
        z = gx3[:, t].float() + linear(to_compute(hp), w_gru, None, out_fp32=True, wcols=(Hd, Hd + D))
        h = gru_gates(z, gamma, beta, hp)
        outs.append(h)
    return torch.stack(outs, 1).reshape(B * T, D)


class _GruScanStream(Function):
    """The same recurrence for hidden sizes whose W_h cannot stay resident in shared memory (BASELINE configs[4], D = 4096): the forward is ONE
    persistent cooperative kernel that streams the bf16 weights once per time step (csrc/rssm_scan_stream.cu), the backward a second one that
    streams the transposed matrix (D % 512 == 0; other sizes walk the steps with the GRU-gate backward kernel and one M = B dgrad GEMM per
    step); the W_h gradient is one GEMM over all B*T rows afterwards."""

    @staticmethod
    def forward(ctx, gx, w_gru, gamma, beta, mask, h0, B, T, Hd):
        gx = gx.contiguous().float()
        D = gx.shape[1] // 3
        dev = gx.device
        wc = compute_weight(w_gru)                       # bf16 shadow [3D, Hd + D]
        wh = wc[:, Hd:Hd + D]
        deter = torch.empty(B * T, D, device=dev, dtype=torch.float32)
        hprev = torch.empty(B * T, D, device=dev, dtype=torch.bfloat16)
        z = torch.empty(B * T, 3 * D, device=dev, dtype=torch.bfloat16)
        stats = torch.empty(B * T, 2, device=dev, dtype=torch.float32)
        h0c = None if h0 is None else h0.contiguous().float()
        maskc = None if mask is None else mask.contiguous().float()
        ws = _scan_workspace(B, T, D, dev)
        lib.genrl_rssm_scan_stream_fwd(_st(), B, T, D, _p(gx), _p(wh), _ld(wh), _p(gamma), _p(beta), _p(maskc), _p(h0c), _p(deter), _p(hprev), D,
                                       _p(z), 3 * D, _p(stats), _p(ws))
        ctx.save_for_backward(z, stats, deter, hprev, maskc, h0c)
        ctx.w, ctx.gamma, ctx.beta, ctx.cfg = w_gru, gamma, beta, (B, T, D, Hd)
        return deter

    @staticmethod
    def backward(ctx, ddeter):
        z, stats, deter, hprev, mask, h0c = ctx.saved_tensors
        B, T, D, Hd = ctx.cfg
        w, dev = ctx.w, z.device
        with use_dtype(torch.bfloat16):
            wh = compute_weight(w)[:, Hd:Hd + D]
            dd3 = ddeter.contiguous().float().view(B, T, D)
            det3, z3, st3 = deter.view(B, T, D), z.view(B, T, 3 * D), stats.view(B, T, 2)
            m3 = None if mask is None else mask.view(B, T, 1)
            dz = torch.empty(B * T, 3 * D, device=dev, dtype=torch.bfloat16)
            dz3 = dz.view(B, T, 3 * D)
            train = w.requires_grad
            gb = bb = dg = db = None
            if train:
                gb, dg = _grad_buf(ctx.gamma)
                bb, db = _grad_buf(ctx.beta)
            dh0 = None
            want_dh0 = h0c is not None and ctx.needs_input_grad[5]
            # one persistent kernel that streams W_h^T once per step (csrc/rssm_scan_stream.cu); 2 = shape not eligible -> step by step below
            rc = 2
            if D % 512 == 0:
                wht = wh.t().contiguous()
                dh0_buf = torch.empty(B, D, device=dev, dtype=torch.float32) if want_dh0 else None
                rc = lib.load().genrl_rssm_scan_stream_bwd(_st(), B, T, D, _p(dd3), _p(z), 3 * D, _p(stats), _p(wht), 3 * D, _p(ctx.gamma),
                                                           _p(ctx.beta), _p(mask), _p(deter), _p(h0c), _p(dz), 3 * D, _p(dh0_buf), _p(gb), _p(bb),
                                                           _p(_scan_workspace(B, T, D, dev)))
            if rc == 0:
                lib.launches += 1
                dh0 = dh0_buf
            elif rc != 2:
                raise GenrlError(f"genrl_rssm_scan_stream_bwd failed: {lib.load().genrl_last_error().decode()}")
            carry = None
            for t in (range(T - 1, -1, -1) if rc == 2 else ()):
                dh = dd3[:, t].contiguous() if carry is None else dd3[:, t] + carry
                hp = det3[:, t - 1] if t > 0 else (h0c if h0c is not None else torch.zeros(B, D, device=dev, dtype=torch.float32))
                if m3 is not None:
                    hp = hp * m3[:, t]
                hp = hp.contiguous()
                st_t = st3[:, t].contiguous()
                zt, dzt = z3[:, t], dz3[:, t]
                dhp = torch.empty(B, D, device=dev, dtype=torch.float32)
                lib.genrl_gru_bwd(_st(), _p(dh), D, _p(zt), _ld(zt), _p(ctx.gamma), _p(ctx.beta), _p(hp), D, _p(st_t), _p(dzt), _ld(dzt), _p(dhp), D,
                                  _p(gb), _p(bb), B, D, _CODE[torch.bfloat16])
                # d(m_t h_{t-1}) = (1 - u) dh + dz_t W_h;  times the reset mask = the gradient reaching h_{t-1}
                gemm(dzt, wh, out=dhp, trans_w=True, accumulate=True)
                carry = dhp if m3 is None else dhp * m3[:, t]
            if rc == 2:
                dh0 = carry if want_dh0 else None
            dw = None
            if train:
                gbuf, dw_ret = _grad_buf(w)
                gemm(dz, hprev, out=gbuf[:, Hd:], trans_a=True, trans_w=True, accumulate=True)
                dw = dw_ret
        ctx.dz_compute = dz
        return cast_to(dz, torch.float32), dw, dg, db, None, dh0, None, None, None


def scan_stream_supported(B, D):
    """the weight-streaming scan: bf16 mode, at most 16 sequences per GPU, D a multiple of 128 with D/32 CTAs co-resident"""
    import os
    return (_State.dtype == torch.bfloat16 and os.environ.get("GENRL_SCAN_STREAM", "1") != "0"
            and bool(lib.load().genrl_rssm_scan_stream_supported(int(B), int(D))))


def gru_scan(gx, w_gru, gamma, beta, mask, h0, B, T, Hd):
    """gx fp32 [B*T, 3D] (x-part pre-activations incl. nothing else), returns deter fp32 [B*T, D] (row n = b*T + t)."""
    if not scan_supported(gx.shape[1] // 3):
        if scan_stream_supported(B, gx.shape[1] // 3):
            return _GruScanStream.apply(gx, w_gru, gamma, beta, mask, h0, B, T, Hd)
        return _gru_scan_stepwise(gx, w_gru, gamma, beta, mask, h0, B, T, Hd)
    return _GruScan.apply(gx, w_gru, gamma, beta, mask, h0, B, T, Hd)


# ---------------------------------------------------------------------------------------------------------------
# noise providers: in-kernel Philox (product) or an explicit tape of reference draws (parity tests)
# ---------------------------------------------------------------------------------------------------------------
class PhiloxNoise:
    """Fast mode: noise is generated inside the kernels (Philox4x32-10).  key = seed (+ a device-side step counter that
    `tick()` bumps once per update, so captured CUDA graphs draw fresh noise on every replay), counter = per-call offset."""

    def __init__(self, seed=0):
        self.seed = int(seed)
        self.offset = 0
        self._step = {}

    def step_counter(self, device):
        k = str(device)
        c = self._step.get(k)
        if c is None:
            c = torch.zeros(1, device=device, dtype=torch.int64)
            self._step[k] = c
        return c

    def tick(self, device):
        """new update call: bump the device-side key, restart the per-call offsets (identical in eager and graph mode)"""
        lib.genrl_counter_add(_st(), self.step_counter(device).data_ptr(), 1)
        self.offset = 0

    def exp(self, rows, K, device):
        self.offset += 1
        return None, self.seed, self.offset

    def randn(self, shape, device):
        self.offset += 1
        out = torch.empty(shape, device=device, dtype=torch.float32)
        lib.genrl_fill_normal(_st(), out.data_ptr(), out.numel(), self.seed, self.offset, self.step_counter(device).data_ptr())
        return out

    def randn_into(self, out):
        """standard-normal fill of a preallocated contiguous fp32 buffer"""
        self.offset += 1
        lib.genrl_fill_normal(_st(), out.data_ptr(), out.numel(), self.seed, self.offset, self.step_counter(out.device).data_ptr())
        return out

    def randint(self):
        return None

    def observe(self, B, T, S, K, device):
        return self.exp(B * T * S, K, device), self.exp(B * T * S, K, device)


class TapeNoise:
    """Parity mode: replays the draws recorded from the reference/oracle run, in the reference's order."""

    def __init__(self, items, device):
        self.items = list(items)
        self.pos = 0
        self.device = device

    def _next(self, kind):
        k, v = self.items[self.pos]
        if k != kind:
            raise GenrlError(f"noise tape out of sync at {self.pos}: wanted {kind}, tape has {k}")
        self.pos += 1
        return v

    def exp(self, rows, K, device):
        v = self._next("exp")
        return v.to(device=device, dtype=torch.float32).reshape(rows, K).contiguous(), 0, 0

    def randn(self, shape, device):
        return self._next("randn").to(device=device, dtype=torch.float32).reshape(shape).contiguous()

    def randn_into(self, out):
        out.copy_(self._next("randn").to(device=out.device, dtype=torch.float32).reshape(out.shape))
        return out

    def randint(self):
        return self._next("randint")

    def tick(self, device):
        pass

    def observe(self, B, T, S, K, device):
        """reference order per step: randint, prior draw, posterior draw (dreamer_utils.py:437-438, :477)"""
        pri, pos = [], []
        for _ in range(T):
            self._next("randint")
            pri.append(self._next("exp").reshape(B, S, K))
            pos.append(self._next("exp").reshape(B, S, K))
        to = lambda lst: torch.stack(lst, 1).to(device=device, dtype=torch.float32).reshape(B * T * S, K).contiguous()
        return (to(pos), 0, 0), (to(pri), 0, 0)


noise = PhiloxNoise(0)


def noise_step_ptr(device):
    """device pointer of the step counter that keys the in-kernel generators (None for tape replay)"""
    if isinstance(noise, PhiloxNoise):
        return noise.step_counter(device).data_ptr()
    return None


def noise_tick(device):
    if isinstance(noise, PhiloxNoise):
        noise.tick(device)


def set_noise(provider):
    global noise
    noise = provider
    return provider
