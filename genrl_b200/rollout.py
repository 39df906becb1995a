"""The imagination rollout as ONE autograd node with explicit buffers (reference WorldModel.imagine, agent/dreamer.py:254-287, over
EnsembleRSSM.img_step dreamer_utils.py:459-487, the actor MLP :718-747 / :814-819 and OneHotDist :177-197).

The op-level path builds the rollout out of ~45 autograd nodes per step: concatenations that materialise [stoch | action], [x | h] and
[stoch | deter], dtype casts, `torch.stack` of the per-step lists, and — in the backward — autograd's own gradient-accumulation adds.
Here the H steps run over pre-laid-out buffers:

  slab  [H+1, N, a0 + S*K + D + Hd]   compute dtype   one row per (step, sample): [ action_{i+1} (padded to a0) | stoch_i | deter_i | x_{i+1} ].
                                            The actor reads feat_i = [stoch | deter], the input block [action | stoch], the GRU GEMM [deter | x]
                                            — column windows of the same row (pointer + leading dimension), with W_in / W_gru column-permuted
                                            once per call; the sampler, the GRU gate kernel, the actor head and the input block's chain kernel
                                            write straight into their ranges, so no copy kernel runs between the steps
  stoch / deter / logit / action [H+1, N, .] fp32   the API's `seq` entries, written in place by the same kernels (no torch.stack)
  per-step saved tensors (pre-LayerNorm activations, statistics, GRU pre-activations, head outputs, noise) in [H, N, .] slabs

The backward is the hand-sequenced reverse loop over the recurrence only: gradient sums are folded into the GRU backward kernel (up to four
sources summed on the fly).  The actor reads sg(feat_{i-1}), so nothing in the recurrence depends on its backward: it runs ONCE over all H*N
rows after the loop (head, per layer one LayerNorm backward incl. parameter gradients, one dgrad and one wgrad GEMM at 36 000 rows).
Nothing goes through autograd inside the rollout.  ~11 launches per step forward instead of ~45, no ATen kernels.

Gradients flow out of `feat`, `stoch` and `policy_ent` (the only rollout outputs a GenRL loss reads: dreamer.py:392-438,
tools/genrl_utils.py:250-277) into the ACTOR's parameters; the world model is frozen inside the behaviour update (dreamer.py:376-384) and
the start state is detached (dreamer.py:78-79).  Anything else falls back to the op-level path in WorldModel.imagine.
"""
import torch
from torch.autograd import Function

from . import ops
from ._lib import lib
from .ops import _st, _p, _ld, _CODE, ACT_SILU, F32


def _pad8(n):
    return (n + 7) // 8 * 8


def supported(policy, rssm):
    """the explicit rollout covers the GenRL configuration: LayerNorm MLP actor with a 'normal' head, frozen single-member RSSM"""
    from .agent import dreamer_utils as common
    if not isinstance(policy, common.MLP) or not isinstance(rssm, common.EnsembleRSSM):
        return False
    if policy._norm != "layer" or policy._out._dist != "normal" or not isinstance(policy._act, torch.nn.SiLU):
        return False
    if any(p.requires_grad for p in rssm.parameters()):
        return False
    if any(p.requires_grad and ops.param_info(p) is None for p in policy.parameters()):
        return False            # parameter gradients are written straight into the ParamStore's flat buffers
    return True


class _Win:
    """column window [c0, c1) of a [steps, N, ld] buffer: per-step views and the rows of several steps as one strided matrix"""

    def __init__(self, t, c0, c1):
        self.t, self.c0, self.c1 = t, c0, c1

    def __getitem__(self, i):
        return self.t[i][:, self.c0:self.c1]

    def all(self, lo=0, hi=None):
        hi = self.t.shape[0] if hi is None else hi
        return self.t[lo:hi].reshape(-1, self.t.shape[2])[:, self.c0:self.c1]


class _Slab:
    """[steps, N, cols] buffer in a given dtype whose rows keep a TMA-legal leading dimension"""

    def __init__(self, steps, N, cols, dtype, device, zero=False):
        ld = _pad8(cols) if dtype == torch.bfloat16 else cols
        self.t = (torch.zeros if (zero or ld != cols) else torch.empty)(steps, N, ld, device=device, dtype=dtype)
        self.cols = cols

    def __getitem__(self, i):
        return self.t[i][:, :self.cols]

    def all(self, lo=0, hi=None):
        """rows of steps lo..hi-1 as one [steps*N, cols] matrix (contiguous slab)"""
        hi = self.t.shape[0] if hi is None else hi
        return self.t[lo:hi].reshape(-1, self.t.shape[2])[:, :self.cols]


def _ln_fwd(x, gamma, beta, eps, y, stats):
    lib.genrl_ln_act_fwd(_st(), _p(x), _ld(x), _p(gamma), _p(beta), float(eps), ACT_SILU, _p(y), _ld(y), _p(stats), x.shape[0], x.shape[1],
                         _CODE[x.dtype])


def _ln_bwd(dy, pre, gamma, beta, stats, dpre):
    lib.genrl_ln_act_bwd(_st(), _p(dy), _ld(dy), _p(pre), _ld(pre), _p(gamma), _p(beta), _p(stats), ACT_SILU, _p(dpre), _ld(dpre), None, None,
                         None, pre.shape[0], pre.shape[1], _CODE[pre.dtype])


def _vec_ok(t):
    """16-byte vector access: aligned base, leading dimension a multiple of 16 bytes"""
    return t.data_ptr() % 16 == 0 and (_ld(t) * t.element_size()) % 16 == 0


def _copy(src, dst, accumulate=False):
    lib.genrl_gather_rows(_st(), _p(src), _ld(src), _CODE[src.dtype], None, None, _p(dst), _ld(dst), _CODE[dst.dtype], dst.shape[0], src.shape[1],
                          int(accumulate))


class _Rollout(Function):
    @staticmethod
    def forward(ctx, stoch0, deter0, logit0, policy, rssm, horizon, eval_policy, *actor_params):
        ctx.set_materialize_grads(False)       # unused outputs (stoch with the env reward, ...) arrive as None, not as zero slabs
        cd = ops.compute_dtype()
        dev = deter0.device
        S, K, D, Hd = rssm._stoch, rssm._discrete, rssm._deter, rssm._hidden
        SK = S * K
        N = deter0.shape[0]
        H = horizon
        L, U, A = policy._layers, policy._units, int(policy._out._out.weight.shape[0])
        noise = ops.noise
        # ---- weights in compute layout -----------------------------------------------------------------------------------------------------
        W = ops.compute_weight
        dense = [getattr(policy, f"dense{l}") for l in range(L)]
        norms = [getattr(policy, f"norm{l}")._layer for l in range(L)]
        w_head = torch.cat([W(policy._out._out.weight), W(policy._out._std.weight)], 0).contiguous()            # [2A, U]
        b_head = torch.cat([policy._out._out.bias.detach(), policy._out._std.bias.detach()], 0).contiguous()
        lin_in, ln_in = rssm._img_in[0], rssm._img_in[1]._layer
        lin_out, ln_out = rssm._ensemble_img_out[0][0], rssm._ensemble_img_out[0][1]._layer
        lin_dist = rssm._ensemble_img_dist[0]
        w_gru, ln_gru = rssm._cell._layer.weight, rssm._cell._norm
        # ---- buffers -----------------------------------------------------------------------------------------------------------------------
        # ONE row per (step, sample) holds every operand the next step's GEMMs read, so no copy kernel runs between the steps:
        #   [ action_{i+1} (padded to a0) | stoch_i (SK) | deter_i (D) | x_{i+1} (Hd) ]
        #   img_in input  [action | stoch] = cols [0, a0+SK)        (W_in with its columns permuted to [action | 0 | stoch])
        #   actor input   feat_i           = cols [a0, a0+SK+D)
        #   GRU input     [deter | x]      = cols [a0+SK, a0+SK+D+Hd)   (W_gru with its columns permuted to [h | x])
        a0 = _pad8(A)
        o_s, o_d, o_x, o_e = a0, a0 + SK, a0 + SK + D, a0 + SK + D + Hd
        slab = torch.empty(H + 1, N, o_e, device=dev, dtype=cd)
        slab[:, :, :a0].zero_()                                   # padded action columns meet zero weights: they must be finite
        feat = _Win(slab, o_s, o_x)
        w_in_c, w_gru_c = W(lin_in.weight), W(w_gru)
        w_in_p = torch.cat([w_in_c[:, SK:SK + A], w_in_c.new_zeros(Hd, a0 - A), w_in_c[:, :SK]], 1)       # [Hd, a0 + SK]
        w_gru_p = torch.cat([w_gru_c[:, Hd:Hd + D], w_gru_c[:, :Hd]], 1)                                   # [3D, D + Hd]
        stoch = torch.empty(H + 1, N, SK, device=dev, dtype=torch.float32)
        deter = torch.empty(H + 1, N, D, device=dev, dtype=torch.float32)
        logit = torch.empty(H + 1, N, SK, device=dev, dtype=torch.float32)
        action = torch.zeros(H + 1, N, A, device=dev, dtype=torch.float32)
        ent = torch.empty(max(H, 1), N, device=dev, dtype=torch.float32)
        a_pre = [_Slab(H, N, U, cd, dev) for _ in range(L)]
        a_y = [_Slab(H, N, U, cd, dev) for _ in range(L)]
        a_st = [torch.empty(H, N, 2, device=dev, dtype=torch.float32) for _ in range(L)]
        out2 = torch.empty(H, N, 2 * A, device=dev, dtype=torch.float32)
        eps = torch.empty(H, N, A, device=dev, dtype=torch.float32)
        in_pre, in_st = _Slab(H, N, Hd, cd, dev), torch.empty(H, N, 2, device=dev, dtype=torch.float32)
        z, g_st = _Slab(H, N, 3 * D, cd, dev), torch.empty(H, N, 2, device=dev, dtype=torch.float32)
        out_pre, out_st = _Slab(H, N, Hd, cd, dev), torch.empty(H, N, 2, device=dev, dtype=torch.float32)
        y_out = _Slab(1, N, Hd, cd, dev)[0]
        # ---- step 0 = the start state ------------------------------------------------------------------------------------------------------
        stoch[0].copy_(stoch0.reshape(N, SK))
        deter[0].copy_(deter0)
        logit[0].copy_(logit0.reshape(N, SK))
        _copy(stoch[0], slab[0][:, o_s:o_d]); _copy(deter[0], slab[0][:, o_d:o_x])
        if not isinstance(noise, ops.PhiloxNoise):       # dreamer.py:259-260 samples (and discards) one action at the start state: keep the
            noise.randn_into(torch.empty(N, A, device=dev, dtype=torch.float32))      # replayed reference draws in step
        step_ptr = ops.noise_step_ptr(dev)
        big = N >= ops.CHAIN_MIN_ROWS
        aligned = o_e % 8 == 0 and SK % 8 == 0 and D % 8 == 0
        chain_a = big and ops.chain_supported(U, cd) and all(n.eps == norms[0].eps for n in norms) and aligned
        chain_h = big and ops.chain_supported(Hd, cd) and aligned
        for i in range(1, H + 1):
            j = i - 1
            rj, ri = slab[j], slab[i]
            # actor on sg(feat_{i-1})                                                                     dreamer.py:265-266, dreamer_utils.py:738-747
            x = rj[:, o_s:o_x]
            if chain_a:
                # the whole trunk in one persistent cluster launch (csrc/mlp_chain.cu)
                ops.mlp_chain_fwd(x, [W(d.weight) for d in dense], [d.bias.detach() if d.bias is not None else None for d in dense],
                                  [n.weight for n in norms], [n.bias for n in norms], norms[0].eps, ACT_SILU,
                                  [a_pre[l][j] for l in range(L)], [a_y[l][j] for l in range(L)], [a_st[l][j] for l in range(L)])
                x = a_y[L - 1][j]
            else:
                for l in range(L):
                    ops.gemm(x, W(dense[l].weight), out=a_pre[l][j], bias=dense[l].bias.detach() if dense[l].bias is not None else None)
                    _ln_fwd(a_pre[l][j], norms[l].weight, norms[l].bias, norms[l].eps, a_y[l][j], a_st[l][j])
                    x = a_y[l][j]
            ops.gemm(x, w_head, out=out2[j], bias=b_head)
            e = None
            if not eval_policy:
                e = noise.randn_into(eps[j])
            lib.genrl_actor_head_fwd(_st(), _p(out2[j]), 2 * A, _p(e), int(bool(eval_policy)), _p(action[i]), A, _p(rj[:, :A]), _ld(rj), _p(ent[j]),
                                     N, A, _CODE[cd])
            # img_step                                                                                     dreamer_utils.py:459-487
            xin, xdst = rj[:, :o_d], rj[:, o_x:o_e]
            if chain_h:
                ops.mlp_chain_fwd(xin, [w_in_p], [lin_in.bias.detach()], [ln_in.weight], [ln_in.bias], ln_in.eps, ACT_SILU,
                                  [in_pre[j]], [xdst], [in_st[j]])
            else:
                ops.gemm(xin, w_in_p, out=in_pre[j], bias=lin_in.bias.detach())
                _ln_fwd(in_pre[j], ln_in.weight, ln_in.bias, ln_in.eps, xdst, in_st[j])
            ops.gemm(rj[:, o_d:o_e], w_gru_p, out=z[j])
            lib.genrl_gru_fwd(_st(), _p(z[j]), _ld(z[j]), _p(ln_gru.weight), _p(ln_gru.bias), _p(deter[j]), D, _p(deter[i]), D,
                              _p(ri[:, o_d:o_x]), _ld(ri), _p(g_st[j]), N, D, _CODE[cd])
            if chain_h:
                ops.mlp_chain_fwd(ri[:, o_d:o_x], [W(lin_out.weight)], [lin_out.bias.detach()], [ln_out.weight], [ln_out.bias], ln_out.eps, ACT_SILU,
                                  [out_pre[j]], [y_out], [out_st[j]])
            else:
                ops.gemm(ri[:, o_d:o_x], W(lin_out.weight), out=out_pre[j], bias=lin_out.bias.detach())
                _ln_fwd(out_pre[j], ln_out.weight, ln_out.bias, ln_out.eps, y_out, out_st[j])
            ops.gemm(y_out, W(lin_dist.weight), out=logit[i], bias=lin_dist.bias.detach())
            noise.randint()
            nz, seed, off = noise.exp(N * S, K, dev)
            lib.genrl_cat_sample_fwd(_st(), _p(logit[i]), _p(nz), seed, off, step_ptr if nz is None else None, 1, _p(ri[:, o_s:o_d]), _ld(ri),
                                     _p(stoch[i]), SK, None, N, S, K, _CODE[cd])
        feat_out = slab[:, :, o_s:o_x]
        ctx.mods = (policy, rssm, dense, norms, w_head, w_in_p, w_gru_p, a0)
        ctx.dims = (S, K, D, Hd, N, H, L, U, A, cd, bool(eval_policy))
        ctx.bufs = dict(feat=feat, deter=deter, logit=logit, a_pre=a_pre, a_y=a_y, a_st=a_st, out2=out2, eps=eps, in_pre=in_pre, in_st=in_st, z=z,
                        g_st=g_st, out_pre=out_pre, out_st=out_st)
        ctx.n_params = len(actor_params)
        stoch_v, logit_v = stoch.reshape(H + 1, N, S, K), logit.reshape(H + 1, N, S, K)
        ctx.mark_non_differentiable(deter, logit_v, action)
        return feat_out, stoch_v, deter, logit_v, action, (ent if H > 0 else ent[:0])

    @staticmethod
    def backward(ctx, dfeat, dstoch, _ddeter, _dlogit, _daction, dent):
        policy, rssm, dense, norms, w_head, w_in_p, w_gru_p, a0 = ctx.mods
        S, K, D, Hd, N, H, L, U, A, cd, eval_policy = ctx.dims
        b = ctx.bufs
        SK = S * K
        dev = b["deter"].device
        W = ops.compute_weight
        with ops.use_dtype(cd):
            lin_in, ln_in = rssm._img_in[0], rssm._img_in[1]._layer
            lin_out, ln_out = rssm._ensemble_img_out[0][0], rssm._ensemble_img_out[0][1]._layer
            lin_dist = rssm._ensemble_img_dist[0]
            ln_gru = rssm._cell._norm
            if dfeat is not None:
                dfeat = dfeat if (dfeat.dtype == cd and dfeat.stride(-1) == 1) else dfeat.to(cd).contiguous()
            if dstoch is not None:
                dstoch = dstoch.reshape(H + 1, N, SK)
                dstoch = dstoch if dstoch.stride(-1) == 1 else dstoch.contiguous()
            if dent is not None:
                dent = dent.contiguous().float()
            dout2c = _Slab(H, N, 2 * A, cd, dev, zero=True)
            G = [torch.zeros(N, D, device=dev, dtype=torch.float32), torch.empty(N, D, device=dev, dtype=torch.float32)]     # d loss / d deter_i
            dlogit = torch.empty(N, SK, device=dev, dtype=torch.float32)
            dlogit_c = dlogit if cd == torch.float32 else _Slab(1, N, SK, cd, dev)[0]
            dy_out = _Slab(1, N, Hd, cd, dev)[0]
            dpre_out = _Slab(1, N, Hd, cd, dev)[0]
            dz = _Slab(1, N, 3 * D, cd, dev)[0]
            dx = _Slab(1, N, Hd, cd, dev)[0]
            dxh = [_Slab(1, N, D + Hd, cd, dev)[0], _Slab(1, N, D + Hd, cd, dev)[0]]      # [d deter_{i-1} | d x_i] (the permuted weight's order)
            fold_ok = D == 1024 and cd == torch.bfloat16
            pending_h = None
            dgo = _Slab(1, N, D, cd, dev)[0]                               # d deter_i through the prior's output block
            merged_gru_dgrad = cd == torch.bfloat16 and N >= 1024
            dpre_in = _Slab(1, N, Hd, cd, dev)[0]
            # d loss / d [stoch_{i-1} | action_i] of EVERY step: the stoch part is the next iteration's carry, the action part is all the actor's
            # backward needs, so the actor's whole backward runs ONCE over the H*N rows after the loop (below)
            dxin = _Slab(H, N, a0 + SK, cd, dev)                       # [d action_i (a0) | d stoch_{i-1}]
            train = H > 0 and dense[0].weight.requires_grad

            def actor_backward():
                """The actor reads sg(feat_{i-1}) (dreamer.py:265): no gradient leaves its trunk, so nothing in the recurrence's reverse loop
                depends on it.  Its backward of all H steps is therefore ONE pass over the contiguous [H*N, .] slabs: head backward, then per
                layer one LayerNorm backward (row pass + dgamma / dbeta / dbias column partials in the same launch), one dgrad GEMM and one
                wgrad GEMM at 36 000 rows — instead of H x (4 + 4 + 2) launches at 2 400 rows."""
                st = _st()
                M = H * N
                e = None if eval_policy else b["eps"].reshape(M, A)
                if e is None:
                    e = torch.zeros(M, A, device=dev, dtype=torch.float32)
                DA = dxin.all()[:, :A]
                DO = dout2c.all()
                lib.genrl_actor_head_bwd(st, _p(b["out2"]), 2 * A, _p(e), _p(DA), _ld(DA), _p(dent), 1.0 if dent is not None else 0.0,
                                         _p(DO), _ld(DO), _CODE[cd], M, A, _CODE[cd])
                DY = _Slab(1, M, U, cd, dev)[0]
                DP = _Slab(1, M, U, cd, dev)[0]
                ops.gemm(DO, w_head, out=DY, trans_w=True)
                XH = b["a_y"][L - 1].all()
                for part, lin in ((0, policy._out._out), (1, policy._out._std)):
                    gbuf, _ = ops._grad_buf(lin.weight)
                    dsl = DO[:, part * A:(part + 1) * A]
                    ops.gemm(dsl, XH, out=gbuf.view(A, U), trans_a=True, trans_w=True, accumulate=True)
                    bbuf, _ = ops._grad_buf(lin.bias)
                    lib.genrl_colsum(st, _p(dsl), _ld(dsl), _p(bbuf), dsl.shape[0], A, _CODE[cd])
                for l in range(L - 1, -1, -1):
                    PRE = b["a_pre"][l].all()
                    gb, _ = ops._grad_buf(norms[l].weight)
                    bb, _ = ops._grad_buf(norms[l].bias)
                    lb = ops._grad_buf(dense[l].bias)[0] if dense[l].bias is not None else None
                    lib.genrl_ln_act_bwd(st, _p(DY), _ld(DY), _p(PRE), _ld(PRE), _p(norms[l].weight), _p(norms[l].bias),
                                         _p(b["a_st"][l].reshape(-1, 2)), ACT_SILU, _p(DP), _ld(DP), _p(gb), _p(bb), _p(lb), M, U, _CODE[cd])
                    X = b["feat"].all(0, H) if l == 0 else b["a_y"][l - 1].all()
                    gbuf, _ = ops._grad_buf(dense[l].weight)
                    ops.gemm(DP, X, out=gbuf.view(U, -1), trans_a=True, trans_w=True, accumulate=True)
                    if l > 0:
                        ops.gemm(DP, W(dense[l].weight), out=DY, trans_w=True)

            have_carry = False
            cur = 0
            for i in range(H, 0, -1):
                j = i - 1
                # ---- d/d stoch_i : feat slice + explicit stoch output + the next step's input projection ------------------------------------
                srcs = []
                if dfeat is not None:
                    srcs.append(dfeat[i][:, :SK])
                if dstoch is not None:
                    srcs.append(dstoch[i])
                if have_carry:
                    srcs.append(dxin[i][:, a0:])
                gh = G[cur]
                # further contributions to d loss / d deter_i, summed inside the GRU backward kernel when it can (else by accumulate-copies)
                extras = []
                if dfeat is not None:
                    extras.append(dfeat[i][:, SK:])
                if pending_h is not None:
                    extras.append(pending_h)
                pending_h = None
                fold = fold_ok and all(e_.dtype == cd and _vec_ok(e_) for e_ in extras)
                if not fold:
                    for e_ in extras:
                        _copy(e_, gh, accumulate=True)
                    extras = []
                if srcs:
                    if K == 32 and len(srcs) <= 2 and all(s_.dtype == cd and _vec_ok(s_) for s_ in srcs) and _vec_ok(dlogit_c):
                        # both contributions summed on the fly, gradient written in the activation dtype: one launch
                        s2 = srcs[1] if len(srcs) == 2 else None
                        lib.genrl_cat_sample_bwd2(_st(), _p(b["logit"][i]), _p(srcs[0]), _ld(srcs[0]), _p(s2), _ld(s2) if s2 is not None else 0,
                                                  _p(dlogit_c), _ld(dlogit_c), N, S, K, _CODE[cd])
                    else:
                        for n_, s_ in enumerate(srcs):          # the straight-through backward is linear in d stoch: accumulate the contributions
                            lib.genrl_cat_sample_bwd(_st(), _p(b["logit"][i]), _p(s_), _ld(s_), 1, _p(dlogit), int(n_ > 0), N, S, K, _CODE[s_.dtype])
                        if dlogit_c is not dlogit:
                            _copy(dlogit, dlogit_c)
                    ops.gemm(dlogit_c, W(lin_dist.weight), out=dy_out, trans_w=True)
                    _ln_bwd(dy_out, b["out_pre"][j], ln_out.weight, ln_out.bias, b["out_st"][j], dpre_out)
                    if fold:
                        # activation-dtype output (the plain epilogue), summed inside the GRU backward kernel below
                        ops.gemm(dpre_out, W(lin_out.weight), out=dgo, trans_w=True)
                        extras.append(dgo)
                    else:
                        ops.gemm(dpre_out, W(lin_out.weight), out=gh, trans_w=True, accumulate=True)
                # ---- GRU cell -------------------------------------------------------------------------------------------------------------
                gn = G[1 - cur]
                if extras:
                    ex = extras + [None] * (3 - len(extras))
                    lib.genrl_gru_bwd_sum(_st(), _p(gh), D, _p(ex[0]), _ld(ex[0]), _p(ex[1]), _ld(ex[1]) if ex[1] is not None else 0,
                                          _p(ex[2]), _ld(ex[2]) if ex[2] is not None else 0, _p(b["z"][j]), _ld(b["z"][j]),
                                          _p(ln_gru.weight), _p(ln_gru.bias), _p(b["deter"][j]), D, _p(b["g_st"][j]), _p(dz), _ld(dz), _p(gn), D,
                                          N, D, _CODE[cd])
                else:
                    lib.genrl_gru_bwd(_st(), _p(gh), D, _p(b["z"][j]), _ld(b["z"][j]), _p(ln_gru.weight), _p(ln_gru.bias), _p(b["deter"][j]), D,
                                      _p(b["g_st"][j]), _p(dz), _ld(dz), _p(gn), D, None, None, N, D, _CODE[cd])
                if i > 1 and merged_gru_dgrad:
                    # one dgrad GEMM for [x | h] (N = Hd + D: enough 256x256 tiles for the 2-CTA kernel); the h half joins the carry inside the
                    # next iteration's GRU backward (or by an accumulate-copy when that kernel is not eligible)
                    ops.gemm(dz, w_gru_p, out=dxh[i & 1], trans_w=True)
                    pending_h = dxh[i & 1][:, :D]
                    dx_i = dxh[i & 1][:, D:]
                else:
                    ops.gemm(dz, w_gru_p[:, D:], out=dx, trans_w=True)
                    if i > 1:
                        ops.gemm(dz, w_gru_p[:, :D], out=gn, trans_w=True, accumulate=True)
                    dx_i = dx
                _ln_bwd(dx_i, b["in_pre"][j], ln_in.weight, ln_in.bias, b["in_st"][j], dpre_in)
                ops.gemm(dpre_in, w_in_p, out=dxin[j], trans_w=True)
                have_carry = True
                cur = 1 - cur
            if train:
                actor_backward()
        return (None,) * (7 + ctx.n_params)


def imagine_rows(policy, rssm, start, horizon, eval_policy=False):
    """start: dict of flattened rows {stoch [N,S,K], deter [N,D], logit [N,S,K]} -> seq dict with [H+1, N, ...] entries"""
    params = list(policy.parameters())
    feat, stoch, deter, logit, action, ent = _Rollout.apply(start["stoch"].detach(), start["deter"].detach(), start["logit"].detach(), policy, rssm,
                                                           int(horizon), bool(eval_policy), *params)
    seq = {"stoch": stoch, "deter": deter, "logit": logit, "feat": feat, "action": action}
    if not eval_policy and horizon > 0:
        seq["policy_ent"] = ent
    return seq
