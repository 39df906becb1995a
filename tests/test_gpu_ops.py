"""GPU parity of every C-ABI kernel (forward and hand-written backward) against the CPU oracle / plain torch fp32.

Run on the B200 box:  python -m pytest tests -m gpu -x -q
Tolerances: fp32 kernels 1e-4 relative (north_star), bf16 tensor-core GEMM compared against an fp32 matmul of the same
bf16-rounded operands (so only accumulation order differs), integer outputs (sample indices, buckets) bit-exact.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

cuda = torch.cuda.is_available()
if cuda:
    from genrl_b200 import ops
    from genrl_b200._lib import lib
from oracle import genrl_oracle as O

DEV = "cuda:0"


def close(a, b, rtol=1e-4, atol=1e-5, msg=""):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    scale = float(b.abs().max()) + 1e-12
    err = float((a - b).abs().max())
    assert a.shape == b.shape, (msg, a.shape, b.shape)
    assert err <= atol + rtol * scale, f"{msg}: max abs err {err:.3e} vs scale {scale:.3e}"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=g) * scale


# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ta,tw", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(70, 50, 33), (257, 130, 96), (512, 384, 256)])
def test_gemm_fp32(ta, tw, M, N, K):
    a, w, b = rnd(M, K), rnd(N, K, seed=1), rnd(N, seed=2)
    ref = a @ w.t() + b
    ad = (a.t().contiguous() if ta else a).to(DEV)
    wd = (w.t().contiguous() if tw else w).to(DEV)
    out = ops.gemm(ad, wd, bias=b.to(DEV), trans_a=ta, trans_w=tw)
    close(out, ref, msg="gemm fp32")
    out2 = out.clone()
    ops.gemm(ad, wd, out=out2, trans_a=ta, trans_w=tw, accumulate=True)
    close(out2, ref + a @ w.t(), msg="gemm fp32 accumulate")


@pytest.mark.parametrize("ta,tw", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 136), (2400, 1024, 1024), (1000, 3072, 2048), (64, 255, 1024), (1024, 1034, 2400)])
def test_gemm_bf16_tcgen05(ta, tw, M, N, K):
    a = rnd(M, K).to(torch.bfloat16)
    w = rnd(N, K, seed=1).to(torch.bfloat16)
    b = rnd(N, seed=2)
    ref = a.float() @ w.float().t() + b
    ad = ops.cast_to((a.t().contiguous() if ta else a).to(DEV), torch.bfloat16)
    wd = ops.cast_to((w.t().contiguous() if tw else w).to(DEV), torch.bfloat16)
    out = ops.gemm(ad, wd, bias=b.to(DEV), trans_a=ta, trans_w=tw, out_dtype=torch.float32)
    close(out, ref, rtol=2e-5, atol=1e-4, msg=f"tcgen05 gemm fp32-out ta={ta} tw={tw}")
    outb = ops.gemm(ad, wd, bias=b.to(DEV), trans_a=ta, trans_w=tw, out_dtype=torch.bfloat16)
    close(outb, ref, rtol=1e-2, atol=1e-2, msg="tcgen05 gemm bf16-out")
    acc = torch.full((M, N), 0.5, device=DEV)
    ops.gemm(ad, wd, out=acc, trans_a=ta, trans_w=tw, accumulate=True)          # split-K + atomics path
    close(acc, a.float() @ w.float().t() + 0.5, rtol=2e-5, atol=1e-4, msg="tcgen05 gemm accumulate")


@pytest.mark.parametrize("tw", [False, True])
@pytest.mark.parametrize("M,N,K", [(700, 255, 136), (1000, 384, 1024), (38400, 1024, 1024), (36000, 255, 1024)])
def test_gemm_bf16_persistent_256x128(tw, M, N, K, monkeypatch):
    """persistent 256x128 kernel (two UMMA per k-step sharing the B tile, double-buffered TMEM): forced on small shapes too"""
    monkeypatch.setenv("GENRL_GEMM_PERSIST", "2")
    a = rnd(M, K).to(torch.bfloat16)
    w = rnd(N, K, seed=1).to(torch.bfloat16)
    b = rnd(N, seed=2)
    ad = ops.cast_to(a.to(DEV), torch.bfloat16)
    wd = ops.cast_to((w.t().contiguous() if tw else w).to(DEV), torch.bfloat16)
    ref = ad.float() @ w.to(DEV).float().t() + b.to(DEV)
    for out_dtype in (torch.bfloat16, torch.float32):
        out = ops.gemm(ad, wd, bias=b.to(DEV), trans_w=tw, out_dtype=out_dtype)
        close(out, ref, rtol=2e-2 if out_dtype == torch.bfloat16 else 2e-3, atol=1e-2, msg=f"persistent gemm {out_dtype}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,D,eps,act", [(37, 64, 1e-5, 1), (1000, 1024, 1e-5, 1), (333, 48, 1e-3, 1), (50, 3072, 1e-5, 0)])
def test_ln_act(dtype, M, D, eps, act):
    with ops.use_dtype(dtype):
        x = rnd(M, D).to(dtype)
        g, b = 1 + 0.1 * rnd(D, seed=1), 0.1 * rnd(D, seed=2)
        dy = rnd(M, D, seed=3).to(dtype)
        xr = x.float().clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
        y = F.layer_norm(xr, (D,), gr, br, eps)
        y = F.silu(y) if act else y
        y.backward(dy.float())
        xd = x.to(DEV).requires_grad_(True)
        gd = torch.nn.Parameter(g.to(DEV)); bd = torch.nn.Parameter(b.to(DEV))
        yd = ops.ln_act(xd, gd, bd, eps, act)
        yd.backward(dy.to(DEV))
        tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
        close(yd, y, msg="ln y", **tol)
        close(xd.grad, xr.grad, msg="ln dx", **tol)
        close(gd.grad, gr.grad, msg="ln dgamma", **(tol if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)))
        close(bd.grad, br.grad, msg="ln dbeta", **(tol if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("frozen", [False, True])
@pytest.mark.parametrize("M,D,eps,act,pad", [(20011, 1024, 1e-5, 1, 0), (10000, 1024, 1e-5, 0, 8), (250003, 48, 1e-3, 1, 0), (120001, 96, 1e-3, 1, 0),
                                             (40000, 192, 1e-3, 1, 0), (30000, 384, 1e-3, 1, 16), (40000, 256, 1e-5, 1, 0)])
def test_ln_act_large_rows_bulk_copy_pipeline(dtype, frozen, M, D, eps, act, pad):
    """row counts above 4 tiles per SM run the cp.async.bulk pipelines of csrc/ln_pipe.cuh (forward; backward with and without the parameter
    gradients); ragged last tiles, strided rows (pad > 0: per-row bulk copies) and every channel width of the conv stacks"""
    with ops.use_dtype(dtype):
        x = rnd(M, D).to(dtype)
        g, b = 1 + 0.1 * rnd(D, seed=1), 0.1 * rnd(D, seed=2)
        dy = rnd(M, D, seed=3).to(dtype)
        xr = x.float().to(DEV).requires_grad_(True); gr = g.to(DEV).requires_grad_(True); br = b.to(DEV).requires_grad_(True)
        y = F.layer_norm(xr, (D,), gr, br, eps)
        y = F.silu(y) if act else y
        y.backward(dy.float().to(DEV))
        xbuf = torch.zeros(M, D + pad, device=DEV, dtype=dtype)
        xbuf[:, :D] = x.to(DEV)
        xd = xbuf[:, :D].detach().requires_grad_(True)
        gd, bd = g.to(DEV), b.to(DEV)
        if not frozen:
            gd, bd = torch.nn.Parameter(gd), torch.nn.Parameter(bd)
        yd = ops.ln_act(xd, gd, bd, eps, act)
        dyb = torch.zeros(M, D + pad, device=DEV, dtype=dtype)
        dyb[:, :D] = dy.to(DEV)
        yd.backward(dyb[:, :D])
        tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
        close(yd, y, msg="ln y", **tol)
        close(xd.grad, xr.grad, msg="ln dx", **tol)
        if not frozen:
            # column sums over up to 250 000 rows: compare relative to the column's scale
            scale = float(gr.grad.abs().max())
            ptol = dict(rtol=1e-3, atol=1e-4 * scale) if dtype == torch.float32 else dict(rtol=3e-2, atol=2e-2 * scale)
            close(gd.grad, gr.grad, msg="ln dgamma", **ptol)
            close(bd.grad, br.grad, msg="ln dbeta", **ptol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M", [2400, 2960, 2961, 40000])
def test_ln_act_bwd_frozen_row_variants(dtype, M):
    """frozen-parameter row pass at D = 1024: the 4-warp / five-blocks-per-SM variant (2 368 < M <= 2 960 on 148 SMs: one wave for the
    rollout's 2 400 rows) and the 8-warp variant give dx of the one-pass kernel with parameter gradients"""
    D = 1024
    with ops.use_dtype(dtype):
        x = rnd(M, D).to(dtype)
        g, b = 1 + 0.1 * rnd(D, seed=1), 0.1 * rnd(D, seed=2)
        dy = rnd(M, D, seed=3).to(dtype)
        xr = x.float().clone().requires_grad_(True)
        F.silu(F.layer_norm(xr, (D,), g, b, 1e-5)).backward(dy.float())
        xd = x.to(DEV).requires_grad_(True)
        yd = ops.ln_act(xd, g.to(DEV), b.to(DEV), 1e-5, 1)          # gamma / beta without gradients -> ln_bwd_fast_kernel
        yd.backward(dy.to(DEV))
        tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
        close(xd.grad, xr.grad, msg="ln dx (frozen)", **tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,K,N", [(130, 72, 64), (200, 264, 128), (333, 1034, 512), (2400, 2048, 1024)])
def test_linear_ln_act_and_linear(dtype, M, K, N, monkeypatch):
    """N = 128..1024 in bf16 runs the cluster-fused GEMM + LayerNorm kernel (row statistics merged over DSMEM)"""
    monkeypatch.setattr(ops, "FUSED_LN_MAX_N", 1024)          # also exercise the 8-CTA cluster (off by default: not faster at N = 1024)
    with ops.use_dtype(dtype):
        lin = torch.nn.Linear(K, N); ln = torch.nn.LayerNorm(N)
        x = rnd(M, K).to(dtype)
        dy = rnd(M, N, seed=5).to(dtype)
        xr = x.float().clone().requires_grad_(True)
        wq = lin.weight.detach().to(dtype).float() if dtype != torch.float32 else lin.weight.detach()
        w_ref = wq.clone().requires_grad_(True)
        b_ref = lin.bias.detach().clone().requires_grad_(True)
        g_ref = ln.weight.detach().clone().requires_grad_(True); be_ref = ln.bias.detach().clone().requires_grad_(True)
        y = F.silu(F.layer_norm(F.linear(xr, w_ref, b_ref), (N,), g_ref, be_ref, 1e-5))
        y.backward(dy.float())
        lin_d, ln_d = lin.to(DEV), ln.to(DEV)
        xd = x.to(DEV).requires_grad_(True)
        yd = ops.linear_ln_act(xd, lin_d.weight, lin_d.bias, ln_d.weight, ln_d.bias, 1e-5, 1)
        yd.backward(dy.to(DEV))
        tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
        close(yd, y, msg="lla y", **tol)
        close(xd.grad, xr.grad, msg="lla dx", **tol)
        close(lin_d.weight.grad, w_ref.grad, msg="lla dW", **tol)
        close(lin_d.bias.grad, b_ref.grad, msg="lla db", **tol)
        close(ln_d.weight.grad, g_ref.grad, msg="lla dgamma", **tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gru_gates(dtype):
    M, D = 77, 64
    with ops.use_dtype(dtype):
        z = rnd(M, 3 * D).to(dtype)
        h = rnd(M, D, seed=1)
        g, b = 1 + 0.1 * rnd(3 * D, seed=2), 0.1 * rnd(3 * D, seed=3)
        dh = rnd(M, D, seed=4)
        zr = z.float().clone().requires_grad_(True); hr = h.clone().requires_grad_(True)
        gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
        parts = F.layer_norm(zr, (3 * D,), gr, br, 1e-5)
        r, c, u = torch.chunk(parts, 3, -1)
        r = torch.sigmoid(r); c = torch.tanh(r * c); u = torch.sigmoid(u - 1)
        out = u * c + (1 - u) * hr
        out.backward(dh)
        zd = z.to(DEV).requires_grad_(True); hd = h.to(DEV).requires_grad_(True)
        gd = torch.nn.Parameter(g.to(DEV)); bd = torch.nn.Parameter(b.to(DEV))
        od = ops.gru_gates(zd, gd, bd, hd)
        od.backward(dh.to(DEV))
        tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
        close(od, out, msg="gru h", **tol)
        close(zd.grad, zr.grad, msg="gru dz", **tol)
        close(hd.grad, hr.grad, msg="gru dh", **tol)
        close(gd.grad, gr.grad, msg="gru dgamma", **tol)
        close(bd.grad, br.grad, msg="gru dbeta", **tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gru_gates_frozen_d1024(dtype):
    """imagination path: frozen GRU LayerNorm parameters, D = 1024 (block-per-row backward kernel)"""
    M, D = 301, 1024
    with ops.use_dtype(dtype):
        z = rnd(M, 3 * D).to(dtype)
        h = rnd(M, D, seed=1)
        g, b = 1 + 0.1 * rnd(3 * D, seed=2), 0.1 * rnd(3 * D, seed=3)
        dh = rnd(M, D, seed=4)
        zr = z.float().clone().requires_grad_(True); hr = h.clone().requires_grad_(True)
        parts = F.layer_norm(zr, (3 * D,), g, b, 1e-5)
        r, c, u = torch.chunk(parts, 3, -1)
        r = torch.sigmoid(r); c = torch.tanh(r * c); u = torch.sigmoid(u - 1)
        out = u * c + (1 - u) * hr
        out.backward(dh)
        zd = z.to(DEV).requires_grad_(True); hd = h.to(DEV).requires_grad_(True)
        gd = torch.nn.Parameter(g.to(DEV), requires_grad=False); bd = torch.nn.Parameter(b.to(DEV), requires_grad=False)
        od = ops.gru_gates(zd, gd, bd, hd)
        od.backward(dh.to(DEV))
        tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
        close(od, out, msg="gru h", **tol)
        close(zd.grad, zr.grad, msg="gru dz", **tol)
        close(hd.grad, hr.grad, msg="gru dh", **tol)


@pytest.mark.parametrize("S,K", [(8, 8), (32, 32), (4, 48)])
def test_cat_sample(S, K):
    rows = 50
    logit = rnd(rows, S, K, scale=2.0)
    tape = O.NoiseTape()
    torch.manual_seed(3)
    lr = logit.clone().requires_grad_(True)
    st, idx = O.sample_onehot(lr, tape, True)
    dst = rnd(rows, S, K, seed=9)
    st.backward(dst)
    q = tape.items[0][1]
    ld = logit.reshape(rows, S * K).to(DEV).requires_grad_(True)
    sd, idd = ops.cat_sample(ld, S, K, q.to(DEV).contiguous(), 0, 0, True)
    sd.backward(dst.reshape(rows, S * K).to(DEV))
    assert torch.equal(idd.cpu().long(), idx), "sample indices must be bit-exact"
    assert torch.equal(sd.cpu().reshape(rows, S, K), st.detach())
    close(ld.grad.reshape(rows, S, K), lr.grad, msg="ST grad")
    # mode
    lr2 = logit.clone().requires_grad_(True)
    sm, idxm = O.sample_onehot(lr2, tape, False)
    sm.backward(dst)
    ld2 = logit.reshape(rows, S * K).to(DEV).requires_grad_(True)
    smd, idmd = ops.cat_sample(ld2, S, K, None, 0, 0, False)
    smd.backward(dst.reshape(rows, S * K).to(DEV))
    assert torch.equal(idmd.cpu().long(), idxm)
    close(ld2.grad.reshape(rows, S, K), lr2.grad, msg="mode ST grad")
    # philox path: valid one-hot, roughly follows the distribution
    sp, idp = ops.cat_sample(ld.detach(), S, K, None, 7, 1, True)
    assert float(sp.sum()) == rows * S and int(idp.min()) >= 0 and int(idp.max()) < K
    sp2, idp2 = ops.cat_sample(ld.detach(), S, K, None, 7, 2, True)
    assert not torch.equal(idp, idp2)


@pytest.mark.parametrize("forward,balance,free", [(False, 0.85, 1.0), (True, 0.8, 0.0), (False, 0.5, 1.0)])
def test_kl_balance(forward, balance, free):
    rows, S, K = 40, 8, 8
    post = rnd(rows, S, K, scale=1.5); prior = rnd(rows, S, K, seed=1, scale=1.5)
    pr = post.clone().requires_grad_(True); qr = prior.clone().requires_grad_(True)
    loss, value = O.kl_loss(pr, qr, forward, balance, free)
    (loss * 1.7).backward()
    pd = post.reshape(rows, S * K).to(DEV).requires_grad_(True); qd = prior.reshape(rows, S * K).to(DEV).requires_grad_(True)
    ld, vd, stats = ops.kl_balance(pd, qd, S, K, forward, balance, free)
    (ld * 1.7).sum().backward()
    close(ld.reshape(()), loss, msg="kl loss")
    close(vd, value, msg="kl value")
    close(pd.grad.reshape(rows, S, K), pr.grad, msg="kl dpost")
    close(qd.grad.reshape(rows, S, K), qr.grad, msg="kl dprior")
    ent_post, ent_prior = O.cat_entropy(post).mean(), O.cat_entropy(prior).mean()
    lhs_ent, rhs_ent = (ent_prior, ent_post) if forward else (ent_post, ent_prior)
    close(stats[2], lhs_ent, msg="ent lhs"); close(stats[3], rhs_ent, msg="ent rhs")


def test_twohot():
    rows = 300
    logits = rnd(rows, 255, scale=2.0)
    x = torch.cat([rnd(rows - 6, scale=30.0), torch.tensor([0.0, 1e9, -1e9, 20.0, -20.0, 0.157])])
    lr = logits.clone().requires_grad_(True)
    lp = O.twohot_logprob(lr, x[:, None])
    g = rnd(rows, seed=4)
    lp.backward(g)
    ld = logits.to(DEV).requires_grad_(True)
    lpd = ops.twohot_logprob(ld, x.to(DEV))
    lpd.backward(g.to(DEV))
    close(lpd, lp, msg="twohot logprob")
    close(ld.grad, lr.grad, msg="twohot dlogits")
    lr2 = logits.clone().requires_grad_(True)
    m = O.twohot_mean(lr2).squeeze(-1)
    m.backward(g)
    ld2 = logits.to(DEV).requires_grad_(True)
    md = ops.twohot_mean(ld2)
    md.backward(g.to(DEV))
    close(md, m, msg="twohot mean")
    close(ld2.grad, lr2.grad, msg="twohot mean grad")


def test_actor_head():
    rows, A = 123, 6
    out = rnd(rows, 2 * A)
    eps = rnd(rows, A, seed=2)
    r = out.clone().requires_grad_(True)
    mean = torch.tanh(r[:, :A]); std = 0.9 * torch.sigmoid(r[:, A:] + 2.0) + 0.1
    act = mean + std * eps
    ent = O.normal_entropy(std)
    da, de = rnd(rows, A, seed=3), rnd(rows, seed=4)
    (act * da).sum().backward(retain_graph=True)
    (ent * de).sum().backward()
    d = out.to(DEV).requires_grad_(True)
    ad, ed = ops.actor_head(d, eps.to(DEV), A)
    ((ad * da.to(DEV)).sum() + (ed * de.to(DEV)).sum()).backward()
    close(ad, act, msg="action"); close(ed, ent, msg="entropy"); close(d.grad, r.grad, msg="actor head grad")


def test_lambda_return():
    H, N = 15, 333
    reward = rnd(H + 1, N, 1); value = rnd(H + 1, N, 1, seed=1)
    rr = reward.clone().requires_grad_(True); vr = value.clone().requires_grad_(True)
    tgt = O.lambda_return(rr[:-1], vr[:-1], torch.full((H, N, 1), 0.99), vr[-1], 0.95)
    g = rnd(H, N, 1, seed=2)
    tgt.backward(g)
    rd = reward.to(DEV).requires_grad_(True); vd = value.to(DEV).requires_grad_(True)
    td = ops.lambda_return(rd, vd, 0.99, 0.95)
    td.backward(g.to(DEV))
    close(td, tgt, msg="lambda return")
    close(rd.grad, rr.grad, msg="dreward"); close(vd.grad, vr.grad, msg="dvalue")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,H,Cin,Cout,k", [(3, 64, 3, 8, 4), (2, 31, 8, 16, 4), (2, 14, 16, 24, 4), (5, 6, 24, 32, 4), (3, 31, 48, 96, 4), (2, 14, 96, 192, 4),
                                             (7, 6, 192, 384, 4)])
def test_conv2d_s2(dtype, N, H, Cin, Cout, k):
    with ops.use_dtype(dtype):
        conv = torch.nn.Conv2d(Cin, Cout, k, stride=2)
        u8 = Cin == 3
        if u8:
            x8 = torch.randint(0, 256, (N, Cin, H, H), dtype=torch.uint8)
            x = x8.float() / 255.0 - 0.5
        else:
            x = rnd(N, Cin, H, H).to(dtype).float()
        wq = conv.weight.detach().to(dtype).float()
        xr = x.clone().requires_grad_(True); wr = wq.clone().requires_grad_(True); br = conv.bias.detach().clone().requires_grad_(True)
        y = F.conv2d(xr, wr, br, stride=2)
        dy = rnd(*y.shape, seed=3).to(dtype).float()
        y.backward(dy)
        cd = conv.to(DEV)
        if u8:
            xd = x8.to(DEV)
        else:
            xd = x.permute(0, 2, 3, 1).reshape(N * H * H, Cin).to(DEV).to(dtype).requires_grad_(True)
        yd, Ho, Wo = ops.conv2d_s2(xd, cd.weight, cd.bias, N, H, H, k, u8=u8)
        assert (Ho, Wo) == tuple(y.shape[2:])
        dyd = dy.permute(0, 2, 3, 1).reshape(N * Ho * Wo, Cout).to(DEV).to(dtype)
        yd.backward(dyd)
        tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
        close(yd.reshape(N, Ho, Wo, Cout).permute(0, 3, 1, 2), y, msg="conv y", **tol)
        close(cd.weight.grad, wr.grad, msg="conv dW", **tol)
        close(cd.bias.grad, br.grad, msg="conv db", **tol)
        if not u8:
            close(xd.grad.reshape(N, H, H, Cin).permute(0, 3, 1, 2), xr.grad, msg="conv dx", **tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,Hi,Cin,Cout,k,last", [(3, 1, 32, 16, 5, False), (2, 5, 16, 8, 5, False), (2, 13, 8, 8, 6, False), (2, 30, 8, 3, 6, True),
                                                   (5, 5, 192, 96, 5, False), (3, 13, 96, 48, 6, False), (2, 30, 48, 3, 6, True)])
def test_conv_transpose2d_s2(dtype, N, Hi, Cin, Cout, k, last):
    with ops.use_dtype(dtype):
        conv = torch.nn.ConvTranspose2d(Cin, Cout, k, stride=2)
        x = rnd(N, Cin, Hi, Hi).to(dtype).float()
        wq = conv.weight.detach().to(dtype).float()
        xr = x.clone().requires_grad_(True); wr = wq.clone().requires_grad_(True); br = conv.bias.detach().clone().requires_grad_(True)
        y = F.conv_transpose2d(xr, wr, br, stride=2)
        dy = rnd(*y.shape, seed=3).to(dtype).float()
        y.backward(dy)
        cd = conv.to(DEV)
        xd = x.permute(0, 2, 3, 1).reshape(N * Hi * Hi, Cin).to(DEV).to(dtype).requires_grad_(True)
        yd, Ho, Wo = ops.conv_transpose2d_s2(xd, cd.weight, cd.bias, N, Hi, Hi, k, nchw_out=last)
        assert (Ho, Wo) == tuple(y.shape[2:])
        tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
        if last:
            yd.backward(dy.to(DEV))
            close(yd, y, msg="convT y (nchw)", **tol)
        else:
            yd.backward(dy.permute(0, 2, 3, 1).reshape(N * Ho * Wo, Cout).to(DEV).to(dtype))
            close(yd.reshape(N, Ho, Wo, Cout).permute(0, 3, 1, 2), y, msg="convT y", **tol)
        close(cd.weight.grad, wr.grad, msg="convT dW", **tol)
        close(cd.bias.grad, br.grad, msg="convT db", **tol)
        close(xd.grad.reshape(N, Hi, Hi, Cin).permute(0, 3, 1, 2), xr.grad, msg="convT dx", **tol)


def test_mse_and_means():
    rows = 7
    mean = rnd(rows, 3, 64, 64, scale=0.3)
    obs = torch.randint(0, 256, (rows, 3, 64, 64), dtype=torch.uint8)
    mr = mean.clone().requires_grad_(True)
    like = O.mse_logprob(mr[None], O.preprocess_obs(obs)[None])[0]
    g = rnd(rows, seed=1)
    like.backward(g)
    md = mean.to(DEV).requires_grad_(True)
    ld = ops.mse_logprob(md, obs.to(DEV))
    ld.backward(g.to(DEV))
    close(ld, like, msg="mse like"); close(md.grad, mr.grad, msg="mse grad")
    x = rnd(5000)
    ms = ops.mean_std(x.to(DEV))
    close(ms[0], x.mean(), msg="mean"); close(ms[1], x.std(), msg="std")
    xs = x.to(DEV).requires_grad_(True)
    s = ops.scaled_mean(xs, -2.0)
    s.backward()
    close(s[0], -2 * x.mean(), msg="scaled mean"); close(xs.grad, torch.full_like(x, -2.0 / x.numel()), msg="scaled mean grad")


@pytest.mark.parametrize("n", [37, 1000, 36000])
def test_quantile_ema(n):
    x = rnd(n, scale=3.0)
    ema = torch.tensor([0.3, 2.0])
    ref_ema = ema.clone()
    off, scale = O.reward_ema_update(x, ref_ema)
    ema_d = ema.to(DEV)
    out4 = torch.empty(4, device=DEV)
    lib.genrl_quantile_ema(ops._st(), x.to(DEV).data_ptr(), n, 0.05, 0.95, 1e-2, ema_d.data_ptr(), out4.data_ptr())
    q = torch.quantile(x, torch.tensor([0.05, 0.95]))
    close(out4[2:], q, rtol=1e-6, atol=1e-6, msg="quantiles")
    close(ema_d, ref_ema, rtol=1e-6, atol=1e-7, msg="ema")
    close(out4[0], off, msg="offset"); close(out4[1], scale, msg="scale")


def test_adam_step_matches_oracle():
    from genrl_b200.store import ParamStore
    torch.manual_seed(0)
    lin = torch.nn.Linear(33, 17)
    P = {"w": lin.weight.detach().clone(), "b": lin.bias.detach().clone()}
    st = O.AdamState(1e-3, 1e-8, 1.0, 1e-6)
    lin_d = torch.nn.Linear(33, 17)
    lin_d.load_state_dict(lin.state_dict())
    lin_d.to(DEV)
    store = ParamStore(DEV, torch.float32)
    grp = store.add_group("g", list(lin_d.parameters()))
    for step in range(3):
        gw, gb = rnd(17, 33, seed=step, scale=3.0), rnd(17, seed=step + 10, scale=3.0)
        norm = O.optimizer_step(P, ["w", "b"], {"w": gw, "b": gb}, st)
        ops.param_info(lin_d.weight).grad.copy_(gw.to(DEV)); ops.param_info(lin_d.bias).grad.copy_(gb.to(DEV))
        grp.touched = True
        nd = store.step([grp], 1e-3, 1e-8, 1.0, 1e-6)
        close(nd, norm, msg="grad norm")
        close(lin_d.weight, P["w"], rtol=1e-6, atol=1e-7, msg=f"adam w step {step}")
        close(lin_d.bias, P["b"], rtol=1e-6, atol=1e-7, msg=f"adam b step {step}")
        assert float(grp.grad.abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,D,Hd,resets", [(3, 5, 64, 64, True), (50, 8, 256, 128, True), (7, 6, 1024, 1024, False), (70, 3, 128, 64, True), (4, 3, 1536, 256, True)])
def test_gru_scan_matches_stepwise(dtype, B, T, D, Hd, resets):
    """persistent scan kernel == T sequential GRU steps (forward values, dgx, dW_h, dgamma, dh0)"""
    with ops.use_dtype(dtype):
        torch.manual_seed(1)
        w = torch.randn(3 * D, Hd + D) / math.sqrt(Hd + D)
        g, b = 1 + 0.1 * rnd(3 * D, seed=2), 0.1 * rnd(3 * D, seed=3)
        gx = rnd(B * T, 3 * D, seed=4)
        h0 = rnd(B, D, seed=5) * 0.5
        mask = torch.ones(B, T)
        if resets:
            mask[:, 0] = 0; mask[0, T // 2] = 0
        dd = rnd(B * T, D, seed=6)
        # reference: explicit loop in fp32 autograd
        wr = w.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
        gxr = gx.clone().requires_grad_(True); h0r = h0.clone().requires_grad_(True)
        h = h0r
        outs = []
        gx3 = gxr.reshape(B, T, 3 * D)
        for t in range(T):
            hm = h * mask[:, t:t + 1]
            z = gx3[:, t] + hm @ wr[:, Hd:].t()
            parts = F.layer_norm(z, (3 * D,), gr, br, 1e-5)
            r, c, u = torch.chunk(parts, 3, -1)
            r = torch.sigmoid(r); c = torch.tanh(r * c); u = torch.sigmoid(u - 1)
            h = u * c + (1 - u) * hm
            outs.append(h)
        deter = torch.stack(outs, 1).reshape(B * T, D)
        (deter * dd).sum().backward()
        wd = torch.nn.Parameter(w.to(DEV)); gd = torch.nn.Parameter(g.to(DEV)); bd = torch.nn.Parameter(b.to(DEV))
        gxd = gx.to(DEV).requires_grad_(True); h0d = h0.to(DEV).requires_grad_(True)
        dt = ops.gru_scan(gxd, wd, gd, bd, mask.reshape(-1).to(DEV).contiguous(), h0d, B, T, Hd)
        (dt * dd.to(DEV)).sum().backward()
        tol = dict(rtol=1e-4, atol=2e-5) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
        # fp32 mode: FFMA recurrence; bf16 mode (D % 128 == 0): the recurrent product runs on bf16 tensor cores, state stays fp32
        close(dt, deter, msg="scan deter", **(dict(rtol=1e-4, atol=2e-5) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)))
        close(gxd.grad, gxr.grad, msg="scan dgx", **tol)
        close(h0d.grad, h0r.grad, msg="scan dh0", **tol)
        close(gd.grad, gr.grad, msg="scan dgamma", **tol)
        close(bd.grad, br.grad, msg="scan dbeta", **tol)
        close(wd.grad[:, Hd:], wr.grad[:, Hd:], msg="scan dW_h", **tol)


@pytest.mark.parametrize("B,T,D,Hd,resets", [(16, 4, 2048, 512, True), (7, 3, 4096, 256, True), (1, 5, 1280, 128, False)])
def test_gru_scan_weight_streaming_matches_stepwise(B, T, D, Hd, resets, monkeypatch):
    """hidden sizes above 8 x #SMs (BASELINE configs[4]: D = 4096) run the weight-streaming scan (csrc/rssm_scan_stream.cu): forward values
    and every gradient against T sequential fp32 GRU steps; the op-level path it replaces gives the same numbers"""
    with ops.use_dtype(torch.bfloat16):
        assert not ops.scan_supported(D) and ops.scan_stream_supported(B, D)
        torch.manual_seed(1)
        w = torch.randn(3 * D, Hd + D) / math.sqrt(Hd + D)
        g, b = 1 + 0.1 * rnd(3 * D, seed=2), 0.1 * rnd(3 * D, seed=3)
        gx = rnd(B * T, 3 * D, seed=4)
        h0 = rnd(B, D, seed=5) * 0.5
        mask = torch.ones(B, T)
        if resets:
            mask[:, 0] = 0; mask[0, T // 2] = 0
        dd = rnd(B * T, D, seed=6)
        wq = w.to(torch.bfloat16).float()
        wr = wq.to(DEV).requires_grad_(True); gr = g.to(DEV).requires_grad_(True); br = b.to(DEV).requires_grad_(True)
        gxr = gx.to(DEV).requires_grad_(True); h0r = h0.to(DEV).requires_grad_(True)
        md = mask.to(DEV)
        h = h0r
        outs = []
        gx3 = gxr.reshape(B, T, 3 * D)
        for t in range(T):
            hm = h * md[:, t:t + 1]
            z = gx3[:, t] + hm @ wr[:, Hd:].t()
            parts = F.layer_norm(z, (3 * D,), gr, br, 1e-5)
            r, c, u = torch.chunk(parts, 3, -1)
            r = torch.sigmoid(r); c = torch.tanh(r * c); u = torch.sigmoid(u - 1)
            h = u * c + (1 - u) * hm
            outs.append(h)
        deter = torch.stack(outs, 1).reshape(B * T, D)
        (deter * dd.to(DEV)).sum().backward()

        def run():
            wd = torch.nn.Parameter(w.to(DEV)); gd = torch.nn.Parameter(g.to(DEV)); bd = torch.nn.Parameter(b.to(DEV))
            gxd = gx.to(DEV).requires_grad_(True); h0d = h0.to(DEV).requires_grad_(True)
            dt = ops.gru_scan(gxd, wd, gd, bd, md.reshape(-1).contiguous(), h0d, B, T, Hd)
            (dt * dd.to(DEV)).sum().backward()
            return dt, gxd.grad, h0d.grad, gd.grad, bd.grad, wd.grad[:, Hd:]

        got = run()
        ref = (deter, gxr.grad, h0r.grad, gr.grad, br.grad, wr.grad[:, Hd:])
        tols = [dict(rtol=1e-2, atol=1e-2)] + [dict(rtol=3e-2, atol=3e-2)] * 5
        for name, a, e, tol in zip(("deter", "dgx", "dh0", "dgamma", "dbeta", "dW_h"), got, ref, tols):
            close(a, e, msg="stream scan " + name, **tol)
        monkeypatch.setenv("GENRL_SCAN_STREAM", "0")
        old = run()
        close(got[0], old[0], msg="stream vs op-level deter", rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("M,N,K", [(1024, 1024, 36000), (1024, 2048, 19200), (3072, 1024, 9000), (256, 512, 8192), (128, 128, 300), (1536, 1024, 2400)])
def test_gemm_pair_wgrad_split_k(M, N, K):
    """weight-gradient form of the 2-CTA kernel (csrc/gemm_pair.cu: both operands MN-major, split-K, fp32 atomic epilogue): C += A^T W against
    an fp32 matmul of the same bf16 operands, C pre-filled (accumulate), ragged last k-block"""
    from genrl_b200._lib import lib
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(K, M, generator=g) * 0.5).to(DEV, torch.bfloat16)
    w = (torch.randn(K, N, generator=g) * 0.5).to(DEV, torch.bfloat16)
    c0 = torch.randn(M, N, generator=g).to(DEV)
    c = c0.clone()
    rc = lib.load().genrl_gemm_pair_wgrad(ops._st(), M, N, K, a.data_ptr(), M, w.data_ptr(), N, c.data_ptr(), N)
    assert rc == 0, lib.load().genrl_last_error()
    ref = c0 + a.float().t() @ w.float()
    close(c, ref, rtol=2e-3, atol=2e-2 * math.sqrt(K / 1000), msg="pair wgrad")
    # the dispatcher sends the long-K accumulate form here and gives the same numbers as the 128x128 split-K kernel within fp32 atomics noise
    c2 = c0.clone()
    ops.gemm(a, w, out=c2, trans_a=True, trans_w=True, accumulate=True)
    close(c2, ref, rtol=2e-3, atol=2e-2 * math.sqrt(K / 1000), msg="dispatched wgrad")


@pytest.mark.parametrize("dtype,T,N,E,align", [(torch.float32, 11, 37, 96, True), (torch.float32, 16, 64, 1536, True),
                                               (torch.bfloat16, 16, 50, 1536, True), (torch.float32, 5, 9, 64, False)])
def test_vlm_reward_alignment_and_gradient(dtype, T, N, E, align):
    """fused video-text reward (tools/genrl_utils.py:240-243, :250-277, :344-366) vs the oracle's max_cosine + window alignment on the same
    projections: chosen offsets bit-equal, rewards and d reward / d projection within the mode's tolerance"""
    g = torch.Generator().manual_seed(3)
    pa = torch.randn(T, N, E, generator=g) * (0.5 + torch.rand(T, N, 1, generator=g))
    pt = torch.randn(T, N, E, generator=g)
    # make a clear best offset for some rows: the imagined sequence repeats the target's head with a delay
    for n in range(0, N, 3):
        d = n % max(T - 8, 1) if align else 0
        pa[d:d + min(8, T - d), n] = pt[:min(8, T - d), n] * 1.3 + 0.05 * torch.randn(min(8, T - d), E, generator=g)
    q = lambda x: x.to(dtype).float()
    pa_r = q(pa).clone().requires_grad_(True)
    F_ = 8
    if align:
        win = torch.stack([O.max_cosine(q(pt)[:F_], pa_r[t:t + F_]).mean(0) for t in range(T - F_)], 0)
        b = win.argmax(0)
    else:
        b = torch.zeros(N, dtype=torch.long)
    idx = (torch.arange(T)[:, None] - b[None]).clamp(min=0)
    shifted = torch.gather(q(pt), 0, idx[:, :, None].expand(T, N, E))
    ref = O.max_cosine(shifted, pa_r)
    w = torch.randn(T, N, generator=g)
    (ref * w).sum().backward()
    pa_d = pa.to(DEV, dtype).requires_grad_(True)
    reward, shift = ops.vlm_reward(pa_d, pt.to(DEV, dtype), F_, align)
    (reward * w.to(DEV)).sum().backward()
    assert torch.equal(shift.cpu().long(), b)
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert (reward.cpu() - ref.detach()).abs().max() <= tol
    gscale = float(pa_r.grad.abs().max())
    assert (pa_d.grad.float().cpu() - pa_r.grad).abs().max() <= tol * gscale + (1e-6 if dtype == torch.float32 else 1e-2 * gscale)


@pytest.mark.parametrize("M,K0,N,L", [(2400, 2048, 1024, 4), (300, 1034, 1024, 1), (128, 256, 256, 3), (5000, 512, 512, 2), (77, 1024, 768, 2),
                                      (5000, 1034, 1024, 2), (513, 1024, 1024, 3), (8192, 1024, 1024, 1)])
def test_mlp_chain_fwd_matches_layerwise_reference(M, K0, N, L):
    """persistent cluster chain kernel (csrc/mlp_chain.cu) vs fp32 torch on the same bf16 operands, layer by layer: pre-activations,
    (mean, rstd), activations; each layer's reference consumes the KERNEL's previous activation so that errors do not compound"""
    g = torch.Generator().manual_seed(M + N)
    pad = lambda t: torch.nn.functional.pad(t, (0, (-t.shape[1]) % 8))[:, :t.shape[1]]          # TMA-legal leading dimension
    x = pad((torch.randn(M, K0, generator=g)).to(DEV, torch.bfloat16))
    Ws = [pad((torch.randn(N, K0 if l == 0 else N, generator=g) / math.sqrt(K0 if l == 0 else N)).to(DEV, torch.bfloat16)) for l in range(L)]
    bs = [(torch.randn(N, generator=g) * 0.1).to(DEV) for _ in range(L)]
    gs = [(1 + 0.1 * torch.randn(N, generator=g)).to(DEV) for _ in range(L)]
    bt = [(0.1 * torch.randn(N, generator=g)).to(DEV) for _ in range(L)]
    slab = lambda: [torch.empty(M, N + 8, device=DEV, dtype=torch.bfloat16)[:, :N] for _ in range(L)]
    pre, y = slab(), slab()
    stats = [torch.empty(M, 2, device=DEV) for _ in range(L)]
    with ops.use_dtype(torch.bfloat16):
        ops.mlp_chain_fwd(x, Ws, bs, gs, bt, 1e-5, ops.ACT_SILU, pre, y, stats)
    torch.cuda.synchronize()
    inp = x
    for l in range(L):
        ref_pre = inp.float() @ Ws[l].float().t() + bs[l]
        close(pre[l], ref_pre, rtol=1e-2, atol=2e-2, msg=f"layer {l} pre")
        mean, var = ref_pre.mean(-1), ref_pre.var(-1, unbiased=False)
        close(stats[l][:, 0], mean, rtol=1e-3, atol=1e-3, msg=f"layer {l} mean")
        close(stats[l][:, 1], 1 / torch.sqrt(var + 1e-5), rtol=2e-3, atol=1e-4, msg=f"layer {l} rstd")
        ref_y = F.silu(F.layer_norm(ref_pre, (N,), gs[l], bt[l], 1e-5))
        close(y[l], ref_y, rtol=1e-2, atol=2e-2, msg=f"layer {l} y")
        inp = y[l]


@pytest.mark.parametrize("tw,M,N,K", [(False, 2400, 1024, 1024), (False, 5000, 1024, 2048), (True, 2400, 1024, 1024), (False, 300, 256, 1040),
                                      (True, 38400, 2048, 1024), (False, 1000, 3072, 2048), (False, 257, 384, 64)])
def test_gemm_pair_2cta(tw, M, N, K):
    """2-CTA tcgen05 GEMM (csrc/gemm_pair.cu) vs an fp32 matmul of the same bf16 operands, K-major and MN-major weights, bias, ragged tiles"""
    g = torch.Generator().manual_seed(M + N + K)
    pad = lambda t: torch.nn.functional.pad(t, (0, (-t.shape[1]) % 8))[:, :t.shape[1]]
    a = pad(torch.randn(M, K, generator=g).to(DEV, torch.bfloat16))
    w = pad((torch.randn(K, N, generator=g) if tw else torch.randn(N, K, generator=g)).to(DEV, torch.bfloat16))
    bias = torch.randn(N, generator=g).to(DEV)
    for out_dtype in (torch.bfloat16, torch.float32):
        out = ops.gemm_pair(a, w, bias=bias, trans_w=tw, out_dtype=out_dtype)
        assert out is not None
        ref = a.float() @ (w.float() if tw else w.float().t()) + bias
        close(out, ref, rtol=1e-2 if out_dtype == torch.bfloat16 else 2e-3, atol=1e-1 if out_dtype == torch.bfloat16 else 2e-2, msg=f"pair gemm {out_dtype}")
    # several launches back to back (persistent pipeline state / barrier phases start clean every launch)
    for _ in range(3):
        out2 = ops.gemm_pair(a, w, bias=bias, trans_w=tw, out_dtype=torch.float32)
    close(out2, ref, rtol=2e-3, atol=2e-2, msg="pair gemm repeat")


@pytest.mark.parametrize("ta,tw,M,N,K,acc", [(False, False, 300, 256, 512, False), (False, True, 2400, 1024, 1034, False), (True, True, 384, 520, 2400, True),
                                             (False, False, 64, 1024, 2048, False)])
def test_gemm_fp32_on_tensor_cores_bf16x3(ta, tw, M, N, K, acc):
    """fp32 parity mode above 2^24 MACs: three-term bf16 split of both operands, six tcgen05 products accumulated in fp32 — against a float64
    matmul the error must stay at the fp32 level (1e-6 of the operand scale x sqrt(K)), i.e. what the FFMA kernel gives"""
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((K, M) if ta else (M, K), generator=g)
    w = torch.randn((K, N) if tw else (N, K), generator=g)
    bias = None if acc else torch.randn(N, generator=g)
    c0 = torch.randn(M, N, generator=g)
    ref = (a.t() if ta else a).double() @ (w if tw else w.t()).double() + (c0.double() if acc else bias.double())
    out = c0.clone().to(DEV) if acc else None
    with ops.use_dtype(torch.float32):
        assert M * N * K >= ops.FP32_TC_MIN_WORK and ops.FP32_TC
        res = ops.gemm(a.to(DEV), w.to(DEV), out=out, bias=None if acc else bias.to(DEV), trans_a=ta, trans_w=tw, accumulate=acc)
    err = float((res.double().cpu() - ref).abs().max())
    assert err <= 4e-6 * math.sqrt(K) * 4.0, f"max abs err {err:.3e}"
    simt_like = (a.t() if ta else a) @ (w if tw else w.t())          # torch fp32 on CPU as the yardstick
    err_ref = float((simt_like.double() + (c0.double() if acc else bias.double()) - ref).abs().max())
    assert err <= 8 * err_ref + 1e-5, (err, err_ref)


def test_mlp_chain_is_deterministic_under_repetition():
    """layer l+1 of the chain kernel reads y_l with TMA right after other CTAs of the cluster wrote it with ordinary stores (generic -> async
    proxy fence + cluster barrier): 300 back-to-back launches on changing inputs must reproduce the first result of each input bit for bit,
    and interleaved launches on a second stream must not disturb them"""
    M, K0, N, L = 2400, 2048, 1024, 4
    g = torch.Generator().manual_seed(7)
    xs = [torch.randn(M, K0, generator=g).to(DEV, torch.bfloat16) for _ in range(3)]
    Ws = [(torch.randn(N, K0 if l == 0 else N, generator=g) / math.sqrt(K0 if l == 0 else N)).to(DEV, torch.bfloat16) for l in range(L)]
    bs = [(0.1 * torch.randn(N, generator=g)).to(DEV) for _ in range(L)]
    gs = [(1 + 0.1 * torch.randn(N, generator=g)).to(DEV) for _ in range(L)]

    def run(x):
        pre = [torch.empty(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(L)]
        y = [torch.empty(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(L)]
        st = [torch.empty(M, 2, device=DEV) for _ in range(L)]
        ops.mlp_chain_fwd(x, Ws, bs, gs, bs, 1e-5, ops.ACT_SILU, pre, y, st)
        return y[-1], pre[-1]

    with ops.use_dtype(torch.bfloat16):
        ref = [tuple(t.clone() for t in run(x)) for x in xs]
        side = torch.cuda.Stream()
        junk = torch.randn(4096, 4096, device=DEV)
        bad = 0
        for i in range(300):
            if i % 3 == 0:
                with torch.cuda.stream(side):
                    junk = junk @ junk.t() * 1e-4          # unrelated traffic through L2 on another stream
            y, pre = run(xs[i % 3])
            bad += int(not torch.equal(y, ref[i % 3][0])) + int(not torch.equal(pre, ref[i % 3][1]))
        torch.cuda.synchronize()
    assert bad == 0, f"{bad} of 600 outputs differ from the first run"
