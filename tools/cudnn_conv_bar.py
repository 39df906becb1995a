"""cuDNN bar for the conv stacks: torch F.conv2d / F.conv_transpose2d, bf16, channels_last, cudnn.benchmark, on the
8 layers of the encoder/decoder at the benchmark batch (2 400 frames), forward and backward (dgrad + wgrad) timed with CUDA events,
next to this library's implicit-GEMM kernels on the same shapes.  Output: one table (us per pass)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

torch.backends.cudnn.benchmark = True
dev = "cuda:0"
N = int(os.environ.get("FRAMES", "2400"))
depth = 48
layers = []
h, cin = 64, 3
for i in range(4):
    cout = depth * 2 ** i
    layers.append(("enc%d" % i, "conv", cin, cout, 4, h)); h = (h - 4) // 2 + 1; cin = cout
h, cin = 1, 32 * depth
for i, k in enumerate((5, 5, 6, 6)):
    cout = 3 if i == 3 else depth * 2 ** (2 - i)
    layers.append(("dec%d" % i, "convT", cin, cout, k, h)); h = (h - 1) * 2 + k; cin = cout


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def ours(kind, cin, cout, k, h):
    from genrl_b200 import ops
    with ops.use_dtype(torch.bfloat16):
        if kind == "conv":
            w = torch.randn(cout, cin, k, k, device=dev, requires_grad=True) * 0.05
            w = w.detach().requires_grad_(True)
            b = torch.zeros(cout, device=dev, requires_grad=True)
            if cin == 3:
                x = torch.randint(0, 256, (N, 3, h, h), device=dev, dtype=torch.uint8)
                f = lambda: ops.conv2d_s2(x, w, b, N, h, h, k, u8=True)[0]
            else:
                x = torch.randn(N * h * h, cin, device=dev).bfloat16().requires_grad_(True)
                f = lambda: ops.conv2d_s2(x, w, b, N, h, h, k)[0]
        else:
            w = (torch.randn(cin, cout, k, k, device=dev) * 0.05).requires_grad_(True)
            b = torch.zeros(cout, device=dev, requires_grad=True)
            x = torch.randn(N * h * h, cin, device=dev).bfloat16().requires_grad_(True)
            f = lambda: ops.conv_transpose2d_s2(x, w, b, N, h, h, k, nchw_out=(cout == 3))[0]
        t_f = timeit(f)
        y = f()
        g = torch.randn_like(y)

        def fb():
            y = f()
            y.backward(g)
        t_fb = timeit(fb)
    return t_f, t_fb - t_f


def cudnn(kind, cin, cout, k, h):
    x = torch.randn(N, cin, h, h, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(cin != 3)
    if kind == "conv":
        w = (torch.randn(cout, cin, k, k, device=dev) * 0.05).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        f = lambda: F.conv2d(x, w, None, stride=2)
    else:
        w = (torch.randn(cin, cout, k, k, device=dev) * 0.05).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        f = lambda: F.conv_transpose2d(x, w, None, stride=2)
    t_f = timeit(f)
    y = f()
    g = torch.randn_like(y)

    def fb():
        y = f()
        y.backward(g)
    t_fb = timeit(fb)
    return t_f, t_fb - t_f


print(f"frames {N}; us per pass (fwd | bwd = dgrad + wgrad); cuDNN via torch {torch.__version__}, bf16 channels_last, benchmark=True")
print(f"{'layer':6s} {'shape':28s} {'GFLOP fwd':>9s} | {'ours fwd':>9s} {'cudnn fwd':>9s} | {'ours bwd':>9s} {'cudnn bwd':>9s}")
tot = [0, 0, 0, 0]
for name, kind, cin, cout, k, h in layers:
    ho = (h - k) // 2 + 1 if kind == "conv" else (h - 1) * 2 + k
    macs = (ho * ho * cout * cin * k * k) if kind == "conv" else (h * h * cin * cout * k * k)
    try:
        o = ours(kind, cin, cout, k, h)
    except Exception as e:
        o = (float("nan"), float("nan")); print("ours failed:", e)
    c = cudnn(kind, cin, cout, k, h)
    for i, v in enumerate((o[0], c[0], o[1], c[1])):
        tot[i] += v
    print(f"{name:6s} {kind:5s} {cin:4d}->{cout:4d} k{k} {h:2d}->{ho:2d}   {2e-9 * macs * N:9.2f} | {o[0]:9.1f} {c[0]:9.1f} | {o[1]:9.1f} {c[1]:9.1f}")
print(f"{'sum':6s} {'':28s} {'':9s} | {tot[0]:9.1f} {tot[1]:9.1f} | {tot[2]:9.1f} {tot[3]:9.1f}")
