"""Time the small-N (cout <= 4) pointwise conv kernels on the shapes of final_image_conv (dev tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_gan_mark_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for (n, h, w) in [(64, 128, 128), (32, 256, 256)]:
    cin, cout = 64, 3
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = torch.randn(cout, 1, 1, cin, device="cuda")
    bias = torch.randn(cout, device="cuda")
    out = torch.empty(n, h, w, cout, device="cuda")
    dy = torch.randn(n, h, w, cout, device="cuda")
    dx = torch.empty(n, h, w, cin, device="cuda", dtype=torch.bfloat16)
    dw = torch.empty(cout, 1, 1, cin, device="cuda")
    db = torch.empty(cout, device="cuda")
    mb = x.numel() * 2 / 1e6
    t = timeit(lambda: ops.smalln_fwd(x, wt, bias, 1, 1, 0, out))
    print(f"{n}x{h}x{w} fwd   {t * 1e3:8.1f} us  ({mb / t / 1e3:.2f} TB/s of x)")
    t = timeit(lambda: ops.smalln_dgrad(dy, wt, 1, 1, 0, dx))
    print(f"{n}x{h}x{w} dgrad {t * 1e3:8.1f} us  ({mb / t / 1e3:.2f} TB/s of dx)")
    t = timeit(lambda: ops.smalln_wgrad(dy, x, 1, 1, 0, dw, db))
    print(f"{n}x{h}x{w} wgrad {t * 1e3:8.1f} us  ({mb / t / 1e3:.2f} TB/s of x)")
