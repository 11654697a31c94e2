#!/usr/bin/env python
"""The narrow (<= 128-channel) tensor-core layers that cap configs[2] / [4] (vae-gan-unet.py 256x256): forward, data
gradient and weight gradient of each, timed inside a CUDA graph (dev tool).

    python tools/gpu_narrow_conv.py            # table: ms and TFLOP/s per primitive
    python tools/gpu_narrow_conv.py once IDX   # shape IDX once per primitive, for an `ncu --set full` capture
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_gan_mark_b200.conv import ConvLinear, new_act  # noqa: E402

# (n, h, w, cin, cout, k)
SHAPES = [(32, 256, 256, 64, 64, 3), (64, 128, 128, 64, 64, 3), (32, 128, 128, 128, 128, 3), (32, 128, 128, 64, 128, 3),
          (64, 64, 64, 128, 128, 3), (32, 256, 256, 64, 64, 1), (32, 128, 128, 128, 64, 1), (32, 64, 64, 256, 256, 3)]


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


def setup(n, h, w, cin, cout, k):
    op = ConvLinear(cin, cout, k, k, 1, (k // 2, k // 2))
    g = torch.Generator().manual_seed(1)
    x = new_act(n, h, w, cin, "cuda"); x.copy_(torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16))
    dy = new_act(n, h, w, cout, "cuda"); dy.copy_(torch.randn(n, h, w, cout, generator=g).to(torch.bfloat16))
    wt = (torch.randn(cout, cin, k, k, generator=g) * 0.05).cuda()
    wf, wb = op.prep_fwd(wt), op.prep_bwd(wt)
    out, dx = new_act(n, h, w, cout, "cuda"), new_act(n, h, w, cin, "cuda")
    return op, x, dy, wf, wb, out, dx


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "once":
        shp = SHAPES[int(sys.argv[2])]
        op, x, dy, wf, wb, out, dx = setup(*shp)
        for _ in range(2):
            op.forward(x, wf, None, 0, out=out)
            op.backward_data(dy, wb, (shp[1], shp[2]), out=dx)
            op.backward_weight(dy, x)
        torch.cuda.synchronize()
        print("ok", shp)
        return
    print(f"{'shape (n,h,w,cin,cout,k)':>32s} {'fwd ms':>8s} {'TF/s':>6s} {'dgrad ms':>9s} {'TF/s':>6s} {'wgrad ms':>9s} {'TF/s':>6s} "
          f"{'HBM floor ms (fwd)':>18s}")
    for shp in SHAPES:
        n, h, w, cin, cout, k = shp
        op, x, dy, wf, wb, out, dx = setup(*shp)
        fl = 2.0 * n * h * w * cin * cout * k * k
        tf = timeit(lambda: op.forward(x, wf, None, 0, out=out))
        td = timeit(lambda: op.backward_data(dy, wb, (h, w), out=dx))
        tw = timeit(lambda: op.backward_weight(dy, x))
        floor = (x.numel() + out.numel()) * 2 / 6.5e9
        print(f"{str(shp):>32s} {tf:8.3f} {fl / tf / 1e9:6.0f} {td:9.3f} {fl / td / 1e9:6.0f} {tw:9.3f} {fl / tw / 1e9:6.0f} {floor:18.3f}",
              flush=True)
        del op, x, dy, wf, wb, out, dx
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
