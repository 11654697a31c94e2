"""Achieved HBM bandwidth of the elementwise / normalisation kernels on cfg-2 sized tensors (development tool).

Prints one line per (kernel, shape): algorithmic bytes (each tensor read/written once at its storage dtype), average
CUDA-event time over 10 launches after 3 warm-ups, GB/s and the fraction of the measured copy peak.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from vae_gan_mark_b200 import ops  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timeit(fn, iters=10, warm=3):
    """Average device time of one call: `iters` calls are captured into one CUDA graph so that the host-side cost of
    issuing a call (ctypes + Python, ~20 us) does not hide the duration of short kernels."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
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


def report(name, shape, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print(json.dumps({"kernel": name, "shape": list(shape), "MB": round(nbytes / 1e6, 1), "ms": round(ms, 4),
                      "GB/s": round(gbs, 1), "frac_of_copy_peak": round(gbs / PEAK, 3)}), flush=True)


def main():
    dev = "cuda"
    shapes = [(64, 128, 128, 512), (64, 128, 128, 64), (64, 64, 64, 128), (64, 32, 32, 256), (64, 16, 16, 512)]
    if len(sys.argv) > 1:
        shapes = shapes[:int(sys.argv[1])]
    for (n, h, w, c) in shapes:
        x = torch.randn(n, h, w, c, device=dev).to(BF16)
        dy = torch.randn(n, h, w, c, device=dev).to(BF16)
        dp = torch.randn(n, h // 2, w // 2, c, device=dev).to(BF16)
        gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
        y, dx = torch.empty_like(x), torch.empty_like(x)
        pool = torch.empty(n, h // 2, w // 2, c, device=dev, dtype=BF16)
        dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
        e = x.numel() * 2
        report("norm_stats", x.shape, e, timeit(lambda: ops.norm_stats(x, False)))
        sums = ops.norm_stats(x, False)
        mr = ops.norm_finalize(sums, n * h * w, 1e-5)
        report("norm_apply", x.shape, 2 * e, timeit(lambda: ops.norm_apply(x, mr, gamma, beta, 1, y)))
        report("norm_apply+pool", x.shape, 2.25 * e, timeit(lambda: ops.norm_apply(x, mr, gamma, beta, 1, y, pool)))
        # backward = reduce pass (reads x, dy) + apply pass (reads x, dy, writes dx): 5 tensor passes
        report("norm_backward", x.shape, 5 * e,
               timeit(lambda: ops.norm_backward(x, dy, None, mr, False, gamma, beta, 1, dx, dg, db)))
        report("norm_backward+pool", x.shape, 5.5 * e,
               timeit(lambda: ops.norm_backward(x, dy, dp, mr, False, gamma, beta, 1, dx, dg, db)))
        if c >= 128:
            gb = torch.randn(n, h, w, 2 * c, device=dev).to(BF16)
            dgb = torch.empty_like(gb)
            report("film_fwd", x.shape, 4 * e, timeit(lambda: ops.film_fwd(gb, x, y)))
            report("film_bwd", x.shape, 7 * e, timeit(lambda: ops.film_bwd(gb, x, dy, dgb, dx)))
    # weight re-layout (fp32 OIHW -> bf16 [O][kh][kw][I]) and multi-tensor Adam on a 512x512x3x3 weight
    wt = torch.randn(512, 512, 3, 3, device=dev)
    out = torch.empty(512, 3, 3, 512, device=dev, dtype=BF16)
    report("strided_copy OIHW->OHWI bf16", wt.shape, wt.numel() * 6, timeit(lambda: ops.strided_copy(wt.permute(0, 2, 3, 1), out)))
    out2 = torch.empty(3, 3, 512, 512, device=dev, dtype=BF16)
    report("strided_copy OIHW->HWIO bf16", wt.shape, wt.numel() * 6, timeit(lambda: ops.strided_copy(wt.permute(2, 3, 1, 0), out2)))
    gw = torch.randn(512, 3, 3, 512, device=dev)
    out3 = torch.empty(512, 512, 3, 3, device=dev)
    report("strided_copy OHWI->OIHW fp32", gw.shape, gw.numel() * 8, timeit(lambda: ops.strided_copy(gw.permute(0, 3, 1, 2), out3)))
    # text-map upsample (write-only) and its backward (read-only), image-side im2col
    for (n, h, w, c) in [(64, 128, 128, 512), (64, 64, 64, 512)]:
        tm = torch.randn(n, 1, w // 16, c, device=dev).to(BF16)
        up = torch.empty(n, h, w, c, device=dev, dtype=BF16)
        dtm = torch.empty(n, 1, w // 16, c, device=dev)
        report("upsample_w_fwd", up.shape, up.numel() * 2, timeit(lambda: ops.upsample_w_fwd(tm, up)))
        report("upsample_w_bwd", up.shape, up.numel() * 2, timeit(lambda: ops.upsample_w_bwd(up, dtm)))
    img = torch.randn(64, 128, 128, 8, device=dev).to(BF16)
    col = torch.empty(64, 128, 128, 64, device=dev, dtype=BF16)
    report("im2col 3x3 c4 -> 64", col.shape, col.numel() * 2 + img.numel() * 2, timeit(lambda: ops.im2col(img, 4, 3, 3, 1, 1, col)))
    hw = torch.randn(256, 65536, device=dev)
    hb = torch.empty(256, 65536, device=dev, dtype=BF16)
    report("strided_copy contiguous fp32->bf16 16.7M", hw.shape, hw.numel() * 6, timeit(lambda: ops.strided_copy(hw, hb)))
    t = torch.randn(1024, 1024, 64, device=dev)
    report("torch copy (reference point)", t.shape, t.numel() * 8, timeit(lambda: t.clone()))


if __name__ == "__main__":
    main()
