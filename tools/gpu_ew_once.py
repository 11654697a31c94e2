"""Every HBM-bound kernel of the step (normalisation, FiLM, resize, losses, optimiser) once, at the shapes of the headline
workload (vae-gan-v2 128x128 b64) -- the target of an `ncu --profile-from-start off` capture (dev tool).

Each kernel is launched twice untimed, then five times between CUDA events (plain launches, no CUDA graph: the shapes are
large enough that the ~20 us host cost per call hides behind the previous launch), then ONCE between
cudaProfilerStart / cudaProfilerStop.  Without ncu the script prints the CUDA-event bandwidth of every kernel against the
algorithmic bytes (each tensor read / written once at its storage dtype); under ncu the bracketed launches give
dram__bytes_read/write and gpu__dram_throughput of the same kernels on the same tensors.
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
RT = torch.cuda.cudart()


def run(name, shape, nbytes, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    RT.cudaProfilerStart()
    fn()
    torch.cuda.synchronize()
    RT.cudaProfilerStop()
    gbs = nbytes / ms / 1e6
    print(json.dumps({"kernel": name, "shape": list(shape), "algorithmic_MB": round(nbytes / 1e6, 1), "ms": round(ms, 4),
                      "GB/s": round(gbs, 1), "frac_of_copy_peak": round(gbs / PEAK, 3)}), flush=True)


def main():
    dev = "cuda"
    for (n, h, w, c) in [(64, 128, 128, 512), (64, 128, 128, 64), (64, 64, 64, 128)]:
        x = torch.randn(n, h, w, c, device=dev).to(BF16)
        dy = torch.randn(n, h, w, c, device=dev).to(BF16)
        dp = torch.randn(n, h // 2, w // 2, c, device=dev).to(BF16)
        gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
        y, dx = torch.empty_like(x), torch.empty_like(x)
        pool = torch.empty(n, h // 2, w // 2, c, device=dev, dtype=BF16)
        dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
        e = x.numel() * 2
        run("norm_stats", x.shape, e, lambda: ops.norm_stats(x, False))
        mr = ops.norm_finalize(ops.norm_stats(x, False), n * h * w, 1e-5)
        run("norm_apply(+ReLU)", x.shape, 2 * e, lambda: ops.norm_apply(x, mr, gamma, beta, 1, y))
        if c < 512:
            run("norm_apply(+ReLU)+pool", x.shape, 2.25 * e, lambda: ops.norm_apply(x, mr, gamma, beta, 1, y, pool))
        # backward = reduce pass (reads x, dy) + apply pass (reads x, dy, writes dx): two launches, 5 tensor passes
        run("norm_backward (reduce + apply)", x.shape, 5 * e,
            lambda: ops.norm_backward(x, dy, None, mr, False, gamma, beta, 1, dx, dg, db))
        if c < 512:
            run("norm_backward+pool (reduce + apply)", x.shape, 5.5 * e,
                lambda: ops.norm_backward(x, dy, dp, mr, False, gamma, beta, 1, dx, dg, db))
        if c == 512:
            gb = torch.randn(n, h, w, 2 * c, device=dev).to(BF16)
            dgb = torch.empty_like(gb)
            run("film_fwd", x.shape, 4 * e, lambda: ops.film_fwd(gb, x, y))
            run("film_bwd", x.shape, 7 * e, lambda: ops.film_bwd(gb, x, dy, dgb, dx))
            del gb, dgb
            tm = torch.randn(n, 1, w // 16, c, device=dev).to(BF16)
            dtm = torch.empty(n, 1, w // 16, c, device=dev)
            run("upsample_w_fwd", x.shape, e, lambda: ops.upsample_w_fwd(tm, y))
            run("upsample_w_bwd", x.shape, e, lambda: ops.upsample_w_bwd(dy, dtm))
        del x, dy, dp, y, dx, pool
    # losses at the step's own sizes: L1 over the 64 x 3 x 128 x 128 image (fp32), hinge over the discriminator's patch
    # logits, reparameterisation + KL over the (64, 2 x 128) heads -- all far below the L2 size (latency-, not HBM-bound)
    a, b = torch.rand(64, 3, 128, 128, device=dev), torch.rand(64, 3, 128, 128, device=dev)
    da, one = torch.empty_like(a), torch.ones((), device=dev)
    run("l1_fwd", a.shape, a.numel() * 8, lambda: ops.l1_fwd(a, b))
    run("l1_bwd", a.shape, a.numel() * 12, lambda: ops.l1_bwd(a, b, one, da))
    # the same kernels on a tensor larger than L2 (bandwidth, not latency)
    a2, b2 = torch.rand(64, 3, 512, 512, device=dev), torch.rand(64, 3, 512, 512, device=dev)
    da2 = torch.empty_like(a2)
    run("l1_fwd", a2.shape, a2.numel() * 8, lambda: ops.l1_fwd(a2, b2))
    run("l1_bwd", a2.shape, a2.numel() * 12, lambda: ops.l1_bwd(a2, b2, one, da2))
    del a2, b2, da2
    p = torch.randn(64, 1, 14, 14, device=dev)
    dpp = torch.empty_like(p)
    run("hinge_fwd", p.shape, p.numel() * 4, lambda: ops.hinge_fwd(p, 1))
    run("hinge_bwd", p.shape, p.numel() * 8, lambda: ops.hinge_bwd(p, 1, one, dpp))
    heads = torch.randn(64, 256, device=dev)
    bm, bl, eps = torch.zeros(128, device=dev), torch.zeros(128, device=dev), torch.randn(64, 128, device=dev)
    run("reparam_kl_fwd", heads.shape, heads.numel() * 4 * 2.5, lambda: ops.reparam_kl_fwd(heads, bm, bl, eps))
    # clip + Adam on 64 Mi parameters (the U-Net of configs[4] has three tensors of this order): sumsq reads g; Adam reads
    # p, g, m, v and writes p, m, v = 28 B per parameter
    nparam = 64 << 20
    pw, g, m, v = (torch.randn(nparam, device=dev) * 0.01 for _ in range(4))
    v.abs_()
    nsq = torch.zeros((), device=dev)
    run("sumsq", (nparam,), nparam * 4, lambda: ops.sumsq(g, nsq))
    run("adam_step (with clip)", (nparam,), nparam * 28, lambda: ops.adam_step(pw, g, m, v, 1e-4, 0.5, 0.999, 1e-8, 3, nsq, 1.0))
    print("ok")


if __name__ == "__main__":
    main()
