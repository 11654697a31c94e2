#!/usr/bin/env python
"""Sustained (power-capped steady state) timing of the CTA-pair variants against the single-CTA kernels (dev tool):
each mode runs ~1 s back to back, the last third is timed; modes alternate twice so that order effects show."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_gan_mark_b200 import _lib  # noqa: E402
from vae_gan_mark_b200.conv import ConvLinear, new_act  # noqa: E402


def sustained(fn, seconds=1.0):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    iters = max(30, int(seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
    for _ in range(2 * iters // 3):
        fn()
    e0.record()
    n = iters - 2 * iters // 3
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    lib = _lib.lib()
    for (n, h, w) in ((64, 128, 128), (64, 64, 64)):
        op = ConvLinear(512, 512, 3, 3, 1, (1, 1))
        g = torch.Generator().manual_seed(1)
        x = new_act(n, h, w, 512, "cuda"); x.copy_(torch.randn(n, h, w, 512, generator=g).to(torch.bfloat16))
        dy = new_act(n, h, w, 512, "cuda"); dy.copy_(torch.randn(n, h, w, 512, generator=g).to(torch.bfloat16))
        wt = (torch.randn(512, 512, 3, 3, generator=g) * 0.02).cuda()
        wf = op.prep_fwd(wt)
        out = new_act(n, h, w, 512, "cuda")
        fl = 2.0 * n * h * w * 512 * 512 * 9
        for rnd in range(2):
            for pairs in (0, 1):
                lib.vg_set_fprop_cta_pairs(pairs)
                ms = sustained(lambda: op.forward(x, wf, None, 0, out=out))
                print(json.dumps({"kernel": "fprop", "m": [n, h, w], "pairs": pairs, "round": rnd, "ms": ms, "tflops": fl / ms / 1e9}), flush=True)
        lib.vg_set_fprop_cta_pairs(1)
        for rnd in range(2):
            for pairs in (0, 1):
                lib.vg_set_cta_pairs(pairs)
                ms = sustained(lambda: op.backward_weight(dy, x))
                print(json.dumps({"kernel": "wgrad", "m": [n, h, w], "pairs": pairs, "round": rnd, "ms": ms, "tflops": fl / ms / 1e9}), flush=True)
        lib.vg_set_cta_pairs(1)


if __name__ == "__main__":
    main()
