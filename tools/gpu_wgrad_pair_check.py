#!/usr/bin/env python
"""CTA-pair (cta_group::2) weight-gradient kernel against the single-CTA kernel and torch, plus timings (dev tool).

    python tools/gpu_wgrad_pair_check.py [--time]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_gan_mark_b200 import _lib  # noqa: E402
from vae_gan_mark_b200.conv import ConvLinear, new_act  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


CASES = [  # cin, cout, k, stride, pad, n, h, w
    (512, 512, 3, 1, 1, 4, 32, 32), (512, 512, 3, 1, 1, 64, 128, 128), (256, 512, 3, 1, 1, 8, 16, 16),
    (128, 256, 3, 1, 1, 8, 32, 32), (512, 1024, 1, 1, 0, 8, 32, 32), (64, 256, 1, 1, 0, 4, 16, 16),
    (1024, 384, 3, 1, 1, 3, 8, 8), (128, 256, 3, 2, 1, 5, 16, 16), (512, 256, 2, 2, 0, 4, 16, 16), (192, 320, 3, 1, 1, 2, 12, 20),
]


def main():
    timing = "--time" in sys.argv
    lib = _lib.lib()
    ok = True
    for cin, cout, k, s, pd, n, h, w in CASES:
        if (n, h) == (64, 128) and not timing:
            continue
        op = ConvLinear(cin, cout, k, k, s, (pd, pd))
        g = torch.Generator().manual_seed(cin + cout)
        x = new_act(n, h, w, cin, "cuda")
        x.copy_(torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16))
        oh, ow = op.out_hw(h, w)
        dy = new_act(n, oh, ow, cout, "cuda")
        dy.copy_(torch.randn(n, oh, ow, cout, generator=g).to(torch.bfloat16))
        res, ms = {}, {}
        for pairs in (0, 1):
            lib.vg_set_cta_pairs(pairs)
            res[pairs] = op.backward_weight(dy, x).contiguous().clone()
            if timing:
                for _ in range(3):
                    op.backward_weight(dy, x)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    op.backward_weight(dy, x)
                e1.record()
                torch.cuda.synchronize()
                ms[pairs] = e0.elapsed_time(e1) / 10
        lib.vg_set_cta_pairs(1)
        torch.cuda.synchronize()
        e_pair = rel(res[1], res[0])
        line = {"cin": cin, "cout": cout, "k": k, "stride": s, "n": n, "h": h, "w": w, "pair_vs_single": e_pair}
        if n * h * w <= 65536:
            ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, k, k), dy.float().permute(0, 3, 1, 2),
                                              stride=s, padding=pd)
            line["pair_vs_torch"] = rel(res[1], ref)
            ok = ok and line["pair_vs_torch"] < 2e-3
        if timing:
            fl = 2.0 * n * oh * ow * cout * cin * k * k
            line.update(ms_single=ms[0], ms_pair=ms[1], tflops_single=fl / ms[0] / 1e9, tflops_pair=fl / ms[1] / 1e9)
        ok = ok and e_pair < 1e-5
        print(json.dumps(line), flush=True)
    print("OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
