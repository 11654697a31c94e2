#!/usr/bin/env python
"""CTA-pair (cta_group::2) variants of the weight-gradient and forward kernels against the single-CTA kernels and torch,
plus timings (dev tool).

    python tools/gpu_pair_check.py [--time] [--fprop-only | --wgrad-only]
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


FPROP_CASES = [  # cin, cout, k, stride, pad, n, h, w, bias
    (512, 512, 3, 1, 1, 8, 64, 64, False), (512, 512, 3, 1, 1, 64, 128, 128, False), (512, 512, 3, 1, 1, 64, 64, 64, False),
    (512, 512, 3, 1, 1, 64, 32, 32, False), (256, 512, 3, 1, 1, 37, 16, 16, True),
    (512, 1024, 1, 1, 0, 64, 32, 32, True), (128, 256, 3, 2, 1, 64, 64, 64, True), (512, 384, 3, 1, 1, 33, 24, 20, False),
    (256, 256, 3, 1, 1, 64, 32, 32, False),
]


def fprop_checks(timing):
    import torch.nn.functional as F
    lib = _lib.lib()
    ok = True
    for cin, cout, k, s, pd, n, h, w, bias in FPROP_CASES:
        if (n, h) in ((64, 128), (64, 64)) and cin == 512 and not timing:
            continue
        op = ConvLinear(cin, cout, k, k, s, (pd, pd))
        g = torch.Generator().manual_seed(cin * 7 + cout)
        x = new_act(n, h, w, cin, "cuda")
        x.copy_(torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16))
        wt = (torch.randn(cout, cin, k, k, generator=g) * (cin * k * k) ** -0.5).cuda()
        b = torch.randn(cout, generator=g).cuda() if bias else None
        wf = op.prep_fwd(wt)
        oh, ow = op.out_hw(h, w)
        dy = new_act(n, oh, ow, cout, "cuda")
        dy.copy_(torch.randn(n, oh, ow, cout, generator=g).to(torch.bfloat16))
        wb = op.prep_bwd(wt)
        res, ms = {}, {}
        for pairs in ((1, 0) if "--reverse" in sys.argv else (0, 1)):
            lib.vg_set_fprop_cta_pairs(pairs)
            stats = torch.empty((1, 2, cout), device="cuda")
            y = op.forward(x, wf, b, 1, stats=stats)
            dx = op.backward_data(dy, wb, (h, w))
            res[pairs] = (y.clone(), dx.clone(), stats.clone())
            if timing:
                for _ in range(3):
                    op.forward(x, wf, b, 1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    op.forward(x, wf, b, 1)
                e1.record()
                torch.cuda.synchronize()
                ms[pairs] = e0.elapsed_time(e1) / 10
        lib.vg_set_fprop_cta_pairs(1)
        torch.cuda.synchronize()
        same_y, same_dx = torch.equal(res[0][0], res[1][0]), torch.equal(res[0][1], res[1][1])
        line = {"kernel": "fprop", "cin": cin, "cout": cout, "k": k, "stride": s, "n": n, "h": h, "w": w,
                "y_bit_equal": same_y, "dgrad_bit_equal": same_dx, "stats_rel": rel(res[1][2], res[0][2])}
        if n * h * w <= 300000:
            ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.bfloat16().float(), b, stride=s, padding=pd)).permute(0, 2, 3, 1)
            line["pair_vs_torch"] = rel(res[1][0].float(), ref)
            ok = ok and line["pair_vs_torch"] < 1e-2
        if timing:
            fl = 2.0 * n * oh * ow * cout * cin * k * k
            line.update(ms_single=ms[0], ms_pair=ms[1], tflops_single=fl / ms[0] / 1e9, tflops_pair=fl / ms[1] / 1e9)
        ok = ok and same_y and same_dx and line["stats_rel"] < 1e-5
        print(json.dumps(line), flush=True)
    return ok


def main():
    timing = "--time" in sys.argv
    lib = _lib.lib()
    ok = True
    if "--wgrad-only" not in sys.argv:
        ok = fprop_checks(timing) and ok
    if "--fprop-only" in sys.argv:
        print("OK" if ok else "MISMATCH")
        sys.exit(0 if ok else 1)
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
