#!/usr/bin/env python
"""Data gradient with the forward operand read MN-major vs a transposed K-major copy (dev tool): ms and TFLOP/s."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tools.gpu_narrow_conv import timeit  # noqa: E402
from vae_gan_mark_b200.conv import ConvLinear, new_act  # noqa: E402

SHAPES = [(64, 128, 128, 512, 512, 3, 1), (64, 64, 64, 512, 512, 3, 1), (64, 32, 32, 256, 256, 3, 1), (64, 128, 128, 64, 64, 3, 1),
          (64, 64, 64, 128, 128, 3, 1), (64, 64, 64, 64, 128, 4, 2), (64, 32, 32, 128, 256, 4, 2), (64, 128, 128, 256, 512, 1, 1)]
for (n, h, w, cin, cout, k, s) in SHAPES:
    op = ConvLinear(cin, cout, k, k, s, ((k - 1) // 2, (k - 1) // 2))
    oh, ow = op.out_hw(h, w)
    g = torch.Generator().manual_seed(1)
    dy = new_act(n, oh, ow, cout, "cuda"); dy.copy_(torch.randn(n, oh, ow, cout, generator=g).to(torch.bfloat16))
    wt = (torch.randn(cout, cin, k, k, generator=g) * 0.05).cuda()
    wf, wb = op.prep_fwd(wt), op.prep_bwd(wt)
    dx1, dx2 = new_act(n, h, w, cin, "cuda"), new_act(n, h, w, cin, "cuda")
    fl = 2.0 * n * oh * ow * cin * cout * k * k
    tk = timeit(lambda: op.backward_data(dy, wb, (h, w), out=dx1))
    tm = timeit(lambda: op.backward_data(dy, {"mn": wf}, (h, w), out=dx2))
    err = float((dx1.float() - dx2.float()).abs().max() / dx1.float().abs().max())
    print(f"{str((n, h, w, cin, cout, k, s)):>34s}  K-major {tk:7.3f} ms {fl / tk / 1e9:6.0f} TF/s   MN-major {tm:7.3f} ms {fl / tm / 1e9:6.0f} TF/s   "
          f"max diff {err:.1e}  prefer_mn={op.prefer_mn(n * h * w)}", flush=True)
